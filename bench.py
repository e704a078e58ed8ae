#!/usr/bin/env python
"""bench.py - throughput of the LightPyCL trace step on B200 (and of the reference CPU path).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c3_parabolic_mirror]

Metric (BASELINE.json / reference README.md:232): ray x triangle tests per second
("RI/s" = sum over iterations of live rays x scene triangles, divided by time), plus traced
rays per second.  One STEP = one complete trace of the workload (every generation until the
reference's termination tests fire).  Default workload = C3, BASELINE's "1M rays x ~100k
triangles, depth 10" (parabolic mirror + terminator + measure dome, 100,552 triangles).

  value   whole-job tests/s with the source rays and the scene already resident in HBM
          (CL_Tracer.run after one CL_Tracer.prepare), CUDA-event timed, max over ranks.
  e2e     the same metric through the public call a user makes - CL_Tracer.iterative_tracer(
          light_source, meshes, ...) + get_measured_rays() - with HOST NumPy inputs: host
          flattening, host->device copies of rays and geometry, scene build and the device->host
          read of the measured (pos, pwr) are all inside the timed region.
  roofline            the dominant kernel (nearest-hit search) over its CUDA-event time: the default
                      tensor-core search does 144 tf32 flop per test (3 outputs x K = 24), bound
                      "tensor", peak = MEASURED_PEAKS bf16_tflops / 2; with LPCL_SEARCH_IMPL=simt the
                      packed-FP32 search, 52 FLOP per test against the FP32 lane peak.
  roofline_material   the fused resolve/Fresnel/compaction kernel against measured HBM bandwidth.
  cpu_baseline        the reference kernels (oracle/_ref, compiled from the reference .cl) on
                      the host cores, bounded sample.  N=1, rank 0 only.

  roofline_fp32       SURVEY.md 8(d)'s yardstick for the same kernel: 52 FP32 FLOP per test against the
                      FP32 lane peak (SMs x 128 x 2 x f_SM), with the ncu FMA-pipe utilisation beside it.
  cpu_baseline        (also) records the run-time probe for PyOpenCL / a pocl CPU platform
                       and which baseline ran.

Multi-GPU (one process per GPU, torchrun): the north_star split.  ONE global ray set (the N = 1
workload) is sharded over the ranks by CL_Tracer(shard_input=True), the scene is replicated; the
collectives are the per-iteration all-reduce of the live count / power and - inside the `e2e`
region - the final gather of the measured rays (get_measured_rays(gather=True)): "scaling":
"strong".  The weak-scaling figure (every rank its own full-size ray set, no gather) is kept as the
extra key `weak_scaling` (`--scaling weak` makes it the headline instead).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_TEST = 52.0  # the reference formulation cl:72-98, no early out
METRIC = "ray_triangle_tests_per_s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c3_parabolic_mirror")
    ap.add_argument("--rays", type=int, default=1_000_000, help="rays per source (per GPU)")
    ap.add_argument("--cpu-rays", type=int, default=20_000, help="ray count of the bounded CPU sample")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = one global ray set sharded over the ranks (default); weak = --rays per GPU")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the extra weak-scaling measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def build_workload(name, rays, seed_shift=0):
    from lightpycl_b200 import scenes

    old = scenes.SEED0
    scenes.SEED0 = old + seed_shift  # different rays on every rank (weak scaling)
    try:
        kw = dict(ray_count=rays)
        if name == "c1_nested_cubes":
            kw = dict(ray_count=rays)
        ls, meshes, params = scenes.CONFIGS[name](**kw)
    finally:
        scenes.SEED0 = old
    return ls, meshes, params


def workload_config(name, rays_total, tri_count, params, n_gpus, scaling="strong"):
    return {
        "workload": name,
        "rays": int(rays_total),
        "triangles": int(tri_count),
        "trace_iterations": int(params["trace_iterations"]),
        "trace_until_dissipated": float(params["trace_until_dissipated"]),
        "max_ray_len": float(params["max_ray_len"]),
        "parallelism": ("1 GPU" if n_gpus == 1 else
                        ("one ray set of %d rays sharded over %d GPUs, scene replicated" % (rays_total, n_gpus)
                         if scaling == "strong" else
                         "%d rays per GPU x %d GPUs, scene replicated" % (rays_total // n_gpus, n_gpus))),
        "l2": "flushed (256 MiB write) before every timed step",
    }


def opencl_probe():
    """probe at run time for PyOpenCL and a CPU OpenCL platform (pocl) and
    say which baseline path is used.  Even with both present the reference's PyOpenCL path could not
    run on the GPU box as is: its kernel source lives in /root/reference, which does not travel (and
    must not be copied into this repository), so the baseline is always the same kernel file
    compiled ahead of time by g++ (oracle/_ref)."""
    probe = {"pyopencl": False, "cpu_platforms": [], "used": None}
    try:
        import pyopencl as cl  # noqa: F401

        probe["pyopencl"] = True
        try:
            for plat in cl.get_platforms():
                for dev in plat.get_devices():
                    if dev.type & cl.device_type.CPU:
                        probe["cpu_platforms"].append("%s / %s (%d CUs)" % (plat.name, dev.name, dev.max_compute_units))
        except Exception as exc:  # an ICD loader without platforms raises
            probe["error"] = str(exc)[:200]
    except Exception as exc:
        probe["error"] = "import pyopencl: %s" % (str(exc)[:160],)
    return probe


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.stop_flag, self.source = None, False, "nvidia-smi -lms 100"

    def _start_nvml(self):
        """In-process NVML sampling every ~5 ms: nvidia-smi needs a few hundred ms before its first line, longer than the
        whole timed region of a sharded run (8 GPUs: 5 traces of 6 ms).  Any failure falls back to nvidia-smi."""
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(self.index))
        smax = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)  # fail here, not in the thread
        pynvml.nvmlDeviceGetPowerUsage(handle)
        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
        bits = ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6))  # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap -> row slot

        def loop():
            while not self.stop_flag:
                try:
                    row = [str(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)), str(smax),
                           str(pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0), "", "", "", ""]
                    mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
                    for bit, slot in bits:
                        row[slot] = "Active" if mask & bit else "Not Active"
                    self.rows.append(row)
                except Exception:
                    pass
                time.sleep(0.005)

        self.nvml = threading.Thread(target=loop, daemon=True)
        self.nvml.start()
        self.source = "NVML, every 5 ms"

    def start(self):
        try:
            self._start_nvml()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.nvml.join(timeout=1.0)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                power.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        # under load = samples drawing more than half of the maximum sampled power
        if sm:
            pmax = max(power)
            loaded = [c for c, p in zip(sm, power) if p >= 0.5 * pmax] or sm
            return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                    "samples": len(sm), "power_w_max": pmax, "source": self.source}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_run(name, cpu_rays, steps, warmup):
    """Time the reference's own CPU implementation of the path (the reference .cl compiled by g++
    -O3 + OpenMP through oracle/cl_shim.h: "reference"; or the C restatement: "port") on all host
    cores, each step = one full trace of a bounded ray sample of the same scene."""
    from oracle import oracle

    probe = opencl_probe()
    if oracle.available("reference_fast"):
        backend, kind = "reference_fast", "reference"
        probe["used"] = ("reference .cl compiled by g++ -O3/OpenMP (oracle/_ref/libref_fast.so); PyOpenCL/pocl %s"
                         % ("present but the kernel source does not travel to this box" if probe["pyopencl"]
                            else "not installed"))
    else:
        oracle.build()
        backend, kind = "port_fast", "port"
        probe["used"] = "C restatement of the reference kernels (oracle/trace_oracle.c), g++ -O3/OpenMP"
    cores = oracle.set_threads(backend)  # all host threads (torchrun exports OMP_NUM_THREADS=1)
    ls, meshes, params = build_workload(name, cpu_rays)
    tri = None
    times, tests, rays = [], 0, 0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = oracle.trace(ls, meshes, backend=backend, keep_debug=False, power_sum="f64", **params)
        dt = time.perf_counter() - t0
        tri = out["tri_count"]
        if i >= warmup:
            times.append(dt)
            rays = sum(len(r[0]) for r in out["results"])
            tests = rays * tri
    total = sum(times)
    value = tests * len(times) / total
    n_src = int(sum(int(x.ray_count) for x in ls))
    sample = "%d source rays x %d triangles, full trace (%d ray-iterations/step), %s via g++ -O3/OpenMP" % (
        n_src, tri, rays, "reference .cl" if kind == "reference" else "C restatement")
    return dict(value=value, rays_per_s=rays * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores,
                kind=kind, sample=sample, tri=tri, params=params, ls=ls, cpu_rays=n_src, opencl_probe=probe)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_reference_run(args.workload, args.cpu_rays, max(args.steps, 1), max(args.warmup, 0))
    sources = len(r["ls"])
    # the native arm's workload (what the sample is a sample OF) ...
    cfg = workload_config(args.workload, args.rays * sources, r["tri"], r["params"], args.gpus, args.scaling)
    # ... and what this arm actually traced per step: a bounded sample of its rays on the full scene.  The metric
    # (tests/s of a brute-force loop) is per test, so the sample measures the same quantity.
    cfg["cpu_rays"] = r["cpu_rays"]
    cfg["cpu_sample"] = "bounded sample: %d of the workload's %d source rays, full scene, full depth" % (
        r["cpu_rays"], args.rays * sources)
    cfg["cpu_threads"] = "%d host threads (OpenMP over rays); the parallelism / l2 keys describe the native arm" % r["cores"]
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "tests/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": args.scaling if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "traced_rays_per_s": r["rays_per_s"],
        "cpu_baseline": {"value": r["value"], "unit": "tests/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"], "opencl_probe": r["opencl_probe"]},
        "e2e": {"value": r["value"], "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ native arm
def ncu_summary(prefixes=("r2c", "r2b", "r2", "r1tc")):
    """Numbers from the committed `ncu --set full` capture of this workload's search kernel
    (profiles/<prefix>_ncu_traffic.json, written by tools/summarize_ncu.py): DRAM bytes per launch and
    pipe utilisations.  Read-only context for the roofline objects; never a timing source."""
    out = {"search_tc_kernel": {}, "search_kernel": {}, "step_kernel": {}, "file": None}
    for pre in prefixes:
        path = os.path.join(ROOT, "profiles", pre + "_ncu_traffic.json")
        try:
            cap = json.load(open(path))
        except Exception:
            continue
        for k in ("search_tc_kernel", "search_kernel", "step_kernel"):
            if k in cap.get("kernels", {}) and not out[k]:
                # a capture of a streamed-upload trace holds one generation as several slice launches: their bytes add up
                parts = cap["kernels"][k][: int(cap.get("slices_of_one_generation", 1))]
                v = parts[-1]
                tot = sum(q.get("dram_read_bytes", float("nan")) + q.get("dram_write_bytes", float("nan")) for q in parts)
                out[k] = {"traffic": tot if tot == tot else None, "fma_pipe_pct": v.get("fma_pipe_pct"),
                          "tensor_pipe_pct": v.get("tensor_pipe_pct"), "issue_pct": v.get("issue_pct"),
                          "file": "profiles/" + pre + "_ncu_traffic.json"}
    return out


def main_native(args):
    import torch

    from lightpycl_b200 import iterative_tracer as it

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as td

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    strong = world > 1 and args.scaling == "strong"

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize(device)

    def reduce_ranks(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        td.all_reduce(t, op=getattr(td.ReduceOp, op))
        return float(t.item())

    def gather_lists(values):
        """Every rank's list of per-iteration numbers (same length everywhere) -> (world, iterations) tensor."""
        t = torch.tensor(values, dtype=torch.float64, device=device).reshape(1, -1)
        if world == 1:
            return t.cpu()
        out = [torch.empty_like(t) for _ in range(world)]
        td.all_gather(out, t)
        return torch.cat(out, dim=0).cpu()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def timed_runs(tr, params, steps):
        """`steps` device-resident traces, CUDA-event timed, L2 flushed before each."""
        acc = dict(step_ms=[], search_ms=0.0, resolve_ms=0.0, rays=0, launches=0, resolve_bytes=0.0, n_search=0)
        for _ in range(steps):
            flush.fill_(1)  # evict L2 (126 MB) between timed steps
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tr.run(**params)
            e1.record()
            e1.synchronize()
            acc["step_ms"].append(e0.elapsed_time(e1))
            acc["launches"] += tr.kernel_launches
            acc["rays"] += sum(tr.stats["rays"])
            # rank-LOCAL parents and surviving children of every generation (132 B + 44 B)
            if tr.stats.get("one_launch_generations"):  # small scene traced in one launch: no per-generation kernels
                acc["search_ms"] += sum(kt["search_ms"] for kt in tr.kernel_times)
                acc["n_search"] += len(tr.kernel_times)
                continue
            for kt, live_local in zip(tr.kernel_times, [x for x, n in zip(tr.stats["live_local"], tr.stats["rays"]) if n]):
                acc["search_ms"] += kt["search_ms"]
                acc["resolve_ms"] += kt["resolve_ms"]
                acc["resolve_bytes"] += 132.0 * kt["rays"] + 44.0 * live_local
                acc["n_search"] += kt["tiles"]
        return acc

    # ---- the headline: N = 1 or strong scaling = ONE global ray set (the same rays at every N); weak = own rays per rank
    ls, meshes, params = build_workload(args.workload, args.rays, seed_shift=0 if (strong or world == 1) else 1000 * rank)
    sources = len(ls)
    tr = it.CL_Tracer(platform_name="NVIDIA", device_name="B200", device=device, keep_results=False, profile=True,
                      shard_input=strong)
    tr.prepare(ls, meshes)
    tri = int(tr.tri_count)
    for _ in range(max(args.warmup, 0)):
        tr.run(**params)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    main = timed_runs(tr, params, args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    t_max = reduce_ranks(sum(main["step_ms"]) * 1e-3, "MAX")
    rays_all = reduce_ranks(float(main["rays"]), "SUM")
    value = rays_all * tri / t_max
    per_rank_rays = gather_lists(tr.stats["rays"])  # (world, iterations): load balance of the ray shards
    balance = [round(float(c.max() / max(c.mean(), 1e-9)), 4) for c in per_rank_rays.t()]

    # ---- weak scaling as an extra key (N > 1, strong headline): every rank its own full-size ray set, no gather
    weak = None
    if strong and not args.no_weak:
        lsw, meshes_w, params_w = build_workload(args.workload, args.rays, seed_shift=1000 * rank)
        trw = it.CL_Tracer(platform_name="NVIDIA", device_name="B200", device=device, keep_results=False, profile=True,
                           shard_input=False)
        trw.prepare(lsw, meshes_w)
        trw.run(**params_w)
        barrier()
        w = timed_runs(trw, params_w, min(args.steps, 3))
        barrier()
        tw = reduce_ranks(sum(w["step_ms"]) * 1e-3, "MAX")
        rw = reduce_ranks(float(w["rays"]), "SUM")
        weak = {"value": rw * tri / tw, "unit": "tests/s", "rays_per_gpu": args.rays * sources,
                "ms_per_step": 1e3 * tw / min(args.steps, 3), "steps": min(args.steps, 3),
                "note": "every rank traces its own %d rays (no gather): N independent copies of the N = 1 job"
                        % (args.rays * sources)}
        del trw

    # ---- end to end through the public API with host inputs (+ the gather of the measured rays at N > 1)
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 3))
        times, h2d, d2h, rays_e2e, breakdown, gathered = [], 0, 0, 0, [], 0
        tr2 = it.CL_Tracer(platform_name="NVIDIA", device_name="B200", device=device, keep_results=False,
                           shard_input=strong)
        # warm-up: the same W untimed steps as the device-resident loop (at least two: host result buffers and device
        # tensors of step k are still referenced while step k + 1 allocates, so the pools settle after the second step)
        pos = pwr = None
        for _ in range(max(2, args.warmup)):
            tr2.iterative_tracer(ls, meshes, **params)
            pos, pwr = tr2.get_measured_rays(gather=strong)  # held like in the timed loop: the host pool alternates two buffers
        barrier()
        for _ in range(e2e_steps):
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            tr2.iterative_tracer(ls, meshes, **params)
            t1 = time.perf_counter()
            pos, pwr = tr2.get_measured_rays(gather=strong)
            torch.cuda.synchronize(device)
            times.append(time.perf_counter() - t0)
            breakdown.append({"host_flatten_ms": round(1e3 * tr2.timings.get("host_flatten_s", 0.0), 2),
                              "upload_and_scene_ms": round(1e3 * tr2.timings.get("upload_s", 0.0), 2),
                              "device_loop_ms": round(1e3 * tr2.timings.get("run_s", 0.0), 2),
                              "measured_gather_and_download_ms": round(1e3 * (time.perf_counter() - t1), 2)})
            if os.environ.get("LPCL_GATHER_TIMING") == "1" and strong and world > 1:  # diagnostics (adds synchronisations)
                from lightpycl_b200 import dist as _d

                breakdown[-1]["gather_phases_ms"] = {k: round(1e3 * v, 2) for k, v in _d.GATHER_TIMINGS.items()}
            h2d = tr2.h2d_bytes
            d2h = int(pos.shape[0] * 16 + 64 * len(tr2.stats["rays"]))
            rays_e2e = sum(tr2.stats["rays"])
            gathered = int(pos.shape[0])
        barrier()
        t_e2e = reduce_ranks(sum(times), "MAX")
        e2e_tests = reduce_ranks(float(rays_e2e * e2e_steps), "SUM") * tri
        measured_sum = int(reduce_ranks(float(sum(tr2.stats["measured"])), "SUM"))
        if strong:  # every rank must hold ALL measured rays after the gather
            assert gathered == measured_sum, "gathered %d measured rays, the ranks measured %d" % (gathered, measured_sum)
        e2e = {"value": e2e_tests / t_e2e, "unit": "tests/s",
               "h2d_bytes_per_step": int(reduce_ranks(float(h2d), "SUM")),
               "d2h_bytes_per_step": int(reduce_ranks(float(d2h), "SUM")),
               "ms_per_step": 1e3 * t_e2e / e2e_steps, "steps": e2e_steps, "warmup": max(2, args.warmup),
               "traced_rays_per_s": e2e_tests / tri / t_e2e, "breakdown_per_step_rank0": breakdown,
               "measured_rays": gathered if strong else measured_sum,
               "gather": ("get_measured_rays(gather=True): padded all-gather of every rank's measured records, merged by "
                          "(iteration, lineage); sum of per-rank counts == gathered count asserted") if strong else
                         ("none (N = 1)" if world == 1 else "none (weak scaling: every rank keeps its own measured rays)")}

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_count = torch.cuda.get_device_properties(device).multi_processor_count
    sm_max_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12  # TFLOP/s: SMs x 128 lanes x 2 x f_SM
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ncu = ncu_summary()
    # rank 0's own kernels over rank 0's own work
    search_ms, resolve_ms, n_search = main["search_ms"], main["resolve_ms"], main["n_search"]
    tests_rank0 = main["rays"] * tri
    tests_per_s_kernel = tests_rank0 / (search_ms * 1e-3) if search_ms else 0.0
    search_tflops = FLOP_PER_TEST * tests_per_s_kernel / 1e12
    simt = os.environ.get("LPCL_SEARCH_IMPL", "tc") == "simt"
    kname = "search_kernel" if simt else "search_tc_kernel"
    roofline_fp32 = {
        "kernel": "lpcl::" + kname, "bound": "fp32", "achieved": search_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": search_tflops / fp32_peak if fp32_peak else None, "algorithmic_flop_per_test": FLOP_PER_TEST,
        "ncu_fma_pipe_pct": ncu[kname].get("fma_pipe_pct"), "ncu_file": ncu[kname].get("file"),
        "peak_source": "SMs x 128 FP32 lanes x 2 x sm_max_mhz (%d x 128 x 2 x %.0f MHz); MEASURED_PEAKS.json has no FP32 "
                       "entry" % (sm_count, sm_max_mhz),
        "note": "SURVEY.md 8(d)'s yardstick: the reference formulation's 52 FP32 FLOP per test (cl:72-98, no early out) over "
                "the search kernel's CUDA-event time.  frac > 1 is possible because the kernel does not execute that "
                "formulation: " + ("the packed-FP32 filter executes 19.25 FMA-pipe lane-ops per test" if simt else
                                   "the 15 multiply-adds per test of the filter run on the tensor pipe (roofline), the FP32 "
                                   "lanes only apply the |DEN| - max(|aU|, |aV|) + k >= 0 test to the GEMM output") +
                "; ncu_fma_pipe_pct is the FMA pipe's utilisation in the committed capture (formulation independent)",
    }
    if simt:
        roofline = dict(roofline_fp32)
        roofline["traffic"] = ncu[kname].get("traffic")
        roofline["traffic_note"] = ("dram read+write bytes per launch, ncu --set full; algorithmic DRAM bytes per launch = "
                                    "32 B/ray + 64 B/triangle")
    else:
        # the default search evaluates the filter as a split-TF32 GEMM: 3 outputs x K = 24 multiply-adds per test
        tc_flop = 3 * 24 * 2.0
        tf32_peak = float(peaks.get("bf16_tflops", 1681.8)) / 2.0
        tc_tflops = tc_flop * tests_per_s_kernel / 1e12
        roofline = {
            "kernel": "lpcl::search_tc_kernel", "bound": "tensor", "achieved": tc_tflops, "peak": tf32_peak,
            "unit": "TFLOP/s", "frac": tc_tflops / tf32_peak if tf32_peak else None,
            "traffic": ncu[kname].get("traffic"), "traffic_note": "dram read+write bytes per launch, ncu --set full (%s); "
            "algorithmic DRAM bytes per launch = 32 B/ray + 288 B/triangle + results" % ncu[kname].get("file"),
            "peak_source": "kind::tf32 tcgen05.mma runs at half the bf16 rate: MEASURED_PEAKS.json bf16_tflops / 2, the BURST "
                           "figure (the stricter denominator; peak_sustained / frac_sustained use bf16_tflops_sustained / 2, the "
                           "figure for a kernel inside a long step; tools/ubench_tmem.cu measures 4092 tf32 flop/clk/SM = 1190 "
                           "TFLOP/s at 1965 MHz unthrottled)",
            "peak_sustained": float(peaks.get("bf16_tflops_sustained", 0.0)) / 2.0 or None,
            "frac_sustained": (tc_tflops / (float(peaks["bf16_tflops_sustained"]) / 2.0)) if peaks.get("bf16_tflops_sustained") else None,
            "algorithmic_flop_per_test": tc_flop, "useful_flop_per_test": 36.0,
            "useful_frac": 36.0 * tests_per_s_kernel / 1e12 / tf32_peak if tf32_peak else None,
            "ncu_tensor_pipe_pct": ncu[kname].get("tensor_pipe_pct"),
            "note": "144 = the tensor-core flops this formulation ISSUES per test (3 outputs x K 24 x 2: the hi/lo TF32 split "
                    "of both operands quadruples the K = 6 contraction); the contraction itself is 36 flop per test "
                    "(useful_flop_per_test, useful_frac).  The reference's 52 fp32 FLOP/test yardstick is roofline_fp32. "
                    "avg_launch_ms is one lpcl_trace_search call: scratch initialisation + the deferred-flag sweep "
                    "(search_tc_kernel) + its second pass (resolve_records_kernel, ~1 % on C3).  What keeps the sweep off "
                    "the tensor peak: TMEM read-back, chunk tests and MMAs share one SM (DESIGN.md 4.1), and the SM clock "
                    "under this kernel is 1.82 - 1.84 of 1.965 GHz (sw_power_cap; `clocks`)",
        }
    roofline.update({"launches": n_search, "avg_launch_ms": search_ms / max(n_search, 1),
                     "tests_per_s_kernel": tests_per_s_kernel,
                     "share_of_step": search_ms / sum(main["step_ms"]) if main["step_ms"] else None,
                     "scope": "rank 0's launches over rank 0's rays"})
    resolve_gbs = main["resolve_bytes"] / (resolve_ms * 1e-3) / 1e9 if resolve_ms else 0.0

    line = {
        "metric": METRIC, "value": value, "unit": "tests/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.rays * sources * (1 if (strong or world == 1) else world), tri, params,
                                  world, "strong" if strong else "weak"),
        "traced_rays_per_s": rays_all / t_max,
        "ray_iterations_per_step": int(rays_all) // args.steps,
        "rays_per_iteration_rank0": tr.stats["rays"],
        "rays_per_iteration_max_over_mean": balance,
        "filter_candidates_per_ray": [c / max(r, 1) for c, r in zip(tr.stats["candidates"], tr.stats["rays"])],
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(main["launches"]),
        "roofline": roofline,
        "roofline_fp32": roofline_fp32,
        "roofline_material": {
            "kernel": "lpcl::step_kernel", "bound": "hbm", "achieved": resolve_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": resolve_gbs / hbm_peak if hbm_peak else None, "traffic": ncu["step_kernel"].get("traffic"),
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
            "algorithmic_bytes": "132 B per parent + 44 B per surviving child (SURVEY.md 8(d)), rank-local counts",
            "avg_launch_ms": resolve_ms / max(n_search, 1),
            "share_of_step": resolve_ms / sum(main["step_ms"]) if main["step_ms"] else None,
        },
    }
    if weak is not None:
        line["weak_scaling"] = weak
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.workload, args.cpu_rays, 1, 1)
        line["cpu_baseline"] = {"value": r["value"], "unit": "tests/s", "cores": r["cores"], "kind": r["kind"],
                                "sample": r["sample"], "traced_rays_per_s": r["rays_per_s"],
                                "opencl_probe": r["opencl_probe"]}
    print(json.dumps(line))
    if world > 1:
        td.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    # stdout carries exactly ONE line (the JSON): anything libraries print meanwhile (e.g. NCCL's
    # version banner when the image exports NCCL_DEBUG) is diverted to stderr at the fd level
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _orig_print = print

    def print(*args, **kw):  # noqa: A001 - the final JSON line goes to the real stdout
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        _orig_print(*args, **kw)
        sys.stdout.flush()
        os.dup2(2, 1)

    rc = main_reference(a) if a.impl == "reference" else main_native(a)
    sys.stdout.flush()
    sys.exit(rc)
