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

Multi-GPU: one process per GPU (torchrun); rays are sharded (each rank traces its own 1M rays
=> weak scaling), the only collectives are the per-iteration all-reduce of the live count/power.
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


def workload_config(name, rays, ls, tri_count, params, n_gpus):
    return {
        "workload": name,
        "rays_per_gpu": int(sum(int(s.ray_count) for s in ls)),
        "triangles": int(tri_count),
        "trace_iterations": int(params["trace_iterations"]),
        "trace_until_dissipated": float(params["trace_until_dissipated"]),
        "max_ray_len": float(params["max_ray_len"]),
        "parallelism": "ray-sharded x%d, scene replicated" % n_gpus,
        "l2": "flushed (256 MiB write) before every timed step",
    }


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
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
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
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
                    "samples": len(sm), "power_w_max": pmax}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_run(name, cpu_rays, steps, warmup):
    """Time the reference's own CPU implementation of the path (the reference .cl compiled by g++
    -O3 + OpenMP through oracle/cl_shim.h: "reference"; or the C restatement: "port") on all host
    cores, each step = one full trace of a bounded ray sample of the same scene."""
    from oracle import oracle

    if oracle.available("reference_fast"):
        backend, kind = "reference_fast", "reference"
    else:
        oracle.build()
        backend, kind = "port_fast", "port"
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
    sample = "%d source rays x %d triangles, full trace (%d ray-iterations/step), %s via g++ -O3/OpenMP" % (
        cpu_rays, tri, rays, "reference .cl" if kind == "reference" else "C restatement")
    return dict(value=value, rays_per_s=rays * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores,
                kind=kind, sample=sample, tri=tri, params=params, ls=ls)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_reference_run(args.workload, args.cpu_rays, max(args.steps, 1), max(args.warmup, 0))
    cfg = workload_config(args.workload, args.rays, build_workload(args.workload, 4)[0], r["tri"], r["params"], args.gpus)
    cfg["rays_per_gpu"] = args.rays
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "tests/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "traced_rays_per_s": r["rays_per_s"],
        "cpu_baseline": {"value": r["value"], "unit": "tests/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ native arm
def main_native(args):
    import torch

    from lightpycl_b200 import iterative_tracer as it

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as td

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.pop("NCCL_DEBUG", None)  # no NCCL banner (stdout is diverted to stderr anyway, see __main__)
        torch.cuda.set_device(local_rank)
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        td.all_reduce(t, op=td.ReduceOp.SUM)
        return float(t.item())

    ls, meshes, params = build_workload(args.workload, args.rays, seed_shift=1000 * rank)
    tr = it.CL_Tracer(platform_name="NVIDIA", device_name="B200", device=device, keep_results=False, profile=True,
                      shard_input=False)
    tr.prepare(ls, meshes)
    tri = int(tr.tri_count)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    for _ in range(max(args.warmup, 0)):
        tr.run(**params)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_ms, search_ms, resolve_ms, rays_total, launches = [], 0.0, 0.0, 0, 0
    resolve_bytes = 0.0
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)  # evict L2 (126 MB) between timed steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr.run(**params)
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        launches += tr.kernel_launches
        rays_step = sum(tr.stats["rays"])
        rays_total += rays_step
        for kt, live in zip(tr.kernel_times, tr.stats["live"]):
            search_ms += kt["search_ms"]
            resolve_ms += kt["resolve_ms"]
            resolve_bytes += 132.0 * kt["rays"] + 44.0 * live  # per parent + per surviving child
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    t_local = sum(step_ms) * 1e-3
    t_max = max_over_ranks(t_local)
    rays_all = sum_over_ranks(float(rays_total))
    tests_all = rays_all * tri
    value = tests_all / t_max
    n_search = len(tr.kernel_times) * args.steps

    # ---- end to end through the public API with host inputs
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 3))
        times, h2d, d2h, rays_e2e, breakdown = [], 0, 0, 0, []
        tr2 = it.CL_Tracer(platform_name="NVIDIA", device_name="B200", device=device, keep_results=False,
                           shard_input=False)
        tr2.iterative_tracer(ls, meshes, **params)  # warm-up (allocator, scene)
        tr2.get_measured_rays(gather=False)
        barrier()
        for _ in range(e2e_steps):
            torch.cuda.synchronize(device)
            t0 = time.perf_counter()
            tr2.iterative_tracer(ls, meshes, **params)
            t1 = time.perf_counter()
            pos, pwr = tr2.get_measured_rays(gather=False)
            torch.cuda.synchronize(device)
            times.append(time.perf_counter() - t0)
            breakdown.append({"host_flatten_ms": round(1e3 * tr2.timings.get("host_flatten_s", 0.0), 2),
                              "upload_and_scene_ms": round(1e3 * tr2.timings.get("upload_s", 0.0), 2),
                              "device_loop_ms": round(1e3 * tr2.timings.get("run_s", 0.0), 2),
                              "measured_download_ms": round(1e3 * (time.perf_counter() - t1), 2)})
            h2d = tr2.h2d_bytes
            d2h = int(pos.shape[0] * 16 + 64 * len(tr2.stats["rays"]))
            rays_e2e = sum(tr2.stats["rays"])
        barrier()
        t_e2e = max_over_ranks(sum(times))
        e2e_tests = sum_over_ranks(float(rays_e2e * e2e_steps)) * tri
        e2e = {"value": e2e_tests / t_e2e, "unit": "tests/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t_e2e / e2e_steps, "steps": e2e_steps,
               "traced_rays_per_s": e2e_tests / tri / t_e2e, "breakdown_per_step": breakdown}

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
    # dram bytes per launch from the committed `ncu --set full` capture of this same workload
    traffic = {"search_kernel": None, "search_tc_kernel": None, "step_kernel": None}
    for cap_file in ("r1_ncu_traffic.json", "r1tc_ncu_traffic.json"):  # packed-FP32 search capture, tensor-core search capture
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", cap_file)))
            for k in traffic:
                if k in cap["kernels"] and (traffic[k] is None or cap_file.startswith("r1tc")):
                    v = cap["kernels"][k][0]
                    tot = v.get("dram_read_bytes", float("nan")) + v.get("dram_write_bytes", float("nan"))
                    traffic[k] = tot if tot == tot else traffic[k]  # NaN -> keep
        except Exception:
            pass
    search_tflops = FLOP_PER_TEST * (rays_total * tri) / (search_ms * 1e-3) / 1e12 if search_ms else 0.0
    tests_per_s_kernel = (rays_total * tri) / (search_ms * 1e-3) if search_ms else 0.0
    simt = os.environ.get("LPCL_SEARCH_IMPL", "tc") == "simt"
    if simt:
        roofline = {
            "kernel": "lpcl::search_kernel", "bound": "fp32", "achieved": search_tflops, "peak": fp32_peak,
            "unit": "TFLOP/s", "frac": search_tflops / fp32_peak if fp32_peak else None,
            "traffic": traffic["search_kernel"], "traffic_note": "dram read+write bytes per launch, ncu --set full "
            "(profiles/r1_ncu_traffic.json); algorithmic DRAM bytes per launch = 32 B/ray + 64 B/triangle",
            "peak_source": "SMs x 128 FP32 lanes x 2 x sm_max_mhz (%d x 128 x 2 x %.0f MHz); the packed-FP32 search is "
                           "FMA-pipe bound; MEASURED_PEAKS.json has no FP32 entry" % (sm_count, sm_max_mhz),
            "algorithmic_flop_per_test": FLOP_PER_TEST,
            "note": "LPCL_SEARCH_IMPL=simt: frac uses the reference formulation's 52 FLOP/test (SURVEY.md 8(d)); the kernel "
                    "executes 19.25 FMA-pipe lane-ops (packed FFMA2) + 1 ALU op per test, so frac can exceed 1; ncu "
                    "sm__pipe_fma_cycles_active = 77% (profiles/r1_search_kernel_0_ncu.txt)",
        }
    else:
        # the default search evaluates the filter as a split-TF32 GEMM: 3 outputs x K = 24 multiply-adds per test
        tc_flop = 3 * 24 * 2.0
        tf32_peak = float(peaks.get("bf16_tflops", 1681.8)) / 2.0
        tc_tflops = tc_flop * tests_per_s_kernel / 1e12
        roofline = {
            "kernel": "lpcl::search_tc_kernel", "bound": "tensor", "achieved": tc_tflops, "peak": tf32_peak,
            "unit": "TFLOP/s", "frac": tc_tflops / tf32_peak if tf32_peak else None,
            "traffic": traffic["search_tc_kernel"], "traffic_note": "dram read+write bytes per launch, ncu --set full "
            "(profiles/r1tc_ncu_traffic.json); algorithmic DRAM bytes per launch = 32 B/ray + 288 B/triangle + results",
            "peak_source": "kind::tf32 tcgen05.mma runs at half the bf16 rate: MEASURED_PEAKS.json bf16_tflops / 2 "
                           "(tools/ubench_tmem.cu measures 4092 tf32 flop/clk/SM = 1190 TFLOP/s at 1965 MHz unthrottled)",
            "algorithmic_flop_per_test": tc_flop,
            "note": "flops are the tensor-core work of this formulation (DESIGN.md: DEN/uN/vN as a K=24 split-TF32 GEMM), not "
                    "the reference's 52 fp32 FLOP/test; with the latter the kernel runs at %.1f TFLOP/s = %.2f of the "
                    "%.1f TFLOP/s fp32 lane peak.  The tensor pipe is ~30%% busy: the kernel is bound by the CUDA-core "
                    "epilogue that reads the accumulator back (tcgen05.ld) and applies the product tests "
                    "(profiles/r1_tc_notes.txt)" % (search_tflops, search_tflops / fp32_peak, fp32_peak),
        }
    roofline.update({"launches": n_search, "avg_launch_ms": search_ms / max(n_search, 1),
                     "tests_per_s_kernel": tests_per_s_kernel,
                     "share_of_step": search_ms / sum(step_ms) if step_ms else None})
    resolve_gbs = resolve_bytes / (resolve_ms * 1e-3) / 1e9 if resolve_ms else 0.0

    line = {
        "metric": METRIC, "value": value, "unit": "tests/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.rays, ls, tri, params, world),
        "traced_rays_per_s": rays_all / t_max,
        "ray_iterations_per_step": rays_total // args.steps,
        "rays_per_iteration": tr.stats["rays"],
        "filter_candidates_per_ray": [c / max(r, 1) for c, r in zip(tr.stats["candidates"], tr.stats["rays"])],
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_material": {
            "kernel": "lpcl::step_kernel", "bound": "hbm", "achieved": resolve_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": resolve_gbs / hbm_peak if hbm_peak else None, "traffic": traffic["step_kernel"],
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
            "algorithmic_bytes": "132 B per parent + 44 B per surviving child (SURVEY.md 8(d))",
            "avg_launch_ms": resolve_ms / max(n_search, 1), "share_of_step": resolve_ms / sum(step_ms) if step_ms else None,
        },
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.workload, args.cpu_rays, 1, 1)
        line["cpu_baseline"] = {"value": r["value"], "unit": "tests/s", "cores": r["cores"], "kind": r["kind"],
                                "sample": r["sample"], "traced_rays_per_s": r["rays_per_s"]}
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
