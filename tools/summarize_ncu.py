#!/usr/bin/env python
"""Turn raw ncu output (gpurun_out/) into the small text/JSON summaries committed under profiles/.

    python tools/summarize_ncu.py --tag r1 --launches gpurun_out/launches_r1.csv --report gpurun_out/prof_r1.ncu-rep

* <tag>_launches.csv / <tag>_launches_summary.txt : the `--metrics gpu__time_duration.sum` launch list
  (cold-cache, serialised: compare SHARES) with per-kernel totals.
* <tag>_<kernel>_ncu.txt : key metrics of each kernel in a `--set full` report.
* <tag>_ncu_traffic.json  : dram bytes per launch (read + write) per kernel - bench.py reports it as
  `roofline.traffic`.
"""

import argparse
import csv
import json
import os
import re
import shutil
import subprocess

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_tmem_ldt.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
]
PREFIXES = ["smsp__average_warps_issue_stalled_"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(path, out_dir, tag, note):
    shutil.copy(path, os.path.join(out_dir, "%s_launches.csv" % tag))
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    idx = {h: i for i, h in enumerate(rows[0])}
    tot, by, cnt, lines = 0.0, {}, {}, []
    for r in rows[1:]:
        val, unit = float(r[idx["Metric Value"]].replace(",", "")), r[idx["Metric Unit"]]
        ms = val / 1e6 if unit.startswith("ns") else (val / 1e3 if unit.startswith("us") else val)
        short = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        lines.append("%3s %-70s grid %-16s block %-12s %10.4f ms" % (r[idx["ID"]], short[:70], r[idx["Grid Size"]],
                                                                    r[idx["Block Size"]], ms))
        by[short] = by.get(short, 0) + ms
        cnt[short] = cnt.get(short, 0) + 1
        tot += ms
    with open(os.path.join(out_dir, "%s_launches_summary.txt" % tag), "w") as f:
        f.write("# %s\n# cold-cache, serialised per-launch times: compare SHARES, not absolutes\n" % note)
        f.write("\n".join(lines) + "\n")
        f.write("total %.3f ms\n" % tot)
        for k, v in sorted(by.items(), key=lambda x: -x[1]):
            f.write("%-70s x%-3d %10.4f ms %6.2f%%\n" % (k[:70], cnt[k], v, 100 * v / tot))


def report(path, out_dir, tag, note):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic, seen = {}, {}
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        name = re.sub(r"\(.*", "", rec["Kernel Name"]).replace("void ", "").replace("lpcl::", "")
        short = re.sub(r"[^a-z_]", "", name.split("<")[0])
        k = seen.get(short, 0)
        seen[short] = k + 1
        u = dict(zip(hdr, units))
        rd = float(rec["dram__bytes_read.sum"].replace(",", "")) * UNIT.get(u["dram__bytes_read.sum"], 1)
        wr = float(rec["dram__bytes_write.sum"].replace(",", "")) * UNIT.get(u["dram__bytes_write.sum"], 1)
        dur = float(rec["gpu__time_duration.sum"].replace(",", ""))
        def pct(key):
            try:
                return float(rec[key].replace(",", ""))
            except Exception:
                return None

        traffic.setdefault(short, []).append({"dram_read_bytes": rd, "dram_write_bytes": wr,
                                              "duration": dur, "duration_unit": u["gpu__time_duration.sum"],
                                              "grid": rec["launch__grid_size"],
                                              "fma_pipe_pct": pct("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                                              "tensor_pipe_pct": pct("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                                              "issue_pct": pct("sm__inst_issued.avg.pct_of_peak_sustained_active")})
        with open(os.path.join(out_dir, "%s_%s_%d_ncu.txt" % (tag, short, k)), "w") as f:
            f.write("# %s\n# kernel: %s (launch %d of this kernel in the capture)\n" % (note, rec["Kernel Name"][:200], k))
            for h in hdr:
                if h in KEYS or any(h.startswith(p) for p in PREFIXES):
                    f.write("%-92s %-16s %s\n" % (h, u[h], rec[h]))
    return traffic


def metrics_csv(path, traffic):
    """A `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log:
    fills in kernels whose dram counters came back NaN in the --set full capture (seen for the 65 us
    step kernel)."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    idx = {h: i for i, h in enumerate(rows[0])}
    per = {}
    for r in rows[1:]:
        short = re.sub(r"[^a-z_]", "", re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("lpcl::", "").split("<")[0])
        rec = per.setdefault((short, r[idx["ID"]]), {})
        val = float(r[idx["Metric Value"]].replace(",", "")) * UNIT.get(r[idx["Metric Unit"]], 1)
        rec[r[idx["Metric Name"]]] = (val, r[idx["Metric Unit"]])
    for (short, _id), rec in sorted(per.items(), key=lambda kv: int(kv[0][1])):
        entry = {"dram_read_bytes": rec["dram__bytes_read.sum"][0], "dram_write_bytes": rec["dram__bytes_write.sum"][0],
                 "duration": rec["gpu__time_duration.sum"][0], "duration_unit": rec["gpu__time_duration.sum"][1],
                 "source": "metrics-only capture"}
        cur = traffic.get(short, [])
        if not cur or cur[0]["dram_read_bytes"] != cur[0]["dram_read_bytes"]:  # absent or NaN
            traffic[short] = [entry]
    return traffic


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--launches")
    ap.add_argument("--report")
    ap.add_argument("--metrics-csv", help="metrics-only ncu csv used where the full capture returned NaN dram counters")
    ap.add_argument("--note", default="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e (C3: 1M rays x 100,552 triangles)")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    if a.launches:
        launches(a.launches, a.out, a.tag, "ncu --metrics gpu__time_duration.sum --clock-control none; " + a.note)
    if a.report:
        traffic = report(a.report, a.out, a.tag, "ncu --set full --clock-control none; " + a.note)
        if a.metrics_csv:
            traffic = metrics_csv(a.metrics_csv, traffic)
        with open(os.path.join(a.out, "%s_ncu_traffic.json" % a.tag), "w") as f:
            json.dump({"note": "ncu --set full --clock-control none; " + a.note, "kernels": traffic}, f, indent=1)
