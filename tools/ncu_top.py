"""Print the scheduler summary and the top stalled SASS instructions of one kernel of an .ncu-rep
(needs `ncu` on PATH).  Usage: python tools/ncu_top.py REPORT.ncu-rep [N] [KERNEL_REGEX]
With several kernels in the report, the first launch whose name matches KERNEL_REGEX is used."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sel = ["-k", "regex:" + sys.argv[3], "-c", "1"] if len(sys.argv) > 3 else ["-c", "1"]
det = subprocess.run(["ncu", "-i", rep] + sel + ["--page", "details"], capture_output=True, text=True).stdout
for line in det.splitlines():
    if any(k in line for k in ("Duration", "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Registers Per Thread",
                               "Active Warps Per Scheduler", "Eligible Warps", "highest-utilized")):
        print(line.strip()[:150])
src = subprocess.run(["ncu", "-i", rep] + sel + ["--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the source page lists every kernel of the report: blocks of ("Kernel Name", name) / header / rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1] if len(r) > 1 else "", "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
import re
want = sys.argv[3] if len(sys.argv) > 3 else "."
blk = [b for b in blocks if re.search(want, b["name"])][0]
print("kernel:", blk["name"][:120])
hdr, data = blk["hdr"], blk["rows"]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
num = lambda r, k: int(r[ix[k]] or 0)
tot = sum(num(r, "# Samples") for r in data)
print("total samples", tot)
agg = {s: sum(num(r, s) for r in data) for s in stalls}
print(", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
pipes = {}
for r in data:
    op = r[ix["Source"]].split()
    op = [t for t in op if not t.startswith("@")][0].split(".")[0] if op else ""
    pipes[op] = pipes.get(op, 0) + num(r, "Instructions Executed")
ti = sum(pipes.values())
print("instr mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / ti) for k, v in sorted(pipes.items(), key=lambda x: -x[1])[:12]), "| total warp-instr", ti)
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top_n]:
    st = sorted(((s, num(r, s)) for s in stalls), key=lambda x: -x[1])[:2]
    print(r[ix["Address"]][-5:], "%7d %5.1f%% %10d  %-70s %s" % (num(r, "# Samples"), 100.0 * num(r, "# Samples") / tot,
          num(r, "Instructions Executed"), r[ix["Source"]][:70], " ".join("%s:%d" % (s[6:], v) for s, v in st)))
