#!/usr/bin/env python
"""Text summaries of the ncu artefacts a tools/gpu_final.sh run leaves in gpurun_out/ (written to profiles/).
usage: tools/ncu_summary.py <tag>
  <tag>_launches.csv   -> profiles/<tag>_kernel_shares.txt      (time per kernel over the profiled command)
  <tag>_trace.ncu-rep  -> profiles/<tag>_k_trace_wide_details.txt (every `details` metric of the captured traversal launches)
                          profiles/<tag>_k_trace_wide_lines.txt   (hottest source lines, tools/ncu_lines.py)
  <tag>_others.ncu-rep -> profiles/<tag>_other_kernels.txt        (one line per captured launch of the non-traversal kernels)
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
PEAK = 6538.6


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


# ---- kernel shares
rows = [r for r in csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))) if len(r) > 10]
h = rows[0]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[1:]:
    if r[h.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    t, n = acc.get(r[ki], (0.0, 0)); acc[r[ki]] = (t + ms, n + 1)
tot = sum(t for t, _ in acc.values())
with open(os.path.join(P, f"{tag}_kernel_shares.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum over `bench.py --steps 1 --warmup 1 --spp 2 --no-cpu-baseline` (C2, path+NEE, trace_mode 3; k_build_* = GPU octree build, scene setup outside the timed steps)\n")
    for k, (t, n) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{t:10.3f} ms {n:4d} {100 * t / tot:5.1f}%  {k[:90]}\n")

# ---- traversal details
out = ncu("-i", os.path.join(G, f"{tag}_trace.ncu-rep"), "--page", "details", "--csv")
rows = list(csv.reader(out.splitlines()))
h = rows[0]
idx = {k: h.index(k) for k in ("ID", "Kernel Name", "Section Name", "Metric Name", "Metric Unit", "Metric Value")}
with open(os.path.join(P, f"{tag}_k_trace_wide_details.txt"), "w") as f:
    for r in rows[1:]:
        if len(r) <= idx["Metric Value"] or not r[idx["Metric Name"]]:
            continue
        f.write(" | ".join([r[idx["ID"]], r[idx["Kernel Name"]][:40], r[idx["Section Name"]], r[idx["Metric Name"]], r[idx["Metric Unit"]], r[idx["Metric Value"]]]) + "\n")
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), os.path.join(G, f"{tag}_trace.ncu-rep"), "60", "0"], capture_output=True, text=True).stdout
open(os.path.join(P, f"{tag}_k_trace_wide_lines.txt"), "w").write(lines)

# ---- other kernels
out = ncu("-i", os.path.join(G, f"{tag}_others.ncu-rep"), "--page", "raw", "--csv")
rows = list(csv.reader(out.splitlines()))
h = rows[0]


def col(name):
    return h.index(name) if name in h else -1


c = {k: col(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size")}
units = rows[1]


def to(v, unit, kind):
    v = float(v.replace(",", ""))
    if kind == "time":
        return v / 1e3 if unit.startswith("ns") else v if unit.startswith("us") else v * 1e3 if unit.startswith("ms") else v * 1e6
    return v / 1e6 if unit.startswith("byte") else v / 1e3 if unit.startswith("Kbyte") else v if unit.startswith("Mbyte") else v * 1e3


with open(os.path.join(P, f"{tag}_other_kernels.txt"), "w") as f:
    f.write(f"# ncu --set full, bench.py --steps 1 --warmup 1 --spp 2 (C2, path+NEE): first launches of every non-traversal kernel.  GB/s = (dram read+write)/duration; peak {PEAK} GB/s measured\n")
    f.write("kernel | us | DRAM MB r/w | DRAM GB/s | frac of HBM peak | issue active % | warps active % | regs | grid\n")
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        us = to(r[c["gpu__time_duration.sum"]], units[c["gpu__time_duration.sum"]], "time")
        rd = to(r[c["dram__bytes_read.sum"]], units[c["dram__bytes_read.sum"]], "bytes"); wr = to(r[c["dram__bytes_write.sum"]], units[c["dram__bytes_write.sum"]], "bytes")
        gbs = (rd + wr) / 1e3 / (us / 1e6) if us > 0 else 0
        f.write(f"{r[c['Kernel Name']].split('(')[0][-28:]} | {us:.1f} | {rd:.1f}/{wr:.1f} | {gbs:.0f} | {gbs / PEAK:.2f} | {float(r[c['smsp__issue_active.avg.pct_of_peak_sustained_active']]):.0f} | "
                f"{float(r[c['sm__warps_active.avg.pct_of_peak_sustained_active']]):.0f} | {r[c['launch__registers_per_thread']]} | {r[c['launch__grid_size']]}\n")
print("written:", [n for n in sorted(os.listdir(P)) if n.startswith(tag)])
