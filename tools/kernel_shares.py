#!/usr/bin/env python
"""profiles/r02_kernel_shares.txt: every kernel of a step with its share of the step's kernel time (ncu launch list, cold-cache and
serialised: shares, not absolute times) and the resource that binds it (profiles/r02_traffic.json, from the `ncu --set full` captures).
usage: tools/kernel_shares.py <label>=<launches.csv>[=<traffic key prefix>] ... > profiles/r02_kernel_shares.txt"""
import collections, csv, json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FACTS = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def shares(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[h]; ix = {k: i for i, k in enumerate(hdr)}
    t = collections.defaultdict(float); n = collections.Counter()
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("crt::", "")
        u = r[ix["Metric Unit"]]; v = float(r[ix["Metric Value"]].replace(",", ""))
        v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        t[name] += v; n[name] += 1
    return {k: (v, n[k]) for k, v in t.items() if not k.startswith(("at::", "cub::", "k_build"))}


def binding(f):
    if f is None:
        return "(no --set full capture)"
    issue, thr, dram, l1 = f["issue_slots_busy_pct"], f["active_threads_per_warp"], f["dram_throughput_pct"], f["l1tex_throughput_pct"]
    gbs = f["dram_bytes_per_launch"] / (f["mean_launch_ms"] * 1e-3) / 1e9
    if dram >= 50:
        return f"HBM: {gbs:.0f} GB/s = {gbs / HBM_PEAK:.2f} of the measured {HBM_PEAK:.0f} GB/s (DRAM throughput {dram:.0f} %)"
    if issue >= 60:
        return f"instruction issue: slots {issue:.0f} % busy x {thr:.1f}/32 lanes = {issue / 100 * thr / 32:.2f} of peak thread-instruction rate (L1 {l1:.0f} %, DRAM {dram:.0f} %)"
    return f"latency: issue slots {issue:.0f} % busy, {thr:.1f}/32 lanes, DRAM {dram:.0f} % ({gbs:.0f} GB/s), L1 {l1:.0f} %, occupancy {f['achieved_occupancy_pct']:.0f} %"


for spec in sys.argv[1:]:
    label, path, prefix = (spec.split("=") + [""])[:3]
    s = shares(path)
    tot = sum(v for v, _ in s.values())
    print(f"== {label}   ({os.path.basename(path)}; {sum(c for _, c in s.values())} launches, {tot / 1e3:.2f} ms of kernel time under ncu)")
    groups = collections.OrderedDict()
    for k, (v, c) in sorted(s.items(), key=lambda x: -x[1][0]):
        base = k.split("<")[0]
        g = groups.setdefault(base, [0.0, 0]); g[0] += v; g[1] += c
    for base, (v, c) in groups.items():
        f = FACTS.get(prefix + base) or FACTS.get(base)
        src = "" if f is None else "   [" + f["from"].split(" ")[0] + "]"
        print(f"  {base:22s} {100 * v / tot:5.1f} %  {c:4d} launches   {binding(f)}{src}")
    print()
