#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum[,smsp__thread_inst_executed_per_inst_executed.ratio] --csv`).
usage: tools/launch_shares.py launches.csv [--seq N]"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[h]; ix = {k: i for i, k in enumerate(hdr)}
t = collections.defaultdict(float); n = collections.Counter(); thr = collections.defaultdict(list); seq = []
for r in rows[h + 1:]:
    if len(r) < len(hdr): continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("crt::", "")[:48]
    v = float(r[ix["Metric Value"]].replace(",", ""))
    if r[ix["Metric Name"]] == "gpu__time_duration.sum":
        u = r[ix["Metric Unit"]]; v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        t[name] += v; n[name] += 1; seq.append((name, v))
    else:
        thr[name].append(v)
ours = {k: v for k, v in t.items() if not k.startswith(("at::", "cub::", "k_build"))}
tot = sum(ours.values())
print(f"total {tot / 1e3:.2f} ms over {sum(n[k] for k in ours)} launches (render kernels only)")
for k, v in sorted(ours.items(), key=lambda x: -x[1]):
    a = thr.get(k)
    print(f"{k:48s} {n[k]:5d} {v / 1e3:9.3f} ms {100 * v / tot:5.1f}%" + (f"  threads/inst {sum(a) / len(a):.1f}" if a else ""))
if "--seq" in sys.argv:
    m = int(sys.argv[sys.argv.index("--seq") + 1])
    print([(a[:18], round(b)) for a, b in seq[:m]])
