#!/usr/bin/env python
"""Per-source-line totals of one profiled launch: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [--launch-skip k --launch-count 1] > f.csv`,
then `tools/ncu_source_lines.py f.csv [top]`.  Prints samples, share, no-instruction share, warp instructions and threads per instruction per line."""
import csv, sys, os
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] != "" and r[0].isdigit() and len(r) >= 12:
        try:
            s = int(r[6]); i = int(r[7]); t = int(r[8]) / max(i, 1); ni = int(r[ix["stall_no_inst"]]); lsb = int(r[ix["stall_long_sb"]])
        except ValueError:
            continue
        out.append((s, i, t, ni, lsb, cur, int(r[0]), r[1].strip()[:90]))
tot = sum(o[0] for o in out); toti = sum(o[1] for o in out)
print(f"samples {tot}  warp-instructions {toti}")
byfile = {}
for o in out: byfile[o[5]] = byfile.get(o[5], 0) + o[0]
print("by file:", {k: f"{100 * v / tot:.1f}%" for k, v in sorted(byfile.items(), key=lambda x: -x[1])})
for s, i, t, ni, lsb, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * s / tot:5.1f}% smp  {100 * i / toti:5.1f}% ins  thr {t:4.1f}  noinst {100 * ni / max(s, 1):3.0f}%  lsb {100 * lsb / max(s, 1):3.0f}%  {f}:{ln}  {src}")
