#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an .ncu-rep (needs -lineinfo + --import-source on).
usage: tools/ncu_lines.py rep.ncu-rep [top_n] [kernel_block_index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
blocks, cur, fname, seen = [], None, None, set()
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        key = (fname, r[1])
        if cur is None or key in seen:           # a (file, function) pair repeating = next profiled launch
            cur = {"name": r[1], "hdr": None, "lines": []}
            blocks.append(cur)
            seen = set()
        seen.add(key)
    elif r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and r[0].isdigit() and len(r) == len(cur["hdr"]) and r[7].isdigit():
        cur["lines"].append((fname, r))
b = blocks[which]
h = b["hdr"]
ii, si = h.index("Instructions Executed"), h.index("# Samples")
ti = h.index("Thread Instructions Executed")
print(b["name"], f"({len(blocks)} kernel blocks)")
tot_i = sum(int(r[ii]) for _, r in b["lines"]); tot_s = sum(int(r[si]) for _, r in b["lines"])
print(f"total warp instructions {tot_i}, samples {tot_s}")
print(" inst%  smpl%  lanes  file:line  source")
for f, r in sorted(b["lines"], key=lambda x: -int(x[1][ii]))[:top]:
    n = int(r[ii])
    print(f"{100 * n / tot_i:6.2f} {100 * int(r[si]) / max(tot_s, 1):6.2f} {int(r[ti]) / max(n, 1):6.1f}  {f}:{r[0]:>4}  {r[1].strip()[:120]}")
