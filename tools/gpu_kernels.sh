#!/bin/bash
# Kernel-level evidence for one config (under gpurun): ncu launch list (time per kernel) and one `--set full` capture of the non-traversal
# kernels (and optionally the traversal kernel) of a short run.
# usage: tools/gpu_kernels.sh <tag> <config> <spp> [trace]
set -u
TAG=$1; CFG=$2; SPP=$3; WITH_TRACE=${4:-}
OUT=gpurun_out; mkdir -p $OUT
SMALL="--config $CFG --steps 1 --warmup 1 --spp $SPP --no-cpu-baseline"
python bench.py $SMALL > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $SMALL > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:k_raygen|k_path_shade|k_shadow_resolve|k_path_splat|k_path_nee" -s 0 -c 12 -f -o $OUT/${TAG}_others python bench.py $SMALL > $OUT/${TAG}_ncu3.log 2>&1
echo "ncu others rc=$?"
if [ -n "$WITH_TRACE" ]; then
  ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 0 -c 4 -f -o $OUT/${TAG}_trace python bench.py $SMALL > $OUT/${TAG}_ncu2.log 2>&1
  echo "ncu trace rc=$?"
fi
python - $OUT/${TAG}_launches.csv <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki, vi, ui, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Metric Name")
acc = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum": continue
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    t, n = acc.get(r[ki], (0.0, 0)); acc[r[ki]] = (t + ms, n + 1)
tot = sum(t for t, _ in acc.values())
for k, (t, n) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"{t:10.3f} ms {n:5d} {100 * t / tot:5.1f}%  {k[:100]}")
PY
