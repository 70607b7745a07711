#!/bin/bash
# Final single-GPU evidence of a round (under gpurun).  Part A: full GPU test suite, smoke, every single-GPU config's bench line with its
# CPU baseline, the reference arm, the ncu launch list of the headline command.  Part B (separate call, the .ncu-rep files are large):
# `--set full` captures of the traversal kernel on C2 and C5 and of the shading kernels on C2 / C3.
# Part C: `--set full` captures of the kernels that contain the traversal where the octree is one leaf (C3 staged: k_path_hit and the
# material kernels; C1 fused: k_path_shade) and of k_shadow_resolve on C3.
# usage: tools/gpu_final2.sh <tag> A|B|C
set -u
TAG=$1; PART=$2
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt
SMALL="--steps 1 --warmup 1 --spp 2 --no-cpu-baseline"
if [ "$PART" = "A" ]; then
  timeout 1800 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/${TAG}_pytest.log
  timeout 600 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
  timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
  for c in C1 C3 C4; do
    timeout 900 python bench.py --config $c --steps 3 --warmup 3 --cpu-seconds 8 > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err; echo "bench $c rc=$?"
  done
  timeout 1500 python bench.py --config C5 --spp 16 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_C5_1gpu_16spp.json 2> $OUT/${TAG}_bench_C5.err; echo "bench C5 rc=$?"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference rc=$?"
  python bench.py $SMALL > $OUT/${TAG}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $SMALL > $OUT/${TAG}_ncu1.log 2>&1
  echo "ncu launches rc=$?"
  for f in $OUT/${TAG}_bench.json $OUT/${TAG}_bench_C1.json $OUT/${TAG}_bench_C3.json $OUT/${TAG}_bench_C4.json $OUT/${TAG}_bench_C5_1gpu_16spp.json; do
    python - $f <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); cb = d.get("cpu_baseline") or {}
    print(f"{sys.argv[1].split('_bench')[-1]:22s} {d['value']:.1f} Mpaths/s {d['mrays_per_s']:.1f} Mrays/s | e2e {d['e2e']['value']:.1f} / resident {d['e2e']['resident']['value']:.1f} | {d['ms_per_step']:.2f} ms/step | cpu {cb.get('value')}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
  rm -f $OUT/${TAG}_*.err
elif [ "$PART" = "C" ]; then
  C3S="--config C3 --steps 1 --warmup 1 --spp 8 --no-cpu-baseline"
  ncu --set full --clock-control none -k regex:k_path_hit -s 0 -c 2 -f -o $OUT/${TAG}_C3_hit python bench.py $C3S > $OUT/${TAG}_ncu6.log 2>&1; echo "ncu C3 hit rc=$?"
  ncu --set full --clock-control none -k regex:k_path_shade_mat -s 0 -c 3 -f -o $OUT/${TAG}_C3_mat python bench.py $C3S > $OUT/${TAG}_ncu7.log 2>&1; echo "ncu C3 mat rc=$?"
  ncu --set full --clock-control none -k regex:k_shadow_resolve -s 0 -c 1 -f -o $OUT/${TAG}_C3_resolve python bench.py $C3S > $OUT/${TAG}_ncu8.log 2>&1; echo "ncu C3 resolve rc=$?"
  ncu --set full --clock-control none -k regex:k_path_shade -s 0 -c 2 -f -o $OUT/${TAG}_C1_shade python bench.py --config C1 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu9.log 2>&1; echo "ncu C1 shade rc=$?"
  ls -la $OUT | grep ncu-rep
  du -sh $OUT
else
  ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 0 -c 3 -f -o $OUT/${TAG}_trace python bench.py $SMALL > $OUT/${TAG}_ncu2.log 2>&1; echo "ncu trace rc=$?"
  ncu --set full --clock-control none -k "regex:k_raygen|k_path_shade|k_shadow_resolve|k_path_splat" -s 0 -c 4 -f -o $OUT/${TAG}_others python bench.py $SMALL > $OUT/${TAG}_ncu3.log 2>&1; echo "ncu others rc=$?"
  ncu --set full --clock-control none -k regex:k_trace_wide -s 0 -c 2 -f -o $OUT/${TAG}_C5_trace python bench.py --config C5 $SMALL > $OUT/${TAG}_ncu4.log 2>&1; echo "ncu C5 trace rc=$?"
  ls -la $OUT | grep ncu-rep
  du -sh $OUT
fi
