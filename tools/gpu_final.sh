#!/bin/bash
# Final evidence run on the GPU box (under gpurun): full GPU test suite, default bench line, reference arm, ncu launch list,
# one full capture of the traversal kernel and one of every other kernel of a wave.
# Usage: tools/gpu_final.sh <tag>
set -u
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/${TAG}_pytest.log
python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference rc=$?"
SMALL="--steps 1 --warmup 1 --spp 2 --no-cpu-baseline"
python bench.py $SMALL > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $SMALL > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 2 -c 3 -f -o $OUT/${TAG}_trace python bench.py $SMALL > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu trace rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:k_raygen|k_path_shade|k_shadow_resolve|k_path_splat|k_film|k_rgb2spec" -s 0 -c 10 -f -o $OUT/${TAG}_others python bench.py $SMALL > $OUT/${TAG}_ncu3.log 2>&1
echo "ncu others rc=$?"
ls -la $OUT | tail -15
