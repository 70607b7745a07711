#!/bin/bash
# Runs on the GPU box (under gpurun): tests, headline bench, ncu launch list + one full capture of the traversal kernel.
# Usage: tools/gpu_profile.sh <tag> [extra bench args]
set -u
TAG=${1:-r01}; shift || true
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
python bench.py "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
SMALL="--steps 1 --warmup 1 --spp 2 --no-cpu-baseline $*"
python bench.py $SMALL > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py $SMALL > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
python bench.py $SMALL > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_trace} -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-3} -f -o $OUT/${TAG}_trace python bench.py $SMALL > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
