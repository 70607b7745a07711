#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> <kernel regex> <skip> <count> [bench args]   (runs the plain command first, then ncu --set full)
set -u
TAG=$1; KREGEX=$2; SKIP=$3; COUNT=$4; shift 4
OUT=gpurun_out; mkdir -p $OUT
SMALL="--steps 1 --warmup 1 --spp 2 --no-cpu-baseline $*"
python bench.py $SMALL > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $COUNT -f -o $OUT/${TAG} python bench.py $SMALL > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu.log
