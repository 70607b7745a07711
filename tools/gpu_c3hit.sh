#!/bin/bash
set -u
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
SMALL="--config C3 --steps 1 --warmup 1 --spp 8 --no-cpu-baseline"
python bench.py $SMALL > $OUT/${TAG}_c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_path_hit -s 1 -c 1 -f -o $OUT/${TAG}_c3_hit python bench.py $SMALL > $OUT/${TAG}_c3_ncu2.log 2>&1; echo "ncu hit rc=$?"
