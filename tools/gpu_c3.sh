#!/bin/bash
# C3 (analytic shapes, deep specular paths): launch list + full capture of the surface-record kernel.  usage: tools/gpu_c3.sh <tag>
set -u
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
SMALL="--config C3 --steps 1 --warmup 1 --spp 8 --no-cpu-baseline"
python bench.py $SMALL > $OUT/${TAG}_c3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_c3_launches.csv python bench.py $SMALL > $OUT/${TAG}_c3_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_path_hit -s 1 -c 1 -f -o $OUT/${TAG}_c3_hit python bench.py $SMALL > $OUT/${TAG}_c3_ncu2.log 2>&1; echo "ncu hit rc=$?"
