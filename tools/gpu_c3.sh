#!/bin/bash
# C3 (analytic shapes, deep specular paths): launch list + full captures of the shading kernels.  usage: tools/gpu_c3.sh <tag>
set -u
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
SMALL="--config C3 --steps 1 --warmup 1 --spp 8 --no-cpu-baseline"
python bench.py $SMALL > $OUT/${TAG}_c3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 1200 --csv --log-file $OUT/${TAG}_c3_launches.csv python bench.py $SMALL > $OUT/${TAG}_c3_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_path_hit -s 1 -c 2 -f -o $OUT/${TAG}_c3_hit python bench.py $SMALL > $OUT/${TAG}_c3_ncu2.log 2>&1; echo "ncu hit rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_path_shade_mat -s 3 -c 3 -f -o $OUT/${TAG}_c3_mat python bench.py $SMALL > $OUT/${TAG}_c3_ncu3.log 2>&1; echo "ncu mat rc=$?"
du -sh $OUT
