#!/bin/bash
# Build tuning variants of libcrt_b200.so (compile-time constants of k_trace_wide) into variants/*.so, then restore the default build.
# usage: tools/build_variants.sh "name1:-DCRT_WIDE_LEAF_WAIT=8" "name2:-DCRT_WIDE_STACK=12" ...
# run one with:  CRT_B200_LIB=$PWD/variants/name1.so python bench.py --no-cpu-baseline
set -eu
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  CRT_NVCC_DEFINES="$defs" python -c "from computational_ray_tracer_b200 import build; build.build(force=True)"
  cp computational_ray_tracer_b200/libcrt_b200.so variants/$name.so
  echo "built variants/$name.so with $defs"
done
python -c "from computational_ray_tracer_b200 import build; build.build(force=True)"
echo "default build restored"
