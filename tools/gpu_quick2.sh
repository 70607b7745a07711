#!/bin/bash
# like gpu_quick.sh, then the same bench lines for every variants/*.so on the FIRST spec.  usage: tools/gpu_quick2.sh <tag> "<pytest targets>" "C3:--spp 64" ...
set -u
TAG=$1; TESTS=$2; FIRST=$3
bash tools/gpu_quick.sh "$@"
OUT=gpurun_out
cfg=${FIRST%%:*}; extra=${FIRST#*:}
for v in variants/*.so; do
  [ -f "$v" ] || continue
  name=$(basename $v .so)
  CRT_B200_LIB=$PWD/$v timeout 600 python bench.py --config $cfg --steps 3 --warmup 2 --no-cpu-baseline $extra > $OUT/${TAG}_v_$name.json 2> $OUT/${TAG}_v_$name.err
  python - "$name: $cfg $extra" $OUT/${TAG}_v_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:28s} {d['value']:7.1f} Mpaths/s  {d['mrays_per_s']:7.1f} Mrays/s  e2e {d['e2e']['value']:7.1f}  {d['ms_per_step']:8.2f} ms  film {d['film_checksum']:.6f} launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
