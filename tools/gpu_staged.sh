#!/bin/bash
# staged vs fused shading: parity tests, then C3 (and C1) with both modes and every variants/*.so.  usage: tools/gpu_staged.sh <tag>
set -u
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_path.py tests/test_lights.py tests/test_gpu_full_size.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:14s} {d['value']:7.1f} Mpaths/s  {d['mrays_per_s']:7.1f} Mrays/s  e2e {d['e2e']['value']:7.1f}  {d['ms_per_step']:8.2f} ms  film {d['film_checksum']:.6f} launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for cfg in C3 C1; do
  A="--config $cfg --steps 3 --warmup 2 --no-cpu-baseline"; [ $cfg = C3 ] && A="$A --spp 64"
  for sm in 1 2; do
    timeout 300 python bench.py $A --shade-mode $sm > $OUT/${TAG}_${cfg}_sm$sm.json 2> $OUT/${TAG}_${cfg}_sm$sm.err; summ ${cfg}_mode$sm $OUT/${TAG}_${cfg}_sm$sm.json
  done
done
for v in variants/*.so; do
  name=$(basename $v .so)
  CRT_B200_LIB=$PWD/$v timeout 300 python bench.py --config C3 --spp 64 --steps 3 --warmup 2 --no-cpu-baseline > $OUT/${TAG}_C3_$name.json 2> $OUT/${TAG}_C3_$name.err
  summ C3_$name $OUT/${TAG}_C3_$name.json
done
