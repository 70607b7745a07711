#!/bin/bash
# quick check of a change: selected GPU tests, then bench lines of the named configs.  usage: tools/gpu_quick.sh <tag> "<pytest targets>" "C3:--spp 64" "C2:" ...
set -u
TAG=$1; TESTS=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
if [ -n "$TESTS" ]; then
  timeout 1200 python -m pytest $TESTS -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
fi
for spec in "$@"; do
  cfg=${spec%%:*}; extra=${spec#*:}; name=$(echo "${cfg}_${extra}" | tr -c 'A-Za-z0-9\n' '_')
  timeout 600 python bench.py --config $cfg --steps 3 --warmup 2 --no-cpu-baseline $extra > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  python - "$cfg $extra" $OUT/${TAG}_$name.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:28s} {d['value']:7.1f} Mpaths/s  {d['mrays_per_s']:7.1f} Mrays/s  e2e {d['e2e']['value']:7.1f}  {d['ms_per_step']:8.2f} ms  film {d['film_checksum']:.6f} launches {d['gpu_launches']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
