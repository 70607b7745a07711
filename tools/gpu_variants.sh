#!/bin/bash
# usage: tools/gpu_variants.sh <tag> [--pytest] [bench args]: optional GPU test suite, default-build bench, then every variants/*.so
set -u
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $OUT/${TAG}_gpu.txt
if [ "${1:-}" = "--pytest" ]; then
  shift
  timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
  tail -5 $OUT/${TAG}_pytest.log
fi
summ() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    r = d["roofline"]
    print(f"{sys.argv[1]:10s} {d['value']:7.1f} Mpaths/s  {d['mrays_per_s']:7.1f} Mrays/s  e2e {d['e2e']['value']:7.1f}  film {d['film_checksum']:.6f}  nodes/ray {r['nodes_per_ray']:.1f} tris/ray {r['tris_per_ray']:.1f} retraced {d['exact_retraced_rays']} trace share {r['kernel_share_of_step']:.3f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $OUT/${TAG}_default.json 2> $OUT/${TAG}_default.err; summ default $OUT/${TAG}_default.json
for v in variants/*.so; do
  name=$(basename $v .so)
  CRT_B200_LIB=$PWD/$v timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline "$@" > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  summ $name $OUT/${TAG}_$name.json
done
