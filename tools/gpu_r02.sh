#!/bin/bash
# Evidence run on the GPU box (under gpurun): GPU test suite, smoke, the bench line of every single-GPU config, the reference arm.
# usage: tools/gpu_r02.sh <tag> [configs...]      (default configs: C2 C1 C3 C4)
set -u
TAG=$1; shift
CONFIGS=${*:-C2 C1 C3 C4}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
for c in $CONFIGS; do
  extra=""; [ "$c" != "C2" ] && extra="--steps 3 --warmup 3 --cpu-seconds 8"
  timeout 900 python bench.py --config $c $extra > $OUT/${TAG}_bench_$c.json 2> $OUT/${TAG}_bench_$c.err; echo "bench $c rc=$?"
  python - $OUT/${TAG}_bench_$c.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]; cb = d.get("cpu_baseline") or {}
    print(f"  {d['value']:.1f} Mpaths/s {d['mrays_per_s']:.1f} Mrays/s | e2e {d['e2e']['value']:.1f} (resident {d['e2e']['resident']['value']:.1f}) | ms/step {d['ms_per_step']:.2f} | launches {d['gpu_launches']} | "
          f"trace share {r['kernel_share_of_step']:.3f} alg GB/s {r['achieved']:.0f} | cpu {cb.get('value')} x{d['value'] / cb['value'] if cb.get('value') else 0:.0f} | clocks {d['clocks']}")
except Exception as e:
    print("  FAILED", e)
PY
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference rc=$?"
tail -c 600 $OUT/${TAG}_bench_reference.json
