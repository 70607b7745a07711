#!/bin/bash
# Multi-GPU evidence (run under `gpurun --gpus 8`): the in-library NCCL reduce test on two GPUs, BASELINE config 4 at 1/2/4/8 GPUs,
# config 5 (4K @ 1024 spp, 10 M triangles) on 8 GPUs with both partitions, and the headline config at 1 and 8 GPUs.
# usage: tools/gpu_multi.sh <tag>
set -u
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/${TAG}_multi_pytest.log 2>&1; echo "multi pytest rc=$?"; tail -2 $OUT/${TAG}_multi_pytest.log
run() {   # run <name> <ngpus> <bench args...>
  local name=$1 n=$2; shift 2
  if [ "$n" = "1" ]; then
    timeout 1200 python bench.py --gpus 1 "$@" > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n "$@" > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  fi
  echo "$name rc=$?"
  python - $OUT/${TAG}_$name.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(f"   {d['n_gpus']} GPU(s): {d['value']:.1f} Mpaths/s {d['mrays_per_s']:.1f} Mrays/s | e2e {d['e2e']['value']:.1f} (resident {d['e2e']['resident']['value']:.1f}) | ms/step {d['ms_per_step']:.2f} | film {d['film_checksum']:.6f} | nccl {d.get('nccl')} | {d['config']['partition']}")
except Exception as e:
    print("   FAILED", e)
PY
}
for n in 1 2 4 8; do [ $n -le $NG ] && run C4_n$n $n --config C4 --steps 3 --warmup 3 --no-cpu-baseline; done
[ $NG -ge 8 ] && run C5_n8_spp 8 --config C5 --steps 2 --warmup 1 --no-cpu-baseline --partition spp
[ $NG -ge 8 ] && run C5_n8_tiles 8 --config C5 --steps 2 --warmup 1 --no-cpu-baseline --partition tiles
run C2_n1 1 --config C2 --steps 5 --warmup 3 --no-cpu-baseline
for n in 2 4 8; do [ $n -le $NG ] && run C2_n$n $n --config C2 --steps 5 --warmup 3 --no-cpu-baseline; done
for n in 2 4; do [ $n -le $NG ] && run C2_n${n}_tiles $n --config C2 --steps 5 --warmup 3 --no-cpu-baseline --partition tiles; done
[ $NG -ge 8 ] && run C2_n8_tiles 8 --config C2 --steps 5 --warmup 3 --no-cpu-baseline --partition tiles
ls -la $OUT | grep ${TAG}_ | head -40
