#!/bin/bash
# config 4 at N GPUs, dense vs sparse RPB gradient all-reduce, then the default config-2 bench at N
N=${1:-2}
mkdir -p gpurun_out
for tag in dense sparse; do
  extra=""; [ $tag = sparse ] && extra="--sparse-grads"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload vitb16 --steps 4 --warmup 3 --no-cpu-baseline --no-kernel-rooflines $extra > gpurun_out/vit_n${N}_$tag.json 2> gpurun_out/vit_n${N}_$tag.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/vit_n${N}_$tag.json').read().strip().splitlines()[-1])
    print('vit $tag N=$N', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
except Exception as e:
    print('$tag FAILED', e)
PY
done
bash tools/gpu_multi.sh $N "--no-cpu-baseline --no-kernel-rooflines" final2
