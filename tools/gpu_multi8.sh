#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
bash tools/gpu_multi.sh $N "" graph
start=$(date +%s)
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload vitb16 --steps 3 --warmup 3 > gpurun_out/bench_vit_n$N.json 2> gpurun_out/bench_vit_n$N.err; echo "vit rc=$? wall=$(( $(date +%s) - start ))s"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_vit_n$N.json').read().strip().splitlines()[-1])
    print('vit N=$N value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
except Exception as e:
    print('vit FAILED', e)
PY
grep -v "Warning\|warn" gpurun_out/bench_vit_n$N.err | tail -3
start=$(date +%s)
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/bench_codec_1b.py > gpurun_out/codec_1b_n$N.json 2> gpurun_out/codec_1b_n$N.err; echo "codec rc=$? wall=$(( $(date +%s) - start ))s"
tail -c 1800 gpurun_out/codec_1b_n$N.json
