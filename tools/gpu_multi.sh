#!/bin/bash
# usage: gpu_multi.sh N [workload-args]  -- short timeouts: a hang must not burn N x GPU-minutes
N=${1:-2}
EXTRA=${2:-}
TAG=${3:-graph}
mkdir -p gpurun_out
start=$(date +%s)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "rc=$? wall=$(( $(date +%s) - start ))s" | tee -a gpurun_out/bench_n${N}_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_$TAG.json').read().strip().splitlines()[-1])
    print('$TAG N=$N', 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['config']['execution'], d['clocks'])
except Exception as e:
    print('$TAG FAILED', e)
PY
tail -3 gpurun_out/bench_n${N}_$TAG.err
