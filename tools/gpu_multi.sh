#!/bin/bash
# usage: gpu_multi.sh N  -- short timeouts: a hang must not burn N x GPU-minutes
N=${1:-2}
mkdir -p gpurun_out
run() { # name, extra args
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $N --steps 20 --warmup 5 $2 > gpurun_out/bench_n${N}_$1.json 2> gpurun_out/bench_n${N}_$1.err; echo "rc=$?" >> gpurun_out/bench_n${N}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_$1.json').read().strip().splitlines()[-1])
    print('$1', 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['config']['execution'])
except Exception as e:
    print('$1 FAILED', e)
PY
  tail -4 gpurun_out/bench_n${N}_$1.err
}
run eager "--no-graph" 29511
run graph "" 29512
