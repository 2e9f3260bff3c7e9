#!/bin/bash
# N>1: gradient bucket size of the overlapped all-reduce, config 2 only
N=${1:-8}
for mb in 24 8 4 24 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
    --steps 100 --warmup 10 --no-kernel-rooflines --no-cpu-baseline --no-dropin --no-vitb16 --sustained-steps 0 --bucket-mb $mb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bucket_mb $mb', d['value'], d['ms_per_step'])"
done
