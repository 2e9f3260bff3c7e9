#!/bin/bash
# N=2 A/B of the overlapped gradient all-reduce (config 2 headline + config 4 sub-run)
set -x
for flags in "" "--no-overlap-grads"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
    --steps 20 --warmup 5 --no-kernel-rooflines --no-cpu-baseline --no-dropin $flags > gpurun_out/r02_n2${flags// /_}.json 2> gpurun_out/r02_n2${flags// /_}.err
  tail -c 300 gpurun_out/r02_n2${flags// /_}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_n2${flags// /_}.json").read().strip().splitlines()[-1])
print("$flags", d["value"], d["ms_per_step"], "vit:", d["configs"]["vitb16"]["value"], d["configs"]["vitb16"]["ms_per_step"])
PY
done
