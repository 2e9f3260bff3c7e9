#!/bin/bash
# N=2 run of the default bench line on the final build (config 2 + config-4 sub-run)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
  --steps 20 --warmup 5 --no-kernel-rooflines --no-cpu-baseline --no-dropin > gpurun_out/r02_bench_n2_v3.json 2> gpurun_out/r02_bench_n2_v3.err
echo "rc=$?"; tail -c 400 gpurun_out/r02_bench_n2_v3.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n2_v3.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "vit:", d["configs"]["vitb16"]["value"], d["configs"]["vitb16"]["ms_per_step"], d["clocks"])
PY
