#!/bin/bash
# N=8 run of the default bench line on the final build (config 2 + config-4 sub-run), as the driver launches it
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
  --steps 20 --warmup 5 --no-kernel-rooflines --no-cpu-baseline --no-dropin > gpurun_out/r02_bench_n${N}_v2.json 2> gpurun_out/r02_bench_n${N}_v2.err
echo "rc=$?"; grep -v "Warn\|warn" gpurun_out/r02_bench_n${N}_v2.err | tail -3
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n${N}_v2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "vit:", d["configs"]["vitb16"]["value"], d["configs"]["vitb16"]["ms_per_step"], d["clocks"])
PY
