#!/bin/bash
# A/B of one environment switch inside one gpurun call: tools/gpu_ab_env.sh VAR [workload] [steps]
VAR=$1; WL=${2:-vitb16}; ST=${3:-6}
Q="--no-cpu-baseline --no-kernel-rooflines --no-vitb16 --no-dropin --sustained-steps 0"
for v in 1 0 1 0; do
  env $VAR=$v timeout 600 python bench.py --workload $WL --steps $ST --warmup 3 $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$VAR=$v', '$WL', d['value'], d['ms_per_step'])"
done
