#!/bin/bash
# run-to-run spread of the config-2 step inside one call
Q="--no-cpu-baseline --no-kernel-rooflines --no-vitb16 --no-dropin --sustained-steps 0"
for i in 1 2 3 4 5 6 7 8; do
  timeout 600 python bench.py --workload flickr8k --steps 100 --warmup 10 $Q 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run $i', d['value'], d['ms_per_step'], d.get('final_loss'), d['clocks']['sm_mhz'])"
done
