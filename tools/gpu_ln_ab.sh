#!/bin/bash
# A/B of the fused LayerNorm kernels inside the harness (configs 2 and 4), after the parity tests
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layernorm.py tests/test_gpu_models.py -x -q > gpurun_out/ln_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ln_tests.log
tail -5 gpurun_out/ln_tests.log
Q="--no-cpu-baseline --no-kernel-rooflines --no-vitb16 --no-dropin --sustained-steps 0"
for ln in 1 0; do
  ATQ_FUSED_LAYERNORM=$ln timeout 600 python bench.py --workload vitb16 --steps 6 --warmup 3 $Q > gpurun_out/ln_vit_$ln.json 2> gpurun_out/ln_vit_$ln.err
  ATQ_FUSED_LAYERNORM=$ln timeout 600 python bench.py --workload flickr8k --steps 50 --warmup 5 $Q > gpurun_out/ln_fl_$ln.json 2> gpurun_out/ln_fl_$ln.err
done
for f in gpurun_out/ln_vit_*.json gpurun_out/ln_fl_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["ms_per_step"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
