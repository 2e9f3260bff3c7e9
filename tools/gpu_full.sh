#!/bin/bash
# full GPU suite, smoke, default bench (both arms) -- the round-end sequence
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/full_tests.log
tail -4 gpurun_out/full_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'])
print('cpu', d['cpu_baseline']); print('vit', {k:d['configs']['vitb16'][k] for k in ('value','ms_per_step')}); print('dropin', d.get('dropin'))
print(d['roofline'])
for r in d['rooflines']: print(r['kernel'][:60], '|', r['workload'][:60], '|', r.get('ms'), r.get('frac'))
print(d.get('summary'))
PY
