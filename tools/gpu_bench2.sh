#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench rc=$?" >> gpurun_out/bench_graph.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu-baseline --no-kernel-rooflines > gpurun_out/bench_eager.json 2>> gpurun_out/bench_graph.err
cat gpurun_out/bench_graph.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); [d.pop(k,None) for k in ('rooflines','own_calls')]; print(json.dumps(d))"
tail -5 gpurun_out/bench_graph.err
python -c "import json; d=json.load(open('gpurun_out/bench_eager.json')); print('eager', d['value'], d['ms_per_step'], d['e2e'])"
