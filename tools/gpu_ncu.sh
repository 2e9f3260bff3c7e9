#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python tools/prof_kernels.py all > gpurun_out/prof_plain.log 2>&1 && \
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_kernels.csv python tools/prof_kernels.py all > gpurun_out/ncu1.log 2>&1
python tools/prof_kernels.py gemm > gpurun_out/prof_plain2.log 2>&1 && \
$NCU --set full --import-source on -k regex:tgemm_kernel -c 12 -o gpurun_out/prof_gemm python tools/prof_kernels.py gemm > gpurun_out/ncu2.log 2>&1
python tools/prof_kernels.py stream > gpurun_out/prof_plain3.log 2>&1 && \
$NCU --set full --import-source on -k regex:'select_pass|ternarize|unpack2|split_|abs_stats' -c 12 -o gpurun_out/prof_stream python tools/prof_kernels.py stream > gpurun_out/ncu3.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/bench_plain.log 2>&1 && \
$NCU --metrics gpu__time_duration.sum --csv -c 4000 --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out; tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log gpurun_out/ncu4.log
