#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 120 python tools/prof_kernels.py all > gpurun_out/prof_plain.log 2>&1 && \
timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_kernels.csv python tools/prof_kernels.py all > gpurun_out/ncu1.log 2>&1
timeout 120 python tools/prof_kernels.py gemm > gpurun_out/prof_plain2.log 2>&1 && \
timeout 300 $NCU --set full --import-source on -k regex:tgemm_kernel -c 14 -o gpurun_out/prof_gemm python tools/prof_kernels.py gemm > gpurun_out/ncu2.log 2>&1
timeout 120 python tools/prof_kernels.py stream > gpurun_out/prof_plain3.log 2>&1 && \
timeout 300 $NCU --set full --import-source on -k regex:'select_pass|filter_kernel|sample_|ternarize|unpack2|split_' -c 24 -o gpurun_out/prof_stream python tools/prof_kernels.py stream > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out | head -30; tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
