#!/bin/bash
# first GPU contact: smoke, parity tests (one process per file so a faulting kernel cannot poison the rest)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests/test_gpu_quant.py -m gpu -q --timeout 300 > gpurun_out/pytest_quant.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quant.log
timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout 300 > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gemm.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_quant.log; tail -25 gpurun_out/pytest_gemm.log
