#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_models.py tests/test_gpu_canaries.py -x -q 2>&1 | tail -4
ATQ_SM100_LIB=$PWD/atq-multimodal_b200/atq/libatq_sm100_prof.so timeout 300 python tools/attn_prof.py parity > gpurun_out/attn_prof_parity4.txt 2>&1; cat gpurun_out/attn_prof_parity4.txt | tail -56
timeout 300 python tools/bench_attention.py 2>&1 | tail -8
