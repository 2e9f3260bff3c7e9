#!/bin/bash
# builds atq/libatq_sm100_prof.so: the normal objects + attention_sm100.cu compiled with -DATQ_ATTN_PROF
set -e
cd "$(dirname "$0")/.."
C=atq-multimodal_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -cudart static \
  -DATQ_ATTN_PROF -c $C/attention_sm100.cu -o $C/build/attention_prof.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o atq-multimodal_b200/atq/libatq_sm100_prof.so \
  $C/build/common.o $C/build/streaming.o $C/build/select.o $C/build/gemm_sm100.o $C/build/attention_prof.o $C/build/loss_sm100.o -ldl
