#!/bin/sh
# Debug build of the library with the eigensolver's section timers enabled.
cd "$(dirname "$0")/../dataassimbench_b200/csrc" && mkdir -p ../../build && \
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DDAB_EIG_TIMING -Xcompiler -fPIC -shared \
  -o ../../build/libdab_timing.so engine.cu l96_forecast.cu etkf_contract.cu etkf_eig.cu etkf_transform.cu etkf_batched.cu fp64_peak.cu
