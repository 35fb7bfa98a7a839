#!/bin/sh
# Debug build of the library with the K4 section timers enabled (Jacobi: DAB_EIG_TIMING into lam_out;
# Newton-Schulz: DAB_NS_TIMING, printed by the kernel) and the forecast's per-CTA trace (DAB_L96_TRACE).
cd "$(dirname "$0")/../dataassimbench_b200/csrc" && mkdir -p ../../build && \
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DDAB_EIG_TIMING -DDAB_NS_TIMING -DDAB_L96_TRACE -DDAB_BATCH_TIMING -Xcompiler -fPIC -shared \
  -o ../../build/libdab_timing.so engine.cu l96_forecast.cu l96_forecast_v3.cu l96_forecast_v6.cu l63_forecast.cu var3d_cg.cu etkf_contract.cu etkf_eig.cu etkf_ns.cu etkf_transform.cu etkf_batched.cu fp64_peak.cu
