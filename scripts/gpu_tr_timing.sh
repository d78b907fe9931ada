#!/bin/bash
# Builds an instrumented copy of the library (-DTR_TIMING: %globaltimer stamps inside the sytrd kernels) next to the
# product library (run this where nvcc is, then gpu_tr_timing.py on the GPU box), which prints where the time of one column step goes.  Diagnostic only; the product build is untouched.
set -e
cd "$(dirname "$0")/../gglasso_b200/csrc"
mkdir -p ../../scripts/micro/timing_build
B=../../scripts/micro/timing_build
F="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a"
nvcc $F -DTR_TIMING -c gg_tridiag.cu -o $B/gg_tridiag.o
nvcc $F -c gg_jacobi.cu -o $B/gg_jacobi.o
nvcc $F -c gg_recon.cu -o $B/gg_recon.o
nvcc $F -fmad=false -c gg_elementwise.cu -o $B/gg_elementwise.o
nvcc $F -fmad=false -c gg_capi.cu -o $B/gg_capi.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $B/libtiming.so $B/*.o
cd ../..
echo built scripts/micro/timing_build/libtiming.so
# on the GPU box: GGLASSO_B200_LIB=scripts/micro/timing_build/libtiming.so python scripts/gpu_tr_timing.py
