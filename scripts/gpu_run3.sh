#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( timeout 600 python scripts/gpu_sytrd_check.py gpurun_out/sytrd_check3.json ) > gpurun_out/r3_sytrd.log 2>&1
( timeout 300 python scripts/gpu_kernel_bw.py gpurun_out/kernel_bw3.json ) > gpurun_out/r3_bw.log 2>&1
( time python -m pytest tests/test_gpu_parity.py -q --timeout 900 -p no:cacheprovider -k "prox or golden or trajectory" ) > gpurun_out/r3_tests.log 2>&1
tail -12 gpurun_out/r3_sytrd.log; tail -6 gpurun_out/r3_bw.log; tail -3 gpurun_out/r3_tests.log
