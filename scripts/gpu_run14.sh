#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( time timeout 600 python scripts/gpu_small_configs.py gpurun_out/small_configs14.json ) > gpurun_out/r14_small.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q --timeout 600 -p no:cacheprovider -x ) > gpurun_out/r14_tests.log 2>&1
tail -20 gpurun_out/r14_small.log; tail -5 gpurun_out/r14_tests.log
