#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider ) > gpurun_out/r12_tests.log 2>&1
( time GG_STRESS_TESTS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 500 -p no:cacheprovider -k "concurrent" ) > gpurun_out/r12_stress.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-grid ) > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err
tail -5 gpurun_out/r12_tests.log; tail -3 gpurun_out/r12_stress.log; tail -3 gpurun_out/r12_bench.err; head -c 300 gpurun_out/r12_bench.json
