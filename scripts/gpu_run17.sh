#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_facade.py -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r17_tests.log 2>&1
( timeout 300 python scripts/gpu_kernel_bw.py gpurun_out/kernel_bw17.json ) > gpurun_out/r17_bw.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -q --timeout 900 -p no:cacheprovider ) > gpurun_out/r17_tests_cfg.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r17_bench.json 2> gpurun_out/r17_bench.err
tail -4 gpurun_out/r17_tests.log; cat gpurun_out/r17_bw.log; tail -4 gpurun_out/r17_tests_cfg.log; tail -3 gpurun_out/r17_bench.err; head -c 300 gpurun_out/r17_bench.json
