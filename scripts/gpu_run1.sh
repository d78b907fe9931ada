#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi1.txt 2>&1
( time python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_facade.py -q --timeout 1200 -p no:cacheprovider ) > gpurun_out/t1_configs.log 2>&1
( time python -m pytest tests/test_gpu_parity.py -q --timeout 600 -p no:cacheprovider -k "pack_unpack or check_every or large_K or k_sharded or golden or trajectory" ) > gpurun_out/t1_parity.log 2>&1
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench1.json 2> gpurun_out/bench1.err
( time python bench.py --impl reference --steps 4 --warmup 1 ) > gpurun_out/bench1_ref.json 2> gpurun_out/bench1_ref.err
tail -5 gpurun_out/t1_configs.log gpurun_out/t1_parity.log
