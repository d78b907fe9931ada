#!/bin/bash
# 2-GPU run: NCCL parity of the K-sharded solve, strong-scaling bench at N=2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r13_smi.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 500 -p no:cacheprovider -k "two_ranks_nccl" ) > gpurun_out/r13_nccl_test.log 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r13_bench_n2.json 2> gpurun_out/r13_bench_n2.err
tail -3 gpurun_out/r13_nccl_test.log; tail -5 gpurun_out/r13_bench_n2.err; head -c 400 gpurun_out/r13_bench_n2.json
