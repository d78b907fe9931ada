#!/bin/bash
# usage: gpu_run_multi.sh N   -- strong-scaling bench of the K=20 solve + cfg5 block distribution on N GPUs
N=${1:-8}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/multi_smi_$N.txt 2>&1
CFG5=${2:-1}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/multi_bench_n$N.json 2> gpurun_out/multi_bench_n$N.err
if [ "$CFG5" = "1" ]; then
python -c "
import sys; sys.path.insert(0,'.')
from oracle import ref_inputs
ref_inputs.load('cfg5'); print('cfg5 cached')" > gpurun_out/multi_inputs_$N.log 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/dist_block_sgl.py gpurun_out/multi_cfg5_n$N.json ) > gpurun_out/multi_cfg5_n$N.log 2>&1
fi
grep "^{" gpurun_out/multi_bench_n$N.json | head -c 400; tail -3 gpurun_out/multi_bench_n$N.err
if [ "$CFG5" = "1" ]; then tail -3 gpurun_out/multi_cfg5_n$N.log; fi
exit 0
