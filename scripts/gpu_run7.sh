#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( timeout 300 python scripts/gpu_sytrd_check.py gpurun_out/sytrd_check10.json ) > gpurun_out/r10_sytrd.log 2>&1
grep "_ms\|'ms'" gpurun_out/r10_sytrd.log | tail -20
