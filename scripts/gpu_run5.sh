#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( timeout 600 python scripts/gpu_sytrd_check.py gpurun_out/sytrd_check6.json ) > gpurun_out/r6_sytrd.log 2>&1
grep "_ms\|phases\|'ms'" gpurun_out/r6_sytrd.log | tail -20
