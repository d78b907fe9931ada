#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider ) > gpurun_out/r16_tests.log 2>&1
tail -4 gpurun_out/r16_tests.log
bash scripts/profile_gpu.sh > gpurun_out/r16_profile.log 2>&1
tail -14 gpurun_out/r16_profile.log
