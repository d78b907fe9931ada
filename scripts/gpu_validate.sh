#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider ) > gpurun_out/r18_tests.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r18_smoke.log 2>&1
tail -6 gpurun_out/r18_tests.log; tail -4 gpurun_out/r18_smoke.log
