#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -x ) > gpurun_out/r11_tests.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r11_smoke.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r11_bench.json 2> gpurun_out/r11_bench.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/r11_bench_ref.json 2> gpurun_out/r11_bench_ref.err
tail -5 gpurun_out/r11_tests.log; tail -3 gpurun_out/r11_smoke.log; tail -3 gpurun_out/r11_bench.err; head -c 600 gpurun_out/r11_bench.json
