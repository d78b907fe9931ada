#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( timeout 600 python scripts/gpu_sytrd_check.py gpurun_out/sytrd_check.json ) > gpurun_out/r2_sytrd.log 2>&1
( GG_TR_OLD=1 timeout 300 python scripts/gpu_sytrd_check.py gpurun_out/sytrd_check_old.json ) > gpurun_out/r2_sytrd_old.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck python - <<'PY'
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from gglasso_b200._engine import eigh
rng = np.random.default_rng(0)
for M, p in ((2, 200), (1, 333)):
    A = rng.standard_normal((M, p, p)); A = (A + A.transpose(0, 2, 1)) / 2
    D, Q = eigh(A)
    print(M, p, np.abs(D - np.linalg.eigvalsh(A)).max())
PY
) > gpurun_out/r2_sanitizer.log 2>&1
( time python -m pytest tests/test_gpu_parity.py -q --timeout 900 -p no:cacheprovider -k "eigh" ) > gpurun_out/r2_eigh_tests.log 2>&1
python - <<'PY' > gpurun_out/r2_fingerprint.log 2>&1
import sys; sys.path.insert(0, ".")
from oracle import ref_inputs
import os
fp = "tests/golden/large_inputs.json"
for n in ("cfg4", "cfg3_small", "cfg4_small", "cfg1"):
    S = ref_inputs.load(n, cache=False)
    print(n, ref_inputs.check_fingerprint(n, S, fp))
PY
tail -3 gpurun_out/r2_sytrd.log gpurun_out/r2_eigh_tests.log gpurun_out/r2_fingerprint.log
