#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q --timeout 600 -p no:cacheprovider ) > gpurun_out/r15_tests.log 2>&1
( timeout 300 python scripts/gpu_kernel_bw.py gpurun_out/kernel_bw15.json ) > gpurun_out/r15_bw.log 2>&1
( GG_RECON_OLD=1 timeout 300 python scripts/gpu_kernel_bw.py gpurun_out/kernel_bw15_old.json ) > gpurun_out/r15_bw_old.log 2>&1
( timeout 300 compute-sanitizer --tool memcheck python - <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from gglasso_b200 import _lib
from gglasso_b200._engine import Eigh, to_dev
rng = np.random.default_rng(0)
for M, p in ((2, 100), (1, 258)):
    A = rng.standard_normal((M, p, p)); A = (A + A.transpose(0, 2, 1)) / 2
    D, Q = np.linalg.eigh(A)
    At = to_dev(A, torch.device("cuda", 0)); e = Eigh(M, p, torch.device("cuda", 0)); e.eigh(At)
    out = torch.empty_like(At); e.recon(At, out, 2)
    print(M, p, np.abs(out.cpu().numpy() - A).max())
PY
) > gpurun_out/r15_sanitizer.log 2>&1
tail -4 gpurun_out/r15_tests.log; grep recon gpurun_out/r15_bw.log gpurun_out/r15_bw_old.log; tail -4 gpurun_out/r15_sanitizer.log
