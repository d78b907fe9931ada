"""device-resident ADMM iterations of the cfg3 shape (K=20, p=1000, FGL) for ncu: python scripts/profile_step.py [iters]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gglasso_b200._engine import run_admm
K, p = 20, 1000
rng = np.random.default_rng(0)
S = np.stack([np.cov(rng.standard_normal((p, 2 * p)), bias=True) for _ in range(K)])
Om0 = np.repeat(np.eye(p)[None], K, 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
st, _ = run_admm("mgl", S, Om0, None, None, lambda1=0.05, lambda2=0.01, reg="FGL", tol=0.0, rtol=0.0, max_iter=n,
                 check_every=10 ** 9)
torch.cuda.synchronize()
print("done", n)
