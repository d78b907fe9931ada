"""small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import contextlib, io, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200 import ADMM_MGL, ADMM_SGL, block_SGL
from gglasso_b200._engine import eigh
from gglasso_b200.parallel import ADMM_MGL_dist, grid_search_device
from gglasso_b200.datagen import synthetic_mgl, synthetic_sgl
rng = np.random.default_rng(0)
def q(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
for p in (7, 45, 163, 190):
    A = rng.standard_normal((2, p, p)); A = A + A.transpose(0, 2, 1)
    D, Q = eigh(A); assert np.abs(A[0] @ Q[0] - Q[0] * D[0]).max() < 1e-9
    if p > 160:
        D, Q = eigh(A, nb2=32); assert np.abs(A[0] @ Q[0] - Q[0] * D[0]).max() < 1e-9
S = synthetic_mgl(3, 30, N=100, seed=1); Om = np.repeat(np.eye(30)[None], 3, 0)
for reg in ("GGL", "FGL"):
    q(ADMM_MGL, S, 0.1, 0.05, reg, Om, max_iter=6)
    q(ADMM_MGL, S, 0.1, 0.05, reg, Om, max_iter=4, latent=True, mu1=0.2, measure=True)
    ADMM_MGL_dist(S, 0.1, 0.05, reg, Om, max_iter=4)
q(ADMM_MGL, S, 0.1, 0.05, "GGL", Om, max_iter=3, stopping_criterion="kkt")
S2 = synthetic_mgl(2, 170, N=300, seed=2); Om2 = np.repeat(np.eye(170)[None], 2, 0)
q(ADMM_MGL, S2, 0.1, 0.05, "FGL", Om2, max_iter=3)
q(ADMM_MGL, S2, 0.1, 0.05, "GGL", Om2, max_iter=2, latent=True, mu1=0.2)
Ss = synthetic_sgl(60, N=100, seed=3)
q(ADMM_SGL, Ss, 0.1, np.eye(60), max_iter=5, lambda1_mask=np.ones((60, 60)))
q(ADMM_SGL, Ss, 0.1, np.eye(60), max_iter=4, latent=True, mu1=0.3)
q(block_SGL, Ss, 0.25, np.eye(60), max_iter=20)
grid_search_device(S, np.full(3, 100), "GGL", [0.2, 0.1], [0.05], tol=1e-4, rtol=1e-4)
torch.cuda.synchronize()
print("SANITIZE_RUN_OK")
