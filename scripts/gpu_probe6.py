import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200._engine import Eigh, to_dev
from gglasso_b200.datagen import synthetic_mgl
dev = torch.device("cuda"); out = {}
def tm(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for (M, p) in ((1, 50), (1, 100), (5, 100), (1, 144), (1, 160), (40, 64), (200, 40)):
    A0 = np.eye(p)[None] - synthetic_mgl(M, p, N=2 * p, seed=2, n_blocks=1)
    W = to_dev(A0, dev); e = Eigh(M, p, dev)
    def run():
        A = W.clone(); e.eigh(A)
    out[f"{M}x{p}_ms"] = tm(run)
    A = W.clone(); D = e.eigh(A).cpu().numpy(); V = A.cpu().numpy()
    out[f"{M}x{p}_resid"] = float(max(np.abs(A0[m] @ V[m].T - V[m].T * D[m]).max() for m in range(M)))
print(os.environ.get("GG_JACOBI_MAX"), json.dumps(out))
