"""multi-stream lambda grid at a size that takes the blocked back-transformation path (p >= 256)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200.datagen import synthetic_mgl
from gglasso_b200.parallel import grid_search_device
p = int(os.environ.get("PP", 300)); K = int(os.environ.get("PK", 4)); ns = int(os.environ.get("NS", 3))
S = synthetic_mgl(K, p, N=2 * p, seed=5)
nl = int(os.environ.get('NL', 3))
l1, l2 = np.logspace(-0.5, -1.5, nl), np.logspace(-1, -2, int(os.environ.get('NL2', 3)))
sc1, it1, ix1, _ = grid_search_device(S, np.full(K, 2 * p), "GGL", l1, l2, gamma=0.1, tol=1e-5, rtol=1e-5, n_streams=1)
sc2, it2, ix2, _ = grid_search_device(S, np.full(K, 2 * p), "GGL", l1, l2, gamma=0.1, tol=1e-5, rtol=1e-5, n_streams=ns)
torch.cuda.synchronize()
print("GRID_OK", np.abs(sc1 - sc2).max(), it1.sum(), it2.sum(), ix1, ix2)
