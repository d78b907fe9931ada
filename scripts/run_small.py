import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200._engine import Eigh, to_dev
from gglasso_b200.datagen import synthetic_mgl
dev = torch.device("cuda")
M, p = 5, 100
W = to_dev(np.eye(p)[None] - synthetic_mgl(M, p, N=2 * p, seed=2), dev)
e = Eigh(M, p, dev)
for _ in range(3):
    A = W.clone(); e.eigh(A)
torch.cuda.synchronize()
ws = e.ws.cpu().numpy()
print("done")
