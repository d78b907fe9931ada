"""small-p timings: eigh cold/warm and whole-solver wall time for cfg1/cfg2-like problems vs the CPU oracle."""
import contextlib, io, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200 import ADMM_MGL, ADMM_SGL
from gglasso_b200._engine import Eigh, to_dev
from gglasso_b200.datagen import synthetic_mgl, synthetic_sgl
from oracle import admm_oracle as orc
out = {}
dev = torch.device("cuda")
def tm(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for (M, p) in ((1, 100), (5, 100), (1, 160), (800, 16)):
    W = to_dev(np.eye(p)[None] - synthetic_mgl(M, p, N=2 * p, seed=2, n_blocks=1 if p < 20 else None), dev)
    e = Eigh(M, p, dev)
    def cold():
        A = W.clone(); e.eigh(A)
    out[f"eigh_cold_{M}x{p}_ms"] = tm(cold)
    A = W.clone(); e.eigh(A); V = A.clone()
    W2 = W + 1e-3 * torch.randn_like(W); W2 = (W2 + W2.transpose(1, 2)) / 2
    def warm():
        A = W2.clone(); Vw = V.clone(); e.eigh(A, warm=Vw)
    out[f"eigh_warm_{M}x{p}_ms"] = tm(warm)
# whole solves
S1 = synthetic_sgl(100, N=1000, seed=1)
S2 = synthetic_mgl(5, 100, N=1000, seed=1)
Om2 = np.repeat(np.eye(100)[None], 5, 0)
def wall(fn, n=3):
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()): r = fn()
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best, r
t, (sol, info) = wall(lambda: ADMM_SGL(S1, 0.05, np.eye(100), tol=1e-7, rtol=1e-7, measure=False))
tr, (rs, ri) = wall(lambda: orc.admm_sgl(S1, 0.05, np.eye(100), tol=1e-7, rtol=1e-7), n=2)
out["cfg1_sgl_gpu_s"] = t; out["cfg1_sgl_cpu_s"] = tr; out["cfg1_iters"] = ri["iterations"]
for reg in ("GGL", "FGL"):
    for lat in (False, True):
        t, _ = wall(lambda: ADMM_MGL(S2, 0.05, 0.01, reg, Om2, tol=1e-7, rtol=1e-7, latent=lat, mu1=0.1 if lat else None))
        tr, (rs, ri) = wall(lambda: orc.admm_mgl(S2, 0.05, 0.01, reg, Om2, tol=1e-7, rtol=1e-7, latent=lat, mu1=0.1 if lat else None), n=1)
        out[f"cfg2_{reg}_lat{int(lat)}_gpu_s"] = t; out[f"cfg2_{reg}_lat{int(lat)}_cpu_s"] = tr; out[f"cfg2_{reg}_lat{int(lat)}_iters"] = ri["iterations"]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe3.json", "w"), indent=1)
print(json.dumps(out, indent=1))
