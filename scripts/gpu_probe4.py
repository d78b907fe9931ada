"""achieved HBM bandwidth of the elementwise kernels at cfg3 size (K=20, p=1000), CUDA-event timed, inputs > L2."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200 import _lib
from gglasso_b200._engine import _p
lib = _lib.load(); dev = torch.device("cuda")
K, p = 20, 1000
A = 8.0 * K * p * p
def sym(scale):
    x = torch.randn(K, p, p, dtype=torch.float64, device=dev) * scale
    return ((x + x.transpose(1, 2)) / 2).contiguous()
Om, Omp, X, S, L, Th, W = sym(0.1) + torch.eye(p, device=dev, dtype=torch.float64), sym(0.1), sym(0.05), sym(0.1), sym(0.01), sym(0.1), sym(0.1)
ctrl = torch.zeros(16, dtype=torch.float64, device=dev); ctrl[0] = 1; ctrl[1] = 1
nt = lib.gg_mgl_ntile(p); parts = torch.zeros(nt * nt, 5, dtype=torch.float64, device=dev)
parts2 = torch.zeros(lib.gg_sgl_nparts(p, K) * K, 5, dtype=torch.float64, device=dev)
st = torch.cuda.current_stream().cuda_stream
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6531.9
def tm(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
out = {}
def rec(name, ms, nA):
    gbs = nA * A / ms / 1e6
    out[name] = {"ms": ms, "algorithmic_GB": nA * A / 1e9, "GBps": gbs, "frac_of_measured_hbm": gbs / peak}
rec("build_w (4A)", tm(lambda: lib.gg_build_w(_p(Th), None, _p(X), _p(S), None, _p(ctrl), K, p, K, _p(W), st)), 4)
for reg, nm in ((0, "GGL"), (1, "FGL")):
    rec(f"prox_mgl {nm} fused dual+norms (5A)", tm(lambda: lib.gg_prox_mgl(_p(Om), _p(Omp), None, _p(X), _p(Th), None, _p(ctrl), 0.05, 0.01, reg, K, p, _p(parts), st)), 5)
    rec(f"prox_mgl {nm} latent (5A: Om,L,X -> Th,C)", tm(lambda: lib.gg_prox_mgl(_p(Om), _p(Omp), _p(L), _p(X), _p(Th), _p(W), _p(ctrl), 0.05, 0.01, reg, K, p, _p(parts), st)), 5)
rec("dual_update latent (6A)", tm(lambda: lib.gg_dual_update(_p(X), _p(Om), _p(Omp), _p(Th), _p(L), _p(ctrl), K, p, K, 0, _p(parts2), st)), 6)
ctrlK = torch.zeros(K, 16, dtype=torch.float64, device=dev); ctrlK[:, 0] = 1; ctrlK[:, 1] = 1
rec("prox_sgl fused (5A, 20 problems)", tm(lambda: lib.gg_prox_sgl(_p(Om), _p(Omp), None, _p(X), _p(Th), None, _p(ctrlK), 0.05, None, K, p, _p(parts2), None, st)), 5)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe4.json", "w"), indent=1)
print(json.dumps(out, indent=1))
