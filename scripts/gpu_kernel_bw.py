"""achieved HBM bandwidth of the elementwise kernels at cfg3 size (K=20, p=1000), algorithmic bytes / CUDA-event time.
usage: python scripts/gpu_kernel_bw.py [out.json]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gglasso_b200 import _lib
from gglasso_b200._engine import run_admm, _p
from gglasso_b200._lib import NPART

lib = _lib.load()
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
K, p = 20, 1000
rng = np.random.default_rng(0)
S = np.stack([np.cov(rng.standard_normal((p, 2 * p)), bias=True) for _ in range(K)])
Om0 = np.repeat(np.eye(p)[None], K, 0)
out = {"peak_gbs": peak, "kernels": {}}
A = 8.0 * K * p * p
for reg in ("GGL", "FGL"):
    st, _ = run_admm("mgl", S, Om0, None, None, lambda1=0.05, lambda2=0.01, reg=reg, tol=0.0, rtol=0.0, max_iter=3,
                     check_every=10 ** 9)
    nt = lib.gg_mgl_ntile(p)
    parts = torch.zeros((nt * nt, NPART), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    regi = 0 if reg == "GGL" else 1

    def prox():
        lib.gg_prox_mgl(_p(st.Omega_new), _p(st.Omega), None, _p(st.X), _p(st.Theta), None, _p(st.ctrl), 0.05, 0.01, regi,
                        K, p, _p(parts), stream)

    def buildw():
        lib.gg_build_w(_p(st.Theta), None, _p(st.X), _p(st.S), None, _p(st.ctrl), K, p, K, _p(st.W), stream)

    nu = lib.gg_mgl_upper_nparts(p)
    parts_u = torch.zeros((nu, NPART), dtype=torch.float64, device="cuda")

    def prox_upper():
        lib.gg_prox_mgl_upper(_p(st.Omega_new), _p(st.Omega), _p(st.X), _p(st.Theta), _p(st.ctrl), 0.05, 0.01, regi, K, p,
                              _p(parts_u), stream)

    def buildw_upper():
        lib.gg_build_w_upper(_p(st.Theta), _p(st.X), _p(st.S), None, _p(st.ctrl), K, p, _p(st.W), stream)

    def mirror():
        lib.gg_mirror_upper(_p(st.Theta), _p(st.X), K, p, stream)

    def recon():
        lib.gg_recon(_p(st.W), _p(st.eig.D), None, _p(st.ctrl), K, 0, K, p, _p(st.Omega_new), stream)

    # the *_upper kernels do the work of their full-matrix counterparts on half the entries: their rate is quoted on
    # the ALGORITHMIC bytes of the step they replace (5 A, 4 A), "moved" is what they touch
    for name, fn, nbytes in (("prox_mgl_" + reg, prox, 5 * A), ("build_w", buildw, 4 * A),
                             ("prox_mgl_upper_" + reg, prox_upper, 5 * A), ("build_w_upper", buildw_upper, 4 * A),
                             ("mirror_upper", mirror, 2 * A)):
        ts = []
        for r in range(12):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        out["kernels"][name] = {"ms": ms, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
        print(name, out["kernels"][name], flush=True)
    if reg == "GGL":
        st.omega_step()
        ts = []
        for r in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); recon(); b.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        out["kernels"]["recon"] = {"ms": ms, "tflops": K * p ** 3 / ms / 1e9}
        print("recon", out["kernels"]["recon"], flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/kernel_bw.json", "w"), indent=1)
