"""where the end-to-end time of one public ADMM_MGL call goes at cfg3 size (host timers around the phases of
gglasso_b200/solver/admm_solver.py, with a device synchronisation after each phase).
usage: python scripts/gpu_e2e_breakdown.py [out.json]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gglasso_b200 import _engine as eng
from gglasso_b200 import ADMM_MGL

K, p, iters = 20, 1000, 10
rng = np.random.default_rng(0)
S = np.stack([np.cov(rng.standard_normal((p, 2 * p)), bias=True) for _ in range(K)])


def pinned(a):
    h = torch.empty(a.shape, dtype=torch.float64, pin_memory=True).numpy()
    h[...] = a
    return h


S, Om0 = pinned(S), pinned(np.repeat(np.eye(p)[None], K, 0))
eng.warmup()
out = {}
for rep in range(3):
    t = {}
    sync = torch.cuda.synchronize
    sync(); t0 = time.perf_counter()
    st = eng.AdmmState(S, Om0, None, None, K, 1.0, iters, False)
    sync(); t["state: uploads + buffers + workspace"] = time.perf_counter() - t0
    del st
    sync(); t0 = time.perf_counter()
    st, res = eng.run_admm("mgl", S, Om0, None, None, lambda1=0.05, lambda2=0.01, reg="FGL", tol=0.0, rtol=0.0,
                           max_iter=iters, check_symmetric=True)
    sync(); t["run_admm (state + symmetric checks + loop + mirror)"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    Om = st.final_omega(res["iters"])
    a = [st.asym_max(A) for A in (Om, st.Theta)]
    sync(); t["final_omega + symmetry checks"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    pd = st.posdef_async(st.Theta, res)
    sync(); t["PD certificate"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    outs = eng.to_host_many([Om, st.Theta, st.X])
    sync(); t["D2H of Omega, Theta, X"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    L = np.zeros((K, p, p))
    t["np.zeros for L"] = time.perf_counter() - t0
    del outs, st, Om, L
    sync(); t0 = time.perf_counter()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        sol, info = ADMM_MGL(S, 0.05, 0.01, "FGL", Om0, tol=0.0, rtol=0.0, max_iter=iters)
    sync(); t["public call total"] = time.perf_counter() - t0
    del sol
    out[f"rep{rep}"] = {k: round(v * 1e3, 2) for k, v in t.items()}
    print(out[f"rep{rep}"], flush=True)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/e2e_breakdown.json", "w"), indent=1)
