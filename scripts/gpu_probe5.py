"""host-side phase timing of one ADMM_MGL call at cfg3 (where does e2e time go?)"""
import contextlib, io, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gglasso_b200._engine as eng
from gglasso_b200 import ADMM_MGL
from gglasso_b200.datagen import synthetic_mgl
S = synthetic_mgl(20, 1000, N=2000, seed=1234, kind="fused"); Om = np.repeat(np.eye(1000)[None], 20, 0)
T = {}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize()
        T[key] = T.get(key, 0.0) + time.perf_counter() - t0; return r
    setattr(obj, name, g)
wrap(eng.AdmmState, "__init__", "state_init(H2D+alloc)")
wrap(eng.AdmmState, "min_eig", "min_eig(PD check)")
wrap(eng.AdmmState, "asym_max", "asym_max")
wrap(eng.AdmmState, "final_omega", "final_omega")
import gglasso_b200.solver.admm_solver as am
wrap(am, "to_host", "to_host(D2H)")
wrap(am, "run_admm", "run_admm(total incl. init)")
for rep in range(2):
    T.clear()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        sol, info = ADMM_MGL(S, 0.05, 0.01, "FGL", Om, tol=1e-7, rtol=1e-7)
    T["TOTAL"] = time.perf_counter() - t0
print(json.dumps({k: round(v * 1e3, 1) for k, v in T.items()}, indent=1))
