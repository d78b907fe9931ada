"""ADMM_SGL on single large matrices (the sizes of the reference's published benchmark):
power-law S, N = 1.1 p, lambda1 = 0.05, tol = rtol = 1e-7.  CPU oracle timed for p <= 2000; at p = 5000 the result
is checked through the graphical-lasso optimality conditions instead.  Writes gpurun_out/sgl_large.json."""
import contextlib, io, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200 import ADMM_SGL
from gglasso_b200.datagen import synthetic_sgl
from oracle import admm_oracle as orc
out = {"published_reference": {"p1000_tol1e-7": "3.62 s / 17 it (Opteron 6378, data/synthetic/bm5000.csv:45)",
                               "p5000_tol1e-7": "174.3 s / 9 it (Opteron 6378, data/synthetic/bm5000.csv:171)"}}
def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()): r = fn()
    torch.cuda.synchronize(); return time.perf_counter() - t0, r
wall(lambda: ADMM_SGL(synthetic_sgl(200, N=300, seed=0), 0.1, np.eye(200), max_iter=3))
for p in (1000, 2000, 5000):
    S = synthetic_sgl(p, N=int(1.1 * p), seed=1234, n_blocks=10)
    lam = 0.05
    t, (sol, info) = wall(lambda: ADMM_SGL(S, lam, np.eye(p), tol=1e-7, rtol=1e-7, measure=True))
    t2, (sol, info2) = wall(lambda: ADMM_SGL(S, lam, np.eye(p), tol=1e-7, rtol=1e-7))
    d = dict(gpu_s=t2, gpu_s_with_measure=t, iters=len(info["residual"]), status=info["status"],
             gpu_loop_s=float(np.sum(info["runtime"])), nnz=int(np.count_nonzero(sol["Theta"])))
    if p <= 2000:
        tc, (ref, ri) = wall(lambda: orc.admm_sgl(S, lam, np.eye(p), tol=1e-7, rtol=1e-7))
        d.update(cpu_s=tc, cpu_iters=ri["iterations"], rel_err_theta=float(np.linalg.norm(sol["Theta"] - ref["Theta"]) / np.linalg.norm(ref["Theta"])),
                 pattern_equal=bool(np.array_equal(sol["Theta"] != 0, ref["Theta"] != 0)))
    else:
        # optimality of the graphical lasso at Omega (exactly PD): |S - Omega^-1|_ij <= lam off the support of Theta
        # (up to the ADMM tolerance), (S - Omega^-1)_ij = -lam sign(Theta_ij) on it
        Oi = np.linalg.inv(sol["Omega"])
        Gm = S - Oi
        off = ~np.eye(p, dtype=bool)
        zero = (sol["Theta"] == 0) & off
        nz = (sol["Theta"] != 0) & off
        d.update(kkt_max_violation_zero=float(np.maximum(np.abs(Gm[zero]) - lam, 0).max()),
                 kkt_max_dev_support=float(np.abs(Gm[nz] + lam * np.sign(sol["Theta"][nz])).max()),
                 primal_gap_rel=float(np.linalg.norm(sol["Omega"] - sol["Theta"]) / np.linalg.norm(sol["Theta"])))
    out[f"sgl_p{p}"] = d
    print(f"sgl_p{p}", json.dumps(d), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/sgl_large.json", "w"), indent=1)
