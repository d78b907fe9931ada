"""Time-to-tolerance of every BASELINE.json config on the GPU path vs the CPU oracle port (bounded CPU samples).
Writes gpurun_out/configs.json.  Usage: python scripts/bench_configs.py [--skip-cpu]"""
import contextlib, io, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200 import ADMM_MGL, ADMM_SGL, block_SGL, get_connected_components
from gglasso_b200.parallel import grid_search_dist, grid_search_device
from gglasso_b200.datagen import synthetic_mgl, synthetic_sgl
from oracle import admm_oracle as orc
skip_cpu = "--skip-cpu" in sys.argv
out = {"host_cores": os.cpu_count(), "gpu": torch.cuda.get_device_name(0)}

def wall(fn, n=1):
    best, r = 1e18, None
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()): r = fn()
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best, r

def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

class Out(dict):
    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(self, open("gpurun_out/configs.json", "w"), indent=1)
        print(k, json.dumps(v), flush=True)
out = Out(out)

# warm-up (library load, first-touch)
wall(lambda: ADMM_SGL(synthetic_sgl(50, N=200, seed=0), 0.1, np.eye(50)))

# cfg1
S = synthetic_sgl(100, N=1000, seed=1234); I = np.eye(100)
t, (sol, info) = wall(lambda: ADMM_SGL(S, 0.05, I, tol=1e-7, rtol=1e-7, measure=True), n=3)
tc, (ref, ri) = wall(lambda: orc.admm_sgl(S, 0.05, I, tol=1e-7, rtol=1e-7), n=2)
out["cfg1_sgl_p100"] = dict(gpu_s=t, cpu_s=tc, iters=len(info["residual"]), cpu_iters=ri["iterations"], rel_err_theta=rel(sol["Theta"], ref["Theta"]),
                            pattern_equal=bool(np.array_equal(sol["Theta"] != 0, ref["Theta"] != 0)))
# cfg2
S = synthetic_mgl(5, 100, N=1000, seed=1234); Om = np.repeat(np.eye(100)[None], 5, 0)
for lat in (False, True):
    t, (sol, info) = wall(lambda: ADMM_MGL(S, 0.05, 0.01, "GGL", Om, tol=1e-7, rtol=1e-7, latent=lat, mu1=0.1 if lat else None, measure=True), n=3)
    tc, (ref, ri) = wall(lambda: orc.admm_mgl(S, 0.05, 0.01, "GGL", Om, tol=1e-7, rtol=1e-7, latent=lat, mu1=0.1 if lat else None))
    out[f"cfg2_ggl_K5_p100_latent{int(lat)}"] = dict(gpu_s=t, cpu_s=tc, iters=len(info["residual"]), cpu_iters=ri["iterations"],
        rel_err_theta=rel(sol["Theta"], ref["Theta"]), pattern_equal=bool(np.array_equal(sol["Theta"] != 0, ref["Theta"] != 0)))
# cfg3: full solve to tolerance
S = synthetic_mgl(20, 1000, N=2000, seed=1234, kind="fused"); Om = np.repeat(np.eye(1000)[None], 20, 0)
t, (sol, info) = wall(lambda: ADMM_MGL(S, 0.05, 0.01, "FGL", Om, tol=1e-7, rtol=1e-7, measure=True), n=2)
d = dict(gpu_s=t, iters=len(info["residual"]), status=info["status"], objective=float(info["objective"][-1]), nnz_theta=int(np.count_nonzero(sol["Theta"])),
         gpu_loop_s=float(np.sum(info["runtime"])))
if not skip_cpu:
    tc, (ref, ri) = wall(lambda: orc.admm_mgl(S, 0.05, 0.01, "FGL", Om, tol=1e-7, rtol=1e-7, measure=True))
    d.update(cpu_s=tc, cpu_iters=ri["iterations"], cpu_objective=float(ri["objective"][-1]), rel_err_theta=rel(sol["Theta"], ref["Theta"]),
             rel_err_omega=rel(sol["Omega"], ref["Omega"]), pattern_equal=bool(np.array_equal(sol["Theta"] != 0, ref["Theta"] != 0)))
out["cfg3_fgl_K20_p1000"] = d
del sol
# cfg4: 10x10 grid, GGL K=10 p=500 (GPU: full grid; CPU: first lambda1 column only = 10 solves)
S = synthetic_mgl(10, 500, N=1000, seed=1234); N = np.full(10, 1000)
l1 = np.logspace(0, -3, 10); l2 = np.logspace(-1, -4, 10)
t, (scores, ix, best) = wall(lambda: grid_search_dist(ADMM_MGL, S, N, "GGL", l1, l2, gamma=0.1, tol=1e-7, rtol=1e-7))
d = dict(gpu_grid_s=t, best_ix=[int(ix[0]), int(ix[1])], best_lambda=[float(l1[ix[1]]), float(l2[ix[0]])], nan_scores=int(np.isnan(scores).sum()))
td, (sc_dev, it_dev, ix_dev, _) = wall(lambda: grid_search_device(S, N, "GGL", l1, l2, gamma=0.1, tol=1e-7, rtol=1e-7))
for ns in (2, 5):
    tn, (sc_n, it_n, ix_n, _) = wall(lambda: grid_search_device(S, N, "GGL", l1, l2, gamma=0.1, tol=1e-7, rtol=1e-7, n_streams=ns))
    d[f"gpu_grid_device_resident_{ns}streams_s"] = tn
    d[f"streams{ns}_scores_rel"] = float(np.nanmax(np.abs(sc_n - sc_dev) / np.abs(sc_dev)))
d.update(gpu_grid_device_resident_s=td, device_best_ix=[int(ix_dev[0]), int(ix_dev[1])], total_admm_iterations=int(it_dev.sum()),
         device_vs_host_scores_rel=float(np.nanmax(np.abs(sc_dev - scores) / np.abs(scores))))
if not skip_cpu:
    def cpu_solver(S, a, b, reg, Om0, tol=1e-7, rtol=1e-7, **kw): return orc.admm_mgl(S, a, b, reg, Om0, tol=tol, rtol=rtol)
    j = 4   # one representative column (lambda1 = l1[4])
    tc, (sc_c, ix_c, _) = wall(lambda: grid_search_dist(cpu_solver, S, N, "GGL", l1[j:j + 1], l2, gamma=0.1))
    tg, (sc_g, ix_g, _) = wall(lambda: grid_search_dist(ADMM_MGL, S, N, "GGL", l1[j:j + 1], l2, gamma=0.1))
    d.update(cpu_one_column_s=tc, gpu_one_column_s=tg, column_scores_rel_diff=float(np.nanmax(np.abs(sc_c - sc_g) / np.abs(sc_c))),
             column_best_equal=bool(tuple(ix_c) == tuple(ix_g)))
out["cfg4_grid10x10_ggl_K10_p500"] = d
# cfg5: block_SGL p=5000
S = synthetic_sgl(5000, N=5500, seed=1234, n_blocks=10); I = np.eye(5000)
for lam in (0.1, 0.07):
    numC, comps = get_connected_components(S, lam)
    sizes = sorted((len(c) for c in comps), reverse=True)
    t, sol = wall(lambda: block_SGL(S, lam, I, tol=1e-7, rtol=1e-7))
    d = dict(gpu_s=t, components=int(numC), non_singleton=int(sum(1 for s in sizes if s > 1)), largest=sizes[:3])
    if not skip_cpu:
        tc, ref = wall(lambda: orc.block_sgl(S, lam, I, tol=1e-7, rtol=1e-7))
        d.update(cpu_s=tc, rel_err_theta=rel(sol["Theta"], ref["Theta"]), pattern_equal=bool(np.array_equal(sol["Theta"] != 0, ref["Theta"] != 0)))
    out[f"cfg5_block_sgl_p5000_lam{lam}"] = d
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
print(json.dumps(out, indent=1))
