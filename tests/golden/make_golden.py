"""
Generate the golden fixtures under tests/golden/ by running the REAL reference
(fabian-sp/GGLasso at /root/reference, imported through _refshim.py) in the build container.

    python tests/golden/make_golden.py

The fixtures are committed; this script cannot run on the GPU box (no /root/reference there).
Per-iteration scalars are captured non-invasively by wrapping the module-global
ADMM_stopping_criterion that the reference solvers look up every iteration
(src/gglasso/solver/admm_solver.py:217, single_admm_solver.py:186).
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refshim import load_reference  # noqa: E402

load_reference()
import gglasso.solver.admm_solver as ref_mgl  # noqa: E402
import gglasso.solver.single_admm_solver as ref_sgl  # noqa: E402
from gglasso.helper.data_generation import (generate_precision_matrix, group_power_network,  # noqa: E402
                                            sample_covariance_matrix, time_varying_power_network)
from gglasso.solver import ggl_helper as gh  # noqa: E402
from gglasso.solver.fgl_helper import condat_method  # noqa: E402


class Capture:
    """wraps <module>.ADMM_stopping_criterion and records the per-iteration scalars."""

    def __init__(self, mod):
        self.mod = mod
        self.rows = []

    def __enter__(self):
        self.orig = self.mod.ADMM_stopping_criterion

        def wrapped(Omega, Omega_t_1, Theta, L, X, S, rho, eps_abs, eps_rel, latent=False):
            out = self.orig(Omega, Omega_t_1, Theta, L, X, S, rho, eps_abs, eps_rel, latent)
            self.rows.append([rho, *out, np.linalg.norm(Omega), np.linalg.norm(Theta), np.linalg.norm(L),
                              np.linalg.norm(X), float(np.count_nonzero(Theta))])
            return out

        self.mod.ADMM_stopping_criterion = wrapped
        return self

    def __exit__(self, *a):
        self.mod.ADMM_stopping_criterion = self.orig

    def table(self):
        # columns: rho r s e_pri e_dual |Omega| |Theta| |L| |X| nnz(Theta)
        return np.array(self.rows)


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def run_mgl(name, S, l1, l2, reg, full=True, **kw):
    K, p, _ = S.shape
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    with Capture(ref_mgl) as cap:
        sol, info = quiet(ref_mgl.ADMM_MGL, S, l1, l2, reg, Om0, measure=True, **kw)
    out = dict(S=S, lambda1=l1, lambda2=l2, reg=reg, traj=cap.table(), status=info["status"],
               objective=info["objective"], residual=info["residual"], Theta=sol["Theta"],
               kw=repr(sorted(kw.items())))
    if full:
        out.update(Omega=sol["Omega"], X=sol["X"], L=sol["L"])
    save(name, **out)


def run_sgl(name, S, l1, **kw):
    p = S.shape[0]
    with Capture(ref_sgl) as cap:
        sol, info = quiet(ref_sgl.ADMM_SGL, S, l1, np.eye(p), measure=True, **kw)
    out = dict(S=S, lambda1=l1, traj=cap.table(), status=info["status"], residual=info["residual"],
               Theta=sol["Theta"], Omega=sol["Omega"], X=sol["X"], kw=repr(sorted(kw.items())))
    if "L" in sol:
        out["L"] = sol["L"]
    if "lambda1_mask" in kw:
        out["lambda1_mask"] = kw["lambda1_mask"]
    save(name, **out)


def main():
    rng = np.random.default_rng(20240917)

    # ---- unit-level prox fixtures ---------------------------------------------------------
    K, p = 7, 12
    X = rng.standard_normal((K, p, p)) * 0.3
    X = X + X.transpose(0, 2, 1)
    ys = rng.standard_normal((40, 9))
    ys[10:20] = np.round(ys[10:20], 1)          # ties / plateaus
    ys[20:25] *= 1e-3
    lams = np.abs(rng.standard_normal(40)) * 0.5 + 1e-3
    A = rng.standard_normal((p, p))
    A = A + A.T
    D, Q = np.linalg.eigh(A)
    mask = np.abs(rng.standard_normal((p, p)))
    mask = mask + mask.T
    save("prox_units",
         X=X, l1=0.11, l2=0.23,
         prox_p_ggl=gh.prox_p(X, 0.11, 0.23, "GGL"), prox_p_fgl=gh.prox_p(X, 0.11, 0.23, "FGL"),
         pval_ggl=gh.P_val(X, 0.11, 0.23, "GGL"), pval_fgl=gh.P_val(X, 0.11, 0.23, "FGL"),
         tv_y=ys, tv_lam=lams, tv_x=np.stack([condat_method(y, l) for y, l in zip(ys, lams)]),
         A=A, D=D, Q=Q, od1_scalar=gh.prox_od_1norm(A, 0.4), mask=mask, od1_mask=gh.prox_od_1norm(A, 0.4 * mask),
         phiplus=gh.phiplus(0.7, D, Q), rank_norm=gh.prox_rank_norm(A, 0.9, D=D, Q=Q))

    # ---- reference-test sized MGL problems (tests/test_solvers.py:25-65: p=50,K=3 -> here M=5) ----
    Sigma, _ = group_power_network(50, 3, 5, seed=1234)
    S, _ = sample_covariance_matrix(Sigma, 1000, seed=1234)
    for reg in ("GGL", "FGL"):
        run_mgl(f"mgl_{reg.lower()}_K3_p50", S, 0.05, 0.01, reg, tol=1e-7, rtol=1e-7)
        run_mgl(f"mgl_{reg.lower()}_latent_K3_p50", S, 0.05, 0.01, reg, tol=1e-7, rtol=1e-7, latent=True, mu1=0.1)
    run_mgl("mgl_ggl_K3_p50_maxiter2", S, 0.05, 0.01, "GGL", tol=1e-7, rtol=1e-7, max_iter=2)
    run_mgl("mgl_ggl_K3_p50_fixedrho", S, 0.05, 0.01, "GGL", tol=1e-6, rtol=1e-6, update_rho=False, rho=2.0)
    run_mgl("mgl_fgl_K3_p50_kkt", S, 0.05, 0.01, "FGL", tol=1e-5, stopping_criterion="kkt", update_rho=False)
    run_mgl("mgl_ggl_K3_p50_nsamples", S, 0.05, 0.01, "GGL", tol=1e-7, rtol=1e-7, n_samples=3)

    # ---- cfg2 of BASELINE.json: K=5, p=100 (Theta + trajectories only, to keep fixtures small) ----
    Sigma, _ = group_power_network(100, 5, 10, seed=1234)
    S, _ = sample_covariance_matrix(Sigma, 1000, seed=1234)
    run_mgl("cfg2_ggl", S, 0.05, 0.01, "GGL", full=False, tol=1e-7, rtol=1e-7)
    run_mgl("cfg2_ggl_latent", S, 0.05, 0.01, "GGL", full=False, tol=1e-7, rtol=1e-7, latent=True, mu1=0.1)
    Sigma, _ = time_varying_power_network(100, 5, 10, seed=1234)
    S, _ = sample_covariance_matrix(Sigma, 1000, seed=1234)
    run_mgl("cfg2_fgl_tv", S, 0.05, 0.01, "FGL", full=False, tol=1e-7, rtol=1e-7)

    # ---- cfg1 of BASELINE.json: SGL p=100 ----
    Sigma, _ = generate_precision_matrix(p=100, M=10, style="powerlaw", gamma=2.8, seed=1234)
    S, _ = sample_covariance_matrix(Sigma, 1000, seed=1234)
    run_sgl("cfg1_sgl", S, 0.05, tol=1e-7, rtol=1e-7)
    run_sgl("cfg1_sgl_latent", S, 0.05, tol=1e-7, rtol=1e-7, latent=True, mu1=0.1)
    m = np.ones((100, 100))
    m[:50, 50:] = 0.0
    m[50:, :50] = 0.0
    m[:20, :20] = 2.5
    run_sgl("cfg1_sgl_mask", S, 0.05, tol=1e-7, rtol=1e-7, lambda1_mask=m)
    run_sgl("cfg1_sgl_kkt", S, 0.05, tol=1e-6, stopping_criterion="kkt", max_iter=200)

    # ---- block_SGL (tests/test_solvers.py:123-148 style input: S = A^T A + 90 I, scaled) ----
    np.random.seed(1234)
    p = 100
    Aa = np.random.randn(p, p)
    Sb = Aa.T @ Aa + 90 * np.eye(p)
    Sb = Sb / np.sqrt(np.outer(np.diag(Sb), np.diag(Sb)))
    lam = 0.12
    sol = quiet(ref_sgl.block_SGL, Sb, lam, np.eye(p), tol=1e-9, rtol=1e-9)
    numC, _ = ref_sgl.get_connected_components(Sb, lam)
    full, _ = quiet(ref_sgl.ADMM_SGL, Sb, lam, np.eye(p), tol=1e-9, rtol=1e-9)
    save("block_sgl_p100", S=Sb, lambda1=lam, numC=numC, Theta=sol["Theta"], Omega=sol["Omega"], X=sol["X"],
         Theta_full=full["Theta"])


def fsgl_fixtures():
    """functional SGL (src/gglasso/solver/functional_sgl_admm.py): M=3 blocks, with and without latent, + prox."""
    import gglasso.solver.functional_sgl_admm as ref_f
    rng = np.random.default_rng(77)
    p, M, N = 12, 3, 400
    pM = p * M
    A = rng.standard_normal((pM, pM)) * (rng.random((pM, pM)) < 0.08)
    Prec = A @ A.T * 0.3 + np.eye(pM)
    Sigma = np.linalg.inv(Prec)
    X = rng.multivariate_normal(np.zeros(pM), Sigma, N).T
    S = np.cov(X, bias=True)
    Y = rng.standard_normal((pM, pM))
    Y = Y + Y.T
    out = dict(S=S, M=M, lambda1=0.05, prox_in=Y, prox_out=gh.prox_sum_Frob(Y, M, 0.7))
    for lat in (False, True):
        with Capture(ref_f) as cap:
            sol, info = quiet(ref_f.ADMM_FSGL, S, 0.05, M, np.eye(pM), tol=1e-7, rtol=1e-7, measure=True,
                              latent=lat, mu1=0.2 if lat else None)
        tag = "lat" if lat else "nolat"
        out.update({f"Theta_{tag}": sol["Theta"], f"Omega_{tag}": sol["Omega"], f"X_{tag}": sol["X"],
                    f"traj_{tag}": cap.table(), f"status_{tag}": info["status"], f"residual_{tag}": info["residual"]})
        if lat:
            out["L_lat"] = sol["L"]
    save("fsgl_p12_M3", **out)


def ext_fixtures():
    """non-conforming GGL (src/gglasso/solver/ext_admm_solver.py): three instances sharing part of their variables;
    the bookkeeping array G is built with the reference's own helpers (helper/ext_admm_helper.py)."""
    import pandas as pd
    import gglasso.solver.ext_admm_solver as ref_e
    from gglasso.helper.ext_admm_helper import construct_indexer, create_group_array, check_G, construct_trivial_G
    rng = np.random.default_rng(2024)
    all_vars = np.arange(18)
    subsets = [np.arange(0, 12), np.arange(4, 18), np.r_[0:6, 10:16]]
    N = 300
    Aall = rng.standard_normal((18, 18)) * (rng.random((18, 18)) < 0.15)
    Prec = Aall @ Aall.T * 0.4 + np.eye(18)
    Sig = np.linalg.inv(Prec)
    samples, S = [], {}
    for k, idx in enumerate(subsets):
        X = rng.multivariate_normal(np.zeros(len(idx)), Sig[np.ix_(idx, idx)], N).T
        samples.append(pd.DataFrame(X, index=idx))
        S[k] = np.cov(X, bias=True)
    ix_exist, ix_location = construct_indexer(samples)
    G = quiet(create_group_array, ix_exist, ix_location)
    p = np.array([len(s) for s in subsets])
    check_G(G, p)
    Om0 = {k: np.eye(p[k]) for k in range(3)}
    out = dict(G=G, p=p, lambda1=0.05, lambda2=0.08)
    for k in range(3):
        out[f"S{k}"] = S[k]
    for tag, kw in (("boyd", dict(tol=1e-7, rtol=1e-7)), ("latent", dict(tol=1e-7, rtol=1e-7, latent=True, mu1=0.3)),
                    ("kkt", dict(tol=1e-4, stopping_criterion="kkt", max_iter=400))):
        sol, info = quiet(ref_e.ext_ADMM_MGL, S, 0.05, 0.08, "GGL", Om0, G, measure=True, **kw)
        out[f"status_{tag}"] = info["status"]
        out[f"residual_{tag}"] = info["residual"]
        for name in ("Omega", "Theta", "L", "X0", "X1"):
            for k in range(3):
                out[f"{name}{k}_{tag}"] = sol[name][k]
    # trivial G (conforming) for the consistency check with ADMM_MGL (tests/test_solvers.py:71-120)
    out["G_trivial"] = construct_trivial_G(12, 3)
    save("ext_mgl_K3", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "fsgl":
        fsgl_fixtures()
    elif len(sys.argv) > 1 and sys.argv[1] == "ext":
        ext_fixtures()
    else:
        main()
        fsgl_fixtures()
        ext_fixtures()
