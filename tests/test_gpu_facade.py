"""The drop-in claim, by test: the UNMODIFIED reference façade (glasso_problem.solve / model_selection,
src/gglasso/problem.py:371,551; grid_search / single_grid_search / K_single_grid,
src/gglasso/helper/model_selection.py:55,300,505) runs on top of the B200 solvers after ``gglasso_b200.install()``
and gives what the same façade gives on its own CPU solvers (oracle/_ref, run live on this host) -- the shapes are the
reference's tests/test_problem.py:31-138 (p=20, K=3, N=1000, 4x3 grid), plus a component-rich SGL case that takes
the block_SGL route of single_grid_search.
"""
import contextlib
import io
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from oracle import ref  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device"),
              pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (oracle/make_ref.sh)")]

p, K, N, M = 20, 3, 1000, 4


def _quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **kw)


def _both(run):
    """run(glasso_problem) once on the reference's CPU solvers and once with the B200 solvers installed"""
    import gglasso_b200
    ref.load()
    from gglasso.problem import glasso_problem
    gglasso_b200.uninstall()
    cpu = _quiet(run, glasso_problem)
    patched = gglasso_b200.install()
    assert "gglasso.problem.ADMM_MGL" in patched and "gglasso.helper.model_selection.block_SGL" in patched
    try:
        gpu = _quiet(run, glasso_problem)
    finally:
        gglasso_b200.uninstall()
    return cpu, gpu


def _gen(name, **kw):
    ref.load()
    from gglasso.helper import data_generation as dg
    Sigma, _ = getattr(dg, name)(**kw)
    S, _ = dg.sample_covariance_matrix(Sigma, N, seed=kw["seed"])
    return S


def _same_stats(a, b, rtol=1e-7):
    for key in ("AIC", "SP"):
        np.testing.assert_allclose(a[key], b[key], rtol=rtol, err_msg=key)
    for g in a["BIC"]:
        np.testing.assert_allclose(a["BIC"][g], b["BIC"][g], rtol=rtol, err_msg=f"BIC[{g}]")
    assert a["BEST"] == b["BEST"], (a["BEST"], b["BEST"])


@pytest.mark.parametrize("reg,latent", [("GGL", False), ("FGL", False), ("GGL", True), ("FGL", True)])
def test_mgl_problem_solve_and_model_selection(reg, latent):
    gen = "group_power_network" if reg == "GGL" else "time_varying_power_network"
    S = _gen(gen, p=p, K=K, M=M, seed=123)

    def run(glasso_problem):
        P = glasso_problem(S=S, N=N, reg=reg, latent=latent)
        rp = {"lambda1": 0.01, "lambda2": 0.001}
        if latent:
            rp["mu1"] = 1.0
        P.set_reg_params(rp)
        P.solve(verbose=True)
        first = {"Theta": P.solution.precision_.copy(), "L": P.solution.lowrank_.copy() if latent else None}
        ms = {"lambda1_range": np.logspace(0, -3, 4), "lambda2_range": np.logspace(-1, -3, 3),
              "mu1_range": np.logspace(-2, 0, 4) if latent else None}
        P.model_selection(modelselect_params=ms, method="eBIC", gamma=0.1)
        return first, P.modelselect_stats, P.solution.precision_.copy(), P.reg_params, P.solution.calc_ebic(gamma=0.1)

    (f_c, st_c, Th_c, rp_c, eb_c), (f_g, st_g, Th_g, rp_g, eb_g) = _both(run)
    assert np.linalg.norm(f_g["Theta"] - f_c["Theta"]) <= 1e-6 * np.linalg.norm(f_c["Theta"])
    assert np.array_equal(f_g["Theta"] != 0, f_c["Theta"] != 0)
    if latent:
        assert np.linalg.norm(f_g["L"] - f_c["L"]) <= 1e-6 * max(np.linalg.norm(f_c["L"]), 1.0)
        assert np.array_equal(st_g["RANK"], st_c["RANK"])
    _same_stats(st_g, st_c, rtol=1e-6)
    assert rp_c["lambda1"] == rp_g["lambda1"] and rp_c["lambda2"] == rp_g["lambda2"]
    if latent:
        assert np.array_equal(rp_c["mu1"], rp_g["mu1"])
    assert np.array_equal(Th_g != 0, Th_c != 0)
    assert np.linalg.norm(Th_g - Th_c) <= 1e-6 * np.linalg.norm(Th_c)
    assert abs(eb_g - eb_c) <= 1e-7 * abs(eb_c)


@pytest.mark.parametrize("latent,seed", [(False, 1234), (True, 2345)])
def test_sgl_problem_solve_and_model_selection(latent, seed):
    S = _gen("generate_precision_matrix", p=p, M=2, style="powerlaw", gamma=2.8, prob=0.1, seed=seed)

    def run(glasso_problem):
        P = glasso_problem(S=S, N=N, reg=None, latent=latent)
        rp = {"lambda1": 0.01}
        if latent:
            rp["mu1"] = 1.0
        P.set_reg_params(rp)
        P.solve()
        first = P.solution.precision_.copy()
        P.model_selection(modelselect_params=None, method="eBIC", gamma=0.1)
        return first, P.modelselect_stats, P.solution.precision_.copy(), P.reg_params

    (f_c, st_c, Th_c, rp_c), (f_g, st_g, Th_g, rp_g) = _both(run)
    assert np.linalg.norm(f_g - f_c) <= 1e-6 * np.linalg.norm(f_c) and np.array_equal(f_g != 0, f_c != 0)
    np.testing.assert_allclose(st_g["BIC"][0.1], st_c["BIC"][0.1], rtol=1e-6)
    np.testing.assert_allclose(st_g["SP"], st_c["SP"], rtol=1e-9)
    assert rp_c["lambda1"] == rp_g["lambda1"]
    assert np.linalg.norm(Th_g - Th_c) <= 1e-6 * np.linalg.norm(Th_c) and np.array_equal(Th_g != 0, Th_c != 0)


def test_sgl_model_selection_takes_the_block_route():
    """p=200 with 10 blocks: single_grid_search(use_block=True) calls block_SGL for the larger lambda1 values."""
    ref.load()
    from gglasso.helper import data_generation as dg
    Sigma, _ = dg.generate_precision_matrix(p=200, M=10, style="powerlaw", gamma=2.8, seed=77)
    S, _ = dg.sample_covariance_matrix(Sigma, 400, seed=77)

    def run(glasso_problem):
        P = glasso_problem(S=S, N=400, reg=None, latent=False)
        P.model_selection(modelselect_params={"lambda1_range": np.logspace(-0.3, -1.5, 5)}, method="eBIC", gamma=0.1)
        return P.modelselect_stats, P.solution.precision_.copy()

    (st_c, Th_c), (st_g, Th_g) = _both(run)
    np.testing.assert_allclose(st_g["BIC"][0.1], st_c["BIC"][0.1], rtol=1e-6)
    assert st_c["BEST"] == st_g["BEST"]
    assert np.array_equal(Th_g != 0, Th_c != 0) and np.linalg.norm(Th_g - Th_c) <= 1e-6 * np.linalg.norm(Th_c)


def test_device_scoring_matches_reference_helpers():
    """gglasso_b200.scoring (eBIC / AIC incl. the lambda1_mask variant, robust_logdet's -inf rule, mean_sparsity,
    matrix_rank, tune_threshold) against the reference's helpers (model_selection.py:697-894, utils.py:17-31)."""
    ref.load()
    from gglasso.helper import model_selection as ms
    from gglasso.helper.utils import mean_sparsity
    from gglasso_b200 import scoring
    rng = np.random.default_rng(5)
    K, pp, Ns = 3, 40, np.array([200, 300, 400])
    A = rng.standard_normal((K, pp, 3 * pp))
    S = A @ A.transpose(0, 2, 1) / (3 * pp)
    Theta = np.linalg.inv(S + 0.5 * np.eye(pp))
    Theta[np.abs(Theta) < 0.05] = 0.0
    Theta = (Theta + Theta.transpose(0, 2, 1)) / 2
    Sd, Td = torch.from_numpy(S).cuda(), torch.from_numpy(Theta).cuda()
    for g in (0.0, 0.1, 0.7):
        assert abs(scoring.ebic(Sd, Td, Ns, g) - ms.ebic(S, Theta, Ns, g)) <= 1e-9 * abs(ms.ebic(S, Theta, Ns, g))
    assert abs(scoring.aic(Sd, Td, Ns) - ms.aic(S, Theta, Ns)) <= 1e-9 * abs(ms.aic(S, Theta, Ns))
    mask = 0.5 + 0.5 * rng.random((pp, pp))
    mask = (mask + mask.T) / 2
    want = ms.ebic_single(S[0], Theta[0], 200, 0.3, lambda1_mask=mask)
    assert abs(scoring.ebic(Sd[0], Td[0], 200, 0.3, lambda1_mask=mask) - want) <= 1e-9 * abs(want)
    assert abs(scoring.mean_sparsity(Td) - mean_sparsity(Theta)) < 1e-15
    # not positive definite -> -inf log det -> +inf score, as in robust_logdet
    bad = Theta.copy()
    bad[1] -= 10 * np.eye(pp)
    assert scoring.ebic(Sd, torch.from_numpy(bad).cuda(), Ns, 0.1) == np.inf == ms.ebic(S, bad, Ns, 0.1)
    # rank of low-rank PSD matrices
    U = rng.standard_normal((K, pp, 4))
    L = U @ U.transpose(0, 2, 1)
    L[2] = 0.0
    assert np.array_equal(scoring.matrix_rank(torch.from_numpy(L).cuda()), [np.linalg.matrix_rank(L[k]) for k in range(K)])
    # threshold tuning
    Tt, tau, sc = scoring.tune_threshold(Td[0], Sd[0], 200, method="eBIC", gamma=0.1)
    Tr, taur, scr = ms.tune_threshold(Theta[0], S[0], 200, method="eBIC", gamma=0.1)
    assert tau == taur and np.array_equal(Tt.cpu().numpy() != 0, Tr != 0)
    np.testing.assert_allclose(sc, scr, rtol=1e-9)
    Tm, taus, _ = scoring.tune_multiple_threshold(Td, Sd, Ns, None, method="AIC")
    Tmr, tausr, _ = ms.tune_multiple_threshold(Theta, S, Ns, None, method="AIC")
    assert np.array_equal(taus, tausr) and np.array_equal(Tm.cpu().numpy(), Tmr)


@pytest.mark.parametrize("latent", [False, True])
def test_single_grid_search_on_device_matches_reference_driver(latent):
    """scoring.single_grid_search_device (S, warm starts, scoring on the GPU) against the reference's
    single_grid_search(use_block=False) on its CPU solver (model_selection.py:505-690)."""
    ref.load()
    import gglasso_b200
    gglasso_b200.uninstall()
    from gglasso.helper import data_generation as dg
    from gglasso.helper.model_selection import single_grid_search
    from gglasso_b200.scoring import single_grid_search_device
    Sigma, _ = dg.generate_precision_matrix(p=40, M=2, style="powerlaw", gamma=2.8, prob=0.1, seed=99)
    S, _ = dg.sample_covariance_matrix(Sigma, 500, seed=99)
    lam = np.logspace(-0.5, -2, 5)
    mu = np.logspace(0, -1, 3) if latent else None
    kw = dict(method="eBIC", gamma=0.3, latent=latent, mu_range=mu, store_all=True, tol=1e-8, rtol=1e-8)
    bs_r, est_r, low_r, st_r = _quiet(single_grid_search, S, lam, 500, use_block=False, **kw)
    bs_g, est_g, low_g, st_g = _quiet(single_grid_search_device, S, lam, 500, **kw)
    assert st_r["BEST"] == st_g["BEST"]
    for g in st_r["BIC"]:
        np.testing.assert_allclose(st_g["BIC"][g], st_r["BIC"][g], rtol=1e-7)
    np.testing.assert_allclose(st_g["AIC"], st_r["AIC"], rtol=1e-7)
    assert np.array_equal(st_g["SP"], st_r["SP"]) and np.array_equal(st_g["RANK"], st_r["RANK"])
    assert np.abs(est_g - est_r).max() < 1e-6 and np.array_equal(est_g != 0, est_r != 0)
    if latent:
        assert np.abs(low_g - low_r).max() < 1e-6
    for k in bs_r:
        assert np.abs(bs_g[k] - bs_r[k]).max() < 1e-6, k
