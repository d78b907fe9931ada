"""Edge cases of the reference-signature callables on the GPU path, checked against the CPU oracle."""
import contextlib
import io
import re

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]
TOL = 1e-8


def _quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        out = fn(*a, **kw)
    return out, buf.getvalue()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _cov(rng, p, N):
    X = rng.standard_normal((p, N))
    return np.atleast_2d(np.cov(X, bias=True))


@pytest.mark.parametrize("p", [1, 2, 3, 17])
def test_sgl_tiny_dimensions(p):
    from gglasso_b200 import ADMM_SGL
    from oracle import admm_oracle as orc
    rng = np.random.default_rng(p)
    S = _cov(rng, p, 40) + 0.1 * np.eye(p)
    (sol, info), _ = _quiet(ADMM_SGL, S, 0.1, np.eye(p), tol=1e-9, rtol=1e-9)
    ref, rinfo = orc.admm_sgl(S, 0.1, np.eye(p), tol=1e-9, rtol=1e-9)
    assert info["status"] == rinfo["status"]
    for k in ("Omega", "Theta", "X"):
        assert sol[k].shape == (p, p)
        assert np.abs(sol[k] - ref[k]).max() < 1e-9, k


@pytest.mark.parametrize("reg", ["GGL", "FGL"])
def test_mgl_single_instance_and_two_instances(reg):
    from gglasso_b200 import ADMM_MGL
    from oracle import admm_oracle as orc
    rng = np.random.default_rng(11)
    for K in (1, 2):
        p = 9
        S = np.stack([_cov(rng, p, 60) for _ in range(K)])
        Om0 = np.repeat(np.eye(p)[None], K, 0)
        (sol, info), _ = _quiet(ADMM_MGL, S, 0.08, 0.03, reg, Om0, tol=1e-8, rtol=1e-8)
        ref, rinfo = orc.admm_mgl(S, 0.08, 0.03, reg, Om0, tol=1e-8, rtol=1e-8)
        assert info["status"] == rinfo["status"]
        for k in ("Omega", "Theta", "X", "L"):
            assert np.abs(sol[k] - ref[k]).max() < 1e-9, (K, k)


def test_start_points_fixed_rho_and_nonunit_rho():
    """user supplied Omega_0 / Theta_0 / X_0 (warm start as model_selection does), rho != 1, update_rho=False."""
    from gglasso_b200 import ADMM_MGL
    from oracle import admm_oracle as orc
    rng = np.random.default_rng(5)
    K, p = 3, 25
    S = np.stack([_cov(rng, p, 100) for _ in range(K)])
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    first, _ = orc.admm_mgl(S, 0.1, 0.05, "GGL", Om0, tol=1e-4, rtol=1e-4)
    kw = dict(Theta_0=first["Theta"], X_0=first["X"], rho=3.0, update_rho=False, tol=1e-8, rtol=1e-8)
    (sol, info), _ = _quiet(ADMM_MGL, S, 0.07, 0.05, "GGL", first["Omega"], **kw)
    ref, rinfo = orc.admm_mgl(S, 0.07, 0.05, "GGL", first["Omega"], **kw)
    assert info["status"] == rinfo["status"]
    for k in ("Omega", "Theta", "X"):
        assert _rel(sol[k], ref[k]) < TOL, k
    # inputs are never mutated
    assert np.array_equal(Om0, np.repeat(np.eye(p)[None], K, 0))


def test_verbose_table_and_measure_keys():
    from gglasso_b200 import ADMM_MGL, ADMM_SGL
    rng = np.random.default_rng(1)
    p = 12
    S = _cov(rng, p, 80)
    (sol, info), out = _quiet(ADMM_SGL, S, 0.1, np.eye(p), verbose=True, measure=True, tol=1e-6, rtol=1e-6)
    lines = out.splitlines()
    assert lines[0] == "------------ADMM Algorithm for Single Graphical Lasso----------------"
    assert lines[1] == "%4s\t%10s\t%10s\t%10s\t%10s" % ("iter", "r_t", "s_t", "eps_pri", "eps_dual")
    n = len(info["residual"])
    rows = [l for l in lines[2:] if re.match(r"^\s*\d+\t", l)]
    assert len(rows) == n and rows[0].split("\t")[0].strip() == "0"
    assert lines[-1] == f"ADMM terminated after {n} iterations with status: {info['status']}."
    assert set(info) == {"status", "runtime", "residual"} and len(info["runtime"]) == n
    S3 = np.stack([S, S])
    (sol, info), out = _quiet(ADMM_MGL, S3, 0.1, 0.05, "FGL", np.repeat(np.eye(p)[None], 2, 0), verbose=True)
    assert out.splitlines()[0] == "------------ADMM Algorithm for Multiple Graphical Lasso----------------"
    assert set(info) == {"status"}


def test_latent_with_kkt_and_mu_array():
    from gglasso_b200 import ADMM_MGL
    from oracle import admm_oracle as orc
    rng = np.random.default_rng(2)
    K, p = 2, 14
    S = np.stack([_cov(rng, p, 90) for _ in range(K)])
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    mu = np.array([0.2, 0.35])
    kw = dict(latent=True, mu1=mu, stopping_criterion="kkt", tol=1e-6, max_iter=300, update_rho=False)
    (sol, info), _ = _quiet(ADMM_MGL, S, 0.1, 0.05, "GGL", Om0, measure=True, **kw)
    ref, rinfo = orc.admm_mgl(S, 0.1, 0.05, "GGL", Om0, measure=True, **kw)
    assert info["status"] == rinfo["status"] and len(info["residual"]) == rinfo["iterations"]
    for k in ("Omega", "Theta", "X", "L"):
        assert _rel(sol[k], ref[k]) < 1e-7, k


def test_nonconvergence_status_strings():
    from gglasso_b200 import ADMM_SGL
    from oracle import admm_oracle as orc
    rng = np.random.default_rng(8)
    p = 20
    S = _cov(rng, p, 30)
    for kw in (dict(max_iter=3, tol=1e-12, rtol=1e-12), dict(max_iter=5, tol=1e-3, rtol=1e-14)):
        (sol, info), out = _quiet(ADMM_SGL, S, 0.05, np.eye(p), **kw)
        ref, rinfo = orc.admm_sgl(S, 0.05, np.eye(p), **kw)
        assert info["status"] == rinfo["status"]
        assert info["status"] in ("max iterations reached", "primal optimal", "dual optimal", "optimal")
        assert f"after {rinfo['iterations']} iterations" in out
        assert _rel(sol["Theta"], ref["Theta"]) < TOL


def test_non_finite_input_raises_instead_of_iterating():
    """numpy's eigh raises LinAlgError on a non-finite matrix; here the device stopping test flags the non-finite
    residual (status -1) and the host raises."""
    from gglasso_b200 import ADMM_SGL, ADMM_MGL, GGLassoB200Error
    rng = np.random.default_rng(0)
    p = 20
    S = np.cov(rng.standard_normal((p, 100)), bias=True)
    S[3, 3] = np.nan
    with pytest.raises(GGLassoB200Error, match="non-finite"):
        _quiet(ADMM_SGL, S, 0.1, np.eye(p), max_iter=50)
    S3 = np.stack([np.cov(rng.standard_normal((p, 100)), bias=True) for _ in range(2)])
    S3[1, 0, 1] = S3[1, 1, 0] = np.inf
    with pytest.raises((GGLassoB200Error, AssertionError)):
        _quiet(ADMM_MGL, S3, 0.1, 0.1, "GGL", np.repeat(np.eye(p)[None], 2, 0), max_iter=50)


def test_upload_from_pinned_and_pageable_host_arrays_agree():
    """to_dev: page-locked caller arrays take the single-DMA path, pageable ones the threaded staging path"""
    import torch
    from gglasso_b200._engine import to_dev, require_cuda
    dev = require_cuda()
    rng = np.random.default_rng(5)
    a = rng.standard_normal((3, 700, 700))                       # 11.8 MB: above the staging threshold
    h = torch.empty(a.shape, dtype=torch.float64, pin_memory=True).numpy()
    h[...] = a
    assert torch.from_numpy(h).is_pinned() and not torch.from_numpy(a).is_pinned()
    d1, d2 = to_dev(a, dev), to_dev(h, dev)
    torch.cuda.synchronize()
    assert torch.equal(d1, d2) and np.array_equal(d1.cpu().numpy(), a)
    hp = torch.empty((60, 60), dtype=torch.float64, pin_memory=True).numpy()        # small arrays: plain path
    hp[...] = np.eye(60)
    assert np.array_equal(to_dev(hp, dev).cpu().numpy(), np.eye(60))
