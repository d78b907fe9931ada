"""GPU parity tests (run on the B200 box: pytest -m gpu).

Every test drives the CUDA path through the C ABI (ctypes) or through the reference-signature
callables and checks it against (a) golden fixtures produced by the real reference and (b) the
CPU oracle on the same seeded input.  Tolerances follow BASELINE.json's north_star:
Theta/Omega/L within 1e-8 relative Frobenius per iteration, identical sparsity pattern,
final objective within 1e-6 relative.
"""
import ast
import contextlib
import ctypes
import io
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

PER_ITER_TOL = 1e-8


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        out = fn(*a, **kw)
    return out, buf.getvalue()


def _sym(rng, p, scale=1.0):
    A = rng.standard_normal((p, p)) * scale
    return (A + A.T) / 2


# --------------------------------------------------------------------------------------------
# eigensolver + spectral reconstruction
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p,nb2", [(161, 32), (200, 64), (333, 32), (512, 128), (100, 1), (160, 1), (49, 1)])
def test_eigh_block_jacobi_path(p, nb2):
    """non-default eigensolver paths: block Jacobi (nb2 = 32/64/128) and the forced shared-memory Jacobi (nb2 = 1)"""
    from gglasso_b200._engine import eigh
    rng = np.random.default_rng(p)
    A = np.stack([_sym(rng, p) for _ in range(2)])
    D, Q = eigh(A, nb2=nb2)
    Dref = np.linalg.eigvalsh(A)
    assert np.abs(D - Dref).max() < 1e-11 * p ** 0.5 * (np.abs(Dref).max() + 1)
    for m in range(2):
        assert np.abs(Q[m].T @ Q[m] - np.eye(p)).max() < 1e-12
        assert np.abs(A[m] @ Q[m] - Q[m] * D[m]).max() < 2e-12 * p ** 0.5 * (np.abs(Dref).max() + 1)


@pytest.mark.parametrize("p", [1, 2, 3, 10, 37, 100, 160, 161, 200, 333, 512, 777, 1000])
def test_eigh_matches_lapack(p):
    from gglasso_b200._engine import eigh
    rng = np.random.default_rng(p)
    M = 3 if p > 2 else 2
    A = np.stack([_sym(rng, p) - np.cov(rng.standard_normal((p, 2 * p + 2)), bias=True).reshape(p, p)
                  for _ in range(M)])
    A[0] = np.diag(np.arange(p, dtype=float))            # already diagonal
    if p > 4:
        A[1][:, :2] = 0.0; A[1][:2, :] = 0.0            # noqa: E702  (exact zero rows -> rank deficient)
    D, Q = eigh(A)
    Dref = np.linalg.eigvalsh(A)
    scale = np.abs(Dref).max(axis=1, keepdims=True) + 1.0
    assert np.abs(D - Dref).max() < 1e-11 * p ** 0.5 * scale.max()
    for m in range(M):
        assert np.abs(Q[m].T @ Q[m] - np.eye(p)).max() < 1e-12
        assert np.abs(A[m] @ Q[m] - Q[m] * D[m]).max() < 2e-12 * p ** 0.5 * scale[m, 0]


def test_eigh_clustered_and_scaled():
    from gglasso_b200._engine import eigh
    rng = np.random.default_rng(7)
    for p in (64, 300):
        Q, _ = np.linalg.qr(rng.standard_normal((p, p)))
        d = np.concatenate([np.full(p // 2, 1.0), np.full(p // 4, 1.0 + 1e-9), rng.standard_normal(p - p // 2 - p // 4)])
        for s in (1e-6, 1.0, 1e5):
            A = (Q * (s * d)) @ Q.T
            A = (A + A.T) / 2
            D, V = eigh(A)
            assert np.abs(np.sort(D) - np.sort(s * d)).max() < 1e-11 * s * p ** 0.5
            f = np.exp(-D / s)
            ref = (Q * np.exp(-d)) @ Q.T
            assert _rel((V * f) @ V.T, ref) < 1e-10


@pytest.mark.parametrize("M,p", [(3, 300), (1, 777)])
def test_eigh_merged_chain_option(M, p, monkeypatch):
    """GG_TR_MERGE=1: one launch per column (the CTA finishing a matrix's last tile runs the next column step)"""
    from gglasso_b200._engine import eigh
    monkeypatch.setenv("GG_TR_MERGE", "1")
    rng = np.random.default_rng(p)
    A = np.stack([_sym(rng, p) for _ in range(M)])
    D, Q = eigh(A)
    assert np.abs(D - np.linalg.eigvalsh(A)).max() < 1e-11
    for m in range(M):
        assert np.abs(Q[m].T @ Q[m] - np.eye(p)).max() < 1e-12
        assert np.abs(A[m] @ Q[m] - Q[m] * D[m]).max() < 1e-11


@pytest.mark.parametrize("p", [5, 64, 100, 161, 257, 384])
def test_recon_modes(p):
    from gglasso_b200 import _lib
    from gglasso_b200._engine import Eigh, to_dev, _p
    lib = _lib.load()
    dev = torch.device("cuda")
    rng = np.random.default_rng(p + 1)
    M = 2
    A = np.stack([_sym(rng, p) for _ in range(M)])
    D, Q = np.linalg.eigh(A)
    At = to_dev(A, dev)
    e = Eigh(M, p, dev)
    e.eigh(At)
    out = torch.empty_like(At)
    bnum = to_dev(np.array([0.7, 1.9]), dev)
    for mode, f in ((0, lambda d, b: 0.5 * (np.sqrt(d * d + 4 * b) + d)), (1, lambda d, b: np.maximum(d - b, 0)),
                    (2, lambda d, b: d)):
        e.recon(At, out, mode, bnum=bnum)
        got = out.cpu().numpy()
        for m, b in enumerate((0.7, 1.9)):
            ref = (Q[m] * f(D[m], b)) @ Q[m].T
            assert _rel(got[m], ref) < 1e-12, (mode, m)
            assert np.array_equal(got[m], got[m].T), "reconstruction must be exactly symmetric"


# --------------------------------------------------------------------------------------------
# prox kernels through the C ABI vs golden / oracle
# --------------------------------------------------------------------------------------------
def _ctrl(dev, rho=1.0, n=1):
    from gglasso_b200._lib import CTRL_STRIDE
    c = np.zeros((n, CTRL_STRIDE))
    c[:, 0] = rho
    c[:, 1] = 1.0
    return torch.from_numpy(c).to(dev)


@pytest.mark.parametrize("reg", ["GGL", "FGL"])
def test_prox_mgl_kernel_vs_reference_fixture(golden, reg):
    from gglasso_b200 import _lib
    from gglasso_b200._engine import to_dev, _p
    lib = _lib.load()
    dev = torch.device("cuda")
    g = golden("prox_units")
    Xin = g["X"]
    K, p, _ = Xin.shape
    want = g["prox_p_ggl" if reg == "GGL" else "prox_p_fgl"]
    # latent-mode call: Theta = prox(Omega + L + X) with Omega = Xin, L = X = 0
    Om, Z = to_dev(Xin, dev), torch.zeros((K, p, p), dtype=torch.float64, device=dev)
    Th, C = torch.empty_like(Om), torch.empty_like(Om)
    rc = lib.gg_prox_mgl(_p(Om), _p(Om), _p(Z), _p(Z), _p(Th), _p(C), _p(_ctrl(dev)), float(g["l1"]), float(g["l2"]),
                         0 if reg == "GGL" else 1, K, p, None, 0)
    assert rc == 0
    got = Th.cpu().numpy()
    if reg == "FGL":
        assert np.array_equal(got, want)           # TV scan + soft threshold: bit exact
    else:
        assert np.abs(got - want).max() < 1e-15    # dnrm2 vs sqrt(sum) can differ in the last bit
    assert np.array_equal(C.cpu().numpy(), got - 0.0 - Xin)


@pytest.mark.parametrize("reg,K,p", [("GGL", 2, 16), ("FGL", 5, 33), ("GGL", 20, 70), ("FGL", 20, 70), ("FGL", 31, 17)])
def test_prox_mgl_fused_dual_and_norms(reg, K, p):
    from gglasso_b200 import _lib
    from gglasso_b200._engine import to_dev, _p
    from oracle import admm_oracle as orc
    lib = _lib.load()
    dev = torch.device("cuda")
    rng = np.random.default_rng(K * 100 + p)
    Om = np.stack([_sym(rng, p, 0.3) for _ in range(K)])
    Omp = np.stack([_sym(rng, p, 0.3) for _ in range(K)])
    X = np.stack([_sym(rng, p, 0.2) for _ in range(K)])
    rho, l1, l2 = 2.0, 0.21, 0.13
    want = orc.prox_p(Om + X, l1 / rho, l2 / rho, reg)
    Xn = X + (Om - want)
    nt = lib.gg_mgl_ntile(p)
    parts = torch.zeros((nt * nt, 5), dtype=torch.float64, device=dev)
    dOm, dOmp, dX = to_dev(Om, dev), to_dev(Omp, dev), to_dev(X, dev)
    Th = torch.empty_like(dOm)
    rc = lib.gg_prox_mgl(_p(dOm), _p(dOmp), None, _p(dX), _p(Th), None, _p(_ctrl(dev, rho)), l1, l2,
                         0 if reg == "GGL" else 1, K, p, _p(parts), 0)
    assert rc == 0
    got = Th.cpu().numpy()
    assert np.abs(got - want).max() < 1e-15
    assert np.array_equal(got != 0, want != 0)
    assert np.array_equal(got, got.transpose(0, 2, 1))
    assert np.abs(dX.cpu().numpy() - Xn).max() < 1e-15
    sums = parts.sum(0).cpu().numpy()
    ref = [np.sum(Om ** 2), np.sum(want ** 2), np.sum(Xn ** 2), np.sum((Om - want) ** 2), np.sum((Om - Omp) ** 2)]
    np.testing.assert_allclose(sums, ref, rtol=1e-12)


@pytest.mark.parametrize("reg,K,p", [("GGL", 2, 16), ("FGL", 5, 33), ("GGL", 20, 70), ("FGL", 20, 257), ("FGL", 31, 17),
                                      ("GGL", 3, 1), ("FGL", 7, 100)])
def test_prox_mgl_upper_triangle_kernels(reg, K, p):
    """the upper-triangle loop kernels (gg_prox_mgl_upper, gg_build_w_upper) followed by gg_mirror_upper give the same
    Theta, X and residual sums as the oracle's prox_p + dual update on full matrices; the lower triangles are not
    touched before the mirror pass"""
    from gglasso_b200 import _lib
    from gglasso_b200._engine import to_dev, _p
    from oracle import admm_oracle as orc
    lib = _lib.load()
    dev = torch.device("cuda")
    rng = np.random.default_rng(K * 1000 + p)
    Om, Omp, X, S = (np.stack([_sym(rng, p, 0.3) for _ in range(K)]) for _ in range(4))
    rho, l1, l2 = 2.0, 0.21, 0.13
    want = orc.prox_p(Om + X, l1 / rho, l2 / rho, reg)
    Xn = X + (Om - want)
    n = lib.gg_mgl_upper_nparts(p)
    parts = torch.zeros((n, 5), dtype=torch.float64, device=dev)
    dOm, dOmp, dX, dS = to_dev(Om, dev), to_dev(Omp, dev), to_dev(X, dev), to_dev(S, dev)
    Th = torch.full_like(dOm, 7.0)
    ctrl = _ctrl(dev, rho)
    rc = lib.gg_prox_mgl_upper(_p(dOm), _p(dOmp), _p(dX), _p(Th), _p(ctrl), l1, l2, 0 if reg == "GGL" else 1, K, p,
                               _p(parts), 0)
    assert rc == 0
    iu = np.triu_indices(p)
    il = np.tril_indices(p, -1)
    got, gotX = Th.cpu().numpy(), dX.cpu().numpy()
    assert np.abs(got[:, iu[0], iu[1]] - want[:, iu[0], iu[1]]).max() < 1e-15
    assert np.all(got[:, il[0], il[1]] == 7.0)                       # lower triangle untouched
    assert np.array_equal(gotX[:, il[0], il[1]], X[:, il[0], il[1]])
    sums = parts.sum(0).cpu().numpy()
    ref = [np.sum(Om ** 2), np.sum(want ** 2), np.sum(Xn ** 2), np.sum((Om - want) ** 2), np.sum((Om - Omp) ** 2)]
    np.testing.assert_allclose(sums, ref, rtol=1e-12)
    # W on the upper triangle, with a pending X rescale
    ctrl[0, 1] = 0.5
    nk = to_dev(np.arange(1, K + 1, dtype=np.float64), dev)
    W = torch.full_like(dOm, -3.0)
    Xb = dX.clone()
    assert lib.gg_build_w_upper(_p(Th), _p(Xb), _p(dS), _p(nk), _p(ctrl), K, p, _p(W), 0) == 0
    Wg = W.cpu().numpy()
    Wref = got - 0.5 * gotX - (np.arange(1, K + 1)[:, None, None] / rho) * S
    assert np.abs(Wg[:, iu[0], iu[1]] - Wref[:, iu[0], iu[1]]).max() < 1e-14
    assert np.all(Wg[:, il[0], il[1]] == -3.0)
    assert np.array_equal(Xb.cpu().numpy()[:, iu[0], iu[1]], 0.5 * gotX[:, iu[0], iu[1]])
    # mirror pass
    assert lib.gg_mirror_upper(_p(Th), _p(dX), K, p, 0) == 0
    got, gotX = Th.cpu().numpy(), dX.cpu().numpy()
    assert np.array_equal(got, got.transpose(0, 2, 1)) and np.array_equal(gotX, gotX.transpose(0, 2, 1))
    assert np.abs(got - want).max() < 1e-15 and np.array_equal(got != 0, want != 0)
    assert np.abs(gotX - Xn).max() < 1e-15


@pytest.mark.parametrize("reg,K,p", [("GGL", 4, 120), ("FGL", 5, 97)])
def test_upper_triangle_loop_equals_full_loop(reg, K, p, monkeypatch):
    """ADMM_MGL with the upper-triangle iteration (default for the tridiagonal eigensolver path) and with the
    full-matrix kernels (GG_UPPER=0): same iteration count, same solution"""
    import gglasso_b200
    rng = np.random.default_rng(p)
    S = np.stack([np.cov(rng.standard_normal((p, 3 * p)), bias=True) for _ in range(K)])
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    sols = []
    for flag in ("1", "0"):
        monkeypatch.setenv("GG_UPPER", flag)
        sol, info = gglasso_b200.ADMM_MGL(S, 0.1, 0.05, reg, Om0, tol=1e-8, rtol=1e-8, verbose=False)
        sols.append((sol, info))
    (a, ia), (b, ib) = sols
    assert ia["status"] == ib["status"] == "optimal"
    for k in ("Omega", "Theta", "X"):
        assert np.abs(a[k] - b[k]).max() < 1e-11, k
        assert np.array_equal(a[k], a[k].transpose(0, 2, 1))
    assert np.array_equal(a["Theta"] != 0, b["Theta"] != 0)


@pytest.mark.parametrize("p,masked", [(1, False), (7, False), (100, True), (257, False)])
def test_prox_sgl_kernel(p, masked):
    from gglasso_b200 import _lib
    from gglasso_b200._engine import to_dev, _p
    from oracle import admm_oracle as orc
    lib = _lib.load()
    dev = torch.device("cuda")
    rng = np.random.default_rng(p)
    M = 3
    Om, Omp, X = (np.stack([_sym(rng, p, 0.4) for _ in range(M)]) for _ in range(3))
    rhos = np.array([1.0, 0.5, 4.0])
    lam = 0.17
    mask = np.abs(np.stack([_sym(rng, p) for _ in range(M)])) if masked else None
    ctrl = _ctrl(dev, 1.0, M)
    ctrl[:, 0] = torch.from_numpy(rhos).to(dev)
    n = lib.gg_sgl_nparts(p, M)
    parts = torch.zeros((M, n, 5), dtype=torch.float64, device=dev)
    dOm, dOmp, dX = to_dev(Om, dev), to_dev(Omp, dev), to_dev(X, dev)
    Th = torch.empty_like(dOm)
    lam_mat = to_dev(lam * mask, dev) if masked else None
    assert lib.gg_prox_sgl(_p(dOm), _p(dOmp), None, _p(dX), _p(Th), None, _p(ctrl), lam, _p(lam_mat), M, p,
                           _p(parts), None, 0) == 0
    got, gotX = Th.cpu().numpy(), dX.cpu().numpy()
    sums = parts.sum(1).cpu().numpy()
    for m in range(M):
        l = (1 / rhos[m]) * (lam * mask[m] if masked else lam)
        want = orc.prox_od_1norm(Om[m] + X[m], l)
        assert np.array_equal(got[m], want)
        Xn = X[m] + Om[m] - want
        assert np.array_equal(gotX[m], Xn)
        ref = [np.sum(Om[m] ** 2), np.sum(want ** 2), np.sum(Xn ** 2), np.sum((Om[m] - want) ** 2),
               np.sum((Om[m] - Omp[m]) ** 2)]
        np.testing.assert_allclose(sums[m], ref, rtol=1e-12)


# --------------------------------------------------------------------------------------------
# full solvers vs the real reference (golden fixtures) and vs the oracle, per iteration
# --------------------------------------------------------------------------------------------
def _check_hist(info_hist, traj):
    n = traj.shape[0]
    h = info_hist[:n]
    assert np.array_equal(h[:, 4], traj[:, 0]), "rho sequence differs from the reference"
    np.testing.assert_allclose(h[:, :4], traj[:, 1:5], rtol=1e-7, atol=1e-12)


MGL_CASES = ["mgl_ggl_K3_p50", "mgl_ggl_latent_K3_p50", "mgl_fgl_K3_p50", "mgl_fgl_latent_K3_p50",
             "mgl_ggl_K3_p50_maxiter2", "mgl_ggl_K3_p50_fixedrho", "mgl_ggl_K3_p50_nsamples",
             "cfg2_ggl", "cfg2_ggl_latent", "cfg2_fgl_tv"]


@pytest.mark.parametrize("name", MGL_CASES)
def test_admm_mgl_vs_reference_golden(golden, name):
    from gglasso_b200 import ADMM_MGL
    g = golden(name)
    S = g["S"]
    K, p, _ = S.shape
    kw = dict(ast.literal_eval(str(g["kw"])))
    (sol, info), out = _quiet(ADMM_MGL, S, float(g["lambda1"]), float(g["lambda2"]), str(g["reg"]),
                              np.repeat(np.eye(p)[None], K, 0), measure=True, **kw)
    n = g["traj"].shape[0]
    assert info["status"] == str(g["status"])
    assert f"ADMM terminated after {n} iterations with status: {info['status']}." in out
    assert len(info["residual"]) == n == len(info["runtime"]) == len(info["objective"])
    np.testing.assert_allclose(info["residual"], g["residual"], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(info["objective"], g["objective"], rtol=1e-6)
    assert _rel(sol["Theta"], g["Theta"]) < PER_ITER_TOL
    assert np.array_equal(sol["Theta"] != 0, g["Theta"] != 0), "sparsity pattern differs from the reference"
    assert set(sol) == {"Omega", "Theta", "L", "X"}
    if "Omega" in g.files:
        for k in ("Omega", "X", "L"):
            assert np.linalg.norm(sol[k] - g[k]) <= PER_ITER_TOL * max(1.0, np.linalg.norm(g[k])), k


@pytest.mark.parametrize("reg,latent", [("GGL", False), ("FGL", False), ("GGL", True), ("FGL", True)])
def test_admm_mgl_per_iteration_trajectory_vs_oracle(golden, reg, latent):
    """first 50 iterations: Omega/Theta/L/X within 1e-8 relative Frobenius of the CPU oracle, same rho path."""
    from gglasso_b200._engine import run_admm
    from oracle import admm_oracle as orc
    g = golden("mgl_ggl_K3_p50")
    S = g["S"]
    K, p, _ = S.shape
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    kw = dict(tol=1e-9, rtol=1e-9, max_iter=50)
    otrace = []
    orc.admm_mgl(S, 0.05, 0.01, reg, Om0, latent=latent, mu1=0.1 if latent else None, trace=otrace, **kw)
    trace = []
    st, res = run_admm("mgl", S, Om0, Om0, np.zeros_like(S), lambda1=0.05, lambda2=0.01, reg=reg, latent=latent,
                       mu=0.1 * np.ones(K) if latent else None, trace=trace, check_every=1, **kw)
    assert len(trace) == len(otrace)
    for t, (a, b) in enumerate(zip(trace, otrace)):
        for k in ("Omega", "Theta", "X") + (("L",) if latent else ()):
            assert np.linalg.norm(a[k] - b[k]) <= PER_ITER_TOL * max(np.linalg.norm(b[k]), 1e-3), (t, k)
        assert np.array_equal(a["Theta"] != 0, b["Theta"] != 0), t
    h = res["hist"][0][:len(otrace)]
    assert np.array_equal(h[:, 4], np.array([t["rho"] for t in otrace]))


SGL_CASES = ["cfg1_sgl", "cfg1_sgl_latent", "cfg1_sgl_mask"]


@pytest.mark.parametrize("name", SGL_CASES)
def test_admm_sgl_vs_reference_golden(golden, name):
    from gglasso_b200 import ADMM_SGL
    g = golden(name)
    S = g["S"]
    p = S.shape[0]
    kw = dict(tol=1e-7, rtol=1e-7)
    if "latent" in name:
        kw.update(latent=True, mu1=0.1)
    if "mask" in name:
        kw["lambda1_mask"] = g["lambda1_mask"]
    (sol, info), out = _quiet(ADMM_SGL, S, float(g["lambda1"]), np.eye(p), measure=True, **kw)
    n = g["traj"].shape[0]
    assert info["status"] == str(g["status"]) and len(info["residual"]) == n
    assert f"ADMM terminated after {n} iterations" in out
    np.testing.assert_allclose(info["residual"], g["residual"], rtol=1e-7, atol=1e-12)
    assert ("L" in sol) == ("latent" in name) and "objective" not in info
    for k in ("Theta", "Omega", "X") + (("L",) if "L" in sol else ()):
        assert np.linalg.norm(sol[k] - g[k]) <= PER_ITER_TOL * max(1.0, np.linalg.norm(g[k])), k
    assert np.array_equal(sol["Theta"] != 0, g["Theta"] != 0)


def test_kkt_criterion_vs_reference_golden(golden):
    from gglasso_b200 import ADMM_MGL, ADMM_SGL
    g = golden("mgl_fgl_K3_p50_kkt")
    S = g["S"]
    K, p, _ = S.shape
    (sol, info), _ = _quiet(ADMM_MGL, S, 0.05, 0.01, "FGL", np.repeat(np.eye(p)[None], K, 0), tol=1e-5,
                            stopping_criterion="kkt", update_rho=False, measure=True)
    assert info["status"] == str(g["status"]) and len(info["residual"]) == len(g["residual"])
    np.testing.assert_allclose(info["residual"], g["residual"], rtol=1e-6, atol=1e-12)
    assert _rel(sol["Theta"], g["Theta"]) < PER_ITER_TOL
    g = golden("cfg1_sgl_kkt")
    S = g["S"]
    (sol, info), _ = _quiet(ADMM_SGL, S, 0.05, np.eye(S.shape[0]), tol=1e-6, stopping_criterion="kkt", max_iter=200,
                            measure=True)
    assert info["status"] == str(g["status"]) and len(info["residual"]) == len(g["residual"])
    assert _rel(sol["Theta"], g["Theta"]) < PER_ITER_TOL


def test_block_sgl_vs_reference_golden(golden):
    from gglasso_b200 import block_SGL, get_connected_components
    g = golden("block_sgl_p100")
    S, lam = g["S"], float(g["lambda1"])
    numC, _ = get_connected_components(S, lam)
    assert numC == int(g["numC"])
    sol, _ = _quiet(block_SGL, S, lam, np.eye(S.shape[0]), tol=1e-9, rtol=1e-9)
    assert set(sol) == {"Omega", "Theta", "X"}
    for k in ("Theta", "Omega", "X"):
        assert np.linalg.norm(sol[k] - g[k]) <= PER_ITER_TOL * max(1.0, np.linalg.norm(g[k])), k
    assert np.array_equal(sol["Theta"] != 0, g["Theta"] != 0)


def test_mask_of_zeros_gives_inverse():
    """known-answer test of the reference (tests/test_solvers.py:191-246)."""
    from gglasso_b200 import ADMM_SGL
    rng = np.random.default_rng(0)
    p = 30
    A = rng.standard_normal((4 * p, p))
    S = A.T @ A / (4 * p)
    (sol, info), _ = _quiet(ADMM_SGL, S, 0.1, np.eye(p), tol=1e-10, rtol=1e-10, lambda1_mask=np.zeros((p, p)))
    np.testing.assert_allclose(sol["Theta"], np.linalg.inv(S), atol=1e-4)


@pytest.mark.parametrize("reg,K,p", [("GGL", 4, 200), ("FGL", 3, 333)])
def test_admm_mgl_block_eigh_path_vs_oracle(reg, K, p):
    """p > 160 exercises the block-Jacobi (DMMA) eigensolver inside the loop."""
    from gglasso_b200 import ADMM_MGL
    from gglasso_b200.datagen import synthetic_mgl
    from oracle import admm_oracle as orc
    S = synthetic_mgl(K, p, N=2 * p, seed=3)
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    (sol, info), _ = _quiet(ADMM_MGL, S, 0.05, 0.02, reg, Om0, tol=1e-7, rtol=1e-7, measure=True)
    ref, rinfo = orc.admm_mgl(S, 0.05, 0.02, reg, Om0, tol=1e-7, rtol=1e-7, measure=True)
    assert info["status"] == rinfo["status"] and len(info["residual"]) == rinfo["iterations"]
    for k in ("Omega", "Theta", "X"):
        assert _rel(sol[k], ref[k]) < PER_ITER_TOL, k
    assert np.array_equal(sol["Theta"] != 0, ref["Theta"] != 0)
    assert abs(info["objective"][-1] - rinfo["objective"][-1]) <= 1e-6 * abs(rinfo["objective"][-1])


@pytest.mark.parametrize("reg,K,p", [("GGL", 1, 49), ("FGL", 2, 50), ("GGL", 33, 63), ("FGL", 5, 64), ("GGL", 3, 65),
                                      ("FGL", 90, 49), ("GGL", 2, 48)])
def test_admm_mgl_default_loop_at_size_boundaries_vs_oracle(reg, K, p):
    """default options (upper-triangle iteration for p > 48, full-matrix kernels at p = 48; K up to the tile limit) at
    the sizes where the eigensolver path, the tile shapes and the K limit of the fused prox change"""
    from gglasso_b200 import ADMM_MGL
    from gglasso_b200.datagen import synthetic_mgl
    from oracle import admm_oracle as orc
    S = synthetic_mgl(K, p, N=3 * p, seed=K + p)
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    (sol, info), _ = _quiet(ADMM_MGL, S, 0.08, 0.03, reg, Om0, tol=1e-7, rtol=1e-7)
    ref, rinfo = orc.admm_mgl(S, 0.08, 0.03, reg, Om0, tol=1e-7, rtol=1e-7)
    assert info["status"] == rinfo["status"]
    for k in ("Omega", "Theta", "X"):
        assert _rel(sol[k], ref[k]) < PER_ITER_TOL, k
        assert np.array_equal(sol[k], sol[k].transpose(0, 2, 1)), k
    assert np.array_equal(sol["Theta"] != 0, ref["Theta"] != 0)


def test_input_validation_matches_reference():
    from gglasso_b200 import ADMM_MGL, ADMM_SGL
    S = np.repeat(np.eye(4)[None], 2, 0)
    with pytest.raises(AssertionError):
        ADMM_MGL(S, 0.1, 0.1, "XYZ", S)
    with pytest.raises(AssertionError):
        ADMM_MGL(S, -0.1, 0.1, "GGL", S)
    with pytest.raises(AssertionError):
        ADMM_SGL(np.eye(4), 0.1, np.eye(5))
    with pytest.raises(AssertionError):
        ADMM_SGL(np.eye(4), 0.1, np.eye(4), latent=True)


@pytest.mark.parametrize("reg,latent", [("GGL", False), ("FGL", False), ("FGL", True)])
def test_k_sharded_loop_single_rank_equals_fused_loop(reg, latent):
    """ADMM_MGL_dist at world_size 1 (pack -> band prox -> dual update kernels) must reproduce ADMM_MGL."""
    from gglasso_b200 import ADMM_MGL
    from gglasso_b200.parallel import ADMM_MGL_dist
    from gglasso_b200.datagen import synthetic_mgl
    K, p = 4, 70
    S = synthetic_mgl(K, p, N=2 * p, seed=5, kind="fused" if reg == "FGL" else "group")
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    kw = dict(tol=1e-7, rtol=1e-7, latent=latent)
    (ref, rinfo), _ = _quiet(ADMM_MGL, S, 0.05, 0.02, reg, Om0, measure=True, mu1=0.1 if latent else None, **kw)
    sol, info = ADMM_MGL_dist(S, 0.05, 0.02, reg, Om0, mu1_local=0.1 if latent else None, **kw)
    assert info["status"] == rinfo["status"] and info["iterations"] == len(rinfo["residual"])
    for k in ("Omega", "Theta", "X", "L"):
        assert np.abs(sol[k] - ref[k]).max() < 1e-10, k
    assert np.array_equal(sol["Theta"] != 0, ref["Theta"] != 0)


@pytest.mark.parametrize("K_loc,p,world", [(3, 10, 4), (2, 7, 7), (5, 64, 2), (4, 33, 1), (2, 101, 8)])
def test_pack_unpack_band_kernels(K_loc, p, world):
    """gg_pack_bands / gg_unpack_dual against the host statement of the layout (band_layout_index) and numpy."""
    from gglasso_b200 import _lib
    from gglasso_b200._engine import to_dev, _p
    from gglasso_b200._lib import NPART
    from gglasso_b200.parallel import band_layout_index
    lib = _lib.load()
    dev = torch.device("cuda")
    rng = np.random.default_rng(K_loc * 100 + p)
    Om, Omp, L, X, Th = (rng.standard_normal((K_loc, p, p)) for _ in range(5))
    idx = band_layout_index(K_loc, p, world).reshape(-1)
    ctrl = _ctrl(dev)
    send = torch.zeros(K_loc * p * p, dtype=torch.float64, device=dev)
    for Lh in (None, L):
        assert lib.gg_pack_bands(_p(to_dev(Om, dev)), _p(None if Lh is None else to_dev(Lh, dev)), _p(to_dev(X, dev)),
                                 _p(ctrl), K_loc, p, world, _p(send), 0) == 0
        want = np.empty(K_loc * p * p)
        want[idx] = ((Om + Lh) + X if Lh is not None else Om + X).reshape(-1)
        assert np.array_equal(send.cpu().numpy(), want)
    recv = np.empty(K_loc * p * p)
    recv[idx] = Th.reshape(-1)
    nparts = lib.gg_sgl_nparts(p, K_loc) * K_loc
    for latent in (False, True):
        Xd, Thd = to_dev(X, dev), torch.zeros((K_loc, p, p), dtype=torch.float64, device=dev)
        Cd = torch.zeros_like(Thd) if latent else None
        parts = torch.zeros((nparts, NPART), dtype=torch.float64, device=dev)
        assert lib.gg_unpack_dual(_p(to_dev(recv, dev)), _p(to_dev(Om, dev)), _p(to_dev(Omp, dev)), _p(Xd), _p(Thd), _p(Cd),
                                  _p(ctrl), K_loc, p, world, _p(parts), 0) == 0
        assert np.array_equal(Thd.cpu().numpy(), Th)
        if latent:
            assert np.array_equal(Cd.cpu().numpy(), (Th - X) - Om) and np.array_equal(Xd.cpu().numpy(), X)
        else:
            Xn = X + (Om - Th)
            assert np.array_equal(Xd.cpu().numpy(), Xn)
            want = [np.sum(Om ** 2), np.sum(Th ** 2), np.sum(Xn ** 2), np.sum((Om - Th) ** 2), np.sum((Om - Omp) ** 2)]
            np.testing.assert_allclose(parts.sum(0).cpu().numpy(), want, rtol=1e-12)


@pytest.mark.parametrize("reg,K,p,world", [("GGL", 7, 45, 4), ("FGL", 20, 33, 8), ("FGL", 5, 64, 2), ("GGL", 3, 10, 3)])
def test_peer_memory_exchange_kernels_on_one_device(reg, K, p, world):
    """gg_pack_bands_p2p / gg_prox_band_p2p with every "rank" played in turn on one GPU (the peer pointers are local
    buffers): uneven instance and row partitions, result = prox_p of the whole stack in gg_unpack_dual's layout"""
    import ctypes
    from gglasso_b200 import _lib
    from gglasso_b200._engine import to_dev, _p
    from gglasso_b200.parallel import partition, band_layout_index
    from oracle import admm_oracle as orc
    lib = _lib.load()
    dev = torch.device("cuda")
    rng = np.random.default_rng(K * 10 + p + world)
    Om = np.stack([_sym(rng, p, 0.3) for _ in range(K)])
    X = np.stack([_sym(rng, p, 0.2) for _ in range(K)])
    rho, l1, l2 = 2.0, 0.21, 0.13
    want = orc.prox_p(Om + X, l1 / rho, l2 / rho, reg)
    kparts, rparts = partition(K, world), partition(p, world)
    bands = [torch.full((K * (hi - lo) * p,), np.nan, dtype=torch.float64, device=dev) for lo, hi in rparts]
    backs = [torch.full((max(1, (hi - lo)) * p * p,), np.nan, dtype=torch.float64, device=dev) for lo, hi in kparts]
    arr = ctypes.c_void_p * 16
    band_ptrs = arr(*[b.data_ptr() for b in bands])
    back_ptrs = arr(*[b.data_ptr() for b in backs])
    ctrl = _ctrl(dev, rho)
    for (klo, khi) in kparts:                                   # every rank packs its instances into all band buffers
        if khi == klo:
            continue
        dO, dX = to_dev(Om[klo:khi], dev), to_dev(X[klo:khi], dev)
        assert lib.gg_pack_bands_p2p(_p(dO), None, _p(dX), _p(ctrl), khi - klo, p, world, klo, band_ptrs, 0) == 0
    torch.cuda.synchronize()
    for d, (lo, hi) in enumerate(rparts):                        # band buffer of rank d = rows lo..hi of V, all instances
        got = bands[d].cpu().numpy().reshape(K, hi - lo, p)
        assert np.array_equal(got, (Om + X)[:, lo:hi, :])
        if hi > lo:
            assert lib.gg_prox_band_p2p(_p(bands[d]), back_ptrs, _p(ctrl), l1, l2, 0 if reg == "GGL" else 1, K, hi - lo,
                                        p, lo, world, 0) == 0
    torch.cuda.synchronize()
    for s, (klo, khi) in enumerate(kparts):                      # receive buffer of rank s, in gg_unpack_dual's layout
        if khi == klo:
            continue
        idx = band_layout_index(khi - klo, p, world)
        got = backs[s].cpu().numpy()[idx]
        assert np.abs(got - want[klo:khi]).max() < 1e-15
        assert np.array_equal(got != 0, want[klo:khi] != 0)


@pytest.mark.parametrize("reg,latent", [("FGL", False), ("GGL", True)])
def test_k_sharded_check_every_keeps_converged_state(reg, latent):
    """iterations enqueued after convergence (check_every > 1) are no-ops: Theta / X / Omega are those of the last
    executed iteration (the band buffers persist across iterations)."""
    from gglasso_b200 import ADMM_MGL
    from gglasso_b200.parallel import ADMM_MGL_dist
    from gglasso_b200.datagen import synthetic_mgl
    K, p = 4, 70
    S = synthetic_mgl(K, p, N=2 * p, seed=5, kind="fused" if reg == "FGL" else "group")
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    kw = dict(tol=1e-7, rtol=1e-7, latent=latent)
    (ref, rinfo), _ = _quiet(ADMM_MGL, S, 0.05, 0.02, reg, Om0, measure=True, mu1=0.1 if latent else None, **kw)
    for ce in (4, 7):
        sol, info = ADMM_MGL_dist(S, 0.05, 0.02, reg, Om0, mu1_local=0.1 if latent else None, check_every=ce, **kw)
        assert info["status"] == rinfo["status"] and info["iterations"] == len(rinfo["residual"])
        for k in ("Omega", "Theta", "X", "L"):
            assert np.abs(sol[k] - ref[k]).max() < 1e-10, (ce, k)


@pytest.mark.parametrize("kw", [dict(stopping_criterion="kkt", tol=1e-5, update_rho=False), dict(measure=True, tol=1e-7, rtol=1e-7),
                                dict(latent=True, mu1=0.2, tol=1e-7, rtol=1e-7, verbose=True)])
def test_admm_mgl_large_K_all_options(kw):
    """K > 90 with the KKT criterion, measure=True and latent variables: same behaviour as for small K (header,
    objective, post-loop checks) -- the reference handles any K (admm_solver.py:13-313)."""
    from gglasso_b200 import ADMM_MGL
    from oracle import admm_oracle as orc
    rng = np.random.default_rng(19)
    K, p, N = 101, 10, 50
    S = np.stack([np.cov(rng.standard_normal((p, N)), bias=True) for _ in range(K)])
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    (sol, info), out = _quiet(ADMM_MGL, S, 0.1, 0.05, "GGL", Om0, **kw)
    okw = {k: v for k, v in kw.items() if k not in ("measure", "verbose")}
    ref, rinfo = orc.admm_mgl(S, 0.1, 0.05, "GGL", Om0, **okw)
    assert info["status"] == rinfo["status"] and f"after {rinfo['iterations']} iterations" in out
    for k in ("Omega", "Theta", "X") + (("L",) if kw.get("latent") else ()):
        assert _rel(sol[k], ref[k]) < 1e-7, k
    if kw.get("measure"):
        assert len(info["objective"]) == rinfo["iterations"] == len(info["residual"])
    if kw.get("verbose"):
        assert out.startswith("------------ADMM Algorithm for Multiple Graphical Lasso----------------")


@pytest.mark.parametrize("exchange", ["nccl all-to-all", "peer memory"])
def test_k_sharded_two_ranks_nccl(exchange):
    """2-GPU check (skipped on a 1-GPU box): K-sharded solve vs the single-GPU solve, with the NCCL all-to-all re-tile
    and with the kernels that store straight into the peers' buffers (GG_DIST_P2P=1, symmetric memory)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GG_DIST_P2P="1" if exchange == "peer memory" else "0")
    o = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "scripts", "dist_check.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert "DIST_CHECK_OK" in o.stdout, o.stdout[-2000:] + o.stderr[-2000:]
    assert f"DIST_EXCHANGE {exchange}" in o.stdout, o.stdout[-2000:]


def test_block_sgl_ragged_batches_vs_oracle():
    """many components of different sizes (singletons, 2..45, one > 160): batched ragged solves, per-block rho and
    stopping test on the device, must reproduce the per-block sequential oracle."""
    from gglasso_b200 import block_SGL, get_connected_components
    from oracle import admm_oracle as orc
    rng = np.random.default_rng(42)
    sizes = [1, 1, 2, 2, 3, 5, 5, 8, 13, 17, 21, 33, 45, 170, 1, 4]
    p = sum(sizes)
    S = np.zeros((p, p))
    o = 0
    for s in sizes:
        Z = rng.standard_normal((s, 3 * s + 5))
        B = Z @ Z.T / (3 * s + 5)
        B = B / np.sqrt(np.outer(np.diag(B), np.diag(B)))
        S[o:o + s, o:o + s] = 0.6 * B + 0.4 * np.eye(s)
        o += s
    perm = rng.permutation(p)
    S = S[np.ix_(perm, perm)]
    lam = 0.02
    numC, comps = get_connected_components(S, lam)
    assert numC >= 10 and max(len(c) for c in comps) > 160
    (sol, out) = _quiet(block_SGL, S, lam, np.eye(p), tol=1e-8, rtol=1e-8)
    ref = orc.block_sgl(S, lam, np.eye(p), tol=1e-8, rtol=1e-8)
    for k in ("Theta", "Omega", "X"):
        assert np.linalg.norm(sol[k] - ref[k]) <= PER_ITER_TOL * max(1.0, np.linalg.norm(ref[k])), k
    assert np.array_equal(sol["Theta"] != 0, ref["Theta"] != 0)
    assert out.count("ADMM terminated after") == sum(1 for c in comps if len(c) > 1)


def test_block_sgl_with_mask_and_warm_start_arrays():
    from gglasso_b200 import block_SGL
    from oracle import admm_oracle as orc
    rng = np.random.default_rng(3)
    p = 60
    A = rng.standard_normal((p, 4 * p))
    S = A @ A.T / (4 * p)
    S[np.abs(S) < 0.12] = 0.0
    S = (S + S.T) / 2 + 0.5 * np.eye(p)
    mask = np.ones((p, p))
    mask[:10, :10] = 2.0
    Om0 = np.eye(p) * 1.5
    X0 = np.zeros((p, p))
    sol, _ = _quiet(block_SGL, S, 0.1, Om0, Theta_0=np.eye(p), X_0=X0, tol=1e-8, rtol=1e-8, lambda1_mask=mask)
    ref = orc.block_sgl(S, 0.1, Om0, Theta_0=np.eye(p), X_0=X0, tol=1e-8, rtol=1e-8, lambda1_mask=mask)
    for k in ("Theta", "Omega", "X"):
        assert np.linalg.norm(sol[k] - ref[k]) <= PER_ITER_TOL * max(1.0, np.linalg.norm(ref[k])), k


@pytest.mark.parametrize("reg", ["GGL", "FGL"])
def test_admm_mgl_large_K_uses_band_prox(reg):
    """K beyond the tile-pair kernel's shared-memory budget (K > 90) goes through the row-band prox."""
    from gglasso_b200 import ADMM_MGL
    from oracle import admm_oracle as orc
    rng = np.random.default_rng(9)
    K, p, N = 130, 12, 60
    S = np.stack([np.cov(rng.standard_normal((p, N)), bias=True) for _ in range(K)])
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    (sol, info), out = _quiet(ADMM_MGL, S, 0.1, 0.05, reg, Om0, tol=1e-7, rtol=1e-7)
    ref, rinfo = orc.admm_mgl(S, 0.1, 0.05, reg, Om0, tol=1e-7, rtol=1e-7)
    assert info["status"] == rinfo["status"] and f"after {rinfo['iterations']} iterations" in out
    for k in ("Omega", "Theta", "X"):
        assert _rel(sol[k], ref[k]) < PER_ITER_TOL, k
    assert np.array_equal(sol["Theta"] != 0, ref["Theta"] != 0)


def test_grid_search_on_device_solver_matches_oracle_solver():
    """lambda grid driver (grid_search_dist, single rank) with the B200 solver vs the same driver with the CPU oracle."""
    from gglasso_b200 import ADMM_MGL
    from gglasso_b200.parallel import grid_search_dist
    from oracle import admm_oracle as orc
    rng = np.random.default_rng(3)
    K, p, N = 3, 30, 150
    S = np.stack([np.cov(rng.standard_normal((p, N)), bias=True) for _ in range(K)])
    l1, l2 = np.logspace(-0.5, -1.5, 3), np.logspace(-1, -2, 2)

    def cpu_solver(S, a, b, reg, Om0, tol=1e-7, rtol=1e-7, **kw):
        return orc.admm_mgl(S, a, b, reg, Om0, tol=tol, rtol=rtol)
    (sc_g, ix_g, best_g), _ = _quiet(grid_search_dist, ADMM_MGL, S, np.full(K, N), "GGL", l1, l2, gamma=0.1)
    sc_c, ix_c, best_c = grid_search_dist(cpu_solver, S, np.full(K, N), "GGL", l1, l2, gamma=0.1)
    np.testing.assert_allclose(sc_g, sc_c, rtol=1e-8)
    assert tuple(ix_g) == tuple(ix_c)
    assert _rel(best_g["Theta"], best_c["Theta"]) < PER_ITER_TOL


def test_grid_search_device_resident_matches_host_driver():
    """device-resident grid (scores on the GPU, warm starts never leave the device) vs the host-scored driver."""
    from gglasso_b200 import ADMM_MGL
    from gglasso_b200.parallel import grid_search_device, grid_search_dist
    rng = np.random.default_rng(3)
    K, p, N = 3, 30, 150
    S = np.stack([np.cov(rng.standard_normal((p, N)), bias=True) for _ in range(K)])
    l1, l2 = np.logspace(-0.5, -1.5, 3), np.logspace(-1, -2, 2)
    sc_d, it_d, ix_d, best_d = grid_search_device(S, np.full(K, N), "GGL", l1, l2, gamma=0.1)
    (sc_h, ix_h, best_h), _ = _quiet(grid_search_dist, ADMM_MGL, S, np.full(K, N), "GGL", l1, l2, gamma=0.1)
    np.testing.assert_allclose(sc_d, sc_h, rtol=1e-9)
    assert tuple(ix_d) == tuple(ix_h) and it_d.min() >= 1
    for k in ("Omega", "Theta", "X"):
        assert _rel(best_d[k], best_h[k]) < PER_ITER_TOL, k
    sc_a, _, _, _ = grid_search_device(S, np.full(K, N), "FGL", l1[:1], l2[:1], method="AIC")
    assert np.isfinite(sc_a).all()
    # several columns concurrently on one GPU (worker threads + streams) must give the same table
    sc_s, it_s, ix_s, best_s = grid_search_device(S, np.full(K, N), "GGL", l1, l2, gamma=0.1, n_streams=3)
    np.testing.assert_allclose(sc_s, sc_d, rtol=1e-9)
    assert tuple(ix_s) == tuple(ix_d) and np.array_equal(it_s, it_d)
    assert _rel(best_s["Theta"], best_d["Theta"]) < PER_ITER_TOL


def test_grid_many_concurrent_columns_large_p():
    """eight columns on eight host threads / CUDA streams at a size that takes the full large-p eigensolver path
    (p >= 256: lazy-write sytrd with programmatic dependent launch, D&C, blocked back-transformation): the score
    table, iteration counts and optimum must equal the single-stream run.  (Regression test: a side stream inside
    gg_eigh once corrupted results only under this kind of load.)"""
    from gglasso_b200.datagen import synthetic_mgl
    from gglasso_b200.parallel import grid_search_device
    K, p = 6, 384
    S = synthetic_mgl(K, p, N=2 * p, seed=11)
    l1, l2 = np.logspace(-0.3, -1.7, 8), np.logspace(-1, -2, 2)
    sc1, it1, ix1, b1 = grid_search_device(S, np.full(K, 2 * p), "GGL", l1, l2, gamma=0.1, tol=1e-6, rtol=1e-6)
    for rep in range(2):
        sc8, it8, ix8, b8 = grid_search_device(S, np.full(K, 2 * p), "GGL", l1, l2, gamma=0.1, tol=1e-6, rtol=1e-6,
                                               n_streams=8)
        assert np.isfinite(sc8).all()
        np.testing.assert_allclose(sc8, sc1, rtol=1e-8)
        assert tuple(ix8) == tuple(ix1) and np.abs(it8 - it1).max() <= 1
        assert _rel(b8["Theta"], b1["Theta"]) < 1e-6


def test_admm_fsgl_vs_reference_golden(golden):
    """functional SGL (block-Frobenius prox) vs the real reference's output, with and without latent variables."""
    from gglasso_b200 import ADMM_FSGL
    g = golden("fsgl_p12_M3")
    S, M, lam = g["S"], int(g["M"]), float(g["lambda1"])
    pM = S.shape[0]
    for tag, lat in (("nolat", False), ("lat", True)):
        (sol, info), out = _quiet(ADMM_FSGL, S, lam, M, np.eye(pM), tol=1e-7, rtol=1e-7, measure=True, latent=lat,
                                  mu1=0.2 if lat else None)
        n = g[f"traj_{tag}"].shape[0]
        assert info["status"] == str(g[f"status_{tag}"]) and len(info["residual"]) == n
        assert f"ADMM terminated after {n} iterations" in out
        np.testing.assert_allclose(info["residual"], g[f"residual_{tag}"], rtol=1e-7, atol=1e-12)
        for k in ("Theta", "Omega", "X"):
            ref = g[f"{k}_{tag}"]
            assert np.linalg.norm(sol[k] - ref) <= PER_ITER_TOL * max(1.0, np.linalg.norm(ref)), (tag, k)
        assert np.array_equal(sol["Theta"] != 0, g[f"Theta_{tag}"] != 0)
        assert ("L" in sol) == lat
    assert np.linalg.norm(sol["L"] - g["L_lat"]) <= PER_ITER_TOL * max(1.0, np.linalg.norm(g["L_lat"]))


def test_admm_fsgl_M1_equals_sgl_and_verbose_format():
    """reference tests/test_func_gl.py: for M=1 the functional solver is the single graphical lasso."""
    from gglasso_b200 import ADMM_FSGL, ADMM_SGL
    rng = np.random.default_rng(123)
    p = 20
    S = np.cov(rng.standard_normal((p, 100)), bias=True)
    (s1, _), _ = _quiet(ADMM_SGL, S, 0.01, np.eye(p), tol=1e-10, rtol=1e-10)
    (s2, _), out = _quiet(ADMM_FSGL, S, 0.01, 1, np.eye(p), tol=1e-10, rtol=1e-10, verbose=True)
    assert np.abs(s1["Theta"] - s2["Theta"]).max() < 1e-9
    lines = out.splitlines()
    assert lines[0] == f"Derived a Functional SGL problem of dimensionality p={p}."
    assert lines[2] == "%4s\t%10s\t%10s\t%10s\t%10s\t%10s" % ("iter", "r_t", "s_t", "eps_pri", "eps_dual", "rho")
    assert len(lines[3].split("\t")) == 6


@pytest.mark.parametrize("M,p", [(1, 1289), (2, 2047)])
def test_eigh_large_odd_sizes(M, p):
    """sizes that are neither multiples of the tile sizes nor powers of two (multi-level D&C, ragged tiles)."""
    from gglasso_b200._engine import eigh
    from gglasso_b200.datagen import synthetic_mgl
    A = np.eye(p)[None] - synthetic_mgl(M, p, N=2 * p, seed=4)
    D, Q = eigh(A)
    Dref = np.linalg.eigvalsh(A)
    assert np.abs(D - Dref).max() < 1e-11 * p ** 0.5
    for m in range(M):
        assert np.abs(Q[m].T @ Q[m] - np.eye(p)).max() < 1e-12
        assert np.abs(A[m] @ Q[m] - Q[m] * D[m]).max() < 1e-11


EXT_CASES = [("boyd", dict(tol=1e-7, rtol=1e-7)), ("latent", dict(tol=1e-7, rtol=1e-7, latent=True, mu1=0.3)),
             ("kkt", dict(tol=1e-4, stopping_criterion="kkt", max_iter=400))]


@pytest.mark.parametrize("tag,kw", EXT_CASES)
def test_ext_admm_mgl_vs_reference_golden(golden, tag, kw):
    """non-conforming group graphical lasso (three instances of size 12/14/12 sharing variables) vs the reference."""
    from gglasso_b200 import ext_ADMM_MGL
    g = golden("ext_mgl_K3")
    p = g["p"]
    S = {k: g[f"S{k}"] for k in range(3)}
    Om0 = {k: np.eye(int(p[k])) for k in range(3)}
    (sol, info), out = _quiet(ext_ADMM_MGL, S, float(g["lambda1"]), float(g["lambda2"]), "GGL", Om0, g["G"],
                              measure=True, **kw)
    n = len(g[f"residual_{tag}"])
    assert info["status"] == str(g[f"status_{tag}"]) and len(info["residual"]) == n
    assert f"ADMM terminated after {n} iterations" in out
    np.testing.assert_allclose(info["residual"], g[f"residual_{tag}"], rtol=1e-6, atol=1e-12)
    assert set(sol) == {"Omega", "Theta", "L", "X0", "X1"}
    for name in sol:
        for k in range(3):
            ref = g[f"{name}{k}_{tag}"]
            assert sol[name][k].shape == ref.shape
            assert np.linalg.norm(sol[name][k] - ref) <= PER_ITER_TOL * max(1.0, np.linalg.norm(ref)), (name, k)
    for k in range(3):
        assert np.array_equal(sol["Theta"][k] != 0, g[f"Theta{k}_{tag}"] != 0)


def test_ext_admm_consistent_with_conforming_solver(golden):
    """reference tests/test_solvers.py:71-120: with all variables in all instances (trivial G) and lambda2/sqrt(K)
    the extended solver solves the same problem as ADMM_MGL."""
    from gglasso_b200 import ADMM_MGL, ext_ADMM_MGL
    g = golden("ext_mgl_K3")
    G = g["G_trivial"]
    rng = np.random.default_rng(7)
    K, p = 3, 12
    S = np.stack([np.cov(rng.standard_normal((p, 200)), bias=True) for _ in range(K)])
    Sd = {k: S[k].copy() for k in range(K)}
    Om0 = {k: np.eye(p) for k in range(K)}
    (se, _), _ = _quiet(ext_ADMM_MGL, Sd, 0.05, 0.02 / np.sqrt(K), "GGL", Om0, G, tol=1e-9, rtol=1e-9)
    (sm, _), _ = _quiet(ADMM_MGL, S, 0.05, 0.02, "GGL", np.repeat(np.eye(p)[None], K, 0), tol=1e-9, rtol=1e-9)
    for k in range(K):
        assert np.abs(se["Theta"][k] - sm["Theta"][k]).max() < 1e-4


@pytest.mark.parametrize("p", [6990, 7100])
def test_eigh_size_limits(p):
    """p = 6990: largest size on the tridiagonal divide & conquer path; p = 7100: the documented fallback to the
    block-Jacobi path above 7000. Checked on the device (no host LAPACK at this size)."""
    from gglasso_b200._engine import Eigh
    g = torch.Generator(device="cuda").manual_seed(p)
    A = torch.randn(p, p, dtype=torch.float64, device="cuda", generator=g)
    A = ((A + A.T) / (2 * p ** 0.5)).contiguous()[None]
    W = A.clone()
    e = Eigh(1, p, torch.device("cuda"))
    D = e.eigh(W, stream=torch.cuda.current_stream().cuda_stream)      # W -> Vt (rows = eigenvectors)
    V = W[0]
    resid = (V @ A[0] - D[0][:, None] * V).abs().max().item()
    orth = (V @ V.T - torch.eye(p, dtype=torch.float64, device="cuda")).abs().max().item()
    assert resid < 1e-11 and orth < 1e-11, (resid, orth)
