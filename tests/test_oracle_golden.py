"""Pin the oracle (oracle/admm_oracle.py + oracle/gg_oracle.c) to the real reference.

The fixtures in tests/golden/*.npz were produced by fabian-sp/GGLasso itself
(tests/golden/make_golden.py).  Tolerances: the oracle uses the same LAPACK/BLAS as the
reference, so trajectories agree to rounding (1e-10 relative is generous); sparsity patterns
and iteration counts must be identical.
"""
import ast

import numpy as np
import pytest

from oracle import admm_oracle as orc

RTOL = 1e-10


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_prox_units(golden):
    g = golden("prox_units")
    X, l1, l2 = g["X"], float(g["l1"]), float(g["l2"])
    assert _rel(orc.prox_p(X, l1, l2, "GGL"), g["prox_p_ggl"]) < 1e-15
    assert np.array_equal(orc.prox_p(X, l1, l2, "FGL"), g["prox_p_fgl"])
    assert abs(orc.P_val(X, l1, l2, "GGL") - float(g["pval_ggl"])) < 1e-12
    assert abs(orc.P_val(X, l1, l2, "FGL") - float(g["pval_fgl"])) < 1e-12
    for y, lam, x in zip(g["tv_y"], g["tv_lam"], g["tv_x"]):
        assert np.array_equal(orc.tv1d(y, lam), x)
        assert np.array_equal(orc.tv1d_py(y, lam), x)
    A, D, Q = g["A"], g["D"], g["Q"]
    assert np.array_equal(orc.prox_od_1norm(A, 0.4), g["od1_scalar"])
    assert np.array_equal(orc.prox_od_1norm(A, 0.4 * g["mask"]), g["od1_mask"])
    assert _rel(orc.phiplus(0.7, D, Q), g["phiplus"]) < 1e-15
    assert _rel(orc.prox_rank_norm(D, Q, 0.9), g["rank_norm"]) < 1e-15


def test_prox_p_numpy_fallback_matches_c(golden):
    g = golden("prox_units")
    X, l1, l2 = g["X"], float(g["l1"]), float(g["l2"])
    saved = orc._CLIB
    try:
        orc._CLIB = False
        assert _rel(orc.prox_p(X, l1, l2, "GGL"), g["prox_p_ggl"]) < 1e-15
        assert np.array_equal(orc.prox_p(X, l1, l2, "FGL"), g["prox_p_fgl"])
        assert abs(orc.P_val(X, l1, l2, "FGL") - float(g["pval_fgl"])) < 1e-12
    finally:
        orc._CLIB = saved


def _check_traj(trace, g, status, info):
    traj = g["traj"]          # rho r s e_pri e_dual |Omega| |Theta| |L| |X| nnz
    assert info["status"] == str(g["status"])
    if traj.shape[0]:
        assert len(trace) == traj.shape[0], (len(trace), traj.shape[0])
        mine = np.array([[t["rho"], t["r"], t["s"], t["e_pri"], t["e_dual"], np.linalg.norm(t["Omega"]),
                          np.linalg.norm(t["Theta"]), np.linalg.norm(t["L"]), np.linalg.norm(t["X"])] for t in trace])
        assert np.array_equal(mine[:, 0], traj[:, 0]), "rho sequence differs"
        np.testing.assert_allclose(mine[:, 1:], traj[:, 1:9], rtol=1e-8, atol=1e-13)
        nnz = np.array([np.count_nonzero(t["Theta"]) for t in trace])
        assert np.array_equal(nnz, traj[:, 9].astype(int))


MGL = ["mgl_ggl_K3_p50", "mgl_ggl_latent_K3_p50", "mgl_fgl_K3_p50", "mgl_fgl_latent_K3_p50",
       "mgl_ggl_K3_p50_maxiter2", "mgl_ggl_K3_p50_fixedrho", "mgl_fgl_K3_p50_kkt", "mgl_ggl_K3_p50_nsamples",
       "cfg2_ggl", "cfg2_ggl_latent", "cfg2_fgl_tv"]


@pytest.mark.parametrize("name", MGL)
def test_admm_mgl_matches_reference(golden, name):
    g = golden(name)
    S = g["S"]
    K, p, _ = S.shape
    kw = dict(ast.literal_eval(str(g["kw"])))
    trace = []
    sol, info = orc.admm_mgl(S, float(g["lambda1"]), float(g["lambda2"]), str(g["reg"]),
                             np.repeat(np.eye(p)[None], K, 0), measure=True, trace=trace, **kw)
    if kw.get("stopping_criterion", "boyd") == "boyd":
        _check_traj(trace, g, str(g["status"]), info)
    else:
        assert info["status"] == str(g["status"])
    assert len(info["residual"]) == len(g["residual"])
    np.testing.assert_allclose(info["residual"], g["residual"], rtol=1e-7, atol=1e-13)
    np.testing.assert_allclose(info["objective"], g["objective"], rtol=1e-10)
    assert _rel(sol["Theta"], g["Theta"]) < RTOL
    assert np.array_equal(sol["Theta"] != 0, g["Theta"] != 0)
    if "Omega" in g.files:
        for k in ("Omega", "X", "L"):
            assert np.linalg.norm(sol[k] - g[k]) <= RTOL * max(1.0, np.linalg.norm(g[k]))


SGL = ["cfg1_sgl", "cfg1_sgl_latent", "cfg1_sgl_mask", "cfg1_sgl_kkt"]


@pytest.mark.parametrize("name", SGL)
def test_admm_sgl_matches_reference(golden, name):
    g = golden(name)
    S = g["S"]
    p = S.shape[0]
    kw = dict(ast.literal_eval(str(g["kw"]).replace("array", "").replace("\n", ""))) if "mask" not in name else \
        dict(tol=1e-7, rtol=1e-7)
    if "mask" in name:
        kw["lambda1_mask"] = g["lambda1_mask"]
    trace = []
    sol, info = orc.admm_sgl(S, float(g["lambda1"]), np.eye(p), measure=True, trace=trace, **kw)
    if kw.get("stopping_criterion", "boyd") == "boyd":
        _check_traj(trace, g, str(g["status"]), info)
    else:
        assert info["status"] == str(g["status"])
    np.testing.assert_allclose(info["residual"], g["residual"], rtol=1e-7, atol=1e-13)
    for k in ("Theta", "Omega", "X") + (("L",) if "L" in g.files else ()):
        assert np.linalg.norm(sol[k] - g[k]) <= RTOL * max(1.0, np.linalg.norm(g[k])), k
    assert np.array_equal(sol["Theta"] != 0, g["Theta"] != 0)
    assert ("L" in sol) == ("L" in g.files)


def test_block_sgl_matches_reference(golden):
    g = golden("block_sgl_p100")
    S, lam = g["S"], float(g["lambda1"])
    n, comps = orc.connected_components(S, lam)
    assert n == int(g["numC"]) and n > 1
    sol = orc.block_sgl(S, lam, np.eye(S.shape[0]), tol=1e-9, rtol=1e-9)
    for k in ("Theta", "Omega", "X"):
        assert np.linalg.norm(sol[k] - g[k]) <= RTOL * max(1.0, np.linalg.norm(g[k])), k
    assert np.abs(sol["Theta"] - g["Theta_full"]).max() < 1e-5


def test_mask_of_zeros_gives_inverse():
    # known-answer test of the reference (tests/test_solvers.py:191-246): zero penalty => Theta = inv(S)
    rng = np.random.default_rng(0)
    p = 20
    A = rng.standard_normal((4 * p, p))
    S = A.T @ A / (4 * p)
    sol, info = orc.admm_sgl(S, 0.1, np.eye(p), tol=1e-10, rtol=1e-10, lambda1_mask=np.zeros((p, p)))
    np.testing.assert_allclose(sol["Theta"], np.linalg.inv(S), atol=1e-4)


def test_fsgl_matches_reference(golden):
    g = golden("fsgl_p12_M3")
    S, M, lam = g["S"], int(g["M"]), float(g["lambda1"])
    assert _rel(orc.prox_sum_frob(g["prox_in"], M, 0.7), g["prox_out"]) < 1e-15
    for tag, lat in (("nolat", False), ("lat", True)):
        trace = []
        sol, info = orc.admm_fsgl(S, lam, M, np.eye(S.shape[0]), tol=1e-7, rtol=1e-7, latent=lat,
                                  mu1=0.2 if lat else None, trace=trace)
        traj = g[f"traj_{tag}"]
        assert info["status"] == str(g[f"status_{tag}"]) and len(trace) == traj.shape[0]
        assert np.array_equal(np.array([t["rho"] for t in trace]), traj[:, 0])
        np.testing.assert_allclose(info["residual"], g[f"residual_{tag}"], rtol=1e-7, atol=1e-13)
        for k in ("Theta", "Omega", "X"):
            assert np.linalg.norm(sol[k] - g[f"{k}_{tag}"]) <= RTOL * max(1.0, np.linalg.norm(g[f"{k}_{tag}"])), k
        assert np.array_equal(sol["Theta"] != 0, g[f"Theta_{tag}"] != 0)
    assert np.linalg.norm(sol["L"] - g["L_lat"]) <= RTOL * max(1.0, np.linalg.norm(g["L_lat"]))


def _ext_inputs(g):
    p = g["p"]
    S = {k: g[f"S{k}"] for k in range(len(p))}
    Om0 = {k: np.eye(int(p[k])) for k in range(len(p))}
    return S, Om0, g["G"], float(g["lambda1"]), float(g["lambda2"])


EXT_CASES = [("boyd", dict(tol=1e-7, rtol=1e-7)), ("latent", dict(tol=1e-7, rtol=1e-7, latent=True, mu1=0.3)),
             ("kkt", dict(tol=1e-4, stopping_criterion="kkt", max_iter=400))]


@pytest.mark.parametrize("tag,kw", EXT_CASES)
def test_ext_admm_matches_reference(golden, tag, kw):
    g = golden("ext_mgl_K3")
    S, Om0, G, l1, l2 = _ext_inputs(g)
    sol, info = orc.ext_admm_mgl(S, l1, l2, Om0, G, **kw)
    assert info["status"] == str(g[f"status_{tag}"])
    assert len(info["residual"]) == len(g[f"residual_{tag}"])
    np.testing.assert_allclose(info["residual"], g[f"residual_{tag}"], rtol=1e-6, atol=1e-13)
    for name in ("Omega", "Theta", "L", "X0", "X1"):
        for k in range(3):
            ref = g[f"{name}{k}_{tag}"]
            assert np.linalg.norm(sol[name][k] - ref) <= RTOL * max(1.0, np.linalg.norm(ref)), (name, k)
    for k in range(3):
        assert np.array_equal(sol["Theta"][k] != 0, g[f"Theta{k}_{tag}"] != 0)
