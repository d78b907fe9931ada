"""BASELINE.json configurations at FULL size on the GPU, on inputs from the reference's own seeded generators
(oracle/ref_inputs.py -> oracle/_ref) against fixtures produced by the REAL reference in the build container
(tests/golden/make_golden_large.py).  north_star bar: Theta/Omega/X within 1e-8 relative Frobenius per iteration,
final objective within 1e-6 relative, identical sparsity pattern at tol=1e-7.

The inputs are regenerated here (they are too large to commit) with the host-independent sampler of
oracle/ref_inputs.py; their fingerprints are compared with the ones recorded next to the fixtures, so a host whose
LAPACK rounds differently shows up as a reported deviation, not as a silent one.  (With the reference's default SVD
sampler the survey's container run gave objective 11752.409197493562 / nnz 48024 at cfg3; the Cholesky-sampled input
of the same generators and seeds gives 11752.787900092353 / 47854 -- tests/golden/cfg3_fgl_full.npz.)
"""
import contextlib
import io
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from oracle import ref, ref_inputs  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device"),
              pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (oracle/make_ref.sh)")]

HERE = os.path.dirname(os.path.abspath(__file__))
FP_JSON = os.path.join(HERE, "golden", "large_inputs.json")
TOL = 1e-8


def _quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        out = fn(*a, **kw)
    return out, buf.getvalue()


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _input(name):
    S = ref_inputs.load(name)
    dev = ref_inputs.check_fingerprint(name, S, FP_JSON)
    assert dev < 1e-9, f"{name}: regenerated input deviates from the build container's by {dev:.2e}"
    return S


def _dense(shape, idx, val):
    out = np.zeros(int(np.prod(shape)))
    out[idx] = val
    return out.reshape(shape)


def test_cfg3_full_solve_vs_reference_trajectory(golden):
    """cfg3 (FGL, K=20, p=1000, N=2000, lambda1=.05, lambda2=.01, tol=rtol=1e-7): every iteration of the real
    reference's run -- rho path, r/s/eps, norms of Omega/Theta/X, nnz(Theta), 4096 sampled entries of each array --
    then the public call: status, iteration count, objective per iteration, final Theta (values and exact pattern)."""
    from gglasso_b200 import ADMM_MGL
    from gglasso_b200._engine import run_admm
    g = golden("cfg3_fgl_full")
    S = _input("cfg3")
    K, p, _ = S.shape
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    idx = torch.from_numpy(g["sample_index"]).cuda()
    rows = []

    def grab(st, it):
        rows.append({k: getattr(st, k).reshape(-1)[idx].cpu().numpy() for k in ("Omega", "Theta", "X")}
                    | {"n" + k: float(torch.linalg.norm(getattr(st, k)).item()) for k in ("Omega", "Theta", "X")}
                    | {"nnz": int(torch.count_nonzero(st.Theta).item())})

    st, res = run_admm("mgl", S, Om0, None, None, lambda1=0.05, lambda2=0.01, reg="FGL", tol=1e-7, rtol=1e-7,
                       trace=grab, check_every=1)
    traj = g["traj"]
    n = traj.shape[0]
    assert int(res["iters"][0]) == n == len(rows) and res["status"][0] == str(g["status"])
    h = res["hist"][0][:n]
    assert np.array_equal(h[:, 4], traj[:, 0]), "rho sequence differs from the reference"
    np.testing.assert_allclose(h[:, :4], traj[:, 1:5], rtol=1e-7, atol=1e-12)
    for t, row in enumerate(rows):
        for j, k in ((5, "Omega"), (6, "Theta"), (8, "X")):
            assert abs(row["n" + k] - traj[t, j]) <= TOL * traj[t, j], (t, k, "norm")
            ref_s = g[k + "_s"][t]
            assert np.linalg.norm(row[k] - ref_s) <= TOL * max(np.linalg.norm(ref_s), 1e-3), (t, k, "sampled entries")
        assert row["nnz"] == int(traj[t, 9]), (t, "nnz(Theta)")

    (sol, info), out = _quiet(ADMM_MGL, S, 0.05, 0.01, "FGL", Om0, tol=1e-7, rtol=1e-7, measure=True)
    assert info["status"] == str(g["status"]) and f"ADMM terminated after {n} iterations" in out
    np.testing.assert_allclose(info["objective"], g["objective"], rtol=1e-6)
    assert abs(info["objective"][-1] - float(g["objective"][-1])) <= 1e-6 * abs(float(g["objective"][-1]))
    np.testing.assert_allclose(info["residual"], g["residual"], rtol=1e-7, atol=1e-12)
    Theta_ref = _dense(S.shape, g["theta_idx"], g["theta_val"])
    assert np.array_equal(np.flatnonzero(sol["Theta"].reshape(-1)), g["theta_idx"]), "sparsity pattern differs"
    assert _rel(sol["Theta"], Theta_ref) < TOL
    sidx = g["sample_index"]
    for k, key in (("Omega", "Omega_final_s"), ("X", "X_final_s")):
        assert np.linalg.norm(sol[k].reshape(-1)[sidx] - g[key]) <= TOL * max(np.linalg.norm(g[key]), 1e-3), k
    for i, k in enumerate(("Omega", "Theta", "X")):
        assert abs(np.linalg.norm(sol[k]) - g["norms"][i]) <= TOL * g["norms"][i], k
    assert np.array_equal(sol["Theta"], sol["Theta"].transpose(0, 2, 1))


def test_cfg3_first_iterations_vs_reference_live():
    """the same input through the real reference (oracle/_ref) on this host's cores, two iterations: full arrays."""
    from gglasso_b200 import ADMM_MGL
    ADMM_MGL_ref, _, _ = ref.fresh_solvers()
    S = _input("cfg3")
    K, p, _ = S.shape
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    kw = dict(tol=1e-7, rtol=1e-7, max_iter=2)
    (sol, info), _ = _quiet(ADMM_MGL, S, 0.05, 0.01, "FGL", Om0, **kw)
    (rsol, rinfo), _ = _quiet(ADMM_MGL_ref, S, 0.05, 0.01, "FGL", Om0, **kw)
    assert info["status"] == rinfo["status"] == "max iterations reached"
    for k in ("Omega", "Theta", "X"):
        assert _rel(sol[k], rsol[k]) < TOL, k
    assert np.array_equal(sol["Theta"] != 0, rsol["Theta"] != 0)


@pytest.mark.parametrize("lam", [0.1, 0.05])
def test_cfg5_block_sgl_vs_reference(golden, lam):
    """cfg5: block_SGL on the p=5000 power-law input (N=5500): component count, exact pattern, Theta to 1e-8."""
    from gglasso_b200 import block_SGL, get_connected_components
    g = golden("cfg5_block_sgl")
    tag = str(lam).replace(".", "")
    S = _input("cfg5")
    p = S.shape[0]
    numC, allC = get_connected_components(S, lam)
    assert numC == int(g[f"numC_{tag}"])
    assert np.array_equal(np.sort([len(c) for c in allC])[::-1], g[f"sizes_{tag}"])
    sol, _ = _quiet(block_SGL, S, lam, np.eye(p), tol=1e-7, rtol=1e-7)
    assert set(sol) == {"Omega", "Theta", "X"}
    Theta_ref = _dense(S.shape, g[f"theta_idx_{tag}"], g[f"theta_val_{tag}"])
    assert np.array_equal(np.flatnonzero(sol["Theta"].reshape(-1)), g[f"theta_idx_{tag}"]), "sparsity pattern differs"
    assert _rel(sol["Theta"], Theta_ref) < TOL
    assert abs(np.linalg.norm(sol["Omega"]) - g[f"omega_norm_{tag}"]) <= TOL * g[f"omega_norm_{tag}"]
    assert abs(np.linalg.norm(sol["X"]) - g[f"x_norm_{tag}"]) <= TOL * max(g[f"x_norm_{tag}"], 1.0)


@pytest.mark.parametrize("tag", ["4x3", "10x10"])
def test_cfg4_model_selection_through_the_reference_facade(golden, tag):
    """cfg4: glasso_problem(S, N, reg='GGL').model_selection(...) of the UNMODIFIED reference with the B200 solvers
    installed underneath (same warm-start chain as the reference): BIC/AIC/SP tables, best (lambda1, lambda2), Theta."""
    import gglasso_b200
    if not os.path.isfile(os.path.join(HERE, "golden", "cfg4_grid.npz")):
        pytest.skip("cfg4_grid.npz not generated")
    g = golden("cfg4_grid")
    if f"bic_{tag}" not in g.files:
        pytest.skip(f"cfg4 {tag} grid not in the fixture")
    S = _input("cfg4")
    N = ref_inputs.n_samples("cfg4")
    ref.load()
    from gglasso.problem import glasso_problem
    gglasso_b200.install()
    try:
        P = glasso_problem(S, N, reg="GGL", latent=False, do_scaling=False)
        _quiet(P.model_selection, modelselect_params={"lambda1_range": g[f"l1_{tag}"], "lambda2_range": g[f"l2_{tag}"]},
               method="eBIC", gamma=0.1, tol=1e-7, rtol=1e-7)
    finally:
        gglasso_b200.uninstall()
    st = P.modelselect_stats
    np.testing.assert_allclose(st["BIC"][0.1], g[f"bic_{tag}"], rtol=1e-7)
    np.testing.assert_allclose(st["AIC"], g[f"aic_{tag}"], rtol=1e-7)
    assert np.array_equal(st["SP"], g[f"sp_{tag}"]), "sparsity table differs"
    assert st["BEST"]["lambda1"] == float(g[f"best_l1_{tag}"]) and st["BEST"]["lambda2"] == float(g[f"best_l2_{tag}"])
    Theta_ref = _dense(S.shape, g[f"theta_idx_{tag}"], g[f"theta_val_{tag}"])
    Theta = P.solution.precision_
    assert np.array_equal(Theta != 0, Theta_ref != 0)
    assert _rel(Theta, Theta_ref) < TOL


def test_cfg4_device_grid_agrees_with_reference_optimum(golden):
    """the device-resident, column-sharded grid (own warm-start chains; start points differ from the reference's
    chain, the optimum does not): same best (lambda1, lambda2) and eBIC table to solver tolerance."""
    from gglasso_b200.parallel import grid_search_device
    if not os.path.isfile(os.path.join(HERE, "golden", "cfg4_grid.npz")):
        pytest.skip("cfg4_grid.npz not generated")
    g = golden("cfg4_grid")
    if "bic_10x10" not in g.files:
        pytest.skip("10x10 grid not in the fixture")
    S = _input("cfg4")
    N = ref_inputs.n_samples("cfg4")
    l1, l2 = g["l1_10x10"], g["l2_10x10"]
    scores, iters, ix, best = grid_search_device(S, np.full(S.shape[0], N), "GGL", l1, l2, gamma=0.1, tol=1e-7,
                                                 rtol=1e-7, n_streams=5)
    assert tuple(int(i) for i in ix) == tuple(int(i) for i in g["ix_10x10"])
    # the eBIC counts non-zeros: at the dense end of the grid a few entries near the threshold differ between the two
    # chains of start points at tol=1e-7 (measured: 8 of 100 points, <= 1.0e-3 relative); the rest agrees to 1e-5
    dev = np.abs(scores - g["bic_10x10"]) / np.abs(g["bic_10x10"])
    assert dev.max() < 5e-3 and np.mean(dev < 1e-5) >= 0.85, (dev.max(), np.mean(dev < 1e-5))
