"""ADMM_SGL / block_SGL on B200 -- drop-ins for gglasso.solver.single_admm_solver.

Same signatures, defaults, asserts, printed lines, return dicts and status strings as the reference
(src/gglasso/solver/single_admm_solver.py:15-275 and :326-498); the iteration runs as CUDA kernels.
"""
import contextlib
import io
import warnings
from typing import Optional

import numpy as np
from scipy.sparse.csgraph import connected_components

from .._engine import run_admm, to_host


def ADMM_SGL(S: np.ndarray,
             lambda1: float,
             Omega_0: np.ndarray,
             Theta_0: np.ndarray = np.array([]),
             X_0: np.ndarray = np.array([]),
             rho: float = 1.,
             max_iter: int = 1000,
             tol: float = 1e-7,
             rtol: float = 1e-4,
             stopping_criterion: str = 'boyd',
             update_rho: bool = True,
             verbose: bool = False,
             measure: bool = False,
             latent: bool = False,
             mu1: Optional[float] = None,
             lambda1_mask: Optional[np.ndarray] = None
             ):
    """(Latent variable) Single Graphical Lasso by ADMM; see the reference docstring for the model.

    Returns ``(sol, info)``; ``sol`` has keys Omega, Theta, X (and L iff ``latent``), all (p,p).
    """
    assert Omega_0.shape == S.shape
    assert S.shape[0] == S.shape[1]

    (p, p) = S.shape

    assert lambda1 > 0, "lambda1 should be positive, otherwise using Graphical Lasso is redundant. Specify entries with zero regularization using lambda1_mask."

    lam_mat = None
    if lambda1_mask is not None:
        assert lambda1_mask.shape == (p, p), f"lambda1_mask needs to be of shape (p,p), but is {lambda1_mask.shape}."
        assert np.all(lambda1_mask >= 0), "lambda1_mask needs to be non-negative."
        assert np.all(np.abs(lambda1_mask.T - lambda1_mask) <= 1e-5), "lambda1_mask needs to be symmetric."
        lam_mat = (lambda1 * lambda1_mask)[None]

    assert stopping_criterion in ["boyd", "kkt"]

    mu = None
    if latent:
        assert mu1 is not None
        assert mu1 > 0
        mu = np.array([float(mu1)])

    assert rho > 0, "ADMM penalization parameter must be positive."

    if len(Theta_0) == 0:
        Theta_0 = None          # device-side default: copy of Omega_0
    if len(X_0) == 0:
        X_0 = None              # device-side default: zeros

    st, res = run_admm('sgl', S, Omega_0, Theta_0, X_0, lambda1=float(lambda1), lam_mat=lam_mat, rho=float(rho),
                       max_iter=int(max_iter), tol=tol, rtol=rtol, stopping_criterion=stopping_criterion,
                       update_rho=update_rho, verbose=verbose, measure=measure, latent=latent, mu=mu,
                       header="------------ADMM Algorithm for Single Graphical Lasso----------------")
    n_it = int(res["iters"][0])
    status = res["status"][0]
    print(f"ADMM terminated after {n_it} iterations with status: {status}.")

    Omega_d = st.final_omega(res["iters"])
    ### CHECK FOR SYMMETRY
    for name, A in (("Omega", Omega_d), ("Theta", st.Theta), ("L", st.L)):
        if A is None:
            continue
        dev_max = st.asym_max(A)
        if dev_max > 1e-5:
            warnings.warn(f"{name} variable is not symmetric, largest deviation is {dev_max}.")

    ### CHECK FOR POSDEF
    TL = st.Theta - st.L if latent else st.Theta
    ok, dmin = st.is_posdef(TL, res)
    if not ok:
        print(f"WARNING: Theta (Theta - L resp.) is not positive definite. Solve to higher accuracy! (min EV is {dmin})")
    if latent:
        dmin = st.min_eig(st.L)
        if dmin < -1e-8:
            print(f"WARNING: L is not positive semidefinite. Solve to higher accuracy! (min EV is {dmin})")

    sol = {'Omega': to_host(Omega_d[0]), 'Theta': to_host(st.Theta[0]), 'X': to_host(st.X[0])}
    if latent:
        sol['L'] = to_host(st.L[0])

    if measure:
        info = {'status': status, 'runtime': res["runtime"][:n_it], 'residual': res["residual"][0]}
    else:
        info = {'status': status}
    return sol, info


#######################################################
## BLOCK-WISE GRAPHICAL LASSO AFTER WITTEN ET AL.
#######################################################

def block_SGL(S: np.ndarray,
              lambda1: float,
              Omega_0: np.ndarray,
              Theta_0: Optional[np.ndarray] = None,
              X_0: Optional[np.ndarray] = None,
              rho: float = 1.,
              max_iter: int = 1000,
              tol: float = 1e-7,
              rtol: float = 1e-3,
              stopping_criterion: str = "boyd",
              update_rho: bool = True,
              verbose: bool = False,
              measure: bool = False,
              lambda1_mask: Optional[np.ndarray] = None
              ):
    """Solve the SGL problem on each connected component of ``|S| > lambda1*mask`` (Witten, Friedman, Simon)
    and reassemble.  Returns ``sol`` only (keys Omega, Theta, X), like the reference.
    """
    assert Omega_0.shape == S.shape
    assert S.shape[0] == S.shape[1]

    (p, p) = S.shape

    assert lambda1 > 0, "lambda1 should be positive, otherwise using Graphical Lasso is redundant. Specify entries with zero regularization using lambda1_mask."

    if lambda1_mask is not None:
        assert lambda1_mask.shape == (p, p), f"lambda1_mask needs to be of shape (p,p), but is {lambda1_mask.shape}."
        assert np.all(lambda1_mask >= 0), "lambda1_mask needs to be non-negative."
        assert np.all(np.abs(lambda1_mask.T - lambda1_mask) <= 1e-5), "lambda1_mask needs to be symmetric."
    else:
        lambda1_mask = np.ones((p, p))

    if Theta_0 is None:
        Theta_0 = Omega_0.copy()
    if X_0 is None:
        X_0 = np.zeros((p, p))

    numC, allC = get_connected_components(S, lambda1 * lambda1_mask)
    kw = dict(tol=tol, rtol=rtol, stopping_criterion=stopping_criterion, update_rho=update_rho, rho=rho,
              max_iter=max_iter, verbose=verbose, measure=measure)
    return _solve_components(S, lambda1, lambda1_mask, Omega_0, Theta_0, X_0, allC, range(numC), kw)


def _solve_components(S, lambda1, lambda1_mask, Omega_0, Theta_0, X_0, allC, mine, kw, solver=None):
    """solve the components with indices ``mine`` (all of them for block_SGL, this rank's share for
    parallel.block_SGL_dist) and scatter them into (p,p) arrays that are zero elsewhere."""
    p = S.shape[0]
    mine = set(mine)
    stopping_criterion, verbose, measure = kw["stopping_criterion"], kw["verbose"], kw["measure"]
    sol = {'Omega': np.zeros((p, p)), 'Theta': np.zeros((p, p)), 'X': np.zeros((p, p))}
    lines = {}
    multi = [ci for ci, C in enumerate(allC) if len(C) > 1 and ci in mine]
    for ci, C in enumerate(allC):
        if ci not in mine:
            continue
        if len(C) == 1:
            # single node components have a closed form solution (off-diagonal penalty only)
            closed_sol = 1 / S[C, C]
            sol['Omega'][C, C] = closed_sol
            sol['Theta'][C, C] = closed_sol

    # Components that fit the shared-memory eigensolver are solved as ragged batches (one CTA per block and
    # kernel, per-block rho / stopping test on the device); larger ones one by one on the large-p path.
    batchable = stopping_criterion == "boyd" and not verbose and not measure and solver is None
    small = [ci for ci in multi if len(allC[ci]) <= _BATCH_MAX] if batchable else []
    large = [ci for ci in multi if ci not in set(small)]
    for bucket in _size_buckets([len(allC[ci]) for ci in small]):
        # a batch index travels in gridDim.y / .z (limit 65535) and every problem keeps a (max_iter, 5) residual
        # history on the device: batches are cut so that both stay bounded
        step = max(1, min(_BATCH_PROBLEMS_MAX, (256 << 20) // (40 * int(kw["max_iter"]))))
        for lo in range(0, len(bucket), step):
            ids = [small[k] for k in bucket[lo:lo + step]]
            _solve_batch(S, lambda1, lambda1_mask, Omega_0, Theta_0, X_0, [allC[ci] for ci in ids], ids, sol, lines, kw)
    for ci in large:
        C = allC[ci]
        ix = np.ix_(C, C)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            block_sol, _ = (solver or ADMM_SGL)(S=np.ascontiguousarray(S[ix]),
                                    lambda1=lambda1,
                                    Omega_0=np.ascontiguousarray(Omega_0[ix]),
                                    Theta_0=np.ascontiguousarray(Theta_0[ix]),
                                    X_0=np.ascontiguousarray(X_0[ix]),
                                    lambda1_mask=np.ascontiguousarray(lambda1_mask[ix]), **kw)
        lines[ci] = buf.getvalue()
        for k in sol:
            sol[k][ix] = block_sol[k]
    # the reference prints one termination line per non-trivial block, in component order
    for ci in multi:
        print(lines[ci], end="")
    return sol


_BATCH_MAX = 160          # GG_SMALL_MAX of the eigensolver
_BATCH_PROBLEMS_MAX = 32768   # problems per ragged batch (grid dimension limit 65535)


def _size_buckets(sizes):
    """group indices of blocks by padded size: powers of two up to 32, then steps of 32 (bounded padding waste)."""
    edges = [2, 4, 8, 16, 32, 64, 96, 128, 160]
    buckets = {}
    for k, s in enumerate(sizes):
        e = next(x for x in edges if s <= x)
        buckets.setdefault(e, []).append(k)
    return [buckets[e] for e in sorted(buckets)]


def _solve_batch(S, lambda1, mask, Omega_0, Theta_0, X_0, comps, ids, sol, lines, kw):
    """one ragged batch: blocks padded to the largest size with decoupled unit diagonal entries (S=Omega=Theta=1,
    X=0 there: exact fixed points of the iteration for every rho, excluded from the norms through pvec)."""
    M = len(comps)
    pm = max(len(C) for C in comps)
    eye = np.eye(pm)
    Sb = np.repeat(eye[None], M, 0)
    Ob, Tb, Xb = Sb.copy(), Sb.copy(), np.zeros((M, pm, pm))
    Lb = np.zeros((M, pm, pm))
    for b, C in enumerate(comps):
        n = len(C)
        ix = np.ix_(C, C)
        Sb[b, :n, :n] = S[ix]
        Ob[b, :n, :n] = Omega_0[ix]
        Tb[b, :n, :n] = Theta_0[ix]
        Xb[b, :n, :n] = X_0[ix]
        Lb[b, :n, :n] = lambda1 * mask[ix]
    st, res = run_admm('sgl', Sb, Ob, Tb, Xb, lambda1=float(lambda1), lam_mat=Lb, rho=float(kw["rho"]),
                       max_iter=int(kw["max_iter"]), tol=kw["tol"], rtol=kw["rtol"], update_rho=kw["update_rho"],
                       pvec=[len(C) for C in comps])
    Om = to_host(st.final_omega(res["iters"]))
    Th, Xs = to_host(st.Theta), to_host(st.X)
    for b, (C, ci) in enumerate(zip(comps, ids)):
        n = len(C)
        ix = np.ix_(C, C)
        sol['Omega'][ix] = Om[b, :n, :n]
        sol['Theta'][ix] = Th[b, :n, :n]
        sol['X'][ix] = Xs[b, :n, :n]
        lines[ci] = f"ADMM terminated after {int(res['iters'][b])} iterations with status: {res['status'][b]}.\n"


def get_connected_components(S, lambda1):
    """connected components of the graph with adjacency |S| > lambda1 (self loops added);
    host side, as in the reference (single_admm_solver.py:478-490)."""
    A = (np.abs(S) > lambda1).astype(int)
    np.fill_diagonal(A, 1)
    numC, labelsC = connected_components(A, directed=False, return_labels=True)
    allC = [np.flatnonzero(labelsC == i) for i in range(numC)]
    return numC, allC


def invert_permutation(p):
    """s with s[p[i]] = i for a permutation p of 0..len(p)-1."""
    s = np.empty_like(p)
    s[p] = np.arange(p.size)
    return s
