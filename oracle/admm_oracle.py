"""
oracle/admm_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU (numpy + plain C) restatement of the GGLasso ADMM hot path, used as the parity
checker for the CUDA path and as the timed CPU baseline in bench.py.  The product package
``gglasso_b200`` never imports this module; only tests/, __graft_entry__.smoke() and
bench.py (cpu_baseline leg / --impl reference) may.

It restates the *algorithm* of the reference (fabian-sp/GGLasso v0.2.1) -- it is not a copy.
Every function cites the reference lines it follows (paths relative to /root/reference):

  soft_threshold        src/gglasso/solver/ggl_helper.py:12-14    prox_1norm
  prox_od_1norm         src/gglasso/solver/ggl_helper.py:16-27
  prox_rank_norm        src/gglasso/solver/ggl_helper.py:29-36
  prox_p                src/gglasso/solver/ggl_helper.py:190-207  (+ prox_phi :178-187, prox_phi_ggl :68-71,
                                                                    prox_2norm :38-43, prox_phi_fgl :131-134)
  tv1d                  src/gglasso/solver/fgl_helper.py:11-68    condat_method
  phiplus               src/gglasso/solver/ggl_helper.py:272-303  phip / phiplus
  objective             src/gglasso/solver/ggl_helper.py:266-270  h / f ; :162-176 P_val ; basic_linalg.py:20-33 Gdot
  boyd_residuals        src/gglasso/solver/admm_solver.py:316-331 ; single_admm_solver.py:277-291
  kkt_residual_mgl      src/gglasso/solver/admm_solver.py:333-371
  kkt_residual_sgl      src/gglasso/solver/single_admm_solver.py:293-319
  admm_mgl              src/gglasso/solver/admm_solver.py:13-313
  admm_sgl              src/gglasso/solver/single_admm_solver.py:15-275
  block_sgl             src/gglasso/solver/single_admm_solver.py:326-475 (+ get_connected_components :478-490)
  prox_sum_frob         src/gglasso/solver/ggl_helper.py:45-66
  admm_fsgl             src/gglasso/solver/functional_sgl_admm.py:12-239
  prox_2norm_G          src/gglasso/solver/ext_admm_solver.py:394-453
  ext_admm_mgl          src/gglasso/solver/ext_admm_solver.py:18-392

Third-party arithmetic the reference delegates to (not under /root/reference): numpy.linalg.eigh
(LAPACK dsyevd), numpy matmul (dgemm), scipy.sparse.csgraph.connected_components.  The oracle
calls the same libraries.

Pinning: tests/test_oracle_golden.py checks this file against fixtures produced by the real
reference in the build container (tests/golden/make_golden.py): final solutions, iteration
counts, status strings and per-iteration (r, s, eps_pri, eps_dual, rho) trajectories.
"""
import ctypes
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CLIB = None


def _clib():
    """C helpers (oracle/gg_oracle.c); None if not built -> pure-numpy/Python fallbacks."""
    global _CLIB
    if _CLIB is None:
        path = os.path.join(_HERE, "_build", "libgg_oracle.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            dp = ctypes.POINTER(ctypes.c_double)
            lib.gg_tv1d.argtypes = [dp, ctypes.c_int, ctypes.c_double, dp]
            lib.gg_tv1d.restype = None
            for fn in (lib.gg_prox_p_fgl, lib.gg_prox_p_ggl):
                fn.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, dp]
                fn.restype = None
            lib.gg_pval.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int]
            lib.gg_pval.restype = ctypes.c_double
            _CLIB = lib
        else:
            _CLIB = False
    return _CLIB or None


def build_c():
    """Compile oracle/gg_oracle.c (called by __graft_entry__.build())."""
    import subprocess
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    global _CLIB
    _CLIB = None
    return _clib() is not None


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ----------------------------------------------------------------------------------------
# prox operators
# ----------------------------------------------------------------------------------------
def soft_threshold(v, l):
    return np.sign(v) * np.maximum(np.abs(v) - l, 0.0)


def prox_od_1norm(A, l):
    """soft-threshold every entry, then restore the diagonal; ``l`` scalar or (p,p)."""
    res = soft_threshold(A, l)
    idx = np.arange(min(A.shape))
    res[idx, idx] = A[idx, idx]
    return res


def prox_rank_norm(D, Q, beta):
    """Q diag(max(D-beta,0)) Q^T from a given eigendecomposition."""
    return (Q * np.maximum(D - beta, 0.0)) @ Q.T


def tv1d_py(y, lam):
    """Pure-Python direct 1-D TV prox (fallback for when the C helper is not built)."""
    n = len(y)
    x = np.zeros(n)
    k = k0 = kp = km = 0
    vmin, vmax = y[0] - lam, y[0] + lam
    umin, umax = lam, -lam
    while True:
        if k == n - 1:
            if umin < 0:
                x[k0:km + 1] = vmin
                km += 1
                k = k0 = km
                umin, vmin, umax = lam, y[k], y[k] + lam - vmax
            elif umax > 0:
                x[k0:kp + 1] = vmax
                kp += 1
                k = k0 = kp
                umax, vmax, umin = -lam, y[k], y[k] - lam - vmin
            else:
                x[k0:] = vmin + umin / (k - k0 + 1)
                return x
            if k == n - 1:
                x[k] = vmin + umin
                return x
            continue
        if y[k + 1] + umin - vmin < -lam:
            x[k0:km + 1] = vmin
            km += 1
            k = kp = k0 = km
            vmin, vmax = y[k], y[k] + 2 * lam
            umin, umax = lam, -lam
        elif y[k + 1] + umax - vmax > lam:
            x[k0:kp + 1] = vmax
            kp += 1
            k = km = k0 = kp
            vmin, vmax = y[k] - 2 * lam, y[k]
            umin, umax = lam, -lam
        else:
            k += 1
            umin = umin + y[k] - vmin
            umax = umax + y[k] - vmax
            if umin >= lam:
                vmin += (umin - lam) / (k - k0 + 1)
                umin = lam
                km = k
            if umax <= -lam:
                vmax += (umax + lam) / (k - k0 + 1)
                umax = -lam
                kp = k


def tv1d(y, lam):
    y = np.ascontiguousarray(y, dtype=np.float64)
    lib = _clib()
    if lib is None:
        return tv1d_py(y, lam)
    x = np.empty_like(y)
    lib.gg_tv1d(_dptr(y), len(y), float(lam), _dptr(x))
    return x


def prox_p(X, l1, l2, reg):
    """Per-entry prox across the K instances on the upper triangle, mirrored; diagonal kept."""
    assert np.abs(X - X.transpose(0, 2, 1)).max() <= 1e-5, "input X is not symmetric"
    assert min(l1, l2) > 0
    assert reg in ("GGL", "FGL")
    K, p, _ = X.shape
    X = np.ascontiguousarray(X, dtype=np.float64)
    lib = _clib()
    if lib is not None:
        M = np.empty_like(X)
        (lib.gg_prox_p_ggl if reg == "GGL" else lib.gg_prox_p_fgl)(_dptr(X), K, p, float(l1), float(l2), _dptr(M))
        return M
    iu, ju = np.triu_indices(p, 1)
    V = X[:, iu, ju]                                   # (K, n_entries)
    if reg == "GGL":
        U = soft_threshold(V, l1)
        nrm = np.sqrt((U * U).sum(0))
        a = np.maximum(nrm, l2)
        R = (U * (a - l2)) / a
    else:
        R = np.empty_like(V)
        for e in range(V.shape[1]):
            R[:, e] = soft_threshold(tv1d_py(V[:, e], l2), l1)
    M = np.zeros_like(X)
    M[:, iu, ju] = R
    M[:, ju, iu] = R
    d = np.arange(p)
    M[:, d, d] = X[:, d, d]
    return M


def phip(d, beta):
    return 0.5 * (np.sqrt(d ** 2 + 4 * beta) + d)


def phiplus(beta, D, Q):
    """prox of -beta*logdet from the eigendecomposition A = Q diag(D) Q^T."""
    return (Q * phip(D, beta)) @ Q.T


# ----------------------------------------------------------------------------------------
# objective and stopping criteria
# ----------------------------------------------------------------------------------------
def P_val(X, l1, l2, reg):
    K, p, _ = X.shape
    lib = _clib()
    if lib is not None:
        Xc = np.ascontiguousarray(X, dtype=np.float64)
        return lib.gg_pval(_dptr(Xc), K, p, float(l1), float(l2), 0 if reg == "GGL" else 1)
    iu, ju = np.triu_indices(p, 1)
    V = X[:, iu, ju]
    n1 = np.abs(V).sum(0)
    n2 = np.sqrt((V * V).sum(0)) if reg == "GGL" else np.abs(V[1:] - V[:-1]).sum(0)
    return 2 * float((l1 * n1 + l2 * n2).sum())


def objective(Omega, Theta, S, l1, l2, reg):
    """sum_k -log det Omega_k + <Omega,S> + P(Theta)   (unweighted by n_k, as the reference)."""
    val = (-np.log(np.linalg.det(Omega))).sum() + float(np.sum(Omega * S))
    return val + P_val(Theta, l1, l2, reg)


def boyd_residuals(Omega, Omega_prev, Theta, L, X, rho, eps_abs, eps_rel):
    """r, s, eps_pri, eps_dual with dim = K(p^2+p)/2 (K=1 for 2-D input)."""
    p = Omega.shape[-1]
    K = Omega.shape[0] if Omega.ndim == 3 else 1
    dim = K * ((p ** 2 + p) / 2)
    e_pri = dim * eps_abs + eps_rel * max(np.linalg.norm(Omega), np.linalg.norm(Theta - L))
    e_dual = dim * eps_abs + eps_rel * rho * np.linalg.norm(X)
    r = np.linalg.norm(Omega - Theta + L)
    s = rho * np.linalg.norm(Omega - Omega_prev)
    return r, s, e_pri, e_dual


def kkt_residual_mgl(Omega, Theta, L, X, S, lambda1, lambda2, nk, reg, latent=False, mu1=None):
    """X is the UNSCALED dual (rho * scaled dual)."""
    K = S.shape[0]
    nT = np.linalg.norm(Theta)
    t1 = np.linalg.norm(Theta - prox_p(Theta + X, lambda1, lambda2, reg)) / (1 + nT)
    t2 = np.linalg.norm(Theta - Omega - L) / (1 + nT)
    D, Q = np.linalg.eigh(Omega - nk * S - X)
    proxK = np.stack([phiplus(nk[k, 0, 0], D[k], Q[k]) for k in range(K)])
    t3 = np.linalg.norm(Omega - proxK) / (1 + np.linalg.norm(Omega))
    t4 = 0.0
    if latent:
        D, Q = np.linalg.eigh(L - X)
        proxL = np.stack([prox_rank_norm(D[k], Q[k], mu1[k]) for k in range(K)])
        t4 = np.linalg.norm(L - proxL) / (1 + np.linalg.norm(L))
    return max(t1, t2, t3, t4)


def kkt_residual_sgl(Omega, Theta, L, X, S, lambda1, latent=False, mu1=None):
    nT = np.linalg.norm(Theta)
    t1 = np.linalg.norm(Theta - prox_od_1norm(Theta + X, lambda1)) / (1 + nT)
    t2 = np.linalg.norm(Omega - Theta + L) / (1 + nT)
    D, Q = np.linalg.eigh(Omega - S - X)
    t3 = np.linalg.norm(Omega - phiplus(1, D, Q)) / (1 + np.linalg.norm(Omega))
    t4 = 0.0
    if latent:
        D, Q = np.linalg.eigh(L - X)
        t4 = np.linalg.norm(L - prox_rank_norm(D, Q, mu1)) / (1 + np.linalg.norm(L))
    return max(t1, t2, t3, t4)


def _rho_update(r, s, rho):
    if r >= 10 * s:
        return 2 * rho
    if s >= 10 * r:
        return 0.5 * rho
    return 1.0 * rho


def _final_status(status, crit, r, s, e_pri, e_dual):
    if status == "optimal":
        return status
    if crit == "boyd":
        if r <= e_pri:
            return "primal optimal"
        if s <= e_dual:
            return "dual optimal"
    return "max iterations reached"


# ----------------------------------------------------------------------------------------
# solvers
# ----------------------------------------------------------------------------------------
def admm_mgl(S, lambda1, lambda2, reg, Omega_0, Theta_0=None, X_0=None, n_samples=None,
             tol=1e-5, rtol=1e-4, stopping_criterion="boyd", update_rho=True, rho=1.0,
             max_iter=1000, measure=False, latent=False, mu1=None, trace=None, quiet=True):
    """ADMM for the conforming multiple graphical lasso; returns (sol, info).

    ``trace``: optional list; one dict per iteration with copies of the state *before* the
    rho-rescaling of X (same point at which the reference calls its stopping criterion).
    """
    assert Omega_0.shape == S.shape and S.shape[1] == S.shape[2]
    assert reg in ("GGL", "FGL") and min(lambda1, lambda2) > 0 and rho > 0
    K, p, _ = S.shape
    if latent:
        if isinstance(mu1, float):
            mu1 = mu1 * np.ones(K)
        assert mu1 is not None and np.all(mu1 > 0)
    nk = np.ones((K, 1, 1)) if n_samples is None else n_samples * np.ones((K, 1, 1))

    Omega = Omega_0.copy()
    Theta = Omega_0.copy() if Theta_0 is None or len(Theta_0) == 0 else Theta_0.copy()
    X = np.zeros((K, p, p)) if X_0 is None or len(X_0) == 0 else X_0.copy()
    L = np.zeros((K, p, p))
    runtime = np.zeros(max_iter)
    residual = np.zeros(max_iter)
    obj = np.zeros(max_iter)
    status = ""
    r = s = e_pri = e_dual = np.nan

    for it in range(max_iter):
        t0 = time.time()
        Omega_prev = Omega.copy()
        W = Theta - L - X - (nk / rho) * S
        D, Q = np.linalg.eigh(W)
        for k in range(K):
            Omega[k] = phiplus(nk[k, 0, 0] / rho, D[k], Q[k])
        Theta = prox_p(Omega + L + X, (1 / rho) * lambda1, (1 / rho) * lambda2, reg)
        if latent:
            C = Theta - X - Omega
            D, Q = np.linalg.eigh(C)
            for k in range(K):
                L[k] = prox_rank_norm(D[k], Q[k], mu1[k] / rho)
        X = X + (Omega - Theta + L)
        if measure:
            runtime[it] = time.time() - t0
            obj[it] = objective(Omega, Theta, S, lambda1, lambda2, reg)

        if stopping_criterion == "boyd":
            r, s, e_pri, e_dual = boyd_residuals(Omega, Omega_prev, Theta, L, X, rho, tol, rtol)
            if trace is not None:
                trace.append(dict(Omega=Omega.copy(), Theta=Theta.copy(), L=L.copy(), X=X.copy(),
                                  rho=rho, r=r, s=s, e_pri=e_pri, e_dual=e_dual))
            if update_rho:
                rho_new = _rho_update(r, s, rho)
                X = (rho / rho_new) * X
                rho = rho_new
            residual[it] = max(r, s)
            if r <= e_pri and s <= e_dual:
                status = "optimal"
                break
        else:
            eta = kkt_residual_mgl(Omega, Theta, L, rho * X, S, lambda1, lambda2, nk, reg, latent, mu1)
            if trace is not None:
                trace.append(dict(Omega=Omega.copy(), Theta=Theta.copy(), L=L.copy(), X=X.copy(), rho=rho, eta=eta))
            residual[it] = eta
            if eta <= tol:
                status = "optimal"
                break

    status = _final_status(status, stopping_criterion, r, s, e_pri, e_dual)
    if not quiet:
        print(f"ADMM terminated after {it + 1} iterations with status: {status}.")
    sol = {"Omega": Omega, "Theta": Theta, "L": L, "X": X}
    info = {"status": status, "iterations": it + 1, "rho": rho}
    if measure:
        info.update(runtime=runtime[:it + 1], residual=residual[:it + 1], objective=obj[:it + 1])
    return sol, info


def admm_sgl(S, lambda1, Omega_0, Theta_0=None, X_0=None, rho=1.0, max_iter=1000, tol=1e-7,
             rtol=1e-4, stopping_criterion="boyd", update_rho=True, measure=False, latent=False,
             mu1=None, lambda1_mask=None, trace=None, quiet=True):
    """ADMM for the single graphical lasso; returns (sol, info)."""
    assert Omega_0.shape == S.shape and S.shape[0] == S.shape[1]
    assert lambda1 > 0 and rho > 0
    p = S.shape[0]
    lam = lambda1 * lambda1_mask if lambda1_mask is not None else lambda1
    if latent:
        assert mu1 is not None and mu1 > 0

    Omega = Omega_0.copy()
    Theta = Omega_0.copy() if Theta_0 is None or len(Theta_0) == 0 else Theta_0.copy()
    X = np.zeros((p, p)) if X_0 is None or len(X_0) == 0 else X_0.copy()
    L = np.zeros((p, p))
    runtime = np.zeros(max_iter)
    residual = np.zeros(max_iter)
    status = ""
    r = s = e_pri = e_dual = np.nan

    for it in range(max_iter):
        t0 = time.time()
        W = Theta - L - X - (1 / rho) * S
        D, Q = np.linalg.eigh(W)
        Omega_prev = Omega.copy()
        Omega = phiplus(1 / rho, D, Q)
        Theta = prox_od_1norm(Omega + L + X, (1 / rho) * lam)
        if latent:
            C = Theta - X - Omega
            D1, Q1 = np.linalg.eigh(C)
            L = prox_rank_norm(D1, Q1, mu1 / rho)
        X = X + Omega - Theta + L
        if measure:
            runtime[it] = time.time() - t0

        if stopping_criterion == "boyd":
            r, s, e_pri, e_dual = boyd_residuals(Omega, Omega_prev, Theta, L, X, rho, tol, rtol)
            if trace is not None:
                trace.append(dict(Omega=Omega.copy(), Theta=Theta.copy(), L=L.copy(), X=X.copy(),
                                  rho=rho, r=r, s=s, e_pri=e_pri, e_dual=e_dual))
            if update_rho:
                rho_new = _rho_update(r, s, rho)
                X = (rho / rho_new) * X
                rho = rho_new
            residual[it] = max(r, s)
            if r <= e_pri and s <= e_dual:
                status = "optimal"
                break
        else:
            eta = kkt_residual_sgl(Omega, Theta, L, rho * X, S, lam, latent, mu1)
            if trace is not None:
                trace.append(dict(Omega=Omega.copy(), Theta=Theta.copy(), L=L.copy(), X=X.copy(), rho=rho, eta=eta))
            residual[it] = eta
            if eta <= tol:
                status = "optimal"
                break

    status = _final_status(status, stopping_criterion, r, s, e_pri, e_dual)
    if not quiet:
        print(f"ADMM terminated after {it + 1} iterations with status: {status}.")
    sol = {"Omega": Omega, "Theta": Theta, "X": X}
    if latent:
        sol["L"] = L
    info = {"status": status, "iterations": it + 1, "rho": rho}
    if measure:
        info.update(runtime=runtime[:it + 1], residual=residual[:it + 1])
    return sol, info


def connected_components(S, lam):
    """components of the graph |S| > lam (diagonal forced on); list of index arrays, label order."""
    from scipy.sparse.csgraph import connected_components as cc
    A = (np.abs(S) > lam).astype(int)
    np.fill_diagonal(A, 1)
    n, labels = cc(A, directed=False, return_labels=True)
    return n, [np.flatnonzero(labels == i) for i in range(n)]


def block_sgl(S, lambda1, Omega_0, Theta_0=None, X_0=None, rho=1.0, max_iter=1000, tol=1e-7,
              rtol=1e-3, stopping_criterion="boyd", update_rho=True, lambda1_mask=None):
    """Witten/Friedman/Simon screening: solve each connected component separately, reassemble."""
    p = S.shape[0]
    mask = np.ones((p, p)) if lambda1_mask is None else lambda1_mask
    Theta_0 = Omega_0.copy() if Theta_0 is None else Theta_0
    X_0 = np.zeros((p, p)) if X_0 is None else X_0
    n, comps = connected_components(S, lambda1 * mask)
    out = {k: np.zeros((p, p)) for k in ("Omega", "Theta", "X")}
    for C in comps:
        ix = np.ix_(C, C)
        if len(C) == 1:
            v = 1 / S[C, C]
            out["Omega"][ix] = v
            out["Theta"][ix] = v
        else:
            bs, _ = admm_sgl(S[ix], lambda1, Omega_0[ix], Theta_0[ix], X_0[ix], rho=rho, max_iter=max_iter,
                             tol=tol, rtol=rtol, stopping_criterion=stopping_criterion,
                             update_rho=update_rho, lambda1_mask=mask[ix])
            for k in out:
                out[k][ix] = bs[k]
    return out


# ----------------------------------------------------------------------------------------
# functional SGL
# ----------------------------------------------------------------------------------------
def prox_sum_frob(X, M, l):
    """off-diagonal MxM blocks shrunk in Frobenius norm (from the upper block, mirrored); diagonal blocks kept."""
    pM = X.shape[0]
    assert pM % M == 0
    p = pM // M
    Y = np.zeros((pM, pM))
    for i in range(p):
        for j in range(i, p):
            blk = X[i * M:(i + 1) * M, j * M:(j + 1) * M]
            if i == j:
                Y[i * M:(i + 1) * M, j * M:(j + 1) * M] = blk
            else:
                a = max(np.linalg.norm(blk), l)
                B = blk * (a - l) / a
                Y[i * M:(i + 1) * M, j * M:(j + 1) * M] = B
                Y[j * M:(j + 1) * M, i * M:(i + 1) * M] = B.T
    return Y


def admm_fsgl(S, lambda1, M, Omega_0, Theta_0=None, X_0=None, rho=1.0, max_iter=1000, tol=1e-7, rtol=1e-4,
              update_rho=True, measure=False, latent=False, mu1=None, trace=None):
    """ADMM for the functional single graphical lasso; returns (sol, info)."""
    assert Omega_0.shape == S.shape and S.shape[0] == S.shape[1] and lambda1 > 0 and rho > 0
    pM = S.shape[0]
    assert pM % M == 0
    Omega = Omega_0.copy()
    Theta = Omega_0.copy() if Theta_0 is None or len(Theta_0) == 0 else Theta_0.copy()
    X = np.zeros((pM, pM)) if X_0 is None or len(X_0) == 0 else X_0.copy()
    L = np.zeros((pM, pM))
    residual = np.zeros(max_iter)
    status = ""
    for it in range(max_iter):
        W = Theta - L - X - (1 / rho) * S
        D, Q = np.linalg.eigh(W)
        Omega_prev = Omega.copy()
        Omega = phiplus(1 / rho, D, Q)
        Theta = prox_sum_frob(Omega + L + X, M, (1 / rho) * lambda1)
        if latent:
            C = Theta - X - Omega
            D1, Q1 = np.linalg.eigh(C)
            L = prox_rank_norm(D1, Q1, mu1 / rho)
        X = X + Omega - Theta + L
        r, s, e_pri, e_dual = boyd_residuals(Omega, Omega_prev, Theta, L, X, rho, tol, rtol)
        if trace is not None:
            trace.append(dict(rho=rho, r=r, s=s, e_pri=e_pri, e_dual=e_dual))
        if update_rho:
            rho_new = _rho_update(r, s, rho)
            X = (rho / rho_new) * X
            rho = rho_new
        residual[it] = max(r, s)
        if r <= e_pri and s <= e_dual:
            status = "optimal"
            break
    status = _final_status(status, "boyd", r, s, e_pri, e_dual)
    sol = {"Omega": Omega, "Theta": Theta, "X": X}
    if latent:
        sol["L"] = L
    return sol, {"status": status, "iterations": it + 1, "residual": residual[:it + 1], "rho": rho}


# ----------------------------------------------------------------------------------------
# non-conforming group graphical lasso (ext_ADMM_MGL)
# ----------------------------------------------------------------------------------------
def prox_2norm_G(X, G, l2):
    """group prox through the bookkeeping array G (2,L,K): per group l the member entries X[k][G0,G1] (G != -1)
    are shrunk in Euclidean norm with threshold l2*sqrt(group size) and written back symmetrically."""
    K = len(X)
    out = [X[k].copy() for k in range(K)]
    for l in range(G.shape[1]):
        ks = [k for k in range(K) if G[0, l, k] != -1]
        v = np.array([out[k][G[0, l, k], G[1, l, k]] for k in ks])
        lam = l2 * np.sqrt(len(ks))
        a = max(np.sqrt((v ** 2).sum()), lam)
        z = v * (a - lam) / a
        for k, zk in zip(ks, z):
            out[k][G[0, l, k], G[1, l, k]] = zk
            out[k][G[1, l, k], G[0, l, k]] = zk
    return dict(enumerate(out))


def ext_admm_mgl(S, lambda1, lambda2, Omega_0, G, X0=None, X1=None, tol=1e-5, rtol=1e-4, stopping_criterion="boyd",
                 rho=1.0, max_iter=1000, latent=False, mu1=None):
    """ADMM for the group graphical lasso with non-conforming dimensions (dicts keyed 0..K-1)."""
    K = len(S)
    p = np.array([S[k].shape[0] for k in range(K)])
    lam1 = lambda1 * np.ones(K) if np.isscalar(lambda1) else np.asarray(lambda1, dtype=float)
    if latent:
        mu = mu1 * np.ones(K) if np.isscalar(mu1) else np.asarray(mu1, dtype=float)
    Omega = {k: Omega_0[k].copy() for k in range(K)}
    Theta = {k: Omega_0[k].copy() for k in range(K)}
    Lam = {k: Omega_0[k].copy() for k in range(K)}
    L = {k: np.zeros((p[k], p[k])) for k in range(K)}
    X0 = {k: np.zeros((p[k], p[k])) for k in range(K)} if X0 is None else {k: X0[k].copy() for k in range(K)}
    X1 = {k: np.zeros((p[k], p[k])) for k in range(K)} if X1 is None else {k: X1[k].copy() for k in range(K)}
    residual = np.zeros(max_iter)
    status = ""
    dim = ((p ** 2 + p) / 2).sum()
    r = s = e_pri = e_dual = np.nan
    for it in range(max_iter):
        Omega_prev = {k: Omega[k].copy() for k in range(K)}
        for k in range(K):
            W = Theta[k] - L[k] - X0[k] - (1 / rho) * S[k]
            D, Q = np.linalg.eigh(W)
            Omega[k] = phiplus(1 / rho, D, Q)
        for k in range(K):
            V = (Omega[k] + L[k] + X0[k] + Lam[k] - X1[k]) * 0.5
            Theta[k] = prox_od_1norm(V, lam1[k] / (2 * rho))
        if latent:
            for k in range(K):
                C = Theta[k] - X0[k] - Omega[k]
                C = (C.T + C) / 2
                D, Q = np.linalg.eigh(C)
                L[k] = prox_rank_norm(D, Q, mu[k] / rho)
        Lam_prev = Lam
        Lam = prox_2norm_G({k: Theta[k] + X1[k] for k in range(K)}, G, lambda2 / rho)
        for k in range(K):
            X0[k] = X0[k] + (Omega[k] - Theta[k] + L[k])
            X1[k] = X1[k] + (Theta[k] - Lam[k])
        if stopping_criterion == "boyd":
            nr = np.linalg.norm
            D1 = np.sqrt(sum(nr(Omega[k]) ** 2 + nr(Lam[k]) ** 2 for k in range(K)))
            D2 = np.sqrt(sum(nr(Theta[k] - L[k]) ** 2 + nr(Theta[k]) ** 2 for k in range(K)))
            D3 = np.sqrt(sum(nr(X0[k]) ** 2 + nr(X1[k]) ** 2 for k in range(K)))
            e_pri = dim * tol + rtol * max(D1, D2)
            e_dual = dim * tol + rtol * rho * D3
            r = np.sqrt(sum(nr(Omega[k] - Theta[k] + L[k]) ** 2 + nr(Lam[k] - Theta[k]) ** 2 for k in range(K)))
            s = rho * np.sqrt(sum(nr(Omega[k] - Omega_prev[k]) ** 2 + nr(Lam[k] - Lam_prev[k]) ** 2 for k in range(K)))
            residual[it] = max(r, s)
            if r <= e_pri and s <= e_dual:
                status = "optimal"
                break
        else:
            eta = _ext_kkt(Omega, Theta, L, Lam, {k: rho * X0[k] for k in range(K)}, {k: rho * X1[k] for k in range(K)},
                           S, G, lam1, lambda2, latent, mu if latent else None)
            residual[it] = eta
            if eta <= tol:
                status = "optimal"
                break
    status = _final_status(status, stopping_criterion, r, s, e_pri, e_dual)
    sol = {"Omega": Omega, "Theta": Theta, "L": L, "X0": X0, "X1": X1}
    return sol, {"status": status, "iterations": it + 1, "residual": residual[:it + 1]}


def _ext_kkt(Omega, Theta, L, Lam, X0, X1, S, G, lam1, lambda2, latent, mu):
    K = len(S)
    nr = np.linalg.norm
    t = np.zeros((6, K))
    for k in range(K):
        D, Q = np.linalg.eigh(Omega[k] - S[k] - X0[k])
        t[0, k] = nr(Omega[k] - phiplus(1, D, Q)) / (1 + nr(Omega[k]))
        t[1, k] = nr(Theta[k] - prox_od_1norm(Theta[k] + X0[k] - X1[k], lam1[k])) / (1 + nr(Theta[k]))
        if latent:
            D, Q = np.linalg.eigh(L[k] - X0[k])
            t[2, k] = nr(L[k] - prox_rank_norm(D, Q, mu[k])) / (1 + nr(L[k]))
        t[4, k] = nr(Omega[k] - Theta[k] + L[k]) / (1 + nr(Theta[k]))
        t[5, k] = nr(Lam[k] - Theta[k]) / (1 + nr(Theta[k]))
    V = prox_2norm_G({k: Lam[k] + X1[k] for k in range(K)}, G, lambda2)
    for k in range(K):
        t[3, k] = nr(V[k] - Lam[k]) / (1 + nr(Lam[k]))
    return max(nr(t[i]) for i in range(6))
