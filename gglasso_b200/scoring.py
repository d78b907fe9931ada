"""Model-selection scores on the device.

Device-resident statements of the reference's scoring helpers, so that a grid search can keep Theta / L on the GPU
between solves:

    ebic_single / ebic_array   src/gglasso/helper/model_selection.py:841-869   (+ lambda1_mask variant :846-852)
    aic_single / aic_array     :813-839
    robust_logdet              :884-894   (-inf when the smallest eigenvalue is <= 1e-12)
    sparsity / mean_sparsity   src/gglasso/helper/utils.py:17-31
    np.linalg.matrix_rank(L)   model_selection.py:253 (grid_search), :677 (single_grid_search)
    thresholding / tune_threshold / tune_multiple_threshold   :697-760

Eigenvalues come from the CUDA eigensolver (`gg_eigh`, values only); the remaining terms are reductions over arrays that
are already resident (torch ops on the device: glue between solves, not part of the ADMM iteration).
All functions take torch CUDA tensors (float64) and return Python floats / numpy arrays.
"""
import numpy as np
import torch

from ._engine import Eigh

N_TAU = 20          # model_selection.py:19


def _stack(A):
    return A if A.ndim == 3 else A[None]


def eigvalsh(A, eig=None):
    """eigenvalues (K,p) of a (K,p,p) / (p,p) symmetric stack; A is not modified."""
    A3 = _stack(A)
    K, p, _ = A3.shape
    e = eig if (eig is not None and eig.M == K and eig.p == p) else Eigh(K, p, A3.device)
    W = A3.clone()
    D = e.eigh(W, vectors=0, stream=torch.cuda.current_stream().cuda_stream)
    return D.clone()


def robust_logdet(A, t=1e-12, eig=None):
    """per-matrix log det, -inf where the smallest eigenvalue is <= t (model_selection.py:884-894)."""
    D = eigvalsh(A, eig)
    dmin = D.min(dim=1).values
    ld = torch.log(D.clamp_min(1e-300)).sum(1)
    return torch.where(dmin > t, ld, torch.full_like(ld, -float("inf")))


def _edges(Theta3, lambda1_mask=None):
    K, p, _ = Theta3.shape
    if lambda1_mask is None:
        return (torch.count_nonzero(Theta3.reshape(K, -1), dim=1).to(torch.float64) - p) / 2
    m = torch.as_tensor(lambda1_mask, dtype=torch.float64, device=Theta3.device)
    E = (Theta3 != 0).to(torch.float64) * m
    E = E - torch.diag_embed(torch.diagonal(E, dim1=1, dim2=2))
    return E.sum((1, 2)) / 2


def _nvec(N, K, dev):
    return torch.as_tensor(np.broadcast_to(np.asarray(N, dtype=np.float64), (K,)).copy(), device=dev)


def ebic(S, Theta, N, gamma=0.5, lambda1_mask=None, eig=None):
    """extended BIC summed over the instances of a stack (ebic_single / ebic_array)."""
    S3, T3 = _stack(S), _stack(Theta)
    K, p, _ = S3.shape
    Nd = _nvec(N, K, S3.device)
    inner = (S3 * T3).sum((1, 2))
    val = Nd * inner - Nd * robust_logdet(T3, eig=eig) + _edges(T3, lambda1_mask) * (torch.log(Nd) + 4 * np.log(p) * gamma)
    return float(val.sum().item())


def aic(S, Theta, N, eig=None):
    """AIC summed over the instances of a stack (aic_single / aic_array)."""
    S3, T3 = _stack(S), _stack(Theta)
    K, p, _ = S3.shape
    Nd = _nvec(N, K, S3.device)
    inner = (S3 * T3).sum((1, 2))
    val = Nd * inner - Nd * robust_logdet(T3, eig=eig) + _edges(T3)
    return float(val.sum().item())


def mean_sparsity(Theta):
    """mean off-diagonal ratio of non-zero entries (utils.py:17-31)."""
    T3 = _stack(Theta)
    K, p, _ = T3.shape
    off = torch.count_nonzero(T3.reshape(K, -1), dim=1).to(torch.float64) - p
    return float((off / (p * p - p)).mean().item())


def matrix_rank(L, eig=None):
    """np.linalg.matrix_rank of each (symmetric) matrix of a stack: #singular values > max(sv) * p * eps; for a
    symmetric matrix the singular values are the absolute eigenvalues."""
    D = eigvalsh(L, eig).abs()
    p = D.shape[1]
    tol = D.max(dim=1, keepdim=True).values * p * np.finfo(np.float64).eps
    return (D > tol).sum(1).cpu().numpy()


def thresholding(A, tau):
    """A * (|A| > tau) with the diagonal kept (model_selection.py:697-705)."""
    mask = A.abs() > tau
    eye = torch.eye(A.shape[-1], dtype=torch.bool, device=A.device)
    return A * (mask | eye)


def tune_threshold(Theta, S, N, tau_range=None, method="eBIC", gamma=0.1):
    """best threshold for one (p,p) matrix: all candidates are scored as ONE stack (one batched eigenvalue call
    instead of len(tau_range) host eigvalsh calls).  Returns (thresholded Theta, tau, scores)."""
    if tau_range is None:
        tau_range = np.logspace(-12, -1, N_TAU)
    tau_range = np.asarray(tau_range, dtype=np.float64)
    assert np.all(tau_range > 0)
    p = Theta.shape[-1]
    taus = torch.as_tensor(tau_range, device=Theta.device).reshape(-1, 1, 1)
    eye = torch.eye(p, dtype=torch.bool, device=Theta.device)
    cand = Theta[None] * ((Theta[None].abs() > taus) | eye)                   # (n_tau, p, p)
    n = cand.shape[0]
    ld = robust_logdet(cand)
    inner = (S[None] * cand).sum((1, 2))
    E = _edges(cand)
    pen = E * (np.log(N) + 4 * np.log(p) * gamma) if method == "eBIC" else E
    scores = (N * inner - N * ld + pen).cpu().numpy()
    scores[scores == np.inf] = np.nan
    ix = int(np.nanargmin(scores))
    return cand[ix].clone(), float(tau_range[ix]), scores


def tune_multiple_threshold(Theta, S, N, tau_range=None, method="eBIC", gamma=0.1):
    """per-instance thresholds of a (K,p,p) stack (model_selection.py:738-760)."""
    K = Theta.shape[0]
    Nv = np.broadcast_to(np.asarray(N, dtype=np.float64), (K,))
    out = Theta.clone()
    tau = np.zeros(K)
    score = {}
    for k in range(K):
        out[k], tau[k], score[k] = tune_threshold(Theta[k], S[k], float(Nv[k]), tau_range, method, gamma)
    return out, tau, score


DEFAULT_GAMMAS = [0.1, 0.3, 0.5, 0.7]      # model_selection.py:17


def single_grid_search_device(S, lambda_range, N, method="eBIC", gamma=0.3, latent=False, mu_range=None,
                              thresholding_=False, store_all=True, tol=1e-7, rtol=1e-7, lambda1_mask=None):
    """Device-resident statement of the reference's ``single_grid_search(..., use_block=False)``
    (src/gglasso/helper/model_selection.py:505-690): the lambda1 (x mu1) path of the single graphical lasso with S,
    the warm starts and every scored Theta / L staying on the GPU; scores (eBIC for all default gammas, AIC), sparsity,
    rank of L and the optional threshold tuning come from this module.  Same loop order, same start points
    (Omega_0 <- previous Omega, X_0 = identity at every point), same return layout:
    ``(best_sol, estimates, lowrank, stats)`` with numpy arrays.  (The block-wise route of the reference,
    ``use_block=True``, stays on the host path: ``install()`` + the reference's own driver.)"""
    from ._engine import run_admm, require_cuda, to_dev, to_host_many
    dev = require_cuda()
    S = np.asarray(S, dtype=np.float64)
    p = S.shape[0]
    if latent:
        assert mu_range is not None
        mu_range = np.asarray(mu_range, dtype=np.float64)
    else:
        mu_range = np.array([0])
    lambda_range = np.asarray(lambda_range, dtype=np.float64)
    _L, _M = len(lambda_range), len(mu_range)
    gammas = sorted(set(DEFAULT_GAMMAS + [gamma]))
    MU, LAMB = np.meshgrid(mu_range, lambda_range)
    BIC = {g: np.nan * np.zeros((_L, _M)) for g in gammas}
    AIC = np.nan * np.zeros((_L, _M))
    SP = np.nan * np.zeros((_L, _M))
    RANK = np.zeros((_L, _M))
    TAU = np.zeros((_L, _M)) if thresholding_ else None
    estimates = np.zeros((_L, _M, p, p)) if store_all else None
    lowrank = np.zeros((_L, _M, p, p)) if store_all else None
    S_dev = to_dev(S, dev)[None] if S.nbytes >= (8 << 20) else torch.from_numpy(S).to(dev)[None]
    eye = torch.eye(p, dtype=torch.float64, device=dev)[None]
    mask_dev = None if lambda1_mask is None else torch.as_tensor(np.asarray(lambda1_mask, dtype=np.float64), device=dev)
    Omega_0 = eye
    eig = Eigh(1, p, dev)
    best, curr_min = None, np.inf
    for j in range(_L):
        lam_mat = None if mask_dev is None else (float(lambda_range[j]) * mask_dev)[None].contiguous()
        for m in range(_M):
            st, res = run_admm("sgl", S_dev, Omega_0, None, eye, lambda1=float(lambda_range[j]), lam_mat=lam_mat,
                               tol=tol, rtol=rtol, latent=latent, mu=np.array([mu_range[m]]) if latent else None)
            Omega_0 = st.final_omega(res["iters"])
            Theta = st.Theta[0]
            if latent:
                RANK[j, m] = matrix_rank(st.L, eig)[0]
                if store_all:
                    lowrank[j, m] = st.L[0].cpu().numpy()
            if thresholding_:
                Theta, TAU[j, m], _ = tune_threshold(Theta, S_dev[0], N, None, method, gamma)
            AIC[j, m] = aic(S_dev[0], Theta, N, eig=eig)
            for g in gammas:
                BIC[g][j, m] = ebic(S_dev[0], Theta, N, g, lambda1_mask=lambda1_mask, eig=eig)
            SP[j, m] = mean_sparsity(Theta)
            if store_all:
                estimates[j, m] = Theta.cpu().numpy()
            score = BIC[gamma][j, m] if method == "eBIC" else AIC[j, m]
            if score < curr_min:
                curr_min = score
                best = {"Omega": Omega_0[0].clone(), "Theta": Theta.clone(), "X": st.X[0].clone(),
                        "L": st.L[0].clone() if latent else None}
    AIC[AIC == -np.inf] = np.nan
    for g in gammas:
        BIC[g][BIC[g] == -np.inf] = np.nan
    table = AIC if method == "AIC" else BIC[gamma]
    ix = np.unravel_index(np.nanargmin(table), table.shape)
    stats = {"BIC": BIC, "AIC": AIC, "SP": SP, "RANK": RANK, "LAMBDA": LAMB, "MU": MU, "TAU": TAU,
             "BEST": {"lambda1": LAMB[ix], "mu1": MU[ix]}, "GAMMA": gammas}
    best_sol = {}
    if best is not None:
        keys = [k for k in ("Omega", "Theta", "X", "L") if best[k] is not None]
        for k, a in zip(keys, to_host_many([best[k] for k in keys])):
            best_sol[k] = a
    return best_sol, estimates, lowrank, stats
