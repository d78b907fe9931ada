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
