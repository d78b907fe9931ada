"""Synthetic inputs for tests and bench.py (host side, numpy only).

The reference's generators (src/gglasso/helper/data_generation.py) need networkx and live in the
reference tree, which is not available on the GPU box; this module produces inputs of the same
*kind* -- block-diagonal sparse precision matrices whose blocks are random preferential-attachment
trees (power-law degrees), K related instances, and the biased sample covariance of N Gaussian
draws -- with its own construction.  It is an input generator, not part of the solver.
"""
import numpy as np


def powerlaw_precision(p, n_blocks=10, seed=0):
    """sparse SPD precision matrix: ``n_blocks`` diagonal blocks, each a weighted random tree."""
    rng = np.random.default_rng(seed)
    L = p // n_blocks
    assert L * n_blocks == p and L >= 2
    A = np.zeros((p, p))
    for b in range(n_blocks):
        deg = np.ones(L)
        o = b * L
        for v in range(1, L):
            # preferential attachment -> heavy tailed degree distribution
            u = rng.choice(v, p=deg[:v] / deg[:v].sum())
            w = rng.uniform(0.1, 0.4) * rng.choice([-1.0, 1.0])
            A[o + u, o + v] = A[o + v, o + u] = w
            deg[u] += 1
            deg[v] += 1
    rs = 1.5 * np.abs(A).sum(1) + 1e-10
    A = A / rs[:, None]
    A = 0.5 * (A + A.T) + np.eye(p)
    dmin = np.linalg.eigvalsh(A).min()      # hubs can break diagonal dominance after symmetrisation
    if dmin < 0.05:
        A += (0.1 + abs(dmin)) * np.eye(p)
    return A


def instance_precisions(K, p, n_blocks=10, seed=0, kind="group"):
    """K related precision matrices.  kind='group': one random block switched off per instance;
    kind='fused': blocks switch at half time and one block decays along k (time-varying network)."""
    rng = np.random.default_rng(seed + 1)
    base = powerlaw_precision(p, n_blocks, seed)
    L = p // n_blocks
    out = np.repeat(base[None], K, 0)
    for k in range(K):
        if kind == "group":
            b = rng.integers(n_blocks) if K > 1 else -1
        else:
            b = 1 if k <= K / 2 else 0
            sl = slice(2 * L, 3 * L)
            blk = out[k, sl, sl]
            d = np.diag(blk).copy()
            blk *= np.exp(-0.5 * k)
            np.fill_diagonal(blk, d)
        if b >= 0:
            sl = slice(b * L, (b + 1) * L)
            out[k, sl, sl] = np.eye(L)
    return out


def sample_cov(Theta, N, seed=0):
    """biased sample covariance of N draws from N(0, Theta^{-1}) for each instance."""
    rng = np.random.default_rng(seed + 2)
    Theta3 = Theta if Theta.ndim == 3 else Theta[None]
    out = np.empty_like(Theta3)
    for k in range(Theta3.shape[0]):
        p = Theta3.shape[1]
        # x = L^{-T} z has covariance (L L^T)^{-1} = Theta^{-1}
        Lc = np.linalg.cholesky(Theta3[k])
        Z = rng.standard_normal((p, N))
        Xs = np.linalg.solve(Lc.T, Z)
        Xs -= Xs.mean(1, keepdims=True)
        out[k] = Xs @ Xs.T / N
    return out if Theta.ndim == 3 else out[0]


def synthetic_mgl(K, p, N=None, seed=0, kind="group", n_blocks=None):
    """(K,p,p) stack of sample covariance matrices for a group / fused MGL problem."""
    if n_blocks is None:
        n_blocks = next(b for b in (10, 8, 5, 4, 3, 2, 1) if p % b == 0 and p // b >= 2)
    N = N or 2 * p
    return sample_cov(instance_precisions(K, p, n_blocks, seed, kind), N, seed)


def synthetic_sgl(p, N=None, seed=0, n_blocks=None):
    if n_blocks is None:
        n_blocks = next(b for b in (10, 8, 5, 4, 3, 2, 1) if p % b == 0 and p // b >= 2)
    N = N or int(1.1 * p)
    return sample_cov(powerlaw_precision(p, n_blocks, seed), N, seed)
