"""ADMM_FSGL on B200 -- drop-in for gglasso.solver.functional_sgl_admm.ADMM_FSGL
(src/gglasso/solver/functional_sgl_admm.py:12-239): functional single graphical lasso, i.e. the SGL loop with
the block-Frobenius prox ``prox_sum_Frob`` (src/gglasso/solver/ggl_helper.py:45-66).  Same signature, asserts,
printed lines, warnings and return dicts; the iteration runs as CUDA kernels (gg_prox_fsgl + the shared loop)."""
import warnings
from typing import Optional

import numpy as np

from .._engine import run_admm, to_host


def ADMM_FSGL(S: np.ndarray,
              lambda1: float,
              M: int,
              Omega_0: np.ndarray,
              Theta_0: np.ndarray = np.array([]),
              X_0: np.ndarray = np.array([]),
              rho: float = 1.,
              max_iter: int = 1000,
              tol: float = 1e-7,
              rtol: float = 1e-4,
              update_rho: bool = True,
              verbose: bool = False,
              measure: bool = False,
              latent: bool = False,
              mu1: Optional[float] = None
              ):
    """(Latent variable) Functional Single Graphical Lasso by ADMM on (p*M, p*M) inputs; see the reference
    docstring for the model.  Returns ``(sol, info)`` with keys Omega, Theta, X (and L iff ``latent``)."""
    assert Omega_0.shape == S.shape
    assert S.shape[0] == S.shape[1]
    assert lambda1 > 0

    mu = None
    if latent:
        assert mu1 is not None
        assert mu1 > 0
        mu = np.array([float(mu1)])

    (pM, pM) = S.shape
    assert pM % M == 0
    p = int(pM / M)

    if verbose:
        print(f"Derived a Functional SGL problem of dimensionality p={p}.")

    assert rho > 0, "ADMM penalization parameter must be positive."

    if len(Theta_0) == 0:
        Theta_0 = None
    if len(X_0) == 0:
        X_0 = None

    st, res = run_admm('sgl', S, Omega_0, Theta_0, X_0, lambda1=float(lambda1), rho=float(rho),
                       max_iter=int(max_iter), tol=tol, rtol=rtol, update_rho=update_rho, verbose=verbose,
                       measure=measure, latent=latent, mu=mu, Mblk=int(M), print_rho=True,
                       header="------------ADMM Algorithm for Functional Single Graphical Lasso----------------")
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
    dmin = st.min_eig(TL)
    if dmin <= 0:
        warnings.warn(f"Theta (Theta - L resp.) is not positive definite. Solve to higher accuracy! (min EV is {dmin})")
    if latent:
        dmin = st.min_eig(st.L)
        if dmin < -1e-8:
            warnings.warn(f"L is not positive semidefinite. Solve to higher accuracy! (min EV is {dmin})")

    sol = {'Omega': to_host(Omega_d[0]), 'Theta': to_host(st.Theta[0]), 'X': to_host(st.X[0])}
    if latent:
        sol['L'] = to_host(st.L[0])
    if measure:
        info = {'status': status, 'runtime': res["runtime"][:n_it], 'residual': res["residual"][0]}
    else:
        info = {'status': status}
    return sol, info
