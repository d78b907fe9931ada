"""ADMM_MGL on B200 -- drop-in for gglasso.solver.admm_solver.ADMM_MGL.

Same signature, defaults, asserts, printed termination line, return dicts and status strings as
the reference (src/gglasso/solver/admm_solver.py:13-313); the iteration itself runs as CUDA
kernels (see gglasso_b200/_engine.py).  numpy in, numpy out.
"""
import warnings
from typing import Optional, Union

import numpy as np

from .._engine import run_admm, to_host


def ADMM_MGL(S: np.ndarray,
             lambda1: float,
             lambda2: float,
             reg: str,
             Omega_0: np.ndarray,
             Theta_0: np.ndarray = np.array([]),
             X_0: np.ndarray = np.array([]),
             n_samples: Optional[Union[int, np.ndarray]] = None,
             tol: float = 1e-5,
             rtol: float = 1e-4,
             stopping_criterion: str = 'boyd',
             update_rho: bool = True,
             rho: float = 1.,
             max_iter: int = 1000,
             verbose: bool = False,
             measure: bool = False,
             latent: bool = False,
             mu1: Optional[Union[float, np.ndarray]] = None
             ):
    """(Latent variable) Multiple Graphical Lasso by ADMM; see the reference docstring for the model.

    Returns ``(sol, info)``: ``sol`` has keys Omega, Theta, L, X of shape (K,p,p); ``info['status']`` is one of
    'optimal', 'primal optimal', 'dual optimal', 'max iterations reached'; with ``measure=True`` also
    'runtime', 'residual', 'objective' (length = number of iterations).
    """
    assert Omega_0.shape == S.shape
    assert S.shape[1] == S.shape[2]
    assert reg in ['GGL', 'FGL']
    assert min(lambda1, lambda2) > 0
    assert stopping_criterion in ['boyd', 'kkt']

    (K, p, p) = S.shape

    assert rho > 0, "ADMM penalization parameter must be positive."

    mu = None
    if latent:
        if isinstance(mu1, float):
            mu1 = mu1 * np.ones(K)
        assert mu1 is not None
        assert np.all(mu1 > 0)
        mu = np.asarray(mu1, dtype=np.float64).reshape(K)

    # n_samples None -> weights 1; int -> same weight for all k (admm_solver.py:133-139)
    if n_samples is None:
        nk = None
    elif isinstance(n_samples, (int, np.integer)):
        nk = float(n_samples) * np.ones(K)
    else:
        nk = np.asarray(n_samples, dtype=np.float64).reshape(K)
        assert len(nk) == K

    if len(Theta_0) == 0:
        Theta_0 = None          # device-side default: copy of Omega_0
    if len(X_0) == 0:
        X_0 = None              # device-side default: zeros
    # prox_p asserts symmetry of its input (ggl_helper.py:193); with symmetric S and start points the
    # iterates stay symmetric, so the check is done once on the inputs -- on the device, after the upload
    # (run_admm(check_symmetric=True)), not with host-side temporaries.

    st, res = run_admm('mgl', S, Omega_0, Theta_0, X_0, lambda1=float(lambda1), lambda2=float(lambda2), reg=reg,
                       rho=float(rho), max_iter=int(max_iter), tol=tol, rtol=rtol,
                       stopping_criterion=stopping_criterion, update_rho=update_rho, verbose=verbose,
                       measure=measure, latent=latent, mu=mu, nk=nk, check_symmetric=True,
                       header="------------ADMM Algorithm for Multiple Graphical Lasso----------------")
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

    ### CHECK FOR POSDEF (eigenvalue computations are enqueued first and overlap the D2H copies of the solution)
    TL = st.Theta - st.L if latent else st.Theta
    loop_done = st.mark()
    pd_min = st.posdef_async(TL, res)
    psd_min = st.posdef_async(st.L) if latent else None
    outs = st.to_host_overlapped([Omega_d, st.Theta, st.X] + ([st.L] if latent else []), after=loop_done)
    if pd_min is not None and float(pd_min.item()) <= 0:
        print("WARNING: Theta (Theta - L resp.) is not positive definite. Solve to higher accuracy!")
    if latent:
        if float(psd_min.item()) < -1e-5:
            print("WARNING: L is not positive semidefinite. Solve to higher accuracy!")

    sol = {'Omega': outs[0], 'Theta': outs[1], 'L': outs[3] if latent else np.zeros((K, p, p)), 'X': outs[2]}
    if measure:
        info = {'status': status,
                'runtime': res["runtime"][:n_it],
                'residual': res["residual"][0],
                'objective': res["objective"][:n_it]}
    else:
        info = {'status': status}
    return sol, info
