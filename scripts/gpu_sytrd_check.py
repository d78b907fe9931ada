"""GPU check + timing of the blocked tridiagonalisation (gg_sytrd_blocked.cuh) against LAPACK and the per-column path.
usage: python scripts/gpu_sytrd_check.py [out.json]"""
import ctypes, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gglasso_b200 import _lib
from gglasso_b200._engine import Eigh, to_dev, _p

lib = _lib.load()
dev = torch.device("cuda")
out = {"checks": [], "timing": []}


def sym(rng, M, p):
    A = rng.standard_normal((M, p, p))
    return (A + A.transpose(0, 2, 1)) / 2


def check(M, p, env):
    for k, v in env.items():
        os.environ[k] = str(v)
    rng = np.random.default_rng(p * 7 + M)
    A = sym(rng, M, p)
    At = to_dev(A, dev)
    e = Eigh(M, p, dev)
    D = e.eigh(At, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    D = D.cpu().numpy(); Vt = At.cpu().numpy()
    Dref = np.linalg.eigvalsh(A)
    derr = float(np.abs(np.sort(D, 1) - Dref).max())
    orth = max(float(np.abs(Vt[m] @ Vt[m].T - np.eye(p)).max()) for m in range(M))
    res = max(float(np.abs(Vt[m] @ A[m] - D[m][:, None] * Vt[m]).max()) for m in range(M))
    rec = {"M": M, "p": p, "env": env, "eig_err": derr, "orth": orth, "resid": res}
    print(rec, flush=True)
    out["checks"].append(rec)
    for k in env:
        os.environ.pop(k, None)


def timeit(M, p, env, reps=5):
    for k, v in env.items():
        os.environ[k] = str(v)
    rng = np.random.default_rng(1)
    A = to_dev(sym(rng, M, p), dev)
    e = Eigh(M, p, dev)
    st = torch.cuda.current_stream().cuda_stream
    rec = {"M": M, "p": p, "env": env}
    for name, fn in (("eigh", lambda W: e.eigh(W, stream=st)),
                     ("sytrd", lambda W: lib.gg_sytrd_profile(_p(W), _p(e.D), M, p, _p(e.ws), e.ws_bytes, 0, st)),
                     ("panels", lambda W: lib.gg_sytrd_profile(_p(W), _p(e.D), M, p, _p(e.ws), e.ws_bytes, 1, st)),
                     ("syr2k", lambda W: lib.gg_sytrd_profile(_p(W), _p(e.D), M, p, _p(e.ws), e.ws_bytes, 2, st))):
        best = 1e9
        for r in range(reps):
            W = A.clone()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(W); b.record()
            torch.cuda.synchronize()
            if r:
                best = min(best, a.elapsed_time(b))
        rec[name + "_ms"] = best
    print(rec, flush=True)
    out["timing"].append(rec)
    for k in env:
        os.environ.pop(k, None)


if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sytrd_check.json"
    old = os.environ.get("GG_TR_OLD", "0") == "1"
    sizes = [(2, 145), (3, 200), (2, 333), (20, 500), (5, 777), (20, 1000), (1, 1289), (2, 2047)]
    if not old:
        for M, p in sizes:
            check(M, p, {})
        check(3, 400, {"GG_TR_NB": 32})
        check(3, 400, {"GG_TR_CS": 4})
        check(3, 400, {"GG_TR_CS": 8, "GG_TR_NB": 32})
        check(2, 1000, {"GG_TR_CS": 16})
        check(20, 1000, {"GG_TR_NB": 32})
    else:
        check(20, 1000, {})
    for M, p, env in ([(20, 1000, {}), (10, 500, {}), (3, 1000, {}), (2, 1000, {}), (1, 1289, {})] if old else
                      [(20, 1000, {}), (20, 1000, {"GG_TR_NB": 32}), (20, 1000, {"GG_TR_CS": 4}), (10, 500, {}),
                       (10, 500, {"GG_TR_NB": 32}), (3, 1000, {}), (2, 1000, {}), (5, 1000, {}), (1, 1289, {}), (40, 1000, {})]):
        timeit(M, p, env)
    json.dump(out, open(path, "w"), indent=1)
