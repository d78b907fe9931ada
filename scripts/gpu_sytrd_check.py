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
            check(M, p, {"GG_TR_BLOCKED": 1})                    # every column below the tail on the blocked path
        check(20, 1000, {})                                       # default: per-column chain with L2 hints + tail
        check(40, 1000, {})
        check(3, 2500, {})
        check(20, 1000, {"GG_TR_SWITCH_MB": 24})                  # blocked part + chain + tail
        check(3, 400, {"GG_TR_NB": 32, "GG_TR_BLOCKED": 1})
        check(3, 400, {"GG_TR_CS": 4, "GG_TR_BLOCKED": 1})
        check(3, 400, {"GG_TR_CS": 8, "GG_TR_NB": 32, "GG_TR_SWITCH_MB": 1})
        check(2, 1000, {"GG_TR_CS": 16, "GG_TR_SWITCH_MB": 4})
        check(20, 1000, {"GG_TR_NB": 32, "GG_TR_SWITCH_MB": 30})
    else:
        check(20, 1000, {})
    for M, p, env in ([(20, 1000, {}), (10, 500, {}), (3, 1000, {}), (2, 1000, {}), (1, 1289, {})] if old else
                      [(20, 1000, {}), (20, 1000, {"GG_TR_L2MB": 0}), (20, 1000, {"GG_TR_L2MB": 32}), (20, 1000, {"GG_TR_L2MB": 48}),
                       (20, 1000, {"GG_TR_L2MB": 80}), (20, 1000, {"GG_TR_L2MB": 96}), (20, 1000, {"GG_TR_SWITCH_MB": 48}),
                       (20, 1000, {"GG_TR_BLOCKED": 1}), (10, 500, {}), (3, 1000, {}),
                       (40, 1000, {}), (40, 1000, {"GG_TR_L2MB": 0}), (40, 1000, {"GG_TR_L2MB": 96}),
                       (10, 1000, {}), (10, 1000, {"GG_TR_L2MB": 0}), (5, 2000, {}), (5, 2000, {"GG_TR_L2MB": 0}), (1, 5000, {})]):
        timeit(M, p, env)
    if not old:
        names = ["X3", "householder+g", "strip product", "X1", "correction", "X2", "w+next row", "prologue", "epilogue"]
        out["phases"] = []
        for M, p, env in [(20, 1000, {"GG_TR_BLOCKED": 1})]:
            os.environ["GG_TR_TIMING"] = "1"
            for k, v in env.items():
                os.environ[k] = str(v)
            A = to_dev(sym(np.random.default_rng(1), M, p), dev)
            e = Eigh(M, p, dev)
            buf = (ctypes.c_ulonglong * 16)()
            for rep in range(2):
                W = A.clone()
                torch.cuda.synchronize()
                lib.gg_sytrd_phase_clock(buf)
                lib.gg_sytrd_profile(_p(W), _p(e.D), M, p, _p(e.ws), e.ws_bytes, 0, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
            lib.gg_sytrd_phase_clock(buf)
            rec = {"M": M, "p": p, "env": env, "ms": {nm: buf[i] / 1e6 for i, nm in enumerate(names)}}
            print(rec, flush=True)
            out["phases"].append(rec)
            for k in list(env) + ["GG_TR_TIMING"]:
                os.environ.pop(k, None)
    json.dump(out, open(path, "w"), indent=1)
