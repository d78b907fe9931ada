"""per-column chain of the tridiagonalisation: two launches per column (GG_TR_MERGE=0) against the merged step kernel
(default): correctness against LAPACK and CUDA-event timing of gg_eigh and of the tridiagonalisation alone.
usage: python scripts/gpu_chain_merge.py [out.json]"""
import json, os, sys
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


for M, p in ((20, 1000), (3, 1000), (10, 500), (1, 1384), (2, 2047), (4, 300), (7, 145), (1, 2200)):
    rng = np.random.default_rng(p + M)
    A = sym(rng, M, p)
    Dref = np.linalg.eigvalsh(A)
    rec = {"M": M, "p": p}
    for merge in (0, 1):
        os.environ["GG_TR_MERGE"] = str(merge)
        e = Eigh(M, p, dev)
        st = torch.cuda.current_stream().cuda_stream
        At = to_dev(A, dev)
        D = e.eigh(At, stream=st)
        torch.cuda.synchronize()
        Dh, Vt = D.cpu().numpy(), At.cpu().numpy()
        rec[f"eig_err_{merge}"] = float(np.abs(np.sort(Dh, 1) - Dref).max())
        rec[f"resid_{merge}"] = max(float(np.abs(Vt[m] @ A[m] - Dh[m][:, None] * Vt[m]).max()) for m in range(M))
        rec[f"orth_{merge}"] = max(float(np.abs(Vt[m] @ Vt[m].T - np.eye(p)).max()) for m in range(M))
        A0 = to_dev(A, dev)
        for name, fn in (("eigh", lambda W: e.eigh(W, stream=st)),
                         ("sytrd", lambda W: lib.gg_sytrd_profile(_p(W), _p(e.D), M, p, _p(e.ws), e.ws_bytes, 0, st))):
            ts = []
            for r in range(6):
                W = A0.clone()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(W); b.record()
                torch.cuda.synchronize()
                if r:
                    ts.append(a.elapsed_time(b))
            rec[f"{name}_ms_{merge}"] = float(np.median(ts))
    print(rec, flush=True)
    out["checks"].append(rec)
os.environ.pop("GG_TR_MERGE", None)
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/chain_merge.json", "w"), indent=1)
