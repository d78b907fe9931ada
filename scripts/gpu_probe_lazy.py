"""Probe for the lazy write-back depth of the tridiagonalisation (env GG_TR_LAZY, GG_SV_OCC): accuracy + stage times."""
import json
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200 import _lib
from gglasso_b200._engine import Eigh, _p

tag = sys.argv[1]
M, p = int(os.environ.get("PM", 20)), int(os.environ.get("PP", 1000))
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(7)
A = torch.randn(M, p, p, dtype=torch.float64, device=dev, generator=g)
A = ((A + A.transpose(1, 2)) / 2).contiguous()
e = Eigh(M, p, dev)
lib = _lib.load()
stream = torch.cuda.current_stream().cuda_stream
W = A.clone()
D = e.eigh(W, stream=stream).clone()
V = W
resid = (torch.bmm(V, A) - D[:, :, None] * V).abs().max().item()
orth = (torch.bmm(V, V.transpose(1, 2)) - torch.eye(p, dtype=torch.float64, device=dev)).abs().max().item()
out = {"tag": tag, "M": M, "p": p, "resid": resid, "orth": orth}
np.save(f"gpurun_out/lazy_D_{tag}.npy", D.cpu().numpy())


def timeit(fn, reps=6):
    best = 1e30
    for r in range(reps):
        W.copy_(A)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if r >= 1:
            best = min(best, a.elapsed_time(b))
    return best


out["eigh_ms"] = timeit(lambda: e.eigh(W, stream=stream))
for which, name in ((2, "symv_ms"), (1, "col_ms"), (3, "sytrd_ms")):
    out[name] = timeit(lambda: lib.gg_sytrd_profile(_p(W), _p(e.D), M, p, _p(e.ws), e.ws_bytes, which, stream))
print(json.dumps(out))
open("gpurun_out/lazy_probe.jsonl", "a").write(json.dumps(out) + "\n")
