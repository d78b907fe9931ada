"""Concurrent stress of the batched eigensolver: T host threads, one CUDA stream and workspace each, repeat the
decomposition of the same finite input and report the first non-finite / inaccurate result."""
import os, sys, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200._engine import Eigh

M, p = int(os.environ.get("PM", 10)), int(os.environ.get("PP", 500))
T, R = int(os.environ.get("NT", 5)), int(os.environ.get("REPS", 150))
vectors = int(os.environ.get("VEC", 1))
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(3)
A = torch.randn(M, p, p, dtype=torch.float64, device=dev, generator=g)
A = ((A + A.transpose(1, 2)) / 2).contiguous()
ref = torch.linalg.eigvalsh(A)
torch.cuda.synchronize()
fails = []


def worker(t):
    torch.cuda.set_device(dev)
    with torch.cuda.stream(torch.cuda.Stream(device=dev)):
        e = Eigh(M, p, dev)
        W = torch.empty_like(A)
        s = torch.cuda.current_stream().cuda_stream
        for r in range(R):
            W.copy_(A)
            D = e.eigh(W, vectors=vectors, stream=s)
            Ds = torch.sort(D, dim=1).values
            err = (Ds - ref).abs().amax(dim=1)
            bad = ~(err < 1e-9)
            if bool(bad.any()):
                fails.append((t, r, torch.nonzero(bad).flatten().tolist(), err[bad].tolist()[:3]))
                return
        torch.cuda.current_stream().synchronize()


th = [threading.Thread(target=worker, args=(t,)) for t in range(T)]
[x.start() for x in th]
[x.join() for x in th]
print("STRESS", "T", T, "reps", R, "vectors", vectors, "FAILS" if fails else "OK", fails[:4], flush=True)
os._exit(0)
