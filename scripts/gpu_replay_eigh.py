"""Replays eigendecomposition inputs captured with GG_DEBUG_KEEP_INPUT=1 (see gglasso_b200/_engine.py): checks the
input itself and repeats the decomposition to see whether a failure is data dependent."""
import glob, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200._engine import Eigh

for f in sorted(glob.glob("gpurun_out/dbg_W_*.npy")):
    i = f.split("_")[-1].split(".")[0]
    W = np.load(f)
    mpp, vectors, M, p = [int(x) for x in np.load(f"gpurun_out/dbg_args_{i}.npy")]
    fin = np.isfinite(W).all()
    asym = np.abs(W - W.transpose(0, 2, 1)).max() if fin else float("nan")
    print(f"[{i}] M={M} p={p} vectors={vectors} finite={fin} max|W|={np.nanmax(np.abs(W)):.3e} asym={asym:.2e}", flush=True)
    if not fin:
        bad = ~np.isfinite(W)
        print("    non-finite entries:", bad.sum(), "matrices:", np.unique(np.nonzero(bad)[0])[:10])
        continue
    dev = torch.device("cuda")
    Wd = torch.as_tensor(W, device=dev)
    e = Eigh(M, p, dev)
    ref = np.linalg.eigvalsh(W)
    for rep in range(int(os.environ.get("REPS", 6))):
        A = Wd.clone()
        D = e.eigh(A, vectors=vectors, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        Dh = np.sort(D.cpu().numpy(), axis=1)
        ok = np.isfinite(Dh).all()
        err = np.abs(Dh - ref).max() if ok else float("nan")
        print(f"    rep {rep}: finite={ok} max eigenvalue error={err:.2e}", flush=True)
