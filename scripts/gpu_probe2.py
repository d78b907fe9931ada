"""stage timings of the tridiagonal eigensolver (GG_TR_STOP=1: sytrd only, 2: +D&C, 0: full)."""
import json, os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from gglasso_b200._engine import Eigh, to_dev
    from gglasso_b200.datagen import synthetic_mgl
    dev = torch.device("cuda")
    out = {}
    for (M, p) in ((20, 1000), (10, 500), (1, 1289), (1, 2000), (4, 300)):
        Wd = to_dev(np.eye(p)[None] - synthetic_mgl(M, p, N=2 * p, seed=2), dev)
        e = Eigh(M, p, dev)
        ts = []
        for i in range(4):
            A = Wd.clone(); torch.cuda.synchronize()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); e.eigh(A); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        out[f"{M}x{p}"] = min(ts[1:])
    print(json.dumps(out))
else:
    res = {}
    for stop in ("1", "2", "0"):
        env = dict(os.environ, GG_TR_STOP=stop)
        o = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        try:
            res[{"1": "sytrd", "2": "sytrd+dc", "0": "full"}[stop]] = json.loads(o.stdout.strip().splitlines()[-1])
        except Exception:
            res[stop] = o.stdout[-500:] + o.stderr[-1500:]
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/probe2.json", "w"), indent=1)
    print(json.dumps(res, indent=1))
