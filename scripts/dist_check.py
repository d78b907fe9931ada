"""torchrun --nproc-per-node N scripts/dist_check.py : K-sharded ADMM_MGL_dist vs the single-GPU ADMM_MGL.
Every rank solves the full problem alone (reference) and its shard cooperatively; prints max deviations."""
import contextlib, io, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gglasso_b200 import ADMM_MGL
from gglasso_b200.parallel import ADMM_MGL_dist, partition
from gglasso_b200.datagen import synthetic_mgl

def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    out = {}
    for (reg, K, p, latent) in (("GGL", 5, 60, False), ("FGL", 7, 200, False), ("FGL", 4, 90, True)):
        S = synthetic_mgl(K, p, N=2 * p, seed=11, kind="fused" if reg == "FGL" else "group")
        Om0 = np.repeat(np.eye(p)[None], K, 0)
        kw = dict(tol=1e-7, rtol=1e-7)
        with contextlib.redirect_stdout(io.StringIO()):
            ref, rinfo = ADMM_MGL(S, 0.05, 0.02, reg, Om0, measure=True, latent=latent, mu1=0.1 if latent else None, **kw)
        lo, hi = partition(K, world)[rank]
        sol, info = ADMM_MGL_dist(S[lo:hi], 0.05, 0.02, reg, Om0[lo:hi], K_total=K, latent=latent,
                                  mu1_local=0.1 if latent else None, **kw)
        err = {k: float(np.abs(sol[k] - ref[k][lo:hi]).max()) for k in ("Omega", "Theta", "X", "L")}
        t = torch.tensor([max(err.values())], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"{reg}_K{K}_p{p}_lat{int(latent)}"] = dict(maxerr=float(t.item()), iters=info["iterations"], ref_iters=len(rinfo["residual"]),
                                                       status=info["status"], ref_status=rinfo["status"],
                                                       pattern_equal=bool(np.array_equal(sol["Theta"] != 0, ref["Theta"][lo:hi] != 0)))
    if rank == 0:
        from gglasso_b200 import parallel
        print("DIST_EXCHANGE", parallel.LAST_EXCHANGE)
        print("DIST_CHECK", json.dumps(out))
        ok = all(v["maxerr"] < 1e-9 and v["iters"] == v["ref_iters"] and v["status"] == v["ref_status"] and v["pattern_equal"] for v in out.values())
        print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAIL")
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
