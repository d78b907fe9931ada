"""torchrun --nproc-per-node N scripts/dist_block_sgl.py out.json : cfg5 (block_SGL, p=5000) with the connected
components distributed over the ranks (parallel.block_SGL_dist) against the single-GPU block_SGL and the reference fixture."""
import contextlib, io, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gglasso_b200 import block_SGL
from gglasso_b200.parallel import block_SGL_dist
from oracle import ref_inputs


def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    S = ref_inputs.load("cfg5")
    p = S.shape[0]
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cfg5_block_sgl.npz"))
    out = {"world": world}
    for lam in (0.1, 0.05):
        tag = str(lam).replace(".", "")
        for rep in range(2):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sol = block_SGL_dist(S, lam, np.eye(p), tol=1e-7, rtol=1e-7)
            torch.cuda.synchronize()
            t_dist = time.perf_counter() - t0
        t = torch.tensor([t_dist], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rec = {"dist_seconds": float(t.item())}
        if rank == 0:
            for rep in range(2):
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    one = block_SGL(S, lam, np.eye(p), tol=1e-7, rtol=1e-7)
                rec["single_gpu_seconds"] = time.perf_counter() - t0
            ref = np.zeros(p * p); ref[g[f"theta_idx_{tag}"]] = g[f"theta_val_{tag}"]; ref = ref.reshape(p, p)
            rec["theta_vs_single_gpu"] = float(np.abs(sol["Theta"] - one["Theta"]).max())
            rec["theta_rel_err_vs_reference"] = float(np.linalg.norm(sol["Theta"] - ref) / np.linalg.norm(ref))
            rec["pattern_identical_to_reference"] = bool(np.array_equal(sol["Theta"] != 0, ref != 0))
            rec["reference_seconds"] = float(g[f"wall_{tag}"]); rec["reference_cores"] = int(g["ref_cores"])
            print(lam, rec, flush=True)
        out[f"lambda1_{lam}"] = rec
    if rank == 0:
        json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dist_block_sgl.json", "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
