import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200.parallel import KShard
from gglasso_b200 import _lib
from gglasso_b200._engine import _p
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
lib = _lib.load()
K, p = 20, 1000
sh = KShard(K * world, p)
x = torch.randn(K, p, p, dtype=torch.float64, device="cuda")
def tm(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
out = {"p2p": torch.cuda.can_device_access_peer(0, 1) if world > 1 else None}
out["to_band_ms"] = tm(lambda: sh.to_band(x))
b = sh.to_band(x)
out["from_band_ms"] = tm(lambda: sh.from_band(b))
send = torch.randn(K * p * p, dtype=torch.float64, device="cuda"); recv = torch.empty_like(send)
out["a2a_equal_160MB_ms"] = tm(lambda: dist.all_to_all_single(recv, send))
out["allreduce_5_ms"] = tm(lambda: dist.all_reduce(torch.zeros(5, device="cuda", dtype=torch.float64)))
ctrl = torch.zeros(16, dtype=torch.float64, device="cuda"); ctrl[0] = 1; ctrl[1] = 1
T = torch.empty_like(b)
out["prox_band_ms"] = tm(lambda: lib.gg_prox_band(_p(b), _p(T), _p(ctrl), 0.05, 0.01, 1, K * world, sh.nb, p, sh.r_lo, torch.cuda.current_stream().cuda_stream))
h = torch.empty(K, p, p, dtype=torch.float64)
out["d2h_pageable_160MB_ms"] = tm(lambda: x.cpu(), n=3)
hp = torch.empty(K, p, p, dtype=torch.float64, pin_memory=True)
out["d2h_pinned_160MB_ms"] = tm(lambda: hp.copy_(x), n=3)
xn = x.cpu().numpy()
out["h2d_pageable_160MB_ms"] = tm(lambda: torch.from_numpy(xn).to("cuda"), n=3)
if rank == 0: print("DIAG", json.dumps(out))
dist.destroy_process_group()
