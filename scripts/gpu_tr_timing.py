"""Reads the %globaltimer stamps written by the -DTR_TIMING build (see gpu_tr_timing.sh) for one sytrd of a batch."""
import json
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gglasso_b200 import _lib
from gglasso_b200._engine import Eigh, _p

M, p = int(os.environ.get("PM", 20)), int(os.environ.get("PP", 1000))
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(7)
A = torch.randn(M, p, p, dtype=torch.float64, device=dev, generator=g)
A = ((A + A.transpose(1, 2)) / 2).contiguous()
e = Eigh(M, p, dev)
lib = _lib.load()
stream = torch.cuda.current_stream().cuda_stream
W = A.clone()
for rep in range(3):
    W.copy_(A)
    torch.cuda.synchronize()
    rc = lib.gg_sytrd_profile(_p(W), _p(e.D), M, p, _p(e.ws), e.ws_bytes, 0, stream)
    torch.cuda.synchronize()
    assert rc == 0


def al(x):
    return (x + 255) // 256 * 256


off = 2 * al(8 * M * p * p)
st = e.ws[off:off + 8 * p * 16].view(torch.int64).cpu().numpy().reshape(p, 2, 8)
js = p - 144
col, sv = st[:js, 0, :6].astype(np.float64), st[:js, 1, :7].astype(np.float64)
names_c = ["wait", "skip", "loads+red1", "red2", "stores"]
names_s = ["wait", "skip", "vec->smem", "tile landed", "sums+sync", "atomics"]
out = {}
for lo, hi in ((1, 100), (100, 400), (400, 600), (600, js - 1)):
    c, s = col[lo:hi], sv[lo:hi]
    d = {"col_total_us": float(np.mean(c[:, 5] - c[:, 0])) / 1e3,
         "col_stages_us": {n: float(np.mean(c[:, i + 1] - c[:, i])) / 1e3 for i, n in enumerate(names_c)},
         "symv_cta0_total_us": float(np.mean(s[:, 6] - s[:, 0])) / 1e3,
         "symv_stages_us": {n: float(np.mean(s[:, i + 1] - s[:, i])) / 1e3 for i, n in enumerate(names_s)},
         "col_end_to_symv_waitdone_us": float(np.mean(s[:, 1] - c[:, 5])) / 1e3,
         "symv_waitdone_to_next_col_waitdone_us": float(np.mean(col[lo + 1:hi + 1, 1] - s[:, 1])) / 1e3,
         "per_column_us": float(np.mean(col[lo + 1:hi + 1, 1] - c[:, 1])) / 1e3}
    out[f"j{lo}-{hi}"] = d
per = (col[1:js, 1] - col[:js - 1, 1]) / 1e3                     # column period vs j (us)
symv_span = (col[1:js, 1] - sv[:js - 1, 1]) / 1e3              # symv wait-done -> next column step released
out["per_column_us_every_16"] = [[int(j), int(p - j - 1), round(float(np.median(per[j:j + 16])), 2),
                                  round(float(np.median(symv_span[j:j + 16])), 2)] for j in range(0, js - 17, 16)]
print(json.dumps({k: v for k, v in out.items() if k != "per_column_us_every_16"}, indent=1))
for row in out["per_column_us_every_16"]:
    print("j=%4d t=%4d period %6.2f us  symv->col %6.2f us" % tuple(row))
json.dump(out, open("gpurun_out/tr_timing.json", "w"), indent=1)
