import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import gglasso_b200._engine as eng
try:
    r = bench.grid_bench(1, 0, lambda: None)
    torch.cuda.synchronize()
    print("GRID_OK", r)
except Exception as ex:
    print("GRID_FAIL", type(ex).__name__, str(ex)[:300])
    for i, (e, buf) in enumerate(eng._DEBUG_BUFFERS[-12:]):
        np.save(f"gpurun_out/dbg_W_{i}.npy", buf.numpy())
        np.save(f"gpurun_out/dbg_args_{i}.npy", np.array([e._dbg_args[0], e._dbg_args[1], e.M, e.p]))
        if e._dbg_args[2] is not None:
            np.save(f"gpurun_out/dbg_ctrl_{i}.npy", e._dbg_args[2])
    print("saved", len(eng._DEBUG_BUFFERS[-12:]))
    os._exit(0)
