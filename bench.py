"""bench.py -- headline benchmark of the B200 ADMM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg3|cfg4col]

Metric (BASELINE.json): ADMM iterations/sec for the fused multiple graphical lasso K=20, p=1000
(cfg3: ADMM_MGL, reg='FGL', lambda1=0.05, lambda2=0.01, N=2000 samples per instance, rho=1, update_rho).
A "step" is one ADMM iteration over the whole (K,p,p) stack.

  value : iterations/sec with S resident in HBM, timed with CUDA events around exactly --steps
          iterations (each iteration reads/writes 160 MB arrays: working set >> 126 MB L2)
  e2e   : same metric through the public reference-signature call ADMM_MGL(S_host, ...) with
          max_iter = --steps: host->device copy of S/Omega_0, the iterations, post-loop checks and
          the device->host copy of sol are all inside the timed region
  roofline : dominant kernel of the step (tr_symv_kernel: trailing-matrix pass of the tridiagonalisation, HBM bound)
  cpu_baseline : the oracle port (numpy/LAPACK + C prox; same algorithm as the reference) timed on
          the host cores for a bounded number of iterations of the same workload

N > 1 (torchrun, one rank per GPU): ONE fused-MGL problem with K = 20*N instances, 20 per rank (weak
scaling).  Everything per instance stays local; the cross-instance TV prox needs all K values of an entry,
so V = Omega + X is re-tiled instance-layout -> row-band layout with an NCCL all-to-all, the prox runs on the
band and Theta travels back with a second all-to-all; 5 residual sums are all-reduced.  `value` counts units of
(K=20, p=1000) stacks processed per second over all ranks = N * iterations/sec.
--impl reference : rank 0 times the CPU oracle port on the same config.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(K=20, p=1000, N=2000, lambda1=0.05, lambda2=0.01, reg="FGL", seed=1234)


def make_input(cfg):
    from gglasso_b200.datagen import synthetic_mgl
    return synthetic_mgl(cfg["K"], cfg["p"], N=cfg["N"], seed=cfg["seed"], kind="fused")


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region: ONE long-running `nvidia-smi -lms 200`
    process (started before, killed after) -- no per-sample fork that could stall the launching thread."""

    def __init__(self, idx):
        self.idx, self.rows, self.proc = idx, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_reference_rate(S, cfg, iters):
    """oracle port (CPU): iterations/sec for `iters` iterations of the same workload, all host threads
    (torchrun exports OMP_NUM_THREADS=1; the BLAS pool is widened explicitly)."""
    from oracle import admm_oracle as orc
    orc.build_c()
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    K, p = cfg["K"], cfg["p"]
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    small = np.repeat(np.eye(8)[None], 2, 0)
    orc.admm_mgl(small, 0.1, 0.1, cfg["reg"], small, max_iter=2)           # warm-up
    t0 = time.perf_counter()
    _, info = orc.admm_mgl(S, cfg["lambda1"], cfg["lambda2"], cfg["reg"], Om0, tol=1e-7, rtol=1e-7, max_iter=iters)
    dt = time.perf_counter() - t0
    return info["iterations"] / dt, info["iterations"], dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S = make_input(CFG)
    steps = max(1, min(args.steps, 4))
    if args.warmup > 0:
        cpu_reference_rate(S, CFG, 1)
    rate, n, dt = cpu_reference_rate(S, CFG, steps)
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": "admm_iters_per_sec", "value": rate, "unit": "iter/s",
            "n_gpus": args.gpus, "steps": n, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 / rate,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg3: ADMM_MGL FGL K=20 p=1000 N=2000 lambda1=0.05 lambda2=0.01", **CFG},
            "cpu_baseline": {"value": rate, "unit": "iter/s", "cores": cores, "kind": "port",
                             "sample": f"{n} ADMM iterations of the full K=20 p=1000 workload ({dt:.1f} s)"},
            "e2e": {"value": rate, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-iters", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-grid", action="store_true", help="skip the secondary cfg4 lambda-grid measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gglasso_b200 import ADMM_MGL, _lib
    from gglasso_b200._engine import run_admm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.load()
    import gglasso_b200._engine as _eng
    _eng.warmup()                              # one-time process init (pinned staging buffers), outside timed regions
    cfg = dict(CFG)
    cfg["seed"] = CFG["seed"] + rank          # each rank: its own replica of the workload
    S = make_input(cfg)
    K, p = cfg["K"], cfg["p"]
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    steps, warm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing: exactly `steps` iterations -------------------------
    # run_admm is the engine behind ADMM_MGL; tol=0 disables the stopping test so that exactly
    # warm+steps iterations execute; inputs are uploaded before the timed region.
    import gglasso_b200._engine as eng
    marks = {}
    orig_step = eng.AdmmState.omega_step
    count = {"n": 0}

    def stepped(self):
        if count["n"] == warm:
            barrier()
            marks["e0"] = torch.cuda.Event(enable_timing=True)
            marks["e0"].record()
        count["n"] += 1
        return orig_step(self)

    eng.AdmmState.omega_step = stepped
    from gglasso_b200.parallel import ADMM_MGL_dist, run_admm_mgl_dist
    with ClockSampler(local) as clk:
        if world == 1:
            st, res = run_admm("mgl", S, Om0, None, None, lambda1=cfg["lambda1"], lambda2=cfg["lambda2"],
                               reg=cfg["reg"], tol=0.0, rtol=0.0, max_iter=warm + steps, check_every=10 ** 9)
        else:
            # one MGL problem with K = 20*N instances, 20 per rank: per-instance work stays local, the
            # cross-instance prox goes through two all-to-all re-tiles per iteration (weak scaling)
            st, _ = run_admm_mgl_dist(S, cfg["lambda1"], cfg["lambda2"], cfg["reg"], Om0, K_total=K * world,
                                      tol=0.0, rtol=0.0, max_iter=warm + steps, check_every=10 ** 9)
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        barrier()
    eng.AdmmState.omega_step = orig_step
    ms = marks["e0"].elapsed_time(e1)
    sweeps = [0] * steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * steps / (ms / 1e3)

    # ---------------- kernel-level roofline of the dominant kernel -------------------------------
    roof = kernel_roofline(st, sweeps, ms / steps) if rank == 0 else None

    # ---------------- end to end through the public API (host buffers) ---------------------------
    dt = None
    for rep in range(2):          # first call = warm-up (lazy kernel module loading, allocator); second is reported
        barrier()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            if world == 1:
                sol, info = ADMM_MGL(S, cfg["lambda1"], cfg["lambda2"], cfg["reg"], Om0, tol=0.0, rtol=0.0,
                                     max_iter=steps)
            else:
                sol, info = ADMM_MGL_dist(S, cfg["lambda1"], cfg["lambda2"], cfg["reg"], Om0, K_total=K * world,
                                          tol=0.0, rtol=0.0, max_iter=steps, check_every=10 ** 9)
                sol.pop("L")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * steps / float(t.item())
    h2d = world * (S.nbytes + Om0.nbytes) / steps
    d2h = world * sum(v.nbytes for k, v in sol.items() if not (k == "L")) / steps

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            rate, n, cdt = cpu_reference_rate(make_input(CFG), CFG, args.cpu_iters)
            cpu = {"value": rate, "unit": "iter/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": f"{n} ADMM iterations of the full K=20 p=1000 workload ({cdt:.1f} s)"}
        launches = launches_per_iter(p, K, sweeps)
        line = {"metric": "admm_iters_per_sec", "value": value, "unit": "iter/s", "n_gpus": world, "steps": steps,
                "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "cfg3: ADMM_MGL FGL K=20 p=1000 N=2000 lambda1=0.05 lambda2=0.01", **CFG,
                           "l2": "inputs larger than L2 (each (K,p,p) FP64 array is 160 MB; >10 arrays per step)",
                           "K_total": K * world, "partition": "K-sharded (20 instances per GPU); cross-instance prox via 2 all-to-all re-tiles per iteration" if world > 1 else "single GPU",
                           "eigh": "sytrd + divide&conquer + ormtr (hand-written)"},
                "e2e": {"value": e2e, "unit": "iter/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches, "clocks": clk.summary(), "roofline": roof, "cpu_baseline": cpu, "grid": None}

    # secondary measurement, last: the headline line above is complete before it starts and is printed even if the
    # grid run fails (single process; with several ranks a failure surfaces through torchrun)
    if not args.no_grid:
        if world == 1:
            try:
                line["grid"] = grid_bench(world, rank, barrier)
            except Exception as ex:                                   # noqa: BLE001
                line["grid"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
                print(json.dumps(line), flush=True)
                os._exit(0)
        else:
            g = grid_bench(world, rank, barrier)
            if rank == 0:
                line["grid"] = g
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def grid_bench(world, rank, barrier):
    """secondary measurement named by BASELINE.json's metric: the 10x10 lambda1 x lambda2 model-selection grid
    (cfg4: GGL, K=10, p=500, N=1000, eBIC gamma=0.1, tol=rtol=1e-7), device resident (scores and warm starts stay on
    the GPU), lambda1 columns sharded over the ranks and 5 columns concurrently per GPU; time = max over ranks."""
    import torch
    import torch.distributed as dist
    from gglasso_b200.datagen import synthetic_mgl
    from gglasso_b200.parallel import grid_search_device
    Sg = synthetic_mgl(10, 500, N=1000, seed=1234)
    Ng = np.full(10, 1000)
    l1, l2 = np.logspace(0, -3, 10), np.logspace(-1, -4, 10)
    grid_search_device(Sg, Ng, "GGL", l1[4:5], l2[:2], gamma=0.1, tol=1e-5, rtol=1e-5)        # warm-up
    barrier()
    t0 = time.perf_counter()
    scores, iters, ix, best = grid_search_device(Sg, Ng, "GGL", l1, l2, gamma=0.1, tol=1e-7, rtol=1e-7, n_streams=5)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"workload": "cfg4: 10x10 lambda grid, GGL K=10 p=500 N=1000, eBIC(0.1), tol=rtol=1e-7", "seconds": float(t.item()),
            "admm_iterations": int(iters.sum()), "grid_points": int(scores.size), "streams_per_gpu": 5,
            "best_lambda": [float(l1[ix[1]]), float(l2[ix[0]])], "scaling": "strong (columns sharded over ranks)"}


def launches_per_iter(p, K, sweeps):
    """kernel launches inside the timed region (all are kernels of libgglasso_b200.so)."""
    levels = 0
    while ((p + (1 << levels) - 1) >> levels) > 32:
        levels += 1
    sytrd = 2 * (p - 144) + 1 if p > 144 else 1               # column + trailing-matrix launch per column, smem tail
    ormtr = 3 + 2 * ((p - 1 + 127) // 128) if p >= 256 else 2   # blocked: gram, X, X*V, then 2 GEMMs per 128-block
    eigh = 4 + sytrd + 4 + 5 * levels + 1 + ormtr              # setup, sytrd, zero/scale/tear/leaves, merges, unscale
    per_iter = 1 + eigh + 1 + 1 + 1                            # build_w, eigh, recon, prox+dual, stop
    return int(per_iter * len(sweeps))


def kernel_roofline(st, sweeps, ms_per_step):
    """Dominant kernel of the step: tr_symv_kernel (trailing-matrix pass of the Householder
    tridiagonalisation: applies the pending rank-2 update to the upper triangle and accumulates the full
    symmetric A*v from that one half-matrix sweep; HBM/L2 bound).
    Timed live with CUDA events on the launch stream through gg_sytrd_profile(which=2), which issues
    exactly the (p-1) symv launches of one eigendecomposition of the batch."""
    import ctypes
    import torch
    from gglasso_b200 import _lib
    from gglasso_b200._engine import _p
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs")
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    if hbm is None:
        hbm, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    lib = _lib.load()
    M, p = st.M, st.p
    stream = torch.cuda.current_stream().cuda_stream
    times = {}
    for which in (2, 1, 0):
        best = 1e30
        for rep in range(3):
            W = (st.Theta - st.X - st.S).contiguous()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = lib.gg_sytrd_profile(_p(W), _p(st.eig.D), M, p, _p(st.eig.ws), st.eig.ws_bytes, which, stream)
            b.record()
            torch.cuda.synchronize()
            assert rc == 0, rc
            if rep >= 1:
                best = min(best, a.elapsed_time(b))
        times[which] = best
    n_launch = p - 1
    # algorithmic bytes of launch j: one read of the UPPER TRIANGLE of the t x t trailing block (t = p-j-1) of each
    # of the M matrices, 8 * t(t+1)/2 bytes, plus the same again on the passes that store it (every q-th pass,
    # q = gg_sytrd_write_depth(); same schedule as the host loop in gg_tridiag.cu)
    q = int(lib.gg_sytrd_write_depth())
    total_bytes, kb, n_write = 0.0, 0, 0
    for j in range(p - 1):
        write = (j - kb) >= q
        total_bytes += (16.0 if write else 8.0) * M * (p - j - 1) * (p - j) / 2
        if write:
            kb, n_write = j, n_write + 1
    achieved = total_bytes / (times[2] * 1e-3) / 1e9            # GB/s over all symv launches of one eigh
    return {"bound": "hbm", "kernel": "tr_symv_kernel", "achieved": achieved, "peak": hbm, "unit": "GB/s",
            "frac": achieved / hbm, "traffic": None, "peak_source": peak_src,
            "launches_per_step": n_launch, "write_depth": q, "write_passes": n_write,
            "avg_ms_per_launch": times[2] / n_launch,
            "algorithmic_bytes_per_launch_avg": total_bytes / n_launch,
            "symv_ms_per_step": times[2], "col_kernels_ms_per_step": times[1], "sytrd_ms_per_step": times[0],
            "share_of_step": times[2] / ms_per_step,
            "note": "trailing blocks shrink from 160 MB to 0 over the launches; blocks below ~126 MB total are L2 resident, "
                    "so late launches can exceed the HBM figure"}


if __name__ == "__main__":
    main()
