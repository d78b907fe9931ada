"""bench.py -- headline benchmark of the B200 ADMM hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-cpu] [--no-grid] [--no-weak]

Metric (BASELINE.json): ADMM iterations/sec AND time-to-tolerance for the fused multiple graphical lasso K=20, p=1000
(cfg3: ADMM_MGL, reg='FGL', lambda1=0.05, lambda2=0.01, N=2000 samples per instance, rho=1, update_rho=True), plus the
10x10 lambda grid (cfg4) as a secondary block.  Inputs come from the REFERENCE's own seeded generators
(time_varying_power_network(1000, 20, 10, seed=1234) + the sample_covariance_matrix recipe with N=2000, seed=1234 and a
host-independent Cholesky sampler, run from oracle/_ref through oracle/ref_inputs.py, cached under /tmp) -- input
preparation only, outside every timed region.
A "step" is one ADMM iteration over the whole (K,p,p) stack.

  value        iterations/sec with S resident in HBM, CUDA events around exactly --steps iterations (stopping test
               disabled by tol=rtol=0; every iteration streams >10 arrays of 160 MB: working set >> 126 MB L2)
  e2e          the same through the public reference-signature call ADMM_MGL(S_host, ...) with max_iter=--steps:
               H2D of S / Omega_0, the iterations, post-loop checks and the D2H of sol are inside the timed region
  time_to_tol  wall seconds of the public call to tol=rtol=1e-7 (host buffers in, host buffers out), its iteration
               count, and the final objective / sparsity pattern checked against the real reference's fixture
  roofline     dominant kernel of the step, timed live with CUDA events on the launch stream
  cpu_baseline the REAL reference (oracle/_ref: numpy/LAPACK + numba) on this host's cores: one full solve to
               tol=rtol=1e-7 of the same input (11 iterations); value = its in-loop iterations/sec, plus its wall time

N > 1 (torchrun, one rank per GPU) -- STRONG scaling of the same workload: ONE K=20, p=1000 problem, the instances
sharded 10/10, 5x4 or 3,3,3,3,2,2,2,2 over the ranks.  Everything per instance stays local; the cross-instance TV prox
needs all K values of an entry, so V = Omega + X is re-tiled instance layout -> row-band layout with an NCCL all-to-all,
the prox runs on the band and Theta travels back with a second all-to-all; 5 residual sums are all-reduced.  `value` =
iterations/sec of that one solve.  The sharded solve is checked inside the bench against the reference fixture
(`dist_check`).  The previous weak-scaling run (K = 20*N) is kept as the extra key `weak`.
--impl reference : rank 0 times the real reference (CPU) on the same config; other ranks exit.
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
WORKLOAD = ("cfg3: ADMM_MGL FGL K=20 p=1000 N=2000 lambda1=0.05 lambda2=0.01, inputs from the reference generators "
            "time_varying_power_network(1000,20,10,seed=1234) + sample_covariance_matrix(N=2000,seed=1234)")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_input(name="cfg3"):
    """reference-generated input (see module docstring); local rank 0 fills the cache, the others wait for the file"""
    from oracle import ref_inputs
    local = int(os.environ.get("LOCAL_RANK", "0"))
    path = os.path.join(ref_inputs.CACHE, name + "_S.npy")
    if local != 0:
        t0 = time.time()
        while not os.path.isfile(path) and time.time() - t0 < 600:
            time.sleep(0.5)
    return ref_inputs.load(name)


def input_check(name, S):
    from oracle import ref_inputs
    try:
        return ref_inputs.check_fingerprint(name, S, os.path.join(GOLDEN, "large_inputs.json"))
    except Exception as ex:                                   # noqa: BLE001
        return f"unavailable: {type(ex).__name__}"


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


# ------------------------------------------------------------------------------------------------------------------
# CPU: the real reference (oracle/_ref), timed on the host cores
# ------------------------------------------------------------------------------------------------------------------
def _widen_threads():
    try:                                   # torchrun exports OMP_NUM_THREADS=1; the BLAS pool is widened explicitly
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass


def reference_solver():
    """(callable, kind): the real reference's ADMM_MGL from oracle/_ref, else the oracle port"""
    from oracle import ref
    if ref.available():
        ADMM_MGL_ref, _, _ = ref.fresh_solvers()
        return ADMM_MGL_ref, "reference"
    from oracle import admm_oracle as orc
    orc.build_c()

    def port(S, l1, l2, reg, Om0, tol=1e-7, rtol=1e-7, max_iter=1000, measure=False, **kw):
        t0 = time.perf_counter()
        sol, info = orc.admm_mgl(S, l1, l2, reg, Om0, tol=tol, rtol=rtol, max_iter=max_iter)
        n = info["iterations"]
        return sol, {"status": info["status"], "runtime": np.full(n, (time.perf_counter() - t0) / n)}
    return port, "port"


def cpu_reference(S, cfg, max_iter, tol):
    """one call of the reference solver: (in-loop iterations/sec, iterations, wall seconds, kind)"""
    _widen_threads()
    solver, kind = reference_solver()
    K, p = cfg["K"], cfg["p"]
    Om0 = np.repeat(np.eye(p)[None], K, 0)
    small = np.repeat(np.eye(8)[None], 2, 0)
    with contextlib.redirect_stdout(io.StringIO()):
        solver(small + 0.1, 0.1, 0.1, cfg["reg"], small, max_iter=2, measure=True)          # numba compilation
        t0 = time.perf_counter()
        _, info = solver(S, cfg["lambda1"], cfg["lambda2"], cfg["reg"], Om0, tol=tol, rtol=tol, max_iter=max_iter,
                         measure=True)
        wall = time.perf_counter() - t0
    rt = np.asarray(info["runtime"], dtype=float)
    return rt, wall, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S = load_input("cfg3")
    # the reference needs ~2-4 s per iteration here: bound the sample so that the arm ends within a few minutes
    warm = min(args.warmup, 3)
    steps = max(1, min(args.steps, 12))
    rt, wall, kind = cpu_reference(S, CFG, warm + steps, 0.0)
    used = rt[warm:warm + steps]
    rate = len(used) / float(used.sum())
    cores = os.cpu_count()
    line = {"impl": "reference", "metric": "admm_iters_per_sec", "value": rate, "unit": "iter/s",
            "n_gpus": args.gpus, "steps": int(len(used)), "warmup": warm, "ms_per_step": 1e3 / rate,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, **CFG},
            "cpu_baseline": {"value": rate, "unit": "iter/s", "cores": cores, "kind": kind,
                             "sample": f"{len(used)} ADMM iterations (after {warm} warm-up iterations) of the full K=20 "
                                       f"p=1000 workload, in-loop time of the reference's own measure=True clock; "
                                       f"requested steps/warmup {args.steps}/{args.warmup} bounded to keep the arm "
                                       f"within minutes ({wall:.1f} s wall)"},
            "e2e": {"value": rate, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-grid", action="store_true", help="skip the secondary cfg4 lambda-grid measurement")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the extra weak-scaling run (K = 20*N)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gglasso_b200 import ADMM_MGL, _lib
    import gglasso_b200._engine as eng
    from gglasso_b200.parallel import ADMM_MGL_dist, run_admm_mgl_dist, partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    eng.warmup()                              # one-time process init (pinned staging buffers), outside timed regions
    cfg = dict(CFG)
    S_full = load_input("cfg3")
    in_dev = input_check("cfg3", S_full) if rank == 0 else None
    K, p = cfg["K"], cfg["p"]
    k_lo, k_hi = partition(K, world)[rank]
    def pinned(a):
        """host copy in page-locked memory (the e2e contract: inputs start in pinned host memory; the solver's upload
        then is one DMA per array instead of a staged copy)"""
        h = torch.empty(a.shape, dtype=torch.float64, pin_memory=True).numpy()
        h[...] = a
        return h

    S = pinned(S_full[k_lo:k_hi])
    Om0 = pinned(np.repeat(np.eye(p)[None], k_hi - k_lo, 0))
    steps, warm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_loop(S_loc, Om_loc, K_total):
        """exactly `steps` device-resident iterations (after `warm`), CUDA events on the launch stream, max over ranks"""
        marks, count = {}, {"n": 0}
        orig_step = eng.AdmmState.omega_step

        def stepped(self, *a):
            if count["n"] == warm:
                barrier()
                marks["e0"] = torch.cuda.Event(enable_timing=True)
                marks["e0"].record()
            count["n"] += 1
            return orig_step(self, *a)

        eng.AdmmState.omega_step = stepped
        l0 = lib.gg_launch_count()
        try:
            if world == 1:
                st, _ = eng.run_admm("mgl", S_loc, Om_loc, None, None, lambda1=cfg["lambda1"], lambda2=cfg["lambda2"],
                                     reg=cfg["reg"], tol=0.0, rtol=0.0, max_iter=warm + steps, check_every=10 ** 9)
            else:
                st, _ = run_admm_mgl_dist(S_loc, cfg["lambda1"], cfg["lambda2"], cfg["reg"], Om_loc, K_total=K_total,
                                          tol=0.0, rtol=0.0, max_iter=warm + steps, check_every=10 ** 9)
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            barrier()
        finally:
            eng.AdmmState.omega_step = orig_step
        launches = (lib.gg_launch_count() - l0) * steps // (warm + steps)
        return st, max_over_ranks(marks["e0"].elapsed_time(e1)), int(launches)

    # ---------------- device-resident timing -----------------------------------------------------------------------
    with ClockSampler(local) as clk:
        st, ms, launches = timed_loop(S, Om0, K)
    value = steps / (ms / 1e3)

    # ---------------- kernel-level roofline of the dominant kernel (rank 0) ----------------------------------------
    roof = kernel_roofline(st, ms / steps) if rank == 0 else None
    blocked = blocked_path_numbers(st, ms / steps) if (rank == 0 and world == 1) else None
    prox = prox_roofline(st, cfg["reg"]) if (rank == 0 and world == 1) else None
    del st

    # ---------------- end to end through the public API (host buffers), `steps` iterations -------------------------
    def public_call(max_iter, tol):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            if world == 1:
                sol, info = ADMM_MGL(S, cfg["lambda1"], cfg["lambda2"], cfg["reg"], Om0, tol=tol, rtol=tol,
                                     max_iter=max_iter)
            else:
                sol, info = ADMM_MGL_dist(S, cfg["lambda1"], cfg["lambda2"], cfg["reg"], Om0, K_total=K, tol=tol,
                                          rtol=tol, max_iter=max_iter, check_every=10 ** 9 if tol == 0.0 else 1)
        if "iterations" not in info:                   # the reference-signature call reports them in its printed line
            words = buf.getvalue().split("ADMM terminated after ")[-1].split()
            info = dict(info, iterations=int(words[0]) if words and words[0].isdigit() else None)
        return sol, info

    dt = None
    sol = None
    for rep in range(3):          # two warm-up calls (lazy module loading; the results live in page-locked memory from
        del sol                   # torch's caching host allocator, which is filled by the first calls); third reported
        barrier()
        t0 = time.perf_counter()
        sol, info = public_call(steps, 0.0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    e2e = steps / max_over_ranks(dt)
    h2d = (S_full.nbytes + S_full.nbytes) / steps                       # S and Omega_0 of all ranks
    d2h = 3 * S_full.nbytes / steps                                     # Omega, Theta, X of all ranks

    # ---------------- time to tolerance through the public API + parity against the reference fixture --------------
    del sol
    barrier()
    t0 = time.perf_counter()
    sol, info = public_call(1000, 1e-7)
    torch.cuda.synchronize()
    ttt = max_over_ranks(time.perf_counter() - t0)
    ttt_info = {"seconds": ttt, "tol": 1e-7, "status": info["status"],
                "iterations": int(info["iterations"]) if "iterations" in info else None,
                "what": "public call, host buffers in and out, H2D + loop + post-checks + D2H"}
    check = fixture_check(sol["Theta"], k_lo, k_hi, p, world)
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, check)
        check = {"pattern_identical": all(g["pattern_identical"] for g in gathered),
                 "theta_rel_err": float(np.sqrt(sum(g["err2"] for g in gathered) / sum(g["ref2"] for g in gathered))),
                 "fixture": gathered[0]["fixture"]}
    else:
        check = {"pattern_identical": check["pattern_identical"],
                 "theta_rel_err": float(np.sqrt(check["err2"] / check["ref2"])), "fixture": check["fixture"]}
    assert check["pattern_identical"] and check["theta_rel_err"] < 1e-8, check
    del sol

    # ---------------- extra: weak scaling (K = 20 per rank), as measured in round 1 --------------------------------
    weak = None
    if world > 1 and not args.no_weak:
        Omf = np.repeat(np.eye(p)[None], K, 0)
        _, wms, _ = timed_loop(S_full, Omf, K * world)
        weak = {"K_total": K * world, "ms_per_step": wms / steps, "units_per_s": world * steps / (wms / 1e3),
                "what": "ONE fused-MGL problem with K = 20 per rank (20*N instances), device resident"}

    line = None
    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            rt, wall, kind = cpu_reference(S_full, CFG, 1000, 1e-7)
            cpu = {"value": len(rt) / float(rt.sum()), "unit": "iter/s", "cores": os.cpu_count(), "kind": kind,
                   "time_to_tol_s": wall, "iterations_to_tol": int(len(rt)),
                   "sample": f"one full solve to tol=rtol=1e-7 of the same input: {len(rt)} iterations, in-loop "
                             f"{rt.sum():.1f} s (reference's measure=True clock), {wall:.1f} s wall incl. objective and checks"}
            ttt_info["cpu_seconds"] = wall
        shard = "single GPU" if world == 1 else (
            f"strong scaling: the K=20 instances sharded {[hi - lo for lo, hi in partition(K, world)]} over {world} "
            "GPUs; cross-instance prox via 2 NCCL all-to-all re-tiles per iteration, 5-double all-reduce")
        line = {"metric": "admm_iters_per_sec", "value": value, "unit": "iter/s", "n_gpus": world, "steps": steps,
                "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, **CFG,
                           "l2": "inputs larger than L2 (each (K,p,p) FP64 array is 160 MB; >10 arrays per step)",
                           "host_polling": "value: the timed region enqueues all steps without reading the device "
                                           "(tol = 0); e2e and time_to_tol: public call, control block read every "
                                           "iteration as in normal use",
                           "K_total": K, "partition": shard, "input_fingerprint_dev": in_dev,
                           "eigh": "sytrd (per-column chain, lazy write-back) + divide&conquer + blocked ormtr, hand-written"},
                "e2e": {"value": e2e, "unit": "iter/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "time_to_tol": ttt_info, "parity": check,
                "gpu_launches": launches * world, "clocks": clk.summary(), "roofline": roof, "cpu_baseline": cpu,
                "weak": weak, "blocked_sytrd": blocked, "prox": prox, "grid": None}

    # secondary measurement, last: the headline line above is complete before it starts and is printed even if the
    # grid run fails (single process; with several ranks a failure surfaces through torchrun)
    if not args.no_grid:
        if world == 1:
            try:
                line["grid"] = grid_bench(world, rank, barrier, max_over_ranks)
            except Exception as ex:                                   # noqa: BLE001
                line["grid"] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
                print(json.dumps(line), flush=True)
                os._exit(0)
        else:
            g = grid_bench(world, rank, barrier, max_over_ranks)
            if rank == 0:
                line["grid"] = g
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def fixture_check(Theta, k_lo, k_hi, p, world):
    """final Theta of the solve to tol=1e-7 against the real reference's result (tests/golden/cfg3_fgl_full.npz)"""
    g = np.load(os.path.join(GOLDEN, "cfg3_fgl_full.npz"))
    idx, val = g["theta_idx"], g["theta_val"]
    lo, hi = k_lo * p * p, k_hi * p * p
    m = (idx >= lo) & (idx < hi)
    ref = np.zeros((k_hi - k_lo) * p * p)
    ref[idx[m] - lo] = val[m]
    mine = Theta.reshape(-1)
    return {"pattern_identical": bool(np.array_equal(mine != 0, ref != 0)),
            "err2": float(np.sum((mine - ref) ** 2)), "ref2": float(np.sum(ref ** 2)),
            "fixture": f"tests/golden/cfg3_fgl_full.npz (real reference, {len(g['objective'])} iterations, "
                       f"objective {float(g['objective'][-1])!r})"}


def grid_bench(world, rank, barrier, max_over_ranks):
    """secondary measurement named by BASELINE.json's metric: the 10x10 lambda1 x lambda2 model-selection grid
    (cfg4: GGL, K=10, p=500, N=1000, eBIC gamma=0.1, tol=rtol=1e-7; input from group_power_network(500,10,10,seed=1234)),
    device resident (scores and warm starts stay on the GPU), lambda1 columns sharded over the ranks and 5 columns
    concurrently per GPU; time = max over ranks.  The eBIC table is compared with the real reference's."""
    from gglasso_b200.parallel import grid_search_device
    Sg = load_input("cfg4")
    Ng = np.full(10, 1000)
    l1, l2 = np.logspace(0, -3, 10), np.logspace(-1, -4, 10)
    grid_search_device(Sg, Ng, "GGL", l1[4:5], l2[:2], gamma=0.1, tol=1e-5, rtol=1e-5)        # warm-up
    barrier()
    t0 = time.perf_counter()
    scores, iters, ix, best = grid_search_device(Sg, Ng, "GGL", l1, l2, gamma=0.1, tol=1e-7, rtol=1e-7, n_streams=5)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    out = {"workload": "cfg4: 10x10 lambda grid, GGL K=10 p=500 N=1000, eBIC(0.1), tol=rtol=1e-7, reference-generated input",
           "seconds": dt, "admm_iterations": int(iters.sum()), "grid_points": int(scores.size), "streams_per_gpu": 5,
           "best_lambda": [float(l1[ix[1]]), float(l2[ix[0]])], "scaling": "strong (columns sharded over ranks)",
           "start_points": "every lambda1 column starts from the identity (the reference chains columns too)"}
    gpath = os.path.join(GOLDEN, "cfg4_grid.npz")
    if os.path.isfile(gpath):
        g = np.load(gpath)
        if "bic_10x10" in g.files:
            out["vs_reference_grid"] = {
                "same_best_index": bool(tuple(int(i) for i in ix) == tuple(int(i) for i in g["ix_10x10"])),
                "max_rel_dev_ebic": float(np.nanmax(np.abs(scores - g["bic_10x10"]) / np.abs(g["bic_10x10"]))),
                "reference_wall_s": float(g["wall_10x10"]), "reference_cores": int(g["ref_cores"])}
    return out


def prox_roofline(st, reg):
    """The HBM-bound kernel class of the step (fused prox + dual update + residual sums, 5 A bytes):
    prox_mgl_upper_kernel, timed live with CUDA events on the buffers of the finished run.  `achieved` is the
    algorithmic 5 A over the kernel time (the kernel itself moves 2.5 A: it works on upper triangles);
    `traffic` is the DRAM traffic of the committed ncu --set full capture of the same kernel at this size."""
    import torch
    from gglasso_b200 import _lib
    from gglasso_b200._engine import _p
    from gglasso_b200._lib import NPART
    lib = _lib.load()
    M, p = st.M, st.p
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        hbm, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    parts = torch.zeros((lib.gg_mgl_upper_nparts(p), NPART), dtype=torch.float64, device=st.dev)
    from gglasso_b200._lib import C_DONE
    ctrl = st.ctrl.clone()
    ctrl[:, C_DONE] = 0.0                              # (the kernel is a no-op for a finished problem)
    stream = torch.cuda.current_stream().cuda_stream
    ts = []
    for r in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = lib.gg_prox_mgl_upper(_p(st.Omega_new), _p(st.Omega), _p(st.X), _p(st.Theta), _p(ctrl), CFG["lambda1"],
                                   CFG["lambda2"], 0 if reg == "GGL" else 1, M, p, _p(parts), stream)
        b.record()
        torch.cuda.synchronize()
        assert rc == 0, rc
        if r >= 2:
            ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    alg = 5.0 * 8.0 * M * p * p
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))["prox_mgl_upper_kernel"]
        if tr["K"] == M and tr["p"] == p:
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        pass
    return {"bound": "hbm", "kernel": f"prox_mgl_upper_kernel<{reg}>", "achieved": alg / ms / 1e6, "peak": hbm,
            "unit": "GB/s", "frac": alg / ms / 1e6 / hbm, "traffic": traffic, "ms": ms, "algorithmic_bytes": alg,
            "moved_bytes": alg / 2, "frac_of_moved_bytes": alg / 2 / ms / 1e6 / hbm, "peak_source": peak_src,
            "note": "works on upper triangles only: 2.5 A moved for the 5 A step it replaces"}


def kernel_roofline(st, ms_per_step):
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
    # DRAM traffic per launch from the committed ncu --set full capture (one read/read/write cycle at t = 640), scaled
    # to the average launch by the ratio traffic / algorithmic bytes of the captured launches
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))["tr_symv_kernel"]
        ratio = ((2 * tr["read_pass_dram_bytes"] + tr["write_pass_dram_bytes"]) /
                 (2 * tr["algorithmic_read_pass_bytes"] + tr["algorithmic_write_pass_bytes"]))
        traffic = ratio * total_bytes / n_launch
        traffic_src = f"{ratio:.3f} x algorithmic, " + tr["source"]
    except Exception:
        pass
    return {"bound": "hbm", "kernel": "tr_symv_kernel", "achieved": achieved, "peak": hbm, "unit": "GB/s",
            "frac": achieved / hbm, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "launches_per_step": n_launch, "write_depth": q, "write_passes": n_write,
            "avg_ms_per_launch": times[2] / n_launch,
            "algorithmic_bytes_per_launch_avg": total_bytes / n_launch,
            "symv_ms_per_step": times[2], "col_kernels_ms_per_step": times[1], "sytrd_ms_per_step": times[0],
            "share_of_step": times[2] / ms_per_step,
            "note": "trailing blocks shrink from 160 MB to 0 over the launches; blocks below ~126 MB total are L2 resident, "
                    "so late launches can exceed the HBM figure"}


def blocked_path_numbers(st, ms_per_step):
    """Dominant kernel of the step: sytrd_panel_kernel, the panel kernel of the blocked tridiagonalisation (one launch
    per 16 columns, one thread-block cluster per matrix; per column every cluster streams the upper triangle of its
    trailing matrix once -- from shared memory, L2 or HBM -- and exchanges partial results over distributed shared
    memory).  Timed live with CUDA events on the launch stream through gg_sytrd_profile(which=1), which issues exactly
    the panel launches of one tridiagonalisation of the batch; which=2 issues the DMMA rank-2k updates, which=0 both
    plus the shared-memory tail."""
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
    os.environ["GG_TR_BLOCKED"] = "1"        # (read per call by the library) every column below the tail on the blocked path
    for which in (1, 2, 0):
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
    os.environ.pop("GG_TR_BLOCKED", None)
    tail, nb = 144, 16
    ncols = p - tail
    n_launch = (ncols + nb - 1) // nb
    # algorithmic bytes: column j reads the UPPER TRIANGLE of the t x t trailing block (t = p-j-1) of each of the M
    # matrices once, 8 * t(t+1)/2 bytes (a streaming FP64 pass moves 8 B per element read)
    total_bytes = sum(8.0 * M * (p - j - 1) * (p - j) / 2 for j in range(ncols))
    achieved = total_bytes / (times[1] * 1e-3) / 1e9
    flops_syr2k = sum(2.0 * M * (2 * nb) * (p - j0 - min(nb, ncols - j0)) ** 2 / 2 for j0 in range(0, ncols, nb))
    return {"what": "alternative blocked tridiagonalisation (not the default: the per-column chain is faster at this shape), timed the same way", "bound": "hbm", "kernel": "sytrd_panel_kernel", "achieved": achieved, "peak": hbm, "unit": "GB/s",
            "frac": achieved / hbm, "traffic": None, "peak_source": peak_src,
            "launches_per_step": n_launch, "columns_per_launch": nb, "avg_ms_per_launch": times[1] / n_launch,
            "algorithmic_bytes_per_launch_avg": total_bytes / n_launch,
            "panel_ms_per_step": times[1], "syr2k_ms_per_step": times[2], "sytrd_ms_per_step": times[0],
            "syr2k_tflops": flops_syr2k / (times[2] * 1e-3) / 1e12,
            "share_of_step": times[1] / ms_per_step,
            "note": "latency bound, not bandwidth bound: two cluster barriers and ~10 dependent shared-memory phases per "
                    "column; part of every strip is resident in shared memory or pinned in L2, so the HBM figure is an "
                    "upper bound on the traffic, not the traffic"}


if __name__ == "__main__":
    main()
