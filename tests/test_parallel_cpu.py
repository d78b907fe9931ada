"""Host-side logic of the multi-GPU paths, exercised on CPU with world_size-2 gloo process groups:
partitioning, the K-layout <-> row-band re-tile (all_to_all_single), the sharded lambda grid and LPT block assignment."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gglasso_b200.parallel import (KShard, assign_blocks, band_layout_index, block_SGL_dist, ebic_mgl, grid_search_dist,
                                    partition)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def test_partition():
    for n in (0, 1, 7, 20, 1000):
        for parts in (1, 2, 3, 8):
            pr = partition(n, parts)
            assert len(pr) == parts and pr[0][0] == 0 and pr[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(pr, pr[1:]))
            sizes = [hi - lo for lo, hi in pr]
            assert max(sizes) - min(sizes) <= 1
    assert partition(20, 8) == [(0, 3), (3, 6), (6, 9), (9, 12), (12, 14), (14, 16), (16, 18), (18, 20)]


def _retile_worker(rank, world, port, K, p, q):
    _init(rank, world, port)
    try:
        full = torch.arange(K * p * p, dtype=torch.float64).view(K, p, p)
        sh = KShard(K, p)
        loc = full[sh.k_lo:sh.k_hi].clone()
        band = sh.to_band(loc)
        ok1 = torch.equal(band, full[:, sh.r_lo:sh.r_hi, :])
        back = sh.from_band(band * 2.0)
        ok2 = torch.equal(back, 2.0 * loc)
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        sh.allreduce_sum(t)
        q.put((rank, bool(ok1), bool(ok2), float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K,p", [(5, 7), (2, 3), (20, 16), (3, 1)])
def test_k_to_band_retile_gloo_ws2(K, p):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_retile_worker, args=(r, 2, port, K, p, q)) for r in range(2)]
    [pr.start() for pr in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [pr.join(timeout=60) for pr in procs]
    assert res == [(0, True, True, 3.0), (1, True, True, 3.0)]


def test_retile_single_process():
    full = torch.randn(4, 6, 6, dtype=torch.float64)
    sh = KShard(4, 6)
    assert torch.equal(sh.to_band(full), full)
    assert torch.equal(sh.from_band(sh.to_band(full)), full)


def _oracle_solver(S, l1, l2, reg, Omega_0, tol=1e-7, rtol=1e-7, **kw):
    from oracle import admm_oracle as orc
    return orc.admm_mgl(S, l1, l2, reg, Omega_0, tol=tol, rtol=rtol, **kw)


def _grid_inputs():
    rng = np.random.default_rng(3)
    K, p, N = 2, 12, 80
    S = np.stack([np.cov(rng.standard_normal((p, N)), bias=True) for _ in range(K)])
    return S, np.full(K, N), np.logspace(-0.5, -1.5, 3), np.logspace(-1, -2, 2)


def _grid_worker(rank, world, port, q):
    _init(rank, world, port)
    try:
        S, N, l1, l2 = _grid_inputs()
        scores, ix, best = grid_search_dist(_oracle_solver, S, N, "GGL", l1, l2, gamma=0.1, tol=1e-6, rtol=1e-6)
        q.put((rank, scores, ix, best["Theta"]))
    finally:
        dist.destroy_process_group()


def test_grid_search_sharded_gloo_ws2_matches_single_process():
    S, N, l1, l2 = _grid_inputs()
    scores1, ix1, best1 = grid_search_dist(_oracle_solver, S, N, "GGL", l1, l2, gamma=0.1, tol=1e-6, rtol=1e-6)
    assert not np.isnan(scores1).any()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grid_worker, args=(r, 2, port, q)) for r in range(2)]
    [pr.start() for pr in procs]
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    [pr.join(timeout=60) for pr in procs]
    for rank, scores, ix, theta in res:
        np.testing.assert_allclose(scores, scores1, rtol=1e-9)
        assert tuple(ix) == tuple(ix1)
        np.testing.assert_allclose(theta, best1["Theta"], atol=1e-10)


def test_ebic_matches_reference_formula():
    rng = np.random.default_rng(0)
    p, N = 6, 50
    S = np.cov(rng.standard_normal((p, N)), bias=True)[None]
    Theta = np.linalg.inv(S[0])
    Theta[np.abs(Theta) < 0.5] = 0
    Theta = (Theta + Theta.T) / 2 + 2 * np.eye(p)
    E = (np.count_nonzero(Theta) - p) / 2
    want = N * np.sum(S[0] * Theta) - N * np.linalg.slogdet(Theta)[1] + E * (np.log(N) + 4 * np.log(p) * 0.3)
    assert abs(ebic_mgl(S, Theta[None], np.array([N]), 0.3) - want) < 1e-9 * abs(want)
    assert ebic_mgl(S, -Theta[None], np.array([N]), 0.3) == np.inf


def test_assign_blocks_lpt():
    sizes = [1289, 19, 19, 18, 7, 5, 3, 2, 2]
    owner = assign_blocks(sizes, 4)
    assert owner[0] != owner[1]                      # the dominant block sits alone first
    load = np.zeros(4)
    for s, o in zip(sizes, owner):
        load[o] += s ** 3
    assert load.max() == 1289 ** 3                    # nothing else lands on the big block's rank
    assert set(assign_blocks([5, 5, 5, 5], 2)) == {0, 1}


def _oracle_sgl(S, lambda1, Omega_0, Theta_0=None, X_0=None, **kw):
    from oracle import admm_oracle as orc
    kw.pop("verbose", None)
    return orc.admm_sgl(S, lambda1, Omega_0, Theta_0, X_0, **kw)


def _block_inputs():
    rng = np.random.default_rng(4)
    sizes = [1, 2, 3, 6, 9, 1, 4]
    p = sum(sizes)
    S = np.zeros((p, p))
    o = 0
    for s in sizes:
        Z = rng.standard_normal((s, 3 * s + 4))
        B = Z @ Z.T / (3 * s + 4)
        S[o:o + s, o:o + s] = 0.5 * B / np.sqrt(np.outer(np.diag(B), np.diag(B))) + 0.5 * np.eye(s)
        o += s
    perm = rng.permutation(p)
    return S[np.ix_(perm, perm)], 0.02


def _block_worker(rank, world, port, q):
    _init(rank, world, port)
    try:
        S, lam = _block_inputs()
        sol = block_SGL_dist(S, lam, np.eye(S.shape[0]), tol=1e-8, rtol=1e-8, solver=_oracle_sgl)
        q.put((rank, sol["Theta"], sol["Omega"], sol["X"]))
    finally:
        dist.destroy_process_group()


def test_block_sgl_distributed_gloo_ws2_matches_oracle():
    from oracle import admm_oracle as orc
    S, lam = _block_inputs()
    ref = orc.block_sgl(S, lam, np.eye(S.shape[0]), tol=1e-8, rtol=1e-8)
    one = block_SGL_dist(S, lam, np.eye(S.shape[0]), tol=1e-8, rtol=1e-8, solver=_oracle_sgl)   # world = 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_block_worker, args=(r, 2, port, q)) for r in range(2)]
    [pr.start() for pr in procs]
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    [pr.join(timeout=60) for pr in procs]
    for got in [(0, one["Theta"], one["Omega"], one["X"])] + res:
        np.testing.assert_allclose(got[1], ref["Theta"], atol=1e-12)
        np.testing.assert_allclose(got[2], ref["Omega"], atol=1e-12)
        np.testing.assert_allclose(got[3], ref["X"], atol=1e-12)


def _packed_exchange_worker(rank, world, port, K, p, q):
    """the exchange of run_admm_mgl_dist with its preallocated buffers: the layout gg_pack_bands writes
    (band_layout_index) is exactly the split layout of the two all_to_all_single calls"""
    _init(rank, world, port)
    try:
        full = torch.arange(K * p * p, dtype=torch.float64).view(K, p, p)
        sh = KShard(K, p)
        loc = full[sh.k_lo:sh.k_hi].clone()
        idx = torch.from_numpy(band_layout_index(sh.K_loc, p, world)).view(-1)
        send = torch.empty(sh.K_loc * p * p, dtype=torch.float64)
        send[idx] = loc.view(-1)                                   # what gg_pack_bands does
        band = torch.empty(K * sh.nb * p, dtype=torch.float64)
        loc_split = [sh.K_loc * (hi - lo) * p for lo, hi in sh.rparts]
        band_split = [(khi - klo) * sh.nb * p for klo, khi in sh.kparts]
        dist.all_to_all_single(band, send, band_split, loc_split)
        ok1 = torch.equal(band.view(K, sh.nb, p), full[:, sh.r_lo:sh.r_hi, :])
        back = torch.empty(sh.K_loc * p * p, dtype=torch.float64)
        dist.all_to_all_single(back, band * 3.0, loc_split, band_split)
        ok2 = torch.equal(back[idx].view(sh.K_loc, p, p), 3.0 * loc)   # what gg_unpack_dual reads
        q.put((rank, bool(ok1), bool(ok2)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K,p", [(5, 7), (20, 16), (3, 2)])
def test_packed_band_exchange_gloo_ws2(K, p):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_packed_exchange_worker, args=(r, 2, port, K, p, q)) for r in range(2)]
    [pr.start() for pr in procs]
    res = [q.get(timeout=120) for _ in range(2)]
    [pr.join(60) for pr in procs]
    assert all(ok1 and ok2 for _, ok1, ok2 in res), res


def test_band_layout_index_is_a_permutation():
    for K_loc, p, world in [(1, 1, 1), (3, 10, 4), (2, 7, 7), (5, 9, 2), (10, 16, 8)]:
        idx = band_layout_index(K_loc, p, world)
        assert np.array_equal(np.sort(idx.reshape(-1)), np.arange(K_loc * p * p))
        if world == 1:
            assert np.array_equal(idx.reshape(-1), np.arange(K_loc * p * p))
