"""Multi-rank host logic on CPU: world_size-2 gloo processes exercise chain sharding and the ensemble gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_partitions_exactly():
    from tonga_b200.dist import shard
    for n in (1, 7, 8, 1024, 8192, 1000):
        for w in (1, 2, 3, 4, 8):
            blocks = [shard(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_history(global_ids, H=3, KC=8, R=5):
    """Deterministic per-chain 'history' that depends only on the GLOBAL chain id (as the Philox streams do)."""
    n = len(global_ids)
    g = np.asarray(global_ids, dtype=np.float64)
    return {"n_hist": torch.full((n,), H, dtype=torch.int32),
            "K": torch.from_numpy((5 + (np.asarray(global_ids)[:, None] + np.arange(H)[None]) % 4).astype(np.int32)),
            "cells": torch.from_numpy(g[:, None, None, None] + np.arange(H)[None, :, None, None] * 0.5 + np.zeros((n, H, 4, KC))),
            "phi": torch.from_numpy(100.0 + g[:, None] + np.arange(H)[None] * 0.25),
            "ptS": torch.from_numpy(g[:, None, None] * 0.1 + np.zeros((n, H, R)))}


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tonga_b200.dist import gather_ensembles, shard
        start, count = shard(n_total, world, rank)
        local = _fake_history(list(range(start, start + count)))
        counts = [shard(n_total, world, r)[1] for r in range(world)]
        ens = gather_ensembles(local, counts)
        from tonga_b200.dist import allreduce_sums
        s1, s2, cnt = allreduce_sums(np.arange(5.0) + rank, np.ones(5) * (rank + 1), 10 + rank)
        assert np.array_equal(s1, 2 * np.arange(5.0) + 1) and np.array_equal(s2, np.full(5, 3.0)) and cnt == 21
        q.put((rank, {k: v.numpy() for k, v in ens.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8])
def test_gather_is_independent_of_world_size(n_total):
    """the gathered ensemble on every rank equals the single-process result (ragged split too)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ref = {k: v.numpy() for k, v in _fake_history(list(range(n_total))).items()}
    for _rank, ens in got:
        for k in ref:
            assert ens[k].shape == ref[k].shape and np.array_equal(ens[k], ref[k]), k


def _ladder_worker(rank, world, port, n_local, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tonga_b200.tempering import geometric_ladder, swap_step_distributed
        rng = np.random.default_rng(42)
        E_all = rng.uniform(50.0, 400.0, world * n_local)            # every rank can regenerate the global energies ...
        beta_all = np.tile(geometric_ladder(T, 50.0), world * n_local // T)
        acc = att = 0
        for step in range(6):
            beta_all, a, t = swap_step_distributed(E_all[rank * n_local:(rank + 1) * n_local], beta_all, T, step, seed=13)  # ... but sends only its own
            acc += a; att += t
            E_all = E_all + 3.0 * np.sin(np.arange(len(E_all)) + step)   # new energies per sweep, the same on every rank
        q.put((rank, beta_all, acc, att))
    finally:
        dist.destroy_process_group()


def test_cross_rank_ladder_takes_the_single_rank_decisions():
    """BASELINE config 5 across GPUs: a 256-rung ladder split over 2 ranks (128 replicas each).  Every rank all-gathers the
    energies and applies the deterministic sweep; the betas must equal the single-process result bit for bit on both ranks."""
    from tonga_b200.tempering import geometric_ladder, swap_step
    world, n_local, T = 2, 128, 256
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ladder_worker, args=(r, world, port, n_local, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    rng = np.random.default_rng(42)
    E = rng.uniform(50.0, 400.0, world * n_local)
    beta = np.tile(geometric_ladder(T, 50.0), world * n_local // T)
    acc = att = 0
    for step in range(6):
        beta, a, t = swap_step(E, beta, T, step, seed=13)
        acc += a; att += t
        E = E + 3.0 * np.sin(np.arange(len(E)) + step)
    assert att == 3 * 128 + 3 * 127 and 0 < acc <= att
    for _rank, b, a, t in got:
        assert np.array_equal(b, beta) and (a, t) == (acc, att)


# ---------------------------------------------------------------------------------------------------- ray sharding (config 3)
def test_ray_shard_tile_ranges_partition_the_tiles():
    """tonga_shard_range (host arithmetic of the C ABI): contiguous, disjoint, covering, sizes within one tile of each other."""
    from tonga_b200.dist import shard_tiles
    for n_tiles in (0, 1, 5, 17, 2441):
        for w in (1, 2, 3, 8, 16):
            rng = [shard_tiles(n_tiles, w, r) for r in range(w)]
            assert rng[0][0] == 0 and rng[-1][1] == n_tiles
            for (a0, a1), (b0, b1) in zip(rng, rng[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in rng]
            assert max(sizes) - min(sizes) <= 1


def _handle_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tonga_b200.dist import exchange_handles
        mine = bytes((rank * 64 + i) % 251 for i in range(64))  # stands in for a cudaIpcMemHandle_t (64 opaque bytes)
        q.put((rank, exchange_handles(mine)))
    finally:
        dist.destroy_process_group()


def test_ipc_handle_exchange_over_gloo():
    """connect_ray_shards' host half: every rank ends up with every rank's 64-byte handle, in rank order."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_handle_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = [bytes((r * 64 + i) % 251 for i in range(64)) for r in range(world)]
    for _rank, hs in got:
        assert hs == want
