"""Chain sharding across the GPUs of one box + gathering the posterior ensembles.

The reference's only parallelism is `pmap` over independent chains (main_inversion.jl:15): chains never interact, so the
path shards with NO data-path collective.  One process per GPU (torchrun); rank g owns the contiguous block of chains
`shard(n_total, world, g)`; device Philox streams are keyed by the GLOBAL chain id, so every chain produces the same
samples whatever the GPU count.  torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is used only to
gather the thinned ensembles (model_hist) -- the analogue of pmap returning `models` to the master.
"""
from __future__ import annotations

import numpy as np


def shard(n_total: int, world: int, rank: int):
    """Contiguous block of chains for `rank`: (first global chain index, count).  Sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


class DeviceArray:
    """Zero-copy view of library-owned device memory for torch (`torch.as_tensor(DeviceArray(...), device='cuda')`)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def history_tensors(chains, device):
    """The packed on-device history of a `Chains` batch as torch tensors (zero-copy views of the library's buffers)."""
    import torch
    p = chains.device_ptrs()
    n, H, KC, R = chains.n, chains.hist_cap, chains.KC, chains.ctx.R
    mk = lambda key, shape, ts: torch.as_tensor(DeviceArray(p[key], shape, ts), device=device)
    return {"n_hist": mk("n_hist", (n,), "<i4"), "K": mk("hist_K", (n, H), "<i4"), "cells": mk("hist_cells", (n, H, 4, KC), "<f8"),
            "phi": mk("hist_phi", (n, H), "<f8"), "ptS": mk("hist_ptS", (n, H, R), "<f8")}


def gather_ensembles(local: dict, counts, group=None):
    """All-gather per-rank ensembles whose leading dimension is the rank's chain count (ragged across ranks).

    local : dict name -> tensor [n_local, ...] (CPU tensors under gloo, CUDA tensors under NCCL)
    counts: chains per rank (len = world).  Returns dict name -> tensor [sum(counts), ...] in global chain order.
    Ragged blocks are padded to max(counts) for the collective (all_gather needs equal shapes) and trimmed after."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return {k: v.clone() for k, v in local.items()}
    nmax, even = int(max(counts)), len(set(int(c) for c in counts)) == 1
    out = {}
    for k, v in local.items():
        if even:  # the usual case: one collective straight into the result, no padding, no concatenation
            send = v.contiguous()
            res = torch.empty((world * nmax,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            dist.all_gather_into_tensor(res, send, group=group)
            out[k] = res
            continue
        pad = torch.zeros((nmax,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        pad[:v.shape[0]] = v
        buf = torch.empty((world * nmax,) + tuple(pad.shape[1:]), dtype=v.dtype, device=v.device)  # gloo wants the concatenated shape
        dist.all_gather_into_tensor(buf, pad, group=group)
        buf = buf.view((world, nmax) + tuple(pad.shape[1:]))
        out[k] = torch.cat([buf[r, :int(c)] for r, c in enumerate(counts)], dim=0)
    return out


def pack_history(chains, device, last: int = 50):
    """The last `last` kept models of every chain of this rank as ONE dense tensor [n, last, 2 + 4 kmax] of doubles
    (K, phi, then x / y / z / zeta of the valid nuclei; kmax = largest nCells among them) -- the record that is gathered for a
    model.jld-equivalent ensemble.  The fixed-capacity device history holds 4 KC doubles of nuclei per model whatever K is."""
    import torch
    hv = history_tensors(chains, device)
    last = max(0, min(int(chains.hist_cap), int(last)))
    K = hv["K"][:, :last]
    kmax = max(int(K.max().item()) if K.numel() else 1, 1)
    rec = torch.empty((chains.n, last, 2 + 4 * kmax), dtype=torch.float64, device=device)
    rec[:, :, 0] = K.to(torch.float64)
    rec[:, :, 1] = hv["phi"][:, :last]
    rec[:, :, 2:] = hv["cells"][:, :last, :, :kmax].reshape(chains.n, last, 4 * kmax)
    return rec


def allreduce_accumulators(acc, device, group=None):
    """Sum the {sum, sum^2, count} accumulators of tonga_chains_raster over the ranks: ONE all-reduce for all slices.
    acc: list of (s1, s2, count) per slice -> the same list, summed."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return acc
    flat = np.concatenate([np.concatenate([s1, s2, [float(c)]]) for s1, s2, c in acc])
    t = torch.from_numpy(flat).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t = t.cpu().numpy()
    out, o = [], 0
    for s1, s2, c in acc:
        m = len(s1)
        out.append((t[o:o + m], t[o + m:o + 2 * m], int(round(t[o + 2 * m]))))
        o += 2 * m + 1
    return out


def run_sharded(TD_parameters, dataStruct, n_chains_total: int, seed: int = 20260000, first_chain: int = 1, gather: bool = True):
    """The whole chain farm on all ranks of the current process group: build_starting, n_iter iterations, thinning, then
    (optionally) the gathered ensemble on every rank.  Call under torchrun with one process per GPU."""
    import os
    import torch
    import torch.distributed as dist
    from .api import Chains, Context
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    start, count = shard(n_chains_total, world, rank)
    ctx = Context(dataStruct, TD_parameters, device=local_rank)
    ch = Chains(ctx, count, chain_id0=first_chain + start, seed=seed)
    ch.build_starting()
    ch.run(int(TD_parameters.n_iter))
    if not gather:
        return ch, None
    dev = torch.device("cuda", local_rank)
    local = {k: v.clone() for k, v in history_tensors(ch, dev).items()}
    counts = [shard(n_chains_total, world, r)[1] for r in range(world)]
    return ch, gather_ensembles(local, counts)


def ensemble_to_numpy(ens: dict) -> dict:
    return {k: v.detach().cpu().numpy() for k, v in ens.items()}


def allreduce_sums(s1, s2, count, device=None, group=None):
    """Sum the posterior accumulators of tonga_chains_raster across ranks (plain sums -> one all-reduce)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return s1, s2, count
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    t = torch.from_numpy(np.concatenate([s1, s2, [float(count)]])).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t = t.cpu().numpy()
    n = len(s1)
    return t[:n], t[n:2 * n], int(round(t[-1]))


# ---------------------------------------------------------------------------------------------------- ray sharding (config 3)
def shard_tiles(n_tiles: int, world: int, rank: int):
    """Own tile range [t0, t1) of `rank` in a ray-sharded streamed batch (the library's tonga_shard_range: host arithmetic)."""
    import ctypes as C
    from . import _lib
    t0, t1 = C.c_int32(), C.c_int32()
    _lib.check(_lib.load().tonga_shard_range(int(n_tiles), int(rank), int(world), C.byref(t0), C.byref(t1)))
    return t0.value, t1.value


def exchange_handles(handle: bytes, group=None):
    """All-gather one 64-byte CUDA IPC handle per rank -> list of `world` handles (works under gloo and NCCL)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    out = torch.empty(world * len(handle), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    flat = bytes(out.cpu().numpy().tobytes())
    return [flat[i * len(handle):(i + 1) * len(handle)] for i in range(world)]


def connect_ray_shards(chains, device: int, group=None):
    """Turn `chains` (a STREAMED batch created identically on every rank of the process group) into this rank's part of a
    ray-sharded batch: the ranks exchange the CUDA IPC handles of their exchange blocks (64 bytes each) and map each other's
    blocks, so that the candidate pass can store its rays' (t*, misfit term) straight into every peer over NVLink."""
    import ctypes as C
    import torch.distributed as dist
    from . import _lib
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    base, _ = chains.shard_init(rank, world)
    h = C.create_string_buffer(64)
    _lib.check(lib.tonga_ipc_export(C.c_void_p(base), h))
    handles = exchange_handles(h.raw, group)
    bases = []
    for r in range(world):
        if r == rank:
            bases.append(base)
            continue
        q = C.c_void_p()
        _lib.check(lib.tonga_ipc_open(int(device), handles[r], C.byref(q)))
        bases.append(q.value)
    chains.shard_connect(bases)
    chains._shard_peer_maps = [(int(device), b) for r, b in enumerate(bases) if r != rank]  # closed by disconnect_ray_shards
    dist.barrier(group)
    return bases


def disconnect_ray_shards(chains, group=None):
    """Unmap the peers' exchange blocks (call on every rank before the batches are destroyed)."""
    import ctypes as C
    import torch.distributed as dist
    from . import _lib
    lib = _lib.load()
    dist.barrier(group)  # nobody is still storing into a block that is about to go away
    for device, b in getattr(chains, "_shard_peer_maps", []):
        _lib.check(lib.tonga_ipc_close(device, C.c_void_p(b)))
    chains._shard_peer_maps = []
