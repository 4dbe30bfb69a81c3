"""Parallel tempering over the chain batch (BASELINE.json config 5) -- an EXTENSION: the reference has no tempering
(SURVEY.md F6), so nothing here enters the parity claims.

Replicas keep their state; what is exchanged between adjacent rungs of a ladder is the inverse temperature beta
(swap temperatures, not states: SURVEY 5.8).  The sweep itself runs ON THE DEVICE (tonga_chains_temper_swap); swap_step below
is its NumPy mirror.  The tempered target of a replica is prior x L^beta with
-log L = E = phi/2 + R*log(noise) (Gaussian likelihood with the hierarchical noise factor; R data).  A swap of rungs i, j is
accepted with probability min(1, exp((beta_i - beta_j) * (E_i - E_j))).  Decisions are a pure function of
(E, beta, seed, step) drawn from a counter-based generator, so every rank of a multi-GPU run takes the same decisions
from all-gathered (E, beta) without further communication.  The device side of the extension is the per-chain beta the
sampler kernel applies to every misfit term (tonga_chains_set_beta); only beta == 1 replicas append to the history.
"""
from __future__ import annotations

import numpy as np


def geometric_ladder(n_temps: int, t_max: float = 50.0) -> np.ndarray:
    """beta_t = 1 / T_t with T geometric in [1, t_max] (config 5: 32 temperatures, T in [1, 50])."""
    if n_temps == 1:
        return np.ones(1)
    return 1.0 / (t_max ** (np.arange(n_temps) / (n_temps - 1)))


def energy(phi: np.ndarray, noise: np.ndarray, n_data: int) -> np.ndarray:
    return 0.5 * np.asarray(phi) + n_data * np.log(np.asarray(noise))


_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32(key: int, c0, c1, c2, c3):
    """Philox4x32-10 on NumPy arrays of counters (the library's device generator, tonga_internal.cuh) -> 4 uint32 arrays."""
    c = [np.asarray(v, dtype=np.uint64) & 0xFFFFFFFF for v in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c


def swap_uniforms(n_lad: int, ladder_size: int, step: int, seed: int) -> np.ndarray:
    """u[ladder, rung] in (0, 1) exactly as tg_temper_swap_kernel draws them: counter (step lo, step hi, ladder, rung | 1 << 30)."""
    lad, rung = np.meshgrid(np.arange(n_lad, dtype=np.uint64), np.arange(ladder_size, dtype=np.uint64), indexing="ij")
    w = philox4x32(seed & 0xFFFFFFFFFFFFFFFF, step & 0xFFFFFFFF, (step >> 32) & 0xFFFFFFFF, lad, rung | 0x40000000)
    return (((w[0] << np.uint64(32)) | w[1]) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53 + 2.0 ** -54


def swap_step(E: np.ndarray, beta: np.ndarray, ladder_size: int, step: int, seed: int = 0):
    """One even/odd sweep of swap attempts inside every ladder of `ladder_size` consecutive replicas -- the NumPy mirror of the
    device kernel (tg_temper_swap_kernel: same generator, same order of operations), kept for tests and documentation.

    E, beta: [n] (n % ladder_size == 0).  Pairs are adjacent in TEMPERATURE order (rank of beta inside the ladder), parity
    alternates with `step`.  Returns (new beta [n], accepted pairs, attempted pairs)."""
    E = np.asarray(E, dtype=np.float64)
    beta = np.array(beta, dtype=np.float64)
    n = len(beta)
    assert n % ladder_size == 0
    n_lad = n // ladder_size
    u = swap_uniforms(n_lad, ladder_size, step, seed)
    acc = att = 0
    for l in range(n_lad):
        sl = slice(l * ladder_size, (l + 1) * ladder_size)
        b, e = beta[sl], E[sl]
        order = np.argsort(-b, kind="stable")  # coldest (beta = 1) first
        for t in range(step % 2, ladder_size - 1, 2):
            i, j = order[t], order[t + 1]
            att += 1
            loga = (b[i] - b[j]) * (e[i] - e[j])
            if np.log(u[l, t]) < min(0.0, loga):
                b[i], b[j] = b[j], b[i]
                acc += 1
        beta[sl] = b
    return beta, acc, att


def run_tempered(chains, n_iter: int, ladder_size: int, swap_every: int = 100, t_max: float = 50.0, seed: int = 0, group=None):
    """Run `chains` for n_iter iterations with a swap sweep every `swap_every`, everything on the device: the sweep reads the
    replicas' phi / noise from device state (no host round trip); with a torch.distributed process group (one rank per GPU) the
    ladders span the ranks -- global replica r = rank * n + local index, ladder_size may exceed n -- and each sweep is preceded
    by ONE all-gather of the ranks' (phi, noise) scalars over NCCL.  Returns dict(swap_rate, beta (this rank's), accepted, attempted)."""
    n = chains.n
    world, rank = 1, 0
    if group is not None or _dist_ready():
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_all = n * world
    assert n_all % ladder_size == 0
    beta_all = np.tile(geometric_ladder(ladder_size, t_max), n_all // ladder_size)
    chains.set_beta(beta_all[rank * n:(rank + 1) * n])
    chains.temper_stats(reset=True)
    gathered = None
    if world > 1:
        gathered = _GatheredScalars(chains, world, rank, beta_all, group)
    done, step = 0, 0
    while done < n_iter:
        k = min(swap_every, n_iter - done)
        chains.run(k)
        done += k
        if gathered is None:
            chains.temper_swap(ladder_size, step, seed)
        else:
            gathered.swap(ladder_size, step, seed)
        step += 1
    acc, att = chains.temper_stats()
    return dict(swap_rate=acc / max(att, 1), beta=chains.get_beta(), accepted=acc, attempted=att)


def _dist_ready() -> bool:
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:
        return False


class _GatheredScalars:
    """Cross-rank ladders: device buffers [2][world * n] for the all-gathered (phi, noise) and [world * n] for beta (kept
    identical on every rank: each rank applies the same deterministic sweep to it)."""

    def __init__(self, chains, world, rank, beta_all, group):
        import torch
        from .dist import DeviceArray
        self.chains, self.world, self.rank, self.group = chains, world, rank, group
        n = chains.n
        dev = torch.device("cuda", chains.ctx.device)
        p = chains.scalar_ptrs()
        self.phi_v = torch.as_tensor(DeviceArray(p["phi"], (n,), "<f8"), device=dev)
        self.noise_v = torch.as_tensor(DeviceArray(p["noise"], (n,), "<f8"), device=dev)
        self.send = torch.empty((2, n), dtype=torch.float64, device=dev)
        self.recv = torch.empty((world * 2, n), dtype=torch.float64, device=dev)  # rank-major blocks of (phi, noise)
        self.all = torch.empty((2, world * n), dtype=torch.float64, device=dev)
        self.beta = torch.as_tensor(np.ascontiguousarray(beta_all), device=dev)
        self.torch = torch

    def swap(self, ladder_size, step, seed):
        import torch.distributed as dist
        torch, n = self.torch, self.chains.n
        self.chains.ctx.synchronize()  # the library's stream has produced phi / noise
        self.send[0].copy_(self.phi_v)
        self.send[1].copy_(self.noise_v)
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        self.all.copy_(self.recv.view(self.world, 2, n).permute(1, 0, 2).reshape(2, self.world * n))
        torch.cuda.current_stream().synchronize()  # the sweep runs on the library's stream
        self.chains.temper_swap(ladder_size, step, seed, n_all=self.world * n, phi_all=self.all[0].data_ptr(), noise_all=self.all[1].data_ptr(),
                                beta_all=self.beta.data_ptr(), offset=self.rank * n)


def swap_step_distributed(E_local: np.ndarray, beta_all: np.ndarray, ladder_size: int, step: int, seed: int = 0, group=None):
    """Host-side twin of the cross-rank sweep (CPU tests, gloo): all-gather the ranks' energies, apply swap_step to the global
    replica order (rank-major), return (beta_all, accepted, attempted) -- identical on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    send = torch.from_numpy(np.ascontiguousarray(E_local, dtype=np.float64))
    recv = torch.empty((world * send.shape[0],) + tuple(send.shape[1:]), dtype=torch.float64)
    dist.all_gather_into_tensor(recv, send, group=group)
    return swap_step(recv.reshape(-1).numpy(), beta_all, ladder_size, step, seed)
