"""Parallel tempering over the chain batch (BASELINE.json config 5) -- an EXTENSION: the reference has no tempering
(SURVEY.md F6), so nothing here enters the parity claims.

Replicas keep their state; what is exchanged between adjacent rungs of a ladder is the inverse temperature beta
(swap temperatures, not states: SURVEY 5.8).  The tempered target of a replica is prior x L^beta with
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


def swap_step(E: np.ndarray, beta: np.ndarray, ladder_size: int, step: int, seed: int = 0):
    """One even/odd sweep of swap attempts inside every ladder of `ladder_size` consecutive replicas.

    E, beta: [n] (n % ladder_size == 0).  Pairs are adjacent in TEMPERATURE order (rank of beta inside the ladder), parity
    alternates with `step`.  Returns (new beta [n], accepted pairs, attempted pairs)."""
    E = np.asarray(E, dtype=np.float64)
    beta = np.array(beta, dtype=np.float64)
    n = len(beta)
    assert n % ladder_size == 0
    n_lad = n // ladder_size
    u = np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, step & 0xFFFFFFFFFFFFFFFF])).random((n_lad, ladder_size))
    acc = att = 0
    for l in range(n_lad):
        sl = slice(l * ladder_size, (l + 1) * ladder_size)
        b, e = beta[sl], E[sl]
        order = np.argsort(-b, kind="stable")  # coldest (beta = 1) first
        for t in range(step % 2, ladder_size - 1, 2):
            i, j = order[t], order[t + 1]
            att += 1
            loga = (b[i] - b[j]) * (e[i] - e[j])
            if np.log(max(u[l, t], 1e-300)) < min(0.0, loga):
                b[i], b[j] = b[j], b[i]
                acc += 1
        beta[sl] = b
    return beta, acc, att


def run_tempered(chains, n_iter: int, ladder_size: int, swap_every: int = 100, t_max: float = 50.0, seed: int = 0,
                 gather=None):
    """Run `chains` (n % ladder_size == 0 replicas) for n_iter iterations with a swap sweep every `swap_every`.

    gather(E, beta) -> (E_all, beta_all, offset) lets a multi-GPU caller all-gather the scalars so that ladders may span
    ranks; by default ladders are local.  Returns dict(swap_rate, beta)."""
    n = chains.n
    assert n % ladder_size == 0
    beta = np.tile(geometric_ladder(ladder_size, t_max), n // ladder_size)
    chains.set_beta(beta)
    R = chains.ctx.R
    done, step, acc, att = 0, 0, 0, 0
    while done < n_iter:
        k = min(swap_every, n_iter - done)
        chains.run(k)
        done += k
        st = chains.state(want_ptS=False)
        E = energy(st["phi"], st["noise"], R)
        if gather is None:
            beta, a, t = swap_step(E, beta, ladder_size, step, seed)
        else:
            E_all, beta_all, off = gather(E, beta)
            beta_all, a, t = swap_step(E_all, beta_all, ladder_size, step, seed)
            beta = beta_all[off:off + n]
        acc += a
        att += t
        step += 1
        chains.set_beta(beta)
    return dict(swap_rate=acc / max(att, 1), beta=beta)
