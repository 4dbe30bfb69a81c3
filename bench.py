#!/usr/bin/env python
"""bench.py -- RJ-MCMC proposals/s (all chains, 381-ray Tonga) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--chains C] [--iters I] [--total-chains T] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the shipped 381-ray Tonga geometry (R = 381, P = 16 845 ray points, S = 16 464
segments), 1024 chains per GPU (chain-sharded, weak scaling: 1024*N chains on N GPUs; `--total-chains T` fixes the total
instead: strong scaling), uniform prior, K in [5, 100], incremental Voronoi update per proposal, proposals drawn on the
device (Philox).  One STEP = `--iters` proposals for every chain (default 1000, the reference's n_iter) STARTING FROM THE
COMMITTED WARM START MODELS (tonga_b200/datasets/warm_start_1024.npz: the chains' states after 20 000 iterations, mean
nCells 6.2) -- the same start models, iteration count and data in the B200 arm and in the CPU reference arm.
A proposal = one iteration of TD_inversion_function.jl:70.

  value  : proposals/s with the chain state resident in HBM when the timed region starts (device time of the sampler's
           kernels, CUDA events on the library's own stream, max over ranks);
  e2e    : the same metric through the host-buffer C-ABI calls a Julia caller makes per TD_inversion_function batch:
           upload the start models from host memory (+ full evaluate), run the iterations, download the thinned
           model_hist (nuclei, zeta, phi, ptS of every kept model) and the final models to host memory;
  roofline: dominant kernel = tg_sampler_kernel; algorithmic flops / bytes per proposal from SURVEY.md 8(d)
           (birth 16P+T, death 8P+T, move 24P+T, change T flop with T = 5S+4R; P owner bytes read per evaluated
           proposal + P written per accepted one), weighted by the measured action mix; `bound` names the larger fraction;
  cpu_baseline / `--impl reference`: the C oracle (a literal restatement of the reference algorithm; julia is not
           installed) on the host cores: one chain per thread as pmap does, a full evaluate per proposal;
  fresh  : the same step from build_starting (K log-uniform in [5, 100]) instead of the warm start, both arms;
  config3: BASELINE config 3 (100k synthetic rays, 2e7 points): full evaluate at K = 100 / 500 / 2000 and the STREAMED
           sampler's proposals/s, each with its SURVEY 8(d) roofline fraction;
  config5: BASELINE config 5 (32 temperatures x 32 ladders per GPU, sigma move, swap sweep on the device every 100
           iterations; ladders span the ranks under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "mcmc-in-tonga_b200"), os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "RJ-MCMC proposals/sec (all chains, 381-ray Tonga)"
UNIT = "proposals/s"
DATA = ("shipped 381-ray Tonga geometry (tonga_b200/datasets/tonga381.npz), slowness synthesised from ak135; start models "
        "tonga_b200/datasets/warm_start_1024.npz; device Philox proposals (B200 arm) / xoshiro (reference arm)")


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 - 0.3 <= t <= t1 + 0.3):
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "samples": len(sm), "reasons": sorted(reasons)}


def setup_data():
    from tonga_b200.data import load_tonga381, load_warm_start
    from tonga_b200.structs import define_TDstructrure
    p = define_TDstructrure()
    return load_tonga381(p=p), p, load_warm_start()


def workload_config(args, world, p):
    """The SAME dict in both arms: what one step is."""
    per_gpu = args.chains if args.total_chains <= 0 else args.total_chains // world
    return {"workload": "381-ray Tonga inversion (R=381, P=16845, S=16464), 1024 batched chains per B200, incremental Voronoi update per proposal",
            "chains_per_gpu": per_gpu, "iters_per_step": args.iters, "start_models": "warm_start_1024.npz (chain c starts from model c mod 1024)",
            "burn_in": args.iters // 2, "keep_each": int(p.keep_each), "kept_models_per_chain_per_step": (args.iters - args.iters // 2) // int(p.keep_each),
            "prior": "uniform", "cells": [p.min_cells, p.max_cells],
            "l2": "256 MB buffer written between timed steps (L2 flush)"}


def oracle_setup(ds, p, n_iter):
    import oracle as O
    box = (ds.xVec.min(), ds.xVec.max(), ds.yVec.min(), ds.yVec.max(), ds.zVec.min(), ds.zVec.max())
    op = O.make_params(box, sig=p.sig, zeta_scale=p.zeta_scale, max_sig=p.max_sig, n_iter=n_iter, burn_in=n_iter / 2, keep_each=p.keep_each,
                       min_cells=p.min_cells, max_cells=p.max_cells, prior=p.prior)
    od = O.Data(ds.rayX, ds.rayY, ds.rayZ, ds.rayL, ds.rayU, ds.tS, ds.allSig)
    return O, op, od


def cpu_reference_arm(ds, p, warm, n_iter, threads, chains=None, seed=1, fresh=False):
    """The reference algorithm on the host cores: the C oracle's chain farm (main_inversion.jl:15 pmap analogue), started from
    the warm start models (the B200 arm's) or, fresh=True, from build_starting."""
    O, op, od = oracle_setup(ds, p, n_iter)
    chains = chains or threads
    t0 = time.perf_counter()
    if fresh:
        O.chain_farm(op, od, chains, n_iter, threads, seed)
    else:
        idx = np.arange(chains) % len(warm["K"])
        O.chain_farm_from(op, od, warm["K"][idx], warm["cells"][idx], n_iter, threads, seed)
    dt = time.perf_counter() - t0
    return chains * n_iter / dt, dt, chains


def algorithmic_work(counts, P, S, R):
    """flops and state bytes for the measured action mix.  counts[n,3,5] = proposed/accepted/evaluated."""
    c = counts.sum(0).astype(np.float64)
    T = 5.0 * S + 4.0 * R
    per_eval = np.array([16.0 * P + T, 8.0 * P + T, T, 24.0 * P + T, 4.0 * R])  # birth, death, change, move, sigma
    flops = float((c[2] * per_eval).sum())
    bytes_ = float(c[2, :4].sum() * P + c[1, :4].sum() * P)
    return flops, bytes_, c


# ---------------------------------------------------------------------------------------------------- extra legs (B200 arm)
def leg_config3(local_rank, peaks, fp64_peak, fp32_peak, R=100000, chains=64, iters=30, world=1):
    """BASELINE config 3: synthetic 3-D model, 100k rays x ~200 points.  tS is synthesised on the GPU (forward model of a random
    200-nucleus model + noise); full evaluate at K = 100 / 500 / 2000; STREAMED sampler started at K = 1000."""
    import copy
    import torch
    from tonga_b200.api import Chains, Context, pack_models
    from tonga_b200.data import synthetic_rays
    from tonga_b200.structs import parameters
    p = parameters()
    t0 = time.perf_counter()
    ds = synthetic_rays(R, seed=3, p=p, n_true=0)
    rng = np.random.default_rng(33)
    box = (ds.xVec.min(), ds.xVec.max(), ds.yVec.min(), ds.yVec.max(), ds.zVec.min(), ds.zVec.max())
    rnd_model = lambda K: [rng.uniform(box[0], box[1], K), rng.uniform(box[2], box[3], K), rng.uniform(box[4], box[5], K), rng.uniform(0, 50, K)]
    ctx0 = Context(ds, p, device=local_rank, device_ingest=True)
    ds.tS = ctx0.evaluate(*rnd_model(200))["ptS"] + rng.normal(0.0, ds.allSig)  # "true" model -> observed t* (data synthesis, untimed)
    ctx0.close()
    p3 = copy.copy(p)
    p3.max_cells, p3.min_cells = 2000, 5
    p3.n_iter, p3.burn_in, p3.keep_each = 1e9, 1e9, 1.0  # no thinning in the timed loop
    ctx = Context(ds, p3, device=local_rank, device_ingest=True)
    t_setup = time.perf_counter() - t0
    out = {"rays": R, "points": int(ctx.P), "segments": int(ctx.S), "setup_s": t_setup, "evaluate": []}
    for K, n in ((100, 4), (500, 2), (2000, 1)):
        cells = np.stack([np.stack(rnd_model(K)) for _ in range(n)])
        Kd = torch.full((n,), K, dtype=torch.int32, device="cuda")
        cd = torch.from_numpy(cells).cuda()
        ptS = torch.zeros((n, ctx.R), dtype=torch.float64, device="cuda")
        phi = torch.zeros(n, dtype=torch.float64, device="cuda")
        run = lambda: ctx.evaluate_batch_dev(n, K, Kd.data_ptr(), cd.data_ptr(), None, ptS.data_ptr(), phi.data_ptr())
        run(); ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            run()
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / 3 / n
        flop = 8.0 * ctx.P * K + 5.0 * ctx.S + 4.0 * ctx.R  # SURVEY 8(d) F_full
        out["evaluate"].append({"K": K, "ms": 1e3 * dt, "pairs_per_s": ctx.P * K / dt, "algorithmic_tflops": flop / dt / 1e12,
                                "frac_fp32_fma_peak": flop / dt / 1e12 / fp32_peak if fp32_peak else None,
                                "frac_fp64_fma_peak": flop / dt / 1e12 / fp64_peak if fp64_peak else None})
    ch = Chains(ctx, chains, seed=11, hist_cap=0, sampler="streamed")
    Kp, cp = pack_models([rnd_model(1000) for _ in range(chains)], Kcap=ch.KC)
    ch.set_models(Kp, cp)
    ch.run(3)
    ch.reset()
    ch.run(iters)
    ms = ch.last_kernel_ms()
    _, counts = ch.stats()
    c = counts.sum(0).astype(np.float64)
    minimal_bytes = 2.0 * ctx.P * (c[2, :4].sum() + c[1, :4].sum())      # u16 owner state read per evaluated / written per accepted proposal
    own_bytes = ctx.P * (18.0 * (c[2, 0] + c[2, 3]) + 2.0 * (c[2, 1] + c[2, 2]) + 18.0 * (c[1, 0] + c[1, 3]) + 2.0 * c[1, 1])  # the kernel's own layout
    mm, dphi, dts = ch.verify()
    out["streamed"] = {"chains": chains, "K_start": 1000, "iterations": iters, "ms_per_iteration": ms / iters, "proposals_per_s": chains * iters / ms * 1e3,
                       "roofline": {"bound": "hbm", "achieved": minimal_bytes / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": minimal_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                    "note": "SURVEY 8(d) minimal bytes (2 B of owner state per point and evaluated / accepted proposal) / time.  The culled point "
                                            "pass (stream_cull.cuh) reads only the rays near the proposal's nucleus (18 B per point of a candidate ray): "
                                            "an unculled pass over the same proposals would move %.0f GB per iteration, the ncu launch list "
                                            "(profiles/r02_stream_cull_launches.md) shows 1.7 GB" % (own_bytes / iters / 1e9)},
                       "verify": {"owner_mismatch": mm, "max_dphi": dphi, "max_dtstar": dts}}
    st_un = ch.state(want_ptS=False)
    ch.close()
    if world > 1:
        # the SAME batch (same chains, seed, start models) ray-sharded over the ranks (SURVEY 8e alternative for config 3): every rank
        # keeps the per-point state of its own ray tiles; (t*, misfit term) of its rays go to every rank's exchange block by P2P
        # stores from inside the candidate pass (tonga_chains_shard_*; no NCCL on the data path).  Strong scaling of one batch.
        import torch.distributed as dist
        from tonga_b200.dist import connect_ray_shards, disconnect_ray_shards
        chs = Chains(ctx, chains, seed=11, hist_cap=0, sampler="streamed")
        connect_ray_shards(chs, local_rank)
        chs.set_models(Kp, cp)
        dist.barrier()
        chs.run(3)
        chs.reset()
        dist.barrier()
        chs.run(iters)
        ms_s = chs.last_kernel_ms()
        st_sh = chs.state(want_ptS=False)
        mm, dphi, dts = chs.verify()
        info = chs.shard_info()
        out["ray_sharded"] = {"chains": chains, "iterations": iters, "ms_per_iteration_rank": ms_s / iters, "own_points": info["point1"] - info["point0"],
                              "bit_identical_to_unsharded": bool(st_sh["phi"].tobytes() == st_un["phi"].tobytes() and np.array_equal(st_sh["K"], st_un["K"])
                                                                 and st_sh["cells"].tobytes() == st_un["cells"].tobytes()),
                              "verify_own_points": {"owner_mismatch": mm, "max_dphi": dphi, "max_dtstar": dts}}
        disconnect_ray_shards(chs)
        chs.close()
    ctx.close()
    return out


def leg_config5(ds, p0, warm, local_rank, rank, world, iters=1000, swap_every=100):
    """BASELINE config 5: 32 temperatures x 32 ladders per GPU (T geometric in [1, 50]), hierarchical sigma move with p = 1/5,
    one swap sweep on the device every 100 iterations; under torchrun the ladders span the ranks (32 * world rungs, one NCCL
    all-gather of (phi, noise) per sweep)."""
    import copy
    import torch
    from tonga_b200.api import Chains, Context
    from tonga_b200.tempering import run_tempered
    p = copy.copy(p0)
    p.max_sig = 2.0
    p.n_iter, p.burn_in = float(iters), float(iters // 2)
    ctx = Context(ds, p, device=local_rank, n_actions=5)
    n = 1024
    ch = Chains(ctx, n, chain_id0=rank * n, seed=555, hist_cap=(iters - iters // 2) // int(p.keep_each) + 1)
    ch.set_models(warm["K"][np.arange(n) % len(warm["K"])], warm["cells"][np.arange(n) % len(warm["K"])])
    ladder = 32 * world
    run_tempered(ch, iters, ladder_size=ladder, swap_every=swap_every, t_max=50.0, seed=7)  # warm-up (ladders equilibrate)
    ch.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = run_tempered(ch, iters, ladder_size=ladder, swap_every=swap_every, t_max=50.0, seed=8)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    _, counts = ch.stats()
    c = counts.sum(0)
    hist_n = ch.history(want_ptS=False)["n_hist"]
    out = {"chains_per_gpu": n, "temperatures": 32, "ladders_per_gpu": 32, "ladder_rungs": ladder, "T_max": 50.0, "iterations": iters, "swap_every": swap_every,
           "proposals_per_s_rank": n * iters / dt, "swap_rate": info["swap_rate"], "swap_attempts": info["attempted"],
           "sigma_move": {"proposed": int(c[0, 4]), "accepted": int(c[1, 4])}, "cold_replicas_kept_models": int(hist_n.sum()),
           "ms": 1e3 * dt, "verify": list(ch.verify())}
    ch.close()
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chains", type=int, default=1024, help="chains per GPU (weak scaling)")
    ap.add_argument("--total-chains", type=int, default=0, help="fixed total number of chains, split over the GPUs (strong scaling)")
    ap.add_argument("--iters", type=int, default=1000, help="proposals per chain per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the fresh-start / config-3 / config-5 legs")
    ap.add_argument("--seed", type=int, default=20260000)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    ds, p, warm = setup_data()
    config = workload_config(args, world, p)

    # ------------------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        ref_chains = 8 * ncores  # a bounded sample of the workload per step: 8 chains per host thread, the B200 arm's per-chain work
        for _ in range(max(args.warmup, 0)):
            cpu_reference_arm(ds, p, warm, max(args.iters // 10, 50), ncores)
        tot_props, tot_t = 0.0, 0.0
        for _ in range(args.steps):
            rate, dt, chains = cpu_reference_arm(ds, p, warm, args.iters, ncores, chains=ref_chains)
            tot_props += chains * args.iters
            tot_t += dt
        val = tot_props / tot_t
        rate_fresh, _, _ = cpu_reference_arm(ds, p, warm, args.iters, ncores, fresh=True)
        sample = (f"{ref_chains} chains (warm start models 0..{ref_chains - 1}) x {args.iters} iterations per step on {ncores} host threads "
                  "(one chain per thread, as pmap does); the per-chain work is the B200 arm's: same start models, same iteration count")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong" if args.total_chains > 0 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": DATA, "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample, "fresh_start_value": rate_fresh,
                             "note": "reference algorithm = C restatement of MCsub.jl / TD_inversion_function.jl (julia not installed); a full evaluate per proposal as the reference does"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from tonga_b200.api import Chains, Context, pinned_empty
    p.n_iter, p.burn_in = float(args.iters), float(args.iters // 2)  # reference ratio: burn-in = n_iter/2, keep every 10th
    ctx = Context(ds, p, device=local_rank)
    n = config["chains_per_gpu"]
    idx = (rank * n + np.arange(n)) % len(warm["K"])
    K0 = pinned_empty((n,), np.int32); K0[:] = warm["K"][idx]
    cells0 = pinned_empty((n, 4, warm["cells"].shape[2]), np.float64); cells0[:] = warm["cells"][idx]
    ch = Chains(ctx, n, chain_id0=rank * n, seed=args.seed)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm ("value"): every step starts from the warm start models, resident in HBM before the timed region
    for _ in range(args.warmup):
        ch.reset(); ch.set_models(K0, cells0); ch.run(args.iters)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t_wall0 = time.time()
    dev_ms = 0.0
    counts = np.zeros((n, 3, 5), np.int64)
    for _ in range(args.steps):
        ch.reset()  # every step is a fresh n_iter-long run: burn-in, thinning and history writes included
        ch.set_models(K0, cells0)  # untimed: owners, t*, phi of the start models established on the device
        flush.zero_()
        torch.cuda.synchronize()
        ch.run(args.iters)
        dev_ms += ch.last_kernel_ms()
        counts += ch.stats()[1]
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    mm, dphi, dts = ch.verify()  # incremental state still equals a full evaluate after the timed region
    st_end = ch.state(want_ptS=False)
    dev_ms_max = max_over_ranks(dev_ms)
    total_props = float(n) * world * args.iters * args.steps
    value = total_props / (dev_ms_max * 1e-3)

    # ---- posterior of BASELINE config 4 (outside the timed regions): maps + gather of the kept models
    from tonga_b200 import api as _api, dist as _tdist
    t0 = time.perf_counter()
    acc = [ch.raster(X, Y, Z) for kind, l0, shape, X, Y, Z in _api.slice_nodes(ds, p)]
    t1 = time.perf_counter()
    posterior = {"slices": len(acc), "nodes": int(sum(len(a_[0]) for a_ in acc)), "raster_ms": 1e3 * (t1 - t0), "backend": "nccl" if world > 1 else "none"}
    if world > 1:
        dev = torch.device("cuda", local_rank)
        red = _tdist.allreduce_accumulators(acc, dev)            # warm-up call (communicator set-up)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); red = _tdist.allreduce_accumulators(acc, dev); e1.record(); torch.cuda.synchronize()
        posterior.update({"allreduce_ms": e0.elapsed_time(e1), "kept_models_all_ranks": int(red[0][2])})
        packed = _tdist.pack_history(ch, dev, last=50)             # [n, last, 2 + 4 kmax] doubles: K, phi, the valid nuclei only
        out_buf = torch.empty((world * packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=dev)
        dist.all_gather_into_tensor(out_buf, packed)               # warm-up
        torch.cuda.synchronize(); dist.barrier()
        e0.record(); dist.all_gather_into_tensor(out_buf, packed); e1.record(); torch.cuda.synchronize()
        g_ms = max_over_ranks(e0.elapsed_time(e1))
        nbytes = packed.numel() * packed.element_size()
        posterior.update({"gather_ms": g_ms, "gather_bytes_per_rank": int(nbytes), "gather_recv_GBps_per_rank": nbytes * (world - 1) / (g_ms * 1e-3) / 1e9,
                          "gathered_chains": int(out_buf.shape[0]), "packed_record_doubles": int(packed.shape[2])})
    else:
        posterior["kept_models_all_ranks"] = int(acc[0][2]) if acc else 0

    # ---- end-to-end arm through host buffers
    # The e2e batch keeps its history in mapped page-locked host memory (TONGA_HISTORY_ON_HOST): the sampler stores every kept
    # model straight into host memory while it runs, so the step's model_hist is on the host when the run returns.
    che = Chains(ctx, n, chain_id0=rank * n, seed=args.seed, host_history=True)
    _, state_buf = che.alloc_buffers(pinned=True)  # page-locked host buffers for the final models, reused every step
    h2d = K0.nbytes + cells0.nbytes
    d2h = 0
    e2e_t = 0.0
    for i in range(2 + args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        che.reset()
        che.set_models(K0, cells0)         # H2D of the start models + full evaluate
        che.run(args.iters)                # the proposal loop; kept models stream to host memory as they are produced
        hist = che.history()               # model_hist (nuclei, zeta, phi, ptS of every kept model): already in host memory
        fin = che.state(out=state_buf)     # D2H of the final models
        chk = float(hist["phi"][:, 0].sum()) + float(fin["phi"].sum())  # the host reads the step's result
        t1 = time.perf_counter()
        tt = max_over_ranks(t1 - t0)
        if i >= 2:
            e2e_t += tt
        nh = np.minimum(hist["n_hist"], che.hist_cap)
        kept_nuclei = int(sum(int(hist["K"][c, :nh[c]].sum()) for c in range(n)))
        d2h = 32 * kept_nuclei + int(nh.sum()) * (8 * ctx.R + 36) + sum(v.nbytes for v in fin.values() if v is not None)  # bytes the GPU wrote to host memory
    assert np.isfinite(chk)
    e2e_val = float(n) * world * args.iters * args.steps / e2e_t
    che.close()

    out = None
    if rank == 0:
        peaks, peaks_src = load_peaks()
        fp64_peak, fp32_peak = ctx.peak_flops()
        flops, bytes_, c = algorithmic_work(counts, ctx.P, ctx.S, ctx.R)
        sec = dev_ms * 1e-3  # rank 0's own kernel time for rank 0's own chains
        ach_tf, ach_gb = flops / sec / 1e12, bytes_ / sec / 1e9
        frac_fp = ach_tf / fp64_peak if fp64_peak > 0 else 0.0
        frac_hbm = ach_gb / peaks["hbm_gbs"]
        props_rank = float(n) * args.iters * args.steps
        traffic_ncu = None
        try:  # DRAM bytes per proposal from this round's committed ncu --set full capture (NOT measured in this run)
            traffic_ncu = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        except Exception:
            pass
        fp = {"achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": frac_fp,
              "peak_source": "FMA-chain microbenchmark run in this process (tonga_peak_flops); FMA = 2 flop"}
        hb = {"achieved": ach_gb, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": frac_hbm, "peak_source": f"MEASURED_PEAKS.json ({peaks_src})"}
        top = fp if frac_fp >= frac_hbm else hb  # report both, name the larger as the bound
        roof = {"bound": "fp64" if top is fp else "hbm", "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
                "traffic": (traffic_ncu["dram_bytes_read_per_launch"] + traffic_ncu["dram_bytes_write_per_launch"]) if traffic_ncu else None,
                "traffic_ncu_capture": traffic_ncu, "peak_source": top["peak_source"],
                "kernel": "tg_sampler_kernel", "launch_ms": dev_ms / args.steps,
                "algorithmic_bytes_per_proposal": bytes_ / props_rank, "algorithmic_flop_per_proposal": flops / props_rank,
                "fp64": fp, "hbm": hb, "fp32_peak_tflops": fp32_peak,
                "note": "contraction depth 3: no tensor cores; the chain state (owners) is shared-memory resident by design, so the HBM figure is "
                        "algorithmic bytes / time; `traffic` = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (1024 chains x 1000 iterations, without history writes) from the committed ncu --set full capture (traffic_ncu_capture), not measured in this run"}
        acc_rate = (c[1, :4] / np.maximum(c[0, :4], 1)).round(4).tolist()
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if args.total_chains > 0 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": DATA, "config": config,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_t / args.steps},
            "gpu_launches": 3 * args.steps,  # per step: tg_pregen_kernel (raw draws), tg_order_kernel (launch order by nCells), tg_sampler_kernel
            "clocks": clocks, "roofline": roof, "chains_total": n * world, "parallelism": f"chain-sharded x{world}",
            "evaluates_per_s": value * float(c[2].sum() / max(c[0].sum(), 1)),  # proposals that needed the forward model (a-priori rejects excluded; rank 0's mix)
            "acceptance": {"birth_death_change_move": acc_rate, "evaluated_fraction": float(c[2].sum() / max(c[0].sum(), 1)),
                           "mean_cells_start": float(K0.mean()), "mean_cells_end": float(st_end["K"].mean())},
            "verify": {"owner_mismatch": mm, "max_dphi": dphi, "max_dtstar": dts},
            "posterior": posterior,
        }
    # ---- extra legs (outside the headline's timed regions)
    if not args.no_extra_legs:
        # fresh start: the same step from build_starting (K log-uniform in [5, 100]) instead of the warm start
        fr_ms = 0.0
        for i in range(3):
            ch.reset(); ch.build_starting(); flush.zero_(); torch.cuda.synchronize()
            ch.run(args.iters)
            if i > 0:
                fr_ms += ch.last_kernel_ms() / 2
        fr = float(n) * world * args.iters / (max_over_ranks(fr_ms) * 1e-3)
        if rank == 0:
            out["fresh"] = {"value": fr, "unit": UNIT, "note": "same step started from build_starting (MCsub.jl:76-121) instead of the warm start models"}
        peaks, _ = load_peaks()
        fp64_peak, fp32_peak = ctx.peak_flops()
        c5 = leg_config5(ds, p, warm, local_rank, rank, world, iters=args.iters)
        c5_rate = float(n) * world * args.iters / (max_over_ranks(c5["ms"]) * 1e-3)
        c3 = leg_config3(local_rank, peaks, fp64_peak, fp32_peak, world=world)
        c3_ms = max_over_ranks(c3["streamed"]["ms_per_iteration"])
        if world > 1:
            rs_ms = max_over_ranks(c3["ray_sharded"]["ms_per_iteration_rank"])
            rs_ok = max_over_ranks(0.0 if c3["ray_sharded"]["bit_identical_to_unsharded"] else 1.0) == 0.0
        if rank == 0:
            c5["proposals_per_s"] = c5_rate
            c3["streamed"]["proposals_per_s_all_gpus"] = c3["streamed"]["chains"] * world / c3_ms * 1e3
            c3["streamed"]["parallelism"] = f"chain-sharded x{world} ({c3['streamed']['chains']} chains per GPU, ray set replicated)"
            if world > 1:
                rs = c3["ray_sharded"]
                rs.update({"ms_per_iteration": rs_ms, "proposals_per_s": rs["chains"] / rs_ms * 1e3, "bit_identical_to_unsharded_all_ranks": rs_ok,
                           "speedup_vs_one_gpu": c3["streamed"]["ms_per_iteration"] / rs_ms,
                           "parallelism": f"ray-sharded x{world}: ONE batch of {rs['chains']} chains, per-point state split by ray tiles; per proposal the "
                                          "candidate pass stores (ray, t*, misfit term) of its dirty rays into every rank's exchange block (P2P stores over "
                                          "NVLink, CUDA IPC mappings), sequence flags + bounded wait in the accept kernel; no NCCL on the data path"})
            out["config5"], out["config3"] = c5, c3
    if rank == 0:
        if not args.no_cpu_baseline:
            rate, dt, chains = cpu_reference_arm(ds, p, warm, args.iters, ncores, chains=48 * ncores)   # ~10 s of host work
            rate1, dt1, _ = cpu_reference_arm(ds, p, warm, args.iters, 1, chains=16)
            ratef, _, _ = cpu_reference_arm(ds, p, warm, args.iters, ncores, fresh=True)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": ncores, "kind": "port",
                                   "sample": f"{chains} chains (warm start models 0..{chains - 1}) x {args.iters} iterations on {ncores} threads ({dt:.1f} s); "
                                             f"single core: {rate1:.1f} proposals/s; from build_starting: {ratef:.1f} proposals/s",
                                   "single_core_value": rate1, "fresh_start_value": ratef}
        print(json.dumps(out))
    ch.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
