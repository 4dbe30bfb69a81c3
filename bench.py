#!/usr/bin/env python
"""bench.py -- RJ-MCMC proposals/s (all chains, 381-ray Tonga) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--chains C] [--iters I] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the shipped 381-ray Tonga geometry (R = 381, P = 16 845 ray points, S = 16 464
segments), 1024 chains per GPU (chain-sharded, weak scaling: 1024*N chains on N GPUs), uniform prior, K in [5, 100],
incremental Voronoi update per proposal, proposals drawn on the device (Philox).  One STEP = `--iters` proposals
for every chain (default 1000, the reference's n_iter).  A proposal = one iteration of TD_inversion_function.jl:70.

  value  : proposals/s with the chain state resident in HBM when the timed region starts (device time of the sampler
           kernel, CUDA events on the library's own stream, max over ranks);
  e2e    : the same metric through the host-buffer C-ABI calls a Julia caller makes per TD_inversion_function batch:
           upload the start models from host memory (+ full evaluate), run the iterations, download the thinned
           model_hist (nuclei, zeta, phi, ptS of every kept model) and the final models to host memory;
  roofline: dominant kernel = tg_sampler_kernel; algorithmic flops / bytes per proposal from SURVEY.md 8(d)
           (birth 16P+T, death 8P+T, move 24P+T, change T flop with T = 5S+4R; P owner bytes read per evaluated
           proposal + P written per accepted one), weighted by the measured action mix;
  cpu_baseline: the C oracle (a literal restatement of the reference algorithm; julia is not installed) on the host
           cores, bounded sample.  `--impl reference` prints that arm alone.
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
WORKLOAD = "381-ray Tonga inversion (R=381, P=16845, S=16464), 1024 batched chains per B200, incremental Voronoi update per proposal"


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
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
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
    from tonga_b200.data import load_tonga381
    from tonga_b200.structs import define_TDstructrure
    p = define_TDstructrure()
    return load_tonga381(p=p), p


def cpu_reference_arm(ds, p, n_iter, threads, chains=None, seed=1):
    """The reference algorithm on the host cores: the C oracle's chain farm (main_inversion.jl:15 pmap analogue)."""
    import oracle as O
    box = (ds.xVec.min(), ds.xVec.max(), ds.yVec.min(), ds.yVec.max(), ds.zVec.min(), ds.zVec.max())
    op = O.make_params(box, sig=p.sig, zeta_scale=p.zeta_scale, max_sig=p.max_sig, n_iter=n_iter, burn_in=n_iter / 2, keep_each=p.keep_each,
                       min_cells=p.min_cells, max_cells=p.max_cells, prior=p.prior)
    od = O.Data(ds.rayX, ds.rayY, ds.rayZ, ds.rayL, ds.rayU, ds.tS, ds.allSig)
    chains = chains or threads
    t0 = time.perf_counter()
    O.chain_farm(op, od, chains, n_iter, threads, seed)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chains", type=int, default=1024, help="chains per GPU")
    ap.add_argument("--iters", type=int, default=1000, help="proposals per chain per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-iters", type=int, default=2500, help="iterations per chain of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=20260000)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    ds, p = setup_data()

    # ------------------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        for _ in range(max(args.warmup, 0)):
            cpu_reference_arm(ds, p, max(args.cpu_iters // 10, 50), ncores)
        tot_props, tot_t = 0.0, 0.0
        for _ in range(args.steps):
            rate, dt, chains = cpu_reference_arm(ds, p, args.cpu_iters, ncores)
            tot_props += chains * args.cpu_iters
            tot_t += dt
        val = tot_props / tot_t
        sample = f"{ncores} chains x {args.cpu_iters} iterations per step on {ncores} host threads (one chain per thread, as pmap does)"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "shipped 381-ray Tonga geometry (tonga_b200/datasets/tonga381.npz), slowness synthesised from ak135",
            "config": {"workload": WORKLOAD, "note": "reference algorithm = C restatement of MCsub.jl/TD_inversion_function.jl (julia not installed); full evaluate per proposal as the reference does"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port", "sample": sample},
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

    from tonga_b200.api import Chains, Context
    p.n_iter, p.burn_in = float(args.iters), float(args.iters // 2)  # reference ratio: burn-in = n_iter/2, keep every 10th
    ctx = Context(ds, p, device=local_rank)
    n = args.chains
    ch = Chains(ctx, n, chain_id0=rank * n, seed=args.seed)
    ch.build_starting()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ("value")
    for _ in range(args.warmup):
        ch.run(args.iters)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    t_wall0 = time.time()
    dev_ms = 0.0
    counts = np.zeros((n, 3, 5), np.int64)
    for _ in range(args.steps):
        flush.zero_()
        ch.reset()  # every step is a fresh n_iter-long run: burn-in, thinning and history writes included
        torch.cuda.synchronize()
        ch.run(args.iters)
        dev_ms += ch.last_kernel_ms()
        counts += ch.stats()[1]
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    mm, dphi, dts = ch.verify()  # incremental state still equals a full evaluate after the timed region
    # posterior maps of BASELINE config 4 (outside the timed regions): rasterise the kept models of the last step on the slices of
    # plot_model_hist (MCsub.jl:753-825) and sum the {count, sum, sum^2} accumulators over the ranks (the path's only collective)
    from tonga_b200 import api as _api, dist as _tdist
    t0 = time.perf_counter()
    acc = []
    for kind, l0, shape, X, Y, Z in _api.slice_nodes(ds, p):
        acc.append(ch.raster(X, Y, Z))
    t1 = time.perf_counter()
    n_nodes = sum(len(a_[0]) for a_ in acc)
    if world > 1:
        dist.barrier()
    t2 = time.perf_counter()
    red = [_tdist.allreduce_sums(s1, s2, cnt, device="cuda") for s1, s2, cnt in acc]
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    posterior = {"slices": len(acc), "nodes": int(n_nodes), "kept_models_all_ranks": int(red[0][2]) if red else 0,
                 "raster_ms": 1e3 * (t1 - t0), "allreduce_ms": 1e3 * (t3 - t2), "backend": "nccl" if world > 1 else "none"}
    if world > 1:  # config 4: gather of the packed kept models (K, nuclei, phi of the last <= 50 kept models of every chain) to all ranks
        hv = _tdist.history_tensors(ch, torch.device("cuda", local_rank))
        last = max(0, min(int(ch.hist_cap), 50))
        local = {k: hv[k][:, :last].contiguous() for k in ("K", "cells", "phi")}
        torch.cuda.synchronize(); dist.barrier()
        t4 = time.perf_counter()
        ens = _tdist.gather_ensembles(local, [n] * world)
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        posterior.update({"gather_ms": 1e3 * (t5 - t4), "gather_bytes_per_rank": int(sum(v.numel() * v.element_size() for v in local.values())),
                          "gathered_chains": int(ens["K"].shape[0])})
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t.item())
    total_props = float(n) * world * args.iters * args.steps
    value = total_props / (dev_ms_max * 1e-3)

    # ---- end-to-end arm through host buffers
    from tonga_b200.api import pinned_empty
    st = ch.state(want_ptS=False)
    K0 = pinned_empty(st["K"].shape, np.int32); K0[:] = st["K"]
    cells0 = pinned_empty(st["cells"].shape, np.float64); cells0[:] = st["cells"]
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
        tt = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        if i >= 2:
            e2e_t += float(tt.item())
        kept = int(np.minimum(hist["n_hist"], che.hist_cap).sum())
        d2h = kept * (32 * che.KC + 8 * ctx.R + 36) + sum(v.nbytes for v in fin.values() if v is not None)  # bytes the GPU wrote to host memory
    assert np.isfinite(chk)
    e2e_val = float(n) * world * args.iters * args.steps / e2e_t
    n_kept = int(hist["n_hist"].min())

    if rank == 0:
        peaks, peaks_src = load_peaks()
        fp64_peak, fp32_peak = ctx.peak_flops()
        flops, bytes_, c = algorithmic_work(counts, ctx.P, ctx.S, ctx.R)
        sec = dev_ms * 1e-3  # rank 0's own kernel time for rank 0's own chains
        ach_tf = flops / sec / 1e12
        ach_gb = bytes_ / sec / 1e9
        frac_fp = ach_tf / fp64_peak if fp64_peak > 0 else None
        frac_hbm = ach_gb / peaks["hbm_gbs"]
        props_rank = float(n) * args.iters * args.steps
        try:  # DRAM bytes per launch from the committed ncu --set full capture, scaled to this launch's proposal count
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj["dram_bytes_per_proposal"] * float(n) * args.iters
        except Exception:
            traffic = None
        roof = {"bound": "hbm", "achieved": ach_gb, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": frac_hbm, "traffic": traffic,
                "peak_source": f"MEASURED_PEAKS.json ({peaks_src})",
                "kernel": "tg_sampler_kernel", "launch_ms": dev_ms / args.steps,
                "algorithmic_bytes_per_proposal": bytes_ / props_rank, "algorithmic_flop_per_proposal": flops / props_rank,
                "fp64": {"achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": frac_fp,
                         "peak_source": "FMA-chain microbenchmark run in this process (tonga_peak_flops); FMA = 2 flop",
                         "note": "exact mode issues DMUL/DADD (1 flop per issue slot): its own ceiling is half the FMA peak"},
                "fp32_peak_tflops": fp32_peak,
                "note": "chain state (owners) is shared-memory resident by design: the HBM figure is algorithmic bytes / time"}
        acc_rate = (c[1, :4] / np.maximum(c[0, :4], 1)).round(4).tolist()
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "shipped 381-ray Tonga geometry (tonga_b200/datasets/tonga381.npz), slowness synthesised from ak135; device Philox proposals",
            "config": {"workload": WORKLOAD, "chains_per_gpu": n, "chains_total": n * world, "iters_per_step": args.iters,
                       "burn_in": int(p.burn_in), "keep_each": int(p.keep_each), "kept_models_per_chain_per_step": n_kept,
                       "prior": "uniform", "cells": [p.min_cells, p.max_cells], "parallelism": f"chain-sharded x{world}",
                       "l2": "256 MB buffer written between timed steps (L2 flush)"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_t / args.steps},
            "gpu_launches": 2 * args.steps,  # per step: tg_order_kernel (launch order by nCells) + tg_sampler_kernel
            "clocks": clocks,
            "roofline": roof,
            "evaluates_per_s": value * float(c[2].sum() / max(c[0].sum(), 1)),  # proposals that needed the forward model (a-priori rejects excluded; rank 0's mix)
            "acceptance": {"birth_death_change_move": acc_rate, "evaluated_fraction": float(c[2].sum() / max(c[0].sum(), 1)),
                           "mean_cells": float(st["K"].mean())},
            "verify": {"owner_mismatch": mm, "max_dphi": dphi, "max_dtstar": dts},
            "posterior": posterior,
        }
        if not args.no_cpu_baseline and world >= 1:
            rate, dt, chains = cpu_reference_arm(ds, p, args.cpu_iters, ncores)
            rate1, dt1, _ = cpu_reference_arm(ds, p, args.cpu_iters, 1, chains=1)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": ncores, "kind": "port",
                                   "sample": f"{chains} chains x {args.cpu_iters} iterations on {ncores} threads ({dt:.1f} s); single core: {rate1:.1f} proposals/s",
                                   "single_core_value": rate1}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
