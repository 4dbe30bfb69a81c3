#!/usr/bin/env python
"""BASELINE.json config 3 at full size: synthetic 3-D model, 100k rays x ~200 points (P ~ 2e7), K in {100, 500, 2000}.

Full evaluate (assignment + integration + misfit) through tonga_evaluate_batch_dev with the inputs resident in HBM; parity of
owners / t* / phi against the CPU oracle on the whole ray set at K = 100 (the oracle needs ~2e9 pair evaluations: seconds).
Usage: python tools/bench_config3.py [R]  (default 100000)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mcmc-in-tonga_b200"), os.path.join(ROOT, "oracle")]
import numpy as np
import torch
from tonga_b200.api import Context
from tonga_b200.data import synthetic_rays
from tonga_b200.structs import parameters

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
p = parameters()
t0 = time.time(); ds = synthetic_rays(R, seed=3, p=p); t_gen = time.time() - t0
t0 = time.time(); ctx = Context(ds, p); t_flat = time.time() - t0
t0 = time.time(); ctx_dev = Context(ds, p, device_ingest=True); t_ingest = time.time() - t0
assert (ctx_dev.P, ctx_dev.S) == (ctx.P, ctx.S)
ctx_dev.close()
box = (ds.xVec.min(), ds.xVec.max(), ds.yVec.min(), ds.yVec.max(), ds.zVec.min(), ds.zVec.max())
fp64_peak, fp32_peak = ctx.peak_flops()
rng = np.random.default_rng(0)
out = dict(R=R, P=int(ctx.P), S=int(ctx.S), gen_s=t_gen, flatten_upload_s=t_flat, device_ingest_s=t_ingest, fp64_peak_tflops=fp64_peak, fp32_peak_tflops=fp32_peak, results=[])
for K, n in ((100, 4), (500, 2), (2000, 1)):
    cells = np.stack([rng.uniform(box[0], box[1], (n, K)), rng.uniform(box[2], box[3], (n, K)), rng.uniform(box[4], box[5], (n, K)),
                      rng.uniform(0, 50, (n, K))], 1)
    Kd = torch.full((n,), K, dtype=torch.int32, device="cuda"); cd = torch.from_numpy(cells).cuda()
    ptS = torch.zeros((n, ctx.R), dtype=torch.float64, device="cuda"); phi = torch.zeros(n, dtype=torch.float64, device="cuda")
    for mode in (False, True):
        ctx.set_exact_only(mode)
        ctx.evaluate_batch_dev(n, K, Kd.data_ptr(), cd.data_ptr(), None, ptS.data_ptr(), phi.data_ptr()); ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.evaluate_batch_dev(n, K, Kd.data_ptr(), cd.data_ptr(), None, ptS.data_ptr(), phi.data_ptr())
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / 3
        out["results"].append(dict(K=K, models=n, mode="exact FP64" if mode else "FP32 screen + FP64 recheck", ms_per_evaluate=1e3 * dt / n,
                                   pairs_per_s=n * ctx.P * K / dt, algorithmic_tflops=n * (8.0 * ctx.P * K + 5.0 * ctx.S + 4.0 * ctx.R) / dt / 1e12))
    ctx.set_exact_only(False)
# parity on the whole set at K = 100 (one model)
import oracle as O
op = O.make_params(box); od = O.Data(ds.rayX, ds.rayY, ds.rayZ, ds.rayL, ds.rayU, ds.tS, ds.allSig)
K = 100
mdl = [rng.uniform(box[0], box[1], K), rng.uniform(box[2], box[3], K), rng.uniform(box[4], box[5], K), rng.uniform(0, 50, K)]
t0 = time.time(); ref = O.evaluate(op, od, *mdl, want_owners=True); t_orc = time.time() - t0
gb = ctx.evaluate_batch(np.array([K], np.int32), np.stack(mdl)[None], want_owners=True)
off = ctx.ray_offsets()
m = ds.rayX.shape[0]
mask = np.arange(m)[:, None] < (off[1:] - off[:-1])[None, :]
flat = ref["owners"].T[mask.T]
out["parity_K100"] = dict(points=int(ctx.P), owner_mismatches=int((gb["owners"][0] != flat).sum()),
                          max_rel_tstar=float(np.max(np.abs(gb["ptS"][0] - ref["ptS"]) / np.maximum(np.abs(ref["ptS"]), 1e-300))),
                          rel_phi=float(abs(gb["phi"][0] - ref["phi"]) / abs(ref["phi"])), oracle_seconds_1core=t_orc)

# ---- the proposal loop at full size: wide sampler (a full forward model per proposal), max_cells = 2000
import copy
from tonga_b200.api import Chains, pack_models
p3 = copy.copy(p); p3.max_cells, p3.min_cells = 2000, 5
p3.n_iter, p3.burn_in, p3.keep_each = 1e9, 1e9, 1.0   # no thinning in the timed loop
ctx3 = Context(ds, p3)
samp = []
for kind, n, K0, iters in (("wide", 8, 100, 40), ("wide", 8, 1000, 20), ("wide", 32, 1000, 10),
                           ("streamed", 8, 100, 200), ("streamed", 8, 1000, 200), ("streamed", 32, 1000, 100), ("streamed", 128, 1000, 50)):
    ch = Chains(ctx3, n, seed=11, hist_cap=0, sampler=kind)
    mdl0 = [[rng.uniform(box[0], box[1], K0), rng.uniform(box[2], box[3], K0), rng.uniform(box[4], box[5], K0), rng.uniform(0, 50, K0)] for _ in range(n)]
    Kp, cp = pack_models(mdl0, Kcap=ch.KC)
    ch.set_models(Kp, cp)
    ch.run(3)
    ch.run(iters)
    ms = ch.last_kernel_ms()
    it, counts = ch.stats()
    mm, dphi, dts = ch.verify() if n <= 32 else (0, 0.0, 0.0)
    samp.append(dict(sampler=kind, chains=n, K_start=K0, iterations=iters, ms_per_iteration=ms / iters, proposals_per_s=n * iters / ms * 1e3,
                     evaluated_fraction=float(counts[:, 2].sum() / max(counts[:, 0].sum(), 1)), accept_fraction=float(counts[:, 1].sum() / max(counts[:, 0].sum(), 1)),
                     verify=dict(owner_mismatches=mm, max_dphi=dphi, max_dtstar=dts)))
    ch.close()
out["samplers"] = samp
# replay of the oracle's chain at full size (1 chain, K ~ 100, a few proposals: the oracle needs ~2e9 pair evaluations each)
op3 = O.make_params(box, max_cells=2000, min_cells=5, n_iter=1e9, burn_in=1e9, keep_each=1.0)
mb = O.ModelBuf(op3.max_cells + 1, od.R).set(*mdl)
assert O.lib().orc_evaluate(op3, od.c, mb.c, None, None) == 1
t0 = time.time(); run = O.chain_run(op3, od, mb, 6, g=O.rng(5)); t_orc = time.time() - t0
ch = Chains(ctx3, 1, hist_cap=0)
assert ch.sampler == "streamed"
Kp, cp = pack_models([mdl], Kcap=ch.KC); ch.set_models(Kp, cp)
got = ch.run(6, recs=run.recs[None], trace=True)
out["streamed_sampler_replay"] = dict(proposals=6, accept_identical=bool(np.array_equal(got["accept"][0], run.accept)), K_identical=bool(np.array_equal(got["K"][0], run.K)),
                                  max_rel_phi=float(np.max(np.abs(got["phi"][0] - run.phi) / np.abs(run.phi))), oracle_seconds_1core=t_orc)
print(json.dumps(out, indent=1))
