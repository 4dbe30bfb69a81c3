#!/usr/bin/env python
"""Streamed sampler at BASELINE.json config 3 size (developer aid; ncu target): python tools/prof_stream.py [R] [chains] [K0] [iters]"""
import copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mcmc-in-tonga_b200")]
import numpy as np
from tonga_b200.api import Chains, Context, pack_models
from tonga_b200.data import synthetic_rays
from tonga_b200.structs import parameters

R = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
K0 = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
p = parameters(); p.max_cells, p.min_cells = 2000, 5
p.n_iter, p.burn_in, p.keep_each = 1e9, 1e9, 1.0
ds = synthetic_rays(R, seed=3, p=p, n_true=0)  # tS = noise only: the data do not matter for the timing
ctx = Context(ds, p)
box = (ds.xVec.min(), ds.xVec.max(), ds.yVec.min(), ds.yVec.max(), ds.zVec.min(), ds.zVec.max())
rng = np.random.default_rng(0)
ch = Chains(ctx, n, seed=11, hist_cap=0, sampler=os.environ.get("TONGA_SAMPLER", "streamed"))
mdl0 = [[rng.uniform(box[0], box[1], K0), rng.uniform(box[2], box[3], K0), rng.uniform(box[4], box[5], K0), rng.uniform(0, 50, K0)] for _ in range(n)]
Kp, cp = pack_models(mdl0, Kcap=ch.KC)
ch.set_models(Kp, cp)
ch.run(3)
ch.run(iters)
ms = ch.last_kernel_ms()
it, counts = ch.stats()
print(f"{ch.sampler}: P={ctx.P} chains={n} K0={K0}: {ms / iters:.3f} ms/iteration, {n * iters / ms * 1e3:.0f} proposals/s, "
      f"{ms / iters / n * 1e3:.1f} us per chain-proposal; proposed/accepted/evaluated by action: {counts.sum(0).tolist()}")

if os.environ.get("PER_ACTION"):
    out = ch.run(iters, record=True)
    base = out["recs"]
    for A in (1, 2, 3, 4):
        recs = base.copy()
        off = recs["action"] != A
        recs["action"][off] = 1; recs["zeta"][off] = -1.0  # a-priori rejected birth: no evaluation
        it0, c0 = ch.stats()
        ch.run(iters, recs=recs)
        ms = ch.last_kernel_ms()
        it1, c1 = ch.stats()
        ev = int((c1 - c0)[:, 2, A - 1].sum()); acc = int((c1 - c0)[:, 1, A - 1].sum())
        print(f"  action {A}: {ev} evaluated, {acc} accepted, {ms:.2f} ms -> {ms / max(ev, 1) * 1e3:.1f} us per evaluated chain-proposal")
