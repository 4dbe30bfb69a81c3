#!/usr/bin/env python
"""Long-run consistency soak: many proposals, then the maintained state must still equal a fresh forward model bit for bit
(tonga_chains_verify) and the three samplers must still agree.  python tools/soak.py [resident_iters] [streamed_iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mcmc-in-tonga_b200")]
import numpy as np
from tonga_b200.api import Chains, Context
from tonga_b200.data import load_tonga381
from tonga_b200.structs import parameters

n_res = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
n_str = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
p = parameters(); ds = load_tonga381(p=p)
ctx = Context(ds, p)
ch = Chains(ctx, 1024, seed=2026, hist_cap=0); ch.build_starting()
t0 = time.time(); done = 0
while done < n_res:
    step = min(100000, n_res - done)
    ch.run(step); done += step
    v = ch.verify()
    assert v == (0, 0.0, 0.0), (done, v)
it, counts = ch.stats()
st = ch.state()
print(f"resident: 1024 chains x {done} iterations = {1024 * done:.3g} proposals in {time.time() - t0:.1f} s; verify clean every 1e5 iterations; "
      f"K in [{st['K'].min()}, {st['K'].max()}], acceptance {np.round(counts[:, 1].sum(0) / np.maximum(counts[:, 0].sum(0), 1), 3)[:4].tolist()}")
ch.close()
ref = None
for kind in ("resident", "streamed", "wide"):
    c = Chains(ctx, 64, seed=7, chain_id0=100, hist_cap=0, sampler=kind); c.build_starting()
    t0 = time.time(); c.run(n_str)
    assert c.verify() == (0, 0.0, 0.0)
    s = c.state()
    sig = (s["K"].tobytes(), s["phi"].tobytes(), s["ptS"].tobytes())
    if ref is None: ref = sig
    assert sig == ref, kind
    print(f"{kind}: 64 chains x {n_str} iterations in {time.time() - t0:.1f} s: verify clean, final state bit-identical to the resident sampler")
    c.close()
