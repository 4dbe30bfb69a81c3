"""Headline workload (warm start models, reset + set_models + run(1000) per step, as bench.py) with and without the cost feedback in the
launch order: python tools/order_feedback_ab.py"""
import os, sys, numpy as np
sys.path.insert(0, "mcmc-in-tonga_b200")
from tonga_b200.api import Chains, Context
from tonga_b200.data import load_tonga381, load_warm_start
from tonga_b200.structs import parameters
p = parameters(); p.n_iter, p.burn_in = 1000.0, 500.0
ds = load_tonga381(p=p); warm = load_warm_start()
ctx = Context(ds, p)
K0, c0 = warm["K"][:1024].astype(np.int32), warm["cells"][:1024]
res = {}
for rep in range(2):
    for fb in ("0", "1"):
        os.environ["TONGA_ORDER_FEEDBACK"] = fb
        ch = Chains(ctx, 1024, seed=20260000)
        ms = []
        for step in range(6):
            ch.reset(); ch.set_models(K0, c0); ch.run(1000); ms.append(ch.last_kernel_ms())
        st = ch.state(want_ptS=False)
        res[fb] = (st["K"].copy(), st["phi"].copy())
        print("feedback", fb, ["%.2f" % m for m in ms], "-> %.2f M/s (median of the last 4)" % (1024 * 1000 / np.median(ms[2:]) / 1e3), "verify", ch.verify())
        ch.close()
print("final states identical with / without feedback:", bool((res["0"][0] == res["1"][0]).all() and (res["0"][1] == res["1"][1]).all()))
