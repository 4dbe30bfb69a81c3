"""Launch order from feedback against the K sort for CONTINUING runs (needs a library built with the feedback order: the experiment of
profiles/README.md round 2; kept as the record of how it was measured): python tools/order_continuing_ab.py"""
import os, sys, numpy as np
sys.path.insert(0, "mcmc-in-tonga_b200")
from tonga_b200.api import Chains, Context
from tonga_b200.data import load_tonga381, load_warm_start
from tonga_b200.structs import parameters
p = parameters(); ds = load_tonga381(p=p); warm = load_warm_start()
ctx = Context(ds, p)
K0, c0 = warm["K"][:1024].astype(np.int32), warm["cells"][:1024]
for start in ("warm", "fresh"):
    for nit in (1000, 250):
        for fb in ("0", "1"):
            os.environ["TONGA_ORDER_FEEDBACK"] = fb
            ch = Chains(ctx, 1024, seed=7, hist_cap=0)
            if start == "warm": ch.set_models(K0, c0)
            else: ch.build_starting()
            ms = []
            for step in range(8):
                ch.run(nit); ms.append(ch.last_kernel_ms())
            print(start, "continuous run(%d) x 8, feedback" % nit, fb, ["%.2f" % m for m in ms], "sum of the last 6: %.2f ms" % sum(ms[2:]))
            ch.close()
