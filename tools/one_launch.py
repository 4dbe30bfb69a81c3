import sys, os, numpy as np
sys.path.insert(0, "mcmc-in-tonga_b200")
from tonga_b200.api import Chains, Context
from tonga_b200.data import load_tonga381
from tonga_b200.structs import parameters
p = parameters(); ds = load_tonga381(p=p)
ctx = Context(ds, p); ch = Chains(ctx, 1024, seed=1, hist_cap=0); ch.build_starting(); ch.run(3000)
ch.run(1000); print(os.environ.get("TONGA_B200_LIB"), ch.last_kernel_ms())
