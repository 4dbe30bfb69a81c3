#!/usr/bin/env python
"""Generate the warm start models of the headline benchmark (needs a GPU): 1024 chains of the 381-ray Tonga inversion, seed
20260000, build_starting + 20 000 iterations -> tonga_b200/datasets/warm_start_1024.npz (K[1024], cells[1024, 4, Kmax]).
bench.py starts BOTH arms (B200 and CPU reference) from these models so that they time the same chain states.
Under gpurun: python tools/make_warm_start.py gpurun_out/warm_start_1024.npz ; then copy the file into the datasets directory."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mcmc-in-tonga_b200"))
from tonga_b200.api import Chains, Context  # noqa: E402
from tonga_b200.data import load_tonga381  # noqa: E402
from tonga_b200.structs import define_TDstructrure  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mcmc-in-tonga_b200", "tonga_b200", "datasets", "warm_start_1024.npz")
p = define_TDstructrure()
ds = load_tonga381(p=p)
ctx = Context(ds, p)
ch = Chains(ctx, 1024, chain_id0=0, seed=20260000, hist_cap=0)
ch.build_starting()
for _ in range(10):
    ch.run(2000)
assert ch.verify() == (0, 0.0, 0.0)
st = ch.state(want_ptS=False)
K = st["K"].astype(np.int32)
kmax = int(K.max())
cells = np.ascontiguousarray(st["cells"][:, :, :kmax])
for i, k in enumerate(K):
    cells[i, :, k:] = 0.0
np.savez_compressed(out, K=K, cells=cells, phi=st["phi"], seed=np.int64(20260000), iterations=np.int64(20000))
print("wrote", out, "K mean %.2f max %d" % (K.mean(), kmax), "phi mean %.1f" % st["phi"].mean(), os.path.getsize(out), "bytes")
