import sys, numpy as np
sys.path.insert(0, "mcmc-in-tonga_b200")
from tonga_b200.api import Chains, Context
from tonga_b200.data import load_tonga381
from tonga_b200.structs import parameters
from scipy import stats
p = parameters(); ds = load_tonga381(p=p)
ctx = Context(ds, p); ch = Chains(ctx, 256, seed=1, hist_cap=0); ch.build_starting()
st0 = ch.state()
out = ch.run(400, record=True, trace=True)
recs = out["recs"]
# change proposals at iteration 0 relative to the start model: (zeta_new - zeta[idx]) / sig_zeta ~ N(0,1)
sig_zeta = p.zeta_scale * p.sig / 100
z = []
for c in range(256):
    r = recs[c, 0]
    if r["action"] == 3: z.append((r["zeta"] - st0["cells"][c, 3, r["idx"]]) / sig_zeta)
# moves: x displacement
mv = []
for c in range(256):
    r = recs[c, 0]
    if r["action"] == 4: mv.append((r["x"] - st0["cells"][c, 0, r["idx"]]) / ((p.sig / 100) * (ds.xVec.max() - ds.xVec.min())))
z = np.array(z); mv = np.array(mv)
print("change n=%d mean %.3f std %.3f ; move n=%d mean %.3f std %.3f" % (len(z), z.mean(), z.std(), len(mv), mv.mean(), mv.std()))
# all iterations: birth zeta - aux? aux not recorded; use u uniformity and action frequencies
print("actions", np.bincount(recs["action"].ravel())[1:] / recs.size)
