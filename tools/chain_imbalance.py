import sys, numpy as np
sys.path.insert(0, "mcmc-in-tonga_b200")
from tonga_b200.api import Chains, Context
from tonga_b200.data import load_tonga381
from tonga_b200.structs import parameters
p = parameters(); ds = load_tonga381(p=p)
ctx = Context(ds, p); ch = Chains(ctx, 1024, seed=1, hist_cap=0); ch.build_starting(); ch.run(3000)
ch.profile(True); ch.run(1000); cyc = ch.profile(False, read=True)
tot = cyc[:, :9].sum(1)
K = ch.state()["K"]
print("per-chain total cycles: mean %.3g  max %.3g  min %.3g  max/mean %.3f  p95/mean %.3f" % (tot.mean(), tot.max(), tot.min(), tot.max() / tot.mean(), np.percentile(tot, 95) / tot.mean()))
print("corr(total, K) = %.3f ; K mean %.1f max %d" % (np.corrcoef(tot, K)[0, 1], K.mean(), K.max()))
order = np.argsort(tot)[::-1][:8]; print("slowest chains K:", K[order], (tot[order] / tot.mean()).round(3))
names = ["A", "B2", "C", "D+E", "F4", "G", "F1", "F2", "B1"]
o = np.argsort(tot)
for lab, sel in (("fastest 10%", o[:102]), ("middle 10%", o[460:562]), ("slowest 10%", o[-102:])):
    print(lab, "K mean %.1f" % K[sel].mean(), {n: int(v) for n, v in zip(names, cyc[sel, :9].mean(0) / 1000)}, "total/iter %d" % (tot[sel].mean() / 1000))
it, counts = ch.stats()
