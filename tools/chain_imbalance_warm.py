"""Per-chain busy cycles and per-phase cycles of the HEADLINE workload (warm start models, as bench.py): python tools/chain_imbalance_warm.py"""
import sys, numpy as np
sys.path.insert(0, "mcmc-in-tonga_b200")
from tonga_b200.api import Chains, Context
from tonga_b200.data import load_tonga381, load_warm_start
from tonga_b200.structs import parameters
p = parameters(); p.n_iter, p.burn_in = 1000.0, 500.0
ds = load_tonga381(p=p); warm = load_warm_start()
ctx = Context(ds, p); ch = Chains(ctx, 1024, seed=20260000)
ch.set_models(warm["K"][:1024].astype(np.int32), warm["cells"][:1024]); ch.run(1000)
ch.reset(); ch.set_models(warm["K"][:1024].astype(np.int32), warm["cells"][:1024])
ch.profile(True); ch.run(1000); cyc = ch.profile(False, read=True); ms = ch.last_kernel_ms()
tot = cyc[:, :9].sum(1)
K = ch.state()["K"]
print("instrumented launch %.1f ms; per-chain total cycles: mean %.3g  max %.3g  min %.3g  max/mean %.3f  p95/mean %.3f  p50/mean %.3f" % (ms, tot.mean(), tot.max(), tot.min(), tot.max() / tot.mean(), np.percentile(tot, 95) / tot.mean(), np.percentile(tot, 50) / tot.mean()))
print("corr(total, K) = %.3f ; K mean %.1f min %d max %d" % (np.corrcoef(tot, K)[0, 1], K.mean(), K.min(), K.max()))
names = ["A", "B2", "C", "D+E", "F4", "G", "F1", "F2", "B1"]
o = np.argsort(tot)
print("all chains", {n: int(v) for n, v in zip(names, cyc[:, :9].mean(0) / 1000)}, "total/iter %d" % (tot.mean() / 1000))
for lab, sel in (("fastest 10%", o[:102]), ("middle 10%", o[460:562]), ("slowest 10%", o[-102:])):
    print(lab, "K mean %.1f" % K[sel].mean(), {n: int(v) for n, v in zip(names, cyc[sel, :9].mean(0) / 1000)}, "total/iter %d" % (tot[sel].mean() / 1000))
print("launch length in cycles / slowest chain:", ms * 1e-3 * 1.965e9 / tot.max())
b, sm = cyc[:, 14].astype(int), cyc[:, 15].astype(int)
o = np.argsort(b)
print("SM of CTA 0..15:", sm[o][:16], " CTA 148..163:", sm[o][148:164], " CTA 296..311:", sm[o][296:312])
print("chains per SM: min %d max %d; SMs used %d" % (np.bincount(sm).min(), np.bincount(sm).max(), len(np.unique(sm))))
same = np.mean([sm[o][i] == sm[o][i + 148] for i in range(1024 - 148)]); print("fraction of CTAs i, i+148 on the same SM: %.3f" % same)
# per-SM load and finish
tot_sm = np.bincount(sm, weights=tot); mx_sm = np.array([tot[sm == s].max() if (sm == s).any() else 0 for s in range(sm.max() + 1)])
print("per-SM sum of chain cycles: max/mean %.3f ; per-SM slowest chain: max/mean %.3f" % (tot_sm.max() / tot_sm[tot_sm > 0].mean(), mx_sm.max() / mx_sm[mx_sm > 0].mean()))
