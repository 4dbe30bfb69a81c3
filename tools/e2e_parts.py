import sys, time, numpy as np
sys.path.insert(0, "mcmc-in-tonga_b200")
from tonga_b200.api import Chains, Context, pinned_empty
from tonga_b200.data import load_tonga381, load_warm_start
from tonga_b200.structs import parameters
p = parameters(); p.n_iter, p.burn_in = 1000.0, 500.0
ds = load_tonga381(p=p); warm = load_warm_start()
ctx = Context(ds, p); n = 1024
K0 = pinned_empty((n,), np.int32); K0[:] = warm["K"][:n]
cells0 = pinned_empty((n, 4, warm["cells"].shape[2]), np.float64); cells0[:] = warm["cells"][:n]
for host_hist in (True, False):
    che = Chains(ctx, n, seed=20260000, host_history=host_hist)
    _, state_buf = che.alloc_buffers(pinned=True)
    T = {k: [] for k in ("reset", "set_models", "run", "kernel_ms", "history", "state", "total")}
    for i in range(6):
        ctx.synchronize(); t0 = time.perf_counter()
        che.reset(); ctx.synchronize(); t1 = time.perf_counter()
        che.set_models(K0, cells0); ctx.synchronize(); t2 = time.perf_counter()
        che.run(1000); t3 = time.perf_counter()
        hist = che.history(); t4 = time.perf_counter()
        fin = che.state(out=state_buf); t5 = time.perf_counter()
        for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, che.last_kernel_ms() * 1e-3, t4 - t3, t5 - t4, t5 - t0)): T[k].append(v * 1e3)
    print("host_history", host_hist, {k: round(float(np.median(v[2:])), 3) for k, v in T.items()})
    che.close()
