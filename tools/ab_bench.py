"""A/B of two builds of libtonga_b200.so on the headline workload (developer aid): python tools/ab_bench.py old.so new.so
Each library is loaded in its own subprocess; prints proposals/s of 1024 chains x 1000 iterations (kernel time, median of 5)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CHILD = r'''
import sys, os, numpy as np
sys.path.insert(0, os.path.join(%(root)r, "mcmc-in-tonga_b200"))
from tonga_b200 import _lib
_lib.LIB_PATH = %(lib)r
import ctypes as C
probe = C.CDLL(%(lib)r)
for name in list(_lib.SYMBOLS):
    if not hasattr(probe, name): _lib.SYMBOLS.pop(name)
from tonga_b200.api import Chains, Context
from tonga_b200.data import load_tonga381
from tonga_b200.structs import parameters
p = parameters(); ds = load_tonga381(p=p)
ctx = Context(ds, p); ch = Chains(ctx, 1024, seed=1, hist_cap=0); ch.build_starting(); ch.run(3000)
ms = []
for _ in range(5):
    ch.run(1000); ms.append(ch.last_kernel_ms())
print("%(tag)s", "%%.2f M/s" %% (1024 * 1000 / np.median(ms) / 1e3), ["%%.1f" %% m for m in ms])
if os.environ.get("AB_PROFILE"):
    ch.profile(True); pm = []
    for _ in range(3):
        ch.run(1000); pm.append(ch.last_kernel_ms())
    cyc = ch.profile(False, read=True) / 3
    print("   instrumented build: %%.2f M/s" %% (1024 * 1000 / np.median(pm) / 1e3), ["%%.1f" %% m for m in pm])
    print("   cycles/iter A,B2,C,D+E,F4,G,F1,F2,B1:", (cyc[:, :9].mean(0) / 1000).round(0), "sum", round(float(cyc[:, :9].mean(0).sum()) / 1000))
'''
root = os.path.dirname(HERE)
for rep in range(int(os.environ.get('REPS', '2'))):
    for lib in sys.argv[1:]:
        path = os.path.abspath(lib)
        subprocess.run([sys.executable, "-c", CHILD % {"root": root, "lib": path, "tag": os.path.basename(lib)}], check=True)
