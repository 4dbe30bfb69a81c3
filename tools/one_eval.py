"""One batched full evaluate at K = 500 on the Tonga set (ncu target for tg_eval_kernel): python tools/one_eval.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mcmc-in-tonga_b200"))
import torch
from tonga_b200.api import Context
from tonga_b200.data import load_tonga381
from tonga_b200.structs import define_TDstructrure
p = define_TDstructrure(); ds = load_tonga381(p=p)
ctx = Context(ds, p)
rng = np.random.default_rng(0)
box = (ds.xVec.min(), ds.xVec.max(), ds.yVec.min(), ds.yVec.max(), ds.zVec.min(), ds.zVec.max())
K, n = 500, 512
cells = np.stack([rng.uniform(box[0], box[1], (n, K)), rng.uniform(box[2], box[3], (n, K)), rng.uniform(box[4], box[5], (n, K)), rng.uniform(0, 50, (n, K))], 1)
Kd = torch.full((n,), K, dtype=torch.int32, device="cuda"); cd = torch.from_numpy(cells).cuda()
ptS = torch.zeros((n, ctx.R), dtype=torch.float64, device="cuda"); phi = torch.zeros(n, dtype=torch.float64, device="cuda")
for _ in range(3):
    ctx.evaluate_batch_dev(n, K, Kd.data_ptr(), cd.data_ptr(), None, ptS.data_ptr(), phi.data_ptr())
ctx.synchronize()
print("ok", float(phi.sum()))
