#!/usr/bin/env python
"""Full-evaluate (assignment + integration) throughput: point-nucleus pairs/s and fraction of the FP64 pipe.

BASELINE.json config 3 regime (assignment-bound, K up to 2000) on the Tonga geometry replicated over many models, inputs
resident in HBM (tonga_evaluate_batch_dev), CUDA events on torch's stream are NOT used: the library's own stream is timed
with its synchronize call around a batch of launches."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "mcmc-in-tonga_b200")]
import numpy as np
import torch
from tonga_b200.api import Context
from tonga_b200.data import load_tonga381
from tonga_b200.structs import define_TDstructrure

p = define_TDstructrure(); ds = load_tonga381(p=p)
ctx = Context(ds, p)
fp64_peak, fp32_peak = ctx.peak_flops()
rng = np.random.default_rng(0)
box = (ds.xVec.min(), ds.xVec.max(), ds.yVec.min(), ds.yVec.max(), ds.zVec.min(), ds.zVec.max())
out = []
for exact in (False, True):
  ctx.set_exact_only(exact)
  for K, n in ((20, 4096), (100, 2048), (500, 512), (2000, 128)):
      cells = np.stack([rng.uniform(box[0], box[1], (n, K)), rng.uniform(box[2], box[3], (n, K)), rng.uniform(box[4], box[5], (n, K)),
                        rng.uniform(0, 50, (n, K))], 1)
      Kd = torch.full((n,), K, dtype=torch.int32, device="cuda")
      cd = torch.from_numpy(cells).cuda()
      ptS = torch.zeros((n, ctx.R), dtype=torch.float64, device="cuda")
      phi = torch.zeros(n, dtype=torch.float64, device="cuda")
      torch.cuda.synchronize()
      for _ in range(2):
          ctx.evaluate_batch_dev(n, K, Kd.data_ptr(), cd.data_ptr(), None, ptS.data_ptr(), phi.data_ptr())
      ctx.synchronize()
      reps = 5
      t0 = time.perf_counter()
      for _ in range(reps):
          ctx.evaluate_batch_dev(n, K, Kd.data_ptr(), cd.data_ptr(), None, ptS.data_ptr(), phi.data_ptr())
      ctx.synchronize()
      dt = (time.perf_counter() - t0) / reps
      pairs = n * ctx.P * K
      flops = n * (8.0 * ctx.P * K + 5.0 * ctx.S + 4.0 * ctx.R)
      out.append(dict(mode='exact FP64' if exact else 'FP32 screen + FP64 recheck', K=K, models=n, ms=dt * 1e3, evaluates_per_s=n / dt, pairs_per_s=pairs / dt, tflops=flops / dt / 1e12,
                      frac_fp64_fma_peak=flops / dt / 1e12 / fp64_peak, frac_fp64_nofma_peak=flops / dt / 1e12 / (fp64_peak / 2), frac_fp32_fma_peak=flops / dt / 1e12 / fp32_peak))
print(json.dumps(dict(fp64_peak_tflops=fp64_peak, fp32_peak_tflops=fp32_peak, results=out), indent=1))
