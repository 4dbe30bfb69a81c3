"""Ray-sharded STREAMED sampler across the GPUs of one node (one process per GPU, peer memory over CUDA IPC):
parity against the unsharded run, then timing.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py [R] [chains] [K] [iters]

Every rank first runs the batch UNSHARDED on its own GPU (the reference), then the ranks connect and run it ray-sharded from
the same start models.  Checked on every rank: proposals, decisions, phi traces, nCells, final models, t*, phi bit-identical;
verify() of the own points.  Prints one JSON line on rank 0.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mcmc-in-tonga_b200"))
import numpy as np
import torch
import torch.distributed as dist

from tonga_b200.api import Chains, Context, pack_models
from tonga_b200.data import synthetic_rays
from tonga_b200.dist import connect_ray_shards, disconnect_ray_shards
from tonga_b200.structs import parameters

R = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
K0 = int(sys.argv[3]) if len(sys.argv) > 3 else 200
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 40

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p = parameters()
p.max_cells, p.min_cells = 2000, 5
p.n_iter, p.burn_in, p.keep_each = float(iters), float(iters // 2), 5.0
ds = synthetic_rays(R, seed=3, p=p, n_true=0)
rng = np.random.default_rng(7)
box = (ds.xVec.min(), ds.xVec.max(), ds.yVec.min(), ds.yVec.max(), ds.zVec.min(), ds.zVec.max())
ctx = Context(ds, p, device=local, device_ingest=True)
models = [[rng.uniform(box[0], box[1], K0), rng.uniform(box[2], box[3], K0), rng.uniform(box[4], box[5], K0), rng.uniform(0, 50, K0)] for _ in range(n)]


def batch():
    ch = Chains(ctx, n, chain_id0=0, seed=11, sampler="streamed")
    Kp, cp = pack_models(models, Kcap=ch.KC)
    ch.set_models(Kp, cp)
    return ch


# ---- unsharded reference on this GPU
ref = batch()
o_ref = ref.run(iters, record=True, trace=True)
ms_ref = ref.last_kernel_ms()
s_ref, h_ref = ref.state(), ref.history()
assert ref.verify() == (0, 0.0, 0.0)
ref.close()

# ---- ray-sharded
ch = batch()
connect_ray_shards(ch, local)
info = ch.shard_info()
dist.barrier()
o = ch.run(iters, record=True, trace=True)
ms1 = ch.last_kernel_ms()
s, h = ch.state(), ch.history()
ok = (o["recs"].tobytes() == o_ref["recs"].tobytes() and np.array_equal(o["accept"], o_ref["accept"]) and o["phi"].tobytes() == o_ref["phi"].tobytes()
      and np.array_equal(o["K"], o_ref["K"]) and s["phi"].tobytes() == s_ref["phi"].tobytes() and s["ptS"].tobytes() == s_ref["ptS"].tobytes()
      and s["cells"].tobytes() == s_ref["cells"].tobytes() and h["ptS"].tobytes() == h_ref["ptS"].tobytes() and np.array_equal(h["n_hist"], h_ref["n_hist"]))
ver = ch.verify()
# ---- timing: a second, longer leg (same lock step on every rank)
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
ch.run(iters)
ms2 = ch.last_kernel_ms()
wall = time.perf_counter() - t0
t = torch.tensor([ms2, float(ok), float(ver == (0, 0.0, 0.0))], dtype=torch.float64, device="cuda")
tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"world": world, "rays": R, "points": int(ctx.P), "chains": n, "K_start": K0, "iterations": iters,
                      "parity_all_ranks": bool(tmin[1].item() == 1.0), "verify_all_ranks": bool(tmin[2].item() == 1.0),
                      "unsharded_ms_per_iteration": ms_ref / iters, "sharded_ms_per_iteration": float(tmax[0].item()) / iters,
                      "speedup": ms_ref / float(tmax[0].item()), "proposals_per_s_sharded": n * iters / float(tmax[0].item()) * 1e3,
                      "own_points_rank0": [info["point0"], info["point1"]], "wall_s_rank0": wall}))
disconnect_ray_shards(ch)
ch.close()
ctx.close()
dist.barrier()
dist.destroy_process_group()
if not (tmin[1].item() == 1.0 and tmin[2].item() == 1.0):
    sys.exit(1)
