"""torchrun check (N GPUs): MLE restarts sharded over ranks give the single-process result, every
rank ends with the bit-identical model, and sharded top-k over a population equals the
single-GPU top-k.  Usage: torchrun --nproc-per-node N tools/dist_fit_check.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import mipego_b200 as mb
from mipego_b200 import dist as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, d, R = 600, 5, 6
rng = np.random.default_rng(0)
X = rng.uniform(-5, 5, (n, d))
y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
y = (y - y.mean()) / y.std()
kw = dict(corr="squared_exponential", theta0=[0.05] * d, thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=1e-6,
          random_start=R, device=local)
np.random.seed(1)
g1 = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), **kw)
D.sharded_fit(g1, X, y, device=dev)
np.random.seed(1)
g2 = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), **kw).fit(X, y)   # all restarts locally
same = np.array_equal(g1.theta_, g2.theta_) and np.array_equal(g1.C, g2.C) and np.array_equal(g1.gamma, g2.gamma)
# every rank holds the same model
t = torch.from_numpy(np.r_[g1.theta_, g1.gamma.ravel()]).to(dev)
ts = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(ts, t)
ident = all(torch.equal(ts[0], u) for u in ts)
Xs = np.random.default_rng(5).uniform(-5, 5, (100001, d))
ei = mb.EI(model=g1, minimize=True)
v, i = D.sharded_topk(ei, Xs, k=8, device=dev)
v1, i1 = ei.topk(Xs, k=8)
topk_ok = np.array_equal(i, i1) and np.array_equal(v, v1)
if rank == 0:
    print("sharded_fit == local fit:", same, "| identical across ranks:", ident, "| sharded top-k == single:", topk_ok,
          "| llf %.6f evals %d" % (float(np.ravel(g1.log_likelihood_)[0]), g1.eval_count))
assert same and ident and topk_ok
dist.barrier()
dist.destroy_process_group()
