"""Multi-GPU plumbing for the GP hot path: one process per GPU, `torch.distributed` (NCCL over
NVLink on the GPU box, gloo in the CPU tests).

The path shards naturally: candidates are independent units, MLE restarts are
independent units.  The only exchanges are
  * one broadcast of the fitted state per fit, so that every rank holds the bit-identical model
    (results are then independent of the number of shards), and
  * an all-gather of each rank's k best (value, global index) pairs, merged identically on every
    rank (value descending, ties -> smaller global index, the order the single-GPU kernel uses).
No collective sits inside the compute kernels; the data path itself needs none.
"""
import numpy as np


def shard_range(m, rank, world):
    """Contiguous, balanced partition of m units: rank r gets [lo, hi)."""
    base, rem = divmod(int(m), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch
    import torch.distributed as dist
    return torch, dist


def broadcast_fitted_state(gp, src=0, device=None):
    """Broadcast (theta, L, gamma, sigma2, beta) of the fitted model on `src` and adopt it on the
    other ranks through `set_fitted_state` (mgp_set_state).  Every rank must hold the same
    training data X, y (they are part of the optimiser state in the BO loop)."""
    torch, dist = _dist()
    rank = dist.get_rank()
    n, d = gp.X.shape
    dev = device if device is not None else "cpu"
    if rank == src:
        pack = [np.asarray(gp._state("theta"), dtype=np.float64).ravel(),
                np.ascontiguousarray(gp.C, dtype=np.float64).ravel(),
                np.asarray(gp.gamma, dtype=np.float64).ravel(),
                np.array([float(np.ravel(gp.sigma2)[0]), float(np.ravel(gp.mean.beta)[0])])]
        flat = np.concatenate(pack)
    else:
        flat = np.empty(d + n * n + n + 2)
    t = torch.from_numpy(flat).to(dev)
    dist.broadcast(t, src=src)
    if rank != src:
        flat = t.cpu().numpy()
        theta, L = flat[:d], flat[d:d + n * n].reshape(n, n)
        gamma, sc = flat[d + n * n:d + n * n + n], flat[-2:]
        gp.set_fitted_state(gp.X, gp.y.ravel(), theta, L, gamma, float(sc[0]), float(sc[1]))
    return gp


def merge_topk(vals, idx, k):
    """Merge per-rank top-k lists.  vals, idx: (world, n_sets, k_r).  Returns (n_sets, k) values
    (descending) and global indices; ties -> smaller index; NaN sorts last; idx < 0 = empty slot."""
    vals = np.asarray(vals, dtype=np.float64)
    idx = np.asarray(idx, dtype=np.int64)
    W, S, K = vals.shape
    v = np.transpose(vals, (1, 0, 2)).reshape(S, W * K)
    i = np.transpose(idx, (1, 0, 2)).reshape(S, W * K)
    v = np.where(np.isnan(v) | (i < 0), -np.inf, v)
    out_v = np.empty((S, k))
    out_i = np.empty((S, k), dtype=np.int64)
    big = np.iinfo(np.int64).max
    for s in range(S):
        order = np.lexsort((np.where(i[s] < 0, big, i[s]), -v[s]))[:k]
        out_v[s], out_i[s] = v[s][order], i[s][order]
    return out_v, out_i


def allgather_topk(local_vals, local_idx, k, offset, device=None):
    """All-gather each rank's (n_sets, k) best candidates (local indices + `offset` = global) and
    merge.  Every rank returns the same (values, global indices)."""
    torch, dist = _dist()
    world = dist.get_world_size()
    dev = device if device is not None else "cpu"
    lv = np.asarray(local_vals, dtype=np.float64)
    li = np.asarray(local_idx, dtype=np.int64)
    mine = torch.from_numpy(np.stack([lv, np.where(li >= 0, li + offset, -1).astype(np.float64)])).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    allv = np.stack([o[0].cpu().numpy() for o in out])
    alli = np.stack([o[1].cpu().numpy() for o in out]).astype(np.int64)
    return merge_topk(allv, alli, k)


def sharded_topk(criterion, X, k=1, params=None, device=None):
    """Population arg-max across ranks: every rank passes the SAME (m, d) candidate array (or
    generates it from the same seed); rank r evaluates rows shard_range(m, r, world) on its GPU."""
    torch, dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(len(X), rank, world)
    v, i = criterion.topk(np.asarray(X)[lo:hi], k=k, params=params)
    if v.shape[1] < k:  # short shard: pad with empty slots
        pad = k - v.shape[1]
        v = np.c_[v, np.full((v.shape[0], pad), -np.inf)]
        i = np.c_[i, np.full((i.shape[0], pad), -1, dtype=np.int64)]
    return allgather_topk(v, i, k, lo, device=device)


def sharded_nll(gp, params, eval_grad=True, device=None):
    """MLE restarts across ranks: rank r evaluates rows shard_range(B, r, world) of `params`;
    the (llf, grad, status) rows are all-gathered so every rank sees all B results."""
    torch, dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    P = np.atleast_2d(np.asarray(params, dtype=np.float64))
    B, n_par = P.shape
    lo, hi = shard_range(B, rank, world)
    width = n_par + 2
    rows = np.zeros((max(1, -(-B // world)), width))
    if hi > lo:
        llf, grad, status = gp.log_likelihood_batch(P[lo:hi], eval_grad=eval_grad)
        rows[:hi - lo, 0] = llf
        rows[:hi - lo, 1] = status
        if eval_grad:
            rows[:hi - lo, 2:] = grad
    dev = device if device is not None else "cpu"
    mine = torch.from_numpy(rows).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    llf_all, st_all, g_all = np.empty(B), np.empty(B, dtype=np.int32), np.zeros((B, n_par))
    for r in range(world):
        a, b = shard_range(B, r, world)
        blk = out[r].cpu().numpy()
        llf_all[a:b], st_all[a:b], g_all[a:b] = blk[:b - a, 0], blk[:b - a, 1].astype(np.int32), blk[:b - a, 2:]
    return llf_all, (g_all if eval_grad else None), st_all
