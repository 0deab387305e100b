"""Multi-GPU plumbing for the GP hot path: one process per GPU, `torch.distributed` (NCCL over
NVLink on the GPU box, gloo in the CPU tests).

The path shards naturally: candidates are independent units, MLE restarts are
independent units.  The only exchanges are
  * one broadcast of the fitted state per fit, so that every rank holds the bit-identical model
    (results are then independent of the number of shards), and
  * an all-gather of each rank's k best (value, global index) pairs, merged identically on every
    rank (value descending, ties -> smaller global index, the order the single-GPU kernel uses).
No collective sits inside the compute kernels; the data path itself needs none.

Two transports:
  * the library's own NCCL communicator (`init_comm`: mgp_comm_init / mgp_bcast_model /
    mgp_allgather_topk in include/mgp.h) -- the model travels HBM -> HBM with no host bounce and no
    re-factorisation on the receivers, the top-k merge runs on the device.  torch.distributed is
    then only the rendezvous that ships the 128-byte NCCL id and a few Python scalars;
  * plain torch.distributed tensors (gloo in the CPU tests, or NCCL) when no communicator was
    created: host-side pack / unpack + `set_fitted_state`.
"""
import ctypes

import numpy as np

_PY_STATE = ("par", "_par_fitted", "theta_", "sigma2", "noise_var", "_env_noise_var", "log_likelihood_",
             "thetaU", "_nugget")


def shard_range(m, rank, world):
    """Contiguous, balanced partition of m units: rank r gets [lo, hi)."""
    base, rem = divmod(int(m), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch
    import torch.distributed as dist
    return torch, dist


def has_comm(gp):
    """True when gp's device handle owns an NCCL communicator (init_comm)."""
    h = getattr(gp, "_h", None)
    return h is not None and h.counter("world") > 1


def init_comm(gp):
    """Create the library's NCCL communicator for gp's handle over the torch.distributed world
    (collective).  The 128-byte ncclUniqueId is made on rank 0 (mgp_comm_unique_id) and shipped with
    torch.distributed -- any backend; that is all torch does for the exchanges that follow."""
    torch, dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    h = gp._handle()
    if world == 1:
        return gp
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        rc = h.lib.mgp_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError("mgp_comm_unique_id failed (%d): %s" % (rc, h.lib.mgp_comm_info(h.h).decode()))
    box = [buf.raw if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    h.check(h.lib.mgp_comm_init(h.h, rank, world, box[0]))
    return gp


def _broadcast_model_nccl(gp, src):
    """HBM -> HBM broadcast of the fitted model through the library's communicator."""
    torch, dist = _dist()
    rank = dist.get_rank()
    h = gp._handle()
    h.check(h.lib.mgp_bcast_model(h.h, int(src)))
    # the Python-side attributes (hyper-parameters, typed scalars): a few hundred bytes
    box = [None]
    if rank == src:
        box[0] = {k: getattr(gp, k) for k in _PY_STATE if hasattr(gp, k)}
        box[0]["beta"] = None if gp.mean.beta is None else np.array(gp.mean.beta, dtype=np.float64)
        box[0]["estimate_trend"] = gp.estimate_trend
    dist.broadcast_object_list(box, src=src)
    if rank != src:
        st = dict(box[0])
        beta, est = st.pop("beta"), st.pop("estimate_trend")
        gp.__dict__.update(st)
        if est and beta is not None:
            gp.mean.beta = beta
        gp._cache = {}
        gp._host_state = None
        gp.is_fitted = True
    return gp


def broadcast_fitted_state(gp, src=0, device=None):
    """Give every rank the bit-identical fitted model of `src`.  Every rank must hold the same
    training data X, y (they are part of the optimiser state in the BO loop; non-source ranks call
    `gp._check_data(X, y)` first).  With a communicator (`init_comm`) the model goes HBM -> HBM
    (mgp_bcast_model); otherwise (theta, L, gamma, sigma2, beta, par) travel as one torch tensor and
    are adopted through `set_fitted_state` (mgp_set_state)."""
    torch, dist = _dist()
    if has_comm(gp):
        return _broadcast_model_nccl(gp, src)
    rank = dist.get_rank()
    n, d = gp.X.shape
    p = int(getattr(gp.mean, "n_dim", 1) or 1)      # trend coefficients (1 constant, d + 1 linear)
    dev = device if device is not None else "cpu"
    # full hyper-parameter vector (theta + sigma2 / alpha / noise_var as the mode has them), so that
    # update(reoptimize=False) works on the receivers in every estimation mode
    n_extra = len(gp._par_names()) - 1 if hasattr(gp, "_par_names") else 0
    if rank == src:
        if n_extra:
            extra = np.asarray(gp._par_vector(), dtype=np.float64)[-n_extra:]
        else:
            extra = np.empty(0)
        pack = [np.asarray(gp._state("theta"), dtype=np.float64).ravel(),
                np.ascontiguousarray(gp.C, dtype=np.float64).ravel(),
                np.asarray(gp.gamma, dtype=np.float64).ravel(),
                np.r_[float(np.ravel(gp.sigma2)[0]), np.ravel(gp.mean.beta).astype(np.float64)], extra]
        flat = np.concatenate(pack)
    else:
        flat = np.empty(d + n * n + n + 1 + p + n_extra)
    t = torch.from_numpy(flat).to(dev)
    dist.broadcast(t, src=src)
    if rank != src:
        flat = t.cpu().numpy()
        theta, L = flat[:d], flat[d:d + n * n].reshape(n, n)
        gamma, sc = flat[d + n * n:d + n * n + n], flat[d + n * n + n:d + n * n + n + 1 + p]
        kw = {}
        if n_extra:
            n_theta = int(np.size(getattr(gp, "theta0", theta)))
            kw["par"] = np.r_[theta[:n_theta] if n_theta == 1 else theta, flat[-n_extra:]]
        gp.set_fitted_state(gp.X, gp.y.ravel(), theta, L, gamma, float(sc[0]),
                            float(sc[1]) if p == 1 else sc[1:].copy(), **kw)
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


def allgather_topk(local_vals, local_idx, k, offset, device=None, gp=None):
    """All-gather each rank's (n_sets, k) best candidates (local indices + `offset` = global) and
    merge.  Every rank returns the same (values, global indices).  With `gp` holding a communicator
    the exchange and the merge run on the device (mgp_allgather_topk)."""
    if gp is not None and has_comm(gp):
        from . import _cabi
        h = gp._handle()
        lv = np.ascontiguousarray(local_vals, dtype=np.float64)
        li = np.ascontiguousarray(local_idx, dtype=np.int64)
        ns, kk = lv.shape
        ov, oi = np.empty((ns, kk)), np.empty((ns, kk), dtype=np.int64)
        h.check(h.lib.mgp_allgather_topk(h.h, ns, kk, _cabi.dptr(lv), li.ctypes.data_as(_cabi.c_lp), int(offset),
                                         _cabi.dptr(ov), oi.ctypes.data_as(_cabi.c_lp)))
        return ov[:, :k], oi[:, :k]
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
    return allgather_topk(v, i, k, lo, device=device, gp=getattr(criterion, "model", None))


def sharded_topk_device(criterion, X_shard, offset, k=1, params=None, out=None):
    """Device-resident population arg-max across ranks: `X_shard` is THIS rank's (m_r, d) CUDA tensor
    of candidates whose row 0 has global index `offset`.  The local top-k (mgp_acq_topk_dev), the
    NCCL all-gather and the merge (mgp_allgather_topk_dev) are enqueued on torch's current stream;
    nothing synchronises.  Returns CUDA tensors (values (n_sets, k), global indices (n_sets, k)),
    identical on every rank.  Needs `init_comm(criterion.model)`."""
    import torch
    gp = criterion.model
    h = gp._handle()
    if h.counter("world") <= 1:
        raise RuntimeError("sharded_topk_device needs dist.init_comm(model)")
    tv, ti = criterion.topk_device(X_shard, k=k, params=params)
    ns = tv.shape[0]
    if out is None:
        out = (torch.empty((ns, k), dtype=torch.float64, device=X_shard.device),
               torch.empty((ns, k), dtype=torch.int64, device=X_shard.device))
    stream = torch.cuda.current_stream(X_shard.device).cuda_stream
    h.check(h.lib.mgp_allgather_topk_dev(h.h, ns, k, tv.data_ptr(), ti.data_ptr(), int(offset),
                                         out[0].data_ptr(), out[1].data_ptr(), stream))
    # tv / ti are read by the enqueued kernels: keep them alive until the stream has passed them
    tv.record_stream(torch.cuda.current_stream(X_shard.device))
    ti.record_stream(torch.cuda.current_stream(X_shard.device))
    return out


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


def sharded_multistart(batch_fn, starts, log10bounds, eval_budget, wait_iter, device=None):
    """Multi-start L-BFGS-B with the restarts partitioned over ranks ("MLE:
    partition the B restart vectors by rank"): rank r runs rows shard_range(R, r, world) of
    `starts` in lock-step on its own GPU, the per-rank winners (f, x) are all-gathered and reduced
    in rank order with the reference's `<=` rule (gpr.py:1045-1052), so every rank returns the
    result a single process would get from all R restarts.  Same signature/return as
    _mle.multistart_lbfgsb."""
    from ._mle import multistart_lbfgsb
    torch, dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    starts = np.atleast_2d(np.asarray(starts, dtype=np.float64))
    R, n_par = starts.shape
    lo, hi = shard_range(R, rank, world)
    row = np.full(n_par + 3, np.nan)
    infos = []
    if hi > lo:
        p, f, n_evals, infos = multistart_lbfgsb(batch_fn, starts[lo:hi], log10bounds, eval_budget,
                                                 wait_iter, mode="batched")
        row[0], row[1], row[2], row[3:] = f, n_evals, 1.0, p
    else:
        row[0], row[1], row[2] = np.inf, 0.0, 0.0
    dev = device if device is not None else "cpu"
    mine = torch.from_numpy(row).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    best_p, best_f, total = None, np.inf, 0
    for r in range(world):
        blk = out[r].cpu().numpy()
        if blk[2] == 0.0:
            continue
        total += int(blk[1])
        if best_p is None or blk[0] <= best_f:
            best_p, best_f = blk[3:].copy(), float(blk[0])
    return best_p, best_f, total, infos


def sharded_fit(gp, X, y, device=None):
    """GaussianProcess.fit with the MLE restarts sharded over the ranks.  Every rank must pass the
    same X, y and hold the same numpy global RNG state (the restart points are drawn from it,
    gpr.py:1035-1036, identically on every rank).  Each rank then builds the fitted state for the
    winning hyper-parameters itself -- same inputs, same kernels, hence the bit-identical model."""
    return gp.fit(X, y, _solve=lambda *a: sharded_multistart(*a, device=device))
