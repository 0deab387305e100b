"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in mipego_b200/dist.py: shard
partition, fitted-state broadcast, all-gather + merge of per-shard top-k, sharded NLL rows.
The compute calls are replaced by numpy stand-ins: this covers the exchange, not the kernels."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeCrit:
    """criterion.topk stand-in: value = -|x|^2 * (1 + set index)."""

    def topk(self, X, k=1, params=None):
        P = np.atleast_1d([1.0] if params is None else params)
        vals = np.stack([-(X ** 2).sum(1) * (1 + s) for s in range(len(P))])
        k = min(k, X.shape[0])
        idx = np.stack([np.lexsort((np.arange(X.shape[0]), -v))[:k] for v in vals])
        return np.take_along_axis(vals, idx, 1), idx


class _Mean:
    beta = np.array([[0.25]])


class _FakeGP:
    def __init__(self, rank, n=6, d=2):
        rng = np.random.default_rng(0)
        self.X = rng.uniform(size=(n, d))
        self.y = rng.uniform(size=(n, 1))
        self.mean = _Mean()
        self.fitted = None
        if rank == 0:
            self._st = dict(theta=np.array([0.3, 0.7]), L=np.tril(rng.uniform(size=(n, n))) + np.eye(n),
                            gamma=rng.uniform(size=n))
            self.sigma2 = np.array([1.5])

    def _state(self, k):
        return self._st[k]

    @property
    def C(self):
        return self._st["L"]

    @property
    def gamma(self):
        return self._st["gamma"].reshape(-1, 1)

    def set_fitted_state(self, X, y, theta, L, gamma, sigma2, beta):
        self.fitted = (theta.copy(), L.copy(), gamma.copy(), sigma2, beta)

    def log_likelihood_batch(self, P, eval_grad=True):
        P = np.atleast_2d(P)
        return -(P ** 2).sum(1), -2 * P, (P[:, 0] > 0.9).astype(np.int32)


def _ms_objective(P):
    """Batched (f, grad) of a multimodal function of log10-parameters (stand-in for -llf)."""
    P = np.atleast_2d(P)
    f = np.sum(P ** 2, axis=1) + np.sum(np.cos(4 * P), axis=1)
    g = 2 * P - 4 * np.sin(4 * P)
    return f, g


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mipego_b200 import dist as D
        rng = np.random.default_rng(7)
        X = rng.uniform(-1, 1, (1001, 3))
        v, i = D.sharded_topk(_FakeCrit(), X, k=4, params=[1.0, 2.0])
        P = rng.uniform(size=(5, 3))
        llf, grad, st = D.sharded_nll(_FakeGP(1), P)
        gp = _FakeGP(rank)
        D.broadcast_fitted_state(gp, src=0)
        # restarts sharded over ranks: 5 starts of a multimodal objective, 3 + 2
        starts = np.random.default_rng(3).uniform(-2, 2, (5, 2))
        ms = D.sharded_multistart(_ms_objective, starts, np.array([[-2.0, 2.0]] * 2), 200, 5)
        q.put((rank, v, i, llf, grad, st, gp.fitted, ms[:3]))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from mipego_b200.dist import shard_range
    for m in (0, 1, 7, 8, 1001):
        for w in (1, 2, 3, 8):
            parts = [shard_range(m, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == m
            assert all(parts[r][1] == parts[r + 1][0] for r in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_merge_topk_order_and_ties():
    from mipego_b200.dist import merge_topk
    vals = np.array([[[3.0, 1.0]], [[3.0, 2.0]]])      # (world=2, sets=1, k=2)
    idx = np.array([[[5, 9]], [[2, 11]]])
    v, i = merge_topk(vals, idx, 3)
    assert v.tolist() == [[3.0, 3.0, 2.0]] and i.tolist() == [[2, 5, 11]]
    vals[0, 0, 0] = np.nan
    v, i = merge_topk(vals, idx, 2)
    assert i.tolist() == [[2, 11]]


@pytest.mark.timeout(120)
def test_world2_gloo_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    X = rng.uniform(-1, 1, (1001, 3))
    ref_v, ref_i = _FakeCrit().topk(X, k=4, params=[1.0, 2.0])
    P = rng.uniform(size=(5, 3))
    from mipego_b200._mle import multistart_lbfgsb
    starts = np.random.default_rng(3).uniform(-2, 2, (5, 2))
    one_p, one_f, one_n, _ = multistart_lbfgsb(_ms_objective, starts, np.array([[-2.0, 2.0]] * 2), 200, 5,
                                               mode="batched")
    for r in res:
        p, f, n_evals = r[7]
        assert f == one_f and np.array_equal(p, one_p) and n_evals == one_n   # == unsharded run
    for rank, v, i, llf, grad, st, fitted, _ms in res:
        assert np.array_equal(i, ref_i) and np.array_equal(v, ref_v)      # same as the unsharded run
        assert np.allclose(llf, -(P ** 2).sum(1)) and np.allclose(grad, -2 * P)
        assert np.array_equal(st, (P[:, 0] > 0.9).astype(np.int32))
    g0 = _FakeGP(0)
    th, L, gm, s2, b = res[1][6]                                            # adopted on rank 1
    assert np.array_equal(th, g0._st["theta"]) and np.array_equal(L, g0._st["L"])
    assert np.array_equal(gm, g0._st["gamma"]) and s2 == 1.5 and b == 0.25
    assert res[0][6] is None                                                # src keeps its own state
