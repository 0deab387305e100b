"""Two ranks on two GPUs (NCCL): the library's own communicator (mgp_comm_init), the HBM -> HBM model
broadcast (mgp_bcast_model) and the device top-k all-gather + merge (mgp_allgather_topk_dev) against
the single-GPU results.  Skipped with fewer than two GPUs (the driver's multi-GPU runs and
`gpurun --gpus 2` exercise it)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    rng = np.random.default_rng(5)
    n, d = 900, 6
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    y = (y - y.mean()) / y.std()
    theta = (0.4 / d) * rng.uniform(0.5, 1.5, d)
    Xs = rng.uniform(-5, 5, (50_001, d))
    ts = np.array([0.7, 2.0, 3.1])
    return X, y, theta, Xs, ts


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import mipego_b200 as mb
        from mipego_b200 import dist as D
        X, y, theta, Xs, ts = _problem()
        d = X.shape[1]
        gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=theta, thetaL=[1e-5] * d,
                                thetaU=[100.0] * d, nugget=1e-6, device=rank)
        if rank == 0:
            gp.fit_fixed(X, y, np.r_[theta, 0.9])
        else:
            gp._check_data(X, y)
        D.init_comm(gp)
        info = gp._handle().lib.mgp_comm_info(gp._handle().h).decode()
        D.broadcast_fitted_state(gp, src=0)
        assert gp.is_fitted and D.has_comm(gp)
        mean, mse = gp.predict(Xs[:4000], eval_MSE=True)
        crit = mb.MGFI(model=gp, minimize=True, plugin=float(y.min()), t=1.0)
        # host-array sharding (every rank passes the whole population)
        hv, hi = D.sharded_topk(crit, Xs, k=5, params=ts)
        # device-resident sharding: this rank's shard only
        lo, hi_ = D.shard_range(len(Xs), rank, world)
        shard = torch.from_numpy(Xs[lo:hi_]).cuda()
        dv, di = D.sharded_topk_device(crit, shard, lo, k=5, params=ts)
        torch.cuda.synchronize()
        # an un-fitted root or mismatching data must fail on EVERY rank, not hang
        other = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=theta,
                                   thetaL=[1e-5] * d, thetaU=[100.0] * d, nugget=1e-6, device=rank)
        other._check_data(X if rank == 0 else X[::-1].copy(), y if rank == 0 else y[::-1].copy())
        D.init_comm(other)
        if rank == 0:
            other.fit_fixed(X, y, np.r_[theta, 0.9])
        try:
            D.broadcast_fitted_state(other, src=0)
            failed = False
        except Exception as e:
            failed = "different training data" in str(e)
        # the noisy-mode hyper-parameters travel with the state: update(reoptimize=False) works on the receiver
        Xn = np.r_[X, np.random.default_rng(9).uniform(-5, 5, (3, d))]
        yn = np.r_[y, 0.1, -0.2, 0.3]
        gp.update(Xn, yn, reoptimize=False)
        m2 = gp.predict(Xs[:1000])
        q.put((rank, info, mean, mse, hv, hi, dv.cpu().numpy(), di.cpu().numpy(), failed, m2,
               int(gp._handle().counter("bcast_bytes"))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_broadcast_and_topk():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    import mipego_b200 as mb
    X, y, theta, Xs, ts = _problem()
    d = X.shape[1]
    ref = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=theta, thetaL=[1e-5] * d,
                             thetaU=[100.0] * d, nugget=1e-6).fit_fixed(X, y, np.r_[theta, 0.9])
    rm, rs = ref.predict(Xs[:4000], eval_MSE=True)
    crit = mb.MGFI(model=ref, minimize=True, plugin=float(y.min()), t=1.0)
    tv, ti = crit.topk(Xs, k=5, params=ts)
    print("NCCL:", res[0][1], "| broadcast payload", res[1][10], "bytes")
    for rank, info, mean, mse, hv, hi, dv, di, failed, m2, nbytes in res:
        # bit-identical model on every rank: the same numbers as the single-GPU run, bit for bit
        assert np.array_equal(mean, rm) and np.array_equal(mse, rs)
        assert np.array_equal(hi, ti) and np.array_equal(hv, tv)
        assert np.array_equal(di, ti) and np.array_equal(dv, tv)
        assert failed, "a data mismatch must raise on every rank"
        assert "world=2" in info and nbytes > 900 * 900 * 8 * 2
    assert np.array_equal(res[0][9], res[1][9])           # the appended model agrees across ranks too
