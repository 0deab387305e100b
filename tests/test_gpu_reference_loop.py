"""The reference's OWN optimisation loop on the drop-ins (north_star: "the BO/ParallelBO loop
calling them"): the unmodified `mipego.BO` / `mipego.ParallelBO` classes (base.py:361-463,
BayesOpt.py:25-31, 84-101) are imported from the reference tree and run with
`mipego_b200.GaussianProcess` as the model and `mipego_b200.InfillCriteria` substituted for the
reference's criteria module -- the two-line switch of INTEGRATION.md.  The flows are the reference's
own smoke tests (test/test_BO.py:99-131 `test_continuous`, :189-221 ParallelBO + MGFI, :16-53
pickling).

Needs the reference tree: present in the build container (/root/reference), absent on the GPU box
unless tools/ship_reference_run.sh has placed its temporary copy; skipped otherwise."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference():
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present (run tools/ship_reference_run.sh for the shipped run)")
    return ref_shim.load_reference()


def _switch(mipego):
    """the criteria module the loop resolves names in."""
    import mipego_b200 as mb
    import mipego.base
    import mipego.BayesOpt
    mipego.base.InfillCriteria = mb.InfillCriteria
    mipego.BayesOpt.InfillCriteria = mb.InfillCriteria
    return mb


def _model(mb, dim, lb, ub, **kw):
    thetaL = 1e-10 * (ub - lb) * np.ones(dim)
    thetaU = 10 * (ub - lb) * np.ones(dim)
    theta0 = np.random.rand(dim) * (thetaU - thetaL) + thetaL
    args = dict(mean=mb.constant_trend(dim, beta=None), corr="squared_exponential", theta0=theta0, thetaL=thetaL,
                thetaU=thetaU, nugget=0, noise_estim=False, optimizer="BFGS", wait_iter=3, random_start=dim,
                likelihood="concentrated", eval_budget=100 * dim)
    args.update(kw)
    return mb.GaussianProcess(**args)


def test_reference_BO_runs_on_the_dropins():
    """test/test_BO.py:99-131 (`test_continuous`): GP + L-BFGS-B MLE, EI maximised by the reference's
    own argmax_restart (per-point criterion(x, return_dx=True) calls, optimizer/utils.py:38-42)."""
    mipego = _reference()
    mb = _switch(mipego)
    from mipego import BO, ContinuousSpace
    np.random.seed(123)
    dim, lb, ub = 5, -1, 5
    calls = []

    def fitness(x):
        x = np.asarray(x)
        calls.append(x.copy())
        return np.sum(x ** 2)

    model = _model(mb, dim, lb, ub)
    opt = BO(search_space=ContinuousSpace([lb, ub]) * dim, obj_fun=fitness, model=model, DoE_size=5, max_FEs=10,
             verbose=False, n_point=1)
    xopt, fopt, stop = opt.run()
    assert len(calls) == 10 and stop["max_FEs"] == 10
    assert model.is_fitted and model.X.shape == (10, dim)
    assert fopt == min(np.sum(np.asarray(c) ** 2) for c in calls)
    assert model._handle().counter("launches") > 100          # the model ran on the device
    assert isinstance(opt._create_acquisition().func, mb.EI)   # the loop resolves the CUDA criteria


def test_reference_ParallelBO_MGFI_and_pickling():
    """ParallelBO(n_point=3, acquisition_fun='MGFI') for 5 iterations (BayesOpt.py:56-101: three
    criteria with sampled t over the same fitted model), then the save / load round trip of
    base.py:575-608 (dill pickles the optimiser including the CUDA-backed model)."""
    mipego = _reference()
    mb = _switch(mipego)
    from mipego import ParallelBO, ContinuousSpace
    np.random.seed(42)
    dim, lb, ub = 3, -5, 5
    n_eval = [0]

    def fitness(x):
        n_eval[0] += 1
        x = np.asarray(x, dtype=float)
        return float(np.sum(x ** 2) + 3 * np.sin(x[0]))

    model = _model(mb, dim, lb, ub, nugget=1e-6, corr="matern", random_start=2)
    opt = ParallelBO(search_space=ContinuousSpace([lb, ub]) * dim, obj_fun=fitness, model=model, DoE_size=6,
                     max_FEs=6 + 3 * 5, verbose=False, n_point=3, acquisition_fun="MGFI", acquisition_par={"t": 2},
                     n_job=1)
    xopt, fopt, stop = opt.run()
    assert n_eval[0] == 21 and model.X.shape[0] == 21 and np.isfinite(fopt)
    import tempfile
    path = os.path.join(tempfile.mkdtemp(), "opt.dump")
    opt.save(path)
    opt2 = ParallelBO.load(path)
    Xs = np.random.default_rng(0).uniform(lb, ub, (50, dim))
    assert np.allclose(opt2.model.predict(Xs), model.predict(Xs), rtol=1e-9, atol=1e-12)


def test_reference_BO_with_batched_argmax_dropin():
    """Same loop with mipego_b200.argmax_restart substituted at base.py:338-347 (all L-BFGS-B
    restarts of the acquisition in lock-step): same signature, the loop is unchanged."""
    mipego = _reference()
    mb = _switch(mipego)
    import mipego.base
    from mipego import BO, ContinuousSpace
    old = mipego.base.argmax_restart
    mipego.base.argmax_restart = mb.argmax_restart
    try:
        np.random.seed(7)
        dim, lb, ub = 4, -5, 5
        model = _model(mb, dim, lb, ub, nugget=1e-6, random_start=3)
        opt = BO(search_space=ContinuousSpace([lb, ub]) * dim, obj_fun=lambda x: float(np.sum(np.asarray(x) ** 2)),
                 model=model, DoE_size=8, max_FEs=14, verbose=False, n_point=1)
        xopt, fopt, stop = opt.run()
        assert model.X.shape[0] == 14 and np.isfinite(fopt)
    finally:
        mipego.base.argmax_restart = old
