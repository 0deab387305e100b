import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, mipego_b200 as mb
from oracle import gp_oracle as O
for n, d, kind, m in ((500, 10, "squared_exponential", 262144), (500, 10, "matern", 262144), (2048, 20, "squared_exponential", 65536), (2048, 40, "matern", 32768), (1024, 64, "absolute_exponential", 32768)):
    rng = np.random.default_rng(1); X = rng.uniform(-5, 5, (n, d)); y = (X**2).sum(1)+3*np.sin(X[:,0]); y=(y-y.mean())/y.std()
    theta = (0.12/d)*rng.uniform(0.5,1.5,d)*(4.0 if kind=="matern" else 1.0)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr=kind, theta0=theta, thetaL=1e-5*np.ones(d), thetaU=100*np.ones(d), nugget=1e-6)
    gp.fit_fixed(X, y, np.r_[theta, 0.9]); h = gp._handle()
    ei = mb.EI(model=gp, minimize=True)
    Xs = rng.uniform(-5, 5, (m, d)); Xd = torch.from_numpy(Xs).cuda()
    st = O.fit_state(X, y, np.r_[theta, 0.9], kind, True, 0.0, "noisy", 1e-6)
    v, dx = ei.evaluate(Xs[:256], return_dx=True)
    rv, rdx = O.acquisition(st, Xs[:256], "EI", None, None, True, return_dx=True)
    err = np.max(np.abs(dx[0]-rdx))/np.abs(rdx).max()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): ei.evaluate_device(Xd, return_dx=True)
        torch.cuda.synchronize()
        h.set_option("time_kernels", 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ei.evaluate_device(Xd, return_dx=True)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/5
    print("n=%d d=%d %s m=%d: %.3f ms/step %.3g cand/s | rgen %.3f trmm %.3f grad %.3f ms | dx err %.2e" % (n, d, kind[:6], m, ms, m/ms*1e3, h.counter("rgen_ns")/5e6, h.counter("trmm_ns")/5e6, h.counter("epilogue_ns")/5e6, err), flush=True)
    h.set_option("time_kernels", 0)
    # single point latency
    x1 = Xs[0]
    for _ in range(5): ei(x1, return_dx=True)
    t0 = time.perf_counter()
    for _ in range(200): ei(x1, return_dx=True)
    print("   single-point EI+dx latency %.1f us" % ((time.perf_counter()-t0)/200*1e6), flush=True)
