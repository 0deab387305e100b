import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, mipego_b200 as mb
cases = [(200, 4, "squared_exponential", 3000), (500, 10, "squared_exponential", 100000), (700, 6, "matern", 20000),
         (1100, 7, "matern", 50000), (4096, 20, "squared_exponential", 151552)]
if len(sys.argv) > 1: cases = cases[:int(sys.argv[1])]
for n, d, kind, m in cases:
    rng = np.random.default_rng(1); X = rng.uniform(-5, 5, (n, d)); y = (X**2).sum(1)+3*np.sin(X[:,0]); y=(y-y.mean())/y.std()
    theta = (0.12/d)*rng.uniform(0.5,1.5,d)*(4.0 if kind=="matern" else 1.0)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr=kind, theta0=theta, thetaL=1e-5*np.ones(d), thetaU=100*np.ones(d), nugget=1e-6)
    gp.fit_fixed(X, y, np.r_[theta, 0.9]); h = gp._handle()
    Xs = rng.uniform(-5, 5, (m, d)); Xs[:50] = X[:50] + 1e-4   # some candidates near training points
    Xd = torch.from_numpy(Xs).cuda()
    ei = mb.EI(model=gp, minimize=True)
    out = {}
    for mode in (0, 1):
        h.set_option("trmm_i8", mode)
        mean, mse = gp.predict_device(Xd); torch.cuda.synchronize()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(2): ei.topk_device(Xd, k=1)
            torch.cuda.synchronize()
            h.set_option("time_kernels", 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): ei.topk_device(Xd, k=1)
            e1.record(); torch.cuda.synchronize()
        out[mode] = (mean.cpu().numpy(), mse.cpu().numpy(), e0.elapsed_time(e1)/3, h.counter("trmm_ns")/3e6, h.counter("rgen_ns")/3e6)
        h.set_option("time_kernels", 0)
    s2 = float(np.ravel(gp.sigma2)[0])
    err = np.max(np.abs(out[1][1]-out[0][1])/(np.abs(out[0][1])+1e-6*s2))
    print("n=%d d=%d %s m=%d: mse rel diff i8 vs dmma %.2e (mean equal %s) | step %.3f -> %.3f ms, trmm %.3f -> %.3f ms (x%.2f), rgen %.3f" %
          (n, d, kind[:6], m, err, np.array_equal(out[0][0], out[1][0]), out[0][2], out[1][2], out[0][3], out[1][3], out[0][3]/out[1][3], out[0][4]), flush=True)
