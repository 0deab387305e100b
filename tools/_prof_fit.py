import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mipego_b200 as mb
for n, d in ((500, 10), (1000, 10)):
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, (n, d)); y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0]); y = (y - y.mean()) / y.std()
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=[0.05] * d,
                            thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=1e-6, random_start=10)
    gp._check_data(X, y)
    for B in (10, 1, 10, 3):
        P = np.c_[0.05 * rng.uniform(0.5, 2, (B, d)), np.full(B, 0.8)]
        gp.log_likelihood_batch(P)
        t0 = time.perf_counter()
        for _ in range(100):
            gp.log_likelihood_batch(P)
        print("n=%d B=%d: %.3f ms per log_likelihood_batch call" % (n, B, (time.perf_counter() - t0) * 10))
    np.random.seed(1)
    t0 = time.perf_counter(); gp.fit(X, y); t1 = time.perf_counter()
    print("fit %.3f s, evals %d, per-restart funcalls %s" % (t1 - t0, gp.eval_count, [i["funcalls"] for i in gp.opt_info_]))
