"""Single-point call latency of the drop-in criteria (the call pattern of the reference's unmodified
argmax_restart: one candidate per call, optimizer/utils.py:38-42)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mipego_b200 as mb

for n, d in ((50, 2), (500, 10), (4096, 20)):
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1)
    y = (y - y.mean()) / y.std()
    theta = (0.12 / d) * rng.uniform(0.5, 1.5, d)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=theta,
                            thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6)
    gp.fit_fixed(X, y, np.r_[theta, 0.9])
    ei = mb.EI(model=gp, minimize=True)
    x = rng.uniform(-5, 5, d)
    for _ in range(20):
        ei(x); ei(x, return_dx=True)
    t0 = time.perf_counter()
    for _ in range(200):
        ei(x)
    t1 = time.perf_counter()
    for _ in range(200):
        ei(x, return_dx=True)
    t2 = time.perf_counter()
    Xb = rng.uniform(-5, 5, (256, d))
    ei(Xb, return_dx=True)
    t3 = time.perf_counter()
    for _ in range(50):
        ei(Xb, return_dx=True)
    t4 = time.perf_counter()
    print("n=%d d=%d: EI(x) %.1f us, EI(x, return_dx) %.1f us per call; 256-point batch with dx %.1f us"
          % (n, d, (t1 - t0) / 200 * 1e6, (t2 - t1) / 200 * 1e6, (t4 - t3) / 50 * 1e6))
    # the same calls through the copy engines (three streams, events) instead of the mapped pinned page
    h = gp._handle()
    v1, g1 = ei(x, return_dx=True)
    h.set_option("zero_copy", 0)
    v0, g0 = ei(x, return_dx=True)
    assert np.array_equal(v0, v1) and np.array_equal(g0, g1)
    t0 = time.perf_counter()
    for _ in range(200):
        ei(x)
    t1 = time.perf_counter()
    for _ in range(200):
        ei(x, return_dx=True)
    t2 = time.perf_counter()
    h.set_option("zero_copy", 1)
    print("            with option zero_copy = 0 (H2D / D2H copies on separate streams): EI(x) %.1f us, with dx %.1f us"
          % ((t1 - t0) / 200 * 1e6, (t2 - t1) / 200 * 1e6))
