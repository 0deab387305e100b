"""Where a single-point EI(x) call spends its time (n = 500, d = 10): the Python class layer, the bare C-ABI call
(ctypes, prebuilt pointers), and the device chain alone (CUDA events around mgp_acq_dev on a torch stream)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mipego_b200 as mb
from mipego_b200 import _cabi

n, d = 500, 10
rng = np.random.default_rng(0)
X = rng.uniform(-5, 5, (n, d))
y = (X ** 2).sum(1)
y = (y - y.mean()) / y.std()
theta = (0.12 / d) * rng.uniform(0.5, 1.5, d)
gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=theta,
                        thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6)
gp.fit_fixed(X, y, np.r_[theta, 0.9])
ei = mb.EI(model=gp, minimize=True)
h = gp._handle()
x = np.ascontiguousarray(rng.uniform(-5, 5, (1, d)))
val = np.empty((1, 1))
dx = np.empty((1, 1, d))
P = np.zeros(1)
N = 2000
for _ in range(50):
    ei(x[0])
t0 = time.perf_counter()
for _ in range(N):
    ei(x[0])
t_py = (time.perf_counter() - t0) / N
px, pv, pp, pdx = _cabi.dptr(x), _cabi.dptr(val), _cabi.dptr(P), _cabi.dptr(dx)
plug = float(ei._plugin_user())
t0 = time.perf_counter()
for _ in range(N):
    h.lib.mgp_acq(h.h, 1, px, _cabi.ACQ["EI"], 1, pp, plug, 1, pv, None)
t_c = (time.perf_counter() - t0) / N
t0 = time.perf_counter()
for _ in range(N):
    h.lib.mgp_acq(h.h, 1, px, _cabi.ACQ["EI"], 1, pp, plug, 1, pv, pdx)
t_cdx = (time.perf_counter() - t0) / N
t0 = time.perf_counter()
for _ in range(N):
    h.counter("launches")
t_noop = (time.perf_counter() - t0) / N
xd = torch.from_numpy(x).cuda()
out = torch.empty((1, 1), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(20):
    h.lib.mgp_acq_dev(h.h, 1, xd.data_ptr(), _cabi.ACQ["EI"], 1, pp, plug, 1, out.data_ptr(), None, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    h.lib.mgp_acq_dev(h.h, 1, xd.data_ptr(), _cabi.ACQ["EI"], 1, pp, plug, 1, out.data_ptr(), None, st)
e1.record()
torch.cuda.synchronize()
t_dev = e0.elapsed_time(e1) / 200 * 1e-3
t0 = time.perf_counter()
for _ in range(200):
    h.lib.mgp_acq_dev(h.h, 1, xd.data_ptr(), _cabi.ACQ["EI"], 1, pp, plug, 1, out.data_ptr(), None, st)
    torch.cuda.synchronize()
t_dev_sync = (time.perf_counter() - t0) / 200
print("EI(x) through the class %.1f us | bare mgp_acq %.1f us (with dx %.1f us) | a no-op ABI call %.1f us | "
      "mgp_acq_dev back to back (device chain, pipelined) %.1f us | mgp_acq_dev + synchronize %.1f us"
      % (t_py * 1e6, t_c * 1e6, t_cdx * 1e6, t_noop * 1e6, t_dev * 1e6, t_dev_sync * 1e6))
