"""Likelihood + gradient at BO-loop sizes (the regime bound by the serial chain of short launches): device time per
batched call (CUDA-graph replay, CUDA events on the launching stream) with the triangular inversion after the
factorisation (trtri_rowwise = 0) and beside it (1).  The diagonal-block kernel version is a process-wide switch
(MGP_DIAG_V1=1 selects version 1), so run the script once per version.
Output: profiles/r02_nll_chain.txt."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mipego_b200 as mb

d = 20
ver = "v1" if os.environ.get("MGP_DIAG_V1", "") == "1" else "v2"
sizes = [int(a) for a in sys.argv[1:]] or [500, 1024, 2048]
for n in sizes:
    rng = np.random.default_rng(30)
    X = rng.uniform(-5, 5, (n, d)); y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0]); y = (y - y.mean()) / y.std()
    for B in (1, 4, 8, 64):
        if B == 64 and n > 2048:
            continue
        theta = (0.12 / d) * 10.0 ** rng.uniform(-0.5, 0.5, (B, d)); P = np.c_[theta, rng.uniform(0.5, 1.0, B)]
        gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=P[0, :d],
                                thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6)
        gp._check_data(X, y); h = gp._handle()
        Pd = torch.from_numpy(P).cuda(); st = torch.cuda.Stream()
        res = {}
        for row in (0, 1):
            h.set_option("trtri_rowwise", row)
            with torch.cuda.stream(st):
                for _ in range(4): out = gp.log_likelihood_batch_device(Pd)
                torch.cuda.synchronize()
                reps = max(5, int(300 / (max(n, 512) / 512) ** 3 / max(B / 8, 1)))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps): gp.log_likelihood_batch_device(Pd)
                e1.record(); torch.cuda.synchronize()
            res[row] = (e0.elapsed_time(e1) / reps, out[0].cpu().numpy().copy())
        fl = (n ** 3 + 3.5 * n * n * d) * B
        dl = float(np.max(np.abs(res[0][1] - res[1][1]) / np.abs(res[0][1])))
        print("diag %s n=%d B=%d: after %.3f ms | beside %.3f ms (%.1f%% of 37.07 TF) | llf rel diff %.1e" %
              (ver, n, B, res[0][0], res[1][0], fl / min(res[0][0], res[1][0]) * 1e-9 / 37.07 * 100, dl), flush=True)
