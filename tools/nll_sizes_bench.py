"""Likelihood + gradient at n = 512 ... 4096 with 8 and 64 parameter vectors per call: device time (CUDA events on the
launching stream) and host time of the call, eager launches and CUDA-graph replay.  Output: profiles/r02_nll_sizes.txt."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, mipego_b200 as mb
d = 20
for n in (512, 1024, 2048, 4096):
    rng = np.random.default_rng(30); X = rng.uniform(-5, 5, (n, d)); y = (X**2).sum(1)+3*np.sin(X[:,0]); y=(y-y.mean())/y.std()
    for B in (8, 64):
        theta = (0.12/d)*10.0**rng.uniform(-0.5, 0.5, (B, d)); P = np.c_[theta, rng.uniform(0.5,1.0,B)]
        gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=P[0,:d], thetaL=1e-5*np.ones(d), thetaU=100*np.ones(d), nugget=1e-6)
        gp._check_data(X, y); h = gp._handle()
        Pd = torch.from_numpy(P).cuda(); st = torch.cuda.Stream()
        res = {}
        for graphs in (0, 1):
            h.set_option("nll_graphs", graphs)
            with torch.cuda.stream(st):
                for _ in range(4): gp.log_likelihood_batch_device(Pd)
                torch.cuda.synchronize()
                reps = max(3, int(200 / (n/512)**3 / (B/8)))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps): gp.log_likelihood_batch_device(Pd)
                e1.record(); torch.cuda.synchronize()
            dev_ms = e0.elapsed_time(e1)/reps
            t0 = time.perf_counter()
            for _ in range(reps): gp.log_likelihood_batch(P)
            host_ms = (time.perf_counter()-t0)/reps*1e3
            res[graphs] = (dev_ms, host_ms)
        fl = (n**3 + 3.5*n*n*d)*B
        print("n=%d B=%d: eager dev %.3f ms host %.3f ms | graph dev %.3f ms host %.3f ms | graph %.0f eval/s = %.1f%% of 37.07 TF" % (n, B, res[0][0], res[0][1], res[1][0], res[1][1], B/res[1][0]*1e3, fl/res[1][0]*1e-9/37.07*100), flush=True)
