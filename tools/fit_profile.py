"""Where the time of GaussianProcess.fit goes (n = 500, d = 10, 10 lock-step restarts): wall time, the share spent
inside mgp_nll_batch, and a cProfile listing.  Output committed as profiles/r02_fit_profile.txt."""
import os, sys, time, cProfile, pstats
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mipego_b200 as mb
n, d, R = 500, 10, 10
rng = np.random.default_rng(0); X = rng.uniform(-5, 5, (n, d)); y = (X**2).sum(1)+3*np.sin(X[:,0]); y=(y-y.mean())/y.std()
def mk():
    np.random.seed(1)
    return mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=[0.05]*d, thetaL=[1e-4]*d, thetaU=[10.0]*d, nugget=1e-6, random_start=R)
g = mk(); g.fit(X, y)
g = mk(); g._h = None
t0=time.perf_counter(); g.fit(X,y); print("fit", (time.perf_counter()-t0)*1e3, "ms", g.eval_count, "evals")
g2 = mk(); g2._h = g._h
# time spent inside the library calls
import mipego_b200.gpr as G
orig = G.GaussianProcess.log_likelihood_batch
acc = [0.0, 0, 0]
def timed(self, P, eval_grad=True):
    t=time.perf_counter(); r=orig(self, P, eval_grad); acc[0]+=time.perf_counter()-t; acc[1]+=1; acc[2]+=len(np.atleast_2d(P)); return r
G.GaussianProcess.log_likelihood_batch = timed
t0=time.perf_counter(); g2.fit(X,y); tot=time.perf_counter()-t0
print("fit %.1f ms: %d batch calls (%d evals) take %.1f ms (%.3f ms per call)" % (tot*1e3, acc[1], acc[2], acc[0]*1e3, acc[0]/acc[1]*1e3))
G.GaussianProcess.log_likelihood_batch = orig
g3 = mk(); g3._h = g._h
pr = cProfile.Profile(); pr.enable(); g3.fit(X, y); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
