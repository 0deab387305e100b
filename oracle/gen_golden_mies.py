"""Golden fixtures for the vectorised MIES: the UNMODIFIED reference
`mipego.optimizer.mies.MIES` (optimizer/mies.py:21-316) is run for a few generations on a box with
its random draws SCRIPTED, and every generation's offspring are frozen.

TEST INFRASTRUCTURE ONLY (container-only: needs /root/reference).  Run:  python oracle/gen_golden_mies.py

The reference draws per offspring, in this order (mies.py:306-309, :160-172, :210-218):
  randint(0, mu) -> p1, randint(0, mu) -> p2, [randn(dim) if p1 != p2: dominant recombination],
  randn() (global step-size factor), randn(N_r) (individual factors), randn(N_r) (the mutation).
`mip-ego_b200/optimizer.py:mies_argmax` draws whole (runs, lambda[, d]) arrays per generation instead.
The script pre-generates those arrays, feeds them to the reference in its call order through
patched module-level `randint / randn`, and stores arrays + outcomes in tests/golden/mies_*.npz;
tests/test_optimizer_cpu.py replays the same arrays through the vectorised implementation and requires
the same offspring in every generation, the same incumbent, evaluation count and stop reason.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class Script(object):
    """Hands the pre-generated numbers to the reference in ITS call order."""

    def __init__(self, P1, P2, Ztake, Ztau, Ztaup, Zmut):
        self.q = []
        G, lam = P1.shape
        for g in range(G):
            for i in range(lam):
                self.q.append(("int", int(P1[g, i])))
                self.q.append(("int", int(P2[g, i])))
                if P1[g, i] != P2[g, i]:
                    self.q.append(("vec", Ztake[g, i].copy()))
                self.q.append(("scalar", float(Ztau[g, i])))
                self.q.append(("vec", Ztaup[g, i].copy()))
                self.q.append(("vec", Zmut[g, i].copy()))
        self.pos = 0

    def _next(self, kind):
        k, v = self.q[self.pos]
        assert k == kind, "reference drew %s where the script expected %s (draw %d)" % (kind, k, self.pos)
        self.pos += 1
        return v

    def randint(self, lo, hi=None, size=None):
        assert size is None
        return self._next("int")

    def randn(self, *shape):
        if len(shape) == 0:
            return self._next("scalar")
        v = self._next("vec")
        assert v.shape == tuple(shape)
        return v


def case(name, seed, minimize, G=12, mu=4, lam=10):
    import mipego.optimizer.mies as ref_mies
    from mipego.SearchSpace import ContinuousSpace
    rng = np.random.default_rng(seed)
    lb = np.array([-5.0, -2.0, 0.0, 1.0])
    ub = np.array([5.0, 3.0, 4.0, 9.0])
    d = len(lb)
    centre = lb + (ub - lb) * np.array([0.3, 0.9, 0.05, 0.5])     # optimum near two faces: reflections happen
    sgn = 1.0 if minimize else -1.0

    def obj(x):
        x = np.asarray(x, dtype=np.float64)
        return float(sgn * np.sum((x - centre) ** 2 * np.arange(1, d + 1)))

    P1 = rng.integers(0, mu, (G, lam))
    P2 = rng.integers(0, mu, (G, lam))
    Ztake, Ztaup = rng.standard_normal((G, lam, d)), rng.standard_normal((G, lam, d))
    Ztau = rng.standard_normal((G, lam))
    Zmut = 3.0 * rng.standard_normal((G, lam, d))                 # wide steps: leave the box now and then
    space = ContinuousSpace([[float(a), float(b)] for a, b in zip(lb, ub)])
    np.random.seed(seed)
    max_eval = mu + G * lam - 1                                    # stop() fires after G generations at the latest
    opt = ref_mies.MIES(space, obj, max_eval=max_eval, minimize=minimize, mu_=mu, lambda_=lam)
    pop0 = np.array([[float(v) for v in opt.pop[k][:d]] for k in range(mu)])
    sig0 = np.array([[float(v) for v in opt.pop[k][d:2 * d]] for k in range(mu)])
    fit0 = np.array(opt.fitness, dtype=np.float64)
    script = Script(P1, P2, Ztake, Ztau, Ztaup, Zmut)
    saved = (ref_mies.randint, ref_mies.randn)
    ref_mies.randint, ref_mies.randn = script.randint, script.randn
    offspring = []
    select0 = opt.select

    def select_and_record():
        offspring.append(np.array([[float(v) for v in opt.offspring[k][:2 * d]] for k in range(lam)]))
        select0()
    opt.select = select_and_record
    try:
        xopt, fopt, stop = opt.optimize()
    finally:
        ref_mies.randint, ref_mies.randn = saved
    off = np.array(offspring)
    reasons = sorted(k for k, v in stop.items() if v is True)
    np.savez(os.path.join(OUT, name + ".npz"), lb=lb, ub=ub, centre=centre, minimize=minimize, mu=mu, lam=lam,
             max_eval=max_eval, pop0=pop0, sig0=sig0, fit0=fit0, P1=P1, P2=P2, Ztake=Ztake, Ztau=Ztau, Ztaup=Ztaup,
             Zmut=Zmut, offspring=off, xopt=np.array(xopt, dtype=np.float64), fopt=float(fopt),
             funcalls=int(stop["funcalls"]), reasons=np.array(reasons), generations=len(offspring))
    print("%s: %d generations, fopt %.6g, funcalls %d, stop %s, draws used %d / %d" %
          (name, len(offspring), fopt, stop["funcalls"], reasons, script.pos, len(script.q)))


def main():
    os.makedirs(OUT, exist_ok=True)
    load_reference()
    case("mies_min", seed=901, minimize=True)
    case("mies_max", seed=902, minimize=False)


if __name__ == "__main__":
    main()
