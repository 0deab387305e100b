#!/usr/bin/env python
"""bench.py -- the two halves of BASELINE.json's metric on B200:
"GP posterior+EI candidates/sec at n,d; GP NLL+grad evals/sec (1/2/4/8 B200)".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload c1|c2|c2dx|c3|c5|nll] [--no-extra]

Headline line (unchanged since round 1): configuration C2 of BASELINE.json (n=500, d=10, 10^6
synthetic candidates per GPU, squared-exponential kernel, ordinary kriging, EI), weak scaling.
A "step" is one pass of the hot path (K1 cross-correlation tiles -> K2c triangular product with
fused sum of squares -> K3 MSE + EI epilogue -> top-1) over one batch of candidates.
  value : candidates/s with the candidates already resident in HBM (device-pointer C ABI,
          CUDA events on the launching stream, L2 flushed between steps), whole job over N ranks;
  e2e   : the same through the public host API (pinned host arrays in, results out: H2D and D2H
          inside the timed region);
  roofline : the dominant kernel (k_trmm) timed live with CUDA events inside the library,
          algorithmic FP64 flops / measured DMMA peak;
  cpu_baseline : the numpy/scipy oracle PORT of the reference path timed on the host cores on a
          bounded sample (the reference itself is pure Python under /root/reference, which does not
          exist on the GPU box) -- a reported baseline only.
With no --workload the same JSON line carries `extra`: full sub-lines (value, ms_per_step,
roofline, e2e, cpu_baseline, clocks) for the other named shapes of the metric --
  extra.c3   ParallelBO MGFI x 8 at n=4096, d=20: 10^7 candidates SPLIT over the N ranks (strong
             scaling), per-step NCCL all-gather + device merge of the per-rank top-1, the model
             broadcast HBM -> HBM reported separately as broadcast_ms;
  extra.nll  the likelihood half of the metric: 64 theta-vectors of the n=4096, d=20 MLE split over
             the N ranks (strong scaling), plus 8 evaluations per GPU at n=1024 and n=2048 (the
             sizes a BO loop lives at);
  extra.c5   n=8192, d=40 Matern, simple kriging, 2^18 candidates per GPU;
  extra.c3_i8, extra.c5_i8, extra.c2_i8  the same shapes with K2c on the INT8 tensor pipe (option trmm_i8: Ozaki
             split into 8 slices, tcgen05.mma kind::i8, TMEM accumulators) -- same parity gates, reported
             against the FP64 DMMA roofline (frac > 1) and the INT8 one;
  extra.c2dx EI with its gradient (return_dx) at the C2 shape;
  extra.fit  wall time of GaussianProcess.fit (n=500, d=10, 10 lock-step restarts), with the reference's
             sequential restart loop on the oracle port as its CPU baseline;
  extra.per_point  single-point criterion(x) / criterion(x, return_dx=True) calls per second, the call
             pattern of the reference's own optimisers, GPU drop-in and oracle port.
N > 1 (torchrun): the model is fitted on rank 0 and broadcast through the library's own NCCL
communicator (mgp_bcast_model); candidates / restarts are sharded; per-rank top-1 pairs are
all-gathered and merged on the device every step (mgp_allgather_topk_dev).
`--impl reference` times the reference's CPU algorithm (oracle port, all host threads) on C2.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_PEAK_FALLBACK_TFLOPS = 37.0  # measured DMMA m8n8k4 peak on this pool (profiles/fp64_peak_r01.json)

WORKLOADS = {
    # BASELINE.json configs; m = candidates per GPU per step (weak) or in total (strong)
    "c1": dict(n=50, d=2, m=10_000, kind="squared_exponential", crit="EI", sets=1, nugget=1e-6, beta=None,
               desc="GP fit n=50 d=2, EI over 10^4 candidates (BASELINE configs[0], the reference's CPU-runnable case)"),
    "c2": dict(n=500, d=10, m=1_000_000, kind="squared_exponential", crit="EI", sets=1, nugget=1e-6, beta=None,
               desc="GP predict+EI d=10 n=500 over 10^6 synthetic candidates per GPU (BASELINE configs[1])"),
    "c2_i8": dict(n=500, d=10, m=1_000_000, kind="squared_exponential", crit="EI", sets=1, nugget=1e-6, beta=None, i8=8,
                  desc="GP predict+EI d=10 n=500 over 10^6 synthetic candidates per GPU (BASELINE configs[1]); "
                       "K2c on the INT8 tensor pipe (Ozaki split, 8 slices, tcgen05 + TMEM)"),
    "c2dx": dict(n=500, d=10, m=262_144, kind="squared_exponential", crit="EI", sets=1, nugget=1e-6, beta=None, dx=True,
                 desc="GP predict+EI with gradient (return_dx) d=10 n=500 over 2^18 candidates per GPU (BASELINE configs[1] shape)"),
    "c3": dict(n=4096, d=20, m=10_000_000, kind="squared_exponential", crit="MGFI", sets=8, nugget=1e-6, beta=None,
               strong=True,
               desc="ParallelBO MGFI n_point=8 d=20 n=4096, 10^7 candidates split over the GPUs (BASELINE configs[2])"),
    "c3_i8": dict(n=4096, d=20, m=10_000_000, kind="squared_exponential", crit="MGFI", sets=8, nugget=1e-6, beta=None,
                  strong=True, i8=8,
                  desc="ParallelBO MGFI n_point=8 d=20 n=4096, 10^7 candidates split over the GPUs (BASELINE configs[2]); "
                       "K2c on the INT8 tensor pipe (Ozaki split, 8 slices, tcgen05 + TMEM)"),
    "c5_i8": dict(n=8192, d=40, m=262_144, kind="matern", crit="EI", sets=1, nugget=0.0, beta=0.0, i8=8,
                  desc="BBOB-shaped surrogate d=40 n=8192 Matern-3/2 simple kriging, 2^18 candidates per GPU "
                       "(BASELINE configs[4]); K2c on the INT8 tensor pipe (Ozaki split, 8 slices, tcgen05 + TMEM)"),
    "c5": dict(n=8192, d=40, m=262_144, kind="matern", crit="EI", sets=1, nugget=0.0, beta=0.0,
               desc="BBOB-shaped surrogate d=40 n=8192 Matern-3/2 simple kriging, EI over CMA-ES-sized "
                    "populations, 2^18 candidates per GPU (BASELINE configs[4])"),
}
STRONG_BLOCKS = 64     # the strong-scaling population is 64 seeded blocks dealt to the ranks


def fp64_peak():
    for name in ("fp64_peak_r02.json", "fp64_peak_r01.json"):
        p = os.path.join(ROOT, "profiles", name)
        try:
            with open(p) as f:
                j = json.load(f)
            return float(j["dmma884_sustained_tflops"]), "measured DMMA m8n8k4 (tools/fp64_peak.cu, %s)" % name
        except Exception:
            continue
    return FP64_PEAK_FALLBACK_TFLOPS, "fallback (earlier DMMA measurement on this pool)"


def trmm_traffic(workload, chunk):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_trmm launch from the committed
    `ncu --set full` capture of this workload (profiles/r0*_trmm_traffic.json), or None."""
    for name in ("r02_trmm_traffic.json", "r01_trmm_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f).get(workload)
            if t and int(t["chunk"]) == int(chunk):
                return float(t["bytes_per_launch"])
        except Exception:
            pass
    return None


def synth_model(n, d, kind, seed=10):
    """SURVEY section 8d synthetic training set and hyper-parameters (no fitting here)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    y = (y - y.mean()) / y.std()
    theta = (0.12 / d) * rng.uniform(0.5, 1.5, d) * (4.0 if kind == "matern" else 1.0)
    return X, y, theta


def flops_per_candidate(n, d, dx=False):
    # F_pred = n^2 + 3nd + 6n; with the gradient F_pred_dx = 2n^2 + 11nd + 6n
    return 2 * n * n + 11 * n * d + 6 * n if dx else n * n + 3 * n * d + 6 * n


class ClockSampler(threading.Thread):
    """One `nvidia-smi --query-gpu=... -lms 50` process started BEFORE the timed region (the recipe
    of B200_PROFILING.md) and read line by line: no fork happens while steps are being launched."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows, self.proc = index, False, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        super().start()

    def run(self):
        if self.proc is None:
            return
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) >= 7 and not self.stop_flag:
                self.rows.append(f)
            if self.stop_flag:
                break

    def wait_first(self, timeout=2.0):
        """Block (before the timed region) until nvidia-smi has delivered its first sample."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
                self.proc.wait(timeout=5)
            except Exception:
                pass

    def summary(self):
        rows = []
        for r in self.rows:
            try:
                rows.append((float(r[0]), float(r[1]), float(r[2]), r[3:7]))
            except ValueError:
                pass
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(r[0] for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3][i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][1], "reasons": reasons,
                "samples": len(rows), "power_w_max": max(r[2] for r in rows)}


def model_par(wl, theta):
    """Hyper-parameter vector of a workload in the reference's layout (gpr.py:814,855)."""
    return np.r_[theta, 0.9] if wl["nugget"] > 0 else np.asarray(theta)


def oracle_state(wl, X, y, theta):
    from oracle import gp_oracle as O
    return O.fit_state(X, y, model_par(wl, theta), wl["kind"], wl["beta"] is None, wl["beta"] or 0.0,
                       "noisy" if wl["nugget"] > 0 else "noiseless", wl["nugget"])


def set_times(sets):
    """ParallelBO's sampled MGFI temperatures t_j = exp(log 2 + 0.5 z_j) (BayesOpt.py:65-68)."""
    return np.exp(np.log(2) + 0.5 * np.random.default_rng(22).standard_normal(sets))


def cpu_port_rate(wl, sample, seed=10, st=None):
    """Reference algorithm (oracle port) on the host cores: vectorised predict + acquisition in
    chunks of 2000 candidates (the best-case CPU path); with the gradient the
    reference's per-point path vectorised over the chunk."""
    from oracle import gp_oracle as O
    use_all_host_threads()
    n, d, kind, crit, sets = wl["n"], wl["d"], wl["kind"], wl["crit"], wl["sets"]
    X, y, theta = synth_model(n, d, kind, seed)
    if st is None:
        st = oracle_state(wl, X, y, theta)
    Xs = np.random.default_rng(seed + 1).uniform(-5, 5, (sample, d))
    plugin = float(y.min())
    ts = set_times(sets)
    t0 = time.perf_counter()
    for s in range(0, sample, 2000):
        xs = Xs[s:s + 2000]
        for t in (ts if sets > 1 else [ts[0]]):   # one posterior, `sets` epilogues
            O.acquisition(st, xs, crit, None if crit == "EI" else float(t), plugin, True,
                          return_dx=bool(wl.get("dx")))
    dt = time.perf_counter() - t0
    return sample / dt, dt


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms are entitled to every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_sample_size(wl):
    if wl.get("dx"):
        return 4000
    return min(wl["m"], 100_000) if wl["n"] <= 1000 else (1000 if wl["n"] <= 4096 else 400)


def run_reference(args, wl, rank):
    if rank != 0:
        return
    sample = min(wl["m"], 40_000) if wl["n"] <= 1000 else (1000 if wl["n"] <= 4096 else 400)
    if wl.get("dx"):
        sample = 4000
    for _ in range(args.warmup):
        cpu_port_rate(wl, min(sample, 4000))
    t_all = 0.0
    for _ in range(args.steps):
        r, dt = cpu_port_rate(wl, sample)
        t_all += dt
    value = sample * args.steps / t_all
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": "GP posterior+EI candidates/sec", "value": value, "unit": "candidates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "n": wl["n"], "d": wl["d"], "candidates_per_step": sample,
                   "criterion": wl["crit"], "n_sets": wl["sets"], "kernel": wl["kind"]},
        "cpu_baseline": {"value": value, "unit": "candidates/s", "cores": cores, "kind": "port",
                         "sample": "%d candidates per step, vectorised predict+%s in chunks of 2000 "
                                   "(oracle PORT of gpr.py:392-483 + InfillCriteria.py -- the reference itself is pure "
                                   "Python and absent on the GPU box; numpy/scipy BLAS threads=%d)"
                                   % (sample, wl["crit"], cores)},
        "e2e": {"value": value, "unit": "candidates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def nll_flops(n, d):
    return float(n) ** 3 + 3.5 * float(n) ** 2 * d  # SURVEY section 8d F_nll (potrf + inverse + contraction)


def nll_problem(n, d, B, seed=30):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    y = (y - y.mean()) / y.std()
    # log10-uniform around the well-conditioned region of SURVEY 8d (theta ~ 0.12/d), plus sigma2
    theta = (0.12 / d) * 10.0 ** np.random.default_rng(seed + 1).uniform(-0.5, 0.5, (B, d))
    P = np.c_[theta, np.random.default_rng(seed + 2).uniform(0.5, 1.0, B)]
    return X, y, P


def cpu_nll_rate(n, d, reps):
    from oracle import gp_oracle as O
    use_all_host_threads()
    X, y, P = nll_problem(n, d, max(1, reps))
    t0 = time.perf_counter()
    for b in range(reps):
        O.log_likelihood_concentrated(X, y, P[b], "squared_exponential", True, 0.0, "noisy", 1e-6, eval_grad=True)
    dt = time.perf_counter() - t0
    return reps / dt, dt


def run_reference_nll(args):
    n, d = args.nll_n, 20
    ns = min(n, 4096)
    r, dt = cpu_nll_rate(ns, d, 1)
    r *= (float(ns) / n) ** 3     # bounded sample at n <= 4096, O(n^3) cost scaled to n beyond
    line = {"impl": "reference", "metric": "GP NLL+grad evals/sec", "value": r, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "GP NLL+grad n=%d d=%d" % (n, d), "n": n, "d": d, "evals_per_step": 1},
            "cpu_baseline": {"value": r, "unit": "evals/s", "cores": blas_threads(), "kind": "port",
                             "sample": "1 evaluation of log_likelihood_concentrated(eval_grad=True) at n=%d, "
                                       "oracle port, rate scaled by (n_sample/n)^3" % ns},
            "e2e": {"value": r, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist
        import mipego_b200 as mb
        self.torch, self.dist, self.mb = torch, dist, mb
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: mipego_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        # an explicit (non-legacy) stream: the legacy default stream implicitly synchronises with other
        # blocking streams and would serialise the library's concurrent batch-group streams
        self.work = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.work)
        self.flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=self.dev)  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def build_model(ctx, wl):
    """Hyper-parameters fixed by the workload; the state is built by the library on rank 0
    (mgp_finalize_fit) and broadcast HBM -> HBM through the library's NCCL communicator, so that
    every rank holds the bit-identical model.  Returns (gp, X, y, theta, broadcast info)."""
    from mipego_b200 import dist as D
    mb = ctx.mb
    n, d, kind = wl["n"], wl["d"], wl["kind"]
    X, y, theta = synth_model(n, d, kind)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=wl["beta"]), corr=kind, theta0=theta,
                            thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d),
                            nugget=wl["nugget"] if wl["nugget"] > 0 else None, device=ctx.local)
    if ctx.rank == 0:
        gp.fit_fixed(X, y, model_par(wl, theta))
    else:
        gp._check_data(X, y)
    bc = None
    if ctx.world > 1:
        D.init_comm(gp)
        D.broadcast_fitted_state(gp, src=0)          # warm-up: NCCL channel set-up, receiver allocations
        ts = []
        for _ in range(3):
            ctx.barrier()
            t0 = time.perf_counter()
            D.broadcast_fitted_state(gp, src=0)      # synchronous on return (header checks + stream sync)
            ts.append(ctx.max_over_ranks(time.perf_counter() - t0))
        bc = {"broadcast_ms": 1e3 * min(ts), "broadcast_bytes": int(gp._handle().counter("bcast_bytes")),
              "transport": "mgp_bcast_model: ncclBroadcast of the device-resident model (L, L^-1, packed operands, "
                           "vectors), no host bounce, no re-factorisation on the receivers; max over ranks, best of 3"}
    return gp, X, y, theta, bc


def candidates(ctx, wl):
    """This rank's candidates.  Weak scaling: m per rank, seed 11 + rank.  Strong scaling: the global
    population is STRONG_BLOCKS seeded blocks of m / STRONG_BLOCKS candidates; rank r holds the blocks
    shard_range(STRONG_BLOCKS, r, world) -- the same population whatever the number of GPUs."""
    from mipego_b200.dist import shard_range
    d = wl["d"]
    if not wl.get("strong"):
        return np.random.default_rng(11 + ctx.rank).uniform(-5, 5, (wl["m"], d)), ctx.rank * wl["m"]
    per = wl["m"] // STRONG_BLOCKS
    lo, hi = shard_range(STRONG_BLOCKS, ctx.rank, ctx.world)
    parts = [np.random.default_rng([21, b]).uniform(-5, 5, (per, d)) for b in range(lo, hi)]
    return (np.concatenate(parts) if parts else np.empty((0, d))), lo * per


def bench_posterior(ctx, args, name, steps, warmup, cpu=True):
    torch, mb = ctx.torch, ctx.mb
    from mipego_b200 import dist as D
    wl = WORKLOADS[name]
    n, d, kind, dx = wl["n"], wl["d"], wl["kind"], bool(wl.get("dx"))
    gp, X, y, theta, bc = build_model(ctx, wl)
    h = gp._handle()
    if args.chunk:
        h.set_option("chunk", args.chunk)
    i8 = int(wl.get("i8", 0))
    if i8:
        h.set_option("trmm_i8", i8)
    plugin = float(y.min())
    ts = set_times(wl["sets"])
    crit = mb.EI(model=gp, minimize=True, plugin=plugin) if wl["crit"] == "EI" else \
        mb.MGFI(model=gp, minimize=True, plugin=plugin, t=float(ts[0]))
    params = None if wl["sets"] == 1 else ts
    Xs_np, offset = candidates(ctx, wl)
    m = Xs_np.shape[0]                              # this rank's candidates
    m_total = wl["m"] if wl.get("strong") else wl["m"] * ctx.world
    Xs_host = torch.from_numpy(Xs_np).pin_memory()
    Xs_np = Xs_host.numpy()
    Xs_dev = Xs_host.to(ctx.dev, non_blocking=True)
    topv = torch.empty((wl["sets"], 1), dtype=torch.float64, device=ctx.dev)
    topi = torch.empty((wl["sets"], 1), dtype=torch.int64, device=ctx.dev)
    gv, gi = torch.empty_like(topv), torch.empty_like(topi)
    val_dev = torch.empty((wl["sets"], m), dtype=torch.float64, device=ctx.dev) if dx else None
    dx_dev = torch.empty((wl["sets"], m, d), dtype=torch.float64, device=ctx.dev) if dx else None

    def step():
        if dx:
            crit.evaluate_device(Xs_dev, params=params, out=val_dev, return_dx=True, out_dx=dx_dev)
        elif ctx.world > 1:
            D.sharded_topk_device(crit, Xs_dev, offset, k=1, params=params, out=(gv, gi))
        else:
            crit.topk_device(Xs_dev, k=1, params=params, out=(topv, topi))

    for _ in range(warmup):
        step()
    ctx.barrier()

    def timed_pass():
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        ctx.barrier()
        for k in range(steps):
            ctx.flush.zero_()
            ev[k][0].record()
            step()
            ev[k][1].record()
        ctx.barrier()
        return sum(a.elapsed_time(b) for a, b in ev)

    launches0 = h.counter("launches")
    sampler = ClockSampler(ctx.local)
    sampler.start()
    sampler.wait_first()
    ms = timed_pass()
    sampler.stop()
    launches = h.counter("launches") - launches0
    # second, separate pass with CUDA events around every kernel launch (inside the library, on the
    # launching stream): per-kernel durations for the roofline; not used for `value`
    h.set_option("time_kernels", 1)
    steps_timed, steps = steps, (1 if wl.get("strong") else steps)     # a 10^7-candidate step is seconds long
    ms_k = timed_pass()
    steps_k, steps = steps, steps_timed
    trmm_ns, trmm_calls = h.counter("trmm_ns"), h.counter("trmm_calls")
    rgen_ns, epi_ns, topk_ns = h.counter("rgen_ns"), h.counter("epilogue_ns"), h.counter("topk_ns")
    h.set_option("time_kernels", 0)
    ms_max = ctx.max_over_ranks(ms)
    value = m_total * steps / (ms_max * 1e-3)
    top1 = None
    if not dx:
        tv_, ti_ = (gv, gi) if ctx.world > 1 else (topv, topi + offset)
        top1 = {"value": float(tv_[0, 0].item()), "global_index": int(ti_[0, 0].item())}

    # ---- e2e: public host API, pinned host candidates in, results out, every step -------------------
    if dx:
        vals = torch.empty((wl["sets"], m), dtype=torch.float64).pin_memory().numpy()
        grads = torch.empty((wl["sets"], m, d), dtype=torch.float64).pin_memory().numpy()
        e2e_call = lambda: crit.evaluate(Xs_np, params=params, return_dx=True, out=vals, out_dx=grads)  # noqa: E731
        api = "%s.evaluate(pinned host ndarray, return_dx=True, out=, out_dx= pinned) (mgp_acq)" % wl["crit"]
        d2h = m * wl["sets"] * 8 * (1 + d)
    elif wl.get("strong"):
        e2e_call = lambda: crit.topk(Xs_np, k=1, params=params)  # noqa: E731
        api = "%s.topk(pinned host shard, k=1, params=8 t-values) (mgp_acq_topk) + mgp_allgather_topk" % wl["crit"]
        d2h = wl["sets"] * 16
        vals = None
    else:
        vals = torch.empty((wl["sets"], m), dtype=torch.float64).pin_memory().numpy()   # pinned result buffer
        e2e_call = lambda: crit.evaluate(Xs_np, params=params, out=vals)  # noqa: E731
        api = "%s.evaluate(pinned host ndarray, out=pinned host ndarray) (mgp_acq)" % wl["crit"]
        d2h = m * wl["sets"] * 8
    crit.evaluate(Xs_np[:min(m, 200000)], params=params, return_dx=dx)  # warm the staging buffers
    ctx.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        r = e2e_call()
        if wl.get("strong") and ctx.world > 1:
            D.allgather_topk(r[0], r[1], 1, offset, gp=gp)
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e_value = m_total * steps / e2e_s

    line = None
    if ctx.rank == 0:
        err = None
        st = None
        if cpu:
            # cpu_baseline leg (the only place the oracle is touched): besides being timed below, the
            # CPU port checks a sample of what the timed GPU path produced (not timed)
            from oracle import gp_oracle as O
            st = oracle_state(wl, X, y, theta)
            sub = Xs_np[:2000 if n <= 4096 else 500]
            if vals is None:
                got = crit.evaluate(sub, params=params)[0]
            else:
                got = vals[0][:len(sub)]
            rv = O.acquisition(st, sub, wl["crit"], None if wl["crit"] == "EI" else float(ts[0]), plugin, True)
            err = float(np.max(np.abs(got - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max())))
        peak, peak_src = fp64_peak()
        chunk = h.counter("chunk")
        per_launch_c = m * steps_k / max(1, trmm_calls)
        fl_c = (2.0 if dx else 1.0) * float(n) * n     # the n^2 term of F_pred (2 n^2 of F_pred_dx: W = R^-1 r)
        flops_per_launch = per_launch_c * fl_c
        achieved = flops_per_launch / (trmm_ns / max(1, trmm_calls) * 1e-9) / 1e12 if trmm_ns else None
        line = {
            "metric": "GP posterior+EI candidates/sec", "value": value, "unit": "candidates/s",
            "n_gpus": ctx.world, "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
            "higher_is_better": True, "scaling": "strong" if wl.get("strong") else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "n": n, "d": d, "candidates_per_step_total": m_total,
                       "candidates_per_step_rank0": m, "criterion": wl["crit"], "n_sets": wl["sets"],
                       "return_dx": dx, "kernel": kind, "chunk": chunk,
                       "l2": "256 MB buffer written between timed steps (L2 flush)",
                       "flops_per_candidate": flops_per_candidate(n, d, dx),
                       "sample_max_rel_err_vs_oracle": err, "top1": top1,
                       "collective": None if ctx.world == 1 or dx else
                       "per step: mgp_allgather_topk_dev (ncclAllGather of n_sets x 16 B per rank + device merge)"},
            "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": int(m * d * 8),
                    "d2h_bytes_per_step": int(d2h), "api": api},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor",
                         "kernel": "k_trmm<true> (W = R^-1 r)" if dx else "k_trmm<false> (T = L^-1 r, fused sum of squares)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": trmm_traffic(name, chunk),
                         "algorithmic_bytes_8d": 8.0 * (d + wl["sets"]) * per_launch_c,
                         "peak_source": "FP64 " + peak_src + "; MEASURED_PEAKS.json has no FP64 entry",
                         "flops_per_launch": flops_per_launch, "launches_timed": int(trmm_calls),
                         "avg_launch_ms": trmm_ns / max(1, trmm_calls) * 1e-6,
                         "share_of_step": {"k_rgen": rgen_ns * 1e-6 / ms_k, "k_trmm": trmm_ns * 1e-6 / ms_k,
                                           "k_grad" if dx else "k_epilogue": epi_ns * 1e-6 / ms_k,
                                           "k_topk": topk_ns * 1e-6 / ms_k},
                         "whole_step_frac_of_peak": value / ctx.world * flops_per_candidate(n, d, dx) / (peak * 1e12)},
            "clocks": sampler.summary(),
        }
        if i8:
            # the same n^2 FP64-equivalent flops per candidate, evaluated as 36 (8 slices) or 34 (7) exact int8
            # slice products: reported against BOTH rooflines (frac > 1 = beyond the FP64 DMMA peak)
            pairs = 36 if i8 == 8 else 34
            # per candidate and slice product: n^2 / 2 multiply-adds over the triangle = n^2 int8 ops
            int8_ops = pairs * float(n) * n * (value / ctx.world)             # per GPU per second
            fused = bool(h.counter("trmm_i8_fuse")) and i8 == 8 and kind != "absolute_exponential" and d <= 40
            line["roofline"]["kernel"] = ("%sk_trmm_i8<%d> (T = L^-1 r as %d exact int8 slice products, tcgen05.mma kind::i8, "
                                          "TMEM accumulators, FP64 recombination%s)"
                                          % ("" if fused else "k_oz_slice_r + ", i8, pairs,
                                             "; the int8 slices of r are written by K1 itself" if fused else ""))
            line["roofline"]["int8"] = {"achieved": int8_ops / 1e12, "peak": 4500.0, "unit": "TOP/s",
                                        "frac": int8_ops / 4.5e15,
                                        "peak_source": "nominal dense INT8 of B200 (no measured INT8 entry in MEASURED_PEAKS.json)",
                                        "note": "ops = 2 x multiply-adds over the lower triangle (n^2 / 2 per slice product and candidate)"}
            line["roofline"]["note"] = "frac is FP64-equivalent n^2 flop per candidate over the measured DMMA peak"
        if bc:
            line["broadcast"] = bc
        if cpu:
            sample = cpu_sample_size(wl)
            r, dt = cpu_port_rate(wl, sample, st=st)
            line["cpu_baseline"] = {"value": r, "unit": "candidates/s", "cores": blas_threads(), "kind": "port",
                                    "sample": "%d candidates, vectorised predict+%s%s in chunks of 2000 (oracle PORT of "
                                              "the reference's numpy/scipy path, %.1f s)"
                                              % (sample, wl["crit"], " with gradient" if dx else "", dt)}
    del gp, crit
    return line


def bench_nll(ctx, args, n, B_total, steps, warmup, strong, cpu=True, groups=0, cta_cap=0):
    """NLL + gradient evaluations/s at (n, d=20).  strong: B_total theta-vectors split over the ranks
    (BASELINE configs[3]: 64 restarts); otherwise B_total per rank.  Every step all-gathers the
    (llf | grad) rows so that each rank -- i.e. each L-BFGS-B driver -- sees all results."""
    torch, mb = ctx.torch, ctx.mb
    from mipego_b200.dist import shard_range, init_comm
    d = 20
    if strong:
        lo, hi = shard_range(B_total, ctx.rank, ctx.world)
        B_all = B_total
    else:
        lo, hi = ctx.rank * B_total, (ctx.rank + 1) * B_total
        B_all = B_total * ctx.world
    B = hi - lo
    X, y, P_all = nll_problem(n, d, B_all)
    P = np.ascontiguousarray(P_all[lo:hi])
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential",
                            theta0=P_all[0, :d], thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6,
                            device=ctx.local)
    gp._check_data(X, y)
    h = gp._handle()
    if groups:
        h.set_option("nll_groups", groups)
    if cta_cap:
        h.set_option("gemm_cta_cap", cta_cap)
    if ctx.world > 1:
        init_comm(gp)
    P_dev = torch.from_numpy(P).to(ctx.dev)
    width = d + 2                                  # llf | grad (d + 1)
    Bmax = -(-B_all // ctx.world)
    rows = torch.zeros((Bmax, width), dtype=torch.float64, device=ctx.dev)
    gathered = torch.empty((ctx.world, Bmax, width), dtype=torch.float64, device=ctx.dev)
    out = [None]

    def step():
        out[0] = gp.log_likelihood_batch_device(P_dev)
        if ctx.world > 1:
            rows[:B, 0] = out[0][0]
            rows[:B, 1:] = out[0][1]
            st = torch.cuda.current_stream(ctx.dev).cuda_stream
            h.check(h.lib.mgp_allgather_f64_dev(h.h, rows.data_ptr(), Bmax * width, gathered.data_ptr(), st))

    sampler = ClockSampler(ctx.local)
    sampler.start()
    sampler.wait_first()
    for _ in range(warmup):
        step()
    ctx.barrier()
    launches0 = h.counter("launches")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    for k in range(steps):
        ev[k][0].record()
        step()
        ev[k][1].record()
    ctx.barrier()
    sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    if os.environ.get("MGP_BENCH_DEBUG"):
        print("per-step ms:", [round(a.elapsed_time(b), 2) for a, b in ev], file=sys.stderr)
    launches = h.counter("launches") - launches0
    ms_max = ctx.max_over_ranks(ms)
    value = B_all * steps / (ms_max * 1e-3)
    # e2e: host API (parameters in, llf + grad + status out)
    ctx.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        llf, grad, status = gp.log_likelihood_batch(P)
    ctx.barrier()
    e2e = B_all * steps / ctx.max_over_ranks(time.perf_counter() - t0)
    line = None
    if ctx.rank == 0:
        peak, peak_src = fp64_peak()
        achieved = value / ctx.world * nll_flops(n, d) / 1e12
        desc = ("GP NLL+grad n=%d d=%d, %d theta-vectors %s (BASELINE configs[3]: hyper-parameter MLE, 64 restarts)"
                % (n, d, B_total, "split over the GPUs" if strong else "per GPU per step"))
        line = {"metric": "GP NLL+grad evals/sec", "value": value, "unit": "evals/s", "n_gpus": ctx.world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
                "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "n": n, "d": d, "evals_per_step_total": B_all, "evals_per_step_rank0": B,
                           "mode": "noisy nugget 1e-6", "status": [int(v) for v in status][:8], "llf0": float(llf[0]),
                           "l2": "working set 4 x %d MB per evaluation exceeds L2" % (n * n * 8 // 2 ** 20),
                           "collective": None if ctx.world == 1 else
                           "per step: mgp_allgather_f64_dev (ncclAllGather of the llf | grad rows)"},
                "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": int(P.nbytes),
                        "d2h_bytes_per_step": int(llf.nbytes + grad.nbytes + status.nbytes),
                        "api": "GaussianProcess.log_likelihood_batch(host ndarray) (mgp_nll_batch)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "kernel": "whole step (k_dgemm family: potrf panels/updates, trtri, M^T M)",
                             "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                             "traffic": None, "peak_source": "FP64 " + peak_src,
                             "flops_per_eval_Fnll": nll_flops(n, d)},
                "clocks": sampler.summary()}
        if cpu:
            # bounded sample: one evaluation at min(n, 4096) (17 s at n = 4096 on 16 cores); beyond that the
            # O(n^3) cost is scaled to n
            ns = min(n, 4096)
            r, dt = cpu_nll_rate(ns, d, 1)
            scale = (float(ns) / n) ** 3
            line["cpu_baseline"] = {"value": r * scale, "unit": "evals/s", "cores": blas_threads(), "kind": "port",
                                    "sample": "1 evaluation of log_likelihood_concentrated(eval_grad=True) at n=%d "
                                              "(oracle PORT, %.1f s)%s" % (ns, dt, "" if ns == n else
                                                                           ", rate scaled by (%d/%d)^3 to n=%d" % (ns, n, n))}
    del gp
    return line


def bench_fit(ctx, cpu=False):
    """Wall time of GaussianProcess.fit in the reference's loop shape (base.py:516: one fit per BO
    iteration): n=500, d=10, Matern, 10 lock-step L-BFGS-B restarts (gpr.py:964-1101).  Rank 0 only."""
    mb = ctx.mb
    out = None
    if ctx.rank == 0:
        n, d, R = 500, 10, 10
        rng = np.random.default_rng(0)
        X = rng.uniform(-5, 5, (n, d))
        y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
        y = (y - y.mean()) / y.std()
        np.random.seed(1)
        gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=[0.05] * d,
                                thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=1e-6, random_start=R, device=ctx.local)
        gp.fit(X[:64], y[:64])          # warm the library (allocations, first launches)
        ts, evals = [], 0
        for rep in range(3):
            np.random.seed(1)
            gp2 = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=[0.05] * d,
                                     thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=1e-6, random_start=R,
                                     device=ctx.local)
            gp2._h = gp._h
            t0 = time.perf_counter()
            gp2.fit(X, y)
            ts.append(time.perf_counter() - t0)
            evals = gp2.eval_count
        out = {"metric": "GaussianProcess.fit wall time", "value": 1e3 * min(ts), "unit": "ms", "higher_is_better": False,
               "config": {"n": n, "d": d, "kernel": "matern", "restarts": R, "likelihood_evaluations": int(evals),
                          "restart_mode": "batched (lock-step L-BFGS-B, one batched likelihood call per step)"},
               "ms_per_evaluation": 1e3 * min(ts) / max(1, evals), "runs_ms": [1e3 * t for t in ts]}
        if cpu:
            out["cpu_baseline"] = cpu_fit_leg(mb, ctx.local, X, y, d, R)
    ctx.barrier()
    return out


def cpu_fit_leg(mb, device, X, y, d, R, budget_s=20.0):
    """the reference's fit() loop on the host cores -- sequential L-BFGS-B restarts with a
    shared evaluation budget and the wait_iter early stop (gpr.py:1030-1064), every objective call ONE
    likelihood + gradient evaluation of the oracle port -- bounded to `budget_s` seconds of CPU work
    (a full run is 200 n_par = 2200 evaluations at most).  The restart points, bounds and parametrisation
    are the ones GaussianProcess.fit draws (injected through its `_solve` hook)."""
    from oracle import gp_oracle as O
    from scipy.optimize import fmin_l_bfgs_b
    use_all_host_threads()
    stat = {"evals": 0, "restarts": 0, "timeout": False, "t": 0.0}

    class _Timeout(Exception):
        pass

    def solve(batch_fn, starts, log10bounds, eval_budget, wait_iter):
        t0 = time.perf_counter()

        def func(p):
            if time.perf_counter() - t0 > budget_s:
                raise _Timeout()
            stat["evals"] += 1
            llf, g = O.log_likelihood_concentrated(X, y, 10.0 ** np.asarray(p), "matern", True, 0.0, "noisy", 1e-6,
                                                   eval_grad=True)
            return -float(llf), -np.asarray(g, dtype=np.float64).ravel()
        best_p, best_f, wait, budget = np.array(starts[0]), np.inf, 0, eval_budget
        try:
            for it in range(len(starts)):
                p_, f_, info = fmin_l_bfgs_b(func, starts[it], bounds=log10bounds, maxfun=budget)
                stat["restarts"] += 1
                if it == 0 or f_ <= best_f:
                    best_p, best_f, wait = p_, f_, 0
                else:
                    wait += 1
                budget -= info["funcalls"]
                if budget <= 0 or wait >= wait_iter:
                    break
        except _Timeout:
            stat["timeout"] = True
        stat["t"] = time.perf_counter() - t0
        return best_p, best_f, stat["evals"], []

    np.random.seed(1)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=[0.05] * d,
                            thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=1e-6, random_start=R, device=device)
    gp.fit(X, y, _solve=solve)
    return {"value": 1e3 * stat["t"], "unit": "ms", "cores": blas_threads(), "kind": "port",
            "evaluations": stat["evals"], "restarts_completed": stat["restarts"], "stopped_by_time_bound": stat["timeout"],
            "ms_per_evaluation": 1e3 * stat["t"] / max(1, stat["evals"]),
            "sample": "the reference's sequential restart loop (gpr.py:1030-1064) on the oracle PORT of the likelihood, "
                      "same data / bounds / restart points, bounded to %.0f s (BASELINE.md B4)" % budget_s}


def bench_per_point(ctx, cpu):
    """the acquisition exactly as the reference's optimisers call it -- ONE point per call
    (optimizer/utils.py:38-42) -- over 10^3 seeded candidates at the C2 model: criterion(x) and
    criterion(x, return_dx=True) through the drop-in classes on the GPU, and the same calls through the
    oracle port on the host cores.  Rank 0 only."""
    out = None
    if ctx.rank == 0:
        mb = ctx.mb
        wl = WORKLOADS["c2"]
        n, d, kind = wl["n"], wl["d"], wl["kind"]
        X, y, theta = synth_model(n, d, kind)
        gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=wl["beta"]), corr=kind, theta0=theta,
                                thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=wl["nugget"], device=ctx.local)
        gp.fit_fixed(X, y, model_par(wl, theta))
        ei = mb.EI(model=gp, minimize=True, plugin=float(y.min()))
        Xs = np.random.default_rng(11).uniform(-5, 5, (1000, d))
        for x in Xs[:50]:
            ei(x)
            ei(x, return_dx=True)
        t0 = time.perf_counter()
        for x in Xs:
            ei(x)
        t1 = time.perf_counter()
        for x in Xs:
            ei(x, return_dx=True)
        t2 = time.perf_counter()
        out = {"metric": "per-point acquisition calls/sec (the reference optimisers' call pattern)", "unit": "calls/s",
               "value": 1000 / (t1 - t0), "value_with_dx": 1000 / (t2 - t1), "higher_is_better": True,
               "us_per_call": 1e3 * (t1 - t0), "us_per_call_with_dx": 1e3 * (t2 - t1),
               "config": {"n": n, "d": d, "kernel": kind, "criterion": "EI", "points": 1000}}
        if cpu:
            from oracle import gp_oracle as O
            use_all_host_threads()
            st = oracle_state(wl, X, y, theta)
            plugin = float(y.min())
            m = 300
            t0 = time.perf_counter()
            for x in Xs[:m]:
                O.acquisition(st, x[None, :], "EI", None, plugin, True)
            t1 = time.perf_counter()
            for x in Xs[:m]:
                O.acquisition(st, x[None, :], "EI", None, plugin, True, return_dx=True)
            t2 = time.perf_counter()
            out["cpu_baseline"] = {"value": m / (t1 - t0), "value_with_dx": m / (t2 - t1), "unit": "calls/s",
                                   "cores": blas_threads(), "kind": "port",
                                   "sample": "%d single-point EI calls (and %d with return_dx) through the oracle PORT "
                                             "of gpr.py:392-617 + InfillCriteria.py:113-148 (BASELINE.md B1)" % (m, m)}
    ctx.barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS) + ["nll"])
    ap.add_argument("--no-extra", action="store_true", help="headline line only (no extra.* sub-benchmarks)")
    ap.add_argument("--nll-n", type=int, default=4096)
    ap.add_argument("--nll-batch", type=int, default=8, help="theta-vectors per GPU per step (--workload nll)")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--nll-groups", type=int, default=0)
    ap.add_argument("--gemm-cta-cap", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if args.workload == "nll":
            if rank == 0:
                run_reference_nll(args)
        else:
            run_reference(args, WORKLOADS[args.workload or "c2"], rank)
        return

    ctx = Ctx()
    cpu = not args.no_cpu_baseline
    if args.workload == "nll":
        line = bench_nll(ctx, args, args.nll_n, args.nll_batch, args.steps, args.warmup, strong=False, cpu=cpu,
                         groups=args.nll_groups, cta_cap=args.gemm_cta_cap)
    elif args.workload is not None:
        wl = WORKLOADS[args.workload]
        steps = args.steps if not wl.get("strong") else min(args.steps, 3)
        line = bench_posterior(ctx, args, args.workload, steps, args.warmup if not wl.get("strong") else 3, cpu=cpu)
    else:
        line = bench_posterior(ctx, args, "c2", args.steps, args.warmup, cpu=cpu)
        if not args.no_extra:
            # the other named shapes of the metric; CPU legs only at N = 1 (bounded: about 60 s in total)
            cpu_x = cpu and ctx.world == 1
            extra = {}
            t0 = time.time()
            extra["c3"] = bench_posterior(ctx, args, "c3", 2, 3, cpu=cpu_x)
            extra["nll"] = bench_nll(ctx, args, 4096, 64, 3, 3, strong=True, cpu=cpu_x)
            # BO-loop sizes: 8 theta-vectors per GPU per step (a lock-step MLE with few restarts: bound by the
            # serial chain of the factorisation) and 64 (run_sequential.py uses random_start = 10 d)
            extra["nll_n2048"] = bench_nll(ctx, args, 2048, 8, 10, 3, strong=False, cpu=False)
            extra["nll_n2048_b64"] = bench_nll(ctx, args, 2048, 64, 5, 3, strong=False, cpu=False)
            extra["nll_n1024"] = bench_nll(ctx, args, 1024, 8, 20, 3, strong=False, cpu=False)
            extra["nll_n1024_b64"] = bench_nll(ctx, args, 1024, 64, 10, 3, strong=False, cpu=False)
            extra["c5"] = bench_posterior(ctx, args, "c5", 3, 3, cpu=cpu_x)
            # the two trmm-bound shapes again with K2c on the INT8 tensor pipe (opt-in path, same parity gates)
            extra["c3_i8"] = bench_posterior(ctx, args, "c3_i8", 2, 3, cpu=cpu_x)
            extra["c5_i8"] = bench_posterior(ctx, args, "c5_i8", 3, 3, cpu=False)
            extra["c2_i8"] = bench_posterior(ctx, args, "c2_i8", 10, 3, cpu=False)
            extra["c2dx"] = bench_posterior(ctx, args, "c2dx", 5, 3, cpu=cpu_x)
            extra["fit"] = bench_fit(ctx, cpu=cpu_x)
            extra["per_point"] = bench_per_point(ctx, cpu_x)
            if ctx.rank == 0:
                extra["seconds"] = time.time() - t0
                line["extra"] = extra
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
