#!/usr/bin/env python
"""bench.py -- GP posterior + EI candidates/s on the BASELINE.json configuration C2
(n=500, d=10, 10^6 synthetic candidates per GPU, squared-exponential kernel, ordinary kriging).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c1|c2|c3|c5|nll]

A "step" is one pass of the hot path (K1 cross-correlation tiles -> K2c triangular product with
fused sum of squares -> K3 MSE + EI epilogue -> top-1) over one batch of candidates.
  value : candidates/s with the candidates already resident in HBM (device-pointer C ABI,
          CUDA events on the launching stream, L2 flushed between steps), whole job over N ranks;
  e2e   : the same through the public host API (EI.evaluate on a pinned host array: H2D of the
          candidates and D2H of the EI values inside the timed region);
  roofline : the dominant kernel (k_trmm) timed live with CUDA events inside the library,
          algorithmic FP64 flops / measured DMMA peak;
  cpu_baseline : the numpy/scipy oracle port of the reference path timed on the host cores on a
          bounded sample (reported baseline only).
N > 1 (torchrun): the model is fitted state on rank 0 and broadcast (NCCL), candidates are
sharded (weak scaling: 10^6 per rank), per-rank top-1 is all-gathered every step.
`--impl reference` times the reference's CPU algorithm (oracle port; the reference is pure
Python and /root/reference does not exist on the GPU box) on the same config.
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
    # BASELINE.json configs; m = candidates per GPU per step
    "c1": dict(n=50, d=2, m=10_000, kind="squared_exponential", crit="EI", sets=1, nugget=1e-6, beta=None,
               desc="GP fit n=50 d=2, EI over 10^4 candidates (BASELINE configs[0], the reference's CPU-runnable case)"),
    "c2": dict(n=500, d=10, m=1_000_000, kind="squared_exponential", crit="EI", sets=1, nugget=1e-6, beta=None,
               desc="GP predict+EI d=10 n=500 over 10^6 synthetic candidates per GPU (BASELINE configs[1])"),
    "c3": dict(n=4096, d=20, m=1_250_000, kind="squared_exponential", crit="MGFI", sets=8, nugget=1e-6, beta=None,
               desc="ParallelBO MGFI n_point=8 d=20 n=4096, 1.25x10^6 candidates per GPU (BASELINE configs[2] shard)"),
    "c5": dict(n=8192, d=40, m=262_144, kind="matern", crit="EI", sets=1, nugget=0.0, beta=0.0,
               desc="BBOB-shaped surrogate d=40 n=8192 Matern-3/2 simple kriging, EI over CMA-ES-sized "
                    "populations, 2^18 candidates per GPU (BASELINE configs[4])"),
}


def fp64_peak():
    p = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    try:
        with open(p) as f:
            j = json.load(f)
        return float(j["dmma884_sustained_tflops"]), "measured DMMA m8n8k4 (tools/fp64_peak.cu, %s)" % os.path.basename(p)
    except Exception:
        return FP64_PEAK_FALLBACK_TFLOPS, "fallback (earlier DMMA measurement on this pool)"


def trmm_traffic(workload, chunk):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_trmm launch from the committed
    `ncu --set full` capture of this workload (profiles/r01_trmm_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_trmm_traffic.json")) as f:
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


def flops_per_candidate(n, d):
    return n * n + 3 * n * d + 6 * n  # SURVEY section 8d F_pred


class ClockSampler(threading.Thread):
    """One `nvidia-smi --query-gpu=... -lms 100` process started BEFORE the timed region (the recipe
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


def cpu_port_rate(wl, sample, seed=10):
    """Reference algorithm (oracle port) on the host cores: vectorised predict + acquisition in
    chunks of 2000 candidates (the best-case CPU path)."""
    from oracle import gp_oracle as O
    use_all_host_threads()
    n, d, kind, crit, sets = wl["n"], wl["d"], wl["kind"], wl["crit"], wl["sets"]
    X, y, theta = synth_model(n, d, kind, seed)
    st = oracle_state(wl, X, y, theta)
    Xs = np.random.default_rng(seed + 1).uniform(-5, 5, (sample, d))
    plugin = float(y.min())
    ts = np.exp(np.log(2) + 0.5 * np.random.default_rng(22).standard_normal(sets))
    t0 = time.perf_counter()
    for s in range(0, sample, 2000):
        xs = Xs[s:s + 2000]
        if sets == 1:
            O.acquisition(st, xs, crit, None if crit == "EI" else float(ts[0]), plugin, True)
        else:
            # one posterior, `sets` epilogues
            for t in ts:
                O.acquisition(st, xs, crit, float(t), plugin, True)
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


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    sample = min(wl["m"], 40_000) if wl["n"] <= 1000 else (1000 if wl["n"] <= 4096 else 400)
    rates = []
    for _ in range(args.warmup):
        cpu_port_rate(wl, min(sample, 4000))
    t_all = 0.0
    for _ in range(args.steps):
        r, dt = cpu_port_rate(wl, sample)
        rates.append(r)
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
                                   "(oracle port of gpr.py:392-483 + InfillCriteria.py; numpy/scipy BLAS threads=%d)"
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


def run_nll(args, rank, world, local):
    """BASELINE configs[3]: GP hyper-parameter MLE n=4096 d=20, 64 restarts -> 8 NLL+grad per GPU."""
    n, d, B = args.nll_n, 20, args.nll_batch
    desc = "GP NLL+grad n=%d d=%d, %d theta-vectors per GPU per step (BASELINE configs[3]: 64 over 8 GPUs)" % (n, d, B)
    if args.impl == "reference":
        if rank != 0:
            return
        reps = 1
        ns = min(n, 4096)
        r, dt = cpu_nll_rate(ns, d, reps)
        r *= (float(ns) / n) ** 3     # bounded sample at n <= 4096, O(n^3) cost scaled to n beyond
        line = {"impl": "reference", "metric": "GP NLL+grad evals/sec", "value": r, "unit": "evals/s",
                "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "n": n, "d": d, "evals_per_step": reps},
                "cpu_baseline": {"value": r, "unit": "evals/s", "cores": blas_threads(), "kind": "port",
                                 "sample": "%d evaluation(s) of log_likelihood_concentrated(eval_grad=True) at n=%d, "
                                           "oracle port, rate scaled by (n_sample/n)^3" % (reps, ns)},
                "e2e": {"value": r, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    import torch
    import torch.distributed as dist
    import mipego_b200 as mb
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    X, y, P_all = nll_problem(n, d, B * world)
    P = P_all[rank * B:(rank + 1) * B]
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential",
                            theta0=P[0, :d], thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6,
                            device=local)
    gp._check_data(X, y)
    h = gp._handle()
    if args.nll_groups:
        h.set_option("nll_groups", args.nll_groups)
    if args.gemm_cta_cap:
        h.set_option("gemm_cta_cap", args.gemm_cta_cap)
    P_dev = torch.from_numpy(P).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    # an explicit (non-legacy) stream: the legacy default stream implicitly synchronises with other
    # blocking streams and would serialise the library's concurrent batch-group streams
    work = torch.cuda.Stream(device=dev)
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    with torch.cuda.stream(work):
        for _ in range(args.warmup):
            out = gp.log_likelihood_batch_device(P_dev)
        barrier()
        launches0 = h.counter("launches")
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for k in range(args.steps):
            ev[k][0].record()
            out = gp.log_likelihood_batch_device(P_dev)
            if world > 1:
                g = [torch.empty_like(out[0]) for _ in range(world)]
                dist.all_gather(g, out[0])
            ev[k][1].record()
        barrier()
    sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    if os.environ.get("MGP_BENCH_DEBUG"):
        print("per-step ms:", [round(a.elapsed_time(b), 2) for a, b in ev], file=sys.stderr)
    launches = h.counter("launches") - launches0
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    value = world * B * args.steps / (float(t_ms.item()) * 1e-3)
    # e2e: host API (parameters in, llf + grad + status out)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        llf, grad, status = gp.log_likelihood_batch(P)
    barrier()
    t_e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e = world * B * args.steps / float(t_e.item())
    if rank == 0:
        peak, peak_src = fp64_peak()
        achieved = value / world * nll_flops(n, d) / 1e12
        line = {"metric": "GP NLL+grad evals/sec", "value": value, "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t_ms.item()) / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "n": n, "d": d, "evals_per_step_per_gpu": B, "mode": "noisy nugget 1e-6",
                           "status": [int(v) for v in status], "llf0": float(llf[0]),
                           "l2": "working set 4 x %d MB per evaluation exceeds L2" % (n * n * 8 // 2 ** 20)},
                "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": int(P.nbytes),
                        "d2h_bytes_per_step": int(llf.nbytes + grad.nbytes + status.nbytes),
                        "api": "GaussianProcess.log_likelihood_batch(host ndarray) (mgp_nll_batch)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor", "kernel": "whole step (k_dgemm family: potrf panels/updates, trtri, M^T M)",
                             "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                             "traffic": None, "peak_source": "FP64 " + peak_src,
                             "flops_per_eval_Fnll": nll_flops(n, d)},
                "clocks": sampler.summary()}
        if not args.no_cpu_baseline:
            # bounded sample: one evaluation at min(n, 4096) (17 s at n = 4096 on 16 cores); beyond that the
            # O(n^3) cost is scaled to n
            ns = min(n, 4096)
            r, dt = cpu_nll_rate(ns, d, 1)
            scale = (float(ns) / n) ** 3
            line["cpu_baseline"] = {"value": r * scale, "unit": "evals/s", "cores": blas_threads(), "kind": "port",
                                    "sample": "1 evaluation of log_likelihood_concentrated(eval_grad=True) at n=%d "
                                              "(oracle port, %.1f s)%s" % (ns, dt, "" if ns == n else
                                                                           ", rate scaled by (%d/%d)^3 to n=%d" % (ns, n, n))}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["nll"])
    ap.add_argument("--nll-n", type=int, default=4096)
    ap.add_argument("--nll-batch", type=int, default=8, help="theta-vectors per GPU per step (configs[3]: 8)")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--nll-groups", type=int, default=0)
    ap.add_argument("--gemm-cta-cap", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "nll":
        run_nll(args, rank, world, local)
        return
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    import mipego_b200 as mb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mipego_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, d, m, kind = wl["n"], wl["d"], wl["m"], wl["kind"]
    # ---- model: hyper-parameters fixed by the workload, state built by the library itself on rank 0
    # (mgp_finalize_fit) and broadcast so that every rank holds the bit-identical model ------------
    from mipego_b200.dist import broadcast_fitted_state
    X, y, theta = synth_model(n, d, kind)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=wl["beta"]), corr=kind, theta0=theta,
                            thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d),
                            nugget=wl["nugget"] if wl["nugget"] > 0 else None, device=local)
    if rank == 0:
        gp.fit_fixed(X, y, model_par(wl, theta))
    else:
        gp._check_data(X, y)
    if world > 1:
        broadcast_fitted_state(gp, src=0, device=dev)
    h = gp._handle()
    if args.chunk:
        h.set_option("chunk", args.chunk)
    plugin = float(y.min())
    ts = np.exp(np.log(2) + 0.5 * np.random.default_rng(22).standard_normal(wl["sets"]))
    crit = mb.EI(model=gp, minimize=True, plugin=plugin) if wl["crit"] == "EI" else \
        mb.MGFI(model=gp, minimize=True, plugin=plugin, t=float(ts[0]))
    params = None if wl["sets"] == 1 else ts

    # ---- candidates: this rank's shard, pinned host copy + resident device copy -----------------
    rng = np.random.default_rng(11 + rank)
    Xs_host = torch.from_numpy(rng.uniform(-5, 5, (m, d))).pin_memory()
    Xs_dev = Xs_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2
    topv = torch.empty((wl["sets"], 1), dtype=torch.float64, device=dev)
    topi = torch.empty((wl["sets"], 1), dtype=torch.int64, device=dev)
    gathered = [torch.empty((wl["sets"], 2), dtype=torch.float64, device=dev) for _ in range(world)]

    def step():
        crit.topk_device(Xs_dev, k=1, params=params, out=(topv, topi))
        if world > 1:
            mine = torch.cat([topv, (topi + rank * m).to(torch.float64)], dim=1)
            dist.all_gather(gathered, mine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    work = torch.cuda.Stream(device=dev)   # explicit stream (not the legacy default stream)
    torch.cuda.set_stream(work)
    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events per step, L2 flushed in between ----------------------
    def timed_pass():
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for k in range(args.steps):
            flush.zero_()
            ev[k][0].record()
            step()
            ev[k][1].record()
        barrier()
        return sum(a.elapsed_time(b) for a, b in ev)

    launches0 = h.counter("launches")
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    ms = timed_pass()
    sampler.stop()
    launches = h.counter("launches") - launches0
    # second, separate pass with CUDA events around every kernel launch (inside the library, on the
    # launching stream): per-kernel durations for the roofline; not used for `value`
    h.set_option("time_kernels", 1)
    ms_k = timed_pass()
    trmm_ns, trmm_calls = h.counter("trmm_ns"), h.counter("trmm_calls")
    rgen_ns, epi_ns, topk_ns = h.counter("rgen_ns"), h.counter("epilogue_ns"), h.counter("topk_ns")
    h.set_option("time_kernels", 0)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * m * args.steps / (ms_max * 1e-3)

    # ---- e2e: public host API, pinned host candidates in, EI values out, every step ---------------
    Xs_np = Xs_host.numpy()
    vals = torch.empty((wl["sets"], m), dtype=torch.float64).pin_memory().numpy()   # pinned result buffer
    crit.evaluate(Xs_np[:min(m, 200000)], params=params)  # warm the staging buffers
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        crit.evaluate(Xs_np, params=params, out=vals)
    barrier()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * m * args.steps / float(t_e.item())

    if rank == 0:
        err = None
        if not args.no_cpu_baseline:
            # cpu_baseline leg (the only place the oracle is touched): besides being timed below, the
            # CPU port checks a sample of what the timed GPU path produced (not timed)
            from oracle import gp_oracle as O
            sub = Xs_np[:2000 if n <= 4096 else 500]
            rv = O.acquisition(oracle_state(wl, X, y, theta), sub, wl["crit"],
                               None if wl["crit"] == "EI" else float(ts[0]), plugin, True)
            err = float(np.max(np.abs(vals[0][:len(sub)] - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max())))
        peak, peak_src = fp64_peak()
        chunk = h.counter("chunk")
        per_launch_c = m * args.steps / max(1, trmm_calls)
        flops_per_launch = per_launch_c * float(n) * n       # the n^2 term of F_pred
        achieved = flops_per_launch / (trmm_ns / max(1, trmm_calls) * 1e-9) / 1e12 if trmm_ns else None
        line = {
            "metric": "GP posterior+EI candidates/sec", "value": value, "unit": "candidates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "n": n, "d": d, "candidates_per_step_per_gpu": m,
                       "criterion": wl["crit"], "n_sets": wl["sets"], "kernel": kind, "chunk": chunk,
                       "l2": "256 MB buffer written between timed steps (L2 flush)",
                       "flops_per_candidate_Fpred": flops_per_candidate(n, d),
                       "sample_max_rel_err_vs_oracle": err},
            "e2e": {"value": e2e_value, "unit": "candidates/s", "h2d_bytes_per_step": int(m * d * 8),
                    "d2h_bytes_per_step": int(m * wl["sets"] * 8),
                    "api": "%s.evaluate(pinned host ndarray, out=pinned host ndarray) (mgp_acq)" % wl["crit"]},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "k_trmm<false> (T = L^-1 r, fused sum of squares)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": trmm_traffic(args.workload, chunk),
                         "peak_source": "FP64 " + peak_src + "; MEASURED_PEAKS.json has no FP64 entry",
                         "flops_per_launch": flops_per_launch, "launches_timed": int(trmm_calls),
                         "avg_launch_ms": trmm_ns / max(1, trmm_calls) * 1e-6,
                         "share_of_step": {"k_rgen": rgen_ns * 1e-6 / ms_k, "k_trmm": trmm_ns * 1e-6 / ms_k,
                                           "k_epilogue": epi_ns * 1e-6 / ms_k, "k_topk": topk_ns * 1e-6 / ms_k},
                         "whole_step_frac_of_peak": value / world * flops_per_candidate(n, d) / (peak * 1e12)},
            "clocks": sampler.summary(),
        }
        if not args.no_cpu_baseline:
            sample = min(m, 100_000) if n <= 1000 else (1000 if n <= 4096 else 400)
            r, dt = cpu_port_rate(wl, sample)
            line["cpu_baseline"] = {"value": r, "unit": "candidates/s", "cores": blas_threads(), "kind": "port",
                                    "sample": "%d candidates, vectorised predict+%s in chunks of 2000 "
                                              "(oracle port, %.1f s)" % (sample, wl["crit"], dt)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
