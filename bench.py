#!/usr/bin/env python
"""bench.py -- ELBO grad-steps/sec on the configs of BASELINE.json; headline = config 2 (configs[1]):
RepGradELBO + ClosedFormEntropy, MeanFieldGaussian, hierarchical logistic regression
n = 10000, d = 1024 (D = 1025), M = 256 Monte-Carlo samples, Adam(1e-3) + ClipScale + PolynomialAveraging.

One "step" = everything `step` does per iteration except the callback (src/algorithms/common.jl:75-104):
sample -> log-density + gradient on M samples -> entropy -> reduce to grad lambda and ELBO -> (exchange)
-> optimiser + operator + averaging.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2|c3|c4a|c4b|c5]
                  [--shard rows|samples] [--extras / --no-extras]

The printed JSON line is the --config workload (default c2).  With c2 at default settings the line also carries a
compact `configs` dict with the other BASELINE.json configs measured in the same process (bounded step counts).

N > 1 (under torchrun): default n-axis sharding -- every rank holds all M samples and n / N data rows, so each rank
ingests X / N; the partial gradient sums and sum_m log pi are exchanged over NVLink inside the iteration kernel's tail
phase.  --shard samples = M-axis (every rank streams all of X).  c5 is weak scaling: a fixed minibatch
per rank out of one reshuffled epoch, likeadj on the global batch.
"""
import argparse
import json
import os
import sys
import threading
import time

if "reference" in sys.argv:
    # the CPU arm uses every host core: torchrun exports OMP_NUM_THREADS=1 to its children, which throttled the
    # N > 1 reference runs of round 1 to one BLAS thread.  Must happen before numpy loads OpenBLAS.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_FLUSH_BYTES = 256 << 20

CONFIGS = {
    "c2": dict(n=10000, d=1024, M=256, family="meanfield", objective="rep", entropy="ClosedFormEntropy", model="logreg",
               opt=("adam", 1e-3), seed=1, scaling="strong",
               workload="C2: RepGradELBO+ClosedFormEntropy, MeanFieldGaussian, hier. logistic regression n=10000 d=1024 (D=1025), M=256"),
    "c3": dict(n=10000, d=1024, M=256, family="fullrank", objective="rep", entropy="ClosedFormEntropy", model="logreg",
               opt=("adam", 1e-3), seed=1, scaling="strong",
               workload="C3: RepGradELBO+ClosedFormEntropy, FullRankGaussian (lambda in R^1051650), hier. logistic regression n=10000 d=1024, M=256"),
    "c4a": dict(n=100000, d=4096, M=1024, family="meanfield", objective="score", entropy="ClosedFormEntropy", model="gaussglm",
                opt=("dog", 1e-6), seed=2, scaling="strong",
                workload="C4a: ScoreGradELBO (VarGrad), MeanFieldGaussian, Gaussian GLM n=100000 d=4096 (D=4097), M=1024, DoG"),
    "c4b": dict(n=100000, d=4096, M=1024, family="meanfield", objective="rep", entropy="StickingTheLandingEntropy", model="gaussglm",
                opt=("dog", 1e-6), seed=2, scaling="strong",
                workload="C4b: RepGradELBO+StickingTheLanding, MeanFieldGaussian, Gaussian GLM n=100000 d=4096 (D=4097), M=1024, DoG"),
    "c5": dict(n=1000000, d=512, M=512, family="meanfield", objective="rep", entropy="ClosedFormEntropy", model="logreg",
               opt=("adam", 1e-3), seed=3, batch=4096, scaling="weak",
               workload="C5: Subsampled RepGradELBO+ClosedFormEntropy (ReshufflingBatchSubsampling, batch 4096 rows per GPU), "
                        "MeanFieldGaussian, hier. logistic regression n=1000000 d=512 (D=513), M=512"),
}
OPT_DESC = {"adam": "Adam(1e-3)+ClipScale+PolynomialAveraging", "dog": "DoG+ClipScale+PolynomialAveraging"}


def synth(n, d, seed, gaussian=False, rows=None):
    """SURVEY.md 8(d) recipe: X_ij ~ N(0,1)/sqrt(d), last column == 1 (intercept), beta* ~ N(0,1),
    y ~ Bernoulli(sigmoid(X beta*)) or X beta* + N(0,1).  Plain numpy generator: nothing under oracle/ is needed
    to produce the inputs of either arm.  rows = (r0, nr): only that slice (generated block-wise so that every rank
    of a sharded run sees the same global data set without materialising all of it)."""
    beta = np.random.default_rng([seed, 1]).standard_normal(d).astype(np.float32)
    r0, nr = (0, n) if rows is None else rows
    blk = 8192
    X = np.empty((nr, d), np.float32)
    y = np.empty(nr, np.float32)
    for b0 in range((r0 // blk) * blk, r0 + nr, blk):
        rng = np.random.default_rng([seed, 2, b0 // blk])
        Xb = rng.standard_normal((min(blk, n - b0), d), dtype=np.float32) / np.float32(np.sqrt(d))
        Xb[:, d - 1] = 1.0
        lg = Xb @ beta
        yb = (lg + rng.standard_normal(len(lg)).astype(np.float32)).astype(np.float32) if gaussian else \
            (rng.random(len(lg)) < 1.0 / (1.0 + np.exp(-lg))).astype(np.float32)
        lo, hi = max(b0, r0), min(b0 + len(lg), r0 + nr)
        X[lo - r0:hi - r0] = Xb[lo - b0:hi - b0]
        y[lo - r0:hi - r0] = yb[lo - b0:hi - b0]
    return X, y


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:   # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:   # noqa: BLE001
                pass
            time.sleep(0.02)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path, restated (oracle/): Julia is not installed, so
    oracle/_ref cannot be built and the port is what runs, on all host cores of the box (rank 0 only).

    `value` = BEST-EFFORT CPU (the conservative denominator): every step is one WHOLE
    grad-step of the workload -- all M samples in one batched GEMM per pass, analytic gradients, Adam +
    ClipScale + PolynomialAveraging -- for exactly --steps steps after --warmup warm-ups.
    `cpu_baseline.reference_shaped_steps_per_s` = the same whole step run the way the reference is structured
    (one logdensity_and_gradient call per Monte-Carlo sample, src/algorithms/repgradelbo.jl:84-86: M GEMVs over X,
    Float64), timed on a few whole steps (it is ~4x slower, so it gets a smaller step count, never an
    extrapolation from a fraction of a step).
    Configs other than c2 are too large for whole CPU steps (c4: 8e11 FLOP per step): they run on a stated row / sample
    slice and the time is scaled linearly (said so in `sample`)."""
    if rank != 0:
        return
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:   # noqa: BLE001
        pass
    from oracle import family as F, models as Mo, objectives as O, optim as Op, philox as P
    cfg = CONFIGS[args.config]
    n, d, M = cfg["n"], cfg["d"], cfg["M"]
    scale = 1.0
    if args.config in ("c4a", "c4b"):
        n, M, scale = n // 16, M // 16, 256.0
    if args.config == "c5":
        n = cfg["batch"]          # one minibatch per step; the gather is a row slice on the CPU
    X, y = synth(n, d, cfg["seed"], gaussian=cfg["model"] == "gaussglm")
    prob = (Mo.GaussGLM if cfg["model"] == "gaussglm" else Mo.LogReg)(X, y, n_data=cfg["n"])
    D = d + 1
    q0 = F.MeanFieldGaussian(np.zeros(D), np.ones(D)) if cfg["family"] == "meanfield" else \
        F.FullRankGaussian(np.zeros(D), 0.6 * np.eye(D))
    rule = Op.Adam(1e-3) if cfg["opt"][0] == "adam" else Op.DoG()
    op, avg = Op.ClipScale(), Op.PolynomialAveraging()

    def run(per_sample, steps, warm):
        st = Op.sgd_init(q0, rule, avg)

        def grad_fn(params, t):
            eps = P.normal_matrix(cfg["seed"], t - 1, D, M)
            if cfg["objective"] == "score":
                v, g, e = O.scoregrad_value_and_gradient(params, q0, prob, eps)
            else:
                v, g, e = O.repgrad_value_and_gradient(params, q0, prob, eps, cfg["entropy"], per_sample=per_sample)
            return v, g, dict(elbo=e)
        for _ in range(warm):
            Op.sgd_step(st, q0, grad_fn, rule, op, avg)
        t0 = time.perf_counter()
        for _ in range(steps):
            Op.sgd_step(st, q0, grad_fn, rule, op, avg)
        return (time.perf_counter() - t0) / steps

    K, W = max(1, args.steps), max(0, args.warmup)
    batched_s = run(False, K, W) * scale
    cores = os.cpu_count()
    val = 1.0 / batched_s
    cb = {"value": val, "unit": "steps/s", "cores": cores, "kind": "port",
          "sample": f"{K} whole grad-steps (all {M} samples, one batched GEMM per pass, Float64 numpy/OpenBLAS, {cores} "
                    f"threads) after {W} warm-ups: best-effort CPU, the conservative denominator"
                    + (f"; run on n/16 rows x M/16 samples, time scaled by {scale:g}" if scale != 1.0 else "")}
    if args.config == "c2":
        k_ps = max(2, min(K, int(20.0 / max(4.0 * batched_s, 1e-3))))   # ~20 s of per-sample steps, whole steps only
        per_sample_s = run(True, k_ps, 1)
        cb["reference_shaped_steps_per_s"] = 1.0 / per_sample_s
        cb["reference_shaped_sample"] = (f"{k_ps} whole grad-steps of {M} per-sample logdensity_and_gradient calls "
                                         "(M GEMVs over X), no extrapolation")
    line = {
        "impl": "reference", "metric": "ELBO grad-steps/sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": batched_s * 1e3, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "optimizer": OPT_DESC[cfg["opt"][0]]},
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement (oracle/) of AdvancedVI.jl's path; the Julia package itself cannot run here",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Bench:
    """One config on this rank: data, target, objective, fused optimiser state."""

    _bpos = 0

    def __init__(self, name, args, ctx, rank, world, local_rank, gemm):
        import advancedvi_jl_b200 as avi
        from advancedvi_jl_b200 import parallel, _lib as L
        from advancedvi_jl_b200.api import _OptState
        self.avi, self.L, self.name, self.cfg = avi, L, name, CONFIGS[name]
        cfg = self.cfg
        self.ctx, self.rank, self.world, self.gemm = ctx, rank, world, gemm
        n, d, M = cfg["n"], cfg["d"], cfg["M"]
        self.D = D = d + 1
        gaussian = cfg["model"] == "gaussglm"
        cls = avi.GaussGLM if gaussian else avi.LogReg
        self.shard = args.shard if world > 1 else "none"
        self.subsampled = "batch" in cfg
        self.rows_local = n
        if self.subsampled or world == 1 or self.shard == "samples":
            X, y = synth(n, d, cfg["seed"], gaussian)
            self.prob = cls(ctx, X, y, gemm=gemm)
            if self.subsampled and world > 1:
                self.prob.set_data_shard(world, n, include_prior=(rank == 0))
        else:
            r0, nr = parallel.row_shard(n, rank, world, align=32)
            X, y = synth(n, d, cfg["seed"], gaussian, rows=(r0, nr))
            self.prob = cls(ctx, X, y, n_data=n, gemm=gemm)
            self.prob.set_data_shard(world, n, include_prior=(rank == 0))
            self.rows_local = nr
        self.X, self.y = (X, y) if name == "c2" else (None, None)
        if cfg["family"] == "meanfield":
            self.q0 = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
        else:
            self.q0 = avi.FullRankGaussian(np.zeros(D, np.float32), (0.6 * np.eye(D)).astype(np.float32))
        opt = avi.Adam(cfg["opt"][1]) if cfg["opt"][0] == "adam" else avi.DoG(cfg["opt"][1])
        ent = getattr(avi, cfg["entropy"])()
        if cfg["objective"] == "score":
            self.alg = avi.KLMinScoreGradDescent(optimizer=opt, n_samples=M, operator=avi.ClipScale())
        else:
            self.alg = avi.KLMinRepGradDescent(optimizer=opt, entropy=ent, n_samples=M, operator=avi.ClipScale())
        self.obj = avi.Objective(cfg["seed"], self.alg.objective, self.q0, self.prob)
        self.m_loc = M
        if world > 1:
            P = self.obj.P
            # exchange payload: the partial-sum vector (mean-field), or for the full-rank family the D x D contraction(s)
            # under sample sharding / the M x D gradient block under row sharding
            pad = (D + 31) // 32 * 32
            parallel.connect(ctx, max_floats=(4 * pad + 64) if cfg["family"] == "meanfield"
                             else max(2 * D * D + 4 * pad + 64, M * ((D + 3) // 4 * 4)), native=True)
            if self.subsampled or self.shard == "rows":
                self.obj.set_shard_axis(L.SHARD_ROWS)
            else:
                m0, ml = parallel.sample_shard(M, rank, world)
                self.obj.set_sample_shard(m0, ml)
                self.m_loc = ml
        self.state = _OptState(self.alg, self.obj, self.q0)
        self.batches = None
        if self.subsampled:
            # one reshuffled epoch, dealt round-robin to the ranks (reshuffling.jl:27-32)
            sub = avi.ReshufflingBatchSubsampling(np.arange(n), cfg["batch"])
            allb = [bb for _, bb in sub.reshuffle_batches(cfg["seed"], 0)]
            allb = [bb for bb in allb if len(bb) == cfg["batch"]]
            self.batches = parallel.rank_batches(allb, rank, world)

    def minibatches(self, n):
        """Row indices of the next n minibatches of this rank (host side of ReshufflingBatchSubsampling)."""
        idx = np.ascontiguousarray(np.stack([self.batches[(self._bpos + k) % len(self.batches)] for k in range(n)]), dtype=np.int32)
        self._bpos += n
        return idx

    def run_steps(self, n, vals, elbos, idx=None):
        import ctypes as C
        L, nd = self.L, C.c_int32()
        if self.subsampled:
            if idx is None:
                idx = self.minibatches(n)
            L.check(L.lib.avi_opt_steps_subsampled(self.state.h, n, L.iptr(idx), idx.shape[1], L.fptr(vals), L.fptr(elbos),
                                                   C.byref(nd)), self.ctx.h)
        else:
            L.check(L.lib.avi_opt_steps(self.state.h, n, L.fptr(vals), L.fptr(elbos), C.byref(nd)), self.ctx.h)
        assert nd.value == n, f"{self.name}: objective diverged"

    def sharding_desc(self):
        cfg = self.cfg
        if self.world == 1:
            return "none"
        if self.subsampled:
            return (f"weak scaling: {cfg['batch']} minibatch rows per rank per step (global batch {cfg['batch'] * self.world}), every "
                    f"rank holds all {cfg['M']} samples; exchange = [sum g, sum g*eps, sum log pi] over NVLink inside the iteration kernel")
        if self.shard == "rows":
            return (f"n-axis: {self.rows_local} data rows per rank, all {cfg['M']} samples on every rank; exchange = partial "
                    "gradient sums + sum_m log pi over NVLink (peer-memory low-latency push), fused into the iteration kernel's tail phase")
        return f"M-axis: {self.m_loc} samples per rank, every rank streams all of X; exchange = one-shot NVLink all-reduce of the partial sums"

    def close(self):
        self.state.close(); self.obj.close(); self.prob.close()
        if self.world > 1:
            import torch.distributed as dist
            self.ctx.disconnect_peers()
            dist.barrier()


def l2_flush(flush):
    """Evict everything from the 126 MB L2: write 256 MiB, then read 256 MiB.  The read pass matters: a write-only flush
    leaves ~126 MB of DIRTY lines behind, and their write-back (as the timed step pulls X in) would be charged to the
    step as extra DRAM traffic that is not the step's own."""
    flush.zero_()
    flush.sum()


def device_loop(b, K, W, torch, dist, ext, flush, barrier, sampler_index=None):
    """K timed iterations of the fused device loop.  Returns (cold_ms, warm_ms, launches, final_elbo, clocks).
    cold: L2 flushed (256 MiB write, untimed) before every step, per-step CUDA events on the library stream, the steps
    enqueued without a host round trip.  warm: K back-to-back replays, one event pair."""
    vals = np.empty(max(K, W), np.float32)
    elbos = np.empty_like(vals)
    b.run_steps(W, vals, elbos)
    sampler = ClockSampler(sampler_index) if sampler_index is not None else None
    if sampler:
        sampler.start()
    barrier()
    l0 = b.ctx.launch_count()
    if b.subsampled:
        # minibatch indices travel with the call: one blocking call of K steps between two events (X = 2 GB >> L2, the
        # gathered batch is produced inside the step: no flush needed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        idx = b.minibatches(K)
        if b.world > 1:
            b.ctx.comm_barrier()
        e0.record(ext)
        b.run_steps(K, vals, elbos, idx)
        e1.record(ext)
        barrier()
        cold_ms = e0.elapsed_time(e1)
        launches = b.ctx.launch_count() - l0
        warm_ms = cold_ms
    else:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        b.state.steps_begin(K)
        for k in range(K):
            with torch.cuda.stream(ext):
                l2_flush(flush)                                 # evict X, R, Z from L2 (untimed)
            if b.world > 1:
                b.ctx.comm_barrier()    # ranks leave their (untimed) flushes at different times: start the step together,
                                        # otherwise the wait for the slowest peer's flush is charged to the step's exchange
            with torch.cuda.stream(ext):
                ev[k][0].record(ext)
            b.state.steps_enqueue(1)
            ev[k][1].record(ext)
        _, cold_elbos, n_cold = b.state.steps_end()
        assert n_cold == K, f"{b.name}: objective diverged"
        barrier()
        cold_ms = sum(x.elapsed_time(z) for x, z in ev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = b.ctx.launch_count()               # counted over the back-to-back region: the flushed one launches the same
        e0.record(ext)                          # kernels per step plus, with several ranks, the untimed alignment barrier
        b.run_steps(K, vals, elbos)
        e1.record(ext)
        barrier()
        launches = b.ctx.launch_count() - l0
        warm_ms = e0.elapsed_time(e1)
    clocks = None
    if sampler:
        sampler.stop_flag = True
        sampler.join()
        clocks = sampler.summary()
    if b.world > 1:
        t = torch.tensor([cold_ms, warm_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cold_ms, warm_ms = t.tolist()
    return cold_ms, warm_ms, int(launches), float(elbos[K - 1]), clocks


def e2e_estimate_gradient(b, K, W, torch, dist, ext, flush):
    """End to end through the reference-facing boundary with HOST buffers: one `step` (common.jl:75-104) = estimate_gradient!
    (host lambda in, host gradient + value out: both transfers inside the timed region) + the host-side
    Optimisers.update! + ClipScale + PolynomialAveraging (common.jl:91-94), as ONE C-ABI call per iteration
    (avi_hoststep_step: the arrays are bound to a handle once).  Wall clock around the Python call, L2 flushed and the
    device idle before every step.  Returns (seconds for K steps, breakdown dict: the library's own wall clock of the two
    halves, and what the ctypes crossing adds)."""
    avi = b.avi
    hs = avi.HostStep(b.obj, b.alg.optimizer, b.alg.operator, b.alg.averager, b.q0.destructure(), scale_offset=b.D)
    b.obj.seed(b.cfg["seed"], 0)
    tot, t_call, t_upd, t_enq, t_wait = 0.0, 0.0, 0.0, 0.0, 0.0
    for k in range(W + K):
        with torch.cuda.stream(ext):
            l2_flush(flush)
        if b.world > 1:
            b.ctx.comm_barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v, e = hs.step()
        t1 = time.perf_counter()
        if k >= W:
            tc, tu, tq, tw = hs.timing()
            tot += t1 - t0; t_call += tc; t_upd += tu; t_enq += tq; t_wait += tw
    hs.close()
    if b.world > 1:
        t = torch.tensor([tot], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = t.item()
    return tot, {"estimate_gradient_call_us": t_call / K, "of_which_enqueue_us": t_enq / K,
                 "of_which_wait_for_completion_flag_us": t_wait / K, "host_update_us": t_upd / K,
                 "ffi_crossing_us": 1e6 * tot / K - (t_call + t_upd) / K}


def kernel_times(b, torch, ext, flush, names, reps=5):
    """Per-kernel device times: eager launches, CUDA events inside the library around the named kernels, L2 flushed."""
    vals, elbos = np.empty(4, np.float32), np.empty(4, np.float32)
    b.ctx.timing(True)
    for _ in range(reps):
        with torch.cuda.stream(ext):
            l2_flush(flush)
        b.run_steps(1, vals, elbos)
    b.ctx.timing(False)
    out = {}
    for name in names:
        ms, cnt = b.ctx.kernel_time(name)
        if cnt:
            out[name] = ms / cnt
    return out


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:   # noqa: BLE001
        return {}, "fallback (B200_PROFILING.md)"


def roofline_for(b, ktime, peaks, src, x3=False):
    """Dominant kernel of the step against the tensor roofline.  achieved = ALGORITHMIC FLOPs per launch (SURVEY.md 8d:
    4 n d M for the forward + backward contraction pair, 2 n d M for a single contraction kernel; the 3xTF32 mode issues
    3x the tensor work for the same algorithmic figure) / the kernel's average launch duration (CUDA events around the
    eager launch on the library stream, L2 flushed: includes the launch overhead, i.e. conservative)."""
    cfg = b.cfg
    bf16 = peaks.get("bf16_tflops", 1590.0)
    n_loc, d, m_loc = (b.rows_local if not b.subsampled else cfg["batch"]), cfg["d"], b.m_loc
    if "glm_step" in ktime:
        dom, flops, kname = "glm_step", 4.0 * n_loc * d * m_loc, "k_glm_mf_step (sample + forward + backward + tail, one launch)"
    elif "glm_fwd_bwd" in ktime:   # the target's batched logdensity_and_gradient as one persistent launch (full-rank path)
        dom, flops, kname = "glm_fwd_bwd", 4.0 * n_loc * d * m_loc, "k_glm_mf_step<EPI_STORE> (forward + backward, gradient block stored; one launch)"
    else:
        cands = {k: v for k, v in ktime.items() if k in ("glm_fwd", "glm_bwd")}
        if not cands:
            return None
        dom = max(cands, key=cands.get)
        flops, kname = 2.0 * n_loc * d * m_loc, f"k_gemm_tc<{dom}>"
    ach = flops / (ktime[dom] * 1e-3) / 1e12
    peak = bf16 / 2.0 / (3.0 if x3 else 1.0)
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this kernel
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if b.world == 1 and b.name == "c2" and not x3:
            traffic = tr["dram_bytes_per_launch"].get(dom)
    except Exception:   # noqa: BLE001
        pass
    return {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": traffic,
            "peak_note": f"kind::tf32 dense = 1/2 of the {src} bf16 cuBLAS peak ({bf16} TFLOP/s)"
                         + (", / 3 for the 3xTF32 mode (three tensor products per algorithmic product)" if x3 else "")
                         + f"; frac of the bf16 figure itself: {ach / bf16:.4f}",
            "flops_per_launch": flops, "kernel_ms": {k: round(v, 5) for k, v in ktime.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="c2", choices=list(CONFIGS))
    ap.add_argument("--gemm", default="tf32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", dest="extras", action="store_true", default=None,
                    help="also measure the other BASELINE.json configs (default: on for c2)")
    ap.add_argument("--no-extras", dest="extras", action="store_false")
    ap.add_argument("--shard", default="rows", choices=["rows", "samples"],
                    help="multi-GPU axis: data rows (n-axis, default: every rank ingests X/N) or Monte-Carlo samples "
                         "(M-axis: every rank streams all of X)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import advancedvi_jl_b200 as avi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, max(args.warmup, 3)
    cfg = CONFIGS[args.config]
    ctx = avi.Context(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device="cuda")   # (float32: .sum() reads it in place; a uint8
                                                                                 # buffer is first converted into a 2 GB int64 temporary)
    peaks, src = load_peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    b = Bench(args.config, args, ctx, rank, world, local_rank, args.gemm)
    D, P = b.D, b.obj.P
    cold_ms, warm_ms, launches, final_elbo, clocks = device_loop(b, K, W, torch, dist, ext, flush, barrier, local_rank)
    meanfield = cfg["family"] == "meanfield"
    if meanfield and cfg["objective"] == "rep" and not b.subsampled:
        e2e_s, e2e_parts = e2e_estimate_gradient(b, K, W, torch, dist, ext, flush)
        e2e = {"value": K / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": 4 * P, "d2h_bytes_per_step": 4 * (P + 5),
               "path": "avi_hoststep_step = avi_obj_estimate_gradient (estimate_gradient! boundary: host lambda in, host "
                       "gradient + value + completion flag out) + host Adam/ClipScale/averaging (avi_host_update), one "
                       "C-ABI call per step; wall clock around the Python call, L2 flushed and device idle before every step",
               "breakdown": e2e_parts}
    else:
        # the call a user makes for these configs is optimize(): one blocking call per chunk of iterations with the
        # minibatch indices (c5) going host -> device and the ELBO trace coming back
        vals, elbos = np.empty(K, np.float32), np.empty(K, np.float32)
        barrier()
        t0 = time.perf_counter()
        b.run_steps(K, vals, elbos)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = t.item()
        e2e = {"value": K / dt, "unit": "steps/s", "h2d_bytes_per_step": 4 * cfg.get("batch", 0), "d2h_bytes_per_step": 8,
               "path": "avi_opt_steps[_subsampled] (the optimize() loop: minibatch indices host -> device, (value, elbo) trace "
                       "device -> host, one synchronisation per call of K iterations); wall clock"}
    ktime = kernel_times(b, torch, ext, flush, ("glm_step", "glm_fwd_bwd", "sample", "glm_fwd", "glm_bwd", "gemm_store"))
    roofline = roofline_for(b, ktime, peaks, src, x3=args.gemm == "tf32x3")

    # the sample+transform kernel at a bandwidth-relevant size (output 134 MB > 126 MB L2): same kernel, M = 32768, as
    # estimate_objective launches it (forward only: eps is regenerated from Philox where needed, never written)
    sample_large = None
    if world == 1 and args.config == "c2":
        M_big = 32768
        qb = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
        pb = avi.MvNormalDiag(ctx, np.zeros(D, np.float32), np.ones(D, np.float32))
        ob = avi.Objective(cfg["seed"], avi.RepGradELBO(8), qb, pb)
        ob.estimate_objective(cfg["seed"], qb, M_big)                      # sizes the buffers (untimed)
        ctx.timing(True)
        for _ in range(5):
            ob.estimate_objective(cfg["seed"], qb, M_big)
        ctx.timing(False)
        ms_b, cnt_b = ctx.kernel_time("sample")
        ld = (D + 3) // 4 * 4
        per = ms_b / max(cnt_b, 1)
        alg_bytes = 4 * (2 * D + D * M_big)                                # read mu, s; write Z
        moved_bytes = 4 * (2 * D + ld * M_big)                             # what the forward-only sampler moves (z padded to ld; eps is not materialised)
        sample_large = {"M": M_big, "ms": per, "algorithmic_bytes": alg_bytes, "moved_bytes": moved_bytes,
                        "achieved_gbs_algorithmic": alg_bytes / (per * 1e-3) / 1e9 if per > 0 else None,
                        "achieved_gbs_moved": moved_bytes / (per * 1e-3) / 1e9 if per > 0 else None,
                        "frac_algorithmic": alg_bytes / (per * 1e-3) / 1e9 / hbm if per > 0 else None,
                        "frac_moved": moved_bytes / (per * 1e-3) / 1e9 / hbm if per > 0 else None, "peak_gbs": hbm}
        ob.close(); pb.close()
    if roofline is not None and sample_large is not None:
        roofline["sample_kernel_hbm"] = sample_large

    line = {
        "metric": "ELBO grad-steps/sec", "value": K / (cold_ms * 1e-3), "unit": "steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": cold_ms / K, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": {"tf32": "tf32", "tf32x3": "tf32x3", "fp32": "f32"}.get(args.gemm, args.gemm),
        "data": "synthetic",
        "config": {"workload": cfg["workload"], "optimizer": OPT_DESC[cfg["opt"][0]],
                   "l2": ("inputs larger than L2 (X = 2 GB; the gathered minibatch is produced inside the step); one event pair around K steps"
                          if b.subsampled else
                          "flushed between timed steps (256 MiB device write + 256 MiB read so that the lines left in L2 are clean; untimed); per-step CUDA events"),
                   "sharding": b.sharding_desc(),
                   "contraction": {"tf32": "tcgen05 kind::tf32 (operands rounded to nearest TF32, fp32 accumulate)",
                                   "tf32x3": "tcgen05 kind::tf32, 3xTF32 split operands (fp32-grade)",
                                   "fp32": "SIMT fp32"}.get(args.gemm)},
        "value_l2_resident": K / (warm_ms * 1e-3), "ms_per_step_l2_resident": warm_ms / K,
        "e2e": e2e, "gpu_launches": launches, "launches_per_step": launches / K, "final_elbo": final_elbo,
        "clocks": clocks, "roofline": roofline,
    }

    # ---- the fp32-grade tensor-core mode (3xTF32 by K-concatenation) as a complete second record ----
    b3 = None
    if world == 1 and args.config == "c2" and args.gemm == "tf32":
        try:
            b3 = Bench("c2", args, ctx, rank, world, local_rank, "tf32x3")
            c3_ms, w3_ms, l3, fe3, _ = device_loop(b3, K, W, torch, dist, ext, flush, barrier)
            e3_s, e3_parts = e2e_estimate_gradient(b3, K, W, torch, dist, ext, flush)
            k3 = kernel_times(b3, torch, ext, flush, ("glm_step", "glm_fwd_bwd", "glm_fwd", "glm_bwd"))
            line["alt_precision"] = {
                "mode": "tf32x3 (hi/lo split operands, 3 tensor-core products per algorithmic product, fp32-grade)",
                "value": K / (c3_ms * 1e-3), "ms_per_step": c3_ms / K, "value_l2_resident": K / (w3_ms * 1e-3),
                "unit": "steps/s", "e2e": {"value": K / e3_s, "unit": "steps/s", "breakdown": e3_parts},
                "gpu_launches": l3, "final_elbo": fe3, "roofline": roofline_for(b3, k3, peaks, src, x3=True)}
        except Exception as ex:   # noqa: BLE001
            line["alt_precision"] = {"error": repr(ex)}

    # ---- CPU baseline (bounded sample of the same workload) and parity of the timed configuration ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.config == "c2":
        from oracle import family as F, models as Mo, objectives as O, philox as Ph
        probo = Mo.LogReg(b.X, b.y)
        qo = F.MeanFieldGaussian(np.zeros(D), np.ones(D))
        M = cfg["M"]
        epsf = Ph.normal_matrix(cfg["seed"], 0, D, M)
        O.repgrad_value_and_gradient(qo.destructure(), qo, probo, epsf, "ClosedFormEntropy")
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            vo, go, eo = O.repgrad_value_and_gradient(qo.destructure(), qo, probo, epsf, "ClosedFormEntropy")
        bt = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        reps_ps = 4
        for _ in range(reps_ps):
            O.repgrad_value_and_gradient(qo.destructure(), qo, probo, epsf, "ClosedFormEntropy", per_sample=True)
        pt = (time.perf_counter() - t0) / reps_ps
        line["cpu_baseline"] = {"value": 1.0 / bt, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{reps} whole value-and-gradient evaluations (all {M} samples, one batched GEMM per pass, "
                                          "Float64 numpy/OpenBLAS, all cores): best-effort CPU, the conservative denominator",
                                "reference_shaped_steps_per_s": 1.0 / pt,
                                "reference_shaped_sample": f"{reps_ps} whole evaluations as {M} per-sample logdensity_and_gradient calls"}
        # parity of the timed configurations against the fp64 oracle on the same eps (step 0)
        b.obj.seed(cfg["seed"], 0)
        v, g, e = b.obj.estimate_gradient(b.q0.destructure())
        line["parity"] = {"elbo_rel_err_vs_fp64_oracle": abs(v - vo) / abs(vo),
                          "grad_rel_err_vs_fp64_oracle": float(np.linalg.norm(g - go) / np.linalg.norm(go)),
                          "tolerance": {"elbo": 5e-4, "grad": 2e-3}}
        if b3 is not None and "error" not in line.get("alt_precision", {}):
            b3.obj.seed(cfg["seed"], 0)
            v3, g3, e3 = b3.obj.estimate_gradient(b3.q0.destructure())
            line["alt_precision"]["parity"] = {"elbo_rel_err_vs_fp64_oracle": abs(v3 - vo) / abs(vo),
                                               "grad_rel_err_vs_fp64_oracle": float(np.linalg.norm(g3 - go) / np.linalg.norm(go)),
                                               "tolerance": {"elbo": 1e-5, "grad": 5e-5}}
    if b3 is not None:
        b3.close()

    # ---- the other BASELINE.json configs, compact (bounded step counts; same timing rules) ----
    extras = args.extras if args.extras is not None else (args.config == "c2")
    if extras:
        line["configs"] = {}
        b.close()
        b = None
        for name, (k2, w2) in (("c3", (40, 5)), ("c4a", (10, 3)), ("c4b", (10, 3)), ("c5", (100, 10))):
            if name == args.config:
                continue
            try:
                t_setup = time.perf_counter()
                bx = Bench(name, args, ctx, rank, world, local_rank, args.gemm)
                setup_s = time.perf_counter() - t_setup
                cm, wm, ln, fe, _ = device_loop(bx, k2, w2, torch, dist, ext, flush, barrier)
                kt = kernel_times(bx, torch, ext, flush, ("glm_step", "glm_fwd_bwd", "glm_fwd", "glm_bwd"), reps=3)
                rf = roofline_for(bx, kt, peaks, src)
                line["configs"][name] = {"workload": CONFIGS[name]["workload"], "value": k2 / (cm * 1e-3), "unit": "steps/s",
                                         "ms_per_step": cm / k2, "value_l2_resident": k2 / (wm * 1e-3), "steps": k2, "warmup": w2,
                                         "scaling": CONFIGS[name]["scaling"], "sharding": bx.sharding_desc(),
                                         "launches_per_step": ln / k2, "final_elbo": fe, "setup_s": round(setup_s, 1),
                                         "roofline": None if rf is None else {k: rf[k] for k in ("kernel", "achieved", "peak", "frac", "unit")},
                                         "parity": "tests/test_gpu_objective_parity.py (one-step oracle parity at this config's width)"}
                bx.close()
                del bx
            except Exception as ex:   # noqa: BLE001
                line["configs"][name] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if b is not None:
        b.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
