#!/usr/bin/env python
"""bench.py -- ELBO grad-steps/sec on BASELINE.json config 2 (configs[1]):
RepGradELBO + ClosedFormEntropy, MeanFieldGaussian, hierarchical logistic regression
n = 10000, d = 1024 (D = 1025), M = 256 Monte-Carlo samples, Adam(1e-3) + ClipScale + PolynomialAveraging.

One "step" = everything `step` does per iteration except the callback (src/algorithms/common.jl:75-104):
sample -> log-density + gradient on M samples -> entropy -> reduce to grad lambda and ELBO -> (exchange)
-> optimiser + operator + averaging.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (under torchrun): the M samples are sharded over the ranks (strong scaling) and the
partial gradient sums are exchanged by the library's one-shot NVLink all-reduce kernel.
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

N_ROWS, N_FEAT, N_MC = 10000, 1024, 256
SEED = 1
WORKLOAD = "C2: RepGradELBO+ClosedFormEntropy, MeanFieldGaussian, hier. logistic regression n=10000 d=1024 (D=1025), M=256"
L2_FLUSH_BYTES = 256 << 20


def synth(n, d, seed, gaussian=False):
    """SURVEY.md 8(d) recipe: X_ij ~ N(0,1)/sqrt(d), last column == 1 (intercept), beta* ~ N(0,1),
    y ~ Bernoulli(sigmoid(X beta*)) or X beta* + N(0,1).  Plain numpy generator: nothing under oracle/ is needed
    to produce the inputs of either arm."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d), dtype=np.float32) / np.float32(np.sqrt(d))
    X[:, d - 1] = 1.0
    beta = rng.standard_normal(d).astype(np.float32)
    logits = X @ beta
    if gaussian:
        y = (logits + rng.standard_normal(n).astype(np.float32)).astype(np.float32)
    else:
        y = (rng.random(n) < 1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
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
    extrapolation from a fraction of a step)."""
    if rank != 0:
        return
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:   # noqa: BLE001
        pass
    from oracle import family as F, models as Mo, objectives as O, optim as Op, philox as P
    X, y = synth(N_ROWS, N_FEAT, SEED)
    prob = Mo.LogReg(X, y)
    D = N_FEAT + 1
    q0 = F.MeanFieldGaussian(np.zeros(D), np.ones(D))
    rule, op, avg = Op.Adam(1e-3), Op.ClipScale(), Op.PolynomialAveraging()

    def run(per_sample, steps, warm):
        st = Op.sgd_init(q0, rule, avg)

        def grad_fn(params, t):
            eps = P.normal_matrix(SEED, t - 1, D, N_MC)
            v, g, e = O.repgrad_value_and_gradient(params, q0, prob, eps, "ClosedFormEntropy", per_sample=per_sample)
            return v, g, dict(elbo=e)
        for _ in range(warm):
            Op.sgd_step(st, q0, grad_fn, rule, op, avg)
        t0 = time.perf_counter()
        for _ in range(steps):
            Op.sgd_step(st, q0, grad_fn, rule, op, avg)
        return (time.perf_counter() - t0) / steps

    K, W = max(1, args.steps), max(0, args.warmup)
    batched_s = run(False, K, W)
    k_ps = max(2, min(K, int(20.0 / max(4.0 * batched_s, 1e-3))))   # ~20 s of per-sample steps, whole steps only
    per_sample_s = run(True, k_ps, 1)
    cores = os.cpu_count()
    val = 1.0 / batched_s
    line = {
        "impl": "reference", "metric": "ELBO grad-steps/sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": batched_s * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "optimizer": "Adam(1e-3)+ClipScale+PolynomialAveraging"},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": f"{K} whole grad-steps (all {N_MC} samples, one batched GEMM per pass, Float64 "
                                   f"numpy/OpenBLAS, {cores} threads) after {W} warm-ups: best-effort CPU, the conservative "
                                   "denominator",
                         "reference_shaped_steps_per_s": 1.0 / per_sample_s,
                         "reference_shaped_sample": f"{k_ps} whole grad-steps of {N_MC} per-sample logdensity_and_gradient "
                                                    "calls (M GEMVs over X), no extrapolation"},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement (oracle/) of AdvancedVI.jl's path; the Julia package itself cannot run here",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--gemm", default="tf32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from advancedvi_jl_b200 import parallel

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    K, W = args.steps, max(args.warmup, 3)

    X, y = synth(N_ROWS, N_FEAT, SEED)
    ctx = avi.Context(local_rank)
    D = N_FEAT + 1
    P = 2 * D
    q0 = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
    alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-3), n_samples=N_MC, operator=avi.ClipScale())
    rows_local = N_ROWS
    if world > 1 and args.shard == "rows":
        r0, rows_local = parallel.row_shard(N_ROWS, rank, world, align=32)
        prob = avi.LogReg(ctx, X[r0:r0 + rows_local], y[r0:r0 + rows_local], n_data=N_ROWS, gemm=args.gemm)
        prob.set_data_shard(world, N_ROWS, include_prior=(rank == 0))
    else:
        prob = avi.LogReg(ctx, X, y, gemm=args.gemm)
    obj = avi.Objective(SEED, alg.objective, q0, prob)
    if world > 1:
        parallel.connect(ctx, max_floats=4 * 1056 + 64, native=True)
        if args.shard == "rows":
            from advancedvi_jl_b200 import _lib as _L
            obj.set_shard_axis(_L.SHARD_ROWS)
        else:
            m0, ml = parallel.sample_shard(N_MC, rank, world)
            obj.set_sample_shard(m0, ml)
    from advancedvi_jl_b200.api import _OptState
    from advancedvi_jl_b200 import _lib as L
    import ctypes as C
    state = _OptState(alg, obj, q0)

    ext = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    vals = np.empty(max(K, W), np.float32)
    elbos = np.empty_like(vals)
    nd = C.c_int32()

    def run_steps(n):
        L.check(L.lib.avi_opt_steps(state.h, n, L.fptr(vals), L.fptr(elbos), C.byref(nd)), ctx.h)
        assert nd.value == n, "objective diverged"

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident steps: (a) L2 flushed between timed steps, (b) back-to-back graph replays ----
    run_steps(W)
    launches_per_step = None
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l0 = ctx.launch_count()
    # the K iterations are enqueued without a host round trip (avi_opt_steps_begin / _enqueue / _end), so the event
    # pairs bracket device work only: flush | ev0 | one captured iteration | ev1
    state.steps_begin(K)
    for k in range(K):
        with torch.cuda.stream(ext):
            flush.zero_()                                   # evict X, R, Z from L2 (untimed)
            ev[k][0].record(ext)
        state.steps_enqueue(1)
        ev[k][1].record(ext)
    _, cold_elbos, n_cold = state.steps_end()
    assert n_cold == K, "objective diverged"
    barrier()
    launches = ctx.launch_count() - l0
    cold_ms = sum(a.elapsed_time(b) for a, b in ev)
    # back-to-back (L2-resident) replays of the captured iteration, one sync for all K
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    run_steps(K)
    e1.record(ext)
    barrier()
    warm_ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    sampler.join()
    if world > 1:
        t = torch.tensor([cold_ms, warm_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cold_ms, warm_ms = t.tolist()
    final_elbo = float(elbos[K - 1])

    # ---- end-to-end through the reference-facing call: estimate_gradient! with HOST buffers, then the host-side
    # Optimisers.update! + ClipScale + PolynomialAveraging of `step` (common.jl:91-94) as compiled host code
    # (avi_host_update; a numpy version of the same update costs ~45 us per step and would dominate) ----
    host = avi.HostUpdate(alg.optimizer, alg.operator, alg.averager, q0.destructure(), scale_offset=D)
    gbuf = np.empty(P, np.float32)
    obj.seed(SEED, 0)
    e2e_s = 0.0
    for k in range(W + K):
        with torch.cuda.stream(ext):
            flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v, g, e = obj.estimate_gradient(host.lam, out=gbuf)  # H2D lambda, kernels, D2H gradient + value
        host.update(g)
        dt = time.perf_counter() - t0
        if k >= W:
            e2e_s += dt
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()

    # ---- per-kernel device times (eager launches, CUDA events inside the library) -> roofline ----
    ctx.timing(True)
    for _ in range(3):
        with torch.cuda.stream(ext):
            flush.zero_()
        run_steps(1)
    ctx.timing(False)
    ktime = {}
    for name in ("sample", "glm_fwd", "glm_bwd"):
        ms, cnt = ctx.kernel_time(name)
        ktime[name] = ms / max(cnt, 1)

    # the sample+transform kernel at a bandwidth-relevant size (outputs >> 126 MB L2): same kernel, M = 32768
    sample_large = None
    if world == 1:
        M_big = 32768
        qb = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
        pb = avi.MvNormalDiag(ctx, np.zeros(D, np.float32), np.ones(D, np.float32))
        ob = avi.Objective(SEED, avi.RepGradELBO(8), qb, pb)
        ob.estimate_objective(SEED, qb, M_big)                      # sizes the buffers (untimed)
        ctx.timing(True)
        for _ in range(5):
            ob.estimate_objective(SEED, qb, M_big)
        ctx.timing(False)
        ms_b, cnt_b = ctx.kernel_time("sample")
        ld = (D + 3) // 4 * 4
        bytes_b = 4 * (2 * D + 2 * ld * M_big)                      # read mu, s; write Z and eps (materialised)
        sample_large = {"M": M_big, "bytes_per_launch": bytes_b, "ms": ms_b / max(cnt_b, 1),
                        "achieved_gbs": bytes_b / (ms_b / max(cnt_b, 1) * 1e-3) / 1e9 if ms_b > 0 else None}
        ob.close(); pb.close()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:   # noqa: BLE001
        pass
    bf16 = peaks.get("bf16_tflops", 1590.0)
    hbm = peaks.get("hbm_gbs", 6650.0)
    src = "measured" if peaks else "fallback"
    m_loc = N_MC // world if (world > 1 and args.shard == "samples") else N_MC
    flops_fwd = 2.0 * rows_local * N_FEAT * m_loc
    dom = "glm_fwd" if ktime["glm_fwd"] >= ktime["glm_bwd"] else "glm_bwd"
    ach = flops_fwd / (ktime[dom] * 1e-3) / 1e12 if ktime[dom] > 0 else None
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["dram_bytes_per_launch"].get(dom)
        if world > 1:
            traffic = None   # captured at N = 1
    except Exception:   # noqa: BLE001
        pass
    roofline = {"bound": "tensor", "kernel": f"k_gemm_tc<{dom}>", "achieved": ach, "peak": bf16 / 2.0, "unit": "TFLOP/s",
                "frac": (ach / (bf16 / 2.0)) if ach else None, "traffic": traffic,
                "peak_note": f"kind::tf32 dense = 1/2 of the {src} bf16 cuBLAS peak ({bf16} TFLOP/s); "
                             f"frac of the bf16 figure itself: {ach / bf16 if ach else None}",
                "flops_per_launch": flops_fwd,
                "kernel_ms": {k: round(v, 5) for k, v in ktime.items()},
                "sample_kernel_hbm": {"bytes_per_launch": 4 * (2 * D + 2 * D * m_loc),
                                      "achieved_gbs": (4 * (2 * D + 2 * D * m_loc) / (ktime["sample"] * 1e-3) / 1e9)
                                      if ktime["sample"] > 0 else None,
                                      "peak_gbs": hbm, "note": "2 MB launch: latency-bound at this shape (SURVEY F8)",
                                      "large_shape": dict(sample_large, frac=(sample_large["achieved_gbs"] / hbm
                                                          if sample_large and sample_large["achieved_gbs"] else None))
                                      if sample_large else None}}

    line = {
        "metric": "ELBO grad-steps/sec", "value": K / (cold_ms * 1e-3), "unit": "steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": cold_ms / K, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "tf32" if args.gemm == "tf32" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "optimizer": "Adam(1e-3)+ClipScale+PolynomialAveraging",
                   "l2": "flushed between timed steps (256 MiB device write, untimed); per-step CUDA events",
                   "sharding": "none" if world == 1 else
                   (f"n-axis: {rows_local} data rows per rank, all {N_MC} samples on every rank; exchange = one-shot NVLink "
                    f"all-reduce kernels of [sum g, sum g*eps] (2x1056 floats) and log pi ({N_MC} floats)"
                    if args.shard == "rows" else
                    f"M-axis: {m_loc} samples per rank; exchange = one-shot NVLink all-reduce of the partial sums"),
                   "contraction": "tcgen05 kind::tf32 (operands rounded to nearest TF32, fp32 accumulate)"
                   if args.gemm == "tf32" else "SIMT fp32"},
        "value_l2_resident": K / (warm_ms * 1e-3), "ms_per_step_l2_resident": warm_ms / K,
        "e2e": {"value": K / e2e_s, "unit": "steps/s", "h2d_bytes_per_step": 4 * P, "d2h_bytes_per_step": 4 * (P + 5),   # lambda in; gradient + 4 scalars + flag out
                "path": "avi_obj_estimate_gradient (estimate_gradient! boundary, host lambda in / host gradient out) "
                        "+ host Adam/ClipScale/averaging (avi_host_update), L2 flushed between steps"},
        "gpu_launches": int(launches), "final_elbo": final_elbo,
        "clocks": sampler.summary(), "roofline": roofline,
    }

    # the fp32-grade tensor-core mode (3xTF32 by K-concatenation) on the same workload, L2-resident replays
    if world == 1 and args.gemm == "tf32":
        p3 = avi.LogReg(ctx, X, y, gemm="tf32x3")
        o3 = avi.Objective(SEED, alg.objective, q0, p3)
        s3 = _OptState(alg, o3, q0)
        L.check(L.lib.avi_opt_steps(s3.h, W, L.fptr(vals), L.fptr(elbos), C.byref(nd)), ctx.h)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(ext)
        L.check(L.lib.avi_opt_steps(s3.h, K, L.fptr(vals), L.fptr(elbos), C.byref(nd)), ctx.h)
        a1.record(ext)
        barrier()
        line["alt_precision"] = {"mode": "tf32x3 (hi/lo split operands, 3 tensor-core products, fp32-grade)",
                                 "value_l2_resident": K / (a0.elapsed_time(a1) * 1e-3), "unit": "steps/s"}
        s3.close(); o3.close(); p3.close()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import family as F, models as Mo, objectives as O, philox as Ph
        probo = Mo.LogReg(X, y)
        qo = F.MeanFieldGaussian(np.zeros(D), np.ones(D))
        ms = 8
        eps = Ph.normal_matrix(SEED, 0, D, ms)
        O.repgrad_value_and_gradient(qo.destructure(), qo, probo, eps[:, :2], "ClosedFormEntropy", per_sample=True)
        t0 = time.perf_counter()
        reps = 4
        for _ in range(reps):
            vo, go, eo = O.repgrad_value_and_gradient(qo.destructure(), qo, probo, eps, "ClosedFormEntropy", per_sample=True)
        dt = (time.perf_counter() - t0) / reps * (N_MC / ms)
        epsf = Ph.normal_matrix(SEED, 0, D, N_MC)
        t0 = time.perf_counter()
        vo, go, eo = O.repgrad_value_and_gradient(qo.destructure(), qo, probo, epsf, "ClosedFormEntropy")
        bt = time.perf_counter() - t0
        # parity of the timed configuration against the fp64 oracle on the same eps (step 0)
        obj.seed(SEED, 0)
        v, g, e = obj.estimate_gradient(q0.destructure())
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"{reps} x {ms}/{N_MC} per-sample logdensity_and_gradient calls (M GEMVs over X, "
                                          f"Float64, numpy/OpenBLAS), scaled by {N_MC // ms}",
                                "batched_gemm_steps_per_s": 1.0 / bt}
        line["parity"] = {"elbo_rel_err_vs_fp64_oracle": abs(v - vo) / abs(vo),
                          "grad_rel_err_vs_fp64_oracle": float(np.linalg.norm(g - go) / np.linalg.norm(go)),
                          "tolerance": {"elbo": 5e-4, "grad": 2e-3}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    state.close(); obj.close(); prob.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
