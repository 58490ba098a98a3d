"""Multi-rank parity check, run as `torchrun --nproc-per-node N tests/multigpu_check.py` (one rank per GPU).
Sharded results (M-axis and n-axis, native NVLink exchange and the NCCL callback) must reproduce the
single-rank result of the same library up to summation order, and must be bitwise identical across ranks."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import advancedvi_jl_b200 as avi
    from advancedvi_jl_b200 import parallel, _lib as L
    from oracle import models as Mo
    from advancedvi_jl_b200.api import _OptState
    import ctypes as C

    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n, d, M, key = 2000, 96, 64, 11
    X, y = Mo.synth_glm_data(n, d, seed=4)
    D = d + 1
    q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.full(D, 0.3, np.float32))
    lam = q.destructure()
    results = {}

    def gather_equal(a, what):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        for r in range(world):
            assert torch.equal(allt[0], allt[r]), f"{what}: rank {r} differs bitwise from rank 0"

    # reference: unsharded, single rank
    c0 = avi.Context(lr)
    p0 = avi.LogReg(c0, X, y, gemm="tf32")
    for kind in ("rep", "stl", "score"):
        spec = {"rep": avi.RepGradELBO(M), "stl": avi.RepGradELBO(M, avi.StickingTheLandingEntropy()),
                "score": avi.ScoreGradELBO(M)}[kind]
        o0 = avi.Objective(key, spec, q, p0)
        results[kind] = o0.estimate_gradient(lam)
        o0.close()
    alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), n_samples=M, operator=avi.ClipScale())
    _, info0, st0 = avi.optimize(key, alg, 25, p0, q)
    lam0, avg0, _ = st0.params()

    for native in (True, False):
        ctx = avi.Context(lr)
        parallel.connect(ctx, max_floats=4 * 128 + 64, native=native)
        # ---- M-axis ----
        prob = avi.LogReg(ctx, X, y, gemm="tf32")
        m0, ml = parallel.sample_shard(M, rank, world)
        for kind in ("rep", "stl", "score"):
            spec = {"rep": avi.RepGradELBO(M), "stl": avi.RepGradELBO(M, avi.StickingTheLandingEntropy()),
                    "score": avi.ScoreGradELBO(M)}[kind]
            o = avi.Objective(key, spec, q, prob)
            o.set_sample_shard(m0, ml)
            v, g, e = o.estimate_gradient(lam)
            v0, g0, e0 = results[kind]
            tol = 2e-3 if kind == "score" else 2e-5
            assert abs(v - v0) <= tol * abs(v0) + 1e-6, (kind, v, v0)
            assert np.linalg.norm(g - g0) <= tol * np.linalg.norm(g0), (kind, native)
            gather_equal(g, f"grad {kind}")
            o.close()
        _, info, st = avi.optimize(key, alg, 25, prob, q, state=None) if False else (None, None, None)
        obj = avi.Objective(key, alg.objective, q, prob)
        obj.set_sample_shard(m0, ml)
        from advancedvi_jl_b200.api import _OptState
        import ctypes as C
        st = _OptState(alg, obj, q)
        vals, elbos, nd = np.empty(25, np.float32), np.empty(25, np.float32), C.c_int32()
        L.check(L.lib.avi_opt_steps(st.h, 25, L.fptr(vals), L.fptr(elbos), C.byref(nd)), ctx.h)
        assert nd.value == 25
        lam1, avg1, _ = st.params()
        assert np.linalg.norm(lam1 - lam0) <= 1e-4 * np.linalg.norm(lam0), np.linalg.norm(lam1 - lam0)
        gather_equal(lam1, "lambda after 25 sharded steps")
        st.close(); obj.close(); prob.close()
        # ---- n-axis: every rank holds all samples and a slice of the rows ----
        r0, nr = parallel.row_shard(n, rank, world, align=32)
        probr = avi.LogReg(ctx, X[r0:r0 + nr], y[r0:r0 + nr], n_data=n, gemm="tf32")
        probr.set_data_shard(world, n, include_prior=(rank == 0))
        for kind in ("rep", "score"):
            spec = avi.RepGradELBO(M) if kind == "rep" else avi.ScoreGradELBO(M)
            o = avi.Objective(key, spec, q, probr)
            o.set_shard_axis(L.SHARD_ROWS)
            v, g, e = o.estimate_gradient(lam)
            v0, g0, e0 = results[kind]
            tol = 2e-3 if kind == "score" else 2e-5
            assert abs(v - v0) <= tol * abs(v0) + 1e-6, (kind, v, v0)
            assert np.linalg.norm(g - g0) <= tol * np.linalg.norm(g0), (kind, "rows", native)
            gather_equal(g, f"row-sharded grad {kind}")
            o.close()
        # n-axis inside the fused optimiser loop (the bench default): 25 steps, rows sharded, exchange in the tail phase
        objr = avi.Objective(key, alg.objective, q, probr)
        objr.set_shard_axis(L.SHARD_ROWS)
        str_ = _OptState(alg, objr, q)
        L.check(L.lib.avi_opt_steps(str_.h, 25, L.fptr(vals), L.fptr(elbos), C.byref(nd)), ctx.h)
        assert nd.value == 25
        lam2, avg2, _ = str_.params()
        assert np.linalg.norm(lam2 - lam0) <= 1e-4 * np.linalg.norm(lam0), ("rows", np.linalg.norm(lam2 - lam0))
        assert abs(elbos[24] - info0[24]["elbo"]) <= 1e-4 * abs(info0[24]["elbo"])
        gather_equal(lam2, "lambda after 25 row-sharded steps")
        str_.close(); objr.close()
        probr.close()
        ctx.close()
        if rank == 0:
            print(f"multigpu_check ok: world={world} native_exchange={native}", flush=True)
    st0.close(); st0.obj.close(); p0.close()

    # ---- full-rank family: its exchange payloads (the D x D contraction under sample sharding, the M x D gradient block
    # under row sharding) exceed the low-latency lanes and take the pull protocol (comm.cu: k_allreduce_oneshot) ----
    nF, dF, MF = 2000, 160, 128
    XF, yF = Mo.synth_glm_data(nF, dF, seed=5)
    DF = dF + 1
    rngF = np.random.default_rng(3)
    LF = np.tril(0.05 * rngF.standard_normal((DF, DF))).astype(np.float32)
    LF[np.diag_indices(DF)] = 0.4
    qF = avi.FullRankGaussian(0.1 * rngF.standard_normal(DF).astype(np.float32), LF)
    lamF = qF.destructure()
    specs = {"rep": lambda: avi.RepGradELBO(MF), "stl": lambda: avi.RepGradELBO(MF, avi.StickingTheLandingEntropy()),
             "score": lambda: avi.ScoreGradELBO(MF)}
    pF0 = avi.LogReg(c0, XF, yF, gemm="tf32")
    resF = {}
    for kind, mk in specs.items():
        o0 = avi.Objective(key, mk(), qF, pF0)
        resF[kind] = o0.estimate_gradient(lamF)
        o0.close()
    algF = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), n_samples=MF, operator=avi.ClipScale())
    _, infoF0, stF0 = avi.optimize(key, algF, 10, pF0, qF)
    lamF0, _, _ = stF0.params()
    stF0.close(); stF0.obj.close(); pF0.close(); c0.close()
    ctx = avi.Context(lr)
    parallel.connect(ctx, max_floats=4 * 192 + 64 + 2 * DF * DF + MF * 164, native=True)
    vals, elbos, nd = np.empty(10, np.float32), np.empty(10, np.float32), C.c_int32()
    for axis in ("samples", "rows"):
        if axis == "samples":
            prob = avi.LogReg(ctx, XF, yF, gemm="tf32")
            m0, ml = parallel.sample_shard(MF, rank, world)
        else:
            r0, nr = parallel.row_shard(nF, rank, world, align=32)
            prob = avi.LogReg(ctx, XF[r0:r0 + nr], yF[r0:r0 + nr], n_data=nF, gemm="tf32")
            prob.set_data_shard(world, nF, include_prior=(rank == 0))

        def shard(o):
            if axis == "samples":
                o.set_sample_shard(m0, ml)
            else:
                o.set_shard_axis(L.SHARD_ROWS)

        for kind, mk in specs.items():
            o = avi.Objective(key, mk(), qF, prob)
            shard(o)
            v, g, e = o.estimate_gradient(lamF)
            v0, g0, e0 = resF[kind]
            tol = 2e-3 if kind == "score" else 5e-5
            assert abs(v - v0) <= tol * abs(v0) + 1e-6, ("full-rank", axis, kind, v, v0)
            assert np.linalg.norm(g - g0) <= tol * np.linalg.norm(g0), ("full-rank", axis, kind, np.linalg.norm(g - g0), np.linalg.norm(g0))
            gather_equal(g, f"full-rank {axis}-sharded grad {kind}")
            o.close()
        obj = avi.Objective(key, algF.objective, qF, prob)
        shard(obj)
        st = _OptState(algF, obj, qF)
        L.check(L.lib.avi_opt_steps(st.h, 10, L.fptr(vals), L.fptr(elbos), C.byref(nd)), ctx.h)
        assert nd.value == 10
        lam1, _, _ = st.params()
        assert np.linalg.norm(lam1 - lamF0) <= 1e-4 * np.linalg.norm(lamF0), ("full-rank", axis, np.linalg.norm(lam1 - lamF0))
        assert abs(elbos[9] - infoF0[9]["elbo"]) <= 1e-4 * abs(infoF0[9]["elbo"])
        gather_equal(lam1, f"full-rank lambda after 10 {axis}-sharded steps")
        st.close(); obj.close(); prob.close()
    ctx.close()
    if rank == 0:
        print(f"multigpu_check ok: world={world} full-rank family (pull-protocol exchange)", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
