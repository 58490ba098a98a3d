"""World-size-2 checks of the multi-rank host logic under gloo on CPU (no GPU, no compute calls into the
library): shard bounds, key agreement, per-rank minibatches, and that the ONE exchange of the path -- a
sum-all-reduce of the partial gradient sums of SURVEY.md 8e -- reproduces the unsharded oracle gradient for
both sharding axes."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from advancedvi_jl_b200 import parallel
    from oracle import family as F, models as Mo, objectives as O, philox as P
    res = {}
    # key agreement
    res["key"] = parallel.broadcast_key(1234 + 77 * rank)
    # M-axis: partial [sum g, sum g*eps, sum logp] over the local samples, one all-reduce, replicated finalize
    n, d, M = 60, 7, 10
    D = d + 1
    X, y = Mo.synth_glm_data(n, d, seed=5)
    prob = Mo.LogReg(X, y)
    q = F.MeanFieldGaussian(0.1 * np.arange(D), 0.5 + 0.05 * np.arange(D))
    eps = P.normal_matrix(3, 0, D, M)
    m0, ml = parallel.sample_shard(M, rank, world)
    Z = q.rand_from_eps(eps[:, m0:m0 + ml])
    lp, G = prob.logdensity_and_gradient_batch(Z)
    acc = np.concatenate([G.sum(1), (G * eps[:, m0:m0 + ml]).sum(1), [lp.sum()]])
    t = torch.from_numpy(acc.copy())
    dist.all_reduce(t)
    acc = t.numpy()
    g = np.concatenate([-acc[:D] / M, -acc[D:2 * D] / M - 1.0 / q.scale])
    val = -(acc[2 * D] / M + q.entropy())
    vo, go, _ = O.repgrad_value_and_gradient(q.destructure(), q, prob, eps, "ClosedFormEntropy")
    res["m_axis"] = (float(abs(val - vo)), float(np.abs(g - go).max()))
    # n-axis: every rank holds all samples and a row slice; the prior is counted on rank 0 only
    r0, nr = parallel.row_shard(n, rank, world, align=4)
    sub = Mo.LogReg(X[r0:r0 + nr], y[r0:r0 + nr], n_data=n)
    Zf = q.rand_from_eps(eps)
    B, eta = Zf[:d], Zf[d]
    logits = sub.X @ B
    ll, resid = sub._loglik_and_resid(logits)
    w = n / n                      # n_data / rows_global
    lp_part, G_part = w * ll, np.zeros_like(Zf)
    G_part[:d] = w * (sub.X.T @ resid)
    if rank == 0:
        # prior terms only: full log-density minus its likelihood part
        lpf, Gf = prob.logdensity_and_gradient_batch(Zf)
        llf, resf = prob._loglik_and_resid(prob.X @ B)
        G0 = Gf.copy()
        G0[:d] -= prob.X.T @ resf
        lp_part, G_part = lp_part + (lpf - llf), G_part + G0
    t = torch.from_numpy(np.concatenate([lp_part, G_part.reshape(-1)]))
    dist.all_reduce(t)
    lp_all, G_all = t.numpy()[:M], t.numpy()[M:].reshape(D, M)
    lpo, Go = prob.logdensity_and_gradient_batch(Zf)
    res["n_axis"] = (float(np.abs(lp_all - lpo).max()), float(np.abs(G_all - Go).max()))
    # weak-scaling minibatches: disjoint, equal-sized, cover the usable part of the epoch
    batches = [np.arange(8 * k, 8 * k + 8) for k in range(7)]
    mine = parallel.rank_batches(batches, rank, world)
    res["batches"] = [b.tolist() for b in mine]
    out.put((rank, res))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[0]["key"] == got[1]["key"] == 1234
    for r in range(2):
        assert max(got[r]["m_axis"]) < 1e-10
        assert max(got[r]["n_axis"]) < 1e-10
    b0, b1 = got[0]["batches"], got[1]["batches"]
    assert len(b0) == len(b1) == 3
    flat = sorted(sum(b0 + b1, []))
    assert flat == list(range(48))


def test_shard_ranges():
    sys.path.insert(0, ROOT)
    from advancedvi_jl_b200 import parallel
    for total in (0, 1, 7, 256, 10000):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(l for _, l in spans) == total
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(l for _, l in spans) - min(l for _, l in spans) <= 1
    assert parallel.sample_shard(256, 3, 8) == (96, 32)
    spans = [parallel.row_shard(10000, r, 8, align=32) for r in range(8)]
    assert all(s[0] % 32 == 0 for s in spans) and sum(s[1] for s in spans) == 10000
