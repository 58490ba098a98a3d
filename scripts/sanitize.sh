#!/bin/bash
# compute-sanitizer over the hot path: memcheck + racecheck + synccheck on the smoke path and on
# one small case per kernel family (whole-iteration kernel, staged kernels, full-rank), then memcheck on the 2-rank
# parity check when the box has two GPUs.  Run under gpurun; logs land in gpurun_out/sanitize_*.log.
#   usage: scripts/sanitize.sh [memcheck|racecheck|synccheck ...]      (default: all three)
O=gpurun_out; mkdir -p $O
TOOLS=${@:-memcheck racecheck synccheck}
CS=/usr/local/cuda/bin/compute-sanitizer
cat > /tmp/sanitize_case.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
import advancedvi_jl_b200 as avi
rng = np.random.default_rng(0)
n, d, M = 300, 37, 40
X = rng.standard_normal((n, d), dtype=np.float32) / 6.0; y = (rng.random(n) < 0.5).astype(np.float32)
ctx = avi.Context(0)
D = d + 1
for fused in (2, 0):                       # whole-iteration kernel, then one kernel per stage
    prob = avi.LogReg(ctx, X, y, gemm="tf32"); prob.set_fused_step(fused)
    for q in (avi.MeanFieldGaussian(np.zeros(D, np.float32), np.full(D, 0.3, np.float32)),
              avi.FullRankGaussian(np.zeros(D, np.float32), (0.3 * np.eye(D)).astype(np.float32))):
        for ent in (avi.ClosedFormEntropy(), avi.StickingTheLandingEntropy()):
            alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), entropy=ent, n_samples=M, operator=avi.ClipScale())
            _, info, st = avi.optimize(3, alg, 4, prob, q)
            assert all(np.isfinite(i["elbo"]) for i in info)
            st.close(); st.obj.close()
    alg = avi.KLMinScoreGradDescent(optimizer=avi.DoG(1e-2), n_samples=M, operator=avi.ClipScale())
    _, info, st = avi.optimize(3, alg, 3, prob, avi.MeanFieldGaussian(np.zeros(D, np.float32), np.full(D, 0.3, np.float32)))
    st.close(); st.obj.close(); prob.close()
# low-rank family: log q based estimators through the capacitance matrix
probn = avi.MvNormalDiag(ctx, np.linspace(-1, 1, D), np.linspace(0.5, 1.5, D))
ql = avi.LowRankGaussian(np.zeros(D, np.float32), np.full(D, 0.7, np.float32), (0.1 * rng.standard_normal((D, 5))).astype(np.float32))
for spec in (avi.RepGradELBO(M, avi.StickingTheLandingEntropy()), avi.RepGradELBO(M, avi.MonteCarloEntropy()), avi.ScoreGradELBO(M)):
    o = avi.Objective(3, spec, ql, probn); o.estimate_gradient(ql.destructure()); o.close()
avi.estimate_objective(3, avi.RepGradELBO(M, avi.MonteCarloEntropy()), ql, probn)
o = avi.Objective(3, avi.RepGradELBO(M, avi.StickingTheLandingEntropyZeroGradient()), ql, probn); o.estimate_gradient(ql.destructure()); o.close()
# non-Gaussian base distributions (sampler variants, score terms), mean-field and full-rank, loop + estimate_objective
for dist in (avi.Laplace(), avi.TDist(4.0)):
    for scale in (np.full(D, 0.3, np.float32), (0.3 * np.eye(D)).astype(np.float32)):
        qb = avi.MvLocationScale(np.zeros(D, np.float32), scale, dist)
        for spec in (avi.RepGradELBO(M, avi.StickingTheLandingEntropy()), avi.ScoreGradELBO(M)):
            o = avi.Objective(3, spec, qb, probn); o.estimate_gradient(qb.destructure()); o.close()
        alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), n_samples=M, operator=avi.ClipScale())
        _, info, st = avi.optimize(3, alg, 3, probn, qb); st.close(); st.obj.close()
        avi.estimate_objective(3, avi.RepGradELBO(M, avi.MonteCarloEntropy()), qb, probn)
# sampling stages of the measure-space algorithms (Stein estimator, batch-and-match)
qf = avi.FullRankGaussian(np.zeros(D, np.float32), (0.3 * np.eye(D)).astype(np.float32))
o = avi.Objective(3, avi.RepGradELBO(8), qf, probn)
o.gaussian_expectation_gradient_and_hessian(qf, 50); o.rand_batch_match_samples_with_objective(qf, 50); o.close()
probn.close()
# minibatch loop: rows-only device gather + whole-iteration kernel, then the staged kernels (column layout rebuilt on demand)
for fused in (2, 0):
    prob = avi.LogReg(ctx, X, y, gemm="tf32"); prob.set_fused_step(fused)
    sub = avi.ReshufflingBatchSubsampling(np.arange(n), 64)
    alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-2), n_samples=M, operator=avi.ClipScale(), subsampling=sub)
    _, info, st = avi.optimize(3, alg, 6, prob, avi.MeanFieldGaussian(np.zeros(D, np.float32), np.full(D, 0.3, np.float32)))
    assert all(np.isfinite(i["elbo"]) for i in info)
    st.close(); st.obj.close(); prob.close()
print("sanitize case ok")
PY
for t in $TOOLS; do
  if [ -z "$SANITIZE_SKIP_SMOKE" ]; then
  timeout 900 $CS --tool $t --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitize_${t}_smoke.log 2>&1; echo "$t smoke rc=$?"; tail -3 $O/sanitize_${t}_smoke.log
  fi
  timeout 1500 $CS --tool $t --error-exitcode 7 python /tmp/sanitize_case.py > $O/sanitize_${t}_cases.log 2>&1; echo "$t cases rc=$?"; tail -3 $O/sanitize_${t}_cases.log
done
if [ -z "$SANITIZE_SKIP_SMOKE" ] && [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 1500 $CS --tool memcheck --target-processes all --error-exitcode 7 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port 29531 tests/multigpu_check.py > $O/sanitize_memcheck_2rank.log 2>&1; echo "memcheck 2-rank rc=$?"; tail -3 $O/sanitize_memcheck_2rank.log
fi
