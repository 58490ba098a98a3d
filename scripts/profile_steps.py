"""Short driver for ncu: a few fused steps of the C2 workload (see bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import advancedvi_jl_b200 as avi

n, d, M = 10000, 1024, 256
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
family = sys.argv[2] if len(sys.argv) > 2 else "meanfield"
rng = np.random.default_rng(1)
X = rng.standard_normal((n, d), dtype=np.float32) / np.float32(np.sqrt(d)); X[:, d - 1] = 1.0
y = (rng.random(n) < 0.5).astype(np.float32)
ctx = avi.Context(0)
prob = avi.LogReg(ctx, X, y, gemm="tf32")
D = d + 1
if family == "meanfield":
    q0 = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
else:
    q0 = avi.FullRankGaussian(np.zeros(D, np.float32), (0.6 * np.eye(D)).astype(np.float32))
alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-3), n_samples=M, operator=avi.ClipScale())
_, info, st = avi.optimize(1, alg, steps, prob, q0)
print("elbo", info[-1]["elbo"], "launches", ctx.launch_count())
