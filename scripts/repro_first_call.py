"""First call on a fresh target vs an identical second call (full C2 size, estimate_gradient! boundary): a visibility race
between the phases shows up as a mismatch on the FIRST call only (later calls find the previous call's identical data)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import advancedvi_jl_b200 as avi
n, d, M = 10000, 1024, 256
rng = np.random.default_rng(1)
X = rng.standard_normal((n, d), dtype=np.float32) / np.float32(np.sqrt(d)); X[:, d - 1] = 1.0
y = (rng.random(n) < 0.5).astype(np.float32)
D = d + 1
q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32) * 0.1)
lam = q.destructure()
ctx = avi.Context(0)
bad = 0; worst = 0.0
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for t in range(trials):
    prob = avi.LogReg(ctx, X, y, gemm="tf32")
    obj = avi.Objective(1, avi.RepGradELBO(M), q, prob)
    a = obj.estimate_gradient(lam)
    ga = a[1].copy()
    obj.seed(1, 0)
    b = obj.estimate_gradient(lam)
    obj.seed(1, 0)
    c = obj.estimate_gradient(lam)
    e1 = np.linalg.norm(ga - c[1]) / np.linalg.norm(c[1]); e2 = np.linalg.norm(b[1] - c[1]) / np.linalg.norm(c[1])
    if e1 > 0 or e2 > 0: bad += 1
    worst = max(worst, e1, e2)
    print(f"trial {t}: first-vs-third {e1:.2e}  second-vs-third {e2:.2e}  value equal {a[0] == c[0]}")
    obj.close(); prob.close()
print(f"mismatching trials {bad}/{trials}, worst {worst:.2e}")
