"""Per-CTA phase stamps (AVI_STEP_PROF=1) of the gradient-store variant of the persistent kernel (forward + backward of the
GLM target for the full-rank iteration, C3): AVI_STEP_PROF=1 python scripts/step_prof_fr.py [steps]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("AVI_STEP_PROF", "1")
os.environ.setdefault("AVI_NO_GRAPH", "1")
import numpy as np
import advancedvi_jl_b200 as avi
from advancedvi_jl_b200 import _lib as L

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n, d, M = 10000, 1024, 256
rng = np.random.default_rng(1)
X = rng.standard_normal((n, d), dtype=np.float32) / 32.0
y = (rng.random(n) < 0.5).astype(np.float32)
ctx = avi.Context(0); prob = avi.LogReg(ctx, X, y, gemm="tf32")
D = d + 1
q0 = avi.FullRankGaussian(np.zeros(D, np.float32), (0.6 * np.eye(D)).astype(np.float32))
alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-3), n_samples=M, operator=avi.ClipScale())
_, info, st = avi.optimize(1, alg, steps, prob, q0)
buf = np.zeros(160 * 32, np.uint64)
fn = L.lib.avi_step_fused_prof_get
fn.restype = C.c_int32
fn.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int32]
grid = fn(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), 160)
h = buf[:grid * 32].reshape(grid, 32).astype(np.int64)
t0 = h[:, 0][h[:, 0] > 0].min()
names = {0: "entry", 1: "prologue done", 2: "past dependency wait", 3: "snapshot read", 5: "past barrier 0",
         6: "fwd: last operand request", 7: "fwd: first operands landed", 8: "fwd: last MMA issued", 9: "fwd: epilogue math done",
         10: "ring re-carved for the backward phase", 11: "fwd: unit complete", 12: "arrive barrier 1", 13: "past barrier 1",
         14: "bwd: last operand request", 15: "bwd: first operands landed", 16: "bwd: last MMA issued", 17: "bwd: epilogue math done",
         19: "bwd: unit complete", 23: "exit"}
print(f"# gradient-store variant, grid {grid}: ns since the first CTA's entry (min / mean / max over CTAs that stamped)")
for k in sorted(names):
    v = h[:, k][h[:, k] > 0] - t0
    if len(v):
        print(f"{k:2d} {names[k]:32s} {v.min():7d} {int(v.mean()):7d} {v.max():7d}   ({len(v)} CTAs)")
