"""Device timeline of the iteration (C2 workload by default, captured graph): where the microseconds of one step go.
    AVI_TIMELINE=1 python scripts/step_timeline.py [steps] [rows] [fused 0|1|2]
Single-kernel path (csrc/step_fused.cu): per step, ns relative to the first CTA's entry, of
    first CTA past the dependency wait | forward starts (past barrier 0) | backward starts | tail starts
    last CTA leaves the sample phase | the forward phase | the backward phase | last CTA done
Staged path (AVI_FUSED_STEP=0): each kernel's [first CTA entered, first CTA past its dependency wait, last CTA done].
`rows` < 10000 times a row shard of the workload (what one rank of an n-axis run holds)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("AVI_TIMELINE", "1")
import numpy as np
import advancedvi_jl_b200 as avi
from advancedvi_jl_b200 import _lib as L
from advancedvi_jl_b200.api import _OptState

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 48
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
fused = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rng = np.random.default_rng(1)
X = rng.standard_normal((rows, 1024), dtype=np.float32) / 32.0
y = (rng.random(rows) < 0.5).astype(np.float32)
ctx = avi.Context(0); prob = avi.LogReg(ctx, X, y, n_data=10000, gemm="tf32")
prob.set_fused_step(fused)
D = 1025; q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-3), n_samples=256, operator=avi.ClipScale())
obj = avi.Objective(1, alg.objective, q, prob)
st = _OptState(alg, obj, q)
st.steps_begin(steps); st.steps_enqueue(steps); _, _, done = st.steps_end()
hist = np.zeros(64 * 32, np.uint64)
L.check(L.lib.avi_ctx_timeline_get(ctx.h, hist.ctypes.data_as(C.POINTER(C.c_uint64))), ctx.h)
hist = hist.reshape(64, 32).astype(np.int64)
prev0 = None
print(f"# rows {rows}, fused mode {fused}")
for s in range(done - 6, done):
    h = hist[s % 64]
    t0 = h[0]
    if fused and h[11] > 0:
        line = (f"step {s}: past-wait {h[1] - t0:6d} | sample done {h[8] - t0:6d} | fwd starts {h[2] - t0:6d} fwd done {h[9] - t0:6d} | "
                f"bwd starts {h[3] - t0:6d} bwd done {h[10] - t0:6d} | tail starts {h[4] - t0:6d} all done {h[11] - t0:6d}  ")
    else:
        line = f"step {s}: "
        for k, nm in enumerate(["sample", "fwd", "bwd", "tail"]):
            line += f"{nm} [{h[k] - t0:6d} {h[4 + k] - t0:6d} {h[8 + k] - t0:6d}]  "
    if prev0 is not None:
        line += f"period {t0 - prev0} ns"
    prev0 = t0
    print(line)
