"""Device timeline of the fused iteration (C2 workload, captured graph): where the microseconds of one step go.
    AVI_TIMELINE=1 python scripts/step_timeline.py [steps]
Prints, for the last few steps, each kernel's [first CTA entered, first CTA past its dependency wait, last CTA done]
in ns relative to the step's first stamp, and the step period."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("AVI_TIMELINE", "1")
import numpy as np
import advancedvi_jl_b200 as avi
from advancedvi_jl_b200 import _lib as L
from advancedvi_jl_b200.api import _OptState
from oracle import models as Mo

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 48
X, y = Mo.synth_glm_data(10000, 1024, 1)
ctx = avi.Context(0); prob = avi.LogReg(ctx, X, y, gemm="tf32")
D = 1025; q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-3), n_samples=256, operator=avi.ClipScale())
obj = avi.Objective(1, alg.objective, q, prob)
st = _OptState(alg, obj, q)
st.steps_begin(steps); st.steps_enqueue(steps); _, _, done = st.steps_end()
hist = np.zeros(64 * 32, np.uint64)
L.check(L.lib.avi_ctx_timeline_get(ctx.h, hist.ctypes.data_as(C.POINTER(C.c_uint64))), ctx.h)
hist = hist.reshape(64, 32).astype(np.int64)
names = ["sample", "fwd", "bwd", "tail"]
prev0 = None
for s in range(done - 6, done):
    h = hist[s % 64]
    t0 = h[0]
    line = f"step {s}: "
    for k, nm in enumerate(names):
        line += f"{nm} [{h[k] - t0:6d} {h[4 + k] - t0:6d} {h[8 + k] - t0:6d}]  "
    if 0 < h[16] < (1 << 62):
        line += f"tail phases: loaded+logdet {h[16] - t0}, value {h[17] - t0}, updated {h[18] - t0}  "
    if prev0 is not None:
        line += f"period {t0 - prev0} ns"
    prev0 = t0
    print(line)
