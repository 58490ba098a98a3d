"""Per-CTA phase stamps of the whole-iteration kernel (AVI_STEP_PROF=1): which phase is the limiter.
    AVI_STEP_PROF=1 python scripts/step_prof.py [rows] [steps]
Prints, per stamp, min / mean / max over the CTAs in ns since the first CTA's entry, for the LAST launch."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("AVI_STEP_PROF", "1")
os.environ.setdefault("AVI_NO_GRAPH", "1")     # eager launches: the stamp buffer is reset before every launch
import numpy as np
import advancedvi_jl_b200 as avi
from advancedvi_jl_b200 import _lib as L
from advancedvi_jl_b200.api import _OptState

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
cold = len(sys.argv) > 3 and sys.argv[3] == "cold"      # flush L2 (256 MiB write + read) before the profiled launch
rng = np.random.default_rng(1)
X = rng.standard_normal((rows, 1024), dtype=np.float32) / 32.0
y = (rng.random(rows) < 0.5).astype(np.float32)
ctx = avi.Context(0); prob = avi.LogReg(ctx, X, y, n_data=10000, gemm="tf32")
D = 1025; q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-3), n_samples=256, operator=avi.ClipScale())
obj = avi.Objective(1, alg.objective, q, prob)
st = _OptState(alg, obj, q)
if cold:
    import torch
    ext = torch.cuda.ExternalStream(ctx.stream(), device=0)
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    st.steps_begin(steps)
    for k in range(steps):
        with torch.cuda.stream(ext):
            flush.zero_(); flush.sum()
        st.steps_enqueue(1)
    st.steps_end()
else:
    st.steps_begin(steps); st.steps_enqueue(steps); st.steps_end()
buf = np.zeros(160 * 32, np.uint64)
fn = L.lib.avi_step_fused_prof_get
fn.restype = C.c_int32
fn.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int32]
grid = fn(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), 160)
h = buf[:grid * 32].reshape(grid, 32).astype(np.int64)
t0 = h[:, 0][h[:, 0] > 0].min()
names = {0: "entry", 1: "prologue done", 2: "past dependency wait", 3: "snapshot read", 4: "sample phase done", 5: "past barrier 0",
         6: "fwd: last operand request", 7: "fwd: first operands landed", 8: "fwd: last MMA issued", 9: "fwd: epilogue math done",
         10: "ring re-carved for the backward phase", 11: "fwd: unit complete", 12: "arrive barrier 1", 13: "past barrier 1",
         14: "bwd: last operand request", 15: "bwd: first operands landed", 16: "bwd: last MMA issued", 17: "bwd: epilogue math done",
         18: "bwd: slab rows stored", 19: "bwd: unit complete (+combine)", 20: "arrive barrier 2", 21: "past barrier 2",
         24: "tail: inputs loaded, scalars reduced", 25: "tail: value + gradient of the slice", 26: "tail: update stored",
         22: "tail done (next samples drawn)", 23: "exit"}
print(f"# {'COLD (L2 flushed before the launch)' if cold else 'warm'}")
print(f"# rows {rows}, grid {grid}: ns since the first CTA's entry (min / mean / max over CTAs that stamped)")
for k in sorted(names):
    v = h[:, k][h[:, k] > 0] - t0
    if len(v):
        print(f"{k:2d} {names[k]:32s} {v.min():7d} {int(v.mean()):7d} {v.max():7d}   ({len(v)} CTAs)")
