"""Sweep of the tcgen05 kernel tiling / cluster overrides on the C2 workload in ONE process (eager launches,
AVI_NO_GRAPH=1): CUDA-event time per kernel (L2 warm) and the in-kernel phase timeline (AVI_TC_PROF).
usage: AVI_NO_GRAPH=1 AVI_TC_PROF=1 python scripts/tc_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("AVI_NO_GRAPH", "1")
import numpy as np, advancedvi_jl_b200 as avi
from oracle import models as Mo

X, y = Mo.synth_glm_data(10000, 1024, 1)
ctx = avi.Context(0); prob = avi.LogReg(ctx, X, y, gemm="tf32")
D = 1025; q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
obj = avi.Objective(1, avi.RepGradELBO(256), q, prob)
lam = q.destructure()
os.environ["AVI_TC_PROF"] = "1"
v0, g0, _ = obj.estimate_gradient(lam)
KEYS = ["AVI_TC_NT", "AVI_TC_CA", "AVI_TC_CB", "AVI_TC_BCA", "AVI_TC_BCB", "AVI_TC_PAIR", "AVI_TC_DBG"]

def run(**kw):
    for k in KEYS: os.environ.pop(k, None)
    for k, v in kw.items(): os.environ["AVI_TC_" + k] = str(v)
    sys.stderr.write(f"== {kw}\n"); sys.stderr.flush()
    try:
        obj.seed(1, 0)
        os.environ["AVI_TC_PROF"] = "1"
        v, g, _ = obj.estimate_gradient(lam)     # prints the [tc_prof] lines of this configuration
        os.environ.pop("AVI_TC_PROF")
        ctx.timing(True)
        for _ in range(10): obj.estimate_gradient(lam)
        ctx.timing(False)
        f, fc = ctx.kernel_time("glm_fwd"); b, bc = ctx.kernel_time("glm_bwd")
        ok = abs(v - v0) / abs(v0) < 1e-5 and np.linalg.norm(g - g0) / np.linalg.norm(g0) < 1e-5
        sys.stderr.write(f"   fwd {f / fc * 1e3:.2f} us  bwd {b / bc * 1e3:.2f} us  same-result {ok}\n")
    except Exception as e:   # noqa: BLE001
        sys.stderr.write(f"   FAILED {e}\n")

run()
for nt in (64, 96, 128, 144, 192, 256): run(NT=nt)
for ca, cb in ((2, 1), (1, 2), (2, 2), (1, 4), (2, 4), (1, 8)):
    run(CA=ca, CB=cb)
    run(CA=ca, CB=cb, NT=256)
run(PAIR=1); run(PAIR=1, NT=256)
for bca in (2, 4, 8): run(BCA=bca)
run(DBG=1); run(DBG=2); run(DBG=3)
