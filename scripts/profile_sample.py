"""Driver for ncu / timing: the sample+transform kernel at a bandwidth-relevant shape (M = 32768, D = 1025)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import advancedvi_jl_b200 as avi

D = int(sys.argv[2]) if len(sys.argv) > 2 else 1025
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ctx = avi.Context(0)
q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
p = avi.MvNormalDiag(ctx, np.zeros(D, np.float32), np.ones(D, np.float32))
o = avi.Objective(1, avi.RepGradELBO(8), q, p)
o.estimate_objective(1, q, M)
ctx.timing(True)
for _ in range(5):
    o.estimate_objective(1, q, M)
ctx.timing(False)
ms, cnt = ctx.kernel_time("sample")
ld = (D + 3) // 4 * 4
b = 4 * (2 * D + 2 * ld * M)
print(f"sample M={M} D={D}: {ms / cnt * 1e3:.1f} us/launch, {b / (ms / cnt * 1e-3) / 1e9:.0f} GB/s")
