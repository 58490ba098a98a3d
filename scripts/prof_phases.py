"""Prints the in-kernel phase timeline of the tcgen05 kernels (AVI_TC_PROF=n skips the first n launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, advancedvi_jl_b200 as avi
from oracle import models as Mo
X, y = Mo.synth_glm_data(10000, 1024, 1)
ctx = avi.Context(0); prob = avi.LogReg(ctx, X, y, gemm="tf32")
D = 1025; q = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
obj = avi.Objective(1, avi.RepGradELBO(256), q, prob)
for i in range(6): obj.estimate_gradient(q.destructure())
