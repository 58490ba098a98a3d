"""Runs BASELINE.json configs C3, C4a, C4b, C5 at full size on ONE GPU (the 8-GPU configs are executed on a
single device here: same arithmetic, 1/8 of the hardware) and prints one JSON line per config with the
steps/s of the fused device loop, ELBO at the first and last step, and a sanity check against the SIMT-fp32
path on a slice.  Usage: python scripts/run_configs.py [c3] [c4a] [c4b] [c5]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import advancedvi_jl_b200 as avi


def synth_fast(n, d, seed, gaussian=False):
    rng = np.random.default_rng(seed)
    X = (rng.standard_normal((n, d), dtype=np.float32) / np.float32(np.sqrt(d)))
    X[:, d - 1] = 1.0
    beta = rng.standard_normal(d).astype(np.float32)
    logits = X @ beta
    if gaussian:
        y = (logits + rng.standard_normal(n).astype(np.float32)).astype(np.float32)
    else:
        y = (rng.random(n) < 1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
    return X, y


def timed_steps(alg, prob, q0, steps, warm, seed):
    _, info_w, st = avi.optimize(seed, alg, warm, prob, q0)
    prob.ctx.synchronize()
    t0 = time.perf_counter()
    _, info, st = avi.optimize(seed, alg, steps, prob, q0, state=st)
    prob.ctx.synchronize()
    dt = time.perf_counter() - t0
    return dict(steps_per_s=steps / dt, ms_per_step=1e3 * dt / steps, elbo_first=info_w[0]["elbo"],
                elbo_last=info[-1]["elbo"], launches=prob.ctx.launch_count()), st


def main():
    which = sys.argv[1:] or ["c3", "c4a", "c4b", "c5"]
    ctx = avi.Context(0)
    for name in which:
        t_setup = time.perf_counter()
        if name == "c3":
            n, d, M = 10000, 1024, 256
            X, y = synth_fast(n, d, 1)
            prob = avi.LogReg(ctx, X, y, gemm="tf32")
            D = d + 1
            q0 = avi.FullRankGaussian(np.zeros(D, np.float32), (0.6 * np.eye(D)).astype(np.float32))
            alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-3), n_samples=M, operator=avi.ClipScale())
            desc = "C3: RepGradELBO+CFE, FullRankGaussian, logreg n=10000 d=1024 M=256"
            steps, warm = 100, 10
        elif name in ("c4a", "c4b"):
            n, d, M = 100000, 4096, 1024
            X, y = synth_fast(n, d, 2, gaussian=True)
            prob = avi.GaussGLM(ctx, X, y, gemm="tf32")
            D = d + 1
            q0 = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
            if name == "c4a":
                alg = avi.KLMinScoreGradDescent(optimizer=avi.DoG(), n_samples=M, operator=avi.ClipScale())
                desc = "C4a: ScoreGradELBO (VarGrad), MeanField, Gaussian GLM n=1e5 d=4096 M=1024, DoG"
            else:
                alg = avi.KLMinRepGradDescent(optimizer=avi.DoG(), entropy=avi.StickingTheLandingEntropy(), n_samples=M,
                                              operator=avi.ClipScale())
                desc = "C4b: RepGradELBO+StickingTheLanding, MeanField, Gaussian GLM n=1e5 d=4096 M=1024, DoG"
            steps, warm = 30, 5
        elif name == "c5":
            n, d, M, bs = 1000000, 512, 512, 4096
            X, y = synth_fast(n, d, 3)
            prob = avi.LogReg(ctx, X, y, gemm="tf32")
            D = d + 1
            q0 = avi.MeanFieldGaussian(np.zeros(D, np.float32), np.ones(D, np.float32))
            alg = avi.KLMinRepGradDescent(optimizer=avi.Adam(1e-3), n_samples=M, operator=avi.ClipScale(),
                                          subsampling=avi.ReshufflingBatchSubsampling(np.arange(n), bs))
            desc = f"C5: Subsampled RepGradELBO+CFE, MeanField, logreg n=1e6 d=512 M=512, batch {bs}, reshuffling"
            steps, warm = 300, 30
        else:
            continue
        setup_s = time.perf_counter() - t_setup
        res, st = timed_steps(alg, prob, q0, steps, warm, 1)
        res.update(config=desc, setup_s=round(setup_s, 1), steps=steps)
        # sanity: TF32 log-density / gradient vs the exact-fp32 SIMT path on 8 samples and a row slice
        ns = min(n, 20000)
        Zs = (0.1 * np.random.default_rng(0).standard_normal((D, 8))).astype(np.float32)
        cls = avi.GaussGLM if name.startswith("c4") else avi.LogReg
        pa, pb = cls(ctx, X[:ns], y[:ns], gemm="tf32"), cls(ctx, X[:ns], y[:ns], gemm="fp32")
        la, Ga = pa.logdensity_and_gradient(Zs)
        lb, Gb = pb.logdensity_and_gradient(Zs)
        res["slice_logp_rel_err_tf32_vs_fp32"] = float(np.abs(la - lb).max() / np.abs(lb).max())
        res["slice_grad_rel_err_tf32_vs_fp32"] = float(np.linalg.norm(Ga - Gb) / np.linalg.norm(Gb))
        pa.close(); pb.close()
        print(json.dumps(res), flush=True)
        st.close(); st.obj.close(); prob.close()
        del X, y


if __name__ == "__main__":
    main()
