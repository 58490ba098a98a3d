"""ctypes binding of libavi_b200.so (include/avi.h).

There is no CPU fallback: if the shared library is missing this module raises at import,
and every entry point raises `AviError` when the library reports a failure (for example
`avi_ctx_create` without a B200).
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AVI_LIB_PATH") or os.path.join(_HERE, "libavi_b200.so")   # (override: A/B runs of two builds)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C advancedvi.jl_b200/csrc`).  This package has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

# status codes / enums (include/avi.h)
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_COMM, ERR_STATE, ERR_CALLBACK = range(7)
MEANFIELD, FULLRANK, LOWRANK = 0, 1, 2
REPGRAD, SCOREGRAD = 0, 1
ENT_CLOSEDFORM, ENT_MONTECARLO, ENT_STL, ENT_CLOSEDFORM_ZEROGRAD, ENT_STL_ZEROGRAD = range(5)
RULE_DESCENT, RULE_ADAM, RULE_DOG, RULE_DOWG = range(4)
OP_IDENTITY, OP_CLIPSCALE, OP_PROXENTROPY = range(3)
AVG_NONE, AVG_POLYNOMIAL = 0, 1
GLM_BERNOULLI_LOGIT, GLM_GAUSSIAN = 0, 1
GLM_SUBSAMPLING, GLM_BASIC = 0, 1
GEMM_SIMT_FP32, GEMM_TF32, GEMM_TF32X3 = 0, 1, 2
SHARD_NONE, SHARD_SAMPLES, SHARD_ROWS = 0, 1, 2

c_float_p = C.POINTER(C.c_float)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
vp = C.c_void_p

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int32, vp, vp, C.c_int64, vp)
LOGDENSITY_FN = C.CFUNCTYPE(C.c_int32, vp, c_float_p, C.c_int32, c_float_p, c_float_p)

# every symbol include/avi.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "avi_version": (C.c_int32, []),
    "avi_last_error": (C.c_char_p, [vp]),
    "avi_ctx_create": (C.c_int32, [C.c_int32, C.POINTER(vp)]),
    "avi_ctx_destroy": (C.c_int32, [vp]),
    "avi_ctx_synchronize": (C.c_int32, [vp]),
    "avi_ctx_info": (C.c_int32, [vp, c_i32_p, c_i64_p, c_i32_p, c_i32_p]),
    "avi_ctx_launch_count": (C.c_int64, [vp]),
    "avi_ctx_set_allreduce": (C.c_int32, [vp, ALLREDUCE_FN, vp, C.c_int32, C.c_int32]),
    "avi_ctx_stream": (vp, [vp]),
    "avi_ctx_timing": (C.c_int32, [vp, C.c_int32]),
    "avi_ctx_timing_get": (C.c_int32, [vp, C.c_char_p, C.POINTER(C.c_double), c_i64_p]),
    "avi_comm_buffer": (C.c_int32, [vp, C.c_int64, C.c_char_p]),
    "avi_comm_connect": (C.c_int32, [vp, C.c_int32, C.c_int32, C.c_char_p]),
    "avi_comm_disconnect": (C.c_int32, [vp]),
    "avi_comm_barrier": (C.c_int32, [vp]),
    "avi_model_mvnormal_diag_create": (C.c_int32, [vp, c_float_p, c_float_p, C.c_int32, C.POINTER(vp)]),
    "avi_model_glm_create": (C.c_int32, [vp, c_float_p, c_float_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32,
                                         C.c_int32, C.c_int32, C.POINTER(vp)]),
    "avi_model_hostcallback_create": (C.c_int32, [vp, C.c_int32, C.c_int32, LOGDENSITY_FN, vp, C.POINTER(vp)]),
    "avi_model_subsample": (C.c_int32, [vp, c_i32_p, C.c_int64]),
    "avi_model_set_data_shard": (C.c_int32, [vp, C.c_int32, C.c_int64, C.c_int32]),
    "avi_model_dimension": (C.c_int32, [vp]),
    "avi_model_capability": (C.c_int32, [vp]),
    "avi_model_set_gemm_mode": (C.c_int32, [vp, C.c_int32]),
    "avi_model_set_fused_step": (C.c_int32, [vp, C.c_int32]),
    "avi_model_logdensity": (C.c_int32, [vp, vp, C.c_int32, C.c_int32, vp]),
    "avi_model_logdensity_and_gradient": (C.c_int32, [vp, vp, C.c_int32, C.c_int32, vp, vp]),
    "avi_model_logdensity_and_gradient_host": (C.c_int32, [vp, c_float_p, C.c_int32, c_float_p, c_float_p]),
    "avi_model_destroy": (C.c_int32, [vp]),
    "avi_obj_create": (C.c_int32, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]),
    "avi_obj_set_model": (C.c_int32, [vp, vp]),
    "avi_obj_set_base": (C.c_int32, [vp, C.c_int32, C.c_float]),
    "avi_base_constants": (C.c_int32, [C.c_int32, C.c_float, c_float_p, c_float_p]),
    "avi_check_indices": (C.c_int32, [C.POINTER(C.c_int32), C.c_int64, C.c_int64, c_i64_p]),
    "avi_obj_batch_match_samples": (C.c_int32, [vp, c_float_p, C.c_int64, C.c_int32, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p]),
    "avi_obj_seed": (C.c_int32, [vp, C.c_uint64, C.c_uint64]),
    "avi_obj_get_step": (C.c_int32, [vp, C.POINTER(C.c_uint64)]),
    "avi_obj_set_sample_shard": (C.c_int32, [vp, C.c_int32, C.c_int32]),
    "avi_obj_set_shard_axis": (C.c_int32, [vp, C.c_int32]),
    "avi_obj_num_params": (C.c_int64, [vp]),
    "avi_obj_estimate_gradient": (C.c_int32, [vp, c_float_p, C.c_int64, c_float_p, c_float_p, c_float_p]),
    "avi_obj_estimate_objective": (C.c_int32, [vp, c_float_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_uint64, c_float_p]),
    "avi_obj_rand": (C.c_int32, [vp, c_float_p, C.c_int64, c_float_p, c_float_p]),
    "avi_obj_destroy": (C.c_int32, [vp]),
    "avi_shuffle": (C.c_int32, [C.c_uint64, C.c_uint64, C.c_int64, c_i32_p]),
    "avi_opt_create": (C.c_int32, [vp, C.c_int32, c_float_p, C.c_int32, C.c_int32, C.c_float, C.c_int32,
                                   C.c_float, c_float_p, C.c_int64, C.POINTER(vp)]),
    "avi_opt_steps": (C.c_int32, [vp, C.c_int32, c_float_p, c_float_p, c_i32_p]),
    "avi_ctx_timeline_get": (C.c_int32, [vp, C.POINTER(C.c_uint64)]),
    "avi_obj_create_lowrank": (C.c_int32, [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]),
    "avi_obj_gauss_expected_grad_hess": (C.c_int32, [vp, c_float_p, C.c_int64, C.c_int32, c_float_p, c_float_p, c_float_p]),
    "avi_host_update": (C.c_int32, [C.c_int32, c_float_p, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_float, C.c_int64,
                                    C.c_int64, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p]),
    "avi_hoststep_create": (C.c_int32, [vp, C.c_int32, c_float_p, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_float,
                                        C.c_int64, c_float_p, c_float_p, c_float_p, C.POINTER(vp)]),
    "avi_hoststep_step": (C.c_int32, [vp, c_float_p, c_float_p]),
    "avi_hoststep_timing": (C.c_int32, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double)]),
    "avi_hoststep_destroy": (C.c_int32, [vp]),
    "avi_opt_steps_begin": (C.c_int32, [vp, C.c_int32]),
    "avi_opt_steps_enqueue": (C.c_int32, [vp, C.c_int32]),
    "avi_opt_steps_end": (C.c_int32, [vp, c_float_p, c_float_p, c_i32_p]),
    "avi_opt_steps_subsampled": (C.c_int32, [vp, C.c_int32, c_i32_p, C.c_int64, c_float_p, c_float_p, c_i32_p]),
    "avi_opt_get": (C.c_int32, [vp, c_float_p, c_float_p, c_float_p]),
    "avi_opt_iteration": (C.c_int64, [vp]),
    "avi_opt_state_nbytes": (C.c_int64, [vp]),
    "avi_opt_state_export": (C.c_int32, [vp, vp, C.c_int64]),
    "avi_opt_state_import": (C.c_int32, [vp, vp, C.c_int64]),
    "avi_opt_destroy": (C.c_int32, [vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)          # AttributeError here == the library does not export the ABI
    _f.restype = _res
    _f.argtypes = _args


class AviError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libavi_b200 error {code}: {msg}")
        self.code = code


def check(code, ctx=None):
    if code != OK:
        msg = lib.avi_last_error(ctx)
        raise AviError(code, msg.decode() if msg else "")


def fptr(a):
    return a.ctypes.data_as(c_float_p) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(c_i32_p) if a is not None else None
