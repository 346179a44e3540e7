"""ctypes binding of libpacoh_b200.so (C ABI declared in include/pacoh_b200.h).

The product path has no CPU fallback: if the shared library has not been built (``python -m
meta_learning_pacoh_b200.build`` or ``__graft_entry__.build()``) importing this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpacoh_b200.so")

PACOH_MAX_LAYERS = 8
PACOH_OK, PACOH_ERR_INVALID, PACOH_ERR_UNSUPPORTED, PACOH_ERR_WORKSPACE, PACOH_ERR_CUDA = 0, -1, -2, -3, -4
MEAN_ZERO, MEAN_CONSTANT, MEAN_NN = 0, 1, 2
COVAR_SE, COVAR_NN = 0, 1
MAX_PEERS = 8   # PACOH_MAX_PEERS
SVGD_RBF, SVGD_IMQ = 0, 1

# every symbol include/pacoh_b200.h declares (tests/test_capi_symbols.py checks the header against this list)
EXPORTED_SYMBOLS = [
    "pacoh_abi_version", "pacoh_last_error", "pacoh_param_count", "pacoh_hyper_prior_params",
    "pacoh_workspace_bytes", "pacoh_meta_mll_fwd_bwd", "pacoh_meta_mll_fwd_bwd_ragged", "pacoh_mlp_bwd_schedule", "pacoh_logprob_finalize", "pacoh_peer_allreduce_finalize", "pacoh_svgd_workspace_bytes",
    "pacoh_svgd_phi", "pacoh_svgd_kernel_matrix", "pacoh_svgd_phi_apply", "pacoh_vi_sample", "pacoh_vi_grad", "pacoh_ffma_peak_launch", "pacoh_adam_step",
    "pacoh_stage_timing_enable", "pacoh_stage_timing_read", "pacoh_gp_forward", "pacoh_gp_forward_workspace_bytes",
    "pacoh_debug_big_layout", "pacoh_peer_allreduce_finalize_dev", "pacoh_step_prepare", "pacoh_adam_step_dev", "pacoh_adamw_step_dev",
    "pacoh_gp_posterior_workspace_bytes", "pacoh_gp_posterior", "pacoh_pred_metrics",
]


class PacohArch(ctypes.Structure):
    _fields_ = [
        ("input_dim", ctypes.c_int32), ("mean_kind", ctypes.c_int32), ("covar_kind", ctypes.c_int32),
        ("n_mean_layers", ctypes.c_int32), ("mean_layers", ctypes.c_int32 * PACOH_MAX_LAYERS),
        ("n_kernel_layers", ctypes.c_int32), ("kernel_layers", ctypes.c_int32 * PACOH_MAX_LAYERS),
        ("feature_dim", ctypes.c_int32), ("has_outputscale", ctypes.c_int32), ("noise_floor", ctypes.c_float),
    ]


class PacohError(RuntimeError):
    """A libpacoh_b200 entry point returned a negative status."""

    def __init__(self, code, message):
        super().__init__("libpacoh_b200 error %d: %s" % (code, message))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libpacoh_b200.so is not built (%s). Run `python -m meta_learning_pacoh_b200.build`; "
            "there is no CPU fallback for the PACOH hot path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    archp = ctypes.POINTER(PacohArch)
    lib.pacoh_abi_version.restype = ctypes.c_int
    lib.pacoh_abi_version.argtypes = []
    lib.pacoh_last_error.restype = ctypes.c_char_p
    lib.pacoh_last_error.argtypes = []
    lib.pacoh_param_count.restype = i64
    lib.pacoh_param_count.argtypes = [archp]
    lib.pacoh_hyper_prior_params.restype = ctypes.c_int
    lib.pacoh_hyper_prior_params.argtypes = [archp, f32, f32, vp, vp]
    lib.pacoh_workspace_bytes.restype = i64
    lib.pacoh_workspace_bytes.argtypes = [archp, i32, i32, i32]
    lib.pacoh_meta_mll_fwd_bwd.restype = ctypes.c_int
    lib.pacoh_meta_mll_fwd_bwd.argtypes = [archp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    lib.pacoh_debug_big_layout.restype = ctypes.c_int
    lib.pacoh_debug_big_layout.argtypes = [archp, i32, i32, i32, ctypes.POINTER(i64)]
    lib.pacoh_mlp_bwd_schedule.restype = ctypes.c_int
    lib.pacoh_mlp_bwd_schedule.argtypes = [i32, i32, i64, vp, vp, vp]
    lib.pacoh_meta_mll_fwd_bwd_ragged.restype = ctypes.c_int
    lib.pacoh_meta_mll_fwd_bwd_ragged.argtypes = [archp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    lib.pacoh_logprob_finalize.restype = ctypes.c_int
    lib.pacoh_logprob_finalize.argtypes = [i32, i64, vp, vp, vp, f32, f32, vp, vp, vp, vp, vp]
    lib.pacoh_peer_allreduce_finalize.restype = ctypes.c_int
    lib.pacoh_peer_allreduce_finalize.argtypes = [i32, i32, vp, vp, ctypes.c_uint32, i32, i64, vp, vp, vp, f32, f32, vp, vp, vp]
    lib.pacoh_peer_allreduce_finalize_dev.restype = ctypes.c_int
    lib.pacoh_peer_allreduce_finalize_dev.argtypes = [i32, i32, vp, vp, ctypes.c_uint32, vp, vp, i32, i64, vp, vp, vp, f32, f32, vp, vp, vp]
    lib.pacoh_step_prepare.restype = ctypes.c_int
    lib.pacoh_step_prepare.argtypes = [vp, i32, i32, vp, vp, i64, vp, vp, f32, f32, i32, f32, f32, vp]
    lib.pacoh_adam_step_dev.restype = ctypes.c_int
    lib.pacoh_adam_step_dev.argtypes = [i64, vp, vp, f32, vp, vp, f32, f32, f32, vp, vp]
    lib.pacoh_gp_posterior_workspace_bytes.restype = i64
    lib.pacoh_gp_posterior_workspace_bytes.argtypes = [archp, i32, i32, i32, i32, i32]
    lib.pacoh_gp_posterior.restype = ctypes.c_int
    lib.pacoh_gp_posterior.argtypes = [archp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    lib.pacoh_pred_metrics.restype = ctypes.c_int
    lib.pacoh_pred_metrics.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, f32, vp, vp]
    lib.pacoh_adamw_step_dev.restype = ctypes.c_int
    lib.pacoh_adamw_step_dev.argtypes = [i64, vp, vp, f32, vp, vp, f32, f32, f32, f32, vp, vp, vp]
    lib.pacoh_svgd_workspace_bytes.restype = i64
    lib.pacoh_svgd_workspace_bytes.argtypes = [i32, i64]
    lib.pacoh_svgd_phi.restype = ctypes.c_int
    lib.pacoh_svgd_phi.argtypes = [i32, i64, vp, vp, f32, i32, vp, vp, vp, i64, vp]
    lib.pacoh_svgd_kernel_matrix.restype = ctypes.c_int
    lib.pacoh_svgd_kernel_matrix.argtypes = [i32, i64, vp, f32, i32, vp, vp, i64, vp]
    lib.pacoh_svgd_phi_apply.restype = ctypes.c_int
    lib.pacoh_svgd_phi_apply.argtypes = [i32, i64, vp, vp, i32, vp, vp, vp, i64, vp]
    lib.pacoh_vi_sample.restype = ctypes.c_int
    lib.pacoh_vi_sample.argtypes = [i32, i64, vp, vp, vp, vp, vp, vp]
    lib.pacoh_vi_grad.restype = ctypes.c_int
    lib.pacoh_vi_grad.argtypes = [i32, i64, vp, vp, vp, f32, vp, vp, vp]
    lib.pacoh_gp_forward_workspace_bytes.restype = i64
    lib.pacoh_gp_forward_workspace_bytes.argtypes = [archp, i32, i32]
    lib.pacoh_gp_forward.restype = ctypes.c_int
    lib.pacoh_gp_forward.argtypes = [archp, i32, i32, vp, vp, vp, vp, vp, i64, vp]
    lib.pacoh_adam_step.restype = ctypes.c_int
    lib.pacoh_adam_step.argtypes = [i64, vp, vp, f32, vp, vp, f32, f32, f32, f32, i64, vp]
    lib.pacoh_stage_timing_enable.restype = ctypes.c_int
    lib.pacoh_stage_timing_enable.argtypes = [i32]
    lib.pacoh_stage_timing_read.restype = ctypes.c_int
    lib.pacoh_stage_timing_read.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i32)]
    lib.pacoh_ffma_peak_launch.restype = ctypes.c_int
    lib.pacoh_ffma_peak_launch.argtypes = [i32, vp, ctypes.POINTER(ctypes.c_double), vp]
    return lib


lib = _load()


def check(code):
    if code < 0:
        raise PacohError(code, lib.pacoh_last_error().decode("utf-8", "replace"))
    return code
