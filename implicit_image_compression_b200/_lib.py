"""ctypes binding of libsirenb200.so (C ABI: include/siren_b200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 GPU is present, every compute entry
point raises.  Build the library with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsirenb200.so")

PREC_FP32 = 0
PREC_F16TC = 1

# every symbol include/siren_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "sirenb200_version", "sirenb200_last_error", "sirenb200_create", "sirenb200_destroy",
    "sirenb200_workspace_bytes", "sirenb200_set_grid_lut", "sirenb200_set_grid_coords",
    "sirenb200_forward", "sirenb200_forward_backward", "sirenb200_backward", "sirenb200_eval_metrics",
    "sirenb200_adam_step", "sirenb200_apply_mask", "sirenb200_kmeans_quantize",
    "sirenb200_prune_threshold_search",
    "sirenb200_fakequant_per_channel", "sirenb200_launch_count", "sirenb200_profile_enable",
    "sirenb200_profile_read", "sirenb200_debug_timeline", "sirenb200_sched_step", "sirenb200_adam_step_dev",
    "sirenb200_comm_create", "sirenb200_comm_handle", "sirenb200_comm_connect", "sirenb200_comm_allreduce",
    "sirenb200_comm_destroy", "sirenb200_fit_steps", "sirenb200_fit_step", "sirenb200_set_act_quant",
    "sirenb200_fakequant_per_tensor", "sirenb200_set_fourier_encoding", "sirenb200_pack_stream",
]

PROFILE_KINDS = ["weight_staging", "first_layer", "fwd_gemm", "last_layer_loss", "dx_gemm", "dw_gemm",
                 "layer0_grad", "reduce_partials", "finalize"]


class Config(ctypes.Structure):
    _fields_ = [
        ("depth", c_int32), ("hidden", c_int32), ("in_features", c_int32), ("out_features", c_int32),
        ("first_omega", c_float), ("hidden_omega", c_float), ("outermost_linear", c_int32),
        ("height", c_int32), ("width", c_int32), ("row_begin", c_int32), ("row_end", c_int32),
        ("precision", c_int32), ("reserved", c_int32 * 4),
    ]


class FitArgs(ctypes.Structure):
    """sirenb200_fit_t (include/siren_b200.h)."""
    _fields_ = [
        ("n_tensors", c_int32), ("h_params", POINTER(c_void_p)), ("h_grads", POINTER(c_void_p)),
        ("h_exp_avg", POINTER(c_void_p)), ("h_exp_avg_sq", POINTER(c_void_p)), ("h_mask", POINTER(c_void_p)),
        ("h_numel", POINTER(c_int64)), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
        ("sched_state", c_void_p), ("stats", c_void_p), ("loss_ring", c_void_p), ("ring_len", c_int32),
        ("loss_host", c_void_p), ("comm", c_void_p), ("flat", c_void_p), ("flat_n", c_int64),
        ("inv_count", c_float),
    ]


class SirenB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises SirenB200Error if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SirenB200Error(
            f"{LIB_PATH} not found: the CUDA extension is not built (run __graft_entry__.build()); "
            "siren-b200 has no CPU or PyTorch fallback")
    lib = ctypes.CDLL(LIB_PATH)
    vp, vpp = c_void_p, POINTER(c_void_p)
    lib.sirenb200_version.restype = c_int32
    lib.sirenb200_last_error.restype = c_char_p
    lib.sirenb200_launch_count.restype = c_int64
    lib.sirenb200_create.argtypes = [POINTER(Config), POINTER(vp)]
    lib.sirenb200_destroy.argtypes = [vp]
    lib.sirenb200_workspace_bytes.argtypes = [vp]
    lib.sirenb200_workspace_bytes.restype = c_int64
    lib.sirenb200_set_grid_lut.argtypes = [vp, vp, vp]
    lib.sirenb200_set_grid_coords.argtypes = [vp, vp]
    lib.sirenb200_forward.argtypes = [vp, vpp, vp, vp]
    lib.sirenb200_forward_backward.argtypes = [vp, vpp, vp, c_float, vpp, vp, vp]
    lib.sirenb200_backward.argtypes = [vp, vpp, vp, vpp, vp]
    lib.sirenb200_eval_metrics.argtypes = [vp, vp, c_int64, vp, vp]
    lib.sirenb200_adam_step.argtypes = [c_int32, vpp, vpp, vpp, vpp, vpp, POINTER(c_int64), c_float,
                                        c_float, c_float, c_float, c_int32, c_float, vp, c_int32, vp]
    lib.sirenb200_apply_mask.argtypes = [vp, vp, c_int64, vp]
    lib.sirenb200_sched_step.argtypes = [vp, vp, c_float, vp, c_int32, vp, vp]
    lib.sirenb200_adam_step_dev.argtypes = [c_int32, vpp, vpp, vpp, vpp, vpp, POINTER(c_int64), c_float,
                                            c_float, c_float, vp, c_float, vp, c_int32, vp]
    lib.sirenb200_comm_create.argtypes = [c_int32, c_int32, c_int64, vpp]
    lib.sirenb200_comm_handle.argtypes = [vp, vp]
    lib.sirenb200_comm_connect.argtypes = [vp, vp]
    lib.sirenb200_comm_allreduce.argtypes = [vp, vp, c_int64, vp]
    lib.sirenb200_comm_destroy.argtypes = [vp]
    lib.sirenb200_fit_steps.argtypes = [vp, c_int32, vp, POINTER(FitArgs), vp]
    lib.sirenb200_fit_step.argtypes = [vp, vp, POINTER(FitArgs), vp]
    lib.sirenb200_set_fourier_encoding.argtypes = [vp, vp]
    lib.sirenb200_pack_stream.argtypes = [c_int32, vpp, vpp, POINTER(c_int32), POINTER(c_int64), POINTER(c_int64),
                                          POINTER(c_int64), vp, vp]
    lib.sirenb200_set_act_quant.argtypes = [vp, vp, c_int32, c_int32, c_float, c_int32, c_int32]
    lib.sirenb200_fakequant_per_tensor.argtypes = [vp, c_int64, vp, c_int32, c_float, c_int32, c_int32, vp, vp, vp]
    lib.sirenb200_kmeans_quantize.argtypes = [vp, c_int64, c_int32, c_int32, c_float, vp, vp, vp, vp,
                                              vp, vp]
    lib.sirenb200_prune_threshold_search.argtypes = [vp, c_int64, c_int64, c_int64, ctypes.c_double, vp, vp, vp]
    lib.sirenb200_fakequant_per_channel.argtypes = [vp, c_int32, c_int32, vp, vp, c_float, c_float, vp, vp,
                                                    vp, vp]
    lib.sirenb200_debug_timeline.argtypes = [vp, POINTER(c_int64), c_int32]
    lib.sirenb200_profile_enable.argtypes = [vp, c_int32]
    lib.sirenb200_profile_read.argtypes = [vp, POINTER(c_float), POINTER(c_int32), c_int32]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().sirenb200_last_error().decode("utf-8", "replace")
        raise SirenB200Error(f"libsirenb200 error {rc}: {msg}")


def ptr_array(tensors):
    """Host array of device pointers (None -> NULL)."""
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def launch_count():
    return int(load().sirenb200_launch_count())


def require_cuda(t, what):
    if not t.is_cuda:
        raise SirenB200Error(
            f"{what} is on {t.device}: siren-b200 runs on CUDA (sm_100a) only and has no CPU fallback")
