"""ctypes binding of liblisec_b200.so — the only way Python reaches the CUDA kernels.

This is the stub a maintainer of the reference would add next to model_training.py: plain
pointers and sizes, no torch types in any signature. Loading fails loudly when the library has not been built;
there is no Python/CPU implementation to fall back to.
"""
from __future__ import annotations

import ctypes as C
import os

LIB_PATH = os.environ.get("LISEC_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblisec_b200.so")

LISEC_OK = 0
LISEC_ERR_BAD_CONFIG = -2
LISEC_ERR_UNSUPPORTED = -6
LISEC_F32, LISEC_F64, LISEC_BF16 = 0, 1, 2
LISEC_MAX_SWEEPS = 64
ABI_VERSION = 2

STATUS_NAMES = {
    0: "LISEC_OK",
    -1: "LISEC_ERR_BAD_ARG",
    -2: "LISEC_ERR_BAD_CONFIG",
    -3: "LISEC_ERR_CAPACITY",
    -4: "LISEC_ERR_CUDA",
    -5: "LISEC_ERR_STATE",
    -6: "LISEC_ERR_UNSUPPORTED",
}


class LisecError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("%s (%d): %s" % (STATUS_NAMES.get(status, "LISEC_ERR_?"), status, message))
        self.status = status


class lisec_config(C.Structure):
    _fields_ = [
        ("voxel_x", C.c_double),
        ("voxel_y", C.c_double),
        ("voxel_z", C.c_double),
        ("sample_size", C.c_int32),
        ("max_voxel_x", C.c_int32),
        ("max_voxel_y", C.c_int32),
        ("max_voxel_z", C.c_int32),
        ("c1", C.c_int32),
        ("c2", C.c_int32),
        ("c3", C.c_int32),
        ("grid_dtype", C.c_int32),
        ("max_sweeps", C.c_int32),
        ("max_points", C.c_int64),
        ("device", C.c_int32),
        ("fcn_post_dense", C.c_int32),
    ]


_FP = C.POINTER(C.c_float)


class lisec_vfe_weights(C.Structure):
    _fields_ = [
        ("dense_kernel", _FP * 3),
        ("bn_gamma", _FP * 3),
        ("bn_beta", _FP * 3),
        ("bn_mean", _FP * 3),
        ("bn_var", _FP * 3),
        ("bn_epsilon", C.c_float),
        ("reserved", C.c_int32),
        ("post_dense_kernel", _FP * 3),
    ]


class lisec_vfe_train_params(C.Structure):  # DEVICE pointers
    _fields_ = [
        ("dense_kernel", C.c_void_p * 3),
        ("bn_gamma", C.c_void_p * 3),
        ("bn_beta", C.c_void_p * 3),
        ("moving_mean", C.c_void_p * 3),
        ("moving_var", C.c_void_p * 3),
        ("bn_epsilon", C.c_float),
        ("bn_momentum", C.c_float),
    ]


class lisec_vfe_train_grads(C.Structure):  # DEVICE pointers
    _fields_ = [("dkernel", C.c_void_p * 3), ("dgamma", C.c_void_p * 3), ("dbeta", C.c_void_p * 3)]


class lisec_conv_desc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "in_d", "in_h", "in_w", "in_c", "kd", "kh", "kw", "stride_d", "stride_hw", "pad_d", "pad_h", "pad_w",
        "out_c", "n_tiles", "shuffle", "out_pitch", "out_ch_off", "relu", "out_dtype", "tile_w", "tile_h", "m_tiles",
        "in_dtype", "out_split", "group_kh", "reserved")]


class lisec_sensor_pose(C.Structure):
    _fields_ = [("rotation", C.c_double * 9), ("translation", C.c_double * 3)]


LISEC_MAX_ANCHORS = 4


class lisec_rpn_desc(C.Structure):
    _fields_ = [("out_x", C.c_int32), ("out_y", C.c_int32), ("n_anchors", C.c_int32), ("reserved", C.c_int32),
                ("cell_x", C.c_double), ("cell_y", C.c_double), ("anchor_z", C.c_double),
                ("anchors", (C.c_double * 4) * LISEC_MAX_ANCHORS)]


class lisec_nms_desc(C.Structure):
    _fields_ = [("overlap_thresh", C.c_double), ("max_boxes", C.c_int32), ("reserved", C.c_int32),
                ("margin_x", C.c_double), ("margin_y", C.c_double), ("limit_x", C.c_double), ("limit_y", C.c_double)]


_H = C.c_void_p
_VP = C.c_void_p
_I32P = C.POINTER(C.c_int32)
_I64P = C.POINTER(C.c_int64)

# name -> (restype, argtypes); exactly the declarations of include/lisec_b200.h
SIGNATURES = {
    "lisec_abi_version": (C.c_int32, []),
    "lisec_create": (C.c_int32, [C.POINTER(lisec_config), C.POINTER(_H)]),
    "lisec_destroy": (None, [_H]),
    "lisec_last_error": (C.c_char_p, [_H]),
    "lisec_workspace_bytes": (C.c_int64, [_H]),
    "lisec_set_vfe_weights": (C.c_int32, [_H, C.POINTER(lisec_vfe_weights), _VP]),
    "lisec_get_c_empty": (C.c_int32, [_H, _FP]),
    "lisec_voxelize": (C.c_int32, [_H, _VP, C.c_int32, _I64P, C.c_int32, _VP]),
    "lisec_voxel_counts": (C.c_int32, [_H, _I32P, _I64P, _I64P, _I64P, _I64P, _VP]),
    "lisec_voxels_export": (C.c_int32, [_H, _VP, _VP, _VP, _VP, _VP]),
    "lisec_emit_dense_input": (C.c_int32, [_H, _VP, _VP]),
    "lisec_vfe_forward": (C.c_int32, [_H, _VP, _VP]),
    "lisec_scatter_dense": (C.c_int32, [_H, _VP, _VP, _VP]),
    "lisec_vfe_scatter_fused": (C.c_int32, [_H, _VP, _VP]),
    "lisec_frontend_forward": (C.c_int32, [_H, _VP, C.c_int32, _I64P, C.c_int32, _VP, _VP]),
    "lisec_frontend_forward_host": (C.c_int32, [_H, _VP, C.c_int32, _I64P, C.c_int32, _VP, _VP]),
    "lisec_voxel_counts_async": (C.c_int32, [_H, _VP, C.c_int64, _VP]),
    "lisec_last_fused_kernel_ms": (C.c_int32, [_H, _FP]),
    "lisec_debug_trace": (C.c_int32, [_H, _I64P, C.c_int64]),
    "lisec_debug_table": (C.c_int32, [_H, C.c_int32, C.POINTER(C.c_int32), C.c_int64]),
    "lisec_vfe_train_forward": (C.c_int32, [_H, C.POINTER(lisec_vfe_train_params), _VP, _VP]),
    "lisec_vfe_train_backward": (C.c_int32, [_H, C.POINTER(lisec_vfe_train_params), _VP, C.POINTER(lisec_vfe_train_grads), _VP]),
    "lisec_vfe_train_read": (C.c_int32, [_H, C.c_int32, _FP, C.c_int64, _FP, _FP]),
    "lisec_last_launch_count": (C.c_int32, [_H]),
    "lisec_conv_plan_create": (C.c_int32, [C.POINTER(lisec_conv_desc), _VP, _VP, _VP, _VP, _VP, C.POINTER(_H)]),
    "lisec_conv_plan_run": (C.c_int32, [_H, _VP]),
    "lisec_conv_plan_set_gather": (C.c_int32, [_H, _VP, _VP, _VP]),
    "lisec_workspace_pointers": (C.c_int32, [_H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "lisec_conv_plan_output_shape": (C.c_int32, [_H, _I32P]),
    "lisec_heads_combine": (C.c_int32, [_VP, _VP, C.c_int32, _VP, C.c_int32, _VP, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_int32, _VP]),
    "lisec_split_tf32": (C.c_int32, [_VP, _VP, _VP, C.c_int64, _VP]),
    "lisec_conv_plan_destroy": (None, [_H]),
    "lisec_conv_last_error": (C.c_char_p, []),
    "lisec_ingest_lidar": (C.c_int32, [_VP, C.c_int32, _I64P, C.POINTER(lisec_sensor_pose), C.c_int32, _VP, _VP, _I32P]),
    "lisec_ingest_last_error": (C.c_char_p, []),
    "lisec_rpn_decode": (C.c_int32, [C.POINTER(lisec_rpn_desc), _VP, C.c_int64, C.c_int64, _VP, C.c_int64, C.c_int64,
                                     C.c_int32, _VP, _VP, _VP]),
    "lisec_nms_rotated": (C.c_int32, [C.POINTER(lisec_nms_desc), _VP, _VP, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP, _VP]),
    "lisec_decode_last_error": (C.c_char_p, []),
    "lisec_sgd_nesterov": (C.c_int32, [_VP, _VP, _VP, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int32, _VP]),
    "lisec_mse_loss_grad": (C.c_int32, [_VP, _VP, C.c_int64, _VP, _VP, _VP]),
    "lisec_weights_flip_transpose": (C.c_int32, [_VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _VP, _VP]),
    "lisec_train_last_error": (C.c_char_p, []),
    "lisec_bn_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "lisec_bn_train_forward": (C.c_int32, [_VP, C.c_int64, C.c_int32, _VP, _VP, C.c_float, C.c_float, _VP, _VP, C.c_int32,
                                           _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "lisec_bn_train_backward": (C.c_int32, [_VP, _VP, _VP, C.c_int64, C.c_int32, _VP, _VP, _VP, C.c_int32, _VP, _VP, _VP,
                                            _VP, _VP, _VP, _VP]),
    "lisec_channel_sums": (C.c_int32, [_VP, C.c_int64, C.c_int32, _VP, _VP, _VP]),
    "lisec_relu_backward": (C.c_int32, [_VP, _VP, C.c_int64, _VP, _VP]),
    "lisec_relu_backward_f32": (C.c_int32, [_VP, _VP, C.c_int64, _VP, _VP]),
    "lisec_dilate": (C.c_int32, [_VP] + [C.c_int32] * 10 + [_VP, _VP]),
    "lisec_pad_channels_bf16": (C.c_int32, [_VP, C.c_int64, C.c_int32, C.c_int32, _VP, _VP]),
    "lisec_add_bf16": (C.c_int32, [_VP, _VP, C.c_int64, _VP]),
    "lisec_cast_f32_to_bf16": (C.c_int32, [_VP, C.c_int64, _VP, _VP]),
    "lisec_refresh_operands": (C.c_int32, [_VP, C.c_int32, C.c_int64, _VP]),
    "lisec_bn_train_backward_f32": (C.c_int32, [_VP, _VP, _VP, C.c_int64, C.c_int32, _VP, _VP, _VP, C.c_int32, _VP, _VP,
                                                _VP, _VP, _VP, _VP, _VP]),
    "lisec_bn_last_error": (C.c_char_p, []),
    "lisec_conv_wgrad_workspace_bytes": (C.c_int64, [C.POINTER(lisec_conv_desc)]),
    "lisec_conv_wgrad_plan_create": (C.c_int32, [C.POINTER(lisec_conv_desc), _VP, _VP, _VP, _VP, C.POINTER(_H)]),
    "lisec_conv_wgrad_plan_run": (C.c_int32, [_H, _VP]),
    "lisec_conv_wgrad_plan_destroy": (None, [_H]),
    "lisec_wgrad_last_error": (C.c_char_p, []),
}

_lib = None


def _preload_cudart() -> None:
    """liblisec_b200.so links libcudart.so.12 dynamically so that it shares ONE runtime instance (streams, primary
    context) with torch. If the loader cannot find it by soname, load the copy torch ships, or the toolkit's."""
    import glob
    import sys

    cands = []
    for sp in sys.path:
        cands += glob.glob(os.path.join(sp, "nvidia", "cuda_runtime", "lib", "libcudart.so.12*"))
    cands += glob.glob("/usr/local/cuda/lib64/libcudart.so.12*")
    for c in cands:
        try:
            C.CDLL(c, mode=C.RTLD_GLOBAL)
            return
        except OSError:
            continue


def load() -> C.CDLL:
    """dlopen the library and bind every symbol the header declares. Raises if it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "liblisec_b200.so is not built (%s). Run `python -m lisec_b200.build` (needs nvcc); "
            "there is no CPU fallback." % LIB_PATH
        )
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError:
        _preload_cudart()
        lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    got = lib.lisec_abi_version()
    if got != ABI_VERSION:
        raise ImportError("liblisec_b200.so has ABI version %d, binding expects %d" % (got, ABI_VERSION))
    _lib = lib
    return lib


def check(lib: C.CDLL, handle, status: int) -> None:
    if status != LISEC_OK:
        msg = lib.lisec_last_error(handle) if handle else b""
        raise LisecError(status, (msg or b"").decode("utf-8", "replace"))
