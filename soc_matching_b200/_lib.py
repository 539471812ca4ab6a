"""ctypes binding of libsocm_b200.so (the C ABI declared in include/socm_b200.h).

There is deliberately no fallback: if the shared library is missing or a CUDA device is
not available the product raises.  Build with ``python -m soc_matching_b200.build``
(or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SOCM_B200_LIB") or os.path.join(_HERE, "libsocm_b200.so")  # override: A/B runs of dev builds

c_float_p = C.POINTER(C.c_float)


class Setting(C.Structure):
    """socm_setting"""
    _fields_ = [
        ("kind", C.c_int32), ("d", C.c_int32), ("sigma_is_identity", C.c_int32), ("lmbd", C.c_float),
        ("sigma", C.c_void_p), ("sigma_inv", C.c_void_p), ("A", C.c_void_p), ("P", C.c_void_p),
        ("Q", C.c_void_p), ("omega", C.c_void_p), ("kappa", C.c_void_p), ("nu", C.c_void_p),
    ]


class UNet(C.Structure):
    """socm_unet"""
    _fields_ = [("d", C.c_int32), ("h0", C.c_int32), ("h1", C.c_int32), ("h2", C.c_int32),
                ("w", C.c_void_p * 9), ("b", C.c_void_p * 9)]


class AdamTensor(C.Structure):
    """socm_adam_tensor"""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64), ("lr", C.c_float)]


class EmaTensor(C.Structure):
    """socm_ema_tensor"""
    _fields_ = [("grad", C.c_void_p), ("ema_grad", C.c_void_p), ("n", C.c_int64)]


class WarmTable(C.Structure):
    """socm_warm_table"""
    _fields_ = [("A", C.c_void_p), ("c", C.c_void_p)]


class TabControl(C.Structure):
    """socm_tab_control"""
    _fields_ = [("kind", C.c_int32), ("A", C.c_void_p), ("c", C.c_void_p), ("ut", C.c_void_p), ("idx_t", C.c_void_p),
                ("nx", C.c_int32), ("xb", C.c_float), ("dx", C.c_float)]


# name -> (restype, argtypes); must list every symbol include/socm_b200.h declares
_vp, _i32, _u32, _u64, _i64, _f32 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_int64, C.c_float
PROTOTYPES = {
    "socm_abi_version": (C.c_int, []),
    "socm_last_error": (C.c_char_p, []),
    "socm_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "socm_set_default_engine": (C.c_int, [_i32]),
    "socm_rollout_workspace_bytes": (_i64, [C.POINTER(UNet)]),
    "socm_rollout_f32": (C.c_int, [C.POINTER(Setting), C.POINTER(UNet), C.POINTER(WarmTable), _vp, _vp, _vp,
                                   _u64, _u64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _vp]),
    "socm_rollout_tabulated_f32": (C.c_int, [C.POINTER(Setting), C.POINTER(TabControl), _vp, _vp, _vp, _u64, _u64, _i32,
                                             _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _vp]),
    "socm_philox_normal_f32": (C.c_int, [_u64, _u64, _i32, _i32, _i32, _vp, _vp]),
    "socm_unet_forward_f32": (C.c_int, [C.POINTER(UNet), _vp, _i32, _vp, _vp]),
    "socm_target_prep_f32": (C.c_int, [C.POINTER(Setting), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp,
                                       _i32, _vp, _vp]),
    "socm_target_gemm_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "socm_target_gemm_tc_workspace_bytes": (_i64, [_i32, _i32]),
    "socm_target_gemm_tc_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp]),
    "socm_target_gemm_bwd_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "socm_target_gemm_bwd_tc_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "socm_target_gemm_bwd_tc_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp]),
    "socm_target_const_m_f32": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "socm_target_grouped_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "socm_target_grouped_bwd_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "socm_target_adjoint_f32": (C.c_int, [C.POINTER(Setting), _vp, _i32, _i32, _f32, _vp, _i32, _vp]),
    "socm_adam_step_f32": (C.c_int, [C.POINTER(AdamTensor), _i32, C.c_double, C.c_double, C.c_double, _i32, _i32, _vp]),
    "socm_ema_stats_f32": (C.c_int, [C.POINTER(EmaTensor), _i32, _vp, _vp, _vp, _i32, C.c_double, C.c_double, _vp]),
    "socm_loss_workspace_bytes": (_i64, [C.POINTER(UNet), _i32, _i32]),
    "socm_unet_param_count": (_i64, [C.POINTER(UNet)]),
    "socm_unet_loss_fwdbwd_f32": (C.c_int, [C.POINTER(Setting), C.POINTER(UNet), C.POINTER(WarmTable), _vp, _vp,
                                            _vp, _i32, _vp, _vp, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _u32, _vp]),
    "socm_weight_stats_f32": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
}
# test-only entry points (not part of include/socm_b200.h)
DEBUG_PROTOTYPES = {
    "socm_debug_wgrad_tc": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp]),
    "socm_debug_wgrad_tile_bytes": (_i64, []),
    "socm_debug_wgrad_h": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
}

ROLLOUT_FORCE_GENERIC = 1
ROLLOUT_NO_TRAJ = 2
ROLLOUT_FORCE_FFMA = 4
ROLLOUT_F16 = 8
ROLLOUT_TF32 = 16
CONTROL_AFFINE, CONTROL_LOOKUP = 0, 1
LOSS_FORCE_GENERIC = 1
LOSS_FORCE_FFMA = 2
LOSS_FORCE_TC = 4
LOSS_F16 = 8
LOSS_TF32 = 16
TARGET_BWD_TF32, TARGET_BWD_F16 = 2, 4

_lib = None


class SocmError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SocmError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m soc_matching_b200.build` "
            "(there is no CPU fallback)."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(PROTOTYPES.items()) + list(DEBUG_PROTOTYPES.items()):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().socm_last_error().decode(errors="replace")
        raise SocmError(f"libsocm_b200 call failed (status {rc}): {msg}")


def ptr(t):
    """Device pointer of a contiguous fp32/fp64 CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "libsocm_b200 needs contiguous CUDA tensors"
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise SocmError(
            f"{what} lives on {t.device}; soc_matching_b200 runs on CUDA (sm_100a) only -- there is no CPU path."
        )
