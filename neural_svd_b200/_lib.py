"""ctypes binding of libnsvd.so — the thin host layer above the C-ABI (include/nsvd.h).

There is deliberately no fallback: if the library cannot be loaded, or a call fails, a
RuntimeError is raised.  Only pointers, sizes and the CUDA stream handle cross this boundary.
"""
from __future__ import annotations

import ctypes as C
import threading

from . import build as _build

c_f32p = C.POINTER(C.c_float)


class Problem(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("n_copies", C.c_int32), ("n_fourier", C.c_int32),
                ("hidden", C.c_int32), ("potential", C.c_int32), ("has_exp_mask", C.c_int32),
                ("pot_coef", C.c_float), ("scale_kinetic", C.c_float), ("op_scale", C.c_float),
                ("op_shift", C.c_float), ("sampling_sigma", C.c_float), ("hard_mul_const", C.c_float),
                ("importance", C.c_int32), ("box_mask", C.c_int32), ("pot_coef2", C.c_float),
                ("box_lim", C.c_float), ("fd_eps", C.c_float), ("ndim", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("Bff", C.c_void_p), ("W", C.c_void_p * 4), ("b", C.c_void_p * 4),
                ("mask_scales", C.c_void_p)]


class Grads(C.Structure):
    _fields_ = [("dW", C.c_void_p * 4), ("db", C.c_void_p * 4), ("dmask_scales", C.c_void_p)]


POT_HYDROGEN, POT_HARMONIC, POT_HYDROGEN_MOL_ION, POT_INFINITE_WELL, POT_COSINE = 0, 1, 2, 3, 4
IMP_GAUSSIAN, IMP_LAPLACE, IMP_UNIFORM, IMP_NONE = 0, 1, 2, 3
BOX_NONE, BOX_SQRT, BOX_EXP = 0, 1, 2
ENGINE_FP32_SIMT, ENGINE_F16X3_TC = 0, 1
CDK_FINALIZE_SCRATCH = 1024      # NSVD_CDK_FINALIZE_SCRATCH (include/nsvd.h)
# "f16x3": tcgen05 tensor cores, every fp32 operand as two fp16 planes (three bf16 products in the CDK loss);
# "bf16x3" is the round-1 name of the same engine slot and stays accepted.
ENGINES = {"fp32": ENGINE_FP32_SIMT, "fp32_simt": ENGINE_FP32_SIMT, "f16x3": ENGINE_F16X3_TC, "tc": ENGINE_F16X3_TC,
           "bf16x3": ENGINE_F16X3_TC}

_vp, _i32, _i64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
_PB, _PR, _GR = C.POINTER(Problem), C.POINTER(Params), C.POINTER(Grads)

# name -> (restype, argtypes); every symbol declared in include/nsvd.h
SIGNATURES = {
    "nsvd_abi_version": (C.c_int, []),
    "nsvd_build_hash": (C.c_char_p, []),
    "nsvd_struct_size": (C.c_size_t, [C.c_int32]),
    "nsvd_last_error": (C.c_char_p, []),
    "nsvd_launch_count": (C.c_long, []),
    "nsvd_set_tc_microbatch": (None, [C.c_int32]),
    "nsvd_profile_enable": (None, [C.c_int]),
    "nsvd_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_long), C.c_int, C.c_int]),
    "nsvd_device_ok": (C.c_int, [C.c_int]),
    "nsvd_scratch_bytes": (C.c_int, [_PB, C.c_int, C.POINTER(_sz), C.POINTER(_sz)]),
    "nsvd_fwd_streams": (C.c_int, [_PB, _PR, C.c_int, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "nsvd_gram_partials_bytes": (_sz, [_i32, _i32]),
    "nsvd_gram_reduce": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "nsvd_cross_gram": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "nsvd_loss_finalize": (C.c_int, [_vp, _vp, _i32, _i64, _i64, _i64, _vp, _vp, _vp]),
    "nsvd_loss_dF": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, _vp, _vp]),
    "nsvd_mlp_bwd": (C.c_int, [_PB, _PR, C.c_int, _vp, _vp, _vp, _sz, _GR, _vp, _sz, _vp]),
    "nsvd_cdk_work_bytes": (_sz, [_i32, _i32, _i32, C.c_int]),
    "nsvd_cdk_fwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, C.c_int, _vp, _vp, _vp, _sz, _vp]),
    "nsvd_cdk_finalize": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "nsvd_cdk_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i64, C.c_int, _vp, _vp, _vp, _sz, _i32, _vp]),
    "nsvd_cdk_offdiag": (C.c_int, [_vp, _vp, _i32, _i32, _i32, C.c_int, _vp, _vp, _sz, _i32, _vp]),
    "nsvd_rmsprop_ema_step": (C.c_int, [_i32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                        C.POINTER(_i64), C.c_float, C.c_float, C.c_float, C.c_float, _vp]),
    "nsvd_sample_gaussian": (C.c_int, [_vp, _i64, C.c_float, C.c_uint64, C.c_uint64, _vp]),
    "nsvd_sample_points": (C.c_int, [_vp, _i64, _i32, C.c_float, C.c_uint64, C.c_uint64, _vp]),
    "nsvd_linear_work_bytes": (_sz, [_i32, _i32, _i32]),
    "nsvd_linear_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, C.c_float, _vp, _sz, _vp]),
    "nsvd_linear_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, C.c_float, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nsvd_tc_gemm_selftest": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _sz, _vp]),
}

_lock = threading.Lock()
_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the tree is newer than the .so).  Raises if impossible."""
    global _lib
    with _lock:
        if _lib is None:
            path = _build.build()
            lib = C.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)        # AttributeError if a declared symbol is missing
                fn.restype, fn.argtypes = res, args
            if lib.nsvd_abi_version() != 5:
                raise RuntimeError("libnsvd.so ABI version mismatch")
            for which, cls in enumerate((Problem, Params, Grads)):
                if lib.nsvd_struct_size(which) != C.sizeof(cls):
                    raise RuntimeError(f"libnsvd.so struct layout mismatch for {cls.__name__}")
            _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().nsvd_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


KERNEL_CLASSES = ["l0_fwd", "hidden_fwd", "hidden_bwd", "l0_wgrad", "gram_reduce", "loss_dF", "prep", "head_bwd"]


def profile_read(reset=True):
    """{class: (total_ms, launches)} of the kernels timed since nsvd_profile_enable(1)."""
    lib = load()
    n = len(KERNEL_CLASSES)
    ms, cnt = (C.c_double * n)(), (C.c_long * n)()
    check(lib.nsvd_profile_read(ms, cnt, n, int(reset)), "nsvd_profile_read")
    return {k: (ms[i], cnt[i]) for i, k in enumerate(KERNEL_CLASSES)}


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
