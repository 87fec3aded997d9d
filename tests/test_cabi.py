"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/nsvd.h declares (CPU)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT
from neural_svd_b200 import _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "nsvd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nsvd_[A-Za-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.nsvd_abi_version() == 5
    import ctypes as C
    assert [lib.nsvd_struct_size(i) for i in range(3)] == [C.sizeof(_lib.Problem), C.sizeof(_lib.Params), C.sizeof(_lib.Grads)]


def test_library_is_sm100a_and_has_no_torch_dependency():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "c10" not in ldd


def test_argument_errors_are_reported_not_crashed():
    lib = _lib.load()
    pb = _lib.Problem(n_points=8, n_copies=100, n_fourier=64, hidden=128, potential=0, has_exp_mask=0,
                      pot_coef=1, scale_kinetic=1, op_scale=1, op_shift=0, sampling_sigma=1, hard_mul_const=1)
    a, b = ctypes.c_size_t(), ctypes.c_size_t()
    rc = lib.nsvd_scratch_bytes(ctypes.byref(pb), 0, ctypes.byref(a), ctypes.byref(b))
    assert rc == 10001 and b"n_copies" in lib.nsvd_last_error()
    pb.n_copies, pb.hidden = 4, 64
    assert lib.nsvd_scratch_bytes(ctypes.byref(pb), 0, ctypes.byref(a), ctypes.byref(b)) == 10001
    pb.hidden = 128
    assert lib.nsvd_scratch_bytes(ctypes.byref(pb), 0, ctypes.byref(a), ctypes.byref(b)) == 0
    assert a.value > 0 and b.value > 0
    assert lib.nsvd_scratch_bytes(ctypes.byref(pb), 7, ctypes.byref(a), ctypes.byref(b)) == 10001
