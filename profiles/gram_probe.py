"""Direct timing of nsvd_gram_reduce / nsvd_loss_dF through the C-ABI (CUDA events around back-to-back calls)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neural_svd_b200 as N          # noqa: E402
from neural_svd_b200 import _lib     # noqa: E402

lib = _lib.load()
B = 1 << 20
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for L in (16, 64):
    F = torch.randn(B, L, device="cuda")
    TF = torch.randn(B, L, device="cuda")
    v = torch.ones(L, device="cuda")
    terms = torch.empty(2 * L * L + 1, device="cuda")
    part = torch.empty(lib.nsvd_gram_partials_bytes(B, L), dtype=torch.uint8, device="cuda")
    for reps in (1, 20):
        for _ in range(3):
            lib.nsvd_gram_reduce(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), B, L, B // 2, _lib.ptr(terms), _lib.ptr(part), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lib.nsvd_gram_reduce(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), B, L, B // 2, _lib.ptr(terms), _lib.ptr(part), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"L={L} reps={reps}: gram {ms * 1e3:.1f} us/call = {8.0 * B * L / ms / 1e6:.0f} GB/s")
    # K3: dF = -(4/B) v TF + F_half . coef_half   (12 B L bytes)
    coef = torch.randn(2 * L * L + 1, device="cuda")
    dF = torch.empty_like(F)
    for _ in range(3):
        lib.nsvd_loss_dF(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), _lib.ptr(coef), None, B, L, B // 2, B, _lib.ptr(dF), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.nsvd_loss_dF(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), _lib.ptr(coef), None, B, L, B // 2, B, _lib.ptr(dF), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    ref = -(4.0 / B) * v * TF
    ref[:B // 2] += F[:B // 2] @ coef[:L * L].view(L, L)
    ref[B // 2:] += F[B // 2:] @ coef[L * L:2 * L * L].view(L, L)
    err = float((dF - ref).norm() / ref.norm())
    print(f"L={L}: loss_dF {ms * 1e3:.1f} us/call = {12.0 * B * L / ms / 1e6:.0f} GB/s  (rel err vs torch {err:.1e})")
