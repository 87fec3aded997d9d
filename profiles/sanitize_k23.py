"""K2 / K3 kernel variants alone (plain shared-memory kernels) for compute-sanitizer racecheck under gpurun:
    timeout 400 compute-sanitizer --tool racecheck --print-limit 10 python profiles/sanitize_k23.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import neural_svd_b200 as N
from neural_svd_b200 import _lib
from conftest import build_problem
from oracle import nsvd_oracle as O

lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

# K2 / K3 variants through the C-ABI (ragged sizes)
for B, L, b1 in ((4096 + 77, 16, 2001), (700, 64, 351), (333, 40, 111), (300, 24, 149), (129, 33, 65)):
    F, TF = torch.randn(B, L, device="cuda"), torch.randn(B, L, device="cuda")
    v, coef = torch.rand(L, device="cuda"), torch.randn(2 * L * L + 1, device="cuda")
    terms = torch.empty(2 * L * L + 5, device="cuda")
    part = torch.empty(lib.nsvd_gram_partials_bytes(B, L), dtype=torch.uint8, device="cuda")
    dF = torch.empty_like(F)
    _lib.check(lib.nsvd_gram_reduce(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), B, L, b1, _lib.ptr(terms), _lib.ptr(part), st), "k2")
    _lib.check(lib.nsvd_loss_dF(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), _lib.ptr(coef), None, B, L, b1, B, _lib.ptr(dF), st), "k3")
    torch.cuda.synchronize()
    print("k2/k3", B, L, "ok", float(dF.abs().max()), flush=True)

