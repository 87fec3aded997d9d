"""Small invocations of every kernel family touched in round 2 (for compute-sanitizer memcheck under gpurun):
    timeout 400 compute-sanitizer --tool memcheck --print-limit 10 python profiles/sanitize_small.py"""
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

# fused step, tensor-core engine: exact and finite-difference mode, ragged batch over two micro-batches
N.set_engine("f16x3")
N.set_microbatch(256)
for eps in (0.0, 0.01):
    cfg = O.PathConfig.hydrogen(neigs=4, fourier_mapping_size=64)
    method, operator, importance, _ = build_problem(cfg, 0, "cuda", laplacian_eps=eps)
    x = (cfg.sampling_scale * torch.randn(300, 2)).cuda()
    loss, _ = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
    torch.cuda.synchronize()
    print("step eps", eps, "ok", float(loss.detach()), flush=True)
N.set_microbatch(65536)

# CDK loss (with diagnostics) and one dense layer pair of the encoder
f = torch.randn(96, 40, device="cuda", requires_grad=True)
g = torch.randn(96, 40, device="cuda", requires_grad=True)
m = N.NestedLoRAForCDK(None, 40)
out = m.compute_loss(f, g)
out[0].backward()
torch.cuda.synchronize()
print("cdk ok", float(out[0].detach()), flush=True)
from neural_svd_b200.linear import TCLinear
lin = TCLinear(64, 136, fused_act=("leaky", 0.2)).cuda()
xx = torch.randn(50, 64, device="cuda", requires_grad=True)
y = lin(xx)
y.square().sum().backward()
torch.cuda.synchronize()
print("linear ok", float(y.abs().max().detach()), flush=True)
