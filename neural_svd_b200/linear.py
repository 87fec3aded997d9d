"""Dense layer of the CDK encoder on the library's own tcgen05 GEMM kernels.

`TCLinear` is `torch.nn.Linear` (same parameters, same state-dict keys) whose forward / backward are
`nsvd_linear_fwd` / `nsvd_linear_bwd` (include/nsvd.h): y = act(x W^T + b) with the activation fused into the
GEMM epilogue, dx = dz W, dW = dz^T x, db = sum dz.  It replaces the Linear (+ LeakyReLU / ReLU) pairs that
`get_mlp` (examples/models/mlp.py:129-164) stacks for `HeteroNetwork` (examples/models/siam.py:132-165).
PyTorch only supplies device memory, the stream and the autograd graph edge.  There is no CPU path: a CPU
tensor raises unless the host-mirror switch (module-structure tests on the CPU) is on.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

ACT_NONE, ACT_LEAKY = 0, 1
HOST_MIRROR = False          # tests of the module structure on the CPU set this; never set by the product


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_work = {}


def _workspace(lib, rows, in_f, out_f, dev):
    """one growing scratch buffer per device (planes of the operands of ONE layer call)"""
    need = lib.nsvd_linear_work_bytes(rows, in_f, out_f)
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)     # calls on one stream are ordered: one buffer suffices
    buf = _work.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=dev)
        _work[key] = buf
    return buf


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act, slope):
        lib = _lib.load()
        dev = x.device
        x2 = x.reshape(-1, x.shape[-1]).to(torch.float32).contiguous()
        w = weight.to(torch.float32).contiguous()
        b = None if bias is None else bias.to(torch.float32).contiguous()
        rows, in_f = x2.shape
        out_f = w.shape[0]
        y = torch.empty((rows, out_f), dtype=torch.float32, device=dev)
        wk = _workspace(lib, rows, in_f, out_f, dev)
        with torch.cuda.device(dev):
            _lib.check(lib.nsvd_linear_fwd(_lib.ptr(x2), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), rows, in_f, out_f, act,
                                           float(slope), _lib.ptr(wk), wk.numel(), _stream(dev)), "nsvd_linear_fwd")
        ctx.save_for_backward(x2, w, y)
        ctx.meta = (act, float(slope), bias is not None, x.shape)
        return y.reshape(*x.shape[:-1], out_f)

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x2, w, y = ctx.saved_tensors
        act, slope, has_bias, xshape = ctx.meta
        dev = gy.device
        rows, in_f = x2.shape
        out_f = w.shape[0]
        gy2 = gy.reshape(rows, out_f).to(torch.float32).contiguous()
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = torch.empty_like(x2) if need_x else None
        dw = torch.empty_like(w) if need_w else None
        db = torch.empty(out_f, dtype=torch.float32, device=dev) if (has_bias and ctx.needs_input_grad[2]) else None
        wk = _workspace(lib, rows, in_f, out_f, dev)
        with torch.cuda.device(dev):
            _lib.check(lib.nsvd_linear_bwd(_lib.ptr(x2), _lib.ptr(w), _lib.ptr(y), _lib.ptr(gy2), rows, in_f, out_f, act,
                                           slope, _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), _lib.ptr(wk), wk.numel(),
                                           _stream(dev)), "nsvd_linear_bwd")
        return (None if dx is None else dx.reshape(xshape)), dw, db, None, None


class TCLinear(nn.Linear):
    """nn.Linear on the hand-written GEMM; `fused_act` = None | ('leaky', slope) is applied in the GEMM epilogue
    (the activation module that follows in the Sequential is then a `FusedActivation` placeholder)."""

    def __init__(self, in_features, out_features, bias=True, fused_act=None):
        super().__init__(in_features, out_features, bias=bias)
        self.fused_act = fused_act

    def forward(self, x):
        act, slope = (ACT_NONE, 0.0) if self.fused_act is None else (ACT_LEAKY, float(self.fused_act[1]))
        if x.device.type != "cuda":
            if not HOST_MIRROR:
                raise RuntimeError("neural_svd_b200 has no CPU path: TCLinear needs CUDA (sm_100) tensors")
            y = F.linear(x, self.weight, self.bias)
            return F.leaky_relu(y, slope) if act == ACT_LEAKY else y
        with torch.autocast("cuda", enabled=False):      # main_sketchy.py:182 wraps the model in autocast: fp32 here
            return _LinearFn.apply(x, self.weight, self.bias, act, slope)


class FusedActivation(nn.Module):
    """Keeps the position (and so the state-dict indices) of the activation module of the reference's Sequential; the
    activation itself ran in the epilogue of the preceding TCLinear."""

    def __init__(self, name):
        super().__init__()
        self.name = name

    def extra_repr(self):
        return f"{self.name} (fused into the preceding TCLinear)"

    def forward(self, x):
        return x
