"""Fused RMSprop + EMA + cosine schedule, and the on-device sampler ("next" rows).

`FusedRMSpropEMA` reproduces, in ONE kernel launch per step, what the reference loop does with
torch.optim.RMSprop (examples/utils.py:48-57), CosineAnnealingLR (operator/__init__.py:35,71-72) and
torch_ema.ExponentialMovingAverage (operator/__init__.py:36,73).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib


class FusedRMSpropEMA:
    def __init__(self, params, lr=1e-4, alpha=0.999, eps=1e-10, ema_decay=None, num_iters=None):
        self.params = [p for p in params if p.requires_grad]
        if not 1 <= len(self.params) <= 16:
            raise ValueError("FusedRMSpropEMA handles 1..16 parameter tensors")
        for p in self.params:
            if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedRMSpropEMA needs contiguous fp32 CUDA parameters (no CPU path)")
        self.lr0, self.alpha, self.eps = lr, alpha, eps
        self.ema_decay, self.num_iters = ema_decay, num_iters
        self.square_avg = [torch.zeros_like(p) for p in self.params]
        self.shadow = [p.detach().clone() for p in self.params] if ema_decay is not None else None
        self.t = 0              # scheduler steps taken
        self.num_updates = 0    # EMA updates taken
        n = len(self.params)
        self._n = n
        self._p = (C.c_void_p * n)(*[p.data_ptr() for p in self.params])
        self._sq = (C.c_void_p * n)(*[s.data_ptr() for s in self.square_avg])
        self._ema = (C.c_void_p * n)(*[s.data_ptr() for s in self.shadow]) if self.shadow is not None else None
        self._sizes = (C.c_int64 * n)(*[p.numel() for p in self.params])

    def current_lr(self):
        if self.num_iters is None:
            return self.lr0
        return 0.5 * self.lr0 * (1 + math.cos(math.pi * self.t / self.num_iters))     # CosineAnnealingLR, eta_min 0

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self):
        lib = _lib.load()
        grads = []
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("FusedRMSpropEMA.step(): a parameter has no gradient")
            grads.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
        g = (C.c_void_p * self._n)(*[t.data_ptr() for t in grads])
        w = 0.0
        if self.shadow is not None:                      # torch_ema: decay = min(decay, (1+n)/(10+n))
            self.num_updates += 1
            d = min(self.ema_decay, (1 + self.num_updates) / (10 + self.num_updates))
            w = 1.0 - d
        dev = self.params[0].device
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.nsvd_rmsprop_ema_step(self._n, self._p, g, self._sq, self._ema, self._sizes, self.current_lr(),
                                             self.alpha, self.eps, w, st), "nsvd_rmsprop_ema_step")
        self.t += 1


def _pairs(n_points: int, ndim: int) -> int:
    """the device samplers draw iid coordinate PAIRS: (n_points, ndim) coordinates take ceil(n ndim / 2) of them"""
    if ndim not in (2, 3):
        raise NotImplementedError("device samplers: ndim must be 2 or 3")
    return (n_points * ndim + 1) // 2


def sample_gaussian(n_points: int, sigma: float, seed: int, offset: int = 0, device="cuda", ndim: int = 2):
    """x (n_points, ndim) = sigma * N(0, I), generated on the device (the CPU draw of main_pde.py:92-93 remains the
    RNG-parity mode)."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("sample_gaussian has no CPU path")
    npair = _pairs(n_points, ndim)
    x = torch.empty((npair, 2), dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib.nsvd_sample_gaussian(_lib.ptr(x), npair, sigma, seed, offset, st), "nsvd_sample_gaussian")
    return x if ndim == 2 else x.reshape(-1)[:n_points * ndim].reshape(n_points, ndim)


def sample_points(n_points: int, sampling_mode: str, scale: float, seed: int, offset: int = 0, device="cuda",
                  ndim: int = 2):
    """x (n_points, ndim) from the sampler of `sampling_mode` in {'gaussian', 'laplacian', 'uniform'} (main_pde.py:89-118),
    generated on the device; pairs with GaussianImportance / LaplaceImportance / UniformImportance."""
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("sample_points has no CPU path")
    code = {"gaussian": _lib.IMP_GAUSSIAN, "laplacian": _lib.IMP_LAPLACE, "uniform": _lib.IMP_UNIFORM}[sampling_mode]
    npair = _pairs(n_points, ndim)
    x = torch.empty((npair, 2), dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib.nsvd_sample_points(_lib.ptr(x), npair, code, scale, seed, offset, st), "nsvd_sample_points")
    return x if ndim == 2 else x.reshape(-1)[:n_points * ndim].reshape(n_points, ndim)
