"""Host-side mirror of `methods/nestedlora.py` (the hot path's public API), on the sm_100a kernels.

Same names, arguments, return values and error behaviour as the reference:
  get_joint_nesting_masks / get_sequential_nesting_masks   nestedlora.py:40-54
  NestedLoRALossFunctionEVD                                nestedlora.py:67-111
  NestedLoRA (.compute_loss_operator, .forward, masks)     nestedlora.py:167-267
  NestedLoRALossFunctionForCDK / NestedLoRAForCDK          nestedlora.py:270-378
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib, fused


def get_joint_nesting_masks(weights: np.ndarray, set_first_mode_const: bool = False):
    vector_mask = list(np.cumsum(list(weights)[::-1])[::-1])
    if set_first_mode_const:
        vector_mask = [vector_mask[0]] + vector_mask
    vector_mask = torch.tensor(np.array(vector_mask)).float()
    matrix_mask = torch.minimum(vector_mask.unsqueeze(1), vector_mask.unsqueeze(1).T).float()
    return vector_mask, matrix_mask


def get_sequential_nesting_masks(L, set_first_mode_const: bool = False):
    if set_first_mode_const:
        L += 1
    return torch.ones(L), torch.triu(torch.ones(L, L))


def _masks(neigs, step, sequential, set_first_mode_const=False):
    if sequential:
        return get_sequential_nesting_masks(neigs, set_first_mode_const)
    end_indices = list(range(step, neigs + 1, step))
    if neigs not in end_indices:
        end_indices.append(neigs)
    w = np.zeros(neigs)
    w[np.array(end_indices) - 1] = 1.0
    return get_joint_nesting_masks(w / w.sum(), set_first_mode_const)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _dev_f32(t, dev):
    return t.to(device=dev, dtype=torch.float32).contiguous()


class NestedLoRALossFunctionEVD(torch.autograd.Function):
    """Stand-alone loss on given (f, Tf, f1, f2): K2 + finalize forward, K3 backward.

    Gradients are returned per input exactly like the reference's hand-written backward
    (f: operator term, f1/f2: metric terms, Tf: None)."""

    @staticmethod
    def forward(ctx, f, Tf, f1, f2, vector_mask, matrix_mask):
        lib = _lib.load()
        dev = f.device
        fused._require_cuda(dev)
        if f.dim() != 2:
            raise NotImplementedError("matrix-valued outputs (B, L, O) are not supported by the fused loss")
        f, Tf, f1, f2 = (_dev_f32(t.detach(), dev) for t in (f, Tf, f1, f2))
        B, L = f.shape
        B1, B2 = f1.shape[0], f2.shape[0]
        v, Mm = _dev_f32(vector_mask, dev), _dev_f32(matrix_mask, dev)
        st = _stream(dev)
        LL = L * L
        terms = torch.empty(2 * LL + 1, dtype=torch.float32, device=dev)
        tmp = torch.empty(2 * LL + 1, dtype=torch.float32, device=dev)
        part = torch.empty(lib.nsvd_gram_partials_bytes(max(B, B1, B2), L), dtype=torch.uint8, device=dev)
        # three reductions: G1 from f1, G2 from f2, the operator sum from (f, Tf)
        _lib.check(lib.nsvd_gram_reduce(_lib.ptr(f1), _lib.ptr(f1), _lib.ptr(v), B1, L, B1, _lib.ptr(tmp),
                                        _lib.ptr(part), st), "nsvd_gram_reduce")
        terms[:LL].copy_(tmp[:LL])
        _lib.check(lib.nsvd_gram_reduce(_lib.ptr(f2), _lib.ptr(f2), _lib.ptr(v), B2, L, 0, _lib.ptr(tmp),
                                        _lib.ptr(part), st), "nsvd_gram_reduce")
        terms[LL:2 * LL].copy_(tmp[LL:2 * LL])
        _lib.check(lib.nsvd_gram_reduce(_lib.ptr(f), _lib.ptr(Tf), _lib.ptr(v), B, L, B, _lib.ptr(tmp),
                                        _lib.ptr(part), st), "nsvd_gram_reduce")
        terms[2 * LL:].copy_(tmp[2 * LL:])
        loss = torch.empty((), dtype=torch.float32, device=dev)
        coef = torch.empty(2 * LL, dtype=torch.float32, device=dev)
        # B1 + B2 need not equal B here (the reference allows independent f1, f2): normalise separately
        _lib.check(lib.nsvd_loss_finalize(_lib.ptr(terms), _lib.ptr(Mm), L, B1 + B2, B1, B2, _lib.ptr(loss),
                                          _lib.ptr(coef), st), "nsvd_loss_finalize")
        if B1 + B2 != B:   # operator term is a mean over f's rows (nestedlora.py:92)
            loss = loss + (2.0 / (B1 + B2) - 2.0 / B) * terms[2 * LL]
        ctx.save_for_backward(f, Tf, f1, f2, v, coef)
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        lib = _lib.load()
        f, Tf, f1, f2, v, coef = ctx.saved_tensors
        dev = f.device
        B, L = f.shape
        st = _stream(dev)
        gl = _dev_f32(grad_output, dev)
        g_f, g_f1, g_f2 = torch.empty_like(f), torch.empty_like(f1), torch.empty_like(f2)
        # operator part: coef == NULL ; metric parts: TF == NULL
        _lib.check(lib.nsvd_loss_dF(_lib.ptr(f), _lib.ptr(Tf), _lib.ptr(v), None, _lib.ptr(gl), B, L, B, B,
                                    _lib.ptr(g_f), st), "nsvd_loss_dF")
        _lib.check(lib.nsvd_loss_dF(_lib.ptr(f1), None, _lib.ptr(v), _lib.ptr(coef), _lib.ptr(gl), f1.shape[0], L,
                                    f1.shape[0], 1, _lib.ptr(g_f1), st), "nsvd_loss_dF")
        _lib.check(lib.nsvd_loss_dF(_lib.ptr(f2), None, _lib.ptr(v), _lib.ptr(coef), _lib.ptr(gl), f2.shape[0], L,
                                    0, 1, _lib.ptr(g_f2), st), "nsvd_loss_dF")
        return g_f, None, g_f1, g_f2, None, None


class NestedLoRA(nn.Module):
    def __init__(self, model, neigs, step=1, sort=False, sequential=False):
        self.name = "nestedlora"
        super().__init__()
        self.neigs = neigs
        self.sort = sort
        self.eigvals = None
        self.sort_indices = None
        self.sequential = sequential
        self.vector_mask, self.matrix_mask = _masks(neigs, step, sequential)
        self.model = model
        self.data_parallel = None      # neural_svd_b200.dist.PointParallel or None

    def forward(self, *args):
        output = self.model(*args)
        if self.sort_indices is not None and self.training:
            return output[:, self.sort_indices, ...]
        return output

    def register_eigvals(self, eigvals):
        print("NOTE: eigenvalues have been registered!")
        self.eigvals = torch.Tensor(eigvals)
        self.sort_indices = torch.sort(self.eigvals)[1].flip(0)

    def reset_eigvals(self):
        print("NOTE: eigenvalues have been reset!")
        self.eigvals = None
        self.sort_indices = None

    def _compute_loss(self, *args, evd=True) -> torch.Tensor:
        if evd:
            return NestedLoRALossFunctionEVD.apply(*args, self.vector_mask, self.matrix_mask)
        raise NotImplementedError

    def compute_loss_kernel(self, get_approx_kernel_op, x, importance, split_batch: bool, evd: bool = True):
        # no caller and no `get_approx_kernel_op` implementation exist in the reference
        raise NotImplementedError("compute_loss_kernel is not part of the accelerated path")

    def compute_loss_operator(self, operator, x, importance=None, evd: bool = True):
        if not evd:
            raise NotImplementedError
        return fused.compute_loss_operator(self, operator, x, importance, dp=self.data_parallel)


class NestedLoRALossFunctionForCDK(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, g, vector_mask, matrix_mask, set_first_mode_const=True, batch_weights=None, dp=None,
                diagnostics=True):
        lib = _lib.load()
        if batch_weights is not None:
            raise NotImplementedError("batch_weights is unused by the reference's callers and not supported")
        dev = f.device
        fused._require_cuda(dev)
        fd, gd = _dev_f32(f.detach(), dev), _dev_f32(g.detach(), dev)
        B, L = fd.shape
        fc = int(bool(set_first_mode_const))
        Lp = L + fc
        v, Mm = _dev_f32(vector_mask, dev), _dev_f32(matrix_mask, dev)
        if v.numel() != Lp:
            raise ValueError(f"vector_mask has {v.numel()} entries, expected {Lp}")
        st = _stream(dev)
        terms = torch.empty(2 * Lp * Lp + 1, dtype=torch.float32, device=dev)
        rs_joint = torch.empty(B, dtype=torch.float32, device=dev)
        engine = _lib.ENGINES[fused.get_engine()]
        nwork = lib.nsvd_cdk_work_bytes(B, L, fc, engine)
        work = torch.empty(nwork, dtype=torch.uint8, device=dev)
        _lib.check(lib.nsvd_cdk_fwd(_lib.ptr(fd), _lib.ptr(gd), _lib.ptr(v), B, L, fc, engine, _lib.ptr(terms),
                                    _lib.ptr(rs_joint), _lib.ptr(work), nwork, st), "nsvd_cdk_fwd")
        Bg = B
        if dp is not None:                      # rows are sharded: sum the Gram terms, and the row counts (python int)
            dp.allreduce_grads(terms)
            Bg = dp.global_counts(B, B, dev)[0]
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        coef = torch.empty(2 * Lp * Lp, dtype=torch.float32, device=dev)
        # the last NSVD_CDK_FINALIZE_SCRATCH bytes of the work buffer are reserved for the finalize partials (nsvd.h)
        fin = C.c_void_p((work.data_ptr() + nwork - _lib.CDK_FINALIZE_SCRATCH) & ~15)
        _lib.check(lib.nsvd_cdk_finalize(_lib.ptr(terms), _lib.ptr(Mm), Lp, Bg, _lib.ptr(losses), _lib.ptr(coef), fin,
                                         st), "nsvd_cdk_finalize")
        if diagnostics:       # (planes_ready = 1: `work` still holds the operand planes nsvd_cdk_fwd built from fd, gd)
            rs_indep = torch.empty(B * B - B, dtype=torch.float32, device=dev)
            _lib.check(lib.nsvd_cdk_offdiag(_lib.ptr(fd), _lib.ptr(gd), B, L, fc, engine, _lib.ptr(rs_indep),
                                            _lib.ptr(work), nwork, 1, st), "nsvd_cdk_offdiag")
        else:
            rs_indep = torch.empty(0, dtype=torch.float32, device=dev)
        ctx.save_for_backward(fd, gd, v, coef)
        ctx.fc, ctx.Bg, ctx.engine, ctx.work = fc, Bg, engine, work
        ctx.mark_non_differentiable(rs_joint, rs_indep)
        ctx.set_materialize_grads(False)   # no zero-filled gradients for the unused outputs (rs_indep alone is B^2 - B floats)
        return losses[0], losses[1], losses[2], rs_joint, rs_indep

    @staticmethod
    def backward(ctx, grad_output, *args) -> Tuple[torch.Tensor, ...]:
        lib = _lib.load()
        fd, gd, v, coef = ctx.saved_tensors
        dev = fd.device
        B, L = fd.shape
        # like the reference's backward (nestedlora.py:320-332) only the gradient of the first output (loss) is used
        if grad_output is None:
            return None, None, None, None, None, None, None, None
        gl = _dev_f32(grad_output, dev)
        gf, gg = torch.empty_like(fd), torch.empty_like(gd)
        _lib.check(lib.nsvd_cdk_bwd(_lib.ptr(fd), _lib.ptr(gd), _lib.ptr(v), _lib.ptr(coef), _lib.ptr(gl), B, L,
                                    ctx.fc, ctx.Bg, ctx.engine, _lib.ptr(gf), _lib.ptr(gg), _lib.ptr(ctx.work),
                                    ctx.work.numel(), 1, _stream(dev)), "nsvd_cdk_bwd")
        return gf, gg, None, None, None, None, None, None


class NestedLoRAForCDK(nn.Module):
    def __init__(self, model, neigs, step=1, sequential=False, set_first_mode_const=True):
        self.name = "nestedlora"
        super().__init__()
        self.neigs = neigs
        self.sequential = sequential
        self.vector_mask, self.matrix_mask = _masks(neigs, step, sequential, set_first_mode_const)
        self.set_first_mode_const = set_first_mode_const
        self.model = model
        self.data_parallel = None
        self.diagnostics = True        # rs_indep (B^2-B values, 67 MB at B=4096) can be switched off

    def forward(self, *args):
        return self.model(*args)

    def compute_loss(self, f, g, batch_weights=None) -> torch.Tensor:
        # the reference moves its CPU-resident masks to the device on every call (nestedlora.py:283-284: 1 MB of
        # pageable H2D copy for L = 512); here the device copies are cached until the mask tensors change
        v, Mm = (fused._nesting_masks(self, f.device) if f.device.type == "cuda"
                 else (self.vector_mask, self.matrix_mask))
        return NestedLoRALossFunctionForCDK.apply(f, g, v, Mm, self.set_first_mode_const, batch_weights,
                                                  self.data_parallel, self.diagnostics)
