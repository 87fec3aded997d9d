"""CUDA-graph capture of the whole fused step (K1 -> K2 -> finalize -> K3 -> K4) for the launch-bound
small-batch configurations (B = 128 / 512 in the reference's scripts).

    step = GraphedOperatorStep(method, operator, importance, batch_size=512)
    loss = step(x)          # == method.zero_grad(set_to_none=True); loss, _ = method.compute_loss_operator(...);
                            #    loss.backward()  — p.grad of every parameter is (re)written
    optimizer.step()

One graph launch replaces ~16 kernel launches, 10 memsets and the Python/ctypes work between them.
Single-GPU only (the data-parallel path keeps the eager sequence with its two NCCL all-reduces).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, fused
from .operators import describe_importance, describe_operator


class GraphedOperatorStep:
    def __init__(self, method, operator, importance, batch_size: int):
        lib = self.lib = _lib.load()
        self.method = method
        md = self.md = fused.describe_model(method)
        od = describe_operator(operator)
        imp = describe_importance(importance, md["ndim"])
        if getattr(method, "data_parallel", None) is not None:
            raise NotImplementedError("GraphedOperatorStep is single-GPU; use compute_loss_operator with data_parallel")
        dev = self.dev = md["Bff"].device
        fused._require_cuda(dev)
        B, L = batch_size, md["L"]
        self.B, self.L = B, L
        self.engine = fused.engine_for(md)
        self.ndim = md["ndim"]
        self.pb = fused._problem(md, od, imp, B)
        ns, nw = C.c_size_t(), C.c_size_t()
        _lib.check(lib.nsvd_scratch_bytes(C.byref(self.pb), self.engine, C.byref(ns), C.byref(nw)), "nsvd_scratch_bytes")
        f32 = dict(dtype=torch.float32, device=dev)
        self.saved = torch.empty(ns.value, dtype=torch.uint8, device=dev)
        self.work = torch.empty(nw.value, dtype=torch.uint8, device=dev)
        self.partials = torch.empty(lib.nsvd_gram_partials_bytes(B, L), dtype=torch.uint8, device=dev)
        self.x = torch.zeros((B, md["ndim"]), **f32)
        self.F, self.TF, self.dF = (torch.empty((B, L), **f32) for _ in range(3))
        self.terms = torch.empty(2 * L * L + 1, **f32)
        self.coef = torch.empty(2 * L * L, **f32)
        self.loss = torch.empty((), **f32)
        self.v = method.vector_mask.to(**f32).contiguous()
        self.Mm = method.matrix_mask.to(**f32).contiguous()
        self.tensors = md["ws"] + md["bs"] + ([md["scales"]] if md["scales"] is not None else [])
        sizes = [t.numel() for t in self.tensors]
        self.flat = torch.empty(sum(sizes), **f32)
        self.views = [v.view(t.shape) for v, t in zip(self.flat.split(sizes), self.tensors)]
        self.b1 = (B + 1) // 2
        # warm-up on a side stream (first-call attribute setup happens outside the capture), then capture
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            self._enqueue()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._enqueue()

    def _enqueue(self):
        lib, md, pb, dev = self.lib, self.md, self.pb, self.dev
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        pr = fused._params_struct(md)
        B, L, p = self.B, self.L, _lib.ptr
        _lib.check(lib.nsvd_fwd_streams(C.byref(pb), C.byref(pr), self.engine, p(self.x), p(self.F), p(self.TF),
                                        p(self.saved), self.saved.numel(), p(self.work), self.work.numel(), st),
                   "nsvd_fwd_streams")
        _lib.check(lib.nsvd_gram_reduce(p(self.F), p(self.TF), p(self.v), B, L, self.b1, p(self.terms),
                                        p(self.partials), st), "nsvd_gram_reduce")
        _lib.check(lib.nsvd_loss_finalize(p(self.terms), p(self.Mm), L, B, self.b1, B - self.b1, p(self.loss),
                                          p(self.coef), st), "nsvd_loss_finalize")
        _lib.check(lib.nsvd_loss_dF(p(self.F), p(self.TF), p(self.v), p(self.coef), None, B, L, self.b1, B,
                                    p(self.dF), st), "nsvd_loss_dF")
        gr = _lib.Grads()
        for i in range(4):
            gr.dW[i] = self.views[i].data_ptr()
            gr.db[i] = self.views[4 + i].data_ptr()
        gr.dmask_scales = self.views[8].data_ptr() if md["scales"] is not None else None
        _lib.check(lib.nsvd_mlp_bwd(C.byref(pb), C.byref(pr), self.engine, p(self.x), p(self.dF), p(self.saved),
                                    self.saved.numel(), C.byref(gr), p(self.work), self.work.numel(), st), "nsvd_mlp_bwd")

    def __call__(self, x):
        if x.dim() == 3:
            x = x.reshape(x.shape[0], -1)
        if tuple(x.shape) != (self.B, self.ndim):
            raise ValueError(f"GraphedOperatorStep was captured for x of shape ({self.B}, {self.ndim}); got {tuple(x.shape)}")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        for t, g in zip(self.tensors, self.views):
            t.grad = g
        return self.loss

    @property
    def aux(self):
        """dict(f=F, Tf=TF, eigvals=None) of the last replay (static buffers)."""
        return dict(f=self.F, Tf=self.TF, eigvals=None)
