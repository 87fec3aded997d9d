"""Fused NestedLoRA step: the autograd boundary above the C-ABI kernels.

Replaces, for one call of `NestedLoRA.compute_loss_operator` + `loss.backward()`
(methods/nestedlora.py:254-267, examples/operator/__init__.py:62-68):
    K1 nsvd_fwd_streams  ->  K2 nsvd_gram_reduce  -> [all-reduce #1] -> nsvd_loss_finalize
    K3 nsvd_loss_dF      ->  K4 nsvd_mlp_bwd      -> [all-reduce #2]
PyTorch is used for device memory, the stream and torch.distributed only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from .operators import describe_importance, describe_operator

_ENGINE = os.environ.get("NSVD_ENGINE", "f16x3")


def set_engine(name: str):
    """'f16x3' (tcgen05 tensor cores, fp32 operands as fp16 hi/lo planes; default) or 'fp32' (CUDA-core validation
    engine).  'bf16x3' / 'tc' are accepted aliases of 'f16x3'."""
    global _ENGINE
    if name not in _lib.ENGINES:
        raise ValueError(f"unknown engine {name!r}; choose from {sorted(_lib.ENGINES)}")
    _ENGINE = name


def get_engine() -> str:
    return _ENGINE


def set_microbatch(points: int):
    """Points per micro-batch of the tensor-core engine (bounds the size of the stream scratch buffers)."""
    _lib.load().nsvd_set_tc_microbatch(int(points))


def _require_cuda(dev: torch.device):
    if dev.type != "cuda":
        raise RuntimeError("neural_svd_b200 has no CPU path: parameters and inputs must live on a CUDA (sm_100) device")


# ------------------------------------------------------------------------------------------
# model / problem description (duck-typed: the reference's own modules are accepted)
# ------------------------------------------------------------------------------------------
def _param_objects(plist):
    """The Parameter objects of an nn.ParameterList without its per-item string indexing (a few us per item)."""
    d = getattr(plist, "_parameters", None)
    return tuple(d.values()) if d else tuple(plist)


def describe_model(model):
    """Pull the parameter tensors out of WaveFunctions(ParallelMLP(GaussianFourierFeatureTransform)).

    The structural checks run once per module: the description is cached on it and reused while the parameter OBJECTS
    (held by the cache, so their identities cannot be recycled), their dtypes / devices and the scalar settings are
    unchanged.  Data pointers are never cached - the kernels always receive the tensors' current storage."""
    m = getattr(model, "model", model)          # NestedLoRA(...) -> .model
    base = getattr(m, "base", None)
    if base is None or not hasattr(base, "ws") or not hasattr(base, "bs"):
        raise NotImplementedError("fused path needs WaveFunctions(base=ParallelMLP(...)) (--parallel 1)")
    fm = getattr(base, "feature_map", None)
    mask = getattr(m, "boundary_mask", None)
    scales = getattr(mask, "scales", None)
    box = mask if scales is None else getattr(mask, "boundary_mask", None)
    objs = (fm, getattr(fm, "_B", None)) + _param_objects(base.ws) + _param_objects(base.bs) + (scales, box)
    settings = (getattr(fm, "append_raw", False), getattr(base, "weight_normalization", False), getattr(base, "bias", True),
                getattr(m, "hard_mul_const", 1.0), getattr(box, "mode", None), getattr(box, "lim", None),
                tuple((t.dtype, t.device) for t in objs[1:-1] if t is not None))
    hit = m.__dict__.get("_nsvd_md")
    if (hit is not None and hit[1] == settings and len(hit[0]) == len(objs)
            and all(a is b for a, b in zip(hit[0], objs))):
        if fm._B._version == hit[3]:     # (a non-contiguous _B is described through a copy: redo it when _B changes)
            return hit[2]
    md = _describe_model_uncached(m, base, fm, mask, scales, box)
    m.__dict__["_nsvd_md"] = (objs, settings, md, fm._B._version)
    return md


def _describe_model_uncached(m, base, fm, mask, scales, box):
    if fm is None or not hasattr(fm, "_B") or getattr(fm, "append_raw", False):
        raise NotImplementedError("fused path needs the Fourier feature map without raw append")
    if getattr(base, "weight_normalization", False) or not getattr(base, "bias", True):
        raise NotImplementedError("weight_normalization / bias=False are not supported")
    ws, bs = list(base.ws), list(base.bs)
    if len(ws) != 4 or len(bs) != 4:
        raise NotImplementedError("fused path is built for 3 hidden layers (mlp_hidden_dims='128,128,128')")
    Bff = fm._B.detach()
    if not Bff.is_contiguous():          # the deterministic map is built as a transposed view (utils.py:108-111)
        Bff = Bff.contiguous()
    if Bff.shape[0] not in (2, 3):
        raise NotImplementedError("fused path is built for one particle in ndim = 2 or 3")
    L, H, K0 = ws[0].shape
    Mff = Bff.shape[1]
    ok = (H == 128 and K0 == 2 * Mff and tuple(ws[1].shape) == (L, 128, 128) and tuple(ws[2].shape) == (L, 128, 128)
          and tuple(ws[3].shape) == (L, 1, 128) and all(tuple(b.shape) == (L, 128, 1) for b in bs[:3])
          and tuple(bs[3].shape) == (L, 1, 1))
    if not ok:
        raise NotImplementedError("unexpected ParallelMLP parameter shapes for the fused path")
    box_mode, box_lim = _describe_box(box)
    params = [Bff] + ws + bs + ([scales] if scales is not None else [])
    for p in params:
        if p.dtype != torch.float32:
            raise NotImplementedError("fused path is fp32 (reference default, --use_amp off)")
        if not p.is_contiguous():
            raise RuntimeError("parameters must be contiguous")
    return dict(Bff=Bff, ws=ws, bs=bs, scales=scales, L=L, Mff=Mff, ndim=int(Bff.shape[0]), box_mode=box_mode,
                box_lim=box_lim,
                hard_mul_const=float(getattr(m, "hard_mul_const", 1.0)))


def _describe_box(box):
    """(mode, lim) of a DirichletBoundaryMaskBox (pde/boundary.py:16-37), (BOX_NONE, 0) for `lambda x: 1.`."""
    if box is None:
        return _lib.BOX_NONE, 0.0
    mode, lim = getattr(box, "mode", None), getattr(box, "lim", None)
    if mode in ("dir_box_sqrt", "dir_box_exp") and lim is not None:
        return (_lib.BOX_SQRT if mode == "dir_box_sqrt" else _lib.BOX_EXP), float(lim)
    try:
        if float(box(None)) == 1.0:
            return _lib.BOX_NONE, 0.0
    except Exception:
        pass
    raise NotImplementedError("boundary mask must be ExponentialMask, DirichletBoundaryMaskBox, both, or none")


def _problem(md, od, imp, B) -> _lib.Problem:
    return _lib.Problem(n_points=B, n_copies=md["L"], n_fourier=md["Mff"], hidden=128,
                        potential=od["potential"], has_exp_mask=int(md["scales"] is not None),
                        pot_coef=od["pot_coef"], scale_kinetic=od["scale_kinetic"], op_scale=od["op_scale"],
                        op_shift=od["op_shift"], sampling_sigma=imp["sigma"], hard_mul_const=md["hard_mul_const"],
                        importance=imp["importance"], box_mask=md["box_mode"], pot_coef2=od.get("pot_coef2", 0.0),
                        box_lim=md["box_lim"], fd_eps=od.get("fd_eps", 0.0), ndim=md["ndim"])


def _params_struct(md) -> _lib.Params:
    pr = _lib.Params()
    pr.Bff = md["Bff"].data_ptr()
    for i in range(4):
        pr.W[i] = md["ws"][i].data_ptr()
        pr.b[i] = md["bs"][i].data_ptr()
    pr.mask_scales = md["scales"].data_ptr() if md["scales"] is not None else None
    return pr


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _Scratch:
    """Per-owner cache of the kernel scratch buffers (allocated by PyTorch's caching allocator)."""

    def __init__(self):
        self.key = None
        self.saved = self.work = self.partials = None
        self.version = 0

    def ensure(self, lib, pb: _lib.Problem, engine: int, dev):
        ns, nw = C.c_size_t(), C.c_size_t()
        _lib.check(lib.nsvd_scratch_bytes(C.byref(pb), engine, C.byref(ns), C.byref(nw)), "nsvd_scratch_bytes")
        key = (pb.n_points, pb.n_copies, pb.n_fourier, engine, str(dev), ns.value, nw.value)
        if key != self.key:
            self.saved = torch.empty(ns.value, dtype=torch.uint8, device=dev)
            self.work = torch.empty(nw.value, dtype=torch.uint8, device=dev)
            npart = lib.nsvd_gram_partials_bytes(pb.n_points, pb.n_copies)
            self.partials = torch.empty(npart, dtype=torch.uint8, device=dev)
            self.key = key
        return self


def _scratch_of(owner) -> _Scratch:
    sc = owner.__dict__.get("_nsvd_scratch")
    if sc is None:
        sc = _Scratch()
        owner.__dict__["_nsvd_scratch"] = sc
    return sc


def _prep_x(x, dev, ndim=2):
    if x.dim() == 3:
        x = x.reshape(x.shape[0], -1)
    if x.dim() != 2 or x.shape[1] != ndim:
        raise NotImplementedError(f"fused path expects x of shape (B, {ndim}); got {tuple(x.shape)}")
    return x.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()


_warned_3d = False


def engine_for(md) -> int:
    """Engine code of a step: the selected one, except that ndim = 3 (five forward-mode streams: 640 accumulator columns
    per 128 units exceed the 512 TMEM columns the tcgen05 tiles are built around) always runs on the fp32 CUDA-core
    engine of the same library - still hand-written sm_100a kernels on the device, never a CPU path."""
    global _warned_3d
    if md["ndim"] == 3:
        if _ENGINE not in ("fp32", "fp32_simt") and not _warned_3d:
            import warnings
            warnings.warn("neural_svd_b200: ndim = 3 runs on the fp32 CUDA-core engine (the tensor-core engine is 2D)")
            _warned_3d = True
        return _lib.ENGINE_FP32_SIMT
    return _lib.ENGINES[_ENGINE]


def _forward_kernels(lib, owner, md, od, imp, x, engine):
    dev = md["Bff"].device
    _require_cuda(dev)
    x = _prep_x(x, dev, md["ndim"])
    B, L = x.shape[0], md["L"]
    pb = _problem(md, od, imp, B)
    sc = _scratch_of(owner).ensure(lib, pb, engine, dev)
    sc.version += 1
    F = torch.empty((B, L), dtype=torch.float32, device=dev)
    TF = torch.empty((B, L), dtype=torch.float32, device=dev)
    pr = _params_struct(md)
    st = _stream(dev)
    _lib.check(lib.nsvd_fwd_streams(C.byref(pb), C.byref(pr), engine, _lib.ptr(x), _lib.ptr(F), _lib.ptr(TF),
                                    _lib.ptr(sc.saved), sc.saved.numel(), _lib.ptr(sc.work), sc.work.numel(),
                                    st), "nsvd_fwd_streams")
    return x, pb, pr, sc, F, TF, st


def apply_operator(model, operator, x, importance):
    """`operator(model, x, importance) -> (Tf, f)` on the fused forward kernel (no autograd graph)."""
    lib = _lib.load()
    md, od = describe_model(model), describe_operator(operator)
    imp = describe_importance(importance, md["ndim"])
    with torch.no_grad():
        _, _, _, _, F, TF, _ = _forward_kernels(lib, getattr(model, "model", model), md, od, imp, x, engine_for(md))
    return TF, F


def model_values(model, x):
    """hard_mul_const * base(x) * mask(x), shape (B, L) (WaveFunctions.forward, pde/__init__.py:15-16)."""
    lib = _lib.load()
    md = describe_model(model)
    od = dict(potential=_lib.POT_INFINITE_WELL, pot_coef=0.0, scale_kinetic=1.0, op_scale=1.0, op_shift=0.0)
    with torch.no_grad():
        _, _, _, _, F, _, _ = _forward_kernels(lib, getattr(model, "model", model), md, od,
                                               dict(importance=_lib.IMP_NONE, sigma=1.0), x, engine_for(md))
    return F


# ------------------------------------------------------------------------------------------
# the fused training step
# ------------------------------------------------------------------------------------------
class _FusedOperatorStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, method, operator, importance, x, dp, *params):
        lib = _lib.load()
        md, od = describe_model(method), describe_operator(operator)
        imp = describe_importance(importance, md["ndim"])
        engine = engine_for(md)
        x, pb, pr, sc, F, TF, st = _forward_kernels(lib, method, md, od, imp, x, engine)
        dev = F.device
        B, L = F.shape
        b1 = (B + 1) // 2                                     # torch.chunk(f, 2), nestedlora.py:263
        v, Mm = _nesting_masks(method, dev)
        terms = torch.empty(2 * L * L + 5, dtype=torch.float32, device=dev)   # + 4 count floats (data parallel)
        _lib.check(lib.nsvd_gram_reduce(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), B, L, b1, _lib.ptr(terms),
                                        _lib.ptr(sc.partials), st), "nsvd_gram_reduce")
        Bg, B1g, B2g = B, b1, B - b1
        if dp is not None:
            dp.allreduce_terms(terms, B, b1)                  # all-reduce #1; the global counts travel inside it
            Bg = B1g = B2g = 0                                # -> the kernels take them from the device buffer
        loss = torch.empty((), dtype=torch.float32, device=dev)
        coef = torch.empty(2 * L * L + 1, dtype=torch.float32, device=dev)
        _lib.check(lib.nsvd_loss_finalize(_lib.ptr(terms), _lib.ptr(Mm), L, Bg, B1g, B2g, _lib.ptr(loss),
                                          _lib.ptr(coef), st), "nsvd_loss_finalize")
        ctx.state = dict(md=md, pb=pb, sc=sc, version=sc.version, x=x, F=F, TF=TF, v=v, coef=coef, b1=b1,
                         Bg=Bg, engine=engine, dp=dp, nparams=len(params))
        ctx.mark_non_differentiable(F, TF)
        ctx.set_materialize_grads(False)     # no zero-filled (B, L) gradients for the two auxiliary outputs
        return loss, F, TF

    @staticmethod
    def backward(ctx, gloss, gF, gTF):
        lib = _lib.load()
        s = ctx.state
        if gloss is None:                    # only f / Tf were used downstream: they carry no gradient (utils of the aux dict)
            return (None,) * (5 + s["nparams"])
        md, pb, sc = s["md"], s["pb"], s["sc"]
        if sc.version != s["version"]:
            raise RuntimeError("the scratch buffers of this NestedLoRA object were overwritten by a later forward "
                               "call before backward ran; call backward() before the next compute_loss_operator()")
        F, TF = s["F"], s["TF"]
        dev = F.device
        B, L = F.shape
        gl = gloss.to(device=dev, dtype=torch.float32).contiguous()
        st = _stream(dev)                # the backward engine's stream for this device
        dF = torch.empty_like(F)
        _lib.check(lib.nsvd_loss_dF(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(s["v"]), _lib.ptr(s["coef"]), _lib.ptr(gl),
                                    B, L, s["b1"], s["Bg"], _lib.ptr(dF), st), "nsvd_loss_dF")
        tensors = md["ws"] + md["bs"] + ([md["scales"]] if md["scales"] is not None else [])
        sizes = [t.numel() for t in tensors]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        views = [v.view(t.shape) for v, t in zip(flat.split(sizes), tensors)]
        gr = _lib.Grads()
        for i in range(4):
            gr.dW[i] = views[i].data_ptr()
            gr.db[i] = views[4 + i].data_ptr()
        gr.dmask_scales = views[8].data_ptr() if md["scales"] is not None else None
        pr = _params_struct(md)
        _lib.check(lib.nsvd_mlp_bwd(C.byref(pb), C.byref(pr), s["engine"], _lib.ptr(s["x"]), _lib.ptr(dF),
                                    _lib.ptr(sc.saved), sc.saved.numel(), C.byref(gr), _lib.ptr(sc.work),
                                    sc.work.numel(), st), "nsvd_mlp_bwd")
        if s["dp"] is not None:
            s["dp"].allreduce_grads(flat)                      # all-reduce #2
        # params were passed as (Bff, ws0..3, bs0..3[, scales]); Bff gets no gradient (utils.py:116-118)
        return (None, None, None, None, None, None) + tuple(views)


def _sort_permutation(method):
    """sort_indices of register_eigvals() when it is active (NestedLoRA.forward, nestedlora.py:195-200), else None."""
    si = getattr(method, "sort_indices", None)
    if si is None or not getattr(method, "training", True):
        return None
    return si


def _nesting_masks(method, dev):
    """(vector_mask, matrix_mask) on the device.  With registered eigenvalues the reference permutes the model's OUTPUT
    columns, f'[:, j] = f[:, s_j] (nestedlora.py:197-198), before the loss; the loss of f' under (v, M) equals the loss of
    the unpermuted f under v'[s_j] = v[j], M'[s_j, s_k] = M[j, k], so the kernels run on the network's own column order
    with permuted masks and dF comes out in that order too.
    The reference keeps the masks on the CPU and moves them every call (nestedlora.py:87-88: two pageable H2D copies,
    ~60 us of host time per step); here the device copies are cached until the mask / sort_indices objects or their
    contents (tensor version counters) change."""
    vm, mm, si = method.vector_mask, method.matrix_mask, _sort_permutation(method)
    ver = (vm._version, mm._version, -1 if si is None else si._version, dev)
    hit = method.__dict__.get("_nsvd_masks")
    if hit is not None and hit[0] is vm and hit[1] is mm and hit[2] is si and hit[3] == ver:
        return hit[4], hit[5]
    v = vm.to(device=dev, dtype=torch.float32)
    Mm = mm.to(device=dev, dtype=torch.float32)
    if si is not None:
        sd = si.to(device=dev, dtype=torch.long)
        v2, M2 = torch.empty_like(v), torch.empty_like(Mm)
        v2[sd] = v
        M2[sd[:, None], sd[None, :]] = Mm
        v, Mm = v2, M2
    v, Mm = v.contiguous(), Mm.contiguous()
    if v is vm or Mm is mm:              # already on the device in fp32: `.to` returned the caller's own tensors
        v, Mm = v.clone(), Mm.clone()    # (the cache must not alias what the caller may modify between steps)
    method.__dict__["_nsvd_masks"] = (vm, mm, si, ver, v, Mm)
    return v, Mm


def compute_loss_operator(method, operator, x, importance, dp=None):
    """Fused equivalent of NestedLoRA.compute_loss_operator (nestedlora.py:254-267)."""
    md = describe_model(method)
    if dp is not None:
        dp.sync_parameters(method)                            # once per module: replicas start from rank 0's weights
    params = [md["Bff"]] + md["ws"] + md["bs"] + ([md["scales"]] if md["scales"] is not None else [])
    loss, F, TF = _FusedOperatorStep.apply(method, operator, importance, x, dp, *params)
    si = _sort_permutation(method)
    if si is not None:                       # the caller sees the permuted columns, as operator(self, x) returns them
        si = si.to(device=F.device, dtype=torch.long)
        F, TF = F[:, si], TF[:, si]
    return loss, dict(f=F, Tf=TF, eigvals=None)
