"""compute_spectrum_evd on the fused kernels (methods/spectrum.py:29-102).

Same signature and outputs as the reference.  Per grid chunk: the fused forward kernel gives
(T phi, phi); `nsvd_cross_gram` accumulates cov = phi^T phi and the FULL cross Gram
quad = phi^T T phi on the device (with the reference's nan_to_num and origin-row zeroing fused in);
the Rayleigh quotients / norms / optional post-alignment are L x L host work.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from . import _lib, fused


def check_only_one_arg(*args):
    return np.array([int(arg is not None) for arg in args]).sum() == 1


def get_sqrt_weight_func(importance_train, importance_val):
    def sqrt_weight_func(x):
        sqrt_ws_train = 1.0
        sqrt_ws_val = 1.0
        if importance_train is not None:
            sqrt_ws_train = importance_train(x).sqrt()
        if importance_val is not None:
            sqrt_ws_val = importance_val(x).sqrt()
        return sqrt_ws_train, sqrt_ws_val
    return sqrt_weight_func


def post_alignment(eigfuncs, cov, quad):
    """spectrum.py:161-170 (L x L, host)."""
    from scipy.linalg import eigh
    eigvals_cov, eigvecs_cov = eigh(cov)
    whitening = eigvecs_cov @ np.diag(1 / np.sqrt(eigvals_cov)) @ eigvecs_cov.T
    eigvals, V = eigh(whitening @ quad @ whitening)
    eigvals = np.sqrt(eigvals[::-1])
    V = V[:, ::-1]
    eigfuncs = eigfuncs @ (V.T @ whitening).T
    return eigfuncs, eigvals, np.eye(quad.shape[0])


def compute_spectrum_evd(model, dataloader, operator, importance_train=None, importance_val=None,
                         set_first_mode_const=False, post_align=False, normalize=False, sort=False, gpu=None,
                         device=None, return_eigfuncs=True, data_parallel=None):
    assert check_only_one_arg(gpu, device)
    if set_first_mode_const:
        raise NotImplementedError("set_first_mode_const is a CDK option; the operator path never sets it")
    lib = _lib.load()
    dev = torch.device("cuda", gpu) if gpu is not None else torch.device(device)
    fused._require_cuda(dev)
    sqrt_weight_func = get_sqrt_weight_func(importance_train, importance_val)
    start = time.time()
    n = 0
    cov = quad = part = None
    eigfuncs = []
    for (x, _) in dataloader:
        if isinstance(x, list):
            x = x[0]
        x = x.to(dev)
        x2 = x.reshape(x.shape[0], -1).float().contiguous()
        sqrt_ws_train, sqrt_ws_val = sqrt_weight_func(x2)
        Tphi, phi = operator(model, x2, importance=importance_train)       # fused forward kernel
        B, L = phi.shape
        if cov is None:
            cov = torch.zeros(L, L, dtype=torch.float32, device=dev)
            quad = torch.zeros(L, L, dtype=torch.float32, device=dev)
        npart = lib.nsvd_gram_partials_bytes(B, L)
        if part is None or part.numel() < npart:
            part = torch.empty(npart, dtype=torch.uint8, device=dev)
        roww = (sqrt_ws_train / sqrt_ws_val)
        roww = roww.reshape(-1).float().contiguous() if torch.is_tensor(roww) else None
        if return_eigfuncs:       # kept on the device: ONE device-to-host copy after the loop, as spectrum.py:83
            eigfuncs.append(sqrt_ws_train * phi if torch.is_tensor(sqrt_ws_train) else phi)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.nsvd_cross_gram(_lib.ptr(phi), _lib.ptr(Tphi), _lib.ptr(roww), _lib.ptr(x2), B, L,
                                       _lib.ptr(cov), _lib.ptr(quad), _lib.ptr(part), st), "nsvd_cross_gram")
        n += B
    if data_parallel is not None:                                           # grid sharded across ranks
        import torch.distributed as dist
        both = torch.stack([cov, quad])
        dist.all_reduce(both, group=data_parallel.group)
        cnt = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt, group=data_parallel.group)
        cov, quad, n = both[0], both[1], int(cnt)
    cov = (cov / n).cpu().numpy()
    quad = (quad / n).cpu().numpy()
    print(f"Took {time.time() - start}s to compute spectrum with data of size {n}")
    outputs = dict()
    outputs["eigfuncs"] = eigfuncs = torch.cat(eigfuncs, dim=0).cpu().numpy() if eigfuncs else None
    outputs["cov"], outputs["quad"] = cov, quad
    outputs["eigvals"] = eigvals = np.diag(quad) / np.diag(cov)
    outputs["norms"] = norms = np.diag(cov)
    if normalize:
        outputs["cov"] = cov / (np.sqrt(norms[:, np.newaxis]) @ np.sqrt(norms[:, np.newaxis]).T)
        if eigfuncs is not None:
            outputs["eigfuncs"] = eigfuncs / np.sqrt(norms).reshape(1, -1)
    if sort:
        si = np.argsort(eigvals)[::-1]
        outputs["eigvals"] = outputs["eigvals"][si]
        if outputs["eigfuncs"] is not None:
            outputs["eigfuncs"] = outputs["eigfuncs"][:, si, ...]
        outputs["cov"] = outputs["cov"][:, si][si, :]
        outputs["quad"] = outputs["quad"][:, si][si, :]
        outputs["norms"] = outputs["norms"][si]
    if post_align and outputs["eigfuncs"] is not None:
        outputs["eigfuncs_aligned"], outputs["eigvals_aligned"], outputs["cov_aligned"] = post_alignment(
            outputs["eigfuncs"], outputs["cov"], outputs["quad"])
    return outputs
