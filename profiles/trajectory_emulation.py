"""CPU emulation of the fixed-seed training run under different tensor-core arithmetic models
(evidence for DESIGN.md §3; not product code, never imported by the product).

The oracle's contractions are replaced by an emulation of  D = A.B^T  on the tcgen05 pipe:
  * operands rounded to a plane format: 'bb' bf16 hi + bf16 lo (16 bits), 'hh' fp16 hi + fp16 lo with a per-tensor
    power-of-two scale (22 bits, absolute floor 2^-25 of the scaled unit), 'bbb' three bf16 planes (24 bits),
    'exact' (fp64);
  * products hi*hi + lo*hi + hi*lo (+ the 2^-16 terms for 'bbb');
  * accumulation: fp64 ('acc=exact') or a chain of K=16 steps whose running sum is truncated toward zero to
    `nbits` significant bits after every MMA (the measured behaviour of the fp32 accumulator in TMEM:
    error linear in the number of chained MMAs, 1.9e-8 per MMA), restarted every `chain` contraction elements
    with the partial sums added in fp32.
Parameters live in fp32 and are updated by RMSprop exactly as torch.optim.RMSprop does (the test's loop).
usage: python profiles/trajectory_emulation.py MODE [steps]      MODE e.g.  fp32 | exact | bb | hh | hh:chain=256
"""
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import nsvd_oracle as O  # noqa: E402

REAL_EINSUM = np.einsum


def bf16(v32):
    u = v32.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def planes(v, fmt):
    """list of fp64 planes (largest first) and the power-of-two scale they were taken at."""
    v32 = np.ascontiguousarray(v).astype(np.float32)
    if fmt == "exact":
        return [v.astype(np.float64)], 1.0
    if fmt == "bb" or fmt == "bbb":
        out, r = [], v32
        for _ in range(2 if fmt == "bb" else 3):
            p = bf16(r)
            out.append(p.astype(np.float64))
            r = r - p
        return out, 1.0
    if fmt == "hh":
        m = float(np.abs(v32).max())
        s = 2.0 ** math.floor(math.log2(16384.0 / m)) if m > 0 else 1.0
        w = v32 * np.float32(s)
        hi = w.astype(np.float16).astype(np.float32)
        lo = (w - hi).astype(np.float16).astype(np.float32)
        return [hi.astype(np.float64), lo.astype(np.float64)], s
    raise ValueError(fmt)


def trunc_bits(a, nbits):
    """round toward zero to nbits significant bits (fp64 container)."""
    mask = np.int64(-1) << np.int64(53 - nbits)
    return (a.view(np.int64) & mask).view(np.float64)


class Emu:
    def __init__(self, fmt="bb", chain=0, nbits=25):
        self.fmt, self.chain, self.nbits = fmt, chain, nbits

    def products(self, pa, pb):
        """pairs of (A plane, B plane) in issue order: small terms first, hi*hi last (as the kernels do)."""
        if len(pa) == 1:
            return [(pa[0], pb[0])]
        if len(pa) == 2:
            return [(pa[1], pb[0]), (pa[0], pb[1]), (pa[0], pb[0])]
        h, m, l = 0, 1, 2
        order = [(m, m), (h, l), (l, h), (h, m), (m, h), (h, h)]
        return [(pa[i], pb[j]) for i, j in order]

    def mm(self, A, Bm):
        """A (..., M, K) @ Bm (..., K, N) under the arithmetic model."""
        pa, sa = planes(A, self.fmt)
        pb, sb = planes(Bm, self.fmt)
        prods = self.products(pa, pb)
        K = A.shape[-1]
        if not self.chain:
            acc = 0.0
            for a, b in prods:
                acc = acc + np.matmul(a, b)
            return acc / (sa * sb)
        total = None
        for c0 in range(0, K, self.chain):
            c1 = min(K, c0 + self.chain)
            acc = None
            for k0 in range(c0, c1, 16):
                k1 = min(c1, k0 + 16)
                for a, b in prods:
                    part = np.matmul(a[..., k0:k1], b[..., k0:k1, :])
                    acc = part if acc is None else acc + part
                    acc = trunc_bits(np.ascontiguousarray(acc), self.nbits)
            acc32 = acc.astype(np.float32)
            total = acc32 if total is None else (total + acc32).astype(np.float32)
        return total.astype(np.float64) / (sa * sb)

    def einsum(self, spec, a, b):
        if spec == "slbk,lhk->slbh":
            return self.mm(a, np.swapaxes(b, -1, -2)[None])
        if spec == "lbh,bk->lhk":
            return self.mm(np.swapaxes(a, -1, -2), b[None])
        if spec == "lbh,lbk->lhk":
            return self.mm(np.swapaxes(a, -1, -2), b)
        if spec == "lbh,lhk->lbk":
            return self.mm(a, b)
        raise NotImplementedError(spec)


def fast_einsum(spec, a, b):
    if spec == "slbk,lhk->slbh":
        return np.matmul(a, np.swapaxes(b, -1, -2)[None])
    if spec == "lbh,bk->lhk":
        return np.matmul(np.swapaxes(a, -1, -2), b[None])
    if spec == "lbh,lbk->lhk":
        return np.matmul(np.swapaxes(a, -1, -2), b)
    if spec == "lbh,lhk->lbk":
        return np.matmul(a, b)
    return REAL_EINSUM(spec, a, b)


def parse_mode(mode):
    parts = mode.split(":")
    kw = {}
    for p in parts[1:]:
        k, v = p.split("=")
        kw[k] = int(v)
    return parts[0], kw


def run(mode, steps=None, verbose=True):
    import torch
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                             "run_hyd_b128_seq_L16.npz"))
    S, B, seed = int(d["steps"]), int(d["B"]), int(d["seed"])
    S_run = steps or S
    cfg = O.PathConfig.hydrogen(sequential=True)
    params = {k: v.astype(np.float32) for k, v in O.init_params_like_reference(cfg, seed).items()}
    fmt, kw = parse_mode(mode)
    dt = np.float32 if fmt == "fp32" else np.float64
    emu = None if fmt in ("fp32", "exact") else Emu(fmt, kw.get("chain", 0), kw.get("nbits", 25))
    O.np.einsum = emu.einsum if emu else fast_einsum
    g = torch.Generator().manual_seed(4242)
    sq = {k: np.zeros_like(v) for k, v in params.items()}
    lr0, alpha, eps = 1e-4, 0.999, 1e-10
    losses = []
    t0 = time.time()
    try:
        for t in range(S_run):
            x = (cfg.sampling_scale * torch.randn((B, 1, cfg.ndim), generator=g)).reshape(B, -1).numpy()
            out = O.train_step(x.astype(dt), {k: v.astype(dt) for k, v in params.items()}, cfg)
            lr = np.float32(0.5 * lr0 * (1 + math.cos(math.pi * t / S)))
            for k, gk in out["grads"].items():
                gk = gk.reshape(params[k].shape).astype(np.float32)
                sq[k] = (np.float32(alpha) * sq[k] + np.float32(1 - alpha) * gk * gk).astype(np.float32)
                avg = (np.sqrt(sq[k]) + np.float32(eps)).astype(np.float32)
                params[k] = (params[k] - lr * (gk / avg)).astype(np.float32)
            losses.append(float(out["loss"]))
            if verbose and (t % 50 == 0 or t == S_run - 1):
                print(f"  step {t} loss {losses[-1]:.6f} ref64 {d['loss64'][t]:.6f} rel {abs(losses[-1]/d['loss64'][t]-1):.2e} "
                      f"({time.time()-t0:.0f}s)", flush=True)
    finally:
        O.np.einsum = fast_einsum
    res = dict(mode=mode, loss_step1=abs(losses[0] / d["loss64"][0] - 1))
    if S_run == S:
        ge = torch.Generator().manual_seed(777)
        xe = (cfg.sampling_scale * torch.randn((8192, 1, cfg.ndim), generator=ge)).reshape(8192, -1).numpy().astype(np.float64)
        p64 = {k: v.astype(np.float64) for k, v in params.items()}
        u = O.forward_streams(xe, p64, cfg)
        Tf, f, _ = O.operator_apply(xe, u, p64, cfg)
        norms, ray = (f * f).mean(0), (f * Tf).sum(0) / (f * f).sum(0)
        e_n, e_r = np.abs(norms / d["norms64"] - 1), np.abs(ray / d["rayleigh64"] - 1)
        res.update(loss_last=abs(losses[-1] / d["loss64"][-1] - 1), norm_max=e_n.max(), norm_med=np.median(e_n),
                   ray_max=e_r.max(), ray_med=np.median(e_r))
    O.np.einsum = REAL_EINSUM
    return res


def gemm_calibration(nbits=25):
    rng = np.random.default_rng(5)
    for K in (128, 512, 2048):
        A = rng.uniform(-1, 1, (256, K))
        Bm = 0.03 * rng.standard_normal((K, 512))
        ref = A @ Bm
        for fmt in ("bb", "hh"):
            e = Emu(fmt, chain=K, nbits=nbits)
            D = e.mm(A, Bm)
            print(f"K={K} fmt={fmt} nbits={nbits}: rel_err {np.linalg.norm(D - ref) / np.linalg.norm(ref):.3e}")


if __name__ == "__main__":
    if sys.argv[1] == "calib":
        gemm_calibration(int(sys.argv[2]) if len(sys.argv) > 2 else 25)
    else:
        r = run(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
        print(" ".join(f"{k}={v:.2e}" if isinstance(v, float) else f"{k}={v}" for k, v in r.items()))
