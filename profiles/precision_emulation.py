"""CPU emulation of the tensor-core operand formats (evidence for DESIGN.md §3; not product code).

Every dense contraction of the oracle is re-run with both operands rounded to a two-plane split
(hi, lo) and the product hi*hi + lo*hi + hi*lo accumulated in fp64, so only the OPERAND rounding
is modelled.  Modes: bf16+bf16 (the bf16x3 engine of round 1) and bf16+fp16 (hi plane bf16, lo plane
fp16: 8 + 11 significant bits).
usage: python profiles/precision_emulation.py [problem] [B]
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import nsvd_oracle as O


def bf16(v32):
    u = v32.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split(v, lo_fmt):
    v32 = v.astype(np.float32)
    hi = bf16(v32)
    r = v32 - hi
    lo = bf16(r) if lo_fmt == "bf16" else r.astype(np.float16).astype(np.float32)
    return hi.astype(np.float64), lo.astype(np.float64)


def make_einsum(lo_fmt, real):
    def e(spec, a, b):
        ah, al = split(np.ascontiguousarray(a), lo_fmt)
        bh, bl = split(np.ascontiguousarray(b), lo_fmt)
        return real(spec, ah, bh) + real(spec, al, bh) + real(spec, ah, bl)
    return e


def run(cfg, x, params, lo_fmt):
    real = np.einsum
    if lo_fmt is not None:
        O.np.einsum = make_einsum(lo_fmt, real)
    try:
        return O.train_step(x, params, cfg)
    finally:
        O.np.einsum = real


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


if __name__ == "__main__":
    prob = sys.argv[1] if len(sys.argv) > 1 else "hydrogen"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    cfg = getattr(O.PathConfig, prob)(neigs=16, sequential=True)
    params = {k: v.astype(np.float64) for k, v in O.init_params_like_reference(cfg, 0).items()}
    rng = np.random.default_rng(1)
    x = (cfg.sampling_scale * rng.standard_normal((B, 2))).astype(np.float32).astype(np.float64)
    ref = run(cfg, x, params, None)
    for fmt in ("bf16", "fp16"):
        out = run(cfg, x, params, fmt)
        g = {k: rel(out["grads"][k], ref["grads"][k]) for k in ref["grads"]}
        print(f"lo={fmt}: loss {abs(out['loss']/ref['loss']-1):.2e} f {rel(out['f'], ref['f']):.2e} "
              f"Tf {rel(out['Tf'], ref['Tf']):.2e} dF {rel(out['dF'], ref['dF']):.2e} grads max {max(g.values()):.2e} "
              + " ".join(f"{k.split('.')[-2]}{k.split('.')[-1]}={v:.1e}" for k, v in g.items()))
