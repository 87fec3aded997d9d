"""CPU restatement of the fp16 hi/lo operand plan of the tensor-core engine (evidence for DESIGN.md §3 and the spec the
CUDA `plan` kernels follow; not product code, never imported by the product).

Every MMA operand v is stored as two fp16 planes of s*v (hi = fp16(s v), lo = fp16(s v - hi)), s a power of two chosen
from a RIGOROUS bound of |v| (so that |s v| <= 2^15 can never overflow fp16): 22 significant bits for |s v| >= 2^-3 and an
absolute error of 2^-25 below.  The bounds are propagated from the weights alone (forward) and from max|dF| (backward):

  features          |phi_s[j]| <= c_s[j]              c = (1, |B_0j|, |B_1j|, B_0j^2 + B_1j^2)
  layer 0           |z0_s[h]| <= R0_s = max_h sum_j (|W0[h,j]| + |W0[h,M+j]|) c_s[j]  (+ max|b0| for s = 0)
  softplus streams  |a_0| <= z_0 + ln 2,  |a_d| <= z_d,  |a_3| <= z_3 + (z_1^2 + z_2^2)/4
  hidden layer i    |z_s| <= (max_h sum_k |W_i[h,k]|) A_s  (+ max|b_i| for s = 0)
  backward          |dZ2| <= |c| max|dF| max|W3|,  |dZ_{i-1}| <= (max_k sum_j |W_i[j,k]|) |dZ_i|

usage: python profiles/scale_plan_emulation.py [hydrogen|oscillator] [B]
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import nsvd_oracle as O  # noqa: E402

TARGET = 2.0 ** 15
LN2 = math.log(2.0)


def pow2_scale(bound):
    b = np.maximum(np.asarray(bound, np.float64) * 1.0001, 1e-30)
    return np.where(np.asarray(bound) > 0, 2.0 ** np.floor(np.log2(TARGET / b)), 1.0)


def hh(v, s):
    """fp16 hi + fp16 lo planes of s*v, returned as the fp64 value they represent, divided by s again."""
    w = (np.asarray(v, np.float64) * s).astype(np.float32)
    assert np.abs(w).max() <= 65504, "fp16 overflow: the bound was not a bound"
    hi = w.astype(np.float16).astype(np.float32)
    lo = (w - hi).astype(np.float16).astype(np.float32)
    return (hi.astype(np.float64) + lo.astype(np.float64)) / s


def forward_plan(params, cfg):
    """per-copy bounds and scales of every forward operand."""
    Bff = params["model.base.feature_map._B"].astype(np.float64)
    M = Bff.shape[1]
    c = np.stack([np.ones(M), np.abs(Bff[0]), np.abs(Bff[1]), (Bff ** 2).sum(0)])          # (4, M)
    W0 = np.abs(params["model.base.ws.0"].astype(np.float64))
    pair = W0[:, :, :M] + W0[:, :, M:]                                                      # (L, H, M)
    R0 = np.einsum("lhj,sj->lsh", pair, c).max(-1)                                          # (L, 4)
    plan = {"w0": pow2_scale(W0.max((1, 2))[:, None] * c.max(1)[None, :])}                  # folded weights (L, 4)
    Z = R0.copy()
    Z[:, 0] += np.abs(params["model.base.bs.0"]).max((1, 2))
    bounds = {}
    for i in range(3):
        A = Z.copy()
        A[:, 0] = Z[:, 0] + LN2
        A[:, 3] = Z[:, 3] + 0.25 * (Z[:, 1] ** 2 + Z[:, 2] ** 2)
        bounds[f"a{i}"] = A
        plan[f"a{i}"] = pow2_scale(A)
        if i < 2:
            W = np.abs(params[f"model.base.ws.{i + 1}"].astype(np.float64))
            plan[f"w{i + 1}"] = pow2_scale(W.max((1, 2)))
            R = W.sum(2).max(1)                                                             # row L1, max over rows
            Z = R[:, None] * A
            Z[:, 0] += np.abs(params[f"model.base.bs.{i + 1}"]).max((1, 2))
    return plan, bounds


def run(prob="hydrogen", B=512):
    cfg = getattr(O.PathConfig, prob)(neigs=16)
    params = {k: v.astype(np.float64) for k, v in O.init_params_like_reference(cfg, 0).items()}
    rng = np.random.default_rng(1)
    x = (cfg.sampling_scale * rng.standard_normal((B, 2))).astype(np.float32).astype(np.float64)
    ref = O.train_step(x, params, cfg)
    plan, bounds = forward_plan(params, cfg)
    L, D = cfg.neigs, 2
    # ---- forward with quantised operands (accumulation exact)
    Bff = params["model.base.feature_map._B"]
    p = x @ Bff
    s, c = np.sin(p), np.cos(p)
    b2 = (Bff ** 2).sum(0)
    feats = [np.concatenate([s, c], 1), np.concatenate([c * Bff[0], -s * Bff[0]], 1),
             np.concatenate([c * Bff[1], -s * Bff[1]], 1), np.concatenate([-s * b2, -c * b2], 1)]
    phi_q = hh(feats[0], 1.0)                                   # the kernel multiplies [sin; cos] by FOLDED weights
    W0 = params["model.base.ws.0"]
    M = Bff.shape[1]
    Ws, Wc = W0[:, :, :M], W0[:, :, M:]
    fold = [np.concatenate([Ws, Wc], 2), np.concatenate([-Wc * Bff[0], Ws * Bff[0]], 2),
            np.concatenate([-Wc * Bff[1], Ws * Bff[1]], 2), np.concatenate([-b2 * Ws, -b2 * Wc], 2)]
    z = np.stack([np.einsum("bk,lhk->lbh", phi_q, np.stack([hh(fold[st][l], plan["w0"][l, st]) for l in range(L)]))
                  for st in range(4)])                          # (4, L, B, H)
    acts, sigs = [feats[0]], []
    report = []
    for i in range(3):
        z[0] += params[f"model.base.bs.{i}"][:, None, :, 0]
        a, sig = O.softplus_streams(z[0])
        out = np.empty_like(z)
        out[0] = a
        out[1], out[2] = sig * z[1], sig * z[2]
        out[3] = sig * z[3] + sig * (1 - sig) * (z[1] ** 2 + z[2] ** 2)
        acts.append(a)
        sigs.append(sig)
        amax = np.abs(out).max((2, 3)).T                        # (L, 4)
        report.append((f"a{i}", np.log2(bounds[f"a{i}"] / amax)))
        aq = np.stack([np.stack([hh(out[st, l], plan[f"a{i}"][l, st]) for l in range(L)]) for st in range(4)])
        if i < 2:
            W = params[f"model.base.ws.{i + 1}"]
            Wq = np.stack([hh(W[l], plan[f"w{i + 1}"][l]) for l in range(L)])
            z = np.einsum("slbk,lhk->slbh", aq, Wq)
        else:
            W3 = params["model.base.ws.3"]
            z = np.einsum("slbk,lhk->slbh", out, W3)            # head: fp32 registers in the kernel
            z[0] += params["model.base.bs.3"][:, None, :, 0]
        acts[-1] = aq[0]                                        # the backward reads the saved (quantised) value stream
    u = np.transpose(z[..., 0], (0, 2, 1))
    Tf, f, aux = O.operator_apply(x, u, params, cfg)
    v, Mm = O.nesting_masks(cfg.neigs, cfg.sequential, cfg.step)
    loss, lam1, lam2 = O.loss_forward(f, Tf, v, Mm)
    dF = O.loss_dF(f, Tf, v, Mm, lam1, lam2)
    # ---- backward
    cm = cfg.hard_mul_const * aux["m"] * aux["rho"]
    du = (dF * cm).T                                            # (L, B)
    mdF = np.abs(dF).max(0)
    W3 = params["model.base.ws.3"][:, 0, :]
    DZ = abs(cfg.hard_mul_const) * mdF * np.abs(W3).max(1)
    dz = du[:, :, None] * W3[:, None, :] * sigs[2]              # dZ2 (L, B, H)
    grads = {"model.base.ws.3": np.einsum("lb,lbh->lh", du, acts[3])[:, None, :],
             "model.base.bs.3": du.sum(1)[:, None, None]}
    for i in (2, 1, 0):
        s_dz = pow2_scale(DZ)
        report.append((f"dZ{i}", np.log2(DZ / np.abs(dz).max((1, 2)))[:, None]))
        dzq = np.stack([hh(dz[l], s_dz[l]) for l in range(L)])
        grads[f"model.base.bs.{i}"] = dz.sum(1)[:, :, None]
        if i == 0:
            grads["model.base.ws.0"] = np.einsum("lbh,bk->lhk", dzq, phi_q)
        else:
            grads[f"model.base.ws.{i}"] = np.einsum("lbh,lbk->lhk", dzq, acts[i])
            W = params[f"model.base.ws.{i}"]
            Wq = np.stack([hh(W[l], plan[f"w{i}"][l]) for l in range(L)])
            dz = np.einsum("lbh,lhk->lbk", dzq, Wq) * sigs[i - 1]
            DZ = np.abs(W).sum(1).max(1) * DZ                   # column L1, max over columns
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    print(f"{prob} B={B}: loss {abs(loss / ref['loss'] - 1):.2e} f {rel(f, ref['f']):.2e} Tf {rel(Tf, ref['Tf']):.2e} "
          f"dF {rel(dF, ref['dF']):.2e}")
    print("  grads:", " ".join(f"{k.split('.')[-2]}{k.split('.')[-1]}={rel(g.reshape(ref['grads'][k].shape), ref['grads'][k]):.1e}"
                               for k, g in sorted(grads.items())))
    print("  log2(bound / actual max), worst copy per stream (headroom lost to the bound):")
    for name, lg in report:
        print(f"    {name}: " + " ".join(f"{v:5.1f}" for v in lg.max(0)), " min", " ".join(f"{v:5.1f}" for v in lg.min(0)))


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "hydrogen", int(sys.argv[2]) if len(sys.argv) > 2 else 512)
