"""CPU oracle for the NestedLoRA training step (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference's algorithm for the hot path
named in BASELINE.json (`methods/nestedlora.py` applied to the 2D Schroedinger
operators of `examples/operator`).  It is the *checker*: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product (`neural_svd_b200`) never does.

Parity pinning: the reference ships no tests or golden vectors,
so this oracle is pinned against the *reference itself*, imported unmodified
from /root/reference in the build container by `oracle/make_golden.py`
(exact-Laplacian mode, fp32 and fp64); the outputs are committed under
`tests/golden/` and `tests/test_oracle_golden.py` checks this file against them.

Every function cites the reference file:line it restates (paths relative to
the reference root).  dtype follows the inputs: pass float64 arrays for the
"truth" run, float32 arrays to mimic the reference's fp32 arithmetic order
loosely (not bit-exact: BLAS summation order differs).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------
# configuration of the path (what the reference spreads over argparse flags)
# --------------------------------------------------------------------------
@dataclass
class PathConfig:
    """Hyper-parameters of one problem instance.

    Mirrors the fields read by `get_problem` (examples/operator/pde/problems.py:23-130),
    `get_wavefunctions` (examples/operator/pde/__init__.py:19-55) and the Gaussian
    sampler/importance (examples/operator/pde/main_pde.py:89-100).
    """
    potential: str = "hydrogen"          # 'hydrogen' | 'harmonic_oscillator' | 'hydrogen_mol_ion' | 'infinite_well' | 'cosine'
    ndim: int = 2
    neigs: int = 16
    charge: float = 1.0                  # potentials.py:5-8
    k: float = 1.0                       # potentials.py:24-27
    scale_kinetic: float = 1.0           # problems.py:28
    operator_scale: float = 100.0        # examples/__init__.py:7-9
    operator_shift: float = 0.0
    sampling_scale: float = 16.0         # main_pde.py:92-100
    fourier_mapping_size: int = 1024     # examples/utils.py:102-124
    fourier_scale: float = 0.1
    hidden: Tuple[int, ...] = (128, 128, 128)
    hard_mul_const: float = 1.0          # pde/__init__.py:15-16
    apply_exp_mask: bool = False         # pde/boundary.py:39-53
    exp_mask_init_scale: float = 100.0
    sequential: bool = False             # nestedlora.py:183-192
    step: int = 1
    sampling_mode: str = "gaussian"      # 'gaussian' | 'laplacian' | 'uniform' | 'none'  (main_pde.py:89-118)
    apply_boundary: bool = False         # DirichletBoundaryMaskBox(lim, boundary_mode), pde/boundary.py:16-37
    boundary_mode: str = "dir_box_sqrt"
    lim: float = 50.0
    fourier_deterministic: bool = False  # utils.py:106-113
    hydrogen_mol_ion_R: float = 1.0      # problems.py:71

    @staticmethod
    def hydrogen(**kw) -> "PathConfig":
        # scripts/exps/pde/hydrogen.sh:11-65
        return PathConfig(**kw)

    @staticmethod
    def oscillator(**kw) -> "PathConfig":
        # scripts/exps/pde/oscillator.sh:11-67
        base = dict(potential="harmonic_oscillator", operator_scale=1.0, operator_shift=16.0,
                    sampling_scale=4.0, fourier_mapping_size=256, fourier_scale=1.0,
                    apply_exp_mask=True, exp_mask_init_scale=10.0)
        base.update(kw)
        return PathConfig(**base)


# --------------------------------------------------------------------------
# nesting masks  (methods/nestedlora.py:40-54, 183-192, 345-356)
# --------------------------------------------------------------------------
def nesting_masks(neigs: int, sequential: bool, step: int = 1,
                  set_first_mode_const: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """vector_mask (L',) and matrix_mask (L',L') in float32, as the reference builds them."""
    if sequential:
        L = neigs + (1 if set_first_mode_const else 0)          # nestedlora.py:49-54
        return np.ones(L, np.float32), np.triu(np.ones((L, L), np.float32))
    end_indices = list(range(step, neigs + 1, step))            # nestedlora.py:186-191
    if neigs not in end_indices:
        end_indices.append(neigs)
    w = np.zeros(neigs)
    w[np.array(end_indices) - 1] = 1.0
    w = w / w.sum()
    v = np.cumsum(w[::-1])[::-1]                                # nestedlora.py:40-46
    if set_first_mode_const:
        v = np.concatenate([v[:1], v])
    v = v.astype(np.float32)
    M = np.minimum(v[:, None], v[None, :]).astype(np.float32)
    return v, M


# --------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------
def param_names(cfg: PathConfig) -> List[str]:
    """`named_parameters()` order of NestedLoRA(WaveFunctions(ParallelMLP))."""
    n = ["model.base.feature_map._B"]
    nl = len(cfg.hidden) + 1
    n += [f"model.base.ws.{i}" for i in range(nl)]
    n += [f"model.base.bs.{i}" for i in range(nl)]
    if cfg.apply_exp_mask:
        n.append("model.boundary_mask.scales")
    return n


# --------------------------------------------------------------------------
# forward: Fourier features -> 4-stream MLP -> operator
# --------------------------------------------------------------------------
def softplus_streams(z: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """softplus(beta=1, threshold=20) value and sigmoid (torch.nn.Softplus, mlp.py:86)."""
    big = z > 20.0
    zs = np.where(big, 0.0, z)
    a = np.where(big, z, np.maximum(zs, 0) + np.log1p(np.exp(-np.abs(zs))))
    sig = np.where(big, 1.0, 1.0 / (1.0 + np.exp(-zs)))
    return a.astype(z.dtype), sig.astype(z.dtype)


def forward_streams(x: np.ndarray, params: Dict[str, np.ndarray], cfg: PathConfig,
                    keep: bool = False):
    """Value / gradient / Laplacian streams of the L parallel MLPs.

    Restates GaussianFourierFeatureTransform.forward (examples/utils.py:126-143) and
    ParallelMLP.forward (examples/models/mlp.py:204-221) in forward mode; equals the
    reference's autograd `exact_laplacian` (pde/diff_ops.py:54-93) to 1e-15 in fp64.

    Returns u (S=D+2, B, L): streams [value, d/dx_1..d/dx_D, Laplacian] of the raw
    network output, plus (if keep) the value-stream activations for the backward.
    """
    dt = x.dtype
    D = cfg.ndim
    Bff = params["model.base.feature_map._B"].astype(dt)            # (D, M)
    p = x @ Bff                                                     # utils.py:139
    s, c = np.sin(p), np.cos(p)
    S = D + 2
    h = [np.concatenate([s, c], 1)]                                 # utils.py:140
    for d in range(D):
        h.append(np.concatenate([c * Bff[d], -s * Bff[d]], 1))
    b2 = (Bff ** 2).sum(0)
    h.append(np.concatenate([-s * b2, -c * b2], 1))
    h = np.stack(h, 0)                                              # (S, B, 2M)
    h = np.broadcast_to(h[:, None], (S, cfg.neigs) + h.shape[1:])   # (S, L, B, K)
    acts = [h[0, 0]]                                                # Phi (B, 2M), shared by copies
    nl = len(cfg.hidden) + 1
    sigs = []
    for i in range(nl):
        W = params[f"model.base.ws.{i}"].astype(dt)                 # (L, H, K)
        b = params[f"model.base.bs.{i}"].astype(dt)                 # (L, H, 1)
        z = np.einsum("slbk,lhk->slbh", h, W)                       # mlp.py:207-218
        z[0] += b[:, None, :, 0]
        if i < nl - 1:
            a, sig = softplus_streams(z[0])                         # mlp.py:212,220
            out = np.empty_like(z)
            out[0] = a
            for d in range(D):
                out[1 + d] = sig * z[1 + d]
            out[S - 1] = sig * z[S - 1] + sig * (1 - sig) * (z[1:1 + D] ** 2).sum(0)
            h = out
            acts.append(a)
            sigs.append(sig)
        else:
            h = z
    u = np.transpose(h[..., 0], (0, 2, 1))                          # (S, B, L)   mlp.py:221
    if keep:
        return u, acts, sigs
    return u


COSINE_CS_2D = (0.814723686393179, 0.905791937075619)          # problems.py:46


def importance_terms(x: np.ndarray, cfg: PathConfig):
    """(w, grad ln sqrt(w), Lap ln sqrt(w)) of the sampler's density (main_pde.py:89-118); `none` = no re-weighting
    (importance=None, diff_ops.py:10-11).  d|x|/dx = sign(x) with a vanishing second derivative, as autograd has it."""
    B, D = x.shape
    s = cfg.sampling_scale
    if cfg.sampling_mode == "gaussian":
        return importance_gaussian(x, s), -x / (2 * s ** 2), np.full((B,), -D / (2 * s ** 2), x.dtype)
    if cfg.sampling_mode == "laplacian":
        w = np.exp(-np.abs(x).sum(1) / s - D * math.log(2 * s)).astype(x.dtype)
        return w, -np.sign(x) / (2 * s), np.zeros((B,), x.dtype)
    if cfg.sampling_mode == "uniform":
        return np.full((B,), 1.0 / (2 * s) ** D, x.dtype), np.zeros_like(x), np.zeros((B,), x.dtype)
    if cfg.sampling_mode == "none":
        return None, np.zeros_like(x), np.zeros((B,), x.dtype)
    raise NotImplementedError(cfg.sampling_mode)


def box_mask_terms(x: np.ndarray, cfg: PathConfig):
    """DirichletBoundaryMaskBox (pde/boundary.py:16-37): mask (B,), its gradient (B,D) and Laplacian (B,) as autograd
    differentiates it (clamp / maximum pass no gradient outside the box)."""
    B, D = x.shape
    if not cfg.apply_boundary:
        return np.ones((B,), x.dtype), np.zeros_like(x), np.zeros((B,), x.dtype)
    lim = cfg.lim
    inside = (x >= -lim) & (x <= lim)
    xc = np.clip(x, -lim, lim)
    if cfg.boundary_mode == "dir_box_sqrt":
        q = 2 * lim ** 2 - xc ** 2
        t = (np.sqrt(q) - lim) / lim
        on = inside & (t > 0)
        m = np.maximum(t, 0.0)
        d1 = np.where(on, -xc / (lim * np.sqrt(q)), 0.0)
        d2 = np.where(on, -2 * lim / (q * np.sqrt(q)), 0.0)
    elif cfg.boundary_mode == "dir_box_exp":
        a, b = np.exp(xc - lim), np.exp(-xc - lim)
        m = (1 - a) * (1 - b)
        d1 = np.where(inside, b - a, 0.0)
        d2 = np.where(inside, -(a + b), 0.0)
    else:
        raise NotImplementedError(cfg.boundary_mode)
    mask = m.prod(1)
    grad = np.stack([d1[:, i] * np.delete(m, i, 1).prod(1) for i in range(D)], 1)
    lap = sum(d2[:, i] * np.delete(m, i, 1).prod(1) for i in range(D))
    return mask.astype(x.dtype), grad.astype(x.dtype), lap.astype(x.dtype)


def importance_gaussian(x: np.ndarray, sigma: float) -> np.ndarray:
    """N(x; 0, sigma^2 I) as MultivariateNormal.log_prob().exp() (main_pde.py:94-100)."""
    D = x.shape[1]
    logw = -(x ** 2).sum(1) / (2 * sigma ** 2) - 0.5 * D * math.log(2 * math.pi * sigma ** 2)
    return np.exp(logw).astype(x.dtype)


def potential(x: np.ndarray, cfg: PathConfig) -> np.ndarray:
    r = np.sqrt((x ** 2).sum(1))
    if cfg.potential == "hydrogen":
        return -(cfg.charge / r)                                   # potentials.py:5-8
    if cfg.potential == "harmonic_oscillator":
        return cfg.k * r ** 2                                       # potentials.py:24-27
    if cfg.potential == "hydrogen_mol_ion":                         # potentials.py:11-17, charge = 2 * args.charge
        e = np.zeros((x.shape[1],), x.dtype)
        e[-1] = cfg.hydrogen_mol_ion_R
        Z = 2 * cfg.charge
        return -Z / np.sqrt(((x - e) ** 2).sum(1)) - Z / np.sqrt(((x + e) ** 2).sum(1))
    if cfg.potential == "infinite_well":
        return np.zeros((x.shape[0],), x.dtype)                     # potentials.py:20-21
    if cfg.potential == "cosine":
        cs = np.asarray(COSINE_CS_2D, np.float32).astype(x.dtype)   # torch.tensor(cs) is fp32 whatever x is
        return (np.cos(x) * cs[None, :]).sum(1)                     # potentials.py:30-31
    raise NotImplementedError(cfg.potential)


def operator_apply(x: np.ndarray, u: np.ndarray, params: Dict[str, np.ndarray], cfg: PathConfig):
    """(Tf, f, aux) from the raw streams u (S,B,L).

    Restates WaveFunctions.forward (pde/__init__.py:15-16), ExponentialMask.forward
    (pde/boundary.py:46-53), VectorizedLaplacian.__call__ with importance
    (pde/diff_ops.py:9-23), NegativeHamiltonian.__call__ (pde/schrodinger/__init__.py:16-22)
    and OperatorWrapper.__call__ (examples/__init__.py:7-9) via the product rule on
    q = sqrt(w) * mask_l  (SURVEY.md §8a).
    """
    dt = x.dtype
    D = cfg.ndim
    B, L = x.shape[0], cfg.neigs
    r = np.sqrt((x ** 2).sum(1))[:, None]                           # (B,1)
    w, gq, lq = importance_terms(x, cfg)
    if w is None:
        rho = np.ones((B, 1), dt)
    else:
        sqrt_w = np.sqrt(w)[:, None]
        rho = sqrt_w / np.maximum(sqrt_w, 1e-5)                     # diff_ops.py:15-18
    gradQ = gq[:, None, :] * np.ones((1, L, 1), dt)                 # (B,L,D)
    lapQ = lq[:, None] * np.ones((1, L), dt)
    if cfg.apply_exp_mask:
        sc = params["model.boundary_mask.scales"].astype(dt)[None, :]            # (1,L)
        m = np.exp(-r / sc)                                         # boundary.py:48-49
        gradQ = gradQ - x[:, None, :] / (r * sc)[:, :, None]
        lapQ = lapQ - (D - 1) / (r * sc)
    else:
        m = np.ones((B, L), dt)
    mb, gmb, lmb = box_mask_terms(x, cfg)                           # boundary.py:16-37 (and :50-51 under the exp mask)
    ce = cfg.hard_mul_const * m * rho
    uv = u[0]
    gu = np.stack([u[1 + d] for d in range(D)], -1)                 # (B,L,D)
    inner = u[D + 1] + 2 * (gradQ * gu).sum(-1) + uv * (lapQ + (gradQ ** 2).sum(-1))
    inner = (mb[:, None] * inner + 2 * (gmb[:, None, :] * (gu + uv[..., None] * gradQ)).sum(-1)
             + uv * lmb[:, None])
    lap = ce * inner
    m = m * mb[:, None]                                             # total mask (backward: df/du = c m rho)
    f = ce * mb[:, None] * uv
    V = potential(x, cfg)[:, None]
    negH = cfg.scale_kinetic * lap - V * f                          # schrodinger/__init__.py:19-22
    Tf = cfg.operator_scale * negH + cfg.operator_shift * f         # examples/__init__.py:9
    aux = dict(m=m, rho=rho, r=r)
    return Tf.astype(dt), f.astype(dt), aux


# --------------------------------------------------------------------------
# loss forward / custom backward  (methods/nestedlora.py:57-111)
# --------------------------------------------------------------------------
def chunk_sizes(B: int) -> Tuple[int, int]:
    """torch.chunk(f, 2) row split (nestedlora.py:263): ceil(B/2), rest."""
    b1 = (B + 1) // 2
    return b1, B - b1


def gram_terms(f: np.ndarray, Tf: np.ndarray, v: np.ndarray, b1: Optional[int] = None):
    """Un-normalised sums [f1^T f1, f2^T f2, sum_b sum_l v_l f Tf] (the all-reduced buffer)."""
    B = f.shape[0]
    if b1 is None:
        b1 = chunk_sizes(B)[0]
    f1, f2 = f[:b1], f[b1:]
    return f1.T @ f1, f2.T @ f2, float((v[None, :] * f * Tf).sum())


def loss_from_terms(G1, G2, opsum, B, B1, B2, M):
    lam1, lam2 = G1 / B1, G2 / B2                                   # nestedlora.py:10-11
    loss_metric = (M * lam1 * lam2).sum()                           # nestedlora.py:64
    loss_operator = -2.0 * opsum / B                                # nestedlora.py:92
    return loss_operator + loss_metric, lam1, lam2


def loss_forward(f, Tf, v, M):
    B = f.shape[0]
    B1, B2 = chunk_sizes(B)
    G1, G2, opsum = gram_terms(f, Tf, v.astype(f.dtype), B1)
    loss, lam1, lam2 = loss_from_terms(G1, G2, opsum, B, B1, B2, M.astype(f.dtype))
    return loss, lam1, lam2


def loss_dF(f, Tf, v, M, lam1, lam2, B=None, B1=None, B2=None, b1_local=None):
    """Total gradient w.r.t. f implied by the reference's *custom* backward
    (nestedlora.py:98-111): dF = -(4/B) v Tf + [ (2/B1) f1 (M*lam2) ; (2/B2) f2 (M*lam1) ].
    Tf receives no gradient.  (B, B1, B2 are the GLOBAL counts in data-parallel use.)"""
    n = f.shape[0]
    if B is None:
        B = n
        B1, B2 = chunk_sizes(B)
    if b1_local is None:
        b1_local = chunk_sizes(n)[0]
    dt = f.dtype
    dF = -(4.0 / B) * v.astype(dt)[None, :] * Tf
    dF[:b1_local] += (2.0 / B1) * f[:b1_local] @ (M.astype(dt) * lam2)
    dF[b1_local:] += (2.0 / B2) * f[b1_local:] @ (M.astype(dt) * lam1)
    return dF


# --------------------------------------------------------------------------
# MLP backward (value stream only: the implicit autograd pass of SURVEY §8 a12)
# --------------------------------------------------------------------------
def mlp_backward(x, dF, params, cfg: PathConfig, u0, acts, sigs, aux) -> Dict[str, np.ndarray]:
    dt = x.dtype
    L = cfg.neigs
    grads: Dict[str, np.ndarray] = {}
    cm = cfg.hard_mul_const * aux["m"] * aux["rho"]
    du = dF * cm                                                    # (B,L)
    if cfg.apply_exp_mask:
        sc = params["model.boundary_mask.scales"].astype(dt)
        # d f / d s_l = c * rho * u * m * r / s_l^2
        grads["model.boundary_mask.scales"] = (dF * cm * u0 * aux["r"] / sc[None, :] ** 2).sum(0)
    nl = len(cfg.hidden) + 1
    dz = du.T[:, :, None]                                           # (L,B,1)
    for i in reversed(range(nl)):
        W = params[f"model.base.ws.{i}"].astype(dt)                 # (L,H,K)
        a_prev = acts[i]                                            # (B,K) for i=0 else (L,B,K)
        if i == 0:
            grads[f"model.base.ws.{i}"] = np.einsum("lbh,bk->lhk", dz, a_prev)
        else:
            grads[f"model.base.ws.{i}"] = np.einsum("lbh,lbk->lhk", dz, a_prev)
        grads[f"model.base.bs.{i}"] = dz.sum(1)[:, :, None]
        if i > 0:
            da = np.einsum("lbh,lhk->lbk", dz, W)
            dz = da * sigs[i - 1]
    return grads


def weighted_values(y: np.ndarray, params: Dict[str, np.ndarray], cfg: PathConfig) -> np.ndarray:
    """g(y) = sqrt(w(y)) * WaveFunctions(y)  (B, L): the function the finite-difference Laplacian differentiates
    (diff_ops.py:13; without importance g = f, diff_ops.py:10-11)."""
    dt = y.dtype
    u0 = forward_streams(y, params, cfg)[0]                          # (B, L) raw network values
    r = np.sqrt((y ** 2).sum(1))[:, None]
    m = np.ones_like(u0)
    if cfg.apply_exp_mask:
        m = np.exp(-r / params["model.boundary_mask.scales"].astype(dt)[None, :])      # boundary.py:48-49
    mb, _, _ = box_mask_terms(y, cfg)
    w, _, _ = importance_terms(y, cfg)
    sw = np.sqrt(w)[:, None] if w is not None else 1.0
    return (sw * (cfg.hard_mul_const * m * mb[:, None] * u0)).astype(dt)


def operator_apply_fd(x: np.ndarray, params: Dict[str, np.ndarray], cfg: PathConfig, eps: float):
    """(Tf, f) with the FINITE-DIFFERENCE Laplacian the shipped scripts use (`laplacian_eps > 0`):
    VectorizedLaplacian.approx_laplacian (pde/diff_ops.py:25-52) of g = sqrt(w) f, divided by clamp(sqrt w, 1e-5)
    (:15-18), then NegativeHamiltonian / OperatorWrapper as in operator_apply.  The shift vector is built in fp32
    (`torch.zeros((1, D))`, diff_ops.py:43-44) whatever x's dtype is, the division is by the python float eps**2."""
    dt = x.dtype
    D = cfg.ndim
    g0 = weighted_values(x, params, cfg)
    lap = -2.0 * D * g0                                             # diff_ops.py:40
    sh = np.float32(eps).astype(dt)
    for i in range(D):
        e = np.zeros((1, D), dt)
        e[0, i] = sh
        lap = lap + (weighted_values(x + e, params, cfg) + weighted_values(x - e, params, cfg))
    lap = lap / dt.type(eps ** 2)                                   # diff_ops.py:48
    w, _, _ = importance_terms(x, cfg)
    if w is not None:
        sw = np.maximum(np.sqrt(w)[:, None], 1e-5)                  # diff_ops.py:15
        lap, f = lap / sw, g0 / sw
    else:
        f = g0
    V = potential(x, cfg)[:, None]
    negH = cfg.scale_kinetic * lap - V * f
    Tf = cfg.operator_scale * negH + cfg.operator_shift * f
    return Tf.astype(dt), f.astype(dt)


def train_step(x: np.ndarray, params: Dict[str, np.ndarray], cfg: PathConfig, sort_indices=None,
               laplacian_eps: float = 0.0):
    """One loss+grad evaluation == reference `compute_loss_operator` + `loss.backward()`
    (nestedlora.py:254-267, operator/__init__.py:62-68) with the exact Laplacian (`laplacian_eps <= 0`) or the
    finite-difference one.  Either way the gradient flows through the CENTRAL evaluation's value stream only
    (the custom backward returns None for Tf, nestedlora.py:98-111).
    `sort_indices`: the permutation NestedLoRA.forward applies to the model output in training mode after
    register_eigvals() (nestedlora.py:195-206); f, Tf, dF are returned in the permuted column order."""
    u, acts, sigs = forward_streams(x, params, cfg, keep=True)
    Tf, f, aux = operator_apply(x, u, params, cfg)
    if laplacian_eps > 0:
        Tf, f = operator_apply_fd(x, params, cfg, laplacian_eps)
    if sort_indices is not None:
        si = np.asarray(sort_indices)
        f, Tf = f[:, si], Tf[:, si]
    v, M = nesting_masks(cfg.neigs, cfg.sequential, cfg.step)
    loss, lam1, lam2 = loss_forward(f, Tf, v, M)
    dF = loss_dF(f, Tf, v, M, lam1, lam2)
    dF_net = dF
    if sort_indices is not None:
        dF_net = np.empty_like(dF)
        dF_net[:, si] = dF                                          # backward of the column gather
    grads = mlp_backward(x, dF_net, params, cfg, u[0], acts, sigs, aux)
    return dict(loss=loss, f=f, Tf=Tf, dF=dF, grads=grads, lam1=lam1, lam2=lam2)


# --------------------------------------------------------------------------
# CDK loss  (methods/nestedlora.py:270-332)
# --------------------------------------------------------------------------
def cdk_forward_backward(f: np.ndarray, g: np.ndarray, neigs: int, sequential: bool = False,
                         step: int = 1, set_first_mode_const: bool = True, diagnostics: bool = True):
    dt = f.dtype
    v, M = nesting_masks(neigs, sequential, step, set_first_mode_const)
    v, M = v.astype(dt), M.astype(dt)
    if set_first_mode_const:                                        # nestedlora.py:287-290
        one = np.ones((f.shape[0], 1), dt)
        fp, gp = np.concatenate([one, f], 1), np.concatenate([one, g], 1)
    else:
        fp, gp = f, g
    B = fp.shape[0]
    lam_f, lam_g = fp.T @ fp / B, gp.T @ gp / B
    loss_metric = (M * lam_f * lam_g).sum()
    loss_operator = -2.0 * (v[None, :] * fp * gp).sum(1).mean()     # nestedlora.py:301
    out = dict(loss=loss_operator + loss_metric, loss_operator=loss_operator, loss_metric=loss_metric)
    if diagnostics:                                                 # nestedlora.py:303-305
        G = fp @ gp.T
        out["rs_joint"] = np.diag(G).copy()
        n = G.shape[0]
        out["rs_indep"] = G.flatten()[:-1].reshape(n - 1, n + 1)[:, 1:].flatten()  # methods/utils.py:16-22
    gf = -(2.0 / B) * v[None, :] * gp + (2.0 / B) * fp @ (M * lam_g)   # nestedlora.py:320-327
    gg = -(2.0 / B) * v[None, :] * fp + (2.0 / B) * gp @ (M * lam_f)
    if set_first_mode_const:
        gf, gg = gf[:, 1:], gg[:, 1:]
    out["grad_f"], out["grad_g"] = gf, gg
    return out


# --------------------------------------------------------------------------
# spectrum evaluation  (methods/spectrum.py:29-102), uniform validation importance
# --------------------------------------------------------------------------
def spectrum_evd(xs: np.ndarray, params, cfg: PathConfig, lim: float, chunk: int = 4096, normalize: bool = False,
                 sort: bool = False, post_align: bool = False):
    """compute_spectrum_evd (methods/spectrum.py:29-102) with the training importance of `cfg` and the uniform
    validation importance of main_pde.py:128-129 on [-lim, lim]^D.  Chunked like the reference's dataloader so the
    accumulation order of cov / quad is the same."""
    dt = xs.dtype
    L = cfg.neigs
    cov = np.zeros((L, L), dt)
    quad = np.zeros((L, L), dt)
    eigfuncs = []
    for i in range(0, len(xs), chunk):
        x = xs[i:i + chunk]
        u = forward_streams(x, params, cfg)
        Tf, f, _ = operator_apply(x, u, params, cfg)
        w, _, _ = importance_terms(x, cfg)
        sw_tr = np.sqrt(w)[:, None] if w is not None else np.ones((len(x), 1), dt)       # spectrum.py:16-27
        # importance_val builds its constant in fp32 (main_pde.py:128-129); the square root is taken in x's dtype
        sw_va = float(np.sqrt(np.float32(1.0 / (2 * lim) ** cfg.ndim).astype(dt)))
        sw = sw_tr / sw_va                                          # spectrum.py:61
        eigfuncs.append(sw_tr * f)                                  # spectrum.py:65
        phi, Tphi = np.nan_to_num(sw * f), np.nan_to_num(sw * Tf)   # spectrum.py:66-72
        Tphi[np.all(np.isclose(x, 0.0), axis=1)] *= 0.0             # spectrum.py:73
        cov += phi.T @ phi
        quad += phi.T @ Tphi
    cov /= len(xs)
    quad /= len(xs)
    out = dict(eigfuncs=np.concatenate(eigfuncs, 0), cov=cov, quad=quad)
    out["eigvals"] = eigvals = np.diag(quad) / np.diag(cov)         # spectrum.py:86
    out["norms"] = norms = np.diag(cov)                             # spectrum.py:87
    if normalize:                                                   # spectrum.py:88-90
        out["cov"] = cov / (np.sqrt(norms[:, None]) @ np.sqrt(norms[:, None]).T)
        out["eigfuncs"] = out["eigfuncs"] / np.sqrt(norms).reshape(1, -1)
    if sort:                                                        # spectrum.py:91-97
        si = np.argsort(eigvals)[::-1]
        out["eigvals"] = out["eigvals"][si]
        out["eigfuncs"] = out["eigfuncs"][:, si]
        out["cov"] = out["cov"][:, si][si, :]
        out["quad"] = out["quad"][:, si][si, :]
        out["norms"] = out["norms"][si]
    if post_align:                                                  # spectrum.py:98-101, 161-170
        from scipy.linalg import eigh
        ec, Vc = eigh(out["cov"])
        whitening = Vc @ np.diag(1 / np.sqrt(ec)) @ Vc.T
        ev, V = eigh(whitening @ out["quad"] @ whitening)
        out["eigvals_aligned"] = np.sqrt(ev[::-1])
        V = V[:, ::-1]
        out["eigfuncs_aligned"] = out["eigfuncs"] @ (V.T @ whitening).T
        out["cov_aligned"] = np.eye(L)
    return out


# --------------------------------------------------------------------------
# deterministic parameter construction in the reference's RNG draw order
# --------------------------------------------------------------------------
def init_params_like_reference(cfg: PathConfig, seed: int) -> Dict[str, np.ndarray]:
    """Same draws as torch.manual_seed(seed); get_wavefunctions(args) (pde/__init__.py:19-55):
    `_B` first (utils.py:116-118), then ws[0..] (mlp.py:186-188); biases zero; mask scales const."""
    import torch
    g = torch.Generator().manual_seed(seed)
    D, M = cfg.ndim, cfg.fourier_mapping_size
    out = {}
    if cfg.fourier_deterministic:                                   # utils.py:106-113 (no RNG draw)
        out["model.base.feature_map._B"] = (cfg.fourier_scale * torch.cat(
            [i * torch.eye(D) for i in range(1, M + 1)], dim=0).T).numpy()
        M = D * M
    else:
        out["model.base.feature_map._B"] = (2 * torch.pi * cfg.fourier_scale *
                                            torch.randn((D, M), generator=g).float()).numpy()
    prev = 2 * M
    dims = list(cfg.hidden) + [1]
    for i, h in enumerate(dims):
        out[f"model.base.ws.{i}"] = (math.sqrt(2.0 / prev) *
                                     torch.randn(cfg.neigs, h, prev, generator=g)).numpy()
        out[f"model.base.bs.{i}"] = np.zeros((cfg.neigs, h, 1), np.float32)
        prev = h
    if cfg.apply_exp_mask:
        out["model.boundary_mask.scales"] = np.full((cfg.neigs,), cfg.exp_mask_init_scale, np.float32)
    return out
