"""Host-side mirror of the reference's operator objects for the hot path.

  hydrogen_potential / harmonic_oscillator_potential  pde/schrodinger/potentials.py:5-8,24-27
  NegativeHamiltonian                                  pde/schrodinger/__init__.py:4-22
  OperatorWrapper                                      examples/__init__.py:1-9
  GaussianImportance  (the closure of)                 pde/main_pde.py:94-100
  get_problem                                          pde/problems.py:23-130 (sch / hydrogen, oscillator)

`operator(method, x, importance) -> (Tf, f)` keeps the reference's callable protocol; the work is
done by the fused forward-mode kernel (exact Laplacian, no autograd double backward).
"""
from __future__ import annotations

import math
import warnings
from functools import partial

import numpy as np
import torch


def hydrogen_potential(x, charge=1.0):
    x = x.reshape(x.shape[0], -1)
    return -(charge / x.norm(dim=1, p=2)).reshape(-1, 1)


def harmonic_oscillator_potential(x, k=1.0):
    x = x.reshape(x.shape[0], -1)
    return (k * x.norm(dim=1, p=2) ** 2).reshape(-1, 1)


class GaussianImportance:
    """w(x) = N(x; 0, sigma^2 I).  Callable like the reference's `importance_train` closure."""

    def __init__(self, sampling_scale: float, dim: int = 2):
        self.sampling_scale = float(sampling_scale)
        self.dim = dim

    def __call__(self, x):
        x = x.reshape(x.shape[0], -1)
        s2 = self.sampling_scale ** 2
        logw = -(x ** 2).sum(1) / (2 * s2) - 0.5 * self.dim * math.log(2 * math.pi * s2)
        return logw.exp().view(-1, 1)


def make_gaussian_sampler(batch_size, sampling_scale, ndim=2, n_particles=1, generator=None):
    """`make_batch_ftn_train` of main_pde.py:92-93 (CPU randn, shape (B, n_particles, ndim))."""
    def make_batch():
        return sampling_scale * torch.randn((batch_size, n_particles, ndim), generator=generator)
    return make_batch


class NegativeHamiltonian:
    """-H f = kappa * Lap f - V f with importance re-weighting (schrodinger/__init__.py:4-22)."""

    def __init__(self, local_potential_ftn, scale_kinetic=1.0, laplacian_eps=1e-5, n_particles=1):
        self.laplacian_eps = laplacian_eps
        self.local_potential_ftn = local_potential_ftn
        self.scale_kinetic = scale_kinetic
        self.n_particles = n_particles

    def __call__(self, f, xs, importance=None, threshold=1e5):
        from . import fused
        return fused.apply_operator(f, OperatorWrapper(self, 1.0, 0.0), xs, importance)


class OperatorWrapper:
    """Tf = scale * T f + shift * f (examples/__init__.py:1-9)."""

    def __init__(self, operator, scale=1.0, shift=0.0):
        self.operator = operator
        self.scale = scale
        self.shift = shift

    def __call__(self, model, x, importance=None):
        from . import fused
        return fused.apply_operator(model, self, x, importance)


_FD_WARNED = False


def describe_operator(operator):
    """Recognise OperatorWrapper(NegativeHamiltonian(hydrogen|oscillator)) — by duck typing, so the
    reference's own objects are accepted too.  Anything else raises (no fallback)."""
    global _FD_WARNED
    inner = getattr(operator, "operator", None)
    if inner is None or not hasattr(operator, "scale") or not hasattr(operator, "shift"):
        raise NotImplementedError("fused path needs an OperatorWrapper(NegativeHamiltonian(...))")
    pot = getattr(inner, "local_potential_ftn", None)
    if pot is None or not hasattr(inner, "scale_kinetic"):
        raise NotImplementedError(f"unsupported operator {type(inner).__name__}: only NegativeHamiltonian")
    if getattr(inner, "n_particles", 1) != 1:
        raise NotImplementedError("n_particles > 1 is out of scope")
    func = pot.func if isinstance(pot, partial) else pot
    kw = pot.keywords if isinstance(pot, partial) else {}
    name = getattr(func, "__name__", "")
    if name == "hydrogen_potential":
        kind, coef = 0, float(kw.get("charge", 1.0))
    elif name == "harmonic_oscillator_potential":
        kind, coef = 1, float(kw.get("k", 1.0))
    else:
        raise NotImplementedError(f"unsupported potential {name!r}: only hydrogen / harmonic_oscillator")
    eps = getattr(inner, "laplacian_eps", 0.0)
    if eps is not None and eps > 0 and not _FD_WARNED:
        _FD_WARNED = True
        warnings.warn(f"laplacian_eps={eps}: the fused kernel evaluates the EXACT Laplacian in forward mode "
                      "(the limit the reference's finite difference approximates; see DESIGN.md)")
    return dict(potential=kind, pot_coef=coef, scale_kinetic=float(inner.scale_kinetic),
                op_scale=float(operator.scale), op_shift=float(operator.shift))


def describe_importance(importance, dim=2) -> float:
    """sigma of a Gaussian importance density; accepts GaussianImportance or the reference's closure
    over a zero-mean isotropic MultivariateNormal (main_pde.py:94-100)."""
    if importance is None:
        raise NotImplementedError("fused path needs the Gaussian importance density (sampling_mode='gaussian')")
    if hasattr(importance, "sampling_scale"):
        return float(importance.sampling_scale)
    for cell in (getattr(importance, "__closure__", None) or ()):
        obj = cell.cell_contents
        cov = getattr(obj, "covariance_matrix", None)
        if cov is not None and hasattr(obj, "loc"):
            cov = cov.detach().cpu().double()
            loc = obj.loc.detach().cpu().double()
            d = cov.shape[0]
            if d != dim or loc.abs().max() != 0 or not torch.allclose(cov, cov[0, 0] * torch.eye(d, dtype=cov.dtype)):
                raise NotImplementedError("importance must be a zero-mean isotropic Gaussian in 2D")
            return float(cov[0, 0].sqrt())
    raise NotImplementedError("cannot recognise the importance density; pass neural_svd_b200.GaussianImportance(sigma)")


def hydrogen2d_eigvals(neigs, charge=1.0):
    """E_n = -Z^2 / (4 (n + 1/2)^2) with degeneracy 2n+1 (schrodinger/ground_truths.py:120-132)."""
    out, n = [], 0
    while len(out) < neigs:
        out += [-(charge ** 2) / (4 * (n + 0.5) ** 2)] * (2 * n + 1)
        n += 1
    return np.array(out[:neigs])


def oscillator2d_eigvals(neigs, k=1.0):
    """2 sqrt(k) (n + 1) ... in the reference's convention H = -Lap + k r^2: 2n + ndim for k=1
    (ground_truths.py:78-90), degeneracy n+1 in 2D; truncated to neigs here."""
    out, n = [], 0
    while len(out) < neigs:
        out += [math.sqrt(k) * (2 * n + 2)] * (n + 1)
        n += 1
    return np.array(out[:neigs])


def get_problem(args, device=None):
    """pde/problems.py:23-130 for problem='sch', potential_type in {hydrogen, harmonic_oscillator}, ndim=2."""
    if args.problem != "sch" or args.ndim != 2:
        raise NotImplementedError("fused path covers the 2D Schroedinger problems")
    args.n_particles = 1
    if args.potential_type == "hydrogen":
        pot = partial(hydrogen_potential, charge=args.charge)
        gt = -hydrogen2d_eigvals(args.neigs, args.charge)
    elif args.potential_type == "harmonic_oscillator":
        pot = partial(harmonic_oscillator_potential, k=1.0)
        gt = -oscillator2d_eigvals(args.neigs, 1.0)
    else:
        raise NotImplementedError(args.potential_type)
    op = NegativeHamiltonian(pot, scale_kinetic=1.0, laplacian_eps=args.laplacian_eps, n_particles=1)
    op = OperatorWrapper(op, scale=args.operator_scale, shift=args.operator_shift)
    return op, args.operator_scale * gt + args.operator_shift
