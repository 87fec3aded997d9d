"""Host-side mirror of the reference's eigenfunction model for the hot path.

Same class names, constructor arguments, parameter names/shapes and RNG draw order as the
reference, so `state_dict()` round-trips (`model.base.ws.{i}`, `model.base.bs.{i}`,
`model.base.feature_map._B`, `model.boundary_mask.scales`):
  GaussianFourierFeatureTransform  examples/utils.py:90-143
  ParallelMLP                      examples/models/mlp.py:167-221
  DirichletBoundaryMaskBox         examples/operator/pde/boundary.py:16-37
  ExponentialMask                  examples/operator/pde/boundary.py:39-53
  WaveFunctions / get_wavefunctions examples/operator/pde/__init__.py:8-55

These modules only HOLD parameters.  Evaluation happens in the fused sm_100a kernels
(`neural_svd_b200.fused`); `forward` routes there and raises if CUDA is unavailable.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

HIDDEN = 128


def parse_str(dims_str: str):
    return list(map(int, dims_str.split(","))) if dims_str != "" else []


class GaussianFourierFeatureTransform(nn.Module):
    """phi(x) = [sin(x B), cos(x B)], frozen B (utils.py:102-124): 2 pi scale randn(D, M), or the deterministic
    integer modulation scale [1 I, 2 I, ..., M I] of shape (D, D M)."""

    def __init__(self, input_dim, mapping_size=256, scale=10, deterministic=False, append_raw=False):
        super().__init__()
        if append_raw:
            raise NotImplementedError("fused path supports Fourier features without raw append only")
        self.input_dim = input_dim
        self.deterministic = deterministic
        if deterministic:
            self._B = nn.Parameter(
                scale * torch.cat([i * torch.eye(input_dim) for i in range(1, mapping_size + 1)], dim=0).T,
                requires_grad=False)
            self._mapping_size = input_dim * mapping_size
        else:
            self._B = nn.Parameter(2 * torch.pi * scale * torch.randn((input_dim, mapping_size)).float(),
                                   requires_grad=False)
            self._mapping_size = mapping_size
        self.feature_dim = 2 * self._mapping_size
        self.append_raw = False

    def forward(self, x):
        raise RuntimeError("GaussianFourierFeatureTransform is evaluated inside the fused kernel; call the "
                           "wave-function model or NestedLoRA.compute_loss_operator instead")


class ParallelMLP(nn.Module):
    """L independent MLPs with stacked weights ws[i]: (L, h_i, h_{i-1}), bs[i]: (L, h_i, 1) (mlp.py:181-199)."""

    def __init__(self, input_dim, mlp_hidden_dims, output_dim, num_copies, nonlinearity, bias=False,
                 weight_normalization=False, feature_map=None, debug=False):
        super().__init__()
        if nonlinearity != "softplus":
            raise NotImplementedError("fused path implements softplus only (scripts/exps/pde/*.sh)")
        if weight_normalization:
            raise NotImplementedError("weight_normalization is not supported by the fused path")
        if not bias:
            raise NotImplementedError("fused path expects bias=True (get_mlp_eigfuncs default)")
        if feature_map is None:
            raise NotImplementedError("fused path expects Fourier features (use_fourier_feature=1)")
        if list(mlp_hidden_dims) != [HIDDEN] * 3 or output_dim != 1:
            raise NotImplementedError("fused path is built for mlp_hidden_dims='128,128,128', output_dim=1")
        self.feature_map = feature_map
        ws, bs = nn.ParameterList(), nn.ParameterList()
        prev = feature_map.feature_dim
        for hdim in list(mlp_hidden_dims) + [output_dim]:
            if not debug:
                ws.append(nn.Parameter(math.sqrt(2.0 / prev) * torch.randn(num_copies, hdim, prev)))
                bs.append(nn.Parameter(torch.zeros([num_copies, hdim, 1])))
            else:
                ws.append(nn.Parameter(0.1 * torch.ones([num_copies, hdim, prev])))
                bs.append(nn.Parameter(0.1 * torch.ones([num_copies, hdim, 1])))
            prev = hdim
        self.ws, self.bs = ws, bs
        self.bias = bias
        self.weight_normalization = weight_normalization
        self.num_copies = num_copies

    def forward(self, x):
        raise RuntimeError("ParallelMLP is evaluated inside the fused kernel (WaveFunctions.forward)")


class DirichletBoundaryMaskBox(nn.Module):
    """Zero Dirichlet condition on the box [-lim, lim]^D (boundary.py:16-37); evaluated inside the fused kernel."""

    def __init__(self, lim, mode="dir_box_sqrt"):
        super().__init__()
        assert mode in ["dir_box_sqrt", "dir_box_exp"]
        self.lim = lim
        self.mode = mode


class ExponentialMask(nn.Module):
    """mask_l(x) = exp(-|x| / s_l), s trainable, optionally times a box mask (boundary.py:39-53)."""

    def __init__(self, output_dim, init_scale=1000, boundary_mask=None):
        super().__init__()
        if boundary_mask is not None and not isinstance(boundary_mask, DirichletBoundaryMaskBox) \
                and not _is_unit_mask(boundary_mask):
            raise NotImplementedError("ExponentialMask takes a DirichletBoundaryMaskBox or no inner mask")
        self.output_dim = output_dim
        self.scales = nn.Parameter(init_scale * torch.ones(output_dim))
        self.boundary_mask = boundary_mask if isinstance(boundary_mask, DirichletBoundaryMaskBox) else None


def _is_unit_mask(m) -> bool:
    try:
        return float(m(None)) == 1.0
    except Exception:
        return False


class WaveFunctions(nn.Module):
    """f = hard_mul_const * base(x) * boundary_mask(x) (pde/__init__.py:8-16)."""

    def __init__(self, base, boundary_mask, hard_mul_const=1.0):
        super().__init__()
        self.base = base
        if not isinstance(boundary_mask, (ExponentialMask, DirichletBoundaryMaskBox)) \
                and not _is_unit_mask(boundary_mask):
            raise NotImplementedError("only ExponentialMask, DirichletBoundaryMaskBox or no mask")
        self.boundary_mask = boundary_mask
        self.hard_mul_const = hard_mul_const

    def forward(self, x):
        """(B, L) values of the L eigenfunction networks (no gradient: inference helper)."""
        from . import fused
        return fused.model_values(self, x)


def get_mlp_eigfuncs(input_dim, neigs, mlp_hidden_dims, nonlinearity, bias=True, weight_normalization=False,
                     parallel=False, feature_map=None, debug=False):
    """mlp.py:91-126, parallel=True branch only."""
    if not parallel:
        raise NotImplementedError("fused path implements --parallel 1 only (hydrogen.sh:38)")
    return ParallelMLP(input_dim=input_dim, mlp_hidden_dims=parse_str(mlp_hidden_dims), output_dim=1,
                       num_copies=neigs, bias=bias, nonlinearity=nonlinearity,
                       weight_normalization=weight_normalization, feature_map=feature_map, debug=debug)


def get_wavefunctions(args):
    """pde/__init__.py:19-55 for the configurations the fused path covers."""
    if not args.use_fourier_feature:
        raise NotImplementedError("use_fourier_feature=0")
    n_particles = getattr(args, "n_particles", 1)
    feature_map = GaussianFourierFeatureTransform(
        input_dim=args.ndim * n_particles, mapping_size=args.fourier_mapping_size, scale=args.fourier_scale,
        deterministic=args.fourier_deterministic, append_raw=args.fourier_append_raw)
    base = get_mlp_eigfuncs(input_dim=args.ndim * n_particles, neigs=args.neigs,
                            mlp_hidden_dims=args.mlp_hidden_dims, nonlinearity=args.nonlinearity,
                            parallel=args.parallel, feature_map=feature_map)
    if getattr(args, "apply_boundary", False):
        assert args.boundary_mode in ["dir_box_sqrt", "dir_box_exp"]
        boundary_mask = DirichletBoundaryMaskBox(lim=args.lim, mode=args.boundary_mode)
    else:
        boundary_mask = lambda x: 1.0  # noqa: E731
    if args.apply_exp_mask:
        boundary_mask = ExponentialMask(output_dim=args.neigs, init_scale=args.exp_mask_init_scale,
                                        boundary_mask=boundary_mask)
    return WaveFunctions(base, boundary_mask=boundary_mask, hard_mul_const=args.hard_mul_const)
