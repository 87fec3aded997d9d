"""Host-side mirror of the CDK encoder of the reference's Sketchy experiment.

  get_mlp            examples/models/mlp.py:129-164   (Linear / BatchNorm1d / activation stack)
  normalize          examples/models/siam.py:168-186  (l2_ball, l2_sphere, clip, tanh)
  HeteroNetwork      examples/models/siam.py:132-165  (two towers x / y; main_sketchy.py:109-115 builds
                                                       512 -> 8192 -> 512 towers, mu = 16, l2_ball)

The dense layers are `TCLinear` modules (neural_svd_b200/linear.py): forward and backward run on the library's own
tcgen05 GEMM kernels (`nsvd_linear_fwd` / `nsvd_linear_bwd`), with LeakyReLU / ReLU fused into the GEMM epilogue; the
activation module of the reference's Sequential stays as a parameter-free placeholder so that the state-dict keys
(`backbones.x.0.weight`, `backbones.x.2.weight`, ...) match and the reference's checkpoints load unchanged.  The
embeddings feed the fused CDK loss kernels (`NestedLoRAForCDK.compute_loss`, nsvd_cdk_*).
"""
from __future__ import annotations

from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .linear import FusedActivation, TCLinear


def _activation(nonlinearity: str):
    """examples/models/mlp.py:65-88 (the subset that takes no model-specific module)."""
    if nonlinearity == "relu":
        return partial(nn.ReLU, inplace=True)
    if nonlinearity.startswith("lrelu"):
        return partial(nn.LeakyReLU, negative_slope=float(nonlinearity[len("lrelu"):]))
    if nonlinearity.startswith("elu"):
        return partial(nn.ELU, alpha=float(nonlinearity[len("elu"):]))
    table = {"tanh": nn.Tanh, "linear": nn.Identity, "softplus": nn.Softplus}
    if nonlinearity in table:
        return table[nonlinearity]
    raise NotImplementedError(f"activation {nonlinearity!r}")


def _fusable_slope(nonlinearity: str):
    """negative slope when the activation is LeakyReLU / ReLU (fusable into the dense layer's epilogue), else None"""
    if nonlinearity == "relu":
        return 0.0
    if nonlinearity.startswith("lrelu"):
        return float(nonlinearity[len("lrelu"):])
    return None


def get_mlp(sizes, bias=True, nonlinearity="relu", use_bn=True, weight_normalization=False, last_layer_bn=True,
            feature_map=None):
    """Sequential of Linear (+ BatchNorm1d) (+ activation, none after the last layer); `.output_dim` attached."""
    if weight_normalization or feature_map is not None:
        raise NotImplementedError("get_mlp: weight_normalization / feature_map are not used by the CDK encoder")
    act = _activation(nonlinearity)
    sizes = list(sizes)
    if len(sizes) == 1:
        model = nn.BatchNorm1d(sizes[0]) if (use_bn and last_layer_bn) else nn.Identity()
    else:
        layers, n = [], len(sizes) - 1
        fusable = _fusable_slope(nonlinearity)            # ReLU / LeakyReLU run in the GEMM epilogue
        for i in range(n):
            last = i == n - 1
            bn_here = use_bn and (not last or last_layer_bn)
            fuse = fusable is not None and not last and not bn_here
            layers.append(TCLinear(sizes[i], sizes[i + 1], bias=bias, fused_act=("leaky", fusable) if fuse else None))
            if bn_here:
                layers.append(nn.BatchNorm1d(sizes[i + 1]))
            if not last:
                layers.append(FusedActivation(nonlinearity) if fuse else act())
        model = nn.Sequential(*layers)
    model.output_dim = sizes[-1]
    return model


def normalize(z, r_up, regularize_mode):
    """Keep embeddings inside / on the radius-r_up ball (siam.py:168-186)."""
    if not r_up > 0:
        return z
    if regularize_mode == "l2_ball":          # rows with |z| >= r are projected onto the sphere, the others untouched
        inside = (torch.norm(z, p=2, dim=-1) < r_up).to(z.dtype).unsqueeze(1)
        return inside * z + (1 - inside) * r_up * F.normalize(z, p=2, dim=1)
    if regularize_mode == "l2_sphere":
        return r_up * F.normalize(z, p=2, dim=1)
    if regularize_mode == "clip":
        return torch.clip(z, min=-r_up, max=r_up)
    if regularize_mode == "tanh":
        return r_up * torch.tanh(z)
    raise NotImplementedError(regularize_mode)


class HeteroNetwork(nn.Module):
    """Two towers (x: sketches, y: photos): backbone -> projector -> normalize(sqrt(mu))."""

    def __init__(self, backbones, projectors, online_heads=None, mu=1.0, regularize_mode=None):
        super().__init__()
        assert regularize_mode in ["l2_ball", "l2_sphere", "clip", "tanh"]
        self.mu = mu
        self.backbones = nn.ModuleDict({"x": backbones[0], "y": backbones[1]})
        self.projectors = nn.ModuleDict({"x": projectors[0], "y": projectors[1]})
        self.online_heads = nn.ModuleDict({"x": online_heads[0], "y": online_heads[1]}) if online_heads else None
        self.output_dims = {k: (self.backbones[k].output_dim if isinstance(self.projectors[k], nn.Identity)
                                else self.projectors[k].output_dim) for k in self.projectors}
        self.regularize_mode = regularize_mode

    def forward(self, x, y):
        return [*self.forward_single(x, "x"), *self.forward_single(y, "y")]

    def forward_single(self, x, x_or_y, classify=False):
        assert x_or_y in ["x", "y"]
        rep = self.backbones[x_or_y](x)
        emb = normalize(self.projectors[x_or_y](rep), np.sqrt(self.mu), self.regularize_mode)
        if classify:
            return rep, emb, self.online_heads[x_or_y](emb.detach())
        return rep, emb


def get_sketchy_encoder(network_dims="8192,512", mu=16.0, activation="lrelu0.2", use_bn=False,
                        regularize_mode="l2_ball", input_dim=512):
    """The model of main_sketchy.py:107-115 (scripts/exps/sketchy.sh: 512 -> 8192 -> 512 towers, mu = 16)."""
    sizes = [input_dim] + [int(v) for v in network_dims.split(",") if v]
    return HeteroNetwork(
        backbones=[get_mlp(sizes, bias=True, nonlinearity=activation, use_bn=use_bn) for _ in range(2)],
        projectors=[nn.Identity(), nn.Identity()], mu=mu, regularize_mode=regularize_mode)
