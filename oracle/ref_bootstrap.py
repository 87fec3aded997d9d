"""Import the UNMODIFIED reference (jongharyu/neural-svd) on CPU with stub modules.

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` (in the build
container, where /root/reference exists) and by `bench.py --impl reference` /
`cpu_baseline` when a copy of the reference travels as `baseline/_ref/`.  The product
package never imports this file.

Recipe = SURVEY.md Appendix A: the reference needs plotting / EMA / CLI packages that
are not installed here and that the hot path never touches; they are replaced by
MagicMock modules *before* importing.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

_STUBS = ["matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1", "torch_ema",
          "termplotlib", "tensorboardX", "configargparse", "uncertainties", "seaborn", "jsonpickle"]


def find_reference() -> str | None:
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.path.join(here, "..", "baseline", "_ref"), "/root/reference"):
        cand = os.path.abspath(cand)
        if os.path.isfile(os.path.join(cand, "methods", "nestedlora.py")):
            return cand
    return None


def import_reference(root: str | None = None):
    """Returns a namespace with the reference's own classes/functions for the hot path."""
    root = root or find_reference()
    if root is None:
        raise RuntimeError("reference not found (looked in baseline/_ref and /root/reference)")
    for n in _STUBS:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = MagicMock()
    import scipy.special as sp
    if not hasattr(sp, "sph_harm"):          # ground_truths.py:2 imports it; unused for 2D
        sp.sph_harm = None
    if root not in sys.path:
        sys.path.insert(0, root)
    from examples.operator.pde import get_wavefunctions            # pde/__init__.py:19
    from examples.operator.pde.problems import get_problem         # problems.py:23
    from methods.nestedlora import (NestedLoRA, NestedLoRAForCDK, NestedLoRALossFunctionEVD,
                                    NestedLoRALossFunctionForCDK)
    from methods.spectrum import compute_spectrum_evd              # spectrum.py:29
    return SimpleNamespace(root=root, get_wavefunctions=get_wavefunctions, get_problem=get_problem,
                           compute_spectrum_evd=compute_spectrum_evd,
                           NestedLoRA=NestedLoRA, NestedLoRAForCDK=NestedLoRAForCDK,
                           NestedLoRALossFunctionEVD=NestedLoRALossFunctionEVD,
                           NestedLoRALossFunctionForCDK=NestedLoRALossFunctionForCDK)


def reference_args(cfg, laplacian_eps: float = 0.0):
    """SimpleNamespace standing in for main_pde.get_args() (main_pde.py:25-87), from a PathConfig."""
    return SimpleNamespace(
        problem="sch", potential_type=cfg.potential, ndim=cfg.ndim, neigs=cfg.neigs, charge=cfg.charge,
        laplacian_eps=laplacian_eps, operator_scale=cfg.operator_scale, operator_shift=cfg.operator_shift,
        lim=cfg.lim, use_fourier_feature=True, fourier_mapping_size=cfg.fourier_mapping_size,
        fourier_scale=cfg.fourier_scale, fourier_deterministic=cfg.fourier_deterministic, fourier_append_raw=False,
        mlp_hidden_dims=",".join(str(h) for h in cfg.hidden), nonlinearity="softplus", parallel=True,
        apply_boundary=cfg.apply_boundary, boundary_mode=cfg.boundary_mode, apply_exp_mask=cfg.apply_exp_mask,
        exp_mask_init_scale=cfg.exp_mask_init_scale, hard_mul_const=cfg.hard_mul_const,
        hydrogen_mol_ion_R=cfg.hydrogen_mol_ion_R, sampling_mode=cfg.sampling_mode,
        sampling_scale=cfg.sampling_scale, use_gaussian_sampling=cfg.sampling_mode == "gaussian", n_particles=1)


def build_reference_problem(ref, cfg, seed: int, laplacian_eps: float = 0.0, dtype=None):
    """(method, operator, importance) exactly as main_pde.main builds them (main_pde.py:176-212)."""
    import torch
    from torch.distributions import MultivariateNormal
    dtype = dtype or torch.float32
    args = reference_args(cfg, laplacian_eps)
    torch.manual_seed(seed)
    operator, gt = ref.get_problem(args, "cpu")
    model = ref.get_wavefunctions(args)
    method = ref.NestedLoRA(model=model, neigs=cfg.neigs, step=cfg.step, sort=False,
                            sequential=cfg.sequential)
    if dtype != torch.float32:
        method = method.to(dtype)
        method.vector_mask = method.vector_mask.to(dtype)
        method.matrix_mask = method.matrix_mask.to(dtype)
    n = cfg.ndim
    if cfg.sampling_mode == "gaussian":
        mvn = MultivariateNormal(loc=torch.zeros(n, dtype=dtype),
                                 covariance_matrix=cfg.sampling_scale ** 2 * torch.eye(n, dtype=dtype))

        def importance(x):                                         # main_pde.py:97-100
            return mvn.log_prob(x.view(x.shape[0], -1)).exp().view(-1, 1)
    elif cfg.sampling_mode == "laplacian":
        from torch.distributions import Laplace
        lap = Laplace(torch.zeros(n, dtype=dtype), cfg.sampling_scale * torch.ones(n, dtype=dtype))

        def importance(x):                                         # main_pde.py:107-112
            return lap.log_prob(x.view(x.shape[0], -1)).sum(-1).exp().view(-1, 1)
    elif cfg.sampling_mode == "uniform":
        def importance(x):                                         # main_pde.py:116-118
            return (1 / (2 * cfg.sampling_scale) ** cfg.ndim * torch.ones(x.shape[0], 1)).to(dtype)
    elif cfg.sampling_mode == "none":
        importance = None
    else:
        raise NotImplementedError(cfg.sampling_mode)

    return method, operator, importance, gt
