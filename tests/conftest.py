import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100) device; run with -m gpu on the B200 box")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    from oracle import nsvd_oracle as O
    d = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = None
    if "config" in d:
        raw = json.loads(str(d["config"]))
        cfg = O.PathConfig(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in raw.items()})
    return d, cfg


def ref_args(cfg, laplacian_eps=0.0):
    """argparse-like namespace for get_problem / get_wavefunctions from an oracle PathConfig."""
    from types import SimpleNamespace
    return SimpleNamespace(
        problem="sch", potential_type=cfg.potential, ndim=cfg.ndim, neigs=cfg.neigs, charge=cfg.charge,
        laplacian_eps=laplacian_eps, operator_scale=cfg.operator_scale, operator_shift=cfg.operator_shift, lim=cfg.lim,
        use_fourier_feature=True, fourier_mapping_size=cfg.fourier_mapping_size, fourier_scale=cfg.fourier_scale,
        fourier_deterministic=cfg.fourier_deterministic, fourier_append_raw=False,
        mlp_hidden_dims=",".join(str(h) for h in cfg.hidden), nonlinearity="softplus", parallel=True,
        apply_boundary=cfg.apply_boundary, boundary_mode=cfg.boundary_mode, apply_exp_mask=cfg.apply_exp_mask,
        exp_mask_init_scale=cfg.exp_mask_init_scale, hard_mul_const=cfg.hard_mul_const,
        hydrogen_mol_ion_R=cfg.hydrogen_mol_ion_R, sampling_mode=cfg.sampling_mode,
        sampling_scale=cfg.sampling_scale)


def build_problem(cfg, seed, device="cpu", laplacian_eps=0.0):
    """(method, operator, importance) from the product's own mirror classes, reference RNG order."""
    import torch
    import neural_svd_b200 as N
    args = ref_args(cfg, laplacian_eps)
    torch.manual_seed(seed)
    operator, gt = N.get_problem(args)
    model = N.get_wavefunctions(args)
    method = N.NestedLoRA(model=model, neigs=cfg.neigs, step=cfg.step, sort=False, sequential=cfg.sequential)
    method = method.to(device)
    importance = {"gaussian": N.GaussianImportance, "laplacian": N.LaplaceImportance, "uniform": N.UniformImportance,
                  "none": lambda *a: None}[cfg.sampling_mode](cfg.sampling_scale, cfg.ndim)
    return method, operator, importance, gt


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def golden_grad_errors(d, names, grads):
    """relative Frobenius error per gradient tensor against the fixture (full or 4096 sampled entries)."""
    out = {}
    for n in names:
        g = np.asarray(grads[n], np.float64)
        if f"gfull/{n}" in d:
            out[n] = rel(g, d[f"gfull/{n}"])
        elif f"gidx/{n}" in d:
            out[n] = rel(g.reshape(-1)[d[f"gidx/{n}"]], d[f"gval/{n}"])
    return out
