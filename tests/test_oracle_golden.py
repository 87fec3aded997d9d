"""The CPU oracle (oracle/nsvd_oracle.py) against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py: exact Laplacian, fp64).  This is what pins the oracle."""
import numpy as np
import pytest

from conftest import golden_grad_errors, load_golden, rel
from oracle import nsvd_oracle as O

PDE_SMALL = ["hyd_small_odd", "osc_small_seq", "hyd_small_sorted",
             # other potentials, samplers, Dirichlet box masks, deterministic features
             "well_uniform_boxsqrt", "cosine_uniform_detff", "molion_laplace_boxexp_mask", "osc_no_importance",
             # ndim = 3: hydrogen and the H2+ ion (the 3D potentials the reference runs, problems.py:62-71)
             "hyd3d_small", "molion3d_laplace_boxexp"]
PDE_FULL = ["hyd_b128_seq_L16", "osc_b512_jnt_L16", "hyd_b512_jnt_L16"]
# finite-difference Laplacian (laplacian_eps > 0, the scripts' mode): fp64 parity only - FD in fp32 is noise-limited
PDE_FD = ["hyd_small_fd0p1", "osc_small_fd0p01", "hyd_b512_jnt_L16_fd0p01", "hyd3d_small_fd0p05"]


def _run(name, dtype):
    d, cfg = load_golden(name)
    p = O.init_params_like_reference(cfg, int(d["seed"]))
    for n in O.param_names(cfg):                       # regenerated weights == reference's constructor draws
        ck = d[f"pck/{n}"]
        a = p[n].astype(np.float64)
        assert np.allclose([a.sum(), (a * a).sum()], ck, rtol=1e-12, atol=1e-12), n
    p = {k: v.astype(dtype) for k, v in p.items()}
    si = None
    if "eigvals" in d:                                  # register_eigvals(): torch.sort(eigvals)[1].flip(0)
        si = np.argsort(d["eigvals"], kind="stable")[::-1].copy()
    eps = float(d["laplacian_eps"]) if "laplacian_eps" in d else 0.0
    r = O.train_step(d["x"].astype(dtype), p, cfg, sort_indices=si, laplacian_eps=eps)
    return d, cfg, r


@pytest.mark.parametrize("name", PDE_SMALL + PDE_FULL + PDE_FD)
def test_oracle_fp64_matches_reference(name):
    d, cfg, r = _run(name, np.float64)
    fd = name in PDE_FD              # second differences divide fp64 round-off by eps^2
    assert abs(r["loss"] - float(d["loss64"])) <= (1e-8 if fd else 1e-11) * abs(float(d["loss64"]))
    assert rel(r["f"], d["f64"]) < 1e-12
    assert rel(r["Tf"], d["Tf64"]) < (1e-8 if fd else 1e-12)
    errs = golden_grad_errors(d, [n for n in O.param_names(cfg) if n in r["grads"]], r["grads"])
    assert len(errs) >= 8
    assert max(errs.values()) < (1e-8 if fd else 1e-11), errs


@pytest.mark.parametrize("name", PDE_SMALL)
def test_oracle_fp32_within_tolerance(name):
    # north-star tolerance: 1e-4 relative on loss and parameter gradients
    d, cfg, r = _run(name, np.float32)
    assert abs(r["loss"] - float(d["loss64"])) <= 1e-4 * abs(float(d["loss64"]))
    errs = golden_grad_errors(d, [n for n in O.param_names(cfg) if n in r["grads"]], r["grads"])
    assert max(errs.values()) < 1e-4, errs


def test_reference_fp32_self_noise_is_recorded():
    d, cfg = load_golden("hyd_b512_jnt_L16")
    for n in O.param_names(cfg):
        if f"gself/{n}" in d:
            assert float(d[f"gself/{n}"]) < 1e-5


@pytest.mark.parametrize("name", ["cdk_small_seq", "cdk_small_noconst"])
def test_cdk_oracle_small(name):
    d, _ = load_golden(name)
    r = O.cdk_forward_backward(d["f"].astype(np.float64), d["g"].astype(np.float64), int(d["L"]),
                               bool(d["sequential"]), 1, bool(d["const"]))
    assert abs(r["loss"] - float(d["loss64"])) < 1e-12 * abs(float(d["loss64"]))
    assert abs(r["loss_operator"] - float(d["lop64"])) < 1e-12 * max(1, abs(float(d["lop64"])))
    assert rel(r["grad_f"], d["gf64"]) < 1e-12 and rel(r["grad_g"], d["gg64"]) < 1e-12
    assert rel(r["rs_joint"], d["rsj64"]) < 1e-12 and rel(r["rs_indep"], d["rsi64"]) < 1e-12


def test_cdk_oracle_full_size():
    import torch
    d, _ = load_golden("cdk_b4096_L512")
    g = torch.Generator().manual_seed(int(d["seed"]))
    f = torch.randn(int(d["B"]), int(d["L"]), generator=g).numpy()
    gg = torch.randn(int(d["B"]), int(d["L"]), generator=g).numpy()
    a = f.astype(np.float64)
    assert np.allclose([a.sum(), (a * a).sum()], d["fck"], rtol=1e-12)
    r = O.cdk_forward_backward(f.astype(np.float64), gg.astype(np.float64), int(d["L"]), False, 1, True,
                               diagnostics=False)
    assert abs(r["loss"] - float(d["loss64"])) < 1e-11 * abs(float(d["loss64"]))
    assert rel(r["grad_f"].reshape(-1)[d["gidx"]], d["gfval"]) < 1e-11
    assert rel(r["grad_g"].reshape(-1)[d["gidx"]], d["ggval"]) < 1e-11


def test_analytic_spectra_known_answers():
    # hydrogen x100 -> [100, 11.11 x3, 4 x5, 2.0408 x7]; oscillator shifted 16 - [2,4,4,6,6,6,...]
    d, _ = load_golden("hyd_b128_seq_L16")
    gt = d["gt"]
    assert np.allclose(gt[:1], 100.0) and np.allclose(gt[1:4], 100 / 9) and np.allclose(gt[4:9], 4.0)
    assert np.allclose(gt[9:16], 100 / 49)
    d, _ = load_golden("osc_b512_jnt_L16")
    assert np.allclose(d["gt"][:6], 16 - np.array([2, 4, 4, 6, 6, 6]))


SPEC_KEYS = ["cov", "quad", "eigvals", "norms", "eigfuncs"]


def _spec_grid(d):
    ax = np.arange(-float(d["lim"]), float(d["lim"]), float(d["val_eps"]))
    xxs = np.meshgrid(ax, ax)
    return np.array(list(zip(*[xx.flatten() for xx in xxs]))).astype(np.float32)


def spec_close(out, d, tag, tol, aligned_tol=None):
    """compare a compute_spectrum_evd output dict with the fixture group `tag` (e.g. 'all64')."""
    for k in SPEC_KEYS:
        assert rel(out[k], d[f"{tag}/{k}"]) < tol, (tag, k, rel(out[k], d[f"{tag}/{k}"]))
    if f"{tag}/eigvals_aligned" in d:
        at = aligned_tol or tol
        assert rel(out["eigvals_aligned"], d[f"{tag}/eigvals_aligned"]) < at
        assert rel(out["cov_aligned"], d[f"{tag}/cov_aligned"]) < at
        a, b = np.asarray(out["eigfuncs_aligned"], np.float64), d[f"{tag}/eigfuncs_aligned"].astype(np.float64)
        sgn = np.sign((a * b).sum(0))                    # eigenvectors are defined up to sign
        assert rel(a * sgn[None, :], b) < at


@pytest.mark.parametrize("name", ["spec_hyd_small", "spec_osc_small"])
@pytest.mark.parametrize("flags", ["plain", "norm", "all"])
def test_oracle_spectrum_matches_reference(name, flags):
    # methods/spectrum.py:29-102 run by the unmodified reference (oracle/make_golden.py: SPECTRUM_CASES), every output
    # branch: nan_to_num + origin-row zeroing (the grid contains the origin), sqrt_ws re-weighting, normalize, sort,
    # post_align
    d, cfg = load_golden(name)
    p = {k: v.astype(np.float64) for k, v in O.init_params_like_reference(cfg, int(d["seed"])).items()}
    kw = dict(plain={}, norm=dict(normalize=True), all=dict(normalize=True, sort=True, post_align=True))[flags]
    with np.errstate(all="ignore"):
        out = O.spectrum_evd(_spec_grid(d).astype(np.float64), p, cfg, float(d["lim"]), chunk=int(d["chunk"]), **kw)
    spec_close(out, d, flags + "64", 1e-10, aligned_tol=1e-8)
