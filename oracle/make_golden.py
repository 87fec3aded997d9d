"""Generate tests/golden/*.npz by running the UNMODIFIED reference (exact Laplacian, CPU).

Run in the build container only (needs /root/reference):
    python oracle/make_golden.py [case ...]

For every case the reference's own `NestedLoRA.compute_loss_operator` + `loss.backward()`
(methods/nestedlora.py:254-267; laplacian_eps=0 -> pde/diff_ops.py:54-93) is evaluated in
fp64 (truth) and fp32 (to record the reference's own fp32 noise).  Stored per case:
  x, seed, config json, loss, f, Tf, and per gradient tensor: Frobenius norm, 4096 sampled
  entries (fixed indices) or the full tensor when small, and the fp32-vs-fp64 relative error.
Weights are NOT stored (19 MB at L=16): they are regenerated from the seed by
`nsvd_oracle.init_params_like_reference`, whose draw order is checked here against the
reference's constructor through the stored parameter checksums.
"""
from __future__ import annotations

import dataclasses
import json
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import nsvd_oracle as O          # noqa: E402
from oracle import ref_bootstrap as RB       # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
FULL_MAX = 70000       # store a gradient tensor in full when it has at most this many entries
NSAMP = 4096

CASES = {
    # BASELINE.json configs[0]: hydrogen, B=128, sequential, L=16
    "hyd_b128_seq_L16": dict(cfg=O.PathConfig.hydrogen(sequential=True), B=128, seed=0),
    # configs[1]: oscillator, B=512, joint, L=16 (learnable exp mask)
    "osc_b512_jnt_L16": dict(cfg=O.PathConfig.oscillator(), B=512, seed=1),
    # configs[2]: hydrogen, B=512, joint, L=16
    "hyd_b512_jnt_L16": dict(cfg=O.PathConfig.hydrogen(), B=512, seed=2),
    # configs[3] shape at reduced B: hydrogen, L=64, joint
    "hyd_b64_jnt_L64": dict(cfg=O.PathConfig.hydrogen(neigs=64), B=64, seed=3),
    # small cases with full gradients: odd batch (torch.chunk -> 49/48), joint step=2, narrow features
    "hyd_small_odd": dict(cfg=O.PathConfig.hydrogen(neigs=4, fourier_mapping_size=64, step=2), B=97, seed=4),
    "osc_small_seq": dict(cfg=O.PathConfig.oscillator(neigs=4, fourier_mapping_size=64, sequential=True,
                                                      hard_mul_const=0.5), B=64, seed=5),
    # the other single-particle 2D problems of pde/problems.py, samplers of main_pde.py:89-118 and the
    # Dirichlet box masks of pde/boundary.py (small shapes, full gradients)
    "well_uniform_boxsqrt": dict(cfg=O.PathConfig(potential="infinite_well", neigs=4, fourier_mapping_size=64,
                                                  fourier_scale=0.5, operator_scale=1.0, sampling_mode="uniform",
                                                  sampling_scale=2.0, lim=2.0, apply_boundary=True,
                                                  boundary_mode="dir_box_sqrt"), B=96, seed=6),
    "cosine_uniform_detff": dict(cfg=O.PathConfig(potential="cosine", neigs=4, fourier_mapping_size=32,
                                                  fourier_scale=1.0, fourier_deterministic=True, operator_scale=1.0,
                                                  operator_shift=2.0, sampling_mode="uniform",
                                                  sampling_scale=math.pi, lim=math.pi, sequential=True),
                                 B=80, seed=7),
    "molion_laplace_boxexp_mask": dict(cfg=O.PathConfig(potential="hydrogen_mol_ion", neigs=4,
                                                        fourier_mapping_size=64, fourier_scale=0.2,
                                                        operator_scale=10.0, sampling_mode="laplacian",
                                                        sampling_scale=3.0, lim=12.0, apply_boundary=True,
                                                        boundary_mode="dir_box_exp", apply_exp_mask=True,
                                                        exp_mask_init_scale=8.0, hydrogen_mol_ion_R=1.5,
                                                        hard_mul_const=2.0), B=101, seed=8),
    # register_eigvals(): the model's output columns are permuted by sort_indices in training mode (nestedlora.py:195-206)
    "hyd_small_sorted": dict(cfg=O.PathConfig.hydrogen(neigs=6, fourier_mapping_size=64, sequential=True), B=64, seed=13,
                             eigvals=[3.0, 9.0, 1.0, 7.0, 8.0, 2.0]),
    # finite-difference Laplacian (the scripts' mode, diff_ops.py:25-52): eps = 0.1 is main_pde's default, 0.01 the
    # value of scripts/exps/pde/*.sh; the fixture stores the reference's fp64 AND fp32 results (FD in fp32 is noisy)
    "hyd_small_fd0p1": dict(cfg=O.PathConfig.hydrogen(neigs=4, fourier_mapping_size=64), B=96, seed=14, eps=0.1),
    "osc_small_fd0p01": dict(cfg=O.PathConfig.oscillator(neigs=4, fourier_mapping_size=64, sequential=True), B=64,
                             seed=15, eps=0.01),
    "hyd_b512_jnt_L16_fd0p01": dict(cfg=O.PathConfig.hydrogen(), B=512, seed=16, eps=0.01),
    # ndim = 3 (problems.py:62-71: hydrogen and the H2+ ion are the potentials the reference runs in 3D -
    # the oscillator / well / cosine branches assert other dimensions); five forward-mode streams, fp32 CUDA-core engine
    "hyd3d_small": dict(cfg=O.PathConfig.hydrogen(ndim=3, neigs=4, fourier_mapping_size=64, sampling_scale=8.0), B=96,
                        seed=17),
    "molion3d_laplace_boxexp": dict(cfg=O.PathConfig(potential="hydrogen_mol_ion", ndim=3, neigs=4,
                                                     fourier_mapping_size=64, fourier_scale=0.2, operator_scale=10.0,
                                                     sampling_mode="laplacian", sampling_scale=3.0, lim=12.0,
                                                     apply_boundary=True, boundary_mode="dir_box_exp",
                                                     apply_exp_mask=True, exp_mask_init_scale=8.0,
                                                     hydrogen_mol_ion_R=1.5), B=80, seed=19),
    "hyd3d_small_fd0p05": dict(cfg=O.PathConfig.hydrogen(ndim=3, neigs=4, fourier_mapping_size=64, sampling_scale=8.0),
                               B=96, seed=20, eps=0.05),
    "osc_no_importance": dict(cfg=O.PathConfig.oscillator(neigs=4, fourier_mapping_size=64, sampling_mode="none"),
                              B=64, seed=9),
}

CDK_CASES = {
    # configs[4]: B=4096, feature dim 512, L=512 (+1 const mode), joint
    "cdk_b4096_L512": dict(B=4096, L=512, seed=10, sequential=False, const=True),
    "cdk_small_seq": dict(B=96, L=24, seed=11, sequential=True, const=True),
    "cdk_small_noconst": dict(B=64, L=16, seed=12, sequential=False, const=False),
}


# methods/spectrum.py:29-102 on a small validation grid (main_pde.py:119-129) that contains the origin; evaluated by
# the reference's own compute_spectrum_evd with every output branch: plain, normalize, normalize+sort+post_align.
SPECTRUM_CASES = {
    "spec_hyd_small": dict(cfg=O.PathConfig.hydrogen(neigs=4, fourier_mapping_size=64, operator_shift=400.0),
                           seed=11, lim=4.0, val_eps=0.25, chunk=300),
    "spec_osc_small": dict(cfg=O.PathConfig.oscillator(neigs=4, fourier_mapping_size=64, sequential=True,
                                                       operator_shift=150.0), seed=12, lim=3.0, val_eps=0.25,
                           chunk=128),
}
SPECTRUM_FLAGS = {"plain": dict(), "norm": dict(normalize=True),
                  "all": dict(normalize=True, sort=True, post_align=True)}


def checksum(a: np.ndarray):
    a = a.astype(np.float64)
    return np.array([a.sum(), (a * a).sum()])


def run_reference(ref, cfg, seed, x32, dtype, eigvals=None, eps=0.0):
    method, operator, importance, gt = RB.build_reference_problem(ref, cfg, seed, eps, dtype)
    if eigvals is not None:
        method.register_eigvals(eigvals)
    x = torch.from_numpy(x32).to(dtype)
    loss, aux = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
    grads = {n: (p.grad.detach().numpy().copy() if p.grad is not None else None)
             for n, p in method.named_parameters()}
    params = {n: p.detach().numpy().copy() for n, p in method.named_parameters()}
    return dict(loss=float(loss.detach()), f=aux["f"].detach().numpy(), Tf=aux["Tf"].detach().numpy(),
                grads=grads, params=params, gt=gt)


def make_case(ref, name, spec):
    cfg, B, seed = spec["cfg"], spec["B"], spec["seed"]
    g = torch.Generator().manual_seed(1000 + seed)
    if cfg.sampling_mode == "uniform":                              # main_pde.py:114-115
        x32 = (cfg.sampling_scale * (2 * torch.rand((B, 1, cfg.ndim), generator=g) - 1)).reshape(B, -1).numpy()
    elif cfg.sampling_mode == "laplacian":                          # main_pde.py:102-106
        u = torch.rand((B, 1, cfg.ndim), generator=g, dtype=torch.float64) - 0.5
        x32 = (-cfg.sampling_scale * u.sign() * torch.log1p(-2 * u.abs())).float().reshape(B, -1).numpy()
    else:                                                           # main_pde.py:92-93
        x32 = (cfg.sampling_scale * torch.randn((B, 1, cfg.ndim), generator=g)).reshape(B, -1).numpy()
    r64 = run_reference(ref, cfg, seed, x32, torch.float64, spec.get("eigvals"), spec.get("eps", 0.0))
    r32 = run_reference(ref, cfg, seed, x32, torch.float32, spec.get("eigvals"), spec.get("eps", 0.0))
    mine = O.init_params_like_reference(cfg, seed)
    out = dict(x=x32, seed=np.int64(seed), config=json.dumps(dataclasses.asdict(cfg)),
               loss64=np.float64(r64["loss"]), loss32=np.float64(r32["loss"]),
               f64=r64["f"], Tf64=r64["Tf"], f32=r32["f"], Tf32=r32["Tf"],
               gt=np.asarray(r64["gt"], np.float64))
    if spec.get("eigvals") is not None:
        out["eigvals"] = np.asarray(spec["eigvals"], np.float64)
    if spec.get("eps"):
        out["laplacian_eps"] = np.float64(spec["eps"])
        # the reference's own fp32-vs-fp64 noise in finite-difference mode
        out["tf_self"] = np.float64(np.linalg.norm(r32["Tf"].astype(np.float64) - r64["Tf"]) / np.linalg.norm(r64["Tf"]))
        out["loss_self"] = np.float64(abs(r32["loss"] / r64["loss"] - 1))
    rs = np.random.RandomState(seed)
    for n in O.param_names(cfg):
        p_ref = r32["params"][n]
        assert np.array_equal(p_ref, mine[n]), f"init draw order mismatch for {n}"
        out[f"pck/{n}"] = checksum(p_ref)
        g64, g32 = r64["grads"][n], r32["grads"][n]
        if g64 is None:
            continue
        nrm = np.linalg.norm(g64)
        out[f"gnorm/{n}"] = np.float64(nrm)
        out[f"gself/{n}"] = np.float64(np.linalg.norm(g32.astype(np.float64) - g64) / max(nrm, 1e-300))
        if g64.size <= FULL_MAX:
            out[f"gfull/{n}"] = g64
        else:
            idx = rs.choice(g64.size, NSAMP, replace=False).astype(np.int64)
            out[f"gidx/{n}"] = idx
            out[f"gval/{n}"] = g64.reshape(-1)[idx]
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: loss64={r64['loss']:.9g} loss32={r32['loss']:.9g} "
          f"|f|={np.linalg.norm(r64['f']):.4g} |Tf|={np.linalg.norm(r64['Tf']):.4g}")


def make_cdk_case(ref, name, spec):
    B, L, seed = spec["B"], spec["L"], spec["seed"]
    g = torch.Generator().manual_seed(seed)
    f32 = torch.randn(B, L, generator=g)
    g32 = torch.randn(B, L, generator=g)
    out = dict(seed=np.int64(seed), B=np.int64(B), L=np.int64(L), sequential=np.bool_(spec["sequential"]),
               const=np.bool_(spec["const"]))
    res = {}
    for tag, dt in (("64", torch.float64), ("32", torch.float32)):
        method = ref.NestedLoRAForCDK(model=None, neigs=L, step=1, sequential=spec["sequential"],
                                      set_first_mode_const=spec["const"])
        method.vector_mask = method.vector_mask.to(dt)
        method.matrix_mask = method.matrix_mask.to(dt)
        f = f32.to(dt).clone().requires_grad_()
        gg = g32.to(dt).clone().requires_grad_()
        loss, lop, lmet, rsj, rsi = method.compute_loss(f, gg)
        loss.backward()
        res[tag] = dict(loss=float(loss.detach()), lop=float(lop.detach()), lmet=float(lmet.detach()), rsj=rsj.detach().numpy(),
                        rsi=rsi.detach().numpy(), gf=f.grad.numpy(), gg=gg.grad.numpy())
    r = res["64"]
    out.update(loss64=np.float64(r["loss"]), lop64=np.float64(r["lop"]), lmet64=np.float64(r["lmet"]),
               loss32=np.float64(res["32"]["loss"]))
    rs = np.random.RandomState(seed)
    if B * L <= FULL_MAX:
        out.update(f=f32.numpy(), g=g32.numpy(), gf64=r["gf"], gg64=r["gg"], rsj64=r["rsj"], rsi64=r["rsi"])
    else:
        out["fck"], out["gck"] = checksum(f32.numpy()), checksum(g32.numpy())
        idx = rs.choice(B * L, NSAMP, replace=False).astype(np.int64)
        out.update(gidx=idx, gfval=r["gf"].reshape(-1)[idx], ggval=r["gg"].reshape(-1)[idx],
                   gfnorm=np.float64(np.linalg.norm(r["gf"])), ggnorm=np.float64(np.linalg.norm(r["gg"])),
                   rsj64=r["rsj"])
        idx2 = rs.choice(r["rsi"].size, NSAMP, replace=False).astype(np.int64)
        out.update(rsi_idx=idx2, rsi_val=r["rsi"][idx2], rsi_norm=np.float64(np.linalg.norm(r["rsi"])))
    out["gself"] = np.float64(np.linalg.norm(res["32"]["gf"] - r["gf"]) / np.linalg.norm(r["gf"]))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(f"{name}: loss64={r['loss']:.9g} loss32={res['32']['loss']:.9g}")


def spectrum_grid(lim, val_eps):
    ax = np.arange(-lim, lim, val_eps)                              # main_pde.py:121-124
    xxs = np.meshgrid(ax, ax)
    return np.array(list(zip(*[xx.flatten() for xx in xxs]))).astype(np.float32)


def make_spectrum_case(ref, name, spec):
    cfg, seed, lim, chunk = spec["cfg"], spec["seed"], spec["lim"], spec["chunk"]
    grid = spectrum_grid(lim, spec["val_eps"])
    assert (np.abs(grid).sum(1) == 0).any(), "the grid must contain the origin (spectrum.py:73)"
    out = dict(seed=np.int64(seed), config=json.dumps(dataclasses.asdict(cfg)), lim=np.float64(lim),
               val_eps=np.float64(spec["val_eps"]), chunk=np.int64(chunk))
    for tag, dt in (("64", torch.float64), ("32", torch.float32)):
        method, operator, importance, _ = RB.build_reference_problem(ref, cfg, seed, 0.0, dt)
        data = torch.from_numpy(grid).to(dt)

        def loader():                                               # main_pde.py:125-127
            for i in range(0, len(data), chunk):
                yield data[i:i + chunk], 0.

        def importance_val(x):                                      # main_pde.py:128-129
            return (1 / (2 * lim) ** cfg.ndim * torch.ones(x.shape[0], 1)).to(dt).view(-1, 1)

        for fl, kw in SPECTRUM_FLAGS.items():
            o = ref.compute_spectrum_evd(method, loader(), operator, importance_train=importance,
                                         importance_val=importance_val, device="cpu", **kw)
            for k, v in o.items():
                out[f"{fl}{tag}/{k}"] = np.asarray(v)
            if tag == "64":
                print(f"{name}[{fl}]: eigvals {np.round(o['eigvals'], 4)} norms {np.round(o['norms'], 5)}",
                      ("aligned " + str(np.round(o['eigvals_aligned'], 4))) if 'eigvals_aligned' in o else "")
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


def main(argv):
    os.makedirs(GOLD, exist_ok=True)
    ref = RB.import_reference("/root/reference")
    torch.set_num_threads(os.cpu_count() or 1)
    want = set(argv) if argv else None
    for name, spec in CASES.items():
        if want is None or name in want:
            make_case(ref, name, spec)
    for name, spec in CDK_CASES.items():
        if want is None or name in want:
            make_cdk_case(ref, name, spec)
    for name, spec in SPECTRUM_CASES.items():
        if want is None or name in want:
            make_spectrum_case(ref, name, spec)


if __name__ == "__main__":
    main(sys.argv[1:])
