"""End-to-end known-answer test: train the 2D hydrogen problem (scripts/exps/pde/hydrogen.sh hyper-parameters,
L=16, sequential nesting) from random weights with the whole B200 path — device sampler, fused forward/loss/
backward kernels (bf16x3 engine), fused RMSprop/EMA step — and compare the learned eigenvalue estimates with the
ANALYTIC spectrum the reference ships (schrodinger/ground_truths.py:120-132, operator scale 100):
    100, 11.11 x3, 4 x5, 2.04 x7.
1500 steps of 32768 points (~10 s on a B200) resolve the first three shells to a few percent."""
import numpy as np
import pytest
import torch

import neural_svd_b200 as N
from conftest import build_problem
from oracle import nsvd_oracle as O

pytestmark = pytest.mark.gpu


def test_hydrogen_spectrum_emerges_from_training():
    cfg = O.PathConfig.hydrogen(sequential=True)
    steps, B = 1500, 32768
    N.set_engine("f16x3")
    method, operator, importance, gt = build_problem(cfg, 0, "cuda")
    assert np.allclose(gt[:9], [100] + [100 / 9] * 3 + [4] * 5)
    opt = N.FusedRMSpropEMA(method.parameters(), lr=1e-4, alpha=0.999, eps=1e-10, ema_decay=0.995, num_iters=steps)
    first = None
    for it in range(steps):
        x = N.sample_gaussian(B, cfg.sampling_scale, seed=7, offset=it * B)
        opt.zero_grad()
        loss, _ = method.compute_loss_operator(operator, x, importance=importance)
        loss.backward()
        opt.step()
        if it == 0:
            first = float(loss.detach())
    assert np.isfinite(float(loss.detach())) and float(loss.detach()) < first
    # eigenvalue estimator of NestedLoRA (methods/spectrum.py:87): norms_l = E_w[f_l^2] on a fresh large batch
    xe = N.sample_gaussian(1 << 18, cfg.sampling_scale, seed=99)
    Tf, f = operator(method, xe, importance=importance)
    norms = (f.double() ** 2).mean(0).cpu().numpy()
    rayleigh = ((f.double() * Tf.double()).sum(0) / (f.double() ** 2).sum(0)).cpu().numpy()
    print("ground truth:", np.round(gt, 2))
    print("norms       :", np.round(norms, 2))
    print("rayleigh    :", np.round(rayleigh, 2))
    assert abs(norms[0] / 100.0 - 1) < 0.25                                   # 1s state: heavy-tailed estimator
    assert np.all(np.abs(np.sort(norms[1:4])[::-1] / (100 / 9) - 1) < 0.10)    # n = 1 shell (3-fold)
    assert np.all(np.abs(np.sort(norms[4:9])[::-1] / 4.0 - 1) < 0.10)          # n = 2 shell (5-fold)
    assert np.all(norms[9:] < 3.0) and np.all(norms[9:] > 0.2)                 # n = 3 shell still converging


def test_oscillator_spectrum_emerges_from_training():
    # scripts/exps/pde/oscillator.sh: H = -Lap + r^2, shifted operator 16 - H, learnable ExponentialMask (its scale
    # gradients are exercised by the training), sequential nesting.  Analytic: 16 - (2n + 2) -> 14, 12 x2, 10 x3, ...
    cfg = O.PathConfig.oscillator(sequential=True)
    steps, B = 2500, 32768
    N.set_engine("f16x3")
    method, operator, importance, gt = build_problem(cfg, 0, "cuda")
    assert np.allclose(gt[:6], [14, 12, 12, 10, 10, 10])
    s0 = method.model.boundary_mask.scales.detach().clone()
    opt = N.FusedRMSpropEMA(method.parameters(), lr=1e-4, alpha=0.999, eps=1e-10, ema_decay=0.995, num_iters=steps)
    for it in range(steps):
        x = N.sample_gaussian(B, cfg.sampling_scale, seed=3, offset=it * B)
        opt.zero_grad()
        loss, _ = method.compute_loss_operator(operator, x, importance=importance)
        loss.backward()
        opt.step()
    assert np.isfinite(float(loss.detach()))
    assert not torch.equal(method.model.boundary_mask.scales.detach(), s0)      # the mask scales were trained
    xe = N.sample_gaussian(1 << 18, cfg.sampling_scale, seed=99)
    Tf, f = operator(method, xe, importance=importance)
    rayleigh = ((f.double() * Tf.double()).sum(0) / (f.double() ** 2).sum(0)).cpu().numpy()
    print("ground truth:", np.round(gt, 2))
    print("rayleigh    :", np.round(rayleigh, 2))
    # measured: 13.95 11.89 11.91 9.93 9.91 (0.4-0.9 % off); later modes are still sorting themselves out
    assert np.all(np.abs(rayleigh[:5] / gt[:5] - 1) < 0.03)
    assert np.all(rayleigh[5:10] > 5.0) and np.all(rayleigh[5:10] < 10.5)


def test_infinite_well_spectrum_emerges_from_training():
    # SURVEY §8 f-4 families end to end: infinite well on [-1, 1]^2 (problems.py:30-33), uniform sampler and importance
    # (main_pde.py:113-118), Dirichlet box mask 'dir_box_sqrt' (boundary.py:29-31), shifted operator 40 - H.
    # Analytic (ground_truths.py:40-59, L = 2): 40 - (nx^2 + ny^2) pi^2 / 4 -> 35.07, 27.66 x2, 20.26, 15.33 x2, ...
    cfg = O.PathConfig(potential="infinite_well", neigs=8, fourier_mapping_size=256, fourier_scale=0.5,
                       operator_scale=1.0, operator_shift=40.0, sampling_mode="uniform", sampling_scale=1.0, lim=1.0,
                       apply_boundary=True, boundary_mode="dir_box_sqrt", sequential=True)
    steps, B = 2000, 16384
    N.set_engine("f16x3")
    method, operator, importance, gt = build_problem(cfg, 0, "cuda")
    assert np.allclose(gt[:4], 40 - np.array([2, 5, 5, 8]) * np.pi ** 2 / 4)
    opt = N.FusedRMSpropEMA(method.parameters(), lr=1e-3, alpha=0.999, eps=1e-10, ema_decay=0.995, num_iters=steps)
    g = torch.Generator(device="cuda").manual_seed(11)
    for it in range(steps):
        x = cfg.sampling_scale * (2 * torch.rand(B, 2, device="cuda", generator=g) - 1)
        opt.zero_grad()
        loss, _ = method.compute_loss_operator(operator, x, importance=importance)
        loss.backward()
        opt.step()
    assert np.isfinite(float(loss.detach()))
    xe = cfg.sampling_scale * (2 * torch.rand(1 << 18, 2, device="cuda", generator=g) - 1)
    Tf, f = operator(method, xe, importance=importance)
    rayleigh = ((f.double() * Tf.double()).sum(0) / (f.double() ** 2).sum(0)).cpu().numpy()
    print("ground truth:", np.round(gt, 2))
    print("rayleigh    :", np.round(rayleigh, 2))
    # measured: 35.07 27.66 27.66 20.26 15.33 15.33 7.92 7.76: the first seven modes within 1e-3, the last one (second
    # member of a degenerate pair, the slowest to settle) within 2-3 % after 2000 steps
    assert np.all(np.abs(rayleigh[:7] / gt[:7] - 1) < 0.005)
    assert abs(rayleigh[7] / gt[7] - 1) < 0.05
