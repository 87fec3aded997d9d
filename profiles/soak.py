"""Stability soak: N full training steps (device sampler + fused step + fused RMSprop/EMA) at a large batch.
Prints the loss trajectory, checks finiteness, and evaluates the spectrum at the end. Run under gpurun."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import neural_svd_b200 as N
from conftest import build_problem
from oracle import nsvd_oracle as O

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
which = sys.argv[3] if len(sys.argv) > 3 else 'hydrogen'
cfg = O.PathConfig.hydrogen(sequential=True) if which == 'hydrogen' else O.PathConfig.oscillator(sequential=(which == 'osc_seq'))
N.set_engine("f16x3")
method, operator, importance, gt = build_problem(cfg, 0, "cuda")
opt = N.FusedRMSpropEMA(method.parameters(), lr=1e-4, alpha=0.999, eps=1e-10, ema_decay=0.995, num_iters=steps)
losses = []
torch.cuda.synchronize(); t0 = time.perf_counter()
for it in range(steps):
    x = N.sample_gaussian(B, cfg.sampling_scale, seed=7, offset=it * B)
    opt.zero_grad()
    loss, aux = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
    opt.step()
    if it % 100 == 0 or it == steps - 1:
        losses.append((it, float(loss.detach())))
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{steps} steps of {B} points in {dt:.1f} s = {steps * B / dt / 1e6:.2f} M points/s (sampler + step + optimizer)")
print("loss:", " ".join(f"{i}:{l:.1f}" for i, l in losses))
assert all(np.isfinite(l) for _, l in losses)
f = aux["f"]
norms = (f * f).mean(0).cpu().numpy()
ray = ((f * aux["Tf"]).sum(0) / (f * f).sum(0)).cpu().numpy()
print("ground truth     :", np.round(gt, 2))
print("norms (last batch):", np.round(norms, 2))
print("rayleigh          :", np.round(ray, 2))
