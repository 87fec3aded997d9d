"""Fixed-seed, fixed-step TRAINING run of the UNMODIFIED reference (CPU, exact Laplacian) ->
tests/golden/run_hyd_b128_seq_L16.npz.   Run in the build container only.

Protocol (SURVEY "trajectory sensitivity"): hydrogen, B=128, sequential nesting, L=16, S=200 steps of the
reference loop body (examples/operator/__init__.py:55-73 without AMP): RMSprop(lr=1e-4, alpha=0.999,
eps=1e-10, momentum=0) (examples/utils.py:48-57) + CosineAnnealingLR(T=S); raw (non-EMA) parameters.
Batches: sigma * randn((B,1,2)) from a dedicated CPU generator (seed 4242), identical for every run.
Estimator: on a fixed evaluation batch of 8192 Gaussian points (generator seed 777),
   norms_l = mean_b f_bl^2   (NestedLoRA's eigenvalue estimator, spectrum.py:87)
   rayleigh_l = sum_b f_bl Tf_bl / sum_b f_bl^2   (spectrum.py:86)
evaluated with the reference operator in fp64 on the final parameters.  Stored for the fp32 run (what a user
of the reference gets) and the fp64 run (truth); their difference is the reference's own self-noise.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import nsvd_oracle as O          # noqa: E402
from oracle import ref_bootstrap as RB       # noqa: E402

S, B, SEED = 200, 128, 0


def batches(cfg):
    g = torch.Generator().manual_seed(4242)
    return [(cfg.sampling_scale * torch.randn((B, 1, cfg.ndim), generator=g)).reshape(B, -1) for _ in range(S)]


def eval_batch(cfg):
    g = torch.Generator().manual_seed(777)
    return (cfg.sampling_scale * torch.randn((8192, 1, cfg.ndim), generator=g)).reshape(8192, -1)


def estimators(params, cfg, xe):
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    x = xe.numpy().astype(np.float64)
    u = O.forward_streams(x, p64, cfg)
    Tf, f, _ = O.operator_apply(x, u, p64, cfg)
    return (f * f).mean(0), (f * Tf).sum(0) / (f * f).sum(0)


def run(ref, cfg, dtype):
    method, operator, importance, gt = RB.build_reference_problem(ref, cfg, SEED, 0.0, dtype)
    opt = torch.optim.RMSprop(method.parameters(), lr=1e-4, alpha=0.999, eps=1e-10, weight_decay=0, momentum=0.0)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, S)
    losses = []
    for x in batches(cfg):
        opt.zero_grad()
        loss, _ = method.compute_loss_operator(operator, x.to(dtype), importance=importance)
        loss.backward()
        opt.step()
        sched.step()
        losses.append(float(loss.detach()))
    params = {n: p.detach().numpy().astype(np.float64) for n, p in method.named_parameters()}
    return np.array(losses), params, gt


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref = RB.import_reference("/root/reference")
    cfg = O.PathConfig.hydrogen(sequential=True)
    xe = eval_batch(cfg)
    out = dict(steps=np.int64(S), B=np.int64(B), seed=np.int64(SEED))
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        losses, params, gt = run(ref, cfg, dt)
        norms, ray = estimators(params, cfg, xe)
        out[f"loss{tag}"], out[f"norms{tag}"], out[f"rayleigh{tag}"] = losses, norms, ray
        print(tag, "final loss", losses[-1], "norms", np.round(norms, 4)[:6], "rayleigh", np.round(ray, 3)[:6])
    out["gt"] = np.asarray(gt, np.float64)
    print("self-noise norms", np.abs(out["norms32"] / out["norms64"] - 1).max(), "rayleigh",
          np.abs(out["rayleigh32"] / out["rayleigh64"] - 1).max())
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "run_hyd_b128_seq_L16.npz"), **out)


if __name__ == "__main__":
    main()
