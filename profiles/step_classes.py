"""Step time and per-kernel-class times of the headline workload (CUDA events inside the library), a few repetitions.
    python profiles/step_classes.py [points] [neigs] [reps]      # NSVD_* environment switches apply"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import neural_svd_b200 as N
from neural_svd_b200 import _lib
from conftest import build_problem
from oracle import nsvd_oracle as O

pts = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
neigs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = O.PathConfig.hydrogen(neigs=neigs)
N.set_engine("f16x3")
method, operator, importance, _ = build_problem(cfg, 0, "cuda")
x = (cfg.sampling_scale * torch.randn(pts, 2)).cuda()
lib = _lib.load()


def step():
    method.zero_grad(set_to_none=True)
    loss, _ = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
for _ in range(reps):
    lib.nsvd_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    cls = {k: round(v[0] / 5, 3) for k, v in _lib.profile_read().items()}
    lib.nsvd_profile_enable(0)
    print(f"{e0.elapsed_time(e1) / 5:.2f} ms/step", cls, flush=True)
