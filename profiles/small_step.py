"""A few eager loss+grad steps at a small batch (for an ncu launch list of the launch-bound configs).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python profiles/small_step.py 512"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import neural_svd_b200 as N
from conftest import build_problem
from oracle import nsvd_oracle as O

pts = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fd = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
cfg = O.PathConfig.hydrogen()
N.set_engine("f16x3")
method, operator, importance, _ = build_problem(cfg, 0, "cuda", laplacian_eps=fd)
x = (cfg.sampling_scale * torch.randn(pts, 2)).cuda()
for _ in range(steps):
    method.zero_grad(set_to_none=True)
    loss, _ = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
