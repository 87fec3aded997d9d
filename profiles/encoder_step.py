"""A few whole CDK steps with the encoder towers (main_sketchy.py:176-186 shape: two 512 -> 8192 -> 512 towers, B = 4096) for
an ncu launch list; without ncu prints the event-timed step."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import neural_svd_b200 as N

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, L = 4096, 512
g = torch.Generator().manual_seed(10)
net = N.get_sketchy_encoder().cuda()
x, y = torch.randn(B, 512, generator=g).cuda(), torch.randn(B, 512, generator=g).cuda()
m2 = N.NestedLoRAForCDK(model=net, neigs=L, step=1, sequential=False, set_first_mode_const=True).cuda()
m2.diagnostics = False


def step():
    m2.zero_grad(set_to_none=True)
    _, fx, _, fy = m2(x, y)
    loss = m2.compute_loss(fx, fy)[0]
    loss.backward()
    return loss


for _ in range(steps):
    loss = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / steps * 1e3:.1f} us/step, loss {float(loss.detach()):.6f}")
