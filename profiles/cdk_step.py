"""A few CDK loss steps (BASELINE config 5: B = 4096, L = 512, +1 constant mode, joint) for an ncu launch list.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python profiles/cdk_step.py [diag]
Without ncu it prints the event-timed step."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import neural_svd_b200 as N

diag = len(sys.argv) > 1 and sys.argv[1] == "1"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, L = 4096, 512
g = torch.Generator().manual_seed(10)
f = torch.randn(B, L, generator=g).cuda().requires_grad_()
gg = torch.randn(B, L, generator=g).cuda().requires_grad_()
method = N.NestedLoRAForCDK(model=None, neigs=L, step=1, sequential=False, set_first_mode_const=True)
method.diagnostics = diag


def step():
    f.grad = gg.grad = None
    loss = method.compute_loss(f, gg)[0]
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
print(f"diag={diag}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us/step, loss {float(loss):.6f}")
