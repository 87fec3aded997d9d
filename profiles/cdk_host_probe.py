"""Host time of the eager CDK loss step (B = 4096, L = 512): event/wall time per step, wall time per C-ABI call, cProfile."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import neural_svd_b200 as N
from neural_svd_b200 import _lib

diag = len(sys.argv) > 1 and sys.argv[1] == "1"
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


for _ in range(20):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(200):
    step()
e1.record()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t1 = time.perf_counter() - t0
print(f"diag={diag}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us/step (events), host loop {t_host / 200 * 1e6:.1f} us/step, "
      f"wall {t1 / 200 * 1e6:.1f} us/step")
lib = _lib.load()
acc = {}
names = ["nsvd_cdk_fwd", "nsvd_cdk_finalize", "nsvd_cdk_offdiag", "nsvd_cdk_bwd", "nsvd_cdk_work_bytes"]
orig = {}
for n in names:
    fn = getattr(lib, n)
    orig[n] = fn

    def wrap(*a, _f=fn, _n=n):
        t = time.perf_counter()
        r = _f(*a)
        acc[_n] = acc.get(_n, 0.0) + time.perf_counter() - t
        return r
    setattr(lib, n, wrap)
for _ in range(200):
    step()
torch.cuda.synchronize()
for n in names:
    setattr(lib, n, orig[n])
print("host us per call:", {k: round(v / 200 * 1e6, 1) for k, v in acc.items()})
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
