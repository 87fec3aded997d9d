"""Where the host time of an EAGER small-batch step goes (BASELINE configs 1-3 are launch/host bound).
    python profiles/eager_host_probe.py [points]
Prints (1) event-timed ms/step eager, (2) wall time per C-ABI call (ctypes call duration = host cost of the call: tensor-map
encoding + launches, the kernels run asynchronously), (3) the cProfile top of 200 steps."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import neural_svd_b200 as N
from neural_svd_b200 import _lib
from conftest import build_problem
from oracle import nsvd_oracle as O

pts = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfg = O.PathConfig.hydrogen()
N.set_engine("f16x3")
method, operator, importance, _ = build_problem(cfg, 0, "cuda")
xs = [(cfg.sampling_scale * torch.randn(pts, 2)).cuda() for _ in range(4)]
it = [0]


def step():
    method.zero_grad(set_to_none=True)
    loss, _ = method.compute_loss_operator(operator, xs[it[0] % 4], importance=importance)
    loss.backward()
    it[0] += 1


for _ in range(20):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(200):
    step()
e1.record()
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"eager B={pts}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us/step (events), {(t1 - t0) / 200 * 1e6:.1f} us/step (wall)")

# wall time per C-ABI call
lib = _lib.load()
acc = {}
names = ["nsvd_fwd_streams", "nsvd_gram_reduce", "nsvd_loss_finalize", "nsvd_loss_dF", "nsvd_mlp_bwd", "nsvd_scratch_bytes",
         "nsvd_gram_partials_bytes"]
orig = {}
for n in names:
    f = getattr(lib, n)
    orig[n] = f

    def wrap(*a, _f=f, _n=n):
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
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(35)
st.sort_stats("tottime").print_stats(25)
