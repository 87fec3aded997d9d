"""Does a persisting-L2 carve-out (cudaLimitPersistingL2CacheSize) make the evict_last hints of the fused hidden forward's
scratch stick?  Runs profiles/step_classes.py-style timing with the limit at 0 (default) and at the device maximum."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import neural_svd_b200 as N
from neural_svd_b200 import _lib
from conftest import build_problem
from oracle import nsvd_oracle as O

rt = C.CDLL("libcudart.so.12")
cudaLimitPersistingL2CacheSize = 0x06
cudaDevAttrMaxPersistingL2CacheSize = 108
mx = C.c_int()
torch.cuda.init()
torch.zeros(1, device="cuda")
print("attr rc", rt.cudaDeviceGetAttribute(C.byref(mx), cudaDevAttrMaxPersistingL2CacheSize, 0), "max persisting L2 bytes", mx.value)
cfg = O.PathConfig.hydrogen()
N.set_engine("f16x3")
method, operator, importance, _ = build_problem(cfg, 0, "cuda")
x = (cfg.sampling_scale * torch.randn(131072, 2)).cuda()
lib = _lib.load()


def step():
    method.zero_grad(set_to_none=True)
    loss, _ = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()


for limit in (0, 16 << 20, 32 << 20, 48 << 20, 64 << 20, mx.value, 0, mx.value):
    rc = rt.cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, C.c_size_t(limit))
    cur = C.c_size_t()
    rt.cudaDeviceGetLimit(C.byref(cur), cudaLimitPersistingL2CacheSize)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    lib.nsvd_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    cls = {k: round(v[0] / 5, 3) for k, v in _lib.profile_read().items()}
    lib.nsvd_profile_enable(0)
    print(f"limit {cur.value >> 20} MB (rc {rc}): {e0.elapsed_time(e1) / 5:.2f} ms/step  hidden_fwd/l0_fwd {cls['hidden_fwd'] / cls['l0_fwd']:.3f}  "
          f"hidden_bwd/l0_fwd {cls['hidden_bwd'] / cls['l0_fwd']:.3f}", cls, flush=True)
