"""Step latency of the small-batch BASELINE configs (B = 128 / 512 / 4096) on one GPU, both engines.
Run under gpurun: python profiles/latency_probe.py"""
import json
import subprocess
import sys

for pts in (128, 512, 4096):
    for eng in ("f16x3", "fp32"):
        r = subprocess.run([sys.executable, "bench.py", "--steps", "50", "--warmup", "10", "--points", str(pts),
                            "--engine", eng, "--no-cpu-baseline"], capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"B={pts:5d} {eng:7s} device ms/step {d['ms_per_step']:.4f}  e2e ms/step {1e3 * pts / d['e2e']['value']:.4f}  "
                  f"e2e points/s {d['e2e']['value']:.0f}  launches/step {d['gpu_launches'] / 50:.1f}")
        except Exception as e:
            print(pts, eng, "failed", e, r.stderr[-300:])

# CUDA-graph step (GraphedOperatorStep) vs the eager sequence, B = 128 / 512
import os, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import neural_svd_b200 as N
from conftest import build_problem
from oracle import nsvd_oracle as O
for pts in (128, 512, 4096):
    cfg = O.PathConfig.hydrogen()
    N.set_engine("f16x3")
    method, operator, importance, _ = build_problem(cfg, 0, "cuda")
    step = N.GraphedOperatorStep(method, operator, importance, pts)
    x = (cfg.sampling_scale * torch.randn(pts, 2)).pin_memory()
    for _ in range(10):
        float(step(x))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        float(step(x))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200
    print(f"B={pts:5d} bf16x3 graphed e2e ms/step {dt * 1e3:.4f}  points/s {pts / dt:.0f}")
