"""Step latency of the small-batch BASELINE configs (B = 128 / 512 / 4096) on one GPU, both engines.
Run under gpurun: python profiles/latency_probe.py"""
import json
import subprocess
import sys

for pts in (128, 512, 4096):
    for eng in ("bf16x3", "fp32"):
        r = subprocess.run([sys.executable, "bench.py", "--steps", "50", "--warmup", "10", "--points", str(pts),
                            "--engine", eng, "--no-cpu-baseline"], capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"B={pts:5d} {eng:7s} device ms/step {d['ms_per_step']:.4f}  e2e ms/step {1e3 * pts / d['e2e']['value']:.4f}  "
                  f"e2e points/s {d['e2e']['value']:.0f}  launches/step {d['gpu_launches'] / 50:.1f}")
        except Exception as e:
            print(pts, eng, "failed", e, r.stderr[-300:])
