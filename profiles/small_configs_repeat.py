"""The three small-batch BASELINE configs, eager and graphed, three times over (how much of a difference is noise: the eager
step is host-bound, so it follows the host CPU's state; the graphed step repeats to 1 %)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
import neural_svd_b200 as N

dev = torch.device("cuda:0")
for rep in range(3):
    for (kind, B, seq) in (("hydrogen", 128, True), ("hydrogen", 512, False), ("oscillator", 512, False)):
        r = bench.small_config(N, kind, B, 16, seq, dev)
        print(rep, kind, B, "seq" if seq else "jnt", "eager ms", round(r["ms_per_step_eager"], 4), "graphed ms",
              round(r["ms_per_step_graphed"], 4), flush=True)
