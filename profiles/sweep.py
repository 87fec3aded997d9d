"""Parameter sweeps of the tile-order knobs (env vars read once per process). Run under gpurun."""
import json, os, subprocess, sys
def run(env):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "bench.py", "--steps", "6", "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True, env=e)
    d = json.loads(r.stdout.strip().splitlines()[-1])
    k = d["kernels"]
    print(env, round(d["ms_per_step"], 3), "l0", round(k["l0_fwd"]["ms_per_step"], 3), "wgrad", round(k["l0_wgrad"]["ms_per_step"], 3), flush=True)
for mg in ("8", "16", "32", "64"):
    run({"NSVD_L0_MGROUP": mg})
for ks, kg in (("16", "4"), ("32", "2"), ("32", "4"), ("16", "8"), ("8", "8")):
    run({"NSVD_WGRAD_KSLICE": ks, "NSVD_WGRAD_KGROUP": kg})
