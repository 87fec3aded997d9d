"""Device probe: fixed-seed training run of tests/test_gpu_training_run.py under different accumulation-chain lengths
(NSVD_L0_SUBCHUNKS) - prints the worst-mode errors of the two eigenvalue estimators.  Run under gpurun."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch
    import neural_svd_b200 as N
    from conftest import build_problem, load_golden
    from oracle import nsvd_oracle as O
    d, _ = load_golden("run_hyd_b128_seq_L16")
    S, B, seed = int(d["steps"]), int(d["B"]), int(d["seed"])
    cfg = O.PathConfig.hydrogen(sequential=True)
    N.set_engine("f16x3")
    method, operator, importance, _ = build_problem(cfg, seed, "cuda")
    opt = torch.optim.RMSprop(method.parameters(), lr=1e-4, alpha=0.999, eps=1e-10, weight_decay=0, momentum=0.0)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, S)
    g = torch.Generator().manual_seed(4242)
    for _ in range(S):
        x = (cfg.sampling_scale * torch.randn((B, 1, cfg.ndim), generator=g)).reshape(B, -1)
        opt.zero_grad()
        loss, _ = method.compute_loss_operator(operator, x.cuda(), importance=importance)
        loss.backward()
        opt.step()
        sched.step()
    ge = torch.Generator().manual_seed(777)
    xe = (cfg.sampling_scale * torch.randn((8192, 1, cfg.ndim), generator=ge)).reshape(8192, -1)
    params = {n: p.detach().cpu().numpy().astype(np.float64) for n, p in method.named_parameters()}
    x64 = xe.numpy().astype(np.float64)
    u = O.forward_streams(x64, params, cfg)
    Tf, f, _ = O.operator_apply(x64, u, params, cfg)
    norms, ray = (f * f).mean(0), (f * Tf).sum(0) / (f * f).sum(0)
    e_n, e_r = np.abs(norms / d["norms64"] - 1), np.abs(ray / d["rayleigh64"] - 1)
    print(f"sub={os.environ.get('NSVD_L0_SUBCHUNKS')} norms max {e_n.max():.2e} med {np.median(e_n):.2e} | rayleigh max "
          f"{e_r.max():.2e} med {np.median(e_r):.2e} | loss last {abs(float(loss) / d['loss64'][-1] - 1):.2e}", flush=True)


def swap_lib(defines):
    """build a side copy of the library with extra nvcc defines and put it at the in-tree name (returns the backup path)"""
    import shutil
    sys.path.insert(0, ROOT)
    from neural_svd_b200 import build
    csrc = os.path.join(ROOT, "neural_svd_b200", "csrc")
    side = os.path.join(ROOT, "gpurun_out", "libnsvd_variant.so")
    os.makedirs(os.path.dirname(side), exist_ok=True)
    subprocess.run(["nvcc"] + build.NVCC_FLAGS + defines + ["-o", side] + build.SOURCES, cwd=csrc, check=True)
    main_lib = os.path.join(ROOT, "neural_svd_b200", "libnsvd.so")
    shutil.copy(main_lib, main_lib + ".bak")
    shutil.copy(side, main_lib)
    return main_lib


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        # arguments: "SUB" or "SUB,-DFLAG[,-DFLAG2]" (a side build of the library with those defines)
        for spec in sys.argv[1:] or ["4", "8"]:
            sub, *defs = spec.split(",")
            main_lib = swap_lib(defs) if defs else None
            try:
                env = dict(os.environ, NSVD_L0_SUBCHUNKS=sub)
                r = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True)
                out = [l for l in r.stdout.splitlines() if l.startswith("sub=")]
                print(spec, "|", "\n".join(out) if out else r.stderr[-400:], flush=True)
            finally:
                if main_lib:
                    import shutil
                    shutil.move(main_lib + ".bak", main_lib)
