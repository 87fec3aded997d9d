"""Phase timeline of one CTA of hidden_fwd12t_kernel (development probe; run under gpurun): builds a side copy of the
library with -DNSVD_TIMELINE (+ extra defines from argv), runs the forward and prints, per ITEM of block 0 (4 items per
tile: L1h0 L1h1 L2h0 L2h1), SM-clock times in us at 1.9 GHz relative to the first item:
 mma_start (accumulator free) | first stage present | MMAs issued | epilogue start (MMAs retired) | epilogue math+stores
 done | producer: adone wait done | producer: loads issued."""
import ctypes as C
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    csrc = os.path.join(ROOT, "neural_svd_b200", "csrc")
    lib_tl = os.path.join(ROOT, "gpurun_out", "libnsvd_tl.so")
    os.makedirs(os.path.dirname(lib_tl), exist_ok=True)
    from neural_svd_b200 import build
    cmd = ["nvcc"] + build.NVCC_FLAGS + ["-DNSVD_TIMELINE"] + sys.argv[1:] + ["-o", lib_tl] + build.SOURCES
    subprocess.run(cmd, cwd=csrc, check=True)
    main_lib = os.path.join(ROOT, "neural_svd_b200", "libnsvd.so")
    shutil.copy(main_lib, main_lib + ".bak")
    shutil.copy(lib_tl, main_lib)
    try:
        import torch
        import neural_svd_b200 as N
        from neural_svd_b200 import _lib
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from conftest import build_problem
        from oracle import nsvd_oracle as O
        cfg = O.PathConfig.hydrogen()
        method, operator, importance, _ = build_problem(cfg, 0, "cuda")
        x = N.sample_gaussian(65536, cfg.sampling_scale, seed=1)
        for _ in range(2):
            loss, _aux = method.compute_loss_operator(operator, x, importance=importance)
        torch.cuda.synchronize()
        lib = _lib.load()
        buf = (C.c_longlong * (64 * 8))()
        lib.nsvd_debug_timeline.restype = C.c_int
        assert lib.nsvd_debug_timeline(buf) == 0
        rows = [[buf[i * 8 + j] for j in range(8)] for i in range(64)]
        t0 = rows[8][0]
        print("item  mma_start first_stage mma_issued  epi_start   epi_done adone_wait loads_issued")
        for i, r in enumerate(rows[8:40]):
            print(f"{i + 8:4d} " + " ".join(f"{(v - t0) / 1900.0:10.2f}" for v in r[:7]))
    finally:
        shutil.move(main_lib + ".bak", main_lib)


if __name__ == "__main__":
    main()
