"""Phase timeline of one CTA of hidden_fwd12_kernel, the fused hidden forward (development probe; run under gpurun).
Builds a side copy of the library with -DNSVD_TIMELINE, runs one forward and prints, per tile of block 0, the SM-clock
times (relative to the tile's start) of: accumulators free | first stage present | last stage present | MMAs retired |
last TMEM read | epilogue done | all loads issued | weights present.  An ITEM is one layer of one 128-point tile: even items
are layer 1 (operands from DRAM), odd items layer 2 (operands from the L2-resident scratch)."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    csrc = os.path.join(ROOT, "neural_svd_b200", "csrc")
    lib_tl = os.path.join(ROOT, "gpurun_out", "libnsvd_tl.so")
    os.makedirs(os.path.dirname(lib_tl), exist_ok=True)
    from neural_svd_b200 import build
    cmd = ["nvcc"] + build.NVCC_FLAGS + ["-DNSVD_TIMELINE", "-o", lib_tl] + build.SOURCES
    subprocess.run(cmd, cwd=csrc, check=True)
    import shutil
    main_lib = os.path.join(ROOT, "neural_svd_b200", "libnsvd.so")
    shutil.copy(main_lib, main_lib + ".bak")
    shutil.copy(lib_tl, main_lib)          # the package loads the in-tree name
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
        print("item  acc_free  first_stage last_stage  mma_done  last_tmem_rd  epi_done  loads_issued  w_present   (us at 1.9 GHz, "
              "relative to acc_free of item 0)")
        t0 = rows[0][0]
        for i, r in enumerate(rows[:48]):
            print(f"{i:4d} " + " ".join(f"{(v - t0) / 1900.0:10.2f}" for v in r[:8]))
    finally:
        shutil.move(main_lib + ".bak", main_lib)


if __name__ == "__main__":
    main()
