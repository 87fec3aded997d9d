"""Phase timeline of one CTA of hidden_bwd2_kernel (development probe; run under gpurun): builds a side copy of the library
with -DNSVD_TIMELINE, runs one step and prints per 64-point half tile of block 0 (the LAST launch = layer 1 of the last
micro-batch), us at 1.9 GHz relative to the first row shown:
 loads issued | stage full (MMA may start) | MMAs issued | MMAs retired (epilogue starts) | epilogue done | store drained."""
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
    subprocess.run(["nvcc"] + build.NVCC_FLAGS + ["-DNSVD_TIMELINE"] + sys.argv[1:] + ["-o", lib_tl] + build.SOURCES,
                   cwd=csrc, check=True)
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
            method.zero_grad(set_to_none=True)
            loss, _aux = method.compute_loss_operator(operator, x, importance=importance)
            loss.backward()
        torch.cuda.synchronize()
        lib = _lib.load()
        buf = (C.c_longlong * (64 * 8))()
        lib.nsvd_debug_timeline.restype = C.c_int
        assert lib.nsvd_debug_timeline(buf) == 0
        rows = [[buf[i * 8 + j] for j in range(8)] for i in range(64)]
        t0 = rows[8][0]
        print("tile loads_issued stage_full mma_issued mma_retired  epi_done store_drained")
        for i, r in enumerate(rows[8:36]):
            print(f"{i + 8:4d} " + " ".join(f"{(v - t0) / 1900.0:10.2f}" for v in r[:6]))
    finally:
        shutil.move(main_lib + ".bak", main_lib)


if __name__ == "__main__":
    main()
