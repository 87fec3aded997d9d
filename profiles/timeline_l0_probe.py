"""Phase timeline of one CTA pair of the layer-0 forward GEMM (big2s_gemm_kernel; development probe, run under gpurun).
Builds a side copy of the library with -DNSVD_TIMELINE, runs one forward and prints per tile of block 0 (SM clocks):
MMA issuer: tile duration | cycles waiting for a free TMEM buffer | cycles waiting for operands;
epilogue warp 2: first sub-chain ready -> last drain done | time inside the drains | fused math + stores."""
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
    subprocess.run(["nvcc"] + build.NVCC_FLAGS + ["-DNSVD_TIMELINE"] + sys.argv[1:] + ["-o", lib_tl] + build.SOURCES, cwd=csrc, check=True)
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
            N.fused.apply_operator(method, operator, x, importance)
        torch.cuda.synchronize()
        lib = _lib.load()
        buf = (C.c_longlong * (64 * 8))()
        lib.nsvd_debug_timeline_l0.restype = C.c_int
        assert lib.nsvd_debug_timeline_l0(buf) == 0
        rows = [[buf[i * 8 + j] for j in range(8)] for i in range(64)]
        print(f"sub_chunks={os.environ.get('NSVD_L0_SUBCHUNKS', 'default')} defines={sys.argv[1:]}")
        print("tile | mma: start  dur  wait_tmem wait_operands | epi: first_ready(rel mma start) drains_span in_drain math+stores")
        t0 = rows[0][0]
        for i, r in enumerate(rows[:12]):
            print(f"{i:4d} | {r[0] - t0:9d} {r[3] - r[0]:7d} {r[1]:7d} {r[2]:7d} | {r[4] - r[0]:9d} {r[5] - r[4]:9d} {r[7]:7d} {r[6] - r[5]:7d}")
    finally:
        shutil.move(main_lib + ".bak", main_lib)


if __name__ == "__main__":
    main()
