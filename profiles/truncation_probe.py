"""Device probe: is the truncation error of the fp32 accumulator in TMEM a coherent shrink toward zero?
D = A.B^T with fp16 hi/lo planes (operand error < 1e-7) at several chain lengths K; fits the scalar alpha that minimises
|alpha D_device - D_exact| and prints the error before / after.  Run under gpurun."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from neural_svd_b200 import _lib


def run(K, dist):
    lib = _lib.load()
    M, N = 256, 512
    g = torch.Generator().manual_seed(5 + K)
    if dist == "features":
        A = torch.rand(M, K, generator=g) * 2 - 1
        B = 0.03 * torch.randn(N, K, generator=g) * 256.0
    else:       # softplus-like activations (positive) x zero-mean weights
        A = torch.nn.functional.softplus(torch.randn(M, K, generator=g)) * 64.0
        B = 0.125 * torch.randn(N, K, generator=g) * 64.0
    ref = A.double() @ B.double().T
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.full((M, N), float("nan"), device="cuda")
    work = torch.empty(4 * (M * K + N * K) + 4096, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mode = 1 | (2 << 4) | (2 << 6)          # K-major, single CTA, fp16+fp16 planes for both operands
    rc = lib.nsvd_tc_gemm_selftest(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), M, N, K, mode, mode, _lib.ptr(work),
                                   work.numel(), st)
    assert rc == 0, lib.nsvd_last_error()
    torch.cuda.synchronize()
    Dh = D.cpu().double()
    err0 = float((Dh - ref).norm() / ref.norm())
    alpha = float((Dh * ref).sum() / (Dh * Dh).sum())
    err1 = float((alpha * Dh - ref).norm() / ref.norm())
    shrink = float(((Dh.abs() < ref.abs()).double().mean()))
    print(f"{dist:9s} K={K:5d} ({3 * K // 16:4d} MMAs): rel_err {err0:.2e}  alpha-1 {alpha - 1:+.3e}  after scaling {err1:.2e}  "
          f"|D|<|ref| in {100 * shrink:.0f}% of entries  (alpha-1)/MMA {(alpha - 1) / (3 * K / 16):.2e}")


if __name__ == "__main__":
    for dist in ("features", "softplus"):
        for K in (64, 128, 256, 512, 1024, 2048):
            run(K, dist)
