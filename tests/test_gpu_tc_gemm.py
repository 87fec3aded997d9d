"""The tcgen05 bf16x3 GEMM building block (nsvd_tc_gemm_selftest) against torch fp64 matmul:
single-CTA and CTA-pair (cta_group::2) kernels, K-major and MN-major operands, ragged shapes."""
import ctypes as C

import pytest
import torch

from neural_svd_b200 import _lib

pytestmark = pytest.mark.gpu

# mode: 1 = K-major, 0 = MN-major (single CTA); 3 / 2 = the same on the CTA-pair kernel
CASES = [(128, 256, 64, 1), (200, 264, 72, 1), (512, 768, 2048, 1), (128, 256, 64, 0), (128, 264, 200, 0),
         (256, 512, 1024, 0), (256, 256, 64, 3), (200, 264, 72, 3), (1024, 1024, 2048, 3), (128, 256, 64, 2),
         (128, 264, 200, 2), (128, 2048, 1024, 2)]


@pytest.mark.parametrize("M,N,K,mode", CASES)
def test_bf16x3_gemm_block(M, N, K, mode):
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K + mode)
    A, B = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().T
    kmajor = mode & 1
    Ad = (A if kmajor else A.T.contiguous()).cuda()
    Bd = (B if kmajor else B.T.contiguous()).cuda()
    D = torch.full((M, N), float("nan"), device="cuda")
    work = torch.empty(4 * (M * K + N * K) + 4096, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.nsvd_tc_gemm_selftest(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), M, N, K, mode, mode,
                                         _lib.ptr(work), work.numel(), st), "nsvd_tc_gemm_selftest")
    err = float((D.cpu().double() - ref).norm() / ref.norm())
    assert err < 2e-5, err            # two-term bf16 split: ~2^-18 per operand, 5e-6 .. 1e-5 measured
    hi_only = A.bfloat16().double() @ B.bfloat16().double().T
    assert err < 0.01 * float((hi_only - ref).norm() / ref.norm())   # far better than a single bf16 pass
