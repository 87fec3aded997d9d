"""Device probe: do tcgen05 kind::f16 MMAs accept different A / B formats (fp16 x bf16), honour fp16 subnormals,
and what accuracy do the two-plane formats of nsvd_tc.cuh give?  (run under gpurun; not a pytest file)
A in [-1, 1] (features), B ~ 0.03 N(0,1) scaled by 2^8 for the fp16 planes (weights), K = 2048."""
import ctypes as C
import subprocess
import sys

FMT = {"BB": 0, "BH": 1, "HH": 2}


def run_case(fa, fb, four, base, K=2048):
    import torch
    sys.path.insert(0, ".")
    from neural_svd_b200 import _lib
    lib = _lib.load()
    M, N = 256, 512
    g = torch.Generator().manual_seed(5)
    A = torch.rand(M, K, generator=g) * 2 - 1
    B = 0.03 * torch.randn(N, K, generator=g)
    if fb == "HH":
        B = B * 256.0
    if fa == "BH-small":            # gradients: tiny values pre-scaled so the lo plane stays in the fp16 range
        A = A * 1e-6 * 2.0 ** 17
        fa = "BH"
    ref = A.double() @ B.double().T
    kmajor = base & 1
    Ad = (A if kmajor else A.T.contiguous()).cuda()
    Bd = (B if kmajor else B.T.contiguous()).cuda()
    D = torch.full((M, N), float("nan"), device="cuda")
    work = torch.empty(4 * (M * K + N * K) + 4096, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mode = base | (FMT[fa] << 4) | (FMT[fb] << 6) | (four << 8)
    rc = lib.nsvd_tc_gemm_selftest(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), M, N, K, mode, mode, _lib.ptr(work),
                                   work.numel(), st)
    if rc:
        print(f"A={fa} B={fb} four={four} base={base}: rc={rc} {lib.nsvd_last_error()}")
        return
    torch.cuda.synchronize()
    Dh = D.cpu().double()
    print(f"A={fa:8s} B={fb} terms={3 + four} base={base} K={K}: rel_err={float((Dh - ref).norm() / ref.norm()):.3e} "
          f"nans={int(torch.isnan(Dh).sum())}")


if __name__ == "__main__":
    if len(sys.argv) >= 5:
        run_case(sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), *(int(v) for v in sys.argv[5:]))
    else:
        cases = [(a, b, f, base) for base in (1,) for (a, b, f) in
                 [("BB", "BB", 0), ("BB", "BB", 1), ("HH", "HH", 0), ("BH", "HH", 0), ("BH", "HH", 1),
                  ("BH", "BH", 0), ("BH", "BH", 1), ("BH-small", "HH", 1), ("BH-small", "BH", 1)]]
        cases += [("HH", "HH", 0, 1, 128), ("HH", "HH", 0, 1, 512), ("HH", "HH", 0, 1, 8192), ("BB", "BB", 1, 1, 128)]
        for c in cases:
            r = subprocess.run([sys.executable, __file__] + [str(v) for v in c], capture_output=True, text=True,
                               timeout=120)
            out = [l for l in (r.stdout + r.stderr).strip().splitlines() if "arn" not in l]
            print("\n".join(out[-6:]) if out else f"case {c}: no output rc={r.returncode}")
