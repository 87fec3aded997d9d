"""Device probe of the tcgen05 bf16x3 GEMM building block (run under gpurun; not a pytest file).
Each case runs in its own process so a trapped kernel cannot poison the next one."""
import ctypes as C
import subprocess
import sys

CASES = [  # M, N, K, mode: 1 = K-major, 0 = MN-major, 3 / 2 = the same on the CTA-pair (cta_group::2) kernel
    (256, 256, 64, 3), (256, 512, 2048, 3), (200, 264, 72, 3), (1024, 1024, 512, 3), (2048, 4096, 2048, 3),
    (128, 256, 64, 2), (128, 2048, 1024, 2), (128, 264, 200, 2), (128, 512, 4096, 2),
    (128, 256, 64, 1), (128, 256, 256, 1), (256, 512, 2048, 1), (200, 264, 72, 1), (1024, 1024, 512, 1),
    (128, 256, 64, 0), (128, 256, 256, 0), (128, 2048, 1024, 0), (128, 264, 200, 0), (256, 512, 4096, 0),
]


def run_case(M, N, K, kmajor):
    import torch
    sys.path.insert(0, ".")
    from neural_svd_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = (A.double() @ B.double().T)
    mode = kmajor
    kmajor = mode & 1
    Ad = (A if kmajor else A.T.contiguous()).cuda()
    Bd = (B if kmajor else B.T.contiguous()).cuda()
    D = torch.full((M, N), float("nan"), device="cuda")
    work = torch.empty(4 * (M * K + N * K) + 4096, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.nsvd_tc_gemm_selftest(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(D), M, N, K, mode, mode, _lib.ptr(work),
                                   work.numel(), st)
    if rc:
        print(f"case {M}x{N}x{K} mode={mode}: rc={rc} {lib.nsvd_last_error()}")
        return
    torch.cuda.synchronize()
    Dh = D.cpu().double()
    err = (Dh - ref).norm() / ref.norm()
    nan = int(torch.isnan(Dh).sum())
    print(f"case {M}x{N}x{K} mode={mode}: rel_err={err:.3e} nans={nan}")
    if not (err < 1e-4):
        Dh = torch.nan_to_num(Dh)
        e = (Dh - ref).abs()
        bm, bn = max(M // 8, 1), max(N // 8, 1)
        blk = e[:bm * 8, :bn * 8].reshape(8, bm, 8, bn).mean((1, 3))
        print("  mean |err| per 8x8 block grid (rows x cols), ref rms=%.2f" % ref.pow(2).mean().sqrt())
        for r in blk:
            print("   " + " ".join(f"{v:8.2f}" for v in r))
        # is D a copy of a single-pass (hi*hi only) or of a permuted truth?
        hi = A.bfloat16().double() @ B.bfloat16().double().T
        print("  rel err vs hi*hi only: %.3e" % ((Dh - hi).norm() / hi.norm()))
        print("  D[0,:8] =", [round(float(v), 3) for v in Dh[0, :8]])
        print("  ref[0,:8]=", [round(float(v), 3) for v in ref[0, :8]])


if __name__ == "__main__":
    if len(sys.argv) == 5:
        run_case(*map(int, sys.argv[1:]))
    else:
        for c in CASES:
            r = subprocess.run([sys.executable, __file__] + [str(v) for v in c], capture_output=True, text=True,
                               timeout=120)
            out = (r.stdout + r.stderr).strip().splitlines()
            keep = [l for l in out if "Warning" not in l and "warn" not in l]
            print("\n".join(keep[-25:]) if keep else f"case {c}: no output rc={r.returncode}")
