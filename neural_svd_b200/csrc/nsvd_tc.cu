// nsvd_tc.cu — tcgen05 / TMEM / TMA engine (bf16x3) of the NestedLoRA step.
//
// Every dense contraction D = A . B^T with fp32 operands is evaluated on the 5th-generation tensor
// cores as  A_hi B_hi + A_lo B_hi + A_hi B_lo  (bf16 operands, fp32 accumulation in TMEM), where
// v = v_hi + v_lo is the two-term bf16 split (16 significant bits).  Kernels are persistent and
// warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM
// allocator, warps 4..11 = epilogue (TMEM -> registers -> fused math -> global).
#include "nsvd_simt.cuh"
#include "nsvd_tc.cuh"

namespace nsvd {

// ------------------------------------------------------------------------------------------
// host: tensor maps through the driver entry point (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return NSVD_E_NODEVICE;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u)",
              (int)r, base, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
              (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, box0, box1);
    return NSVD_E_BADARG;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// S1: "big GEMM" skeleton.  Tile 128 x 256, K chunk 64, hi/lo planes, 2 smem stages (96 KB each),
// two TMEM accumulator buffers (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1.  kMN = operands are MN-major in global/shared memory (weight-gradient GEMMs).
// ------------------------------------------------------------------------------------------
namespace big {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB per plane
constexpr int B_BYTES = BN * BK * 2;          // 32 KB per plane
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (4 + EPI_WARPS) * 32;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
}  // namespace big

struct BigShape {
  int m_tiles, n_tiles, batches, k_slices;  // tile index: mt fastest, then ks, nt, batch
  int k_chunks_per_slice, k_chunks_total;
  int a_batched, b_batched;                 // third TMA coordinate = batch index or 0
};
struct TileCoord {
  int b, nt, ks, mt;
};
__device__ __forceinline__ TileCoord decode_tile(const BigShape& s, int t) {
  TileCoord c;
  c.mt = t % s.m_tiles;
  t /= s.m_tiles;
  c.ks = t % s.k_slices;
  t /= s.k_slices;
  c.nt = t % s.n_tiles;
  c.b = t / s.n_tiles;
  return c;
}

template <bool kMN, class Epi>
__global__ void __launch_bounds__(big::THREADS, 1)
big_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                const BigShape shape, const Epi epi) {
  using namespace big;
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;               // [STAGES]
  uint64_t* empty = bars + STAGES;     // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES; // [2]
  uint64_t* tempty = tfull + 2;        // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = shape.m_tiles * shape.n_tiles * shape.batches * shape.k_slices;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmAh);
    tma_prefetch_desc(&tmAl);
    tma_prefetch_desc(&tmBh);
    tma_prefetch_desc(&tmBl);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        TileCoord c = decode_tile(shape, t);
        const int kc0 = c.ks * shape.k_chunks_per_slice;
        int kc1 = kc0 + shape.k_chunks_per_slice;
        if (kc1 > shape.k_chunks_total) kc1 = shape.k_chunks_total;
        const int ab = shape.a_batched ? c.b : 0, bb = shape.b_batched ? c.b : 0;
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1, 1);
          uint8_t* sA = smem + stage * STAGE_BYTES;
          uint8_t* sB = sA + 2 * A_BYTES;
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          if (!kMN) {
            tma_load_3d(sA, &tmAh, &full[stage], kc * BK, c.mt * BM, ab);
            tma_load_3d(sA + A_BYTES, &tmAl, &full[stage], kc * BK, c.mt * BM, ab);
            tma_load_3d(sB, &tmBh, &full[stage], kc * BK, c.nt * BN, bb);
            tma_load_3d(sB + B_BYTES, &tmBl, &full[stage], kc * BK, c.nt * BN, bb);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) {
              tma_load_3d(sA + i * 8192, &tmAh, &full[stage], c.mt * BM + i * 64, kc * BK, ab);
              tma_load_3d(sA + A_BYTES + i * 8192, &tmAl, &full[stage], c.mt * BM + i * 64, kc * BK, ab);
            }
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) {
              tma_load_3d(sB + i * 8192, &tmBh, &full[stage], c.nt * BN + i * 64, kc * BK, bb);
              tma_load_3d(sB + B_BYTES + i * 8192, &tmBl, &full[stage], c.nt * BN + i * 64, kc * BK, bb);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, kMN ? 1 : 0, kMN ? 1 : 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        TileCoord c = decode_tile(shape, t);
        const int kc0 = c.ks * shape.k_chunks_per_slice;
        int kc1 = kc0 + shape.k_chunks_per_slice;
        if (kc1 > shape.k_chunks_total) kc1 = shape.k_chunks_total;
        mbar_wait(&tempty[acc], acc_phase ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(&full[stage], phase, 3);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sB = sA + 2 * A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            uint64_t ah, al, bh, bl;
            if (!kMN) {
              ah = make_sdesc_sw128(sA + kk * 32, 16, 1024);
              al = make_sdesc_sw128(sA + A_BYTES + kk * 32, 16, 1024);
              bh = make_sdesc_sw128(sB + kk * 32, 16, 1024);
              bl = make_sdesc_sw128(sB + B_BYTES + kk * 32, 16, 1024);
            } else {
              ah = make_sdesc_sw128(sA + kk * 2048, 8192, 1024);
              al = make_sdesc_sw128(sA + A_BYTES + kk * 2048, 8192, 1024);
              bh = make_sdesc_sw128(sB + kk * 2048, 8192, 1024);
              bl = make_sdesc_sw128(sB + B_BYTES + kk * 2048, 8192, 1024);
            }
            umma_f16(d_tmem, al, bh, idesc, (kc > kc0 || kk > 0) ? 1u : 0u);
            umma_f16(d_tmem, ah, bl, idesc, 1u);
            umma_f16(d_tmem, ah, bh, idesc, 1u);
          }
          umma_commit(&empty[stage]);  // frees the smem stage when these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);      // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps =====================
    const int ewarp = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      TileCoord c = decode_tile(shape, t);
      mbar_wait(&tfull[acc], acc_phase, 4);
      tc_fence_after();
      epi(tmem_base + acc * BN, c, ewarp, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// plain epilogue: D[b][row][col] = acc   (self-test / generic GEMM)
struct StoreEpi {
  float* D;
  int M, N;
  long d_bs;
  int accumulate_atomic;
  __device__ __forceinline__ void operator()(uint32_t tmem_acc, const TileCoord& c, int ewarp, int lane) const {
    const int q = ewarp & 3, half = ewarp >> 2;
    const int row = c.mt * big::BM + q * 32 + lane;
    float* drow = D + (long)c.b * d_bs + (long)row * N;
#pragma unroll 1
    for (int ch = 0; ch < 8; ++ch) {
      const int col0 = half * 128 + ch * 16;
      float v[16];
      tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + col0, v);
      tc::tmem_ld_wait();
      if (row < M) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int col = c.nt * big::BN + col0 + i;
          if (col < N) {
            if (accumulate_atomic) atomicAdd(drow + col, v[i]);
            else drow[col] = v[i];
          }
        }
      }
    }
  }
};

// fp32 -> bf16 hi/lo planes (optionally transposing nothing: same layout)
__global__ void split_planes_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __nv_bfloat16 h, l;
  tc::split_bf16(src[i], h, l);
  hi[i] = h;
  lo[i] = l;
}

template <bool kMN, class Epi>
static int launch_big(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                      const BigShape& shape, const Epi& epi, cudaStream_t st) {
  static bool configured = false;
  auto kern = big_gemm_kernel<kMN, Epi>;
  if (!configured) {
    NSVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, big::SMEM_BYTES));
    configured = true;
  }
  int tiles = shape.m_tiles * shape.n_tiles * shape.batches * shape.k_slices;
  if (tiles <= 0) return 0;
  int grid = tiles < 148 ? tiles : 148;
  kern<<<grid, big::THREADS, big::SMEM_BYTES, st>>>(ah, al, bh, bl, shape, epi);
  NSVD_LAUNCH_CHECK();
  return 0;
}

int tc_gemm_selftest(const float* A, const float* B, float* D, int M, int N, int K, int a_kmajor, int b_kmajor,
                     void* work, size_t work_bytes, cudaStream_t st) {
  NSVD_CHECK_ARG(a_kmajor == b_kmajor, "selftest: both operands must share the major mode");
  NSVD_CHECK_ARG(M % 8 == 0 && N % 8 == 0 && K % 8 == 0, "selftest: M, N, K must be multiples of 8");
  size_t na = (size_t)M * K, nb = (size_t)N * K;
  size_t need = 2 * (na + nb) * sizeof(__nv_bfloat16) + 1024;
  if (work_bytes < need) {
    set_error("selftest work too small: %zu < %zu", work_bytes, need);
    return NSVD_E_WORKSPACE;
  }
  __nv_bfloat16* ah = (__nv_bfloat16*)(((uintptr_t)work + 255) & ~(uintptr_t)255);
  __nv_bfloat16* al = ah + na;
  __nv_bfloat16* bh = al + na;
  __nv_bfloat16* bl = bh + nb;
  split_planes_kernel<<<cdiv((long)na, 256), 256, 0, st>>>(A, ah, al, (long)na);
  NSVD_LAUNCH_CHECK();
  split_planes_kernel<<<cdiv((long)nb, 256), 256, 0, st>>>(B, bh, bl, (long)nb);
  NSVD_LAUNCH_CHECK();
  CUtensorMap mah, mal, mbh, mbl;
  int rc;
  BigShape s{};
  s.m_tiles = cdiv(M, big::BM);
  s.n_tiles = cdiv(N, big::BN);
  s.batches = 1;
  s.k_slices = 1;
  s.k_chunks_total = cdiv(K, big::BK);
  s.k_chunks_per_slice = s.k_chunks_total;
  StoreEpi epi{D, M, N, 0, 0};
  if (a_kmajor) {
    if ((rc = make_tmap_bf16_3d(&mah, ah, K, M, 1, (uint64_t)K * 2, (uint64_t)M * K * 2, 64, big::BM))) return rc;
    if ((rc = make_tmap_bf16_3d(&mal, al, K, M, 1, (uint64_t)K * 2, (uint64_t)M * K * 2, 64, big::BM))) return rc;
    if ((rc = make_tmap_bf16_3d(&mbh, bh, K, N, 1, (uint64_t)K * 2, (uint64_t)N * K * 2, 64, big::BN))) return rc;
    if ((rc = make_tmap_bf16_3d(&mbl, bl, K, N, 1, (uint64_t)K * 2, (uint64_t)N * K * 2, 64, big::BN))) return rc;
    return launch_big<false>(mah, mal, mbh, mbl, s, epi, st);
  }
  if ((rc = make_tmap_bf16_3d(&mah, ah, M, K, 1, (uint64_t)M * 2, (uint64_t)M * K * 2, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&mal, al, M, K, 1, (uint64_t)M * 2, (uint64_t)M * K * 2, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&mbh, bh, N, K, 1, (uint64_t)N * 2, (uint64_t)N * K * 2, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&mbl, bl, N, K, 1, (uint64_t)N * 2, (uint64_t)N * K * 2, 64, 64))) return rc;
  return launch_big<true>(mah, mal, mbh, mbl, s, epi, st);
}

// ------------------------------------------------------------------------------------------
// engine entry points (filled in below)
// ------------------------------------------------------------------------------------------
void tc_scratch_bytes(const nsvd_problem_t& pb, size_t* saved, size_t* work) {
  *saved = 256;
  *work = 256;
  (void)pb;
}
int tc_forward(const nsvd_problem_t&, const nsvd_params_t&, const float*, float*, float*, void*, void*, size_t,
               cudaStream_t) {
  set_error("tcgen05 engine: forward not built yet");
  return NSVD_E_BADARG;
}
int tc_backward(const nsvd_problem_t&, const nsvd_params_t&, const float*, const float*, const void*, nsvd_grads_t&,
                void*, size_t, cudaStream_t) {
  set_error("tcgen05 engine: backward not built yet");
  return NSVD_E_BADARG;
}

}  // namespace nsvd
