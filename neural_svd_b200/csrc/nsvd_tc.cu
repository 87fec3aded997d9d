// nsvd_tc.cu — tcgen05 / TMEM / TMA engine (f16x3) of the NestedLoRA step.
//
// Every dense contraction D = A . B^T with fp32 operands is evaluated on the 5th-generation tensor
// cores as  A_hi B_hi + A_lo B_hi + A_hi B_lo  (16-bit operand planes, fp32 accumulation in TMEM), where
// v = v_hi + v_lo is a two-plane split: fp16 planes of a power-of-two multiple of v on the operator path
// (22 significant bits, see "Operand plan" below), bf16 planes (16 bits) in the CDK loss.  Kernels are persistent and
// warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM
// allocator, warps 4.. = epilogue (TMEM -> registers -> fused math -> global / staged TMA stores):
// 8 epilogue warps in the layer-0 GEMMs (MMA-bound), 16 in the hidden-layer kernels (epilogue-bound).
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

// A training loop encodes the same ~45 descriptors every step (same scratch buffers, same shapes): the encoded maps are
// memoised per host thread in a small direct-mapped table keyed by every argument of the encoding (a descriptor is a
// pure function of them), which takes the driver call out of the launch-bound small-batch steps.
struct TmapKey {
  const void* base;
  uint64_t d0, d1, d2, s1, s2;
  uint32_t box0, box1;
  int swizzle, valid;
  bool operator==(const TmapKey& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && s1 == o.s1 && s2 == o.s2 && box0 == o.box0 &&
           box1 == o.box1 && swizzle == o.swizzle && valid == o.valid;
  }
};
struct TmapSlot {
  TmapKey key;
  CUtensorMap map;
};
constexpr int kTmapSlots = 512;

int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1,
                      int swizzle_bytes) {
  static thread_local TmapSlot* table = nullptr;
  if (!table) table = new TmapSlot[kTmapSlots]();   // (zero-initialised: valid = 0; lives as long as the thread)
  const TmapKey key{base, d0, d1, d2, stride1_bytes, stride2_bytes, box0, box1, swizzle_bytes, 1};
  uint64_t h = (uint64_t)(uintptr_t)base * 0x9E3779B97F4A7C15ull;
  h ^= (d1 + 0x632BE59BD9B4E019ull * d2 + ((uint64_t)box1 << 20) + ((uint64_t)box0 << 8) + (uint64_t)swizzle_bytes) *
       0xC2B2AE3D27D4EB4Full;
  TmapSlot& slot = table[(h >> 40) % kTmapSlots];
  if (slot.key == key) {
    *out = slot.map;
    return 0;
  }
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
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u)",
              (int)r, base, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
              (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, box0, box1);
    return NSVD_E_BADARG;
  }
  slot.key = key;
  slot.map = *out;
  return 0;
}

// The fp32 accumulation in TMEM truncates toward zero: every chained MMA shrinks the accumulator by 1.61e-8 relative,
// coherently - measured 1.53e-8 .. 1.64e-8 per MMA for chains of 12 .. 384 MMAs on two operand distributions
// (profiles/truncation_probe.py; K = 2048: 7.2e-6 -> 3.6e-6 after rescaling, what remains is the incoherent part).  Each
// finished chain is therefore multiplied by 1 + kTruncPerMma * (number of MMAs in the chain) when it leaves TMEM.
constexpr float kTruncPerMma = 1.61e-8f;

// Optional phase timelines of block 0 (development builds: nvcc -DNSVD_TIMELINE; see profiles/timeline_probe.py)
#ifdef NSVD_TIMELINE
__device__ long long g_timeline_l0[64 * 8];
#define NSVD_TL0(tile, slot, val) \
  do { if (blockIdx.x == 0 && (tile) < 64) g_timeline_l0[(tile) * 8 + (slot)] = (val); } while (0)
extern "C" int nsvd_debug_timeline_l0(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_timeline_l0, sizeof(long long) * 64 * 8);
}
#else
#define NSVD_TL0(tile, slot, val) do {} while (0)
#endif

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
  int m_tiles, n_tiles, batches, k_slices;  // tile index: mt (inside an m-group) fastest, then ks, nt, batch, m-group
  int k_chunks_per_slice, k_chunks_total;
  int a_batched, b_batched;                 // third TMA coordinate = batch index or 0
  int m_group;                              // m-tiles per group (0 = all): keeps the A rows of a group L2-resident
                                            // while every (batch, n-tile) of B sweeps over them
  int k_group;                              // k-slices per group (0 = off): same idea for the weight-gradient GEMM,
                                            // whose long dimension is K (points)
  int b_group;                              // batches per group (0 = off), K-major GEMMs with k_slices = 1: a group of
                                            // batches (copies) keeps its B tiles (weights) L2-resident while the m-tiles
                                            // stream past it ONCE per group - order: group, m-tile, (batch, n-tile) fastest
  int a_xbatch;                             // big2s, MN-major: batches of A are column blocks of ONE matrix, batch b starts
                                            // at column b * a_xbatch (0 = batches are the third TMA coordinate)
};
struct TileCoord {
  int b, nt, ks, mt;
};
__device__ __forceinline__ TileCoord decode_tile(const BigShape& s, int t) {
  TileCoord c;
  if (s.b_group > 0) {
    const int G = s.b_group;
    const int per_group_full = G * s.n_tiles * s.m_tiles;
    const int g = t / per_group_full;
    t -= g * per_group_full;
    const int b0 = g * G;
    const int gb = (s.batches - b0) < G ? (s.batches - b0) : G;   // batches in this (possibly last, smaller) group
    const int inner = gb * s.n_tiles;
    c.mt = t / inner;
    t -= c.mt * inner;
    c.nt = t % s.n_tiles;
    c.b = b0 + t / s.n_tiles;
    c.ks = 0;
    return c;
  }
  if (s.k_group > 0) {   // k-slice groups outermost; inside a group: ks fastest, then mt, nt, batch
    const int G = s.k_group;
    const int per_group_full = G * s.m_tiles * s.n_tiles * s.batches;
    const int g = t / per_group_full;
    t -= g * per_group_full;
    const int k0 = g * G;
    const int gk = (s.k_slices - k0) < G ? (s.k_slices - k0) : G;
    c.ks = k0 + t % gk;
    t /= gk;
    c.mt = t % s.m_tiles;
    t /= s.m_tiles;
    c.nt = t % s.n_tiles;
    c.b = t / s.n_tiles;
    return c;
  }
  const int G = s.m_group > 0 ? s.m_group : s.m_tiles;
  const int per_group_full = G * s.k_slices * s.n_tiles * s.batches;
  const int g = t / per_group_full;
  t -= g * per_group_full;
  const int m0 = g * G;
  const int gm = (s.m_tiles - m0) < G ? (s.m_tiles - m0) : G;   // tiles in this (possibly last, smaller) group
  c.mt = m0 + t % gm;
  t /= gm;
  c.ks = t % s.k_slices;
  t /= s.k_slices;
  c.nt = t % s.n_tiles;
  c.b = t / s.n_tiles;
  return c;
}

// FMT = plane format of A | plane format of B << 2 | (add the lo*lo product) << 4, compile-time so that the
// instruction descriptors stay immediates in the MMA issue loop (as kernel arguments they cost 5-6 % of the
// layer-0 GEMMs: the single issuing thread is on the critical path at 98 % tensor-pipe utilisation).
template <bool kMN, class Epi, int FMT = 0>
__global__ void __launch_bounds__(big::THREADS, 1)
big_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                const BigShape shape, const Epi epi) {
  using namespace big;
  using namespace tc;
  constexpr MmaDescs md = make_descs(BM, BN, kMN ? 1 : 0, kMN ? 1 : 0, FMT & 3, (FMT >> 2) & 3, (FMT >> 4) & 1);
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
            umma_f16(d_tmem, al, bh, md.lh, (kc > kc0 || kk > 0) ? 1u : 0u);
            umma_f16(d_tmem, ah, bl, md.hl, 1u);
            if (md.four) umma_f16(d_tmem, al, bl, md.ll, 1u);
            umma_f16(d_tmem, ah, bh, md.hh, 1u);
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

// ------------------------------------------------------------------------------------------
// S1 on CTA pairs (cta_group::2): the two SMs of a cluster share one 256 x 256 tile.  CTA r owns rows
// [128 r, 128 r + 128) of the M tile (K-major) or copy 2 b + r (MN-major weight gradient) and loads only
// half of the B tile (128 of the 256 N rows); the leader issues tcgen05.mma.cta_group::2 (M = 256) which
// reads A from each CTA's own shared memory and B from both.  Per SM this halves the B-operand shared
// memory traffic (operand reads 64 B/clk + TMA fills 42 B/clk instead of 96 + 62), which is what
// limited the single-CTA kernel, and the smaller stage (64 KB) allows a 3-deep ring.
// ------------------------------------------------------------------------------------------
namespace big2 {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 3;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB per plane
constexpr int BH_BYTES = (BN / 2) * BK * 2;   // 16 KB per plane (this CTA's half of B)
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (4 + EPI_WARPS) * 32;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
}  // namespace big2

template <bool kMN, class Epi, int FMT = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(big2::THREADS, 1)
big2_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                 const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                 const BigShape shape, const int batches_valid, const Epi epi) {
  using namespace big2;
  using namespace tc;
  constexpr MmaDescs md = make_descs(2 * BM, BN, kMN ? 1 : 0, kMN ? 1 : 0, FMT & 3, (FMT >> 2) & 3, (FMT >> 4) & 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                   // [STAGES]  (the leader's are used)
  uint64_t* empty = bars + STAGES;         // [STAGES]  (each CTA waits on its own)
  uint64_t* tfull = bars + 2 * STAGES;     // [2]       (each CTA waits on its own)
  uint64_t* tempty = tfull + 2;            // [2]       (the leader's are used, 2 x EPI_WARPS arrivals)
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_tiles = shape.m_tiles * shape.n_tiles * shape.batches * shape.k_slices;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

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
      mbar_init(&tempty[i], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        TileCoord c = decode_tile(shape, t);
        const int kc0 = c.ks * shape.k_chunks_per_slice;
        int kc1 = kc0 + shape.k_chunks_per_slice;
        if (kc1 > shape.k_chunks_total) kc1 = shape.k_chunks_total;
        const int bpair = kMN ? 2 * c.b + (int)rank : c.b;
        const int ab = shape.a_batched ? bpair : 0, bb = shape.b_batched ? c.b : 0;
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1, 31);
          uint8_t* sA = smem + stage * STAGE_BYTES;
          uint8_t* sB = sA + 2 * A_BYTES;
          const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
          if (!kMN) {
            const int arow = (2 * c.mt + (int)rank) * BM, brow = c.nt * BN + (int)rank * (BN / 2);
            tma_load_3d_pair(sA, &tmAh, lead_full, kc * BK, arow, ab);
            tma_load_3d_pair(sA + A_BYTES, &tmAl, lead_full, kc * BK, arow, ab);
            tma_load_3d_pair(sB, &tmBh, lead_full, kc * BK, brow, bb);
            tma_load_3d_pair(sB + BH_BYTES, &tmBl, lead_full, kc * BK, brow, bb);
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              tma_load_3d_pair(sA + i * 8192, &tmAh, lead_full, c.mt * BM + i * 64, kc * BK, ab);
              tma_load_3d_pair(sA + A_BYTES + i * 8192, &tmAl, lead_full, c.mt * BM + i * 64, kc * BK, ab);
              const int ncol = c.nt * BN + ((int)rank * 2 + i) * 64;
              tma_load_3d_pair(sB + i * 8192, &tmBh, lead_full, ncol, kc * BK, bb);
              tma_load_3d_pair(sB + BH_BYTES + i * 8192, &tmBl, lead_full, ncol, kc * BK, bb);
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
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        TileCoord c = decode_tile(shape, t);
        const int kc0 = c.ks * shape.k_chunks_per_slice;
        int kc1 = kc0 + shape.k_chunks_per_slice;
        if (kc1 > shape.k_chunks_total) kc1 = shape.k_chunks_total;
        mbar_wait(&tempty[acc], acc_phase ^ 1, 32);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(&full[stage], phase, 33);
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
              bl = make_sdesc_sw128(sB + BH_BYTES + kk * 32, 16, 1024);
            } else {
              ah = make_sdesc_sw128(sA + kk * 2048, 8192, 1024);
              al = make_sdesc_sw128(sA + A_BYTES + kk * 2048, 8192, 1024);
              bh = make_sdesc_sw128(sB + kk * 2048, 8192, 1024);
              bl = make_sdesc_sw128(sB + BH_BYTES + kk * 2048, 8192, 1024);
            }
            umma_f16_pair(d_tmem, al, bh, md.lh, (kc > kc0 || kk > 0) ? 1u : 0u);
            umma_f16_pair(d_tmem, ah, bl, md.hl, 1u);
            if (md.four) umma_f16_pair(d_tmem, al, bl, md.ll, 1u);
            umma_f16_pair(d_tmem, ah, bh, md.hh, 1u);
          }
          umma_commit_pair(&empty[stage]);   // frees the stage in both CTAs
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_pair(&tfull[acc]);       // accumulators of both CTAs complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps (both CTAs, own TMEM half) =====================
    const int ewarp = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      TileCoord c = decode_tile(shape, t);
      bool valid = true;
      if (kMN) {
        c.b = 2 * c.b + (int)rank;
        valid = c.b < batches_valid;
      } else {
        c.mt = 2 * c.mt + (int)rank;
      }
      mbar_wait(&tfull[acc], acc_phase, 34);
      tc_fence_after();
      if (valid) epi(tmem_base + acc * BN, c, ewarp, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still arrive on our barriers / read our shared memory until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// S1 on CTA pairs with SHORT ACCUMULATION CHAINS (the two layer-0 GEMMs).  The fp32 accumulator in TMEM truncates: the
// error of a chain grows linearly with the number of chained MMAs (1.9e-8 each: K = 2048 -> 7e-6 whatever the operand
// format, profiles/format_probe.py), which alone breaks the fixed-seed trajectory tolerance
// (profiles/trajectory_emulation.py: hh:chain=2048 -> 2.9e-3, chain=256 -> 1.2e-4).  Here a tile's K range is cut into
// sub-chains of `sub_chunks` K chunks; each sub-chain accumulates from zero in one of the two TMEM buffers and is added
// to REGISTER accumulators by the epilogue warps (round-to-nearest fp32 adds) while the next sub-chain runs in the
// other buffer.  16 epilogue warps: warp (q, sub) owns TMEM lanes [32 q, 32 q + 32) and 4 groups of 16 columns chosen
// by the epilogue functor (64 fp32 accumulators per thread); the fused math then runs on the registers, with both TMEM
// buffers already handed back to the MMA issuer.
// ------------------------------------------------------------------------------------------
namespace big2s {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 3;
constexpr int A_BYTES = BM * BK * 2;
constexpr int BH_BYTES = (BN / 2) * BK * 2;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;
constexpr int EPI_WARPS = 16;
constexpr int THREADS = (2 + EPI_WARPS) * 32;   // warp 0: TMA producer + TMEM allocator, warp 1: MMA issuer, 2..17: epilogue
constexpr int EPI_STAGING = 32768;              // output staging of the epilogue (8 KB per TMEM lane quarter)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_STAGING;
// End of the accumulation chain that starts at K chunk ks0 of a tile covering [kc0, kc1): the first two chains of a
// tile take `sub_first` chunks, the others `sub_chunks`.  While the epilogue warps run the fused math and stores of
// the previous tile (10k cycles) the MMA issuer can only fill the two TMEM buffers; two longer first chains give it
// that much look-ahead (6 + 6 chunks = 18k cycles) without lengthening the other chains (timeline_l0_probe.py).
__device__ __forceinline__ int chain_end(int ks0, int kc0, int kc1, int sub_chunks, int sub_first) {
  const int e = ks0 + ((ks0 - kc0) < 2 * sub_first ? sub_first : sub_chunks);
  return e < kc1 ? e : kc1;
}
}  // namespace big2s

template <bool kMN, class Epi, int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(big2s::THREADS, 1)
big2s_gemm_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                  const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
                  const BigShape shape, const int batches_valid, const int sub_chunks, const int sub_first,
                  const Epi epi) {
  using namespace big2s;
  using namespace tc;
  constexpr MmaDescs md = make_descs(2 * BM, BN, kMN ? 1 : 0, kMN ? 1 : 0, FMT & 3, (FMT >> 2) & 3, (FMT >> 4) & 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                   // [STAGES]  (the leader's are used)
  uint64_t* empty = bars + STAGES;         // [STAGES]  (each CTA waits on its own)
  uint64_t* tfull = bars + 2 * STAGES;     // [2]       (each CTA waits on its own)
  uint64_t* tempty = tfull + 2;            // [2]       (the leader's are used, 2 x EPI_WARPS arrivals)
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_tiles = shape.m_tiles * shape.n_tiles * shape.batches * shape.k_slices;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

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
      mbar_init(&tempty[i], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_keep = l2_policy_evict_last();
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        TileCoord c = decode_tile(shape, t);
        const int kc0 = c.ks * shape.k_chunks_per_slice;
        int kc1 = kc0 + shape.k_chunks_per_slice;
        if (kc1 > shape.k_chunks_total) kc1 = shape.k_chunks_total;
        const int bpair = kMN ? 2 * c.b + (int)rank : c.b;
        const int ab = shape.a_batched ? bpair : 0, bb = shape.b_batched ? c.b : 0;
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1, 51);
          uint8_t* sA = smem + stage * STAGE_BYTES;
          uint8_t* sB = sA + 2 * A_BYTES;
          const uint32_t lead_full = mapa_u32(smem_u32(&full[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * STAGE_BYTES);
          if (!kMN) {
            const int arow = (2 * c.mt + (int)rank) * BM, brow = c.nt * BN + (int)rank * (BN / 2);
            tma_load_3d_pair(sA, &tmAh, lead_full, kc * BK, arow, ab);
            tma_load_3d_pair(sA + A_BYTES, &tmAl, lead_full, kc * BK, arow, ab);
            if (shape.b_group > 0) {   // the weight tiles of the copy group are re-read for every m-tile: keep them in the L2
              tma_load_3d_pair_hint(sB, &tmBh, lead_full, kc * BK, brow, bb, pol_keep);
              tma_load_3d_pair_hint(sB + BH_BYTES, &tmBl, lead_full, kc * BK, brow, bb, pol_keep);
            } else {
              tma_load_3d_pair(sB, &tmBh, lead_full, kc * BK, brow, bb);
              tma_load_3d_pair(sB + BH_BYTES, &tmBl, lead_full, kc * BK, brow, bb);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int acol = c.mt * BM + i * 64 + ab * shape.a_xbatch, az = shape.a_xbatch ? 0 : ab;
              tma_load_3d_pair(sA + i * 8192, &tmAh, lead_full, acol, kc * BK, az);
              tma_load_3d_pair(sA + A_BYTES + i * 8192, &tmAl, lead_full, acol, kc * BK, az);
              const int ncol = c.nt * BN + ((int)rank * 2 + i) * 64;
              tma_load_3d_pair(sB + i * 8192, &tmBh, lead_full, ncol, kc * BK, bb);
              tma_load_3d_pair(sB + BH_BYTES + i * 8192, &tmBl, lead_full, ncol, kc * BK, bb);
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
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
#ifdef NSVD_TIMELINE
      int tl_i = 0;
#endif
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        TileCoord c = decode_tile(shape, t);
        const int kc0 = c.ks * shape.k_chunks_per_slice;
        int kc1 = kc0 + shape.k_chunks_per_slice;
        if (kc1 > shape.k_chunks_total) kc1 = shape.k_chunks_total;
#ifdef NSVD_TIMELINE
        long long tl_we = 0, tl_wf = 0, tl_t;
        NSVD_TL0(tl_i, 0, clock64());
#endif
        for (int ks0 = kc0, ks1; ks0 < kc1; ks0 = ks1) {        // one sub-chain per TMEM buffer
          ks1 = chain_end(ks0, kc0, kc1, sub_chunks, sub_first);
#ifdef NSVD_TIMELINE
          tl_t = clock64();
#endif
          mbar_wait(&tempty[acc], acc_phase ^ 1, 52);
#ifdef NSVD_TIMELINE
          tl_we += clock64() - tl_t;
#endif
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          for (int kc = ks0; kc < ks1; ++kc) {
#ifdef NSVD_TIMELINE
            tl_t = clock64();
#endif
            mbar_wait(&full[stage], phase, 53);
#ifdef NSVD_TIMELINE
            tl_wf += clock64() - tl_t;
#endif
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
                bl = make_sdesc_sw128(sB + BH_BYTES + kk * 32, 16, 1024);
              } else {
                ah = make_sdesc_sw128(sA + kk * 2048, 8192, 1024);
                al = make_sdesc_sw128(sA + A_BYTES + kk * 2048, 8192, 1024);
                bh = make_sdesc_sw128(sB + kk * 2048, 8192, 1024);
                bl = make_sdesc_sw128(sB + BH_BYTES + kk * 2048, 8192, 1024);
              }
              umma_f16_pair(d_tmem, al, bh, md.lh, (kc > ks0 || kk > 0) ? 1u : 0u);
              umma_f16_pair(d_tmem, ah, bl, md.hl, 1u);
              umma_f16_pair(d_tmem, ah, bh, md.hh, 1u);
            }
            umma_commit_pair(&empty[stage]);   // frees the stage in both CTAs
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit_pair(&tfull[acc]);       // this sub-chain is complete in both CTAs
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
#ifdef NSVD_TIMELINE
        NSVD_TL0(tl_i, 1, tl_we);
        NSVD_TL0(tl_i, 2, tl_wf);
        NSVD_TL0(tl_i, 3, clock64());
        ++tl_i;
#endif
      }
    }
  } else if (warp >= 2) {
    // ===================== epilogue warps (both CTAs, own TMEM half): drain sub-chains, then the fused math ==========
    // a warp reaches the TMEM lanes [32 (warp % 4), +32): q follows the hardware warp index
    const int q = warp & 3, sub = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
#ifdef NSVD_TIMELINE
    int tl_i = 0;
    const bool tl_on = warp == 2 && lane == 0;
#endif
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      TileCoord c = decode_tile(shape, t);
      const int kc0 = c.ks * shape.k_chunks_per_slice;
      int kc1 = kc0 + shape.k_chunks_per_slice;
      if (kc1 > shape.k_chunks_total) kc1 = shape.k_chunks_total;
#ifdef NSVD_TIMELINE
      long long tl_dr = 0, tl_t = 0;
#endif
      bool valid = true;
      if (kMN) {
        c.b = 2 * c.b + (int)rank;
        valid = c.b < batches_valid;
      } else {
        c.mt = 2 * c.mt + (int)rank;
      }
      float r[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) r[i] = 0.f;
      for (int ks0 = kc0, ks1; ks0 < kc1; ks0 = ks1) {
        ks1 = chain_end(ks0, kc0, kc1, sub_chunks, sub_first);
        mbar_wait(&tfull[acc], acc_phase, 54);
#ifdef NSVD_TIMELINE
        tl_t = clock64();
        if (tl_on && ks0 == kc0) NSVD_TL0(tl_i, 4, tl_t);
#endif
        tc_fence_after();
        const uint32_t tl = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
        const int nk = (ks1 - ks0) * (BK / 16) * 3;   // MMAs in this chain
        const float corr = 1.f + kTruncPerMma * (float)nk;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v[16];
          tmem_ld16(tl + Epi::col0(sub, j), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[j * 16 + i] = fmaf(v[i], corr, r[j * 16 + i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[acc]), 0));
#ifdef NSVD_TIMELINE
        tl_dr += clock64() - tl_t;
#endif
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
#ifdef NSVD_TIMELINE
      if (tl_on) NSVD_TL0(tl_i, 5, clock64());
#endif
      if (valid) epi(r, c, q, sub, lane, smem + STAGES * STAGE_BYTES + 256);
#ifdef NSVD_TIMELINE
      if (tl_on) {
        NSVD_TL0(tl_i, 6, clock64());
        NSVD_TL0(tl_i, 7, tl_dr);
      }
      ++tl_i;
#endif
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still arrive on our barriers / read our shared memory until here
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
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
                                    __nv_bfloat16* __restrict__ lo, long n, int fmt) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint16_t h, l;
  if (fmt == tc::PF_BB) tc::split1<tc::PF_BB>(src[i], h, l);
  else if (fmt == tc::PF_BH) tc::split1<tc::PF_BH>(src[i], h, l);
  else tc::split1<tc::PF_HH>(src[i], h, l);
  reinterpret_cast<uint16_t*>(hi)[i] = h;
  reinterpret_cast<uint16_t*>(lo)[i] = l;
}

template <bool kMN, class Epi, int FMT = 0>
static int launch_big(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                      const BigShape& shape, const Epi& epi, cudaStream_t st) {
  auto kern = big_gemm_kernel<kMN, Epi, FMT>;
  NSVD_SMEM_OPTIN(kern, big::SMEM_BYTES);
  int tiles = shape.m_tiles * shape.n_tiles * shape.batches * shape.k_slices;
  if (tiles <= 0) return 0;
  int grid = tiles < 148 ? tiles : 148;
  kern<<<grid, big::THREADS, big::SMEM_BYTES, st>>>(ah, al, bh, bl, shape, epi);
  NSVD_LAUNCH_CHECK();
  return 0;
}

template <bool kMN, class Epi, int FMT = 0>
static int launch_big2(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                       const BigShape& shape, int batches_valid, const Epi& epi, cudaStream_t st) {
  auto kern = big2_gemm_kernel<kMN, Epi, FMT>;
  NSVD_SMEM_OPTIN(kern, big2::SMEM_BYTES);
  int tiles = shape.m_tiles * shape.n_tiles * shape.batches * shape.k_slices;
  if (tiles <= 0) return 0;
  int clusters = tiles < 74 ? tiles : 74;
  kern<<<2 * clusters, big2::THREADS, big2::SMEM_BYTES, st>>>(ah, al, bh, bl, shape, batches_valid, epi);
  NSVD_LAUNCH_CHECK();
  return 0;
}

constexpr int kFmtHH = tc::PF_HH | (tc::PF_HH << 2);   // fp16 hi/lo planes for both operands

template <bool kMN, class Epi, int FMT>
static int launch_big2s(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                        const BigShape& shape, int batches_valid, int sub_chunks, int sub_first, const Epi& epi,
                        cudaStream_t st) {
  auto kern = big2s_gemm_kernel<kMN, Epi, FMT>;
  NSVD_SMEM_OPTIN(kern, big2s::SMEM_BYTES);
  int tiles = shape.m_tiles * shape.n_tiles * shape.batches * shape.k_slices;
  if (tiles <= 0) return 0;
  int clusters = tiles < 74 ? tiles : 74;
  if (sub_chunks <= 0 || sub_chunks > shape.k_chunks_per_slice) sub_chunks = shape.k_chunks_per_slice;
  if (sub_first < sub_chunks) sub_first = sub_chunks;
  kern<<<2 * clusters, big2s::THREADS, big2s::SMEM_BYTES, st>>>(ah, al, bh, bl, shape, batches_valid, sub_chunks, sub_first,
                                                                epi);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// CTA pairs are the default for the two layer-0 GEMMs; NSVD_TC_PAIR=0 selects the single-CTA kernel.
static bool tc_use_pair() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("NSVD_TC_PAIR");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

int tc_gemm_selftest(const float* A, const float* B, float* D, int M, int N, int K, int a_kmajor, int b_kmajor,
                     void* work, size_t work_bytes, cudaStream_t st) {
  NSVD_CHECK_ARG(a_kmajor == b_kmajor, "selftest: both operands must share the major mode");
  // mode bits: 0 = K-major, 1 = CTA-pair kernel, 4-5 = plane format of A, 6-7 = plane format of B, 8 = add lo*lo
  const int fa = (a_kmajor >> 4) & 3, fb = (a_kmajor >> 6) & 3, four = (a_kmajor >> 8) & 1;
  NSVD_CHECK_ARG(fa <= tc::PF_HH && fb <= tc::PF_HH, "selftest: unknown plane format");
  const int fmt = fa | (fb << 2) | (four << 4);
  // non-default formats exist on the single-CTA K-major kernel only (profiles/format_probe.py)
  NSVD_CHECK_ARG(fmt == 0 || ((a_kmajor & 3) == 1), "selftest: plane formats other than bf16+bf16 need mode 1");
  const bool pair = (a_kmajor & 2) != 0;
  a_kmajor &= 1;
  b_kmajor &= 1;
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
  split_planes_kernel<<<cdiv((long)na, 256), 256, 0, st>>>(A, ah, al, (long)na, fa);
  NSVD_LAUNCH_CHECK();
  split_planes_kernel<<<cdiv((long)nb, 256), 256, 0, st>>>(B, bh, bl, (long)nb, fb);
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
    if (pair) {
      if ((rc = make_tmap_bf16_3d(&mbh, bh, K, N, 1, (uint64_t)K * 2, (uint64_t)N * K * 2, 64, big::BN / 2))) return rc;
      if ((rc = make_tmap_bf16_3d(&mbl, bl, K, N, 1, (uint64_t)K * 2, (uint64_t)N * K * 2, 64, big::BN / 2))) return rc;
      s.m_tiles = cdiv(M, 2 * big::BM);
      return launch_big2<false>(mah, mal, mbh, mbl, s, 1, epi, st);
    }
    if ((rc = make_tmap_bf16_3d(&mbl, bl, K, N, 1, (uint64_t)K * 2, (uint64_t)N * K * 2, 64, big::BN))) return rc;
    switch (fmt) {
      case 0: return launch_big<false>(mah, mal, mbh, mbl, s, epi, st);
      case 16: return launch_big<false, StoreEpi, 16>(mah, mal, mbh, mbl, s, epi, st);            // bf16 planes, 4 products
      case 10: return launch_big<false, StoreEpi, 10>(mah, mal, mbh, mbl, s, epi, st);            // fp16+fp16 both
      case 9: return launch_big<false, StoreEpi, 9>(mah, mal, mbh, mbl, s, epi, st);              // A bf16+fp16, B fp16+fp16
      case 25: return launch_big<false, StoreEpi, 25>(mah, mal, mbh, mbl, s, epi, st);
      case 5: return launch_big<false, StoreEpi, 5>(mah, mal, mbh, mbl, s, epi, st);              // bf16+fp16 both
      case 21: return launch_big<false, StoreEpi, 21>(mah, mal, mbh, mbl, s, epi, st);
      default: set_error("selftest: plane format combination %d not instantiated", fmt); return NSVD_E_BADARG;
    }
  }
  if ((rc = make_tmap_bf16_3d(&mah, ah, M, K, 1, (uint64_t)M * 2, (uint64_t)M * K * 2, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&mal, al, M, K, 1, (uint64_t)M * 2, (uint64_t)M * K * 2, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&mbh, bh, N, K, 1, (uint64_t)N * 2, (uint64_t)N * K * 2, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&mbl, bl, N, K, 1, (uint64_t)N * 2, (uint64_t)N * K * 2, 64, 64))) return rc;
  if (pair) {   // MN-major pair kernel stacks two batches along M: here batch 1 does not exist (zero fill, skipped)
    s.m_tiles = cdiv(M, big::BM);
    s.a_batched = 1;
    NSVD_CHECK_ARG(M <= big::BM, "selftest: MN-major pair mode takes M <= 128");
    return launch_big2<true>(mah, mal, mbh, mbl, s, 1, epi, st);
  }
  return launch_big<true>(mah, mal, mbh, mbl, s, epi, st);
}

// ------------------------------------------------------------------------------------------
// shared epilogue math
// ------------------------------------------------------------------------------------------
// softplus on the value stream and its forward-mode derivative streams.
// Epilogue-rate version: 3 MUFU ops (ex2, lg2, rcp) and ~20 FP32 instructions per (point, unit):
//   e = exp(-|z|), a = max(z,0) + log(1+e), sigma = z>=0 ? 1/(1+e) : e/(1+e).
// For z > 20: e < 2.1e-9 so 1+e == 1 and a == z, sigma == 1 exactly: torch's threshold=20 branch
// falls out without a compare.  Absolute error ~1e-7 (below the 2^-17 operand rounding that follows).
__device__ __forceinline__ void act_streams(float z0, float z1, float z2, float z3, float& a0, float& a1,
                                            float& a2, float& a3) {
  float e = __expf(-fabsf(z0));
  float u = 1.f + e;
#ifdef NSVD_RCP_RN
  float inv = __frcp_rn(u);
#else
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(u));   // one MUFU op, 1 ulp (u in [1, 2])
#endif
  float sg = z0 >= 0.f ? inv : e * inv;
  a0 = fmaxf(z0, 0.f) + __logf(u);
  a1 = sg * z1;
  a2 = sg * z2;
  a3 = fmaf(sg, z3, sg * (1.f - sg) * fmaf(z1, z1, z2 * z2));
}
// the value-stream half of act_streams: a = softplus(z), sg = sigmoid(z) (same arithmetic, so both epilogue orders agree)
__device__ __forceinline__ void softplus_sigmoid(float z0, float& a0, float& sg) {
  float e = __expf(-fabsf(z0));
  float u = 1.f + e;
#ifdef NSVD_RCP_RN
  float inv = __frcp_rn(u);
#else
  float inv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(u));
#endif
  sg = z0 >= 0.f ? inv : e * inv;
  a0 = fmaxf(z0, 0.f) + __logf(u);
}
// sigmoid(z) from a = softplus(z) at epilogue rate: 1 - exp(-a)  (absolute error ~6e-8)
__device__ __forceinline__ float sig_fast(float a) { return 1.f - __expf(-a); }

// 16 fp32 values -> 16 bf16 hi + 16 bf16 lo, two 16-byte stores each (dst 32-byte aligned)
__device__ __forceinline__ void store_split16(const float* v, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) tc::split_bf16x2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  uint4* ph = reinterpret_cast<uint4*>(hi);
  uint4* pl = reinterpret_cast<uint4*>(lo);
  ph[0] = make_uint4(h[0], h[1], h[2], h[3]);
  ph[1] = make_uint4(h[4], h[5], h[6], h[7]);
  pl[0] = make_uint4(l[0], l[1], l[2], l[3]);
  pl[1] = make_uint4(l[4], l[5], l[6], l[7]);
}
// 16 consecutive values hi+lo -> fp32
__device__ __forceinline__ void load_merge16(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* v) {
  uint4 h[2] = {reinterpret_cast<const uint4*>(hi)[0], reinterpret_cast<const uint4*>(hi)[1]};
  uint4 l[2] = {reinterpret_cast<const uint4*>(lo)[0], reinterpret_cast<const uint4*>(lo)[1]};
  const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(h);
  const __nv_bfloat162* ll = reinterpret_cast<const __nv_bfloat162*>(l);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 a = __bfloat1622float2(hh[i]), b = __bfloat1622float2(ll[i]);
    v[2 * i] = a.x + b.x;
    v[2 * i + 1] = a.y + b.y;
  }
}
// fp16 hi/lo (PF_HH) versions
__device__ __forceinline__ void store_split16h(const float* v, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) tc::split2<tc::PF_HH>(v[2 * i], v[2 * i + 1], h[i], l[i]);
  uint4* ph = reinterpret_cast<uint4*>(hi);
  uint4* pl = reinterpret_cast<uint4*>(lo);
  ph[0] = make_uint4(h[0], h[1], h[2], h[3]);
  ph[1] = make_uint4(h[4], h[5], h[6], h[7]);
  pl[0] = make_uint4(l[0], l[1], l[2], l[3]);
  pl[1] = make_uint4(l[4], l[5], l[6], l[7]);
}
__device__ __forceinline__ float merge1h(uint32_t hbits, uint32_t lbits) {   // one fp16 hi + fp16 lo pair -> fp32
  return __half2float(__ushort_as_half((unsigned short)hbits)) + __half2float(__ushort_as_half((unsigned short)lbits));
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// ------------------------------------------------------------------------------------------
// Operand plan.  Every MMA operand v is stored as two fp16 planes of s*v (PF_HH: hi = fp16(s v), lo = fp16(s v - hi)):
// 22 significant bits for |s v| >= 2^-3 and an absolute error of 2^-25 below (fp16 subnormals are honoured by the tensor
// cores).  s is a power of two per (tensor, copy, stream) taken from a RIGOROUS bound of |v|, so that |s v| <= 2^15 can
// never overflow (fp16 max 65504); the bounds are propagated from the weights (forward) and from max|dF| (backward) by
// the small kernels below - nothing is measured on the activations, nothing depends on an earlier step.  The formulas are
// restated and checked on the CPU in profiles/scale_plan_emulation.py:
//   features          |phi_s[j]| <= c_s[j],  c = (1, |B_0j|, |B_1j|, B_0j^2 + B_1j^2)
//   layer 0           |z_s[h]| <= R0_s = max_h sum_j (|W0[h,j]| + |W0[h,M+j]|) c_s[j]   (+ max|b0| for s = 0)
//   softplus streams  a_0 <= z_0 + ln 2,  |a_d| <= z_d,  |a_3| <= z_3 + (z_1^2 + z_2^2) / 4
//   hidden layer i    |z_s| <= (max_h sum_k |W_i[h,k]|) A_s   (+ max|b_i| for s = 0)
//   backward          |dZ2| <= |c| max|dF| max|W3|,  |dZ_{i-1}| <= (max_k sum_j |W_i[j,k]|) |dZ_i|
// Why not bf16 planes (round 1): two bf16 planes carry 16 bits, and the 1e-5 per-step error they leave fails the 1e-3
// fixed-seed trajectory tolerance (tests/test_gpu_training_run.py, profiles/trajectory_emulation.py); three bf16 planes
// would double the MMA count; mixed fp16 x bf16 MMAs are illegal (profiles/format_probe.py).
// ------------------------------------------------------------------------------------------
enum PlanSlot : int {
  PL_SW0 = 0,      // [4] scale of the folded layer-0 weights, per stream
  PL_INV_W0 = 4,   // [4] 1 / PL_SW0 (Phi is stored unscaled)
  PL_SA0 = 8,      // [4] scale of the a0 streams (layer-0 output)
  PL_SW1 = 12,
  PL_U1 = 13,      // [4] 1 / (SA0[s] SW1): layer-1 accumulator -> z
  PL_SA1 = 17,     // [4]
  PL_SW2 = 21,
  PL_U2 = 22,      // [4] 1 / (SA1[s] SW2)
  PL_SA2 = 26,     // value stream only (the derivative streams of a2 never leave the kernel)
  PL_INV_SA0 = 27, PL_INV_SA1 = 28, PL_INV_SA2 = 29,   // 1 / SA_i[0]: the backward reads the saved value streams
  PL_SDZ2 = 30, PL_SDZ1 = 31, PL_SDZ0 = 32,            // scales of the dZ_i planes
  PL_UD2 = 33,     // 1 / (SW2 SDZ2): layer-2 dgrad accumulator -> dA1
  PL_UD1 = 34,
  PL_UW2 = 35,     // 1 / (SDZ2 SA1[0]): layer-2 wgrad accumulator -> dW2
  PL_UW1 = 36,
  PL_UW0 = 37,     // 1 / SDZ0
  PL_STRIDE = 40
};
constexpr float kPlaneTarget = 32768.f;   // 2^15

__device__ __forceinline__ float pow2_scale(float bound) {
  if (!(bound > 0.f) || !isfinite(bound)) return 1.f;
  // 2^floor(log2(target / bound')) from the exponent field; 1.001 covers the rounding of the bound's own sums
  const float q = kPlaneTarget / (bound * 1.001f);
  int e = (int)((__float_as_uint(q) >> 23) & 0xffu) - 127;   // subnormal q -> -127, infinite q -> 128: both clamped
  e = e > 100 ? 100 : (e < -100 ? -100 : e);
  return __uint_as_float((unsigned)(e + 127) << 23);
}

// blocks [0, L*128): one per row (l, h) of W0: rowstat[(l*128 + h)*5 + s] = sum_j (|W0[h,j]| + |W0[h,M+j]|) c_s[j],
// [4] = max|W0[h,:]|.  blocks [L*128, L*128 + 2L): one per hidden matrix W = W_{i+1}[l]:
// hstat[(i*L + l)*3 + {0: max_h sum_k |W[h,k]|, 1: max|W|, 2: max_k sum_h |W[h,k]|}]; one more block: feature bounds
__global__ void __launch_bounds__(256) weight_stats_kernel(const float* __restrict__ W0, const float* __restrict__ Bff,
                                                           const float* __restrict__ W1, const float* __restrict__ W2,
                                                           float* __restrict__ rowstat, float* __restrict__ hstat,
                                                           int L, int M) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((int)blockIdx.x == L * kHidden + 2 * L) {
    // last block: bounds of the feature streams, hstat[6 L + {1, 2, 3}] = max_j {|B_0j|, |B_1j|, B_0j^2 + B_1j^2}
    float c[3] = {0.f, 0.f, 0.f};
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
      const float x0 = Bff[j], x1 = Bff[M + j];
      c[0] = fmaxf(c[0], fabsf(x0));
      c[1] = fmaxf(c[1], fabsf(x1));
      c[2] = fmaxf(c[2], fmaf(x0, x0, x1 * x1));
    }
    __shared__ float cred[8][3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      for (int o = 16; o > 0; o >>= 1) c[q] = fmaxf(c[q], __shfl_xor_sync(0xffffffffu, c[q], o));
      if (lane == 0) cred[warp][q] = c[q];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      float v = cred[0][threadIdx.x];
      for (int w = 1; w < 8; ++w) v = fmaxf(v, cred[w][threadIdx.x]);
      hstat[6L * L + 1 + threadIdx.x] = v;
    }
    return;
  }
  if ((int)blockIdx.x >= L * kHidden) {
    // hidden matrix: coalesced float4 sweep; a warp reads one row per step (row sum by shuffles), a thread always
    // meets the same four columns (column sums in registers)
    const int q = blockIdx.x - L * kHidden, i = q / L, l = q % L;
    const float4* W = reinterpret_cast<const float4*>((i == 0 ? W1 : W2) + (long)l * kHidden * kHidden);
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, rmax = 0.f, mx = 0.f;
#pragma unroll 4
    for (int it = 0; it < kHidden / 8; ++it) {          // 8 rows per step (8 warps)
      const float4 v = W[(it * 8 + warp) * (kHidden / 4) + lane];
      const float a0 = fabsf(v.x), a1 = fabsf(v.y), a2 = fabsf(v.z), a3 = fabsf(v.w);
      cs[0] += a0; cs[1] += a1; cs[2] += a2; cs[3] += a3;
      mx = fmaxf(mx, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)));
      float rs = (a0 + a1) + (a2 + a3);
      for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
      rmax = fmaxf(rmax, rs);
    }
    __shared__ float csum[8][kHidden], wred[8][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) csum[warp][lane * 4 + c] = cs[c];
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
      wred[warp][0] = rmax;
      wred[warp][1] = mx;
    }
    __syncthreads();
    if (warp == 0) {
      float cmax = 0.f;
      for (int c = lane; c < kHidden; c += 32) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += csum[w][c];
        cmax = fmaxf(cmax, t);
      }
      float r = lane < 8 ? wred[lane][0] : 0.f, m = lane < 8 ? wred[lane][1] : 0.f;
      for (int o = 16; o > 0; o >>= 1) {
        cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      }
      if (lane == 0) {
        float* o = hstat + ((long)i * L + l) * 3;
        o[0] = r;
        o[1] = m;
        o[2] = cmax;
      }
    }
    return;
  }
  const long row = blockIdx.x;
  const float* w = W0 + row * 2L * M;
  float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = threadIdx.x * 4; j < M; j += blockDim.x * 4) {
    const float4 s4 = *reinterpret_cast<const float4*>(w + j), c4 = *reinterpret_cast<const float4*>(w + M + j);
    const float4 x4 = *reinterpret_cast<const float4*>(Bff + j), y4 = *reinterpret_cast<const float4*>(Bff + M + j);
    const float sv[4] = {s4.x, s4.y, s4.z, s4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
    const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, yv[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float ws = fabsf(sv[k]), wc = fabsf(cv[k]), pr = ws + wc;
      a[0] += pr;
      a[1] = fmaf(pr, fabsf(xv[k]), a[1]);
      a[2] = fmaf(pr, fabsf(yv[k]), a[2]);
      a[3] = fmaf(pr, fmaf(xv[k], xv[k], yv[k] * yv[k]), a[3]);
      a[4] = fmaxf(a[4], fmaxf(ws, wc));
    }
  }
  __shared__ float red[8][5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    float v = a[i];
    for (int o = 16; o > 0; o >>= 1) {
      float t = __shfl_xor_sync(0xffffffffu, v, o);
      v = i == 4 ? fmaxf(v, t) : v + t;
    }
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float v = red[0][threadIdx.x];
    for (int wv = 1; wv < (int)(blockDim.x >> 5); ++wv)
      v = threadIdx.x == 4 ? fmaxf(v, red[wv][threadIdx.x]) : v + red[wv][threadIdx.x];
    rowstat[row * 5 + threadIdx.x] = v;
  }
}

// forward part of the plan: one warp per copy (grid = ceil(L / 4) blocks of 4 warps), lanes parallel over hidden units
__global__ void __launch_bounds__(128) fwd_plan_kernel(const float* __restrict__ rowstat, const float* __restrict__ hstat,
                                                       const float* __restrict__ Bff, const float* __restrict__ b0,
                                                       const float* __restrict__ b1, const float* __restrict__ b2,
                                                       float* __restrict__ plan, int L, int M) {
  const int lane = threadIdx.x & 31, l = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (l >= L) return;
  auto wmax = [](float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  };
  const float cmax[4] = {1.f, hstat[6L * L + 1], hstat[6L * L + 2], hstat[6L * L + 3]};
  float Z[4] = {0.f, 0.f, 0.f, 0.f}, mW0 = 0.f, mb[3] = {0.f, 0.f, 0.f};
  const float* bias[3] = {b0, b1, b2};
  for (int h = lane; h < kHidden; h += 32) {
    const float* r = rowstat + ((long)l * kHidden + h) * 5;
    for (int s = 0; s < 4; ++s) Z[s] = fmaxf(Z[s], r[s]);
    mW0 = fmaxf(mW0, r[4]);
    for (int i = 0; i < 3; ++i) mb[i] = fmaxf(mb[i], fabsf(bias[i][l * kHidden + h]));
  }
  for (int s = 0; s < 4; ++s) Z[s] = wmax(Z[s]);
  mW0 = wmax(mW0);
  for (int i = 0; i < 3; ++i) mb[i] = wmax(mb[i]);
  if (lane != 0) return;
  float* P = plan + (long)l * PL_STRIDE;
  for (int s = 0; s < 4; ++s) {
    const float sw = pow2_scale(mW0 * cmax[s]);
    P[PL_SW0 + s] = sw;
    P[PL_INV_W0 + s] = 1.f / sw;
  }
  float SA[3][4];
  for (int i = 0; i < 3; ++i) {
    Z[0] += mb[i];
    float A[4] = {Z[0] + 0.6931472f, Z[1], Z[2], Z[3] + 0.25f * (Z[1] * Z[1] + Z[2] * Z[2])};
    for (int s = 0; s < 4; ++s) SA[i][s] = pow2_scale(A[s]);
    if (i < 2) {
      const float R = hstat[((long)i * L + l) * 3 + 0];
      for (int s = 0; s < 4; ++s) Z[s] = R * A[s];
    }
  }
  const float sw1 = pow2_scale(hstat[((long)0 * L + l) * 3 + 1]), sw2 = pow2_scale(hstat[((long)1 * L + l) * 3 + 1]);
  P[PL_SW1] = sw1;
  P[PL_SW2] = sw2;
  for (int s = 0; s < 4; ++s) {
    P[PL_SA0 + s] = SA[0][s];
    P[PL_SA1 + s] = SA[1][s];
    P[PL_U1 + s] = 1.f / (SA[0][s] * sw1);
    P[PL_U2 + s] = 1.f / (SA[1][s] * sw2);
  }
  P[PL_SA2] = SA[2][0];
  P[PL_INV_SA0] = 1.f / SA[0][0];
  P[PL_INV_SA1] = 1.f / SA[1][0];
  P[PL_INV_SA2] = 1.f / SA[2][0];
}

// mdF[l] = max_b |dF[b, l]|  (mdF zero-initialised; non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256) col_absmax_kernel(const float* __restrict__ dF, long n, int L,
                                                         float* __restrict__ mdF) {
  // every thread keeps ONE column: the stride of the grid-stride loop is a multiple of L
  const long stride = ((long)gridDim.x * blockDim.x / L) * L;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (stride == 0 || i >= stride) return;
  float m = 0.f;
  for (; i < n; i += stride) m = fmaxf(m, fabsf(dF[i]));
  atomicMax(reinterpret_cast<unsigned int*>(mdF) + (((long)blockIdx.x * blockDim.x + threadIdx.x) % L), __float_as_uint(m));
}

// backward part of the plan for copy l, from max|dF[:, l]| and max|W3[l, :]|; evaluated redundantly by every block of
// the head backward kernel (its first block per copy publishes the slots the later kernels read)
struct BwdScales {
  float s2, s1, s0;
};
__device__ __forceinline__ BwdScales bwd_scales(const float* __restrict__ hstat, float mdf, float m3, float hard_mul_const,
                                                int l, int L) {
  const float dz2 = fabsf(hard_mul_const) * mdf * m3;            // sigma <= 1, rho <= 1, masks <= 1
  const float dz1 = hstat[((long)1 * L + l) * 3 + 2] * dz2;      // column L1 norm of W2
  const float dz0 = hstat[((long)0 * L + l) * 3 + 2] * dz1;      // column L1 norm of W1
  return BwdScales{pow2_scale(dz2), pow2_scale(dz1), pow2_scale(dz0)};
}
__device__ __forceinline__ void publish_bwd_plan(float* __restrict__ P, const BwdScales& b) {
  P[PL_SDZ2] = b.s2;
  P[PL_SDZ1] = b.s1;
  P[PL_SDZ0] = b.s0;
  P[PL_UD2] = 1.f / (P[PL_SW2] * b.s2);
  P[PL_UD1] = 1.f / (P[PL_SW1] * b.s1);
  P[PL_UW2] = 1.f / (b.s2 * P[PL_SA1]);
  P[PL_UW1] = 1.f / (b.s1 * P[PL_SA0]);
  P[PL_UW0] = 1.f / b.s0;
}

// ------------------------------------------------------------------------------------------
// weight / feature preparation (SIMT, HBM-bound, tiny next to the GEMMs)
// ------------------------------------------------------------------------------------------
// Folded layer-0 weights: all four streams are W'_s . [sin p ; cos p]   (probe10)
//   rows n = l*512 + (h/64)*256 + s*64 + (h%64),  K-major, fp16 hi/lo planes of PL_SW0[s] * W'_s.
// One thread = 8 consecutive features j of one row (l, h): 16-byte loads and stores.
__device__ __forceinline__ void fold_w0_body(long i, const float* __restrict__ W0, const float* __restrict__ Bff,
                                             const float* __restrict__ plan, __nv_bfloat16* __restrict__ hi,
                                             __nv_bfloat16* __restrict__ lo, int L, int M) {
  const int M8 = M >> 3;
  if (i >= (long)L * kHidden * M8) return;
  const int j = (int)(i % M8) * 8;
  const int h = (int)((i / M8) % kHidden);
  const int l = (int)(i / ((long)M8 * kHidden));
  const long K0 = 2L * M;
  const float* wrow = W0 + ((long)l * kHidden + h) * K0;
  float ws[8], wc[8], b0[8], b1[8];
  *reinterpret_cast<float4*>(ws) = *reinterpret_cast<const float4*>(wrow + j);
  *reinterpret_cast<float4*>(ws + 4) = *reinterpret_cast<const float4*>(wrow + j + 4);
  *reinterpret_cast<float4*>(wc) = *reinterpret_cast<const float4*>(wrow + M + j);
  *reinterpret_cast<float4*>(wc + 4) = *reinterpret_cast<const float4*>(wrow + M + j + 4);
  *reinterpret_cast<float4*>(b0) = *reinterpret_cast<const float4*>(Bff + j);
  *reinterpret_cast<float4*>(b0 + 4) = *reinterpret_cast<const float4*>(Bff + j + 4);
  *reinterpret_cast<float4*>(b1) = *reinterpret_cast<const float4*>(Bff + M + j);
  *reinterpret_cast<float4*>(b1 + 4) = *reinterpret_cast<const float4*>(Bff + M + j + 4);
  const long rbase = (long)l * 512 + (h / 64) * 256 + (h % 64);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const float sc = plan[(long)l * PL_STRIDE + PL_SW0 + s];
    float vs[8], vc[8];   // coefficients of sin p_j and cos p_j in stream s
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float nb2 = -(b0[k] * b0[k] + b1[k] * b1[k]);
      const float bd = s == 1 ? b0[k] : b1[k];
      vs[k] = (s == 0 ? ws[k] : s == 3 ? nb2 * ws[k] : -wc[k] * bd) * sc;
      vc[k] = (s == 0 ? wc[k] : s == 3 ? nb2 * wc[k] : ws[k] * bd) * sc;
    }
    uint32_t hs[4], ls[4], hc[4], lc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      tc::split2<tc::PF_HH>(vs[2 * k], vs[2 * k + 1], hs[k], ls[k]);
      tc::split2<tc::PF_HH>(vc[2 * k], vc[2 * k + 1], hc[k], lc[k]);
    }
    const long o = (rbase + s * 64) * K0 + j;
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(hs[0], hs[1], hs[2], hs[3]);
    *reinterpret_cast<uint4*>(lo + o) = make_uint4(ls[0], ls[1], ls[2], ls[3]);
    *reinterpret_cast<uint4*>(hi + o + M) = make_uint4(hc[0], hc[1], hc[2], hc[3]);
    *reinterpret_cast<uint4*>(lo + o + M) = make_uint4(lc[0], lc[1], lc[2], lc[3]);
  }
}

// both hidden weight tensors: fp16 hi/lo planes of PL_SW_i * W_i, as stored (forward B operand, K-major) and
// transposed, WT_i[l][c][r] = W_i[l][r][c] (dgrad B operand, K-major)
struct HiddenPlanes {
  __nv_bfloat16 *hi[2], *lo[2], *hiT[2], *loT[2];
};
// one block = one 32 x 32 tile of one matrix (blk = (which * L + l) * 16 + tile); the transposed planes go through a
// shared-memory tile so that both writes are coalesced
__device__ __forceinline__ void split_w_body(unsigned blk, const float* __restrict__ W1, const float* __restrict__ W2,
                                             const float* __restrict__ plan, const HiddenPlanes& o, int L) {
  __shared__ uint32_t tile[32][33];
  const int tl = blk & 15, l = (blk >> 4) % L, which = (blk >> 4) / L;
  const int r0 = (tl >> 2) * 32, c0 = (tl & 3) * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* W = (which ? W2 : W1) + (long)l * kHidden * kHidden;
  const float sc = plan[(long)l * PL_STRIDE + (which ? PL_SW2 : PL_SW1)];
  uint16_t* hi = reinterpret_cast<uint16_t*>(o.hi[which]) + (long)l * kHidden * kHidden;
  uint16_t* lo = reinterpret_cast<uint16_t*>(o.lo[which]) + (long)l * kHidden * kHidden;
  uint16_t* hiT = reinterpret_cast<uint16_t*>(o.hiT[which]) + (long)l * kHidden * kHidden;
  uint16_t* loT = reinterpret_cast<uint16_t*>(o.loT[which]) + (long)l * kHidden * kHidden;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    uint16_t a, b;
    tc::split1<tc::PF_HH>(W[r * kHidden + c] * sc, a, b);
    hi[r * kHidden + c] = a;
    lo[r * kHidden + c] = b;
    tile[ty + 8 * k][tx] = (uint32_t)a | ((uint32_t)b << 16);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t v = tile[tx][ty + 8 * k];
    const int it = (c0 + ty + 8 * k) * kHidden + r0 + tx;      // WT[c][r] = W[r][c]
    hiT[it] = (uint16_t)(v & 0xffffu);
    loT[it] = (uint16_t)(v >> 16);
  }
}

// Phi = [sin(x B), cos(x B)] as fp16 hi/lo planes (B, 2M), unscaled (|Phi| <= 1)   (examples/utils.py:139-140)
// One thread = 4 consecutive features of one point (8-byte stores).  The phase p = x.B is formed in fp32 exactly
// like the reference; a two-constant Cody-Waite step (k = rint(p / 2pi), r = p - k 2pi_hi - k 2pi_lo with FMAs,
// 3.5e-8 rms error) brings it to [-pi, pi].  kAccurate: sinf / cosf of the reduced argument (1 ulp) - the fp16 hi/lo
// planes carry 2^-22, so the 5e-7 of the MUFU approximations would be the largest error of the whole layer.
__device__ __forceinline__ void features_f16_body(long i, const float* __restrict__ x, const float* __restrict__ Bff,
                                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long P,
                                                  int M) {
  const int M4 = M >> 2;
  if (i >= P * M4) return;
  long p = i / M4;
  int j = (int)(i % M4) * 4;
  const float x0 = x[2 * p], x1 = x[2 * p + 1];
  const float4 b0 = *reinterpret_cast<const float4*>(Bff + j);
  const float4 b1 = *reinterpret_cast<const float4*>(Bff + M + j);
  const float bb0[4] = {b0.x, b0.y, b0.z, b0.w}, bb1[4] = {b1.x, b1.y, b1.z, b1.w};
  float sn[4], cs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float ph = fmaf(x1, bb1[k], x0 * bb0[k]);
    float kq = rintf(ph * 0.15915494309189535f);
    float r = fmaf(-kq, 6.2831855f, ph);          // 2pi_hi = fl(2 pi)
    r = fmaf(-kq, -1.7484555e-07f, r);            // 2pi_lo = 2 pi - 2pi_hi
    sincosf(r, &sn[k], &cs[k]);                   // |r| <= pi: the fast path of sincosf (no Payne-Hanek)
  }
  uint32_t h0, l0, h1, l1;
  long o = p * 2L * M + j;
  tc::split2<tc::PF_HH>(sn[0], sn[1], h0, l0);
  tc::split2<tc::PF_HH>(sn[2], sn[3], h1, l1);
  *reinterpret_cast<uint2*>(hi + o) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo + o) = make_uint2(l0, l1);
  tc::split2<tc::PF_HH>(cs[0], cs[1], h0, l0);
  tc::split2<tc::PF_HH>(cs[2], cs[3], h1, l1);
  *reinterpret_cast<uint2*>(hi + o + M) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo + o + M) = make_uint2(l0, l1);
}

// ---- finite-difference mode (pb.fd_eps > 0, pde/diff_ops.py:25-52): a second forward pass in which the four stream
// slots carry the four SHIFTED point sets x + eps e_0, x - eps e_0, x + eps e_1, x - eps e_1, value stream only.
// Phi of the shifted sets, rows s * P + p of a (4 P, 2 M) matrix, fp16 hi/lo planes
__global__ void features_shift_f16_kernel(const float* __restrict__ x, const float* __restrict__ Bff,
                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long P, int M,
                                          float eps) {
  const int M4 = M >> 2;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * P * M4) return;
  const long row = i / M4;
  const int j = (int)(i % M4) * 4, s = (int)(row / P);
  const long p = row % P;
  float y0 = x[2 * p], y1 = x[2 * p + 1];
  fd_shift(s, eps, y0, y1);
  const float4 b0 = *reinterpret_cast<const float4*>(Bff + j);
  const float4 b1 = *reinterpret_cast<const float4*>(Bff + M + j);
  const float bb0[4] = {b0.x, b0.y, b0.z, b0.w}, bb1[4] = {b1.x, b1.y, b1.z, b1.w};
  float sn[4], cs[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float ph = fmaf(y1, bb1[k], y0 * bb0[k]);
    float kq = rintf(ph * 0.15915494309189535f);
    float r = fmaf(-kq, 6.2831855f, ph);
    r = fmaf(-kq, -1.7484555e-07f, r);
    sincosf(r, &sn[k], &cs[k]);
  }
  uint32_t h0, l0, h1, l1;
  long o = row * 2L * M + j;
  tc::split2<tc::PF_HH>(sn[0], sn[1], h0, l0);
  tc::split2<tc::PF_HH>(sn[2], sn[3], h1, l1);
  *reinterpret_cast<uint2*>(hi + o) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo + o) = make_uint2(l0, l1);
  tc::split2<tc::PF_HH>(cs[0], cs[1], h0, l0);
  tc::split2<tc::PF_HH>(cs[2], cs[3], h1, l1);
  *reinterpret_cast<uint2*>(hi + o + M) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo + o + M) = make_uint2(l0, l1);
}
// W0 itself (not folded with the stream scalings), rows l * 128 + h, fp16 hi/lo planes of PL_SW0[0] * W0
__device__ __forceinline__ void split_w0_body(long i, const float* __restrict__ W0, const float* __restrict__ plan,
                                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int L, long K0) {
  long n = (long)L * kHidden * K0;
  if (i >= n) return;
  const int l = (int)(i / (kHidden * K0));
  uint16_t a, b;
  tc::split1<tc::PF_HH>(W0[i] * plan[(long)l * PL_STRIDE + PL_SW0], a, b);
  reinterpret_cast<uint16_t*>(hi)[i] = a;
  reinterpret_cast<uint16_t*>(lo)[i] = b;
}
// All per-call operand preparation in ONE launch (the small-batch configurations are bound by the number and the
// latency of these kernels): block ranges = features of all points | folded layer-0 weights | hidden weight planes
// (plain + transposed) | unfolded W0 planes (finite-difference pass only).
struct PrepArgs {
  const float *x, *Bff, *W0, *W1, *W2, *plan;
  __nv_bfloat16 *phi_hi, *phi_lo, *w0_hi, *w0_lo, *w0v_hi, *w0v_lo;
  HiddenPlanes hp;
  long B;
  int L, M;
  unsigned n_feat, n_fold, n_split;   // blocks of the first three ranges
};
__global__ void __launch_bounds__(256) prep_operands_kernel(PrepArgs a) {
  unsigned b = blockIdx.x;
  if (b < a.n_feat) return features_f16_body((long)b * 256 + threadIdx.x, a.x, a.Bff, a.phi_hi, a.phi_lo, a.B, a.M);
  b -= a.n_feat;
  if (b < a.n_fold) return fold_w0_body((long)b * 256 + threadIdx.x, a.W0, a.Bff, a.plan, a.w0_hi, a.w0_lo, a.L, a.M);
  b -= a.n_fold;
  if (b < a.n_split) return split_w_body(b, a.W1, a.W2, a.plan, a.hp, a.L);
  b -= a.n_split;
  split_w0_body((long)b * 256 + threadIdx.x, a.W0, a.plan, a.w0v_hi, a.w0v_lo, a.L, 2L * a.M);
}

// softplus alone at epilogue rate (value-only passes)
__device__ __forceinline__ float softplus_fast(float z) { return fmaxf(z, 0.f) + __logf(1.f + __expf(-fabsf(z))); }

// ------------------------------------------------------------------------------------------
// layer-0 forward epilogue (S1, K-major): tile = (copy l, hidden half hc, 128 points)
//   TMEM columns [s*64 + hh]; writes the 4 activation streams as hi/lo planes for layer 1 and
//   the value stream into the saved buffer.
// ------------------------------------------------------------------------------------------
struct L0FwdEpi {
  const float* bias;            // b0 (L,128)
  const float* plan;            // operand plan (PL_* slots per copy)
  __nv_bfloat16 *str_hi, *str_lo;  // [L][4][P][128]
  __nv_bfloat16 *sav_hi, *sav_lo;  // [L][Btot][128]
  int P;
  long Btot, p_off;
  // accumulator group j of epilogue warp `sub` = stream j, hidden units [16 sub, 16 sub + 16) of the tile's 64
  __device__ static __forceinline__ uint32_t col0(int sub, int j) { return (uint32_t)(j * 64 + sub * 16); }
  __device__ __forceinline__ void operator()(float (&r)[64], const TileCoord& c, int q, int sub, int lane,
                                             uint8_t* staging) const {
    const int l = c.b, h0 = c.nt * 64 + sub * 16;
    const float* pl = plan + (long)l * PL_STRIDE;
    const float i0 = __ldg(pl + PL_INV_W0), i1 = __ldg(pl + PL_INV_W0 + 1), i2 = __ldg(pl + PL_INV_W0 + 2),
                i3 = __ldg(pl + PL_INV_W0 + 3);
    const float s0 = __ldg(pl + PL_SA0), s1 = __ldg(pl + PL_SA0 + 1), s2 = __ldg(pl + PL_SA0 + 2),
                s3 = __ldg(pl + PL_SA0 + 3);
    float bs[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + l * kHidden + h0) + i);
      bs[4 * i] = b4.x;
      bs[4 * i + 1] = b4.y;
      bs[4 * i + 2] = b4.z;
      bs[4 * i + 3] = b4.w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float zb = fmaf(r[i], i0, bs[i]);
      float a0, a1, a2, a3;
#ifdef NSVD_PROBE_NOMATH
      a0 = zb; a1 = r[16 + i] * i1; a2 = r[32 + i] * i2; a3 = r[48 + i] * i3;
#else
      act_streams(zb, r[16 + i] * i1, r[32 + i] * i2, r[48 + i] * i3, a0, a1, a2, a3);
#endif
      r[i] = a0 * s0;
      r[16 + i] = a1 * s1;
      r[32 + i] = a2 * s2;
      r[48 + i] = a3 * s3;
    }
    // Stores.  A thread holds 16 units (32 bytes per plane) of ONE point row; written directly, every store
    // instruction would touch 32 different 128-byte lines (measured: 12k of the epilogue's 14k cycles,
    // profiles/timeline_l0_probe.py).  The four warps of a TMEM lane quarter (sub = 0..3: units [16 sub, +16) of the same
    // 32 rows) therefore stage one (stream, plane) at a time in shared memory as [32 rows][128 B] (16-byte pieces
    // XOR-swizzled with the row: conflict-free both ways) and write it back with every instruction covering four
    // full lines.  Two 4 KB buffers per quarter alternate, so one named barrier per round suffices.
    uint8_t* stq = staging + q * 8192;
    // the 32 KB per point written here are read by the next kernel only after gigabytes of other traffic: evict first,
    // so that they do not push the Phi group (re-used by every copy) out of the L2
    const uint64_t pol_once = tc::l2_policy_evict_first();
    const int gt = sub * 32 + lane;                     // thread index inside the quarter group (128 threads)
    const int row_base = c.mt * big::BM + q * 32;       // first point row of the quarter inside the micro-batch
    const uint32_t wr0 = (uint32_t)lane * 128 + ((uint32_t)((sub * 2) ^ (lane & 7)) << 4);
    const uint32_t wr1 = (uint32_t)lane * 128 + ((uint32_t)((sub * 2 + 1) ^ (lane & 7)) << 4);
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      uint32_t h[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) tc::split2<tc::PF_HH>(r[st * 16 + 2 * i], r[st * 16 + 2 * i + 1], h[i], lo[i]);
#pragma unroll
      for (int plane = 0; plane < 2; ++plane) {
        uint8_t* buf = stq + ((st * 2 + plane) & 1) * 4096;
        const uint32_t* w = plane ? lo : h;
        *reinterpret_cast<uint4*>(buf + wr0) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(buf + wr1) = make_uint4(w[4], w[5], w[6], w[7]);
        tc::named_bar_sync(1 + q, 128);
        // destination of this (stream, plane): the value stream lives only in `saved` (the next layer and the backward
        // both read it there); the derivative streams go to the micro-batch stream buffer
        __nv_bfloat16* dst = st == 0 ? (plane ? sav_lo : sav_hi) + ((long)l * Btot + p_off) * kHidden
                                     : (plane ? str_lo : str_hi) + ((long)l * 4 + st) * (long)P * kHidden;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int piece = gt + 128 * k, rr = piece >> 3, cc = piece & 7;
          const int pt = row_base + rr;
#ifdef NSVD_PROBE_NOSTORE
          if (pt < P && r[0] == 12345.678f)
#else
          if (pt < P)
#endif
            tc::st_global_v4_hint(reinterpret_cast<uint8_t*>(dst + (long)pt * kHidden + c.nt * 64) + cc * 16,
                                  *reinterpret_cast<const uint4*>(buf + rr * 128 + ((cc ^ (rr & 7)) << 4)), pol_once);
        }
      }
    }
    tc::named_bar_sync(1 + q, 128);   // the last round's reads are done before the next tile's first write
  }
};

// layer-0 VALUE-ONLY epilogue (finite-difference mode): tile = (256 stacked rows, 2 copies x 128 units);
// a0 = softplus(W0 phi + b0) of stacked row R = s * slot_rows + p goes to dst[l * copy_stride + s * slot_stride + p * 128].
//   shifted pass : 4 P rows (the four shifted point sets) -> slot s of the stream buffer [L][4][P][128]
//   central pass : P rows -> the saved value stream a0 [L][Btot][128] at row p_off + p (dst already offset)
struct L0ValEpi {
  const float* bias;            // b0 (L,128)
  const float* plan;
  __nv_bfloat16 *dst_hi, *dst_lo;
  long rows_total, slot_rows, slot_stride, copy_stride;
  int L;
  __device__ static __forceinline__ uint32_t col0(int sub, int j) { return (uint32_t)(sub * 64 + j * 16); }
  __device__ __forceinline__ void operator()(float (&r)[64], const TileCoord& c, int q, int sub, int lane,
                                             uint8_t*) const {
    const long R = (long)c.mt * big::BM + q * 32 + lane;
    const int l = 2 * c.nt + (sub >> 1), u0 = (sub & 1) * 64;
    if (R >= rows_total || l >= L) return;
    const long s = R / slot_rows, p = R % slot_rows;
    const float inv = __ldg(plan + (long)l * PL_STRIDE + PL_INV_W0), sa = __ldg(plan + (long)l * PL_STRIDE + PL_SA0);
#pragma unroll
    for (int i = 0; i < 64; ++i) r[i] = softplus_fast(fmaf(r[i], inv, __ldg(bias + l * kHidden + u0 + i))) * sa;
    const long o = (long)l * copy_stride + s * slot_stride + p * kHidden + u0;
#pragma unroll
    for (int j = 0; j < 4; ++j) store_split16h(&r[j * 16], dst_hi + o + j * 16, dst_lo + o + j * 16);
  }
};

// layer-0 weight-gradient epilogue (S1, MN-major): tile = (copy l, 256 features, k-slice of points)
//   dW0[l][j][n] += acc / SDZ0[l]   (fp32 vector reductions in L2)
struct L0WgradEpi {
  float* dW0;
  const float* plan;
  int K0;
  __device__ static __forceinline__ uint32_t col0(int sub, int j) { return (uint32_t)(sub * 64 + j * 16); }
  __device__ __forceinline__ void operator()(float (&r)[64], const TileCoord& c, int q, int sub, int lane,
                                             uint8_t*) const {
    const int j = q * 32 + lane;
    float* drow = dW0 + ((long)c.b * kHidden + j) * K0;
    const float un = __ldg(plan + (long)c.b * PL_STRIDE + PL_UW0);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n0 = c.nt * big::BN + sub * 64 + g * 16;
      if (n0 + 16 <= K0) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          red_add_v4(drow + n0 + i, r[g * 16 + i] * un, r[g * 16 + i + 1] * un, r[g * 16 + i + 2] * un,
                     r[g * 16 + i + 3] * un);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (n0 + i < K0) atomicAdd(drow + n0 + i, r[g * 16 + i] * un);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------
// S2-fwd: hidden layer i (128 -> 128) on the 4 streams of one copy; tile = (l, 128 points).
//   smem: W_i[l] hi/lo resident (64 KB) + 2 stages of one stream tile hi/lo (64 KB each).
//   TMEM: 4 x 128 columns (one block per stream).  kLast fuses the 128->1 head, the importance /
//   mask product rule, the potential and the operator scale/shift (F, TF).
// ------------------------------------------------------------------------------------------
namespace hid {
constexpr int PLANE = 128 * 128 * 2;  // one bf16 128x128 tile = two 16 KB K-chunks
constexpr int CHUNK = 16384;          // [128 rows][128 B]: 64 bf16 along K, 128-byte swizzle
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (4 + EPI_WARPS) * 32;
// forward kernel
constexpr int F_EPI_WARPS = 16;
constexpr int F_THREADS = (4 + F_EPI_WARPS) * 32;
constexpr int F_STAGES = 3;
constexpr int F_STAGE_BYTES = 2 * CHUNK;  // hi chunk + lo chunk of one stream
constexpr int F_BOX = 8192;               // staging box [128 rows][64 B] (32 bf16), 64-byte swizzle
constexpr int F_STAGING = 8 * F_BOX;      // 4 streams x {hi, lo}
constexpr int SMEM_FWD = 2 * PLANE + F_STAGES * F_STAGE_BYTES + F_STAGING + 1024 + 1152;
// pipelined backward kernel: half tiles of 64 points in a 2-stage ring
constexpr int HROWS = 64;
constexpr int HCHUNK = HROWS * 128;       // [64 rows][128 B]
constexpr int HPLANE = 2 * HCHUNK;        // one 64 x 128 bf16 plane
constexpr int B2_STAGE = 4 * HPLANE;      // dZ hi | dZ lo | a hi | a lo  = 64 KB
constexpr int SMEM_BWD2 = 2 * PLANE + 2 * B2_STAGE + 2 * HPLANE + 1024 + 1024;   // W | 2 stages | output staging
}  // namespace hid

// Optional phase timeline of block 0 (development builds: nvcc -DNSVD_TIMELINE; see profiles/timeline_probe.py)
#ifdef NSVD_TIMELINE
__device__ long long g_timeline[64 * 8];
#define NSVD_TL(tile, slot) \
  do { if (blockIdx.x == 0 && (tile) < 64) g_timeline[(tile) * 8 + (slot)] = clock64(); } while (0)
extern "C" int nsvd_debug_timeline(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(long long) * 64 * 8);
}
#else
#define NSVD_TL(tile, slot) do {} while (0)
#endif

struct HidFwdArgs {
  int L, P, m_tiles;
  long Btot, p_off;
  const float* bias;                 // b_i (L,128)
  const float* plan;                 // operand plan
  int u_slot, sa_slot;               // PL_U1 / PL_U2 (accumulator -> z, per stream), PL_SA1 / PL_SA2 (output scales)
  // kLast only
  const float* W3;                   // (L,128)
  const float* b3;                   // (L)
  const float* x;                    // (Btot,2)
  const float* mscales;              // (L) or null
  float *F, *TF, *U0;                // (Btot, L)
  nsvd_problem_t pb;
};

// S2-fwd: hidden layer i (128 -> 128) on the 4 streams of one copy; tile = (l, 128 points).
//   smem : W_i[l] hi/lo resident (64 KB) | ring of 3 x 32 KB (stream, K-chunk) operand stages |
//          64 KB output staging (8 boxes of 128 rows x 64 B) drained by TMA bulk stores.
//   TMEM : 4 x 128 columns (one block per stream).
//   kLast fuses the 128 -> 1 head, the importance / mask product rule, the potential and the
//   operator scale / shift (F, TF) and only stores the value stream.
template <bool kLast>
__global__ void __launch_bounds__(hid::F_THREADS, 1)
hidden_fwd_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                  const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                  const __grid_constant__ CUtensorMap tmOh, const __grid_constant__ CUtensorMap tmOl,
                  const __grid_constant__ CUtensorMap tmSh, const __grid_constant__ CUtensorMap tmSl,
                  const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                  const HidFwdArgs args) {
  using namespace hid;
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                       // [hi c0 | hi c1 | lo c0 | lo c1]
  uint8_t* sA = smem + 2 * PLANE;           // ring
  uint8_t* sO = sA + F_STAGES * F_STAGE_BYTES;  // staging boxes
  uint64_t* bars = (uint64_t*)(sO + F_STAGING);
  uint64_t* full = bars;                    // [3]
  uint64_t* empty = bars + 3;               // [3]
  uint64_t* wfull = bars + 6;
  uint64_t* wfree = bars + 7;
  uint64_t* tfull = bars + 8;
  uint64_t* tempty = bars + 9;
  uint32_t* tmem_slot = (uint32_t*)(bars + 10);
  float* bias_s = (float*)(bars + 16);      // [128]
  float* w3_s = bias_s + 128;               // [128] (kLast)
  float* ubuf = (float*)(sO + 6 * F_BOX);   // [128][3][4] head partial sums (kLast: boxes 2..7 are unused)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = args.L * args.m_tiles;
  const int tpc = (T + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * tpc;
  const int t_end = (t_begin + tpc < T) ? t_begin + tpc : T;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmAh);
    tma_prefetch_desc(&tmAl);
    tma_prefetch_desc(&tmWh);
    tma_prefetch_desc(&tmWl);
    tma_prefetch_desc(&tmSh);
    tma_prefetch_desc(&tmSl);
    tma_prefetch_desc(&tmVh);
    tma_prefetch_desc(&tmVl);
    if (!kLast) {
      tma_prefetch_desc(&tmOh);
      tma_prefetch_desc(&tmOl);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < F_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(wfull, 1);
    mbar_init(wfree, 1);
    mbar_init(tfull, 1);
    mbar_init(tempty, F_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, cur_l = -1;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int l = t / args.m_tiles, mt = t % args.m_tiles;
        const int i = t - t_begin;
        if (l != cur_l) {
          if (i > 0) mbar_wait(wfree, (uint32_t)((i - 1) & 1), 10);  // MMAs of the previous tile retired
          mbar_arrive_expect_tx(wfull, 2 * PLANE);
          tma_load_3d(sW, &tmWh, wfull, 0, 0, l);
          tma_load_3d(sW + CHUNK, &tmWh, wfull, 64, 0, l);
          tma_load_3d(sW + PLANE, &tmWl, wfull, 0, 0, l);
          tma_load_3d(sW + PLANE + CHUNK, &tmWl, wfull, 64, 0, l);
          cur_l = l;
        }
        for (int sc = 0; sc < 8; ++sc) {  // (stream, K-chunk)
          const int s = sc >> 1, c = sc & 1;
          mbar_wait(&empty[stage], phase ^ 1, 11);
          uint8_t* d = sA + stage * F_STAGE_BYTES;
          mbar_arrive_expect_tx(&full[stage], F_STAGE_BYTES);
          if (s == 0) {   // input value stream: from `saved` (whole-batch rows), not duplicated in the stream buffer
            tma_load_3d(d, &tmVh, &full[stage], 64 * c, (int)args.p_off + mt * 128, l);
            tma_load_3d(d + CHUNK, &tmVl, &full[stage], 64 * c, (int)args.p_off + mt * 128, l);
          } else {
            tma_load_3d(d, &tmAh, &full[stage], 64 * c, mt * 128, l * 4 + s);
            tma_load_3d(d + CHUNK, &tmAl, &full[stage], 64 * c, mt * 128, l * 4 + s);
          }
          if (++stage == F_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        NSVD_TL(i, 6);   // all loads of the tile issued
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, 128, 0, 0, false, false);   // fp16 x fp16 planes
      int stage = 0, cur_l = -1;
      uint32_t phase = 0, wphase = 0, tphase = 0;
      const uint32_t w_hi = smem_u32(sW), w_lo = w_hi + PLANE;
      for (int t = t_begin; t < t_end; ++t) {
        const int l = t / args.m_tiles;
        if (l != cur_l) {
          mbar_wait(wfull, wphase, 12);
          wphase ^= 1;
          cur_l = l;
        }
        mbar_wait(tempty, tphase ^ 1, 13);
        tc_fence_after();
        NSVD_TL(t - t_begin, 0);   // accumulators handed back
        for (int sc = 0; sc < 8; ++sc) {
          const int s = sc >> 1, c = sc & 1;
          mbar_wait(&full[stage], phase, 14);
          tc_fence_after();
          if (sc == 0) NSVD_TL(t - t_begin, 1);   // first operand stage present
          if (sc == 7) NSVD_TL(t - t_begin, 2);   // last operand stage present
          const uint32_t a_hi = smem_u32(sA + stage * F_STAGE_BYTES), a_lo = a_hi + CHUNK;
          const uint32_t d_tmem = tmem_base + s * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t off = kk * 32;
            uint64_t ah = make_sdesc_sw128(a_hi + off, 16, 1024), al = make_sdesc_sw128(a_lo + off, 16, 1024);
            uint64_t bh = make_sdesc_sw128(w_hi + c * CHUNK + off, 16, 1024);
            uint64_t bl = make_sdesc_sw128(w_lo + c * CHUNK + off, 16, 1024);
            umma_f16(d_tmem, al, bh, idesc, (c > 0 || kk > 0) ? 1u : 0u);
            umma_f16(d_tmem, ah, bl, idesc, 1u);
            umma_f16(d_tmem, ah, bh, idesc, 1u);
          }
          umma_commit(&empty[stage]);
          if (++stage == F_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(tfull);
        umma_commit(wfree);
        tphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // 16 epilogue warps: 4 per TMEM lane quarter, each owning 8 of the 32 hidden units of a round (4 warps per
    // scheduler keep the SFU / conversion latency of the activation math hidden)
    const int ewarp = warp - 4, q = ewarp & 3, sub = ewarp >> 2;
    const int et = threadIdx.x - 128;        // 0..511 among the epilogue threads
    const int row = q * 32 + lane;           // point row inside the tile == TMEM lane
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    // staging address of this thread's 16-byte piece inside a 64-byte swizzled row
    const uint32_t sw = (uint32_t)((row >> 1) & 3);
    const uint32_t piece0 = (uint32_t)row * 64 + (((uint32_t)sub ^ sw) << 4);
    uint32_t tphase = 0;
    int cur_l = -1;
    float un[4] = {1.f, 1.f, 1.f, 1.f}, so[4] = {1.f, 1.f, 1.f, 1.f};   // accumulator -> z ; activation -> stored plane
    for (int t = t_begin; t < t_end; ++t) {
      const int l = t / args.m_tiles, mt = t % args.m_tiles;
      const int pt = mt * 128 + row;
      if (l != cur_l) {  // all epilogue threads are past the previous tile's last staging barrier
        if (et < 128) {
          bias_s[et] = args.bias[l * kHidden + et];
          if (kLast) w3_s[et] = args.W3[l * kHidden + et];
        }
        const float* pl = args.plan + (long)l * PL_STRIDE;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          un[s] = __ldg(pl + args.u_slot + s) * (1.f + kTruncPerMma * 24.f);   // K = 128: chains of 24 MMAs
          so[s] = __ldg(pl + args.sa_slot + (kLast ? 0 : s));
        }
        cur_l = l;
        named_bar_sync(1, F_EPI_WARPS * 32);
      }
      mbar_wait(tfull, tphase, 15);
      tphase ^= 1;
      tc_fence_after();
      if (et == 0) NSVD_TL(t - t_begin, 3);       // MMAs of the tile retired
      float u[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int r = 0; r < 4; ++r) {          // h-quarters of 32 hidden units; this warp takes 8 of them
        const int h0 = r * 32 + sub * 8;
        float z[4][8];
#pragma unroll
        for (int s = 0; s < 4; ++s) tmem_ld8(tl + s * 128 + h0, z[s]);
        tmem_ld_wait();
        if (r == 3) {                        // last TMEM read of this tile: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);
          if (et == 0) NSVD_TL(t - t_begin, 4);   // last TMEM read of the tile
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float zb = fmaf(z[0][i], un[0], bias_s[h0 + i]);
          act_streams(zb, z[1][i] * un[1], z[2][i] * un[2], z[3][i] * un[3], z[0][i], z[1][i], z[2][i], z[3][i]);
        }
        if (kLast) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float w = w3_s[h0 + i];
#pragma unroll
            for (int s = 0; s < 4; ++s) u[s] = fmaf(z[s][i], w, u[s]);
          }
        }
        // The last layer stages the value stream only: two box pairs alternate between rounds, so a round never
        // waits for its predecessor's bulk store (one barrier per round); the other layers need all 8 boxes per
        // round and first let the previous round's stores drain them.
        const int sb = kLast ? 2 * (r & 1) : 0;
        if (!kLast) {
          if (et == 0) tma_store_wait_read();
          named_bar_sync(2, F_EPI_WARPS * 32);
        }
#pragma unroll
        for (int s = 0; s < (kLast ? 1 : 4); ++s) {
          uint32_t h[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split2<PF_HH>(z[s][2 * i] * so[s], z[s][2 * i + 1] * so[s], h[i], lo[i]);
          uint8_t* bh = sO + (sb + 2 * s) * F_BOX;
          uint8_t* bl = bh + F_BOX;
          *reinterpret_cast<uint4*>(bh + piece0) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(bl + piece0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async_smem();
        // kLast: every store issued up to the previous round has read its boxes (so the pair the NEXT round writes,
        // last used two rounds ago, is free once all threads have passed the barrier below)
        if (kLast && et == 0) tma_store_wait_read();
        named_bar_sync(3, F_EPI_WARPS * 32);
        if (et == 0) {
          if (!kLast) {
#pragma unroll
            for (int s = 1; s < 4; ++s) {
              tma_store_3d(&tmOh, sO + (2 * s) * F_BOX, r * 32, mt * 128, l * 4 + s);
              tma_store_3d(&tmOl, sO + (2 * s + 1) * F_BOX, r * 32, mt * 128, l * 4 + s);
            }
          }
          tma_store_3d(&tmSh, sO + sb * F_BOX, r * 32, (int)args.p_off + mt * 128, l);
          tma_store_3d(&tmSl, sO + (sb + 1) * F_BOX, r * 32, (int)args.p_off + mt * 128, l);
          tma_store_commit();
        }
      }
      if (kLast) {
        if (sub != 0) {
#pragma unroll
          for (int s = 0; s < 4; ++s) ubuf[(row * 3 + sub - 1) * 4 + s] = u[s];
        }
        named_bar_sync(1, F_EPI_WARPS * 32);
        if (sub == 0 && pt < args.P) {
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int s = 0; s < 4; ++s) u[s] += ubuf[(row * 3 + j) * 4 + s];
          u[0] += __ldg(args.b3 + l);
          const long pg = args.p_off + pt;
          PointGeom g = point_geom(args.x[2 * pg], args.x[2 * pg + 1], args.pb);
          float f, tf;
          operator_epilogue(g, args.pb, args.pb.has_exp_mask != 0, args.pb.has_exp_mask ? args.mscales[l] : 1.f,
                            u[0], u[1], u[2], u[3], f, tf);
          args.F[pg * args.L + l] = f;
          args.TF[pg * args.L + l] = tf;
          args.U0[pg * args.L + l] = u[0];
        }
        named_bar_sync(1, F_EPI_WARPS * 32);  // ubuf (staging boxes 6-7) is free again
      }
      if (et == 0) NSVD_TL(t - t_begin, 5);       // epilogue of the tile done
    }
    if (et == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// S2-fwd fused: hidden layers 1 AND 2 (+ head + operator) in one persistent kernel.  The three derivative streams of
// a1 go through a private scratch of kHidGroup tile slots per CTA (TMA bulk stores, evict_last; 192 KB per slot) and
// come straight back as the A operand of layer 2; only the value streams (needed by the backward) and F / TF are meant
// to leave the chip: 48 KB per point and 16 copies algorithmic instead of 107 KB with one kernel per layer.  Measured
// (profiles/README.md): single-tile groups keep the 28 MB of scratch L2-resident (60 KB per point of DRAM traffic) but
// stall 4.7 us per tile on the store -> load round trip; two-tile groups (57 MB of scratch, 108 KB per point) put a whole
// item between a tile's stores and their re-load, keep the weights of a layer for two items and are 9-10 % faster: the
// default.  (Keeping a1 in shared memory is not possible: 4 streams x 128 units x 4 B = 2 KB per point next to the
// weight planes.)
//   work items of a CTA, in groups of kHidGroup tiles of its range:  A(t0) [A(t1)] B(t0) [B(t1)]   (A = layer 1, B = layer 2)
//   smem : W planes of the layer in use (64 KB, reloaded when (layer, copy) changes) | ring 3 x 32 KB | 64 KB staging
//   TMEM : 4 x 128 columns (one block per stream), one item at a time; the MMAs of an item are issued stream by stream
//          and commit sfull[s] per stream, so the stream-major epilogue starts on stream 0 a quarter into the MMA phase
//   adone[slot] : the bulk stores of A(slot) have completed (cp.async.bulk.wait_group 0 by the issuing thread) - the
//                 producer waits for it before loading B(slot)'s operands.
// ------------------------------------------------------------------------------------------
struct HidFwd12Maps {
  CUtensorMap a0h, a0l;   // a0 derivative streams   [H][P][4 L]            load box {64, 128}
  CUtensorMap v0h, v0l;   // saved a0 value stream   [H][B][L]              load box {64, 128}
  CUtensorMap w1h, w1l, w2h, w2l;
  CUtensorMap s1h, s1l;   // saved a1 value stream   [H][B][L]              store box {32, 128}, 64-byte swizzle
  CUtensorMap v1h, v1l;   // the same buffer                                 load box {64, 128}
  CUtensorMap s2h, s2l;   // saved a2 value stream                           store box {32, 128}
  CUtensorMap xsh, xsl;   // scratch  [H][128][grid x 2 slots x 4 streams]   store box {32, 128}
  CUtensorMap xlh, xll;   // the same buffer                                 load box {64, 128}
};
struct HidFwd12Args {
  int L, P, m_tiles;
  long Btot, p_off;
  const float *bias1, *bias2;        // b1, b2 (L,128)
  const float* plan;                 // operand plan
  const float* W3;                   // (L,128)
  const float* b3;                   // (L)
  const float* x;                    // (Btot,2)
  const float* mscales;              // (L) or null
  float *F, *TF, *U0;                // (Btot, L)
  nsvd_problem_t pb;
  int vmode;                         // finite-difference mode, values only.  1: the 4 stream slots are the 4 shifted point
                                     // sets; reads U0, writes TF, stores nothing else.  2: central pass, tile = 512 points,
                                     // slot s = its s-th 128-point block; saves a1 / a2, writes F and U0 (no derivative
                                     // streams are needed when TF comes from differences: a quarter of the exact pass)
};

#ifndef NSVD_HID_GROUP
#define NSVD_HID_GROUP 2
#endif
constexpr int kHidGroup = NSVD_HID_GROUP;   // tiles per group (1 or 2): the scratch holds kHidGroup tile slots per CTA

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__global__ void __launch_bounds__(hid::F_THREADS, 1)
hidden_fwd12_kernel(const __grid_constant__ HidFwd12Maps tm, const HidFwd12Args args) {
  using namespace hid;
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                       // [hi c0 | hi c1 | lo c0 | lo c1]
  uint8_t* sA = smem + 2 * PLANE;           // ring
  uint8_t* sO = sA + F_STAGES * F_STAGE_BYTES;  // staging boxes
  uint64_t* bars = (uint64_t*)(sO + F_STAGING);
  uint64_t* full = bars;                    // [3]
  uint64_t* empty = bars + 3;               // [3]
  uint64_t* wfull = bars + 6;
  uint64_t* wfree = bars + 7;
  uint64_t* tempty = bars + 8;
  uint64_t* adone = bars + 9;               // [2]
  uint64_t* sfull = bars + 11;              // [4] accumulator of stream s complete
  uint32_t* tmem_slot = (uint32_t*)(bars + 15);
  float* bias_s = (float*)(bars + 16);      // [128]
  float* w3_s = bias_s + 128;               // [128]
  float* ubuf = (float*)(sO + 6 * F_BOX);   // [128][3][4] head partial sums (layer 2: boxes 2..7 are unused)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = args.L * args.m_tiles;
  const int tpc = (T + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * tpc;
  const int t_end = (t_begin + tpc < T) ? t_begin + tpc : T;
  constexpr int G = kHidGroup;
  const int n_items = t_end > t_begin ? 2 * G * ((t_end - t_begin + G - 1) / G) : 0;
  // item i: group g = i / (2 G), layer = (i / G) & 1 (0: layer 1, 1: layer 2), slot = i % G, tile = t_begin + G g + slot

  if (warp == 0 && lane == 0) {
    const CUtensorMap* m = &tm.a0h;
    for (int i = 0; i < 18; ++i) tma_prefetch_desc(m + i);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < F_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(wfull, 1);
    mbar_init(wfree, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&sfull[i], 1);
    mbar_init(tempty, F_EPI_WARPS);
    mbar_init(&adone[0], 1);
    mbar_init(&adone[1], 1);
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
      int stage = 0, cur_key = -1, done = 0;   // done = items already issued (valid ones)
      uint32_t phase = 0;
      const uint64_t pol_keep = l2_policy_evict_last(), pol_once = l2_policy_evict_first();
      for (int i = 0; i < n_items; ++i) {
        const int g = i / (2 * G), layer = (i / G) & 1, slot = i % G, t = t_begin + G * g + slot;
        if (t >= t_end) continue;
        const int l = t / args.m_tiles, mt = t % args.m_tiles;
        const int key = l * 2 + layer;
        if (key != cur_key) {
          if (done > 0) mbar_wait(wfree, (uint32_t)((done - 1) & 1), 60);   // MMAs of the previous item retired
          mbar_arrive_expect_tx(wfull, 2 * PLANE);
          const CUtensorMap* wh = layer ? &tm.w2h : &tm.w1h;
          const CUtensorMap* wl = layer ? &tm.w2l : &tm.w1l;
          tma_load_3d(sW, wh, wfull, 0, 0, l);
          tma_load_3d(sW + CHUNK, wh, wfull, 64, 0, l);
          tma_load_3d(sW + PLANE, wl, wfull, 0, 0, l);
          tma_load_3d(sW + PLANE + CHUNK, wl, wfull, 64, 0, l);
          cur_key = key;
        }
        if (layer) {   // layer-2 operands of this slot were written by this CTA: wait until those stores have completed
          mbar_wait(&adone[slot], (uint32_t)(g & 1), 61);
          fence_proxy_async_all();
        }
        const int xs = ((int)blockIdx.x * G + slot) * 4;
        const bool vm = args.vmode != 0, cen = args.vmode == 2;
        if (!cen) {
          // L2 prefetch of the layer-1 operands (256 KB from DRAM) of the tile that starts one or two items later:
          // a layer-2 item announces the same slot of the next group, A(t0) of a two-tile group announces t1
          const int tn = layer ? t + G : ((G == 2 && slot == 0) ? t + 1 : -1);
          if (tn >= 0 && tn < t_end) {
            const int ln = tn / args.m_tiles, mn = tn % args.m_tiles;
            for (int sc = 0; sc < 8; ++sc) {
              const int s = sc >> 1, c = sc & 1;
              if (s == 0 && !vm) {
                tma_prefetch_3d(&tm.v0h, 64 * c, (int)args.p_off + mn * 128, ln);
                tma_prefetch_3d(&tm.v0l, 64 * c, (int)args.p_off + mn * 128, ln);
              } else {
                tma_prefetch_3d(&tm.a0h, 64 * c, mn * 128, ln * 4 + s);
                tma_prefetch_3d(&tm.a0l, 64 * c, mn * 128, ln * 4 + s);
              }
            }
          }
        }
        for (int sc = 0; sc < 8; ++sc) {  // (stream, K-chunk)
          const int s = sc >> 1, c = sc & 1;
          mbar_wait(&empty[stage], phase ^ 1, 62);
          uint8_t* d = sA + stage * F_STAGE_BYTES;
          mbar_arrive_expect_tx(&full[stage], F_STAGE_BYTES);
          // operands read once (a0 from DRAM; a1 / the scratch for the last time) leave the L2 first
          if (cen) {             // central value-only pass: slot s = the s-th 128-point block of this 512-point tile
            const int row = (int)args.p_off + (mt * 4 + s) * 128;
            tma_load_3d_hint(d, layer ? &tm.v1h : &tm.v0h, &full[stage], 64 * c, row, l, pol_once);
            tma_load_3d_hint(d + CHUNK, layer ? &tm.v1l : &tm.v0l, &full[stage], 64 * c, row, l, pol_once);
          } else if (s == 0 && !vm) {   // value stream: from `saved` (whole-batch rows)
            tma_load_3d_hint(d, layer ? &tm.v1h : &tm.v0h, &full[stage], 64 * c, (int)args.p_off + mt * 128, l, pol_once);
            tma_load_3d_hint(d + CHUNK, layer ? &tm.v1l : &tm.v0l, &full[stage], 64 * c, (int)args.p_off + mt * 128, l,
                             pol_once);
          } else if (!layer) {
            tma_load_3d_hint(d, &tm.a0h, &full[stage], 64 * c, mt * 128, l * 4 + s, pol_once);
            tma_load_3d_hint(d + CHUNK, &tm.a0l, &full[stage], 64 * c, mt * 128, l * 4 + s, pol_once);
          } else {
            tma_load_3d_hint(d, &tm.xlh, &full[stage], 64 * c, 0, xs + s, pol_keep);
            tma_load_3d_hint(d + CHUNK, &tm.xll, &full[stage], 64 * c, 0, xs + s, pol_keep);
          }
          if (++stage == F_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        NSVD_TL(i, 6);   // all loads of the item issued
        ++done;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // (Issuing the unit dimension in two halves of 64 columns, so that the epilogue of one half overlaps the MMAs of the
      // other, was measured: the A tiles then pass through the ring twice and the kernel, which is bound by the ~42 B/clk
      // per SM of L2 -> SM fill bandwidth, gets slower - 4.5 ms instead of 3.5 ms per 131072 points.)
      constexpr uint32_t idesc = make_idesc_f16(128, 128, 0, 0, false, false);   // fp16 x fp16 planes
      int stage = 0, cur_key = -1;
      uint32_t phase = 0, wphase = 0, tphase = 0;
      const uint32_t w_hi = smem_u32(sW), w_lo = w_hi + PLANE;
      for (int i = 0; i < n_items; ++i) {
        const int g = i / (2 * G), layer = (i / G) & 1, slot = i % G, t = t_begin + G * g + slot;
        if (t >= t_end) continue;
        const int key = (t / args.m_tiles) * 2 + layer;
        if (key != cur_key) {
          mbar_wait(wfull, wphase, 63);
          wphase ^= 1;
          cur_key = key;
        }
        NSVD_TL(i, 7);   // weights of the item present
        mbar_wait(tempty, tphase ^ 1, 64);
        tc_fence_after();
        NSVD_TL(i, 0);   // accumulators handed back
        for (int sc = 0; sc < 8; ++sc) {
          const int s = sc >> 1, c = sc & 1;
          mbar_wait(&full[stage], phase, 65);
          tc_fence_after();
          if (sc == 0) NSVD_TL(i, 1);   // first operand stage present
          if (sc == 7) NSVD_TL(i, 2);   // last operand stage present
          const uint32_t a_hi = smem_u32(sA + stage * F_STAGE_BYTES), a_lo = a_hi + CHUNK;
          const uint32_t d_tmem = tmem_base + s * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t off = kk * 32;
            uint64_t ah = make_sdesc_sw128(a_hi + off, 16, 1024), al = make_sdesc_sw128(a_lo + off, 16, 1024);
            uint64_t bh = make_sdesc_sw128(w_hi + c * CHUNK + off, 16, 1024);
            uint64_t bl = make_sdesc_sw128(w_lo + c * CHUNK + off, 16, 1024);
            umma_f16(d_tmem, al, bh, idesc, (c > 0 || kk > 0) ? 1u : 0u);
            umma_f16(d_tmem, ah, bl, idesc, 1u);
            umma_f16(d_tmem, ah, bh, idesc, 1u);
          }
          umma_commit(&empty[stage]);
          if (c == 1) umma_commit(&sfull[s]);   // the accumulator of stream s is complete: its epilogue round may start
          if (++stage == F_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(wfree);
        tphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== 16 epilogue warps, STREAM-major =====================
    // The MMAs of an item are issued stream by stream, so the accumulator of stream s is complete a quarter / half /
    // three quarters of the way through the MMA phase (sfull[s]).  Round s of the epilogue handles stream s for the whole
    // tile as soon as it is there: rounds 0-2 run UNDER the MMAs of the later streams instead of after them (the
    // unit-major epilogue this replaces needed all four accumulators before its first round: MMA phase and epilogue ran
    // back to back, profiles/README.md timeline).  What the later streams need from the earlier ones stays in TMEM:
    // round 0 writes sigma = sigmoid(z0) over z0 (tcgen05.st), rounds 1-3 read it back, round 3 also re-reads z1, z2.
    // Thread = (point row, 32 consecutive units): the same thread owns the same TMEM cells in every round.
    const int ewarp = warp - 4, q = ewarp & 3, sub = ewarp >> 2;
    const int et = threadIdx.x - 128;        // 0..511 among the epilogue threads
    const int row = q * 32 + lane;           // point row inside the tile == TMEM lane
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * 32);   // + 128 s + 8 k
    const uint32_t sw = (uint32_t)((row >> 1) & 3);
    uint8_t* const my_hi = sO + (2 * sub) * F_BOX + row * 64;   // this thread's 64-byte row in the box of its unit quarter
    uint8_t* const my_lo = my_hi + F_BOX;
    uint32_t tphase = 0;
    int cur_key = -1;
    float un[4] = {1.f, 1.f, 1.f, 1.f}, so[4] = {1.f, 1.f, 1.f, 1.f};
    const bool vm = args.vmode != 0, cen = args.vmode == 2;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_once = l2_policy_evict_first();
    for (int i = 0; i < n_items; ++i) {
      const int g = i / (2 * G), layer = (i / G) & 1, slot = i % G, t = t_begin + G * g + slot;
      if (t >= t_end) continue;
      (void)g;
      const bool last = layer != 0;
      const int l = t / args.m_tiles, mt = t % args.m_tiles;
      const int pt = mt * 128 + row;
      const int key = l * 2 + layer;
      // is this the last layer-1 item of its group (the next valid item is a layer-2 one)?
      const bool closes_a = !last && (slot == G - 1 || t + 1 >= t_end);
      if (key != cur_key) {  // all epilogue threads are past the previous item's last staging barrier
        if (et < 128) {
          bias_s[et] = (last ? args.bias2 : args.bias1)[l * kHidden + et];
          w3_s[et] = args.W3[l * kHidden + et];
        }
        const float* pl = args.plan + (long)l * PL_STRIDE;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int ss = vm ? 0 : s;             // value-only pass: every slot carries a value stream
          un[s] = __ldg(pl + (last ? PL_U2 : PL_U1) + ss) * (1.f + kTruncPerMma * 24.f);   // K = 128: chains of 24 MMAs
          so[s] = last ? __ldg(pl + PL_SA2) : __ldg(pl + PL_SA1 + ss);
        }
        cur_key = key;
        named_bar_sync(1, F_EPI_WARPS * 32);
      }
      float u[4] = {0.f, 0.f, 0.f, 0.f};
      const int xs = ((int)blockIdx.x * G + slot) * 4;
      const bool wide = !last || cen;          // every stream of this item is staged and stored
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        mbar_wait(&sfull[s], tphase, 66);
        tc_fence_after();
        if (s == 0 && et == 0) NSVD_TL(i, 3);  // first accumulator of the item complete
        const bool stores = wide || (s == 0 && !vm);
        if (stores) {      // the staging boxes must have been read by the bulk stores that used them last
          if (et == 0) {
            if (!last && s == 0 && slot == 1) {   // A(t0)'s stores were issued a whole item ago: they are complete
              tma_store_wait_all();
              mbar_arrive(&adone[0]);
            } else {
              tma_store_wait_read();
            }
          }
          named_bar_sync(2, F_EPI_WARPS * 32);
        }
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {          // 8 of this thread's 32 units
          const int h0 = sub * 32 + 8 * k;
          float a[8];
          if (vm) {                            // value-only pass: the four slots are independent value streams
            tmem_ld8(tl + s * 128 + 8 * k, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = softplus_fast(fmaf(a[j], un[s], bias_s[h0 + j]));
          } else if (s == 0) {
            float sg[8];
            tmem_ld8(tl + 8 * k, a);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) softplus_sigmoid(fmaf(a[j], un[0], bias_s[h0 + j]), a[j], sg[j]);
            tmem_st8(tl + 8 * k, sg);          // sigma replaces z0 for rounds 1-3
          } else if (s < 3) {
            float sg[8];
            tmem_ld8(tl + s * 128 + 8 * k, a);
            tmem_ld8(tl + 8 * k, sg);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = sg[j] * (a[j] * un[s]);
          } else {
            float sg[8], z1[8], z2[8];
            tmem_ld8(tl + 3 * 128 + 8 * k, a);
            tmem_ld8(tl + 8 * k, sg);
            tmem_ld8(tl + 128 + 8 * k, z1);
            tmem_ld8(tl + 256 + 8 * k, z2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float d1 = z1[j] * un[1], d2 = z2[j] * un[2];
              a[j] = fmaf(sg[j], a[j] * un[3], sg[j] * (1.f - sg[j]) * fmaf(d1, d1, d2 * d2));
            }
          }
          if (last) {
#pragma unroll
            for (int j = 0; j < 8; ++j) u[s] = fmaf(a[j], w3_s[h0 + j], u[s]);
          }
          if (stores) {
            uint32_t h[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split2<PF_HH>(a[2 * j] * so[s], a[2 * j + 1] * so[s], h[j], lo[j]);
            const uint32_t piece = ((uint32_t)k ^ sw) << 4;
            *reinterpret_cast<uint4*>(my_hi + piece) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(my_lo + piece) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        if (s == 0 && !vm) tmem_st_wait();     // sigma is in TMEM before this thread reads it back
        if (s == 3) {                          // last TMEM access of this item: hand the accumulators back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);
          if (et == 0) NSVD_TL(i, 4);
        }
        if (stores) {
          fence_proxy_async_smem();
          named_bar_sync(3, F_EPI_WARPS * 32);
          if (et == 0) {   // boxes 2 j (hi), 2 j + 1 (lo): units 32 j .. 32 j + 31 of stream s
            const CUtensorMap *mh, *ml;
            int y, z;
            uint64_t pol;
            if (cen) {           // slot s = the s-th 128-point block of the tile: a1 (re-read by layer 2 and the backward) / a2
              mh = last ? &tm.s2h : &tm.s1h;
              ml = last ? &tm.s2l : &tm.s1l;
              y = (int)args.p_off + (mt * 4 + s) * 128;
              z = l;
              pol = last ? pol_once : pol_keep;
            } else if (last) {   // a2 value stream for the backward
              mh = &tm.s2h;
              ml = &tm.s2l;
              y = (int)args.p_off + mt * 128;
              z = l;
              pol = pol_once;
            } else if (s == 0 && !vm) {   // a1 value stream: read back by layer 2 of this tile, then by the backward
              mh = &tm.s1h;
              ml = &tm.s1l;
              y = (int)args.p_off + mt * 128;
              z = l;
              pol = pol_keep;
            } else {             // derivative streams of a1 (shifted value streams in finite-difference mode): scratch
              mh = &tm.xsh;
              ml = &tm.xsl;
              y = 0;
              z = xs + s;
              pol = pol_keep;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              tma_store_3d_hint(mh, sO + (2 * j) * F_BOX, 32 * j, y, z, pol);
              tma_store_3d_hint(ml, sO + (2 * j + 1) * F_BOX, 32 * j, y, z, pol);
            }
            tma_store_commit();
          }
        }
      }
      tphase ^= 1;
      if (et == 0) NSVD_TL(i, 5);           // epilogue of the item done (stores issued)
      if (closes_a && et == 0) {             // every store of this group's layer-1 items has completed
        tma_store_wait_all();
        if (slot == 0) mbar_arrive(&adone[0]);
        else mbar_arrive(&adone[1]);
      }
      if (last) {
        // the head partial sums go through staging boxes 6-7: the stores of this item must have read them
        if (et == 0) tma_store_wait_read();
        named_bar_sync(2, F_EPI_WARPS * 32);
        if (sub != 0) {
#pragma unroll
          for (int s = 0; s < 4; ++s) ubuf[(row * 3 + sub - 1) * 4 + s] = u[s];
        }
        named_bar_sync(1, F_EPI_WARPS * 32);
        if (sub == 0 && (pt < args.P || cen)) {
#pragma unroll
          for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int s = 0; s < 4; ++s) u[s] += ubuf[(row * 3 + j) * 4 + s];
          const long pg = args.p_off + pt;
          const float msc = args.pb.has_exp_mask ? args.mscales[l] : 1.f;
          if (cen) {       // values only: F and U0 of the four blocks (TF comes from the shifted pass)
            const float b3v = __ldg(args.b3 + l);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              const int pts = (mt * 4 + s) * 128 + row;
              if (pts < args.P) {
                const long pgs = args.p_off + pts;
                const float u0 = u[s] + b3v;
                PointGeom gm = point_geom(args.x[2 * pgs], args.x[2 * pgs + 1], args.pb);
                const float mexp = args.pb.has_exp_mask ? expf(-gm.r / msc) : 1.f;
                args.F[pgs * args.L + l] = head_factor(gm, args.pb, mexp) * u0;
                args.U0[pgs * args.L + l] = u0;
              }
            }
          } else if (vm) {        // finite differences of the four shifted values around the central one (first pass)
            const float b3v = __ldg(args.b3 + l);
#pragma unroll
            for (int s = 0; s < 4; ++s) u[s] += b3v;
            args.TF[pg * args.L + l] = fd_operator(args.x[2 * pg], args.x[2 * pg + 1], args.pb,
                                                   args.pb.has_exp_mask != 0, msc, args.U0[pg * args.L + l], u);
          } else {
            u[0] += __ldg(args.b3 + l);
            PointGeom gm = point_geom(args.x[2 * pg], args.x[2 * pg + 1], args.pb);
            float f, tf;
            operator_epilogue(gm, args.pb, args.pb.has_exp_mask != 0, msc, u[0], u[1], u[2], u[3], f, tf);
            args.F[pg * args.L + l] = f;
            args.TF[pg * args.L + l] = tf;
            args.U0[pg * args.L + l] = u[0];
          }
        }
        named_bar_sync(1, F_EPI_WARPS * 32);  // ubuf (staging boxes 6-7) is free again
      }
    }
    if (et == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

struct HidBwdArgs {
  int L, P, m_tiles;
  long Btot, p_off;
  float* dW;                                 // (L,128,128) accumulated with reductions
  float* db_prev;                            // (L,128) accumulated with atomics
  const float* plan;                         // operand plan
  int ud_slot, inv_sa_slot, sdz_slot, uw_slot;   // dgrad accumulator -> dA ; saved a_{i-1} plane -> a ; scale of the
                                                 // dZ_{i-1} planes written ; wgrad accumulator -> dW_i
};
constexpr int kWgradFlush = 4;   // half tiles (64 points) per TMEM accumulation chain of the hidden-layer weight gradient

// ------------------------------------------------------------------------------------------
// S2-bwd: backward of hidden layer i for one copy.  One pass over dZ_i produces BOTH the input gradient
//   dgrad  D1 = dZ_i W_i  (then dZ_{i-1} = D1 (.) sigma(a_{i-1}))  and the weight gradient  wgrad  D2 = dZ_i^T a_{i-1},
// on HALF tiles of 64 points in a 2-stage ring, so the loads of half tile j+1 overlap the MMAs, epilogue and stores
// of half tile j (an earlier version kept three 64 KB tiles resident and ran load -> MMA -> epilogue back to back).
//   smem : W_i^T hi/lo resident (64 KB) | 2 stages x 64 KB (dZ_i hi/lo, a_{i-1} hi/lo of 64 points).
//   dgrad is issued transposed, D1[k][p] = sum_j W_i^T[k][j] dZ_i[p][j]  (M = 128 units, N = 64 points), so the
//   accumulator keeps the full 128-lane datapath busy; TMEM lane = unit k, column = point.  Epilogue thread (k, 16
//   points): dZ_{i-1} = D1 (.) sigma(a_{i-1}) with a read from the stage, written back as hi/lo over the dead dZ_i
//   half tile (2-byte accesses, conflict-free: a warp covers 64 contiguous bytes of one row), db_{i-1} is a
//   per-thread running sum.  Warp 3 drains the staged half tile with TMA stores and releases the stage.
//   wgrad D2[j][k] += dZ_i^T a_{i-1} accumulates in TMEM over the half tiles of a copy (K = 64 points each).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(hid::F_THREADS, 1)
hidden_bwd2_kernel(const __grid_constant__ CUtensorMap tmZh, const __grid_constant__ CUtensorMap tmZl,
                   const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                   const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWl,
                   const __grid_constant__ CUtensorMap tmOh, const __grid_constant__ CUtensorMap tmOl,
                   const HidBwdArgs args) {
  using namespace hid;
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                 // W_i^T hi | lo  (128 rows k, K = j)
  uint8_t* sS = smem + 2 * PLANE;     // 2 stages
  uint8_t* sOut = sS + 2 * B2_STAGE;  // output staging: dZ_{i-1} half tile hi | lo (32 KB), drained by the store warp
  uint64_t* bars = (uint64_t*)(sOut + 2 * HPLANE);
  uint64_t* full = bars;              // [2] stage landed
  uint64_t* mma_done = bars + 2;      // [2] dgrad + wgrad of the stage retired
  uint64_t* empty = bars + 4;         // [2] the 16 epilogue warps are done with the stage and its TMEM buffers
  uint64_t* wfull = bars + 6;
  uint32_t* tmem_slot = (uint32_t*)(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h_tiles = args.m_tiles;   // half tiles (64 points) per copy
  const int T = args.L * h_tiles;
  const int tpc = (T + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * tpc;
  const int t_end = (t_begin + tpc < T) ? t_begin + tpc : T;
  constexpr int NBAR = F_EPI_WARPS * 32 + 32;   // epilogue threads arrive, the store warp syncs

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmZh);
    tma_prefetch_desc(&tmZl);
    tma_prefetch_desc(&tmAh);
    tma_prefetch_desc(&tmAl);
    tma_prefetch_desc(&tmWh);
    tma_prefetch_desc(&tmWl);
    tma_prefetch_desc(&tmOh);
    tma_prefetch_desc(&tmOl);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&mma_done[i], 1);
      mbar_init(&empty[i], F_EPI_WARPS);
    }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // columns [0,64) and [64,128): dgrad buffers; [128,256) and [256,384): wgrad buffers.  The wgrad chain is restarted
  // every kWgradFlush half tiles in the other buffer (TMEM accumulation truncates: 48 chained MMAs = 9e-7) and the
  // finished chain is added to register accumulators by the epilogue threads.  Re-using a wgrad buffer two chains later
  // needs no barrier of its own: the MMAs of half tile j wait for full[j & 1], which is loaded only after the epilogue
  // of half tile j - 2 (including its drain) has released the stage.
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d2 = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int cur_l = -1;
      for (int t = t_begin; t < t_end; ++t) {
        const int j = t - t_begin, s = j & 1, u = j >> 1;
        const int l = t / h_tiles, ht = t % h_tiles;
        if (u > 0) mbar_wait(&empty[s], (uint32_t)((u - 1) & 1), 40);
        if (l != cur_l) {
          // every MMA that reads the resident weights has retired once the previous half tile is drained
          if (j > 0) mbar_wait(&empty[(j - 1) & 1], (uint32_t)(((j - 1) >> 1) & 1), 41);
          mbar_arrive_expect_tx(wfull, 2 * PLANE);
          tma_load_3d(sW, &tmWh, wfull, 0, 0, l);
          tma_load_3d(sW + CHUNK, &tmWh, wfull, 64, 0, l);
          tma_load_3d(sW + PLANE, &tmWl, wfull, 0, 0, l);
          tma_load_3d(sW + PLANE + CHUNK, &tmWl, wfull, 64, 0, l);
          cur_l = l;
        }
        uint8_t* st = sS + s * B2_STAGE;
        const int pz = ht * HROWS, pa = (int)args.p_off + pz;
        mbar_arrive_expect_tx(&full[s], B2_STAGE);
        tma_load_3d(st, &tmZh, &full[s], 0, pz, l);
        tma_load_3d(st + HCHUNK, &tmZh, &full[s], 64, pz, l);
        tma_load_3d(st + HPLANE, &tmZl, &full[s], 0, pz, l);
        tma_load_3d(st + HPLANE + HCHUNK, &tmZl, &full[s], 64, pz, l);
        tma_load_3d(st + 2 * HPLANE, &tmAh, &full[s], 0, pa, l);
        tma_load_3d(st + 2 * HPLANE + HCHUNK, &tmAh, &full[s], 64, pa, l);
        tma_load_3d(st + 3 * HPLANE, &tmAl, &full[s], 0, pa, l);
        tma_load_3d(st + 3 * HPLANE + HCHUNK, &tmAl, &full[s], 64, pa, l);
        NSVD_TL(j, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_d = make_idesc_f16(128, HROWS, 0, 0, false, false);  // dgrad: K-major, M = units, N = points
      constexpr uint32_t idesc_w = make_idesc_f16(128, 128, 1, 1, false, false);    // wgrad: MN-major operands
      const uint32_t w_hi = smem_u32(sW), w_lo = w_hi + PLANE;
      int cur_l = -1, hrun = 0, cidx = -1;
      uint32_t wphase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int j = t - t_begin, s = j & 1, u = j >> 1;
        const int l = t / h_tiles;
        if (l != cur_l) {
          mbar_wait(wfull, wphase, 42);
          wphase ^= 1;
          cur_l = l;
          hrun = 0;
        }
        const bool first_of_chain = (hrun % kWgradFlush) == 0;
        if (first_of_chain) ++cidx;
        const uint32_t tmem_w = tmem_d2 + (uint32_t)(cidx & 1) * 128;
        ++hrun;
        mbar_wait(&full[s], (uint32_t)(u & 1), 43);   // implies stage s and TMEM buffer s were released
        tc_fence_after();
        NSVD_TL(j, 1);
        const uint32_t z_hi = smem_u32(sS + s * B2_STAGE), z_lo = z_hi + HPLANE;
        const uint32_t a_hi = z_hi + 2 * HPLANE, a_lo = z_hi + 3 * HPLANE;
        const uint32_t acc = tmem_base + (uint32_t)s * HROWS;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t ow = (kk >> 2) * CHUNK + (kk & 3) * 32, oz = (kk >> 2) * HCHUNK + (kk & 3) * 32;
          uint64_t ah = make_sdesc_sw128(w_hi + ow, 16, 1024), al = make_sdesc_sw128(w_lo + ow, 16, 1024);
          uint64_t bh = make_sdesc_sw128(z_hi + oz, 16, 1024), bl = make_sdesc_sw128(z_lo + oz, 16, 1024);
          umma_f16(acc, al, bh, idesc_d, kk > 0 ? 1u : 0u);
          umma_f16(acc, ah, bl, idesc_d, 1u);
          umma_f16(acc, ah, bh, idesc_d, 1u);
        }
#pragma unroll
        for (int kk = 0; kk < HROWS / 16; ++kk) {   // K = 64 points, 16 per MMA
          const uint32_t off = kk * 2048;
          uint64_t ah = make_sdesc_sw128(z_hi + off, HCHUNK, 1024), al = make_sdesc_sw128(z_lo + off, HCHUNK, 1024);
          uint64_t bh = make_sdesc_sw128(a_hi + off, HCHUNK, 1024), bl = make_sdesc_sw128(a_lo + off, HCHUNK, 1024);
          umma_f16(tmem_w, al, bh, idesc_w, (first_of_chain && kk == 0) ? 0u : 1u);
          umma_f16(tmem_w, ah, bl, idesc_w, 1u);
          umma_f16(tmem_w, ah, bh, idesc_w, 1u);
        }
        umma_commit(&mma_done[s]);
        NSVD_TL(j, 2);
      }
    }
  } else if (warp == 3) {
    // ===================== store warp: drain the staged half tile, release the stage =====================
    for (int t = t_begin; t < t_end; ++t) {
      const int j = t - t_begin, s = j & 1;
      const int l = t / h_tiles, ht = t % h_tiles;
      named_bar_sync(2 + s, NBAR);
      if (lane == 0) {
        tma_store_3d(&tmOh, sOut, 0, ht * HROWS, l);
        tma_store_3d(&tmOh, sOut + HCHUNK, 64, ht * HROWS, l);
        tma_store_3d(&tmOl, sOut + HPLANE, 0, ht * HROWS, l);
        tma_store_3d(&tmOl, sOut + HPLANE + HCHUNK, 64, ht * HROWS, l);
        tma_store_commit();
        tma_store_wait_read();
        NSVD_TL(j, 5);
      }
      __syncwarp();
      if (t + 1 < t_end) named_bar_arrive(4, NBAR);   // the staging buffer may be written again
    }
    if (lane == 0) tma_store_wait_all();
  } else if (warp >= 4) {
    // ===================== 16 epilogue warps: lane quarter q (32 units) x point group cg (16 points) ==========
    const int ewarp = warp - 4, q = ewarp & 3, cg = ewarp >> 2;
    const int k = q * 32 + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t koff = (uint32_t)((k >> 6) * HCHUNK + (k & 7) * 2);
    const int piece = (k & 63) >> 3;
    float dbacc = 0.f;
    float wacc[32];                 // dW_i[row k][cg*32 .. cg*32+31] of the current copy, summed over finished chains
#pragma unroll
    for (int i = 0; i < 32; ++i) wacc[i] = 0.f;
    int cur_l = -1, hrun = 0, cidx = -1;
    float ud = 1.f, inv_sa = 1.f, sdz = 1.f, uw = 1.f;
    for (int t = t_begin; t < t_end; ++t) {
      const int j = t - t_begin, s = j & 1, u = j >> 1;
      const int l = t / h_tiles;
      if (l != cur_l) {
        const float* pl = args.plan + (long)l * PL_STRIDE;
        ud = __ldg(pl + args.ud_slot) * (1.f + kTruncPerMma * 24.f);   // dgrad: K = 128, chains of 24 MMAs
        inv_sa = __ldg(pl + args.inv_sa_slot);
        sdz = __ldg(pl + args.sdz_slot);
        uw = __ldg(pl + args.uw_slot);
        cur_l = l;
        hrun = 0;
      }
      if ((hrun % kWgradFlush) == 0) ++cidx;
      ++hrun;
      mbar_wait(&mma_done[s], (uint32_t)(u & 1), 45);
      tc_fence_after();
      if (warp == 4 && lane == 0) NSVD_TL(j, 3);
      float v[16];
      tmem_ld16(tl + (uint32_t)(s * HROWS + cg * 16), v);
      const uint8_t* st = sS + s * B2_STAGE;
      tmem_ld_wait();
      uint32_t packed[16];                // hi | lo << 16 of this thread's 16 outputs
      // rows beyond P carry dZ_i = 0 (TMA zero fill), hence D1 = 0 and the product is 0 without a mask
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int pt = cg * 16 + i;
        const uint32_t off = koff + (uint32_t)(pt * 128 + ((piece ^ (pt & 7)) << 4));
        const uint32_t ahb = *reinterpret_cast<const uint16_t*>(st + 2 * HPLANE + off);
        const uint32_t alb = *reinterpret_cast<const uint16_t*>(st + 3 * HPLANE + off);
        const float a = merge1h(ahb, alb) * inv_sa;
        const float val = v[i] * ud * sig_fast(a);
        dbacc += val;
        uint16_t h16, l16;
        split1<PF_HH>(val * sdz, h16, l16);
        packed[i] = (uint32_t)h16 | ((uint32_t)l16 << 16);
      }
      const bool last_of_run = (t + 1 == t_end) || ((t + 1) / h_tiles != l);
      if (last_of_run || (hrun % kWgradFlush) == 0) {       // the wgrad chain in buffer cidx & 1 is complete
        const uint32_t tw = tl + 128 + (uint32_t)(cidx & 1) * 128 + (uint32_t)(cg * 32);
        const int nh = ((hrun - 1) % kWgradFlush) + 1;                  // half tiles in this chain, 12 MMAs each
        const float corr = 1.f + kTruncPerMma * (float)(12 * nh);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          float w[16];
          tmem_ld16(tw + ch * 16, w);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) wacc[ch * 16 + i] = fmaf(w[i], corr, wacc[ch * 16 + i]);
        }
      }
      if (last_of_run) {
        float* drow = args.dW + ((long)l * kHidden + k) * kHidden + cg * 32;   // lane = row j of dW
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          red_add_v4(drow + i, wacc[i] * uw, wacc[i + 1] * uw, wacc[i + 2] * uw, wacc[i + 3] * uw);
          wacc[i] = wacc[i + 1] = wacc[i + 2] = wacc[i + 3] = 0.f;
        }
        atomicAdd(args.db_prev + l * kHidden + k, dbacc);
        dbacc = 0.f;
      }
      // the stage (dZ_i, a_{i-1}) and both TMEM buffers of this half tile are no longer needed: hand them back now - the
      // producer reloads the stage while the outputs are staged and stored from their own buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (j > 0) named_bar_sync(4, NBAR);       // the store of the previous half tile has read the staging buffer
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int pt = cg * 16 + i;
        const uint32_t off = koff + (uint32_t)(pt * 128 + ((piece ^ (pt & 7)) << 4));
        *reinterpret_cast<uint16_t*>(sOut + off) = (uint16_t)(packed[i] & 0xffffu);
        *reinterpret_cast<uint16_t*>(sOut + HPLANE + off) = (uint16_t)(packed[i] >> 16);
      }
      fence_proxy_async_smem();
      if (warp == 4 && lane == 0) NSVD_TL(j, 4);
      named_bar_arrive(2 + s, NBAR);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// head backward (SIMT, HBM-bound): du = dF c m rho ; dZ2 = du W3 (.) sigma(a2) -> hi/lo planes ;
// dW3, db3, db2, dscales accumulated (block partials + atomics).  grid = (point chunks, L)
// ------------------------------------------------------------------------------------------
constexpr int kHeadChunk = 128;  // points per block (32 per warp); small batches use 32 (8 per warp) for more blocks
__global__ void __launch_bounds__(128)
head_bwd_bf16_kernel(const float* __restrict__ dF, const float* __restrict__ U0,
                     const __nv_bfloat16* __restrict__ a2_hi, const __nv_bfloat16* __restrict__ a2_lo,
                     const float* __restrict__ W3, const float* __restrict__ x, const float* __restrict__ mscales,
                     float* __restrict__ plan, const float* __restrict__ hstat, const float* __restrict__ mdf,
                     nsvd_problem_t pb, __nv_bfloat16* __restrict__ dz_hi, __nv_bfloat16* __restrict__ dz_lo,
                     float* __restrict__ dW3, float* __restrict__ db3, float* __restrict__ db2,
                     float* __restrict__ dscales, int P, long Btot, long p_off, int chunk) {
  __shared__ float red[4][2 * kHidden + 2];
  const int l = blockIdx.y, L = pb.n_copies;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p_begin = blockIdx.x * chunk, wpts = chunk >> 2;   // points per warp
  const int p_end = p_begin + chunk < P ? p_begin + chunk : P;
  float w3[4], accW[4] = {0, 0, 0, 0}, accB[4] = {0, 0, 0, 0}, acc_b3 = 0.f, acc_s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) w3[i] = W3[l * kHidden + lane * 4 + i];
  float sc = pb.has_exp_mask ? mscales[l] : 1.f;
  const float inv_sa2 = plan[(long)l * PL_STRIDE + PL_INV_SA2];
  float m3 = fmaxf(fmaxf(fabsf(w3[0]), fabsf(w3[1])), fmaxf(fabsf(w3[2]), fabsf(w3[3])));
  for (int o = 16; o > 0; o >>= 1) m3 = fmaxf(m3, __shfl_xor_sync(0xffffffffu, m3, o));
  const BwdScales bsc = bwd_scales(hstat, mdf[l], m3, pb.hard_mul_const, l, L);
  if (blockIdx.x == 0 && threadIdx.x == 0) publish_bwd_plan(plan + (long)l * PL_STRIDE, bsc);
  const float sdz2 = bsc.s2;
  // each warp owns chunk / 4 consecutive points: lane i evaluates the per-point factor du of point i once, then the
  // warp walks its points with 4 rows of a2 in flight (lane = 4 hidden units)
  {
    const int pw = p_begin + warp * wpts;
    const int pmine = pw + lane;
    float du_l = 0.f;
    if (lane < wpts && pmine < p_end) {
      const long pg = p_off + pmine;
      PointGeom g = point_geom(x[2 * pg], x[2 * pg + 1], pb);
      float m = pb.has_exp_mask ? expf(-g.r / sc) : 1.f;
      float cm = head_factor(g, pb, m);
      du_l = dF[pg * L + l] * cm;
      acc_b3 = du_l;
      if (pb.has_exp_mask) acc_s = du_l * U0[pg * L + l] * g.r / (sc * sc);
    }
    for (int j0 = 0; j0 < wpts && pw + j0 < p_end; j0 += 4) {
      uint2 h[4], lo2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = pw + j0 + u;
        const long oa = ((long)l * Btot + p_off + (p < p_end ? p : pw)) * kHidden + lane * 4;
        h[u] = *reinterpret_cast<const uint2*>(a2_hi + oa);
        lo2[u] = *reinterpret_cast<const uint2*>(a2_lo + oa);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = pw + j0 + u;
        const float du = __shfl_sync(0xffffffffu, du_l, j0 + u);
        if (p >= p_end) continue;
        float a[4];
        {
          float2 t0 = tc::merge2<tc::PF_HH>(h[u].x, lo2[u].x), t1 = tc::merge2<tc::PF_HH>(h[u].y, lo2[u].y);
          a[0] = t0.x * inv_sa2; a[1] = t0.y * inv_sa2; a[2] = t1.x * inv_sa2; a[3] = t1.y * inv_sa2;
        }
        float dz[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          dz[i] = du * w3[i] * sig_fast(a[i]);
          accW[i] = fmaf(du, a[i], accW[i]);
          accB[i] += dz[i];
        }
        uint32_t h01, l01, h23, l23;
        tc::split2<tc::PF_HH>(dz[0] * sdz2, dz[1] * sdz2, h01, l01);
        tc::split2<tc::PF_HH>(dz[2] * sdz2, dz[3] * sdz2, h23, l23);
        const long oz = ((long)l * P + p) * kHidden + lane * 4;
        *reinterpret_cast<uint2*>(dz_hi + oz) = make_uint2(h01, h23);
        *reinterpret_cast<uint2*>(dz_lo + oz) = make_uint2(l01, l23);
      }
    }
    acc_b3 = warp_sum(acc_b3);
    acc_s = warp_sum(acc_s);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[warp][lane * 4 + i] = accW[i];
    red[warp][kHidden + lane * 4 + i] = accB[i];
  }
  if (lane == 0) {
    red[warp][2 * kHidden] = acc_b3;
    red[warp][2 * kHidden + 1] = acc_s;
  }
  __syncthreads();
  const int tix = threadIdx.x;
  float sW = red[0][tix] + red[1][tix] + red[2][tix] + red[3][tix];
  float sB = red[0][kHidden + tix] + red[1][kHidden + tix] + red[2][kHidden + tix] + red[3][kHidden + tix];
  atomicAdd(dW3 + l * kHidden + tix, sW);
  atomicAdd(db2 + l * kHidden + tix, sB);
  if (tix == 0) {
    atomicAdd(db3 + l, red[0][2 * kHidden] + red[1][2 * kHidden] + red[2][2 * kHidden] + red[3][2 * kHidden]);
    if (pb.has_exp_mask && dscales)
      atomicAdd(dscales + l, red[0][2 * kHidden + 1] + red[1][2 * kHidden + 1] + red[2][2 * kHidden + 1] +
                                 red[3][2 * kHidden + 1]);
  }
}

// ------------------------------------------------------------------------------------------
// engine drivers
// ------------------------------------------------------------------------------------------
static int g_tc_micro_batch = 0;   // 0 = not initialised yet
void tc_set_micro_batch(int points) {
  if (points < 128) points = 128;
  g_tc_micro_batch = (points + 127) / 128 * 128;
}
static int tc_micro_batch() {
  if (!g_tc_micro_batch) {
    const char* e = getenv("NSVD_TC_MICROBATCH");
    tc_set_micro_batch(e ? atoi(e) : 65536);
  }
  return g_tc_micro_batch;
}

constexpr int kHidGrid = 148;   // CTAs of the hidden-layer kernels (one per SM)

struct TcLayout {
  // saved (whole batch)
  size_t phi_hi, phi_lo, av_hi[3], av_lo[3], u0, plan, rowstat, hstat, mdf, saved_total;
  // work
  size_t w0_hi, w0_lo, w_hi[2], w_lo[2], wT_hi[2], wT_lo[2], str_hi[2], str_lo[2], dz_hi[2], dz_lo[2], scr_hi, scr_lo, phis_hi, phis_lo,
      w0v_hi, w0v_lo, work_total;
  long P;
};
static TcLayout tc_layout(const nsvd_problem_t& pb) {
  TcLayout t{};
  const size_t B = pb.n_points, L = pb.n_copies, K0 = 2 * (size_t)pb.n_fourier, H = kHidden;
  size_t mb = (size_t)tc_micro_batch();
  t.P = (long)(B < mb ? B : mb);
  const size_t P = (size_t)t.P;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += align_up(bytes, 1024);
    return r;
  };
  t.phi_hi = take(B * K0 * 2);
  t.phi_lo = take(B * K0 * 2);
  for (int i = 0; i < 3; ++i) {
    t.av_hi[i] = take(L * B * H * 2);
    t.av_lo[i] = take(L * B * H * 2);
  }
  t.u0 = take(B * L * 4);
  t.plan = take(L * PL_STRIDE * 4);        // operand plan: written by the forward, completed and read by the backward
  t.rowstat = take(L * H * 5 * 4);
  t.hstat = take((2 * L * 3 + 4) * 4);   // + the three feature-stream bounds
  t.mdf = take(L * 4);
  t.saved_total = o + 1024;
  o = 0;
  t.w0_hi = take(L * 512 * K0 * 2);
  t.w0_lo = take(L * 512 * K0 * 2);
  for (int i = 0; i < 2; ++i) {
    t.w_hi[i] = take(L * H * H * 2);
    t.w_lo[i] = take(L * H * H * 2);
    t.wT_hi[i] = take(L * H * H * 2);
    t.wT_lo[i] = take(L * H * H * 2);
  }
  for (int i = 0; i < 2; ++i) {
    t.str_hi[i] = take(L * 4 * P * H * 2);
    t.str_lo[i] = take(L * 4 * P * H * 2);
  }
  // L2-resident scratch of the fused hidden forward: 148 CTAs x 2 tile slots x 3 derivative streams x (128 x 128) planes
  t.scr_hi = take((size_t)kHidGrid * 8 * hid::PLANE);
  t.scr_lo = take((size_t)kHidGrid * 8 * hid::PLANE);
  if (pb.fd_eps > 0.f) {   // finite-difference pass: features of the 4 shifted point sets, W0 planes (not folded)
    t.phis_hi = take(4 * P * K0 * 2);
    t.phis_lo = take(4 * P * K0 * 2);
    t.w0v_hi = take(L * H * K0 * 2);
    t.w0v_lo = take(L * H * K0 * 2);
  }
  // backward reuses the stream area for the dZ planes (two ping-pong pairs)
  for (int i = 0; i < 2; ++i) {
    t.dz_hi[i] = t.str_hi[i];
    t.dz_lo[i] = t.str_lo[i];
  }
  t.work_total = o + 1024;
  return t;
}

void tc_scratch_bytes(const nsvd_problem_t& pb, size_t* saved, size_t* work) {
  TcLayout t = tc_layout(pb);
  *saved = t.saved_total;
  *work = t.work_total;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static inline uint8_t* align1k(void* p) { return (uint8_t*)(((uintptr_t)p + 1023) & ~(uintptr_t)1023); }
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)

template <bool kLast>
static int launch_hidden_fwd(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh,
                             const CUtensorMap& wl, const CUtensorMap& oh, const CUtensorMap& ol,
                             const CUtensorMap& sh, const CUtensorMap& sl, const CUtensorMap& vh, const CUtensorMap& vl,
                             const HidFwdArgs& a, cudaStream_t st) {
  auto kern = hidden_fwd_kernel<kLast>;
  NSVD_SMEM_OPTIN(kern, hid::SMEM_FWD);
  int T = a.L * a.m_tiles;
  int grid = T < 148 ? T : 148;
  kern<<<grid, hid::F_THREADS, hid::SMEM_FWD, st>>>(ah, al, wh, wl, oh, ol, sh, sl, vh, vl, a);
  NSVD_LAUNCH_CHECK();
  return 0;
}

int tc_forward(const nsvd_problem_t& pb, const nsvd_params_t& pr, const float* x, float* F, float* TF, void* saved_v,
               void* work_v, size_t work_bytes, cudaStream_t st) {
  (void)work_bytes;
  const long B = pb.n_points, L = pb.n_copies, M = pb.n_fourier, K0 = 2 * M, H = kHidden;
  TcLayout t = tc_layout(pb);
  uint8_t* sv = align1k(saved_v);
  uint8_t* wk = align1k(work_v);
  int rc;
  // ---- per-call preparation: operand plan, features of all points, folded / split weights
  float* plan = reinterpret_cast<float*>(sv + t.plan);
  float* rowstat = reinterpret_cast<float*>(sv + t.rowstat);
  float* hstat = reinterpret_cast<float*>(sv + t.hstat);
  {
  ProfScope prep(KC_PREP, st);
  weight_stats_kernel<<<(unsigned)(L * H + 2 * L + 1), 256, 0, st>>>(pr.W[0], pr.Bff, pr.W[1], pr.W[2], rowstat, hstat, (int)L,
                                                                 (int)M);
  NSVD_LAUNCH_CHECK();
  fwd_plan_kernel<<<cdiv(L, 4), 128, 0, st>>>(rowstat, hstat, pr.Bff, pr.b[0], pr.b[1], pr.b[2], plan, (int)L, (int)M);
  NSVD_LAUNCH_CHECK();
  PrepArgs pa{};
  pa.x = x; pa.Bff = pr.Bff; pa.W0 = pr.W[0]; pa.W1 = pr.W[1]; pa.W2 = pr.W[2]; pa.plan = plan;
  pa.phi_hi = BF(sv + t.phi_hi); pa.phi_lo = BF(sv + t.phi_lo);
  pa.w0_hi = BF(wk + t.w0_hi); pa.w0_lo = BF(wk + t.w0_lo);
  for (int i = 0; i < 2; ++i) {
    pa.hp.hi[i] = BF(wk + t.w_hi[i]); pa.hp.lo[i] = BF(wk + t.w_lo[i]);
    pa.hp.hiT[i] = BF(wk + t.wT_hi[i]); pa.hp.loT[i] = BF(wk + t.wT_lo[i]);
  }
  pa.B = B; pa.L = (int)L; pa.M = (int)M;
  pa.n_feat = (unsigned)cdiv(B * (M / 4), 256);
  pa.n_fold = pb.fd_eps > 0.f ? 0u : (unsigned)cdiv(L * H * (M / 8), 256);   // FD mode never runs the 4-stream GEMM
  pa.n_split = (unsigned)(2 * L * 16);   // 32 x 32 tiles of the 128 x 128 hidden matrices
  unsigned n_w0v = 0;
  if (pb.fd_eps > 0.f) {
    pa.w0v_hi = BF(wk + t.w0v_hi); pa.w0v_lo = BF(wk + t.w0v_lo);
    n_w0v = (unsigned)cdiv(L * H * K0, 256);
  }
  prep_operands_kernel<<<pa.n_feat + pa.n_fold + pa.n_split + n_w0v, 256, 0, st>>>(pa);
  NSVD_LAUNCH_CHECK();
  }
  CUtensorMap mW0h, mW0l, mWh[2], mWl[2];
  const uint32_t w0_box = big::BN / 2;   // a CTA of a pair loads half of the 256 W' rows
  if ((rc = make_tmap_bf16_3d(&mW0h, wk + t.w0_hi, K0, 512, L, K0 * 2, 512 * K0 * 2, 64, w0_box))) return rc;
  if ((rc = make_tmap_bf16_3d(&mW0l, wk + t.w0_lo, K0, 512, L, K0 * 2, 512 * K0 * 2, 64, w0_box))) return rc;
  for (int i = 0; i < 2; ++i) {
    if ((rc = make_tmap_bf16_3d(&mWh[i], wk + t.w_hi[i], H, H, L, H * 2, H * H * 2, 64, 128))) return rc;
    if ((rc = make_tmap_bf16_3d(&mWl[i], wk + t.w_lo[i], H, H, L, H * 2, H * H * 2, 64, 128))) return rc;
  }
  for (long p0 = 0; p0 < B; p0 += t.P) {
    const int P = (int)((B - p0) < t.P ? (B - p0) : t.P);
    const int m_tiles = cdiv(P, 128);
    // ---- layer 0: Z0 = Phi . W0f^T  (S1, K-major) with the softplus-stream epilogue
    CUtensorMap mPh, mPl;
    if ((rc = make_tmap_bf16_3d(&mPh, sv + t.phi_hi + p0 * K0 * 2, K0, P, 1, K0 * 2, (uint64_t)P * K0 * 2, 64, big::BM))) return rc;
    if ((rc = make_tmap_bf16_3d(&mPl, sv + t.phi_lo + p0 * K0 * 2, K0, P, 1, K0 * 2, (uint64_t)P * K0 * 2, 64, big::BM))) return rc;
    BigShape s{};
    s.m_tiles = m_tiles;
    s.n_tiles = 2;
    s.batches = (int)L;
    s.k_slices = 1;
    s.k_chunks_total = cdiv(K0, big::BK);
    s.k_chunks_per_slice = s.k_chunks_total;
    s.a_batched = 0;
    s.b_batched = 1;
    static const int mgroup = env_int("NSVD_L0_MGROUP", 16), bgroup = env_int("NSVD_L0_BGROUP", 0);
    s.m_group = mgroup;   // 4096 points x 8 KB of Phi per group
    s.b_group = bgroup;   // alternative order: copies per group, Phi streamed once per group (profiles/README.md)
    L0FwdEpi e0{pr.b[0], plan, BF(wk + t.str_hi[0]), BF(wk + t.str_lo[0]), BF(sv + t.av_hi[0]), BF(sv + t.av_lo[0]), P, B, p0};
    const bool fd = pb.fd_eps > 0.f;
    if (!fd) {   // (finite-difference mode needs no derivative streams: its central pass is value-only, below)
      ProfScope ps(KC_L0_FWD, st);
      // K0 in sub-chains of 4 chunks (K = 256, 48 chained MMAs per TMEM accumulation), the first two of a tile 6 chunks.
      // Chains of 8 chunks miss the fixed-seed trajectory tolerance (profiles/trajectory_probe.py: 1.6e-3 vs 4.7e-4).
      static const int sub = env_int("NSVD_L0_SUBCHUNKS", 4);
      s.m_tiles = cdiv(P, 2 * big::BM);
      static const int sub_first = env_int("NSVD_L0_SUBFIRST", 6);
      if ((rc = launch_big2s<false, L0FwdEpi, kFmtHH>(mPh, mPl, mW0h, mW0l, s, (int)L, sub, sub_first, e0, st))) return rc;
    }
    // ---- hidden layers 1, 2 (+ head + operator)
    static const int fused = env_int("NSVD_HIDDEN_FUSED", 1);
    if (fused) {
      HidFwd12Maps hm;
      const uint64_t PH = (uint64_t)P * H * 2, BH = (uint64_t)B * H * 2;
      const int T = (int)L * m_tiles, grid = T < kHidGrid ? T : kHidGrid;
      if ((rc = make_tmap_bf16_3d(&hm.a0h, wk + t.str_hi[0], H, P, 4 * L, H * 2, PH, 64, 128))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.a0l, wk + t.str_lo[0], H, P, 4 * L, H * 2, PH, 64, 128))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.v0h, sv + t.av_hi[0], H, B, L, H * 2, BH, 64, 128))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.v0l, sv + t.av_lo[0], H, B, L, H * 2, BH, 64, 128))) return rc;
      hm.w1h = mWh[0];
      hm.w1l = mWl[0];
      hm.w2h = mWh[1];
      hm.w2l = mWl[1];
      if ((rc = make_tmap_bf16_3d(&hm.s1h, sv + t.av_hi[1], H, B, L, H * 2, BH, 32, 128, 64))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.s1l, sv + t.av_lo[1], H, B, L, H * 2, BH, 32, 128, 64))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.v1h, sv + t.av_hi[1], H, B, L, H * 2, BH, 64, 128))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.v1l, sv + t.av_lo[1], H, B, L, H * 2, BH, 64, 128))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.s2h, sv + t.av_hi[2], H, B, L, H * 2, BH, 32, 128, 64))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.s2l, sv + t.av_lo[2], H, B, L, H * 2, BH, 32, 128, 64))) return rc;
      const uint64_t nslot = (uint64_t)kHidGrid * 8;
      if ((rc = make_tmap_bf16_3d(&hm.xsh, wk + t.scr_hi, H, 128, nslot, H * 2, hid::PLANE, 32, 128, 64))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.xsl, wk + t.scr_lo, H, 128, nslot, H * 2, hid::PLANE, 32, 128, 64))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.xlh, wk + t.scr_hi, H, 128, nslot, H * 2, hid::PLANE, 64, 128))) return rc;
      if ((rc = make_tmap_bf16_3d(&hm.xll, wk + t.scr_lo, H, 128, nslot, H * 2, hid::PLANE, 64, 128))) return rc;
      HidFwd12Args a{};
      a.L = (int)L;
      a.P = P;
      a.m_tiles = m_tiles;
      a.Btot = B;
      a.p_off = p0;
      a.bias1 = pr.b[1];
      a.bias2 = pr.b[2];
      a.plan = plan;
      a.W3 = pr.W[3];
      a.b3 = pr.b[3];
      a.x = x;
      a.mscales = pr.mask_scales;
      a.F = F;
      a.TF = TF;
      a.U0 = reinterpret_cast<float*>(sv + t.u0);
      a.pb = pb;
      NSVD_SMEM_OPTIN(hidden_fwd12_kernel, hid::SMEM_FWD);
      if (!fd) {
        ProfScope ps(KC_HID_FWD, st);
        hidden_fwd12_kernel<<<grid, hid::F_THREADS, hid::SMEM_FWD, st>>>(hm, a);
        NSVD_LAUNCH_CHECK();
      } else {
        // ---- finite-difference Laplacian (pde/diff_ops.py:25-52), value streams only.
        // (1) central pass: F, U0 and the saved activations (what the backward needs) on the un-folded W0 - a quarter
        //     of the exact pass, since no derivative stream is propagated;
        // (2) shifted pass: the four stream slots carry the four shifted point sets, TF from their differences.
        CUtensorMap mVh, mVl;
        if ((rc = make_tmap_bf16_3d(&mVh, wk + t.w0v_hi, K0, L * H, 1, K0 * 2, L * H * K0 * 2, 64, big::BN / 2))) return rc;
        if ((rc = make_tmap_bf16_3d(&mVl, wk + t.w0v_lo, K0, L * H, 1, K0 * 2, L * H * K0 * 2, 64, big::BN / 2))) return rc;
        static const int sub = env_int("NSVD_L0_SUBCHUNKS", 4), sub_first = env_int("NSVD_L0_SUBFIRST", 6);
        {
          BigShape sc{};
          sc.m_tiles = cdiv(P, 2 * big::BM);
          sc.n_tiles = cdiv(L * H, big::BN);
          sc.batches = 1;
          sc.k_slices = 1;
          sc.k_chunks_total = cdiv(K0, big::BK);
          sc.k_chunks_per_slice = sc.k_chunks_total;
          sc.m_group = 16;
          L0ValEpi ec{pr.b[0], plan, BF(sv + t.av_hi[0]) + p0 * H, BF(sv + t.av_lo[0]) + p0 * H, (long)P, (long)P, 0L,
                      (long)B * H, (int)L};
          ProfScope ps(KC_L0_FWD, st);
          if ((rc = launch_big2s<false, L0ValEpi, kFmtHH>(mPh, mPl, mVh, mVl, sc, 1, sub, sub_first, ec, st))) return rc;
        }
        {
          HidFwd12Args ac = a;
          ac.vmode = 2;
          ac.m_tiles = cdiv(P, 512);
          const int Tc = (int)L * ac.m_tiles;
          ProfScope ps(KC_HID_FWD, st);
          hidden_fwd12_kernel<<<Tc < kHidGrid ? Tc : kHidGrid, hid::F_THREADS, hid::SMEM_FWD, st>>>(hm, ac);
          NSVD_LAUNCH_CHECK();
        }
        const float* xm = x + 2 * p0;
        features_shift_f16_kernel<<<cdiv(4L * P * (M / 4), 256), 256, 0, st>>>(xm, pr.Bff, BF(wk + t.phis_hi),
                                                                              BF(wk + t.phis_lo), P, (int)M, pb.fd_eps);
        NSVD_LAUNCH_CHECK();
        CUtensorMap mXh, mXl;
        if ((rc = make_tmap_bf16_3d(&mXh, wk + t.phis_hi, K0, 4 * (uint64_t)P, 1, K0 * 2, 4 * (uint64_t)P * K0 * 2, 64, big::BM))) return rc;
        if ((rc = make_tmap_bf16_3d(&mXl, wk + t.phis_lo, K0, 4 * (uint64_t)P, 1, K0 * 2, 4 * (uint64_t)P * K0 * 2, 64, big::BM))) return rc;
        BigShape sv2{};
        sv2.m_tiles = cdiv(4L * P, 2 * big::BM);
        sv2.n_tiles = cdiv(L * H, big::BN);
        sv2.batches = 1;
        sv2.k_slices = 1;
        sv2.k_chunks_total = cdiv(K0, big::BK);
        sv2.k_chunks_per_slice = sv2.k_chunks_total;
        sv2.m_group = 16;
        L0ValEpi ev{pr.b[0], plan, BF(wk + t.str_hi[0]), BF(wk + t.str_lo[0]), 4L * P, (long)P, (long)P * H, 4L * P * H, (int)L};
        {
          ProfScope ps(KC_L0_FWD, st);
          if ((rc = launch_big2s<false, L0ValEpi, kFmtHH>(mXh, mXl, mVh, mVl, sv2, 1, sub, sub_first, ev, st))) return rc;
        }
        a.vmode = 1;
        ProfScope ps(KC_HID_FWD, st);
        hidden_fwd12_kernel<<<grid, hid::F_THREADS, hid::SMEM_FWD, st>>>(hm, a);
        NSVD_LAUNCH_CHECK();
      }
      continue;
    }
    // one kernel per layer (NSVD_HIDDEN_FUSED=0): the derivative streams of a1 go through the second stream buffer
    NSVD_CHECK_ARG(!(pb.fd_eps > 0.f), "the finite-difference pass exists in the fused hidden kernel only (NSVD_HIDDEN_FUSED=1)");
    for (int i = 0; i < 2; ++i) {
      CUtensorMap mAh, mAl, mOh, mOl, mSh, mSl;
      if ((rc = make_tmap_bf16_3d(&mAh, wk + t.str_hi[i], H, P, 4 * L, H * 2, (uint64_t)P * H * 2, 64, 128))) return rc;
      if ((rc = make_tmap_bf16_3d(&mAl, wk + t.str_lo[i], H, P, 4 * L, H * 2, (uint64_t)P * H * 2, 64, 128))) return rc;
      // bulk-store maps (64-byte swizzle, boxes of 32 hidden units x 128 points)
      const int o = i == 0 ? 1 : 0;  // layer 2 has no stream output; give it valid (unused) maps
      if ((rc = make_tmap_bf16_3d(&mOh, wk + t.str_hi[o], H, P, 4 * L, H * 2, (uint64_t)P * H * 2, 32, 128, 64))) return rc;
      if ((rc = make_tmap_bf16_3d(&mOl, wk + t.str_lo[o], H, P, 4 * L, H * 2, (uint64_t)P * H * 2, 32, 128, 64))) return rc;
      if ((rc = make_tmap_bf16_3d(&mSh, sv + t.av_hi[i + 1], H, B, L, H * 2, (uint64_t)B * H * 2, 32, 128, 64))) return rc;
      if ((rc = make_tmap_bf16_3d(&mSl, sv + t.av_lo[i + 1], H, B, L, H * 2, (uint64_t)B * H * 2, 32, 128, 64))) return rc;
      // input value stream a_i of this layer, as saved by the previous one (rows = points of the whole batch)
      CUtensorMap mVh, mVl;
      if ((rc = make_tmap_bf16_3d(&mVh, sv + t.av_hi[i], H, B, L, H * 2, (uint64_t)B * H * 2, 64, 128))) return rc;
      if ((rc = make_tmap_bf16_3d(&mVl, sv + t.av_lo[i], H, B, L, H * 2, (uint64_t)B * H * 2, 64, 128))) return rc;
      HidFwdArgs a{};
      a.L = (int)L;
      a.P = P;
      a.m_tiles = m_tiles;
      a.Btot = B;
      a.p_off = p0;
      a.bias = pr.b[i + 1];
      a.plan = plan;
      a.u_slot = i == 0 ? PL_U1 : PL_U2;
      a.sa_slot = i == 0 ? PL_SA1 : PL_SA2;
      ProfScope ps(KC_HID_FWD, st);
      if (i == 0) {
        if ((rc = launch_hidden_fwd<false>(mAh, mAl, mWh[0], mWl[0], mOh, mOl, mSh, mSl, mVh, mVl, a, st))) return rc;
      } else {
        a.W3 = pr.W[3];
        a.b3 = pr.b[3];
        a.x = x;
        a.mscales = pr.mask_scales;
        a.F = F;
        a.TF = TF;
        a.U0 = reinterpret_cast<float*>(sv + t.u0);
        a.pb = pb;
        if ((rc = launch_hidden_fwd<true>(mAh, mAl, mWh[1], mWl[1], mOh, mOl, mSh, mSl, mVh, mVl, a, st))) return rc;
      }
    }
  }
  return 0;
}

// zero up to 10 fp32 buffers in one launch (gradients are accumulated with reductions: they start from zero)
struct ZeroList {
  float* p[10];
  long n[10];
  int count;
};
__global__ void __launch_bounds__(256) zero_list_kernel(ZeroList z) {
  const long stride = (long)gridDim.x * blockDim.x, t0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int b = 0; b < z.count; ++b) {
    float* p = z.p[b];
    const long n = z.n[b];
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      float4* p4 = reinterpret_cast<float4*>(p);
      for (long i = t0; i < n / 4; i += stride) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (long i = (n / 4) * 4 + t0; i < n; i += stride) p[i] = 0.f;
    } else {
      for (long i = t0; i < n; i += stride) p[i] = 0.f;
    }
  }
}

int tc_backward(const nsvd_problem_t& pb, const nsvd_params_t& pr, const float* x, const float* dF,
                const void* saved_v, nsvd_grads_t& gr, void* work_v, size_t work_bytes, cudaStream_t st) {
  (void)work_bytes;
  const long B = pb.n_points, L = pb.n_copies, M = pb.n_fourier, K0 = 2 * M, H = kHidden;
  TcLayout t = tc_layout(pb);
  uint8_t* sv = align1k(const_cast<void*>(saved_v));
  uint8_t* wk = align1k(work_v);
  int rc;
  NSVD_SMEM_OPTIN(hidden_bwd2_kernel, hid::SMEM_BWD2);
  // gradients are accumulated with reductions: start from zero
  // (one launch instead of ten memsets: the small-batch configurations are launch-bound)
  float* mdf = reinterpret_cast<float*>(sv + t.mdf);
  {
    ZeroList z{};
    z.p[0] = gr.dW[0]; z.n[0] = L * H * K0;
    z.p[1] = gr.dW[1]; z.n[1] = L * H * H;
    z.p[2] = gr.dW[2]; z.n[2] = L * H * H;
    z.p[3] = gr.dW[3]; z.n[3] = L * H;
    for (int i = 0; i < 3; ++i) { z.p[4 + i] = gr.db[i]; z.n[4 + i] = L * H; }
    z.p[7] = gr.db[3]; z.n[7] = L;
    z.p[8] = mdf; z.n[8] = L;
    z.count = 9;
    if (pb.has_exp_mask && gr.dmask_scales) { z.p[9] = gr.dmask_scales; z.n[9] = L; z.count = 10; }
    long total = 0;
    for (int i = 0; i < z.count; ++i) total += z.n[i];
    int nb = cdiv(total / 4 + 1, 256);
    zero_list_kernel<<<nb < 148 * 8 ? nb : 148 * 8, 256, 0, st>>>(z);
    NSVD_LAUNCH_CHECK();
  }
  // backward half of the operand plan: scales of the dZ planes from max|dF| per copy
  float* plan = reinterpret_cast<float*>(sv + t.plan);
  {
    const long want = cdiv(B * L, 256 * 8);   // >= 8 elements per thread; small batches get a small grid
    col_absmax_kernel<<<(unsigned)(want < 148 * 2 ? (want > 1 ? want : 1) : 148 * 2), 256, 0, st>>>(dF, B * L, (int)L, mdf);
  }
  NSVD_LAUNCH_CHECK();
  // transposed hidden weights (dgrad B operand, K-major): WT_i[l][k][j] = W_i[l][j][k], written by the forward's
  // operand preparation (the version check of the Python layer guarantees the forward of THIS step wrote them)
  CUtensorMap mWh[2], mWl[2];
  for (int i = 0; i < 2; ++i) {
    if ((rc = make_tmap_bf16_3d(&mWh[i], wk + t.wT_hi[i], H, H, L, H * 2, H * H * 2, 64, 128))) return rc;
    if ((rc = make_tmap_bf16_3d(&mWl[i], wk + t.wT_lo[i], H, H, L, H * 2, H * H * 2, 64, 128))) return rc;
  }
  for (long p0 = 0; p0 < B; p0 += t.P) {
    const int P = (int)((B - p0) < t.P ? (B - p0) : t.P);
    const int m_tiles = cdiv(P, 128);
    // ---- head: dZ2 planes (pair 0), dW3, db3, db2, dscales
    const int hchunk = P <= 8192 ? 32 : kHeadChunk;
    dim3 hg(cdiv(P, hchunk), (unsigned)L);
    {
      ProfScope ps(KC_HEAD_BWD, st);
      head_bwd_bf16_kernel<<<hg, 128, 0, st>>>(dF, reinterpret_cast<const float*>(sv + t.u0), BF(sv + t.av_hi[2]),
                                               BF(sv + t.av_lo[2]), pr.W[3], x, pr.mask_scales, plan,
                                               reinterpret_cast<const float*>(sv + t.hstat), mdf, pb, BF(wk + t.dz_hi[0]),
                                               BF(wk + t.dz_lo[0]), gr.dW[3], gr.db[3], gr.db[2], gr.dmask_scales, P, B, p0, hchunk);
      NSVD_LAUNCH_CHECK();
    }
    // ---- hidden layers 2, 1: dgrad + wgrad in one pass
    int cur = 0;
    for (int i = 2; i >= 1; --i) {
      CUtensorMap mZh, mZl, mAh, mAl;
      const int rows = hid::HROWS;   // points per half tile = TMA box height
      if ((rc = make_tmap_bf16_3d(&mZh, wk + t.dz_hi[cur], H, P, L, H * 2, (uint64_t)P * H * 2, 64, rows))) return rc;
      if ((rc = make_tmap_bf16_3d(&mZl, wk + t.dz_lo[cur], H, P, L, H * 2, (uint64_t)P * H * 2, 64, rows))) return rc;
      if ((rc = make_tmap_bf16_3d(&mAh, sv + t.av_hi[i - 1], H, B, L, H * 2, (uint64_t)B * H * 2, 64, rows))) return rc;
      if ((rc = make_tmap_bf16_3d(&mAl, sv + t.av_lo[i - 1], H, B, L, H * 2, (uint64_t)B * H * 2, 64, rows))) return rc;
      CUtensorMap mOh, mOl;
      if ((rc = make_tmap_bf16_3d(&mOh, wk + t.dz_hi[cur ^ 1], H, P, L, H * 2, (uint64_t)P * H * 2, 64, rows))) return rc;
      if ((rc = make_tmap_bf16_3d(&mOl, wk + t.dz_lo[cur ^ 1], H, P, L, H * 2, (uint64_t)P * H * 2, 64, rows))) return rc;
      HidBwdArgs a{};
      a.L = (int)L;
      a.P = P;
      a.m_tiles = cdiv(P, rows);
      a.Btot = B;
      a.p_off = p0;
      a.dW = gr.dW[i];
      a.db_prev = gr.db[i - 1];
      a.plan = plan;
      a.ud_slot = i == 2 ? PL_UD2 : PL_UD1;
      a.inv_sa_slot = i == 2 ? PL_INV_SA1 : PL_INV_SA0;
      a.sdz_slot = i == 2 ? PL_SDZ1 : PL_SDZ0;
      a.uw_slot = i == 2 ? PL_UW2 : PL_UW1;
      int T = (int)L * a.m_tiles;
      int grid = T < 148 ? T : 148;
      {
        ProfScope ps(KC_HID_BWD, st);
        hidden_bwd2_kernel<<<grid, hid::F_THREADS, hid::SMEM_BWD2, st>>>(mZh, mZl, mAh, mAl, mWh[i - 1], mWl[i - 1], mOh, mOl, a);
        NSVD_LAUNCH_CHECK();
      }
      cur ^= 1;
    }
    // ---- layer 0 weight gradient: dW0[l] += dZ0[l]^T . Phi  (S1, MN-major, K = points in slices)
    CUtensorMap mZh, mZl, mPh, mPl;
    if ((rc = make_tmap_bf16_3d(&mZh, wk + t.dz_hi[cur], H, P, L, H * 2, (uint64_t)P * H * 2, 64, 64))) return rc;
    if ((rc = make_tmap_bf16_3d(&mZl, wk + t.dz_lo[cur], H, P, L, H * 2, (uint64_t)P * H * 2, 64, 64))) return rc;
    if ((rc = make_tmap_bf16_3d(&mPh, sv + t.phi_hi + p0 * K0 * 2, K0, P, 1, K0 * 2, (uint64_t)P * K0 * 2, 64, 64))) return rc;
    if ((rc = make_tmap_bf16_3d(&mPl, sv + t.phi_lo + p0 * K0 * 2, K0, P, 1, K0 * 2, (uint64_t)P * K0 * 2, 64, 64))) return rc;
    BigShape s{};
    s.m_tiles = 1;
    s.n_tiles = cdiv(K0, big::BN);
    s.batches = (int)L;
    s.k_chunks_total = cdiv(P, big::BK);
    static const int kslice = env_int("NSVD_WGRAD_KSLICE", 16), kgroup = env_int("NSVD_WGRAD_KGROUP", 4);
    s.k_chunks_per_slice = kslice;  // 1024 points per accumulation (fp32 TMEM accumulation stays short)
    s.k_slices = cdiv(s.k_chunks_total, s.k_chunks_per_slice);
    s.a_batched = 1;
    s.b_batched = 0;
    s.k_group = kgroup;   // 4096 points per group: dZ0 (all copies) + Phi of a group stay in L2
    L0WgradEpi ew{gr.dW[0], plan, (int)K0};
    {
      ProfScope ps(KC_L0_WGRAD, st);
      static const int sub = env_int("NSVD_WGRAD_SUBCHUNKS", 8);   // 512 points per TMEM accumulation chain
      s.batches = cdiv(L, 2);   // a CTA pair stacks two copies along M
      if ((rc = launch_big2s<true, L0WgradEpi, kFmtHH>(mZh, mZl, mPh, mPl, s, (int)L, sub, sub, ew, st))) return rc;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// CDK loss on the tcgen05 GEMM block (methods/nestedlora.py:270-332; BASELINE config 5:
// B = 4096, L = 512 (+1 constant mode)).  Operand planes are [rows][Lpad] bf16 hi/lo with the
// constant-1 column in front and zero padding up to a multiple of 8 columns (TMA stride rule).
// ------------------------------------------------------------------------------------------
static inline long cdk_lpad(int Lp) { return (Lp + 7) / 8 * 8; }

// planes of the padded inputs: out[b][c] = (fc && c == 0) ? 1 : in[b][c - fc] ; zero for c >= Lp
__global__ void cdk_planes_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long B, int L, int fc, int Lpad) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Lpad) return;
  long b = i / Lpad;
  int c = (int)(i % Lpad);
  float v = 0.f;
  if (c < L + fc) v = (fc && c == 0) ? 1.f : in[b * L + c - fc];
  __nv_bfloat16 a, d;
  tc::split_bf16(v, a, d);
  hi[i] = a;
  lo[i] = d;
}
// Forward preparation in ONE launch (it was five: two padded fp32 copies, the row dots over them, two plane splits):
// a warp per row reads f[b], g[b] once, writes the hi/lo planes of both padded rows and takes the exact fp32 row dots
//   opdot[b] = sum_c v_c Fp[b][c] Gp[b][c],   rs_joint[b] = diag(Fp Gp^T)[b]
// in the lane-strided order of cdk_rowdots_kernel (same values to the last bit).
__global__ void __launch_bounds__(256)
cdk_prep_kernel(const float* __restrict__ f, const float* __restrict__ g, const float* __restrict__ v,
                __nv_bfloat16* __restrict__ fh, __nv_bfloat16* __restrict__ fl, __nv_bfloat16* __restrict__ gh,
                __nv_bfloat16* __restrict__ gl, float* __restrict__ opdot, float* __restrict__ rs_joint, long B, int L,
                int fc, int Lpad) {
  const long b = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int Lp = L + fc;
  float a = 0.f, d = 0.f;
  for (int c = lane; c < Lpad; c += 32) {
    float fv = 0.f, gv = 0.f;
    if (c < Lp) {
      const bool one = fc && c == 0;
      fv = one ? 1.f : f[b * L + c - fc];
      gv = one ? 1.f : g[b * L + c - fc];
      const float p = fv * gv;
      d += p;
      a = fmaf(v[c], p, a);
    }
    __nv_bfloat16 h, l;
    tc::split_bf16(fv, h, l);
    fh[b * Lpad + c] = h;
    fl[b * Lpad + c] = l;
    tc::split_bf16(gv, h, l);
    gh[b * Lpad + c] = h;
    gl[b * Lpad + c] = l;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
  if (lane == 0) {
    opdot[b] = a;
    if (rs_joint) rs_joint[b] = d;
  }
}
// transposed coefficient planes: out[which][c][i] = coef[which][i][c]  (K-major B operand of the backward GEMM)
__global__ void cdk_coefT_planes_kernel(const float* __restrict__ coef, __nv_bfloat16* __restrict__ hi,
                                        __nv_bfloat16* __restrict__ lo, int Lp, int Lpad) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n = 2L * Lp * Lpad;
  if (i >= n) return;
  int k = (int)(i % Lpad);
  int c = (int)((i / Lpad) % Lp);
  int w = (int)(i / ((long)Lpad * Lp));
  float v = k < Lp ? coef[(long)w * Lp * Lp + (long)k * Lp + c] : 0.f;
  __nv_bfloat16 a, d;
  tc::split_bf16(v, a, d);
  hi[i] = a;
  lo[i] = d;
}

// In-register transpose of a 32 x 32 fp32 block held one ROW per lane (a[k] = element (lane, k), what a 32x32b TMEM load
// delivers) into one COLUMN per lane (a[k] = element (k, lane)): five butterfly stages of 16 shuffles with static
// register indices.  The CDK epilogues below write row-major matrices whose row starts are not 16-byte aligned (the
// dropped constant column / the removed diagonal shift every row by one element), so vector stores per row are not
// available; with a column per lane every store and every load of a warp is one contiguous 128-byte segment.  The
// first versions stored one row per lane (32 sectors per instruction): 65 us per backward GEMM, 147 us for Fp Gp^T.
__device__ __forceinline__ void warp_transpose32(float (&a)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((i & s) == 0) {
        const float send = up ? a[i] : a[i + s];
        const float recv = __shfl_xor_sync(0xffffffffu, send, s);
        if (up) a[i] = recv;
        else a[i + s] = recv;
      }
    }
  }
}

// grad[b][c - fc] = gs * ( acc[b][c] - c2 v[c] other[b][c - fc] )
struct CdkBwdEpi {
  float* grad;
  const float* other;
  const float* v;
  const float* gscale;
  int B, L, fc;
  float c2;
  __device__ __forceinline__ void operator()(uint32_t tmem_acc, const TileCoord& c, int ewarp, int lane) const {
    const int q = ewarp & 3, half = ewarp >> 2;
    const int b0 = c.mt * big::BM + q * 32;          // first of this warp's 32 rows
    const float gs = gscale ? gscale[0] : 1.f;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      const int col0 = half * 128 + ch * 32;
      float a[32];
      tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + col0, a);
      tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + col0 + 16, a + 16);
      tc::tmem_ld_wait();
      warp_transpose32(a, lane);                     // a[k] = acc[b0 + k][col0 + lane]
      const int cc = c.nt * big::BN + col0 + lane;
      if (cc >= fc && cc < L + fc) {
        const float cv = c2 * v[cc];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          if (b0 + k < B) {
            const long o = (long)(b0 + k) * L + cc - fc;
            grad[o] = gs * (a[k] - cv * other[o]);
          }
        }
      }
    }
  }
};

// off_diagonal(Fp Gp^T): out[i (B-1) + j - (j > i)] = acc[i][j], i != j   (methods/utils.py:16-22)
struct CdkOffdiagEpi {
  float* out;
  int B;
  __device__ __forceinline__ void operator()(uint32_t tmem_acc, const TileCoord& c, int ewarp, int lane) const {
    const int q = ewarp & 3, half = ewarp >> 2;
    const int i0 = c.mt * big::BM + q * 32;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      const int col0 = half * 128 + ch * 32;
      float a[32];
      tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + col0, a);
      tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + col0 + 16, a + 16);
      tc::tmem_ld_wait();
      warp_transpose32(a, lane);                     // a[k] = acc[i0 + k][col0 + lane]
      const int j = c.nt * big::BN + col0 + lane;
      if (j < B) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int i = i0 + k;
          if (i < B && j != i) __stcs(out + (long)i * (B - 1) + j - (j > i ? 1 : 0), a[k]);   // 67 MB, written once
        }
      }
    }
  }
};

struct CdkLayout {
  size_t opdot, f_hi, f_lo, g_hi, g_lo, c_hi, c_lo, total;
};
static CdkLayout cdk_layout(long B, int L, int fc) {
  CdkLayout t{};
  const long Lp = L + fc, Lpad = cdk_lpad((int)Lp);
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o += align_up(bytes, 1024);
    return r;
  };
  t.opdot = take(B * 4);
  t.f_hi = take(B * Lpad * 2);
  t.f_lo = take(B * Lpad * 2);
  t.g_hi = take(B * Lpad * 2);
  t.g_lo = take(B * Lpad * 2);
  t.c_hi = take(2 * Lp * Lpad * 2);
  t.c_lo = take(2 * Lp * Lpad * 2);
  t.total = o + 1024;
  return t;
}
size_t tc_cdk_work_bytes(int B, int L, int fc) { return cdk_layout(B, L, fc).total; }

static int cdk_make_planes(const float* f, const float* g, long B, int L, int fc, uint8_t* wk, const CdkLayout& t,
                           cudaStream_t st) {
  const int Lpad = (int)cdk_lpad(L + fc);
  long n = B * Lpad;
  cdk_planes_kernel<<<cdiv(n, 256), 256, 0, st>>>(f, BF(wk + t.f_hi), BF(wk + t.f_lo), B, L, fc, Lpad);
  NSVD_LAUNCH_CHECK();
  cdk_planes_kernel<<<cdiv(n, 256), 256, 0, st>>>(g, BF(wk + t.g_hi), BF(wk + t.g_lo), B, L, fc, Lpad);
  NSVD_LAUNCH_CHECK();
  return 0;
}

int tc_cdk_fwd(const float* f, const float* g, const float* v, int B, int L, int fc, float* terms, float* rs_joint,
               void* work, cudaStream_t st) {
  const int Lp = L + fc, Lpad = (int)cdk_lpad(Lp);
  CdkLayout t = cdk_layout(B, L, fc);
  uint8_t* wk = align1k(work);
  int rc;
  // operand planes of the padded rows + exact fp32 row dots (operator term, diag(Fp Gp^T)), one launch
  float* opdot = reinterpret_cast<float*>(wk + t.opdot);
  cdk_prep_kernel<<<cdiv((long)B * 32, 256), 256, 0, st>>>(f, g, v, BF(wk + t.f_hi), BF(wk + t.f_lo), BF(wk + t.g_hi),
                                                           BF(wk + t.g_lo), opdot, rs_joint, B, L, fc, Lpad);
  NSVD_LAUNCH_CHECK();
  // Grams: X^T X with both operands MN-major straight from the [B][Lpad] planes, K = rows in slices
  NSVD_CUDA(cudaMemsetAsync(terms, 0, sizeof(float) * 2L * Lp * Lp, st));
  if ((rc = colsum(opdot, terms + 2L * Lp * Lp, B, 1, 1, 0, 0, st))) return rc;
  for (int which = 0; which < 2; ++which) {
    const uint8_t* ph = wk + (which ? t.g_hi : t.f_hi);
    const uint8_t* pl = wk + (which ? t.g_lo : t.f_lo);
    CUtensorMap mh, ml;
    if ((rc = make_tmap_bf16_3d(&mh, ph, Lpad, B, 1, (uint64_t)Lpad * 2, (uint64_t)B * Lpad * 2, 64, 64))) return rc;
    if ((rc = make_tmap_bf16_3d(&ml, pl, Lpad, B, 1, (uint64_t)Lpad * 2, (uint64_t)B * Lpad * 2, 64, 64))) return rc;
    BigShape s{};
    s.m_tiles = cdiv(Lp, big::BM);
    s.n_tiles = cdiv(Lp, big::BN);
    s.batches = 1;
    s.k_chunks_total = cdiv(B, big::BK);
    s.k_chunks_per_slice = 8;
    s.k_slices = cdiv(s.k_chunks_total, s.k_chunks_per_slice);
    StoreEpi epi{terms + (long)which * Lp * Lp, Lp, Lp, 0, 1};
    if ((rc = launch_big<true>(mh, ml, mh, ml, s, epi, st))) return rc;
  }
  return 0;
}

int tc_cdk_bwd(const float* f, const float* g, const float* v, const float* coef, const float* gscale, int B, int L,
               int fc, long Bg, float* grad_f, float* grad_g, void* work, int planes_ready, cudaStream_t st) {
  const int Lp = L + fc, Lpad = (int)cdk_lpad(Lp);
  CdkLayout t = cdk_layout(B, L, fc);
  uint8_t* wk = align1k(work);
  int rc;
  if (!planes_ready && (rc = cdk_make_planes(f, g, B, L, fc, wk, t, st))) return rc;
  long nc = 2L * Lp * Lpad;
  cdk_coefT_planes_kernel<<<cdiv(nc, 256), 256, 0, st>>>(coef, BF(wk + t.c_hi), BF(wk + t.c_lo), Lp, Lpad);
  NSVD_LAUNCH_CHECK();
  const bool pair = tc_use_pair();
  const float c2 = (float)(2.0 / (double)Bg);
  for (int which = 0; which < 2; ++which) {
    const uint8_t* xh = wk + (which ? t.g_hi : t.f_hi);
    const uint8_t* xl = wk + (which ? t.g_lo : t.f_lo);
    const uint8_t* ch = wk + t.c_hi + (size_t)which * Lp * Lpad * 2;
    const uint8_t* cl = wk + t.c_lo + (size_t)which * Lp * Lpad * 2;
    CUtensorMap mah, mal, mbh, mbl;
    const uint32_t bbox = pair ? big::BN / 2 : big::BN;
    if ((rc = make_tmap_bf16_3d(&mah, xh, Lpad, B, 1, (uint64_t)Lpad * 2, (uint64_t)B * Lpad * 2, 64, big::BM))) return rc;
    if ((rc = make_tmap_bf16_3d(&mal, xl, Lpad, B, 1, (uint64_t)Lpad * 2, (uint64_t)B * Lpad * 2, 64, big::BM))) return rc;
    if ((rc = make_tmap_bf16_3d(&mbh, ch, Lpad, Lp, 1, (uint64_t)Lpad * 2, (uint64_t)Lp * Lpad * 2, 64, bbox))) return rc;
    if ((rc = make_tmap_bf16_3d(&mbl, cl, Lpad, Lp, 1, (uint64_t)Lpad * 2, (uint64_t)Lp * Lpad * 2, 64, bbox))) return rc;
    BigShape s{};
    s.n_tiles = cdiv(Lp, big::BN);
    s.batches = 1;
    s.k_slices = 1;
    s.k_chunks_total = cdiv(Lpad, big::BK);
    s.k_chunks_per_slice = s.k_chunks_total;
    CdkBwdEpi epi{which ? grad_g : grad_f, which ? f : g, v, gscale, B, L, fc, c2};
    if (pair) {
      s.m_tiles = cdiv(B, 2 * big::BM);
      if ((rc = launch_big2<false>(mah, mal, mbh, mbl, s, 1, epi, st))) return rc;
    } else {
      s.m_tiles = cdiv(B, big::BM);
      if ((rc = launch_big<false>(mah, mal, mbh, mbl, s, epi, st))) return rc;
    }
  }
  return 0;
}

int tc_cdk_offdiag(const float* f, const float* g, int B, int L, int fc, float* out, void* work, int planes_ready,
                   cudaStream_t st) {
  const int Lp = L + fc, Lpad = (int)cdk_lpad(Lp);
  CdkLayout t = cdk_layout(B, L, fc);
  uint8_t* wk = align1k(work);
  int rc;
  if (!planes_ready && (rc = cdk_make_planes(f, g, B, L, fc, wk, t, st))) return rc;
  const bool pair = tc_use_pair();
  const uint32_t bbox = pair ? big::BN / 2 : big::BN;
  CUtensorMap mah, mal, mbh, mbl;
  if ((rc = make_tmap_bf16_3d(&mah, wk + t.f_hi, Lpad, B, 1, (uint64_t)Lpad * 2, (uint64_t)B * Lpad * 2, 64, big::BM))) return rc;
  if ((rc = make_tmap_bf16_3d(&mal, wk + t.f_lo, Lpad, B, 1, (uint64_t)Lpad * 2, (uint64_t)B * Lpad * 2, 64, big::BM))) return rc;
  if ((rc = make_tmap_bf16_3d(&mbh, wk + t.g_hi, Lpad, B, 1, (uint64_t)Lpad * 2, (uint64_t)B * Lpad * 2, 64, bbox))) return rc;
  if ((rc = make_tmap_bf16_3d(&mbl, wk + t.g_lo, Lpad, B, 1, (uint64_t)Lpad * 2, (uint64_t)B * Lpad * 2, 64, bbox))) return rc;
  BigShape s{};
  s.n_tiles = cdiv(B, big::BN);
  s.batches = 1;
  s.k_slices = 1;
  s.k_chunks_total = cdiv(Lpad, big::BK);
  s.k_chunks_per_slice = s.k_chunks_total;
  CdkOffdiagEpi epi{out, B};
  if (pair) {
    s.m_tiles = cdiv(B, 2 * big::BM);
    return launch_big2<false>(mah, mal, mbh, mbl, s, 1, epi, st);
  }
  s.m_tiles = cdiv(B, big::BM);
  return launch_big<false>(mah, mal, mbh, mbl, s, epi, st);
}

// ------------------------------------------------------------------------------------------
// Dense layers of the CDK encoder (examples/models/mlp.py:129-164 `get_mlp`, siam.py:132-165
// `HeteroNetwork`; main_sketchy.py:107-115 builds two 512 -> 8192 -> 512 towers) on the CTA-pair GEMM block with
// bounded accumulation chains (big2s), 3 products of fp16 hi/lo planes:
//   forward   y  = act(x W^T + b)          K-major,  K = in features
//   backward  dz = dy * act'(y);  dx = dz W (K-major, K = out features, transposed W planes);
//             dW = dz^T x (MN-major, K = rows, the out features in column blocks of 128 stacked by the CTA pair);
//             db = column sums of dz.
// Every tensor is stored as planes of s v with s = 2^k from its measured max |v| (one reduction pass per tensor), so
// that |s v| <= 2^15: 22 significant bits.  (bf16 planes, 16 bits, need no scale but leave 1e-5 in y - enough for y itself,
// not for the gradients behind the activation kink: with 33 M hidden activations per tower, ~60 of them change sign
// against the fp32 reference and dW of the first layer moves by 1.4e-3, measured.)
// ------------------------------------------------------------------------------------------
struct AbsmaxList {
  const float* p[3];
  long n[3];
  int count;
};
// out[t] = max |p[t][i]|  (out zero-initialised; non-negative floats order like their bit patterns).  16-byte loads,
// four per thread in flight (the scalar grid-stride loop this replaces read a 134 MB activation tensor at 1.3 TB/s).
__global__ void __launch_bounds__(256) absmax_list_kernel(AbsmaxList a, float* __restrict__ out) {
  __shared__ float red[8];
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long)gridDim.x * blockDim.x;
  for (int t = 0; t < a.count; ++t) {
    const float* p = a.p[t];
    const long n = a.n[t];
    float m = 0.f;
    long done = 0;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const float4* p4 = reinterpret_cast<const float4*>(p);
      const long n4 = n >> 2;
      long i = gtid;
      for (; i + 3 * gsz < n4; i += 4 * gsz) {
        const float4 v0 = __ldcs(p4 + i), v1 = __ldcs(p4 + i + gsz), v2 = __ldcs(p4 + i + 2 * gsz),
                     v3 = __ldcs(p4 + i + 3 * gsz);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))));
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w))));
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v2.x), fabsf(v2.y)), fmaxf(fabsf(v2.z), fabsf(v2.w))));
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v3.x), fabsf(v3.y)), fmaxf(fabsf(v3.z), fabsf(v3.w))));
      }
      for (; i < n4; i += gsz) {
        const float4 v = __ldcs(p4 + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      }
      done = n4 << 2;
    }
    for (long i = done + gtid; i < n; i += gsz) m = fmaxf(m, fabsf(p[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
      if (m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(out) + t, __float_as_uint(m));
    }
    __syncthreads();
  }
}

struct LinearEpi {
  float* D;
  const float* bias;        // per output column, or null
  const float *amax_a, *amax_b;   // max |.| of the two operands (their planes carry pow2_scale(max) * v)
  int M, N;                 // valid rows / columns of D
  long ldd;
  int rows_per_batch;       // MN-major: batch b covers the rows [b * rows_per_batch, +128)
  int act;                  // 0 none, 1 leaky ReLU with `slope` (0 = ReLU)
  float slope;
  __device__ static __forceinline__ uint32_t col0(int sub, int j) { return (uint32_t)(sub * 64 + j * 16); }
  __device__ __forceinline__ void operator()(float (&r)[64], const TileCoord& c, int q, int sub, int lane,
                                             uint8_t*) const {
    const long row = (long)c.b * rows_per_batch + (long)c.mt * big::BM + q * 32 + lane;
    if (row >= M) return;
    const float un = 1.f / (pow2_scale(__ldg(amax_a)) * pow2_scale(__ldg(amax_b)));
    float* drow = D + row * ldd;
    const bool vec = (ldd & 3) == 0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n0 = c.nt * big::BN + sub * 64 + g * 16;
      if (n0 < N) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float v = r[g * 16 + i] * un;
          if (bias && n0 + i < N) v += __ldg(bias + n0 + i);
          if (act == 1) v = v > 0.f ? v : v * slope;
          r[g * 16 + i] = v;
        }
        if (vec && n0 + 16 <= N) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(drow + n0 + i) = make_float4(r[g * 16 + i], r[g * 16 + i + 1], r[g * 16 + i + 2],
                                                                    r[g * 16 + i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n0 + i < N) drow[n0 + i] = r[g * 16 + i];
        }
      }
    }
  }
};

// fp16 hi/lo planes of pow2_scale(*amax) * src
// planes of pow2_scale(max) * src; a thread converts 8 consecutive values (two 16-byte loads, one 16-byte store per
// plane) when the three pointers allow it, the remainder one value per thread
__global__ void __launch_bounds__(256) split_scaled_kernel(const float* __restrict__ src, const float* __restrict__ amax,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                           long n, long n_vec) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const float sc = pow2_scale(__ldg(amax));
  if (i < n_vec) {
    const float4 v0 = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 v1 = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint32_t h[4], l[4];
    tc::split2<tc::PF_HH>(v0.x * sc, v0.y * sc, h[0], l[0]);
    tc::split2<tc::PF_HH>(v0.z * sc, v0.w * sc, h[1], l[1]);
    tc::split2<tc::PF_HH>(v1.x * sc, v1.y * sc, h[2], l[2]);
    tc::split2<tc::PF_HH>(v1.z * sc, v1.w * sc, h[3], l[3]);
    reinterpret_cast<uint4*>(hi)[i] = make_uint4(h[0], h[1], h[2], h[3]);
    reinterpret_cast<uint4*>(lo)[i] = make_uint4(l[0], l[1], l[2], l[3]);
    return;
  }
  const long e = 8 * n_vec + (i - n_vec);
  if (e >= n) return;
  uint16_t a, b;
  tc::split1<tc::PF_HH>(src[e] * sc, a, b);
  reinterpret_cast<uint16_t*>(hi)[e] = a;
  reinterpret_cast<uint16_t*>(lo)[e] = b;
}
static int launch_split_scaled(const float* src, const float* amax, __nv_bfloat16* hi, __nv_bfloat16* lo, long n,
                               cudaStream_t st) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(hi) |
                         reinterpret_cast<uintptr_t>(lo)) & 15) == 0;
  const long n_vec = aligned ? n / 8 : 0;
  const long threads = n_vec + (n - 8 * n_vec);
  if (threads <= 0) return 0;
  split_scaled_kernel<<<cdiv(threads, 256), 256, 0, st>>>(src, amax, hi, lo, n, n_vec);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// src (R, C) fp32 -> planes of the TRANSPOSE (C, R); 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) split_scaled_transposed_kernel(const float* __restrict__ src,
                                                                      const float* __restrict__ amax,
                                                                      __nv_bfloat16* __restrict__ hi,
                                                                      __nv_bfloat16* __restrict__ lo, int R, int C) {
  __shared__ uint32_t tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float sc = pow2_scale(__ldg(amax));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    uint16_t a = 0, b = 0;
    if (r < R && c < C) tc::split1<tc::PF_HH>(src[(long)r * C + c] * sc, a, b);
    tile[ty + 8 * k][tx] = (uint32_t)a | ((uint32_t)b << 16);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (c < C && r < R) {
      const uint32_t v = tile[tx][ty + 8 * k];
      reinterpret_cast<uint16_t*>(hi)[(long)c * R + r] = (uint16_t)(v & 0xffffu);
      reinterpret_cast<uint16_t*>(lo)[(long)c * R + r] = (uint16_t)(v >> 16);
    }
  }
}

// dz = dy * act'(y) -> planes scaled by pow2_scale(max |dy|) (|act'| <= 1); db[col] += sum over this block's 64 rows.
// grid (cols / 256, rows / 64)
__global__ void __launch_bounds__(256) linear_dz_kernel(const float* __restrict__ dy, const float* __restrict__ y, int act,
                                                        float slope, const float* __restrict__ amax,
                                                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                        float* __restrict__ db, int M, int N) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= N) return;
  const int r0 = blockIdx.y * 64, r1 = r0 + 64 < M ? r0 + 64 : M;
  const float sc = pow2_scale(__ldg(amax));
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) {
    const long o = (long)r * N + col;
    float g = dy[o];
    if (act == 1 && !(y[o] > 0.f)) g *= slope;
    acc += g;
    uint16_t a, b;
    tc::split1<tc::PF_HH>(g * sc, a, b);
    reinterpret_cast<uint16_t*>(hi)[o] = a;
    reinterpret_cast<uint16_t*>(lo)[o] = b;
  }
  if (db) atomicAdd(db + col, acc);
}

static size_t linear_planes(long rows, long cols) { return align_up((size_t)rows * cols * 2, 1024); }
size_t tc_linear_work_bytes(int rows, int in_f, int out_f) {
  // 1 KB of scales, then (backward needs the most) dz planes, transposed W planes, x planes (hi + lo each)
  return 1024 + 2 * (linear_planes(rows, out_f) + linear_planes(out_f, in_f) + linear_planes(rows, in_f)) + 2048;
}

static int linear_check(int rows, int in_f, int out_f, float slope, const void* work, size_t work_bytes) {
  NSVD_CHECK_ARG(rows >= 1 && in_f >= 8 && out_f >= 8 && in_f % 8 == 0 && out_f % 8 == 0,
                 "dense layer: rows >= 1, in/out features positive multiples of 8 (got %d x %d -> %d)", rows, in_f, out_f);
  NSVD_CHECK_ARG(slope >= -1.f && slope <= 1.f, "dense layer: |slope| must be <= 1 (got %g)", (double)slope);
  NSVD_CHECK_ARG(work != nullptr, "work is NULL");
  if (work_bytes < tc_linear_work_bytes(rows, in_f, out_f)) {
    set_error("dense layer work too small: %zu < %zu", work_bytes, tc_linear_work_bytes(rows, in_f, out_f));
    return NSVD_E_WORKSPACE;
  }
  return 0;
}

constexpr int kLinearSub = 8;   // K chunks (512 terms x 3 products) per TMEM accumulation chain

static int linear_absmax(const float* a, long na, const float* b, long nb, const float* c, long nc, float* out,
                         cudaStream_t st) {
  NSVD_CUDA(cudaMemsetAsync(out, 0, 3 * sizeof(float), st));
  AbsmaxList l{};
  l.p[0] = a; l.n[0] = na;
  l.p[1] = b; l.n[1] = nb;
  l.p[2] = c; l.n[2] = nc;
  l.count = c ? 3 : 2;
  long most = na > nb ? na : nb;
  if (nc > most) most = nc;
  const long want = cdiv(most, 256 * 16);
  absmax_list_kernel<<<(unsigned)(want < 148 * 4 ? (want > 1 ? want : 1) : 148 * 4), 256, 0, st>>>(l, out);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// D (M, N) = A (M, K) . B (N, K)^T from K-major fp16 hi/lo planes
static int linear_gemm_kmajor(const uint8_t* a_hi, const uint8_t* a_lo, const uint8_t* b_hi, const uint8_t* b_lo, int M,
                              int N, int K, const LinearEpi& epi, cudaStream_t st) {
  CUtensorMap mah, mal, mbh, mbl;
  int rc;
  if ((rc = make_tmap_bf16_3d(&mah, a_hi, K, M, 1, (uint64_t)K * 2, (uint64_t)M * K * 2, 64, big::BM))) return rc;
  if ((rc = make_tmap_bf16_3d(&mal, a_lo, K, M, 1, (uint64_t)K * 2, (uint64_t)M * K * 2, 64, big::BM))) return rc;
  if ((rc = make_tmap_bf16_3d(&mbh, b_hi, K, N, 1, (uint64_t)K * 2, (uint64_t)N * K * 2, 64, big::BN / 2))) return rc;
  if ((rc = make_tmap_bf16_3d(&mbl, b_lo, K, N, 1, (uint64_t)K * 2, (uint64_t)N * K * 2, 64, big::BN / 2))) return rc;
  BigShape s{};
  s.m_tiles = cdiv(M, 2 * big::BM);
  s.n_tiles = cdiv(N, big::BN);
  s.batches = 1;
  s.k_slices = 1;
  s.k_chunks_total = cdiv(K, big::BK);
  s.k_chunks_per_slice = s.k_chunks_total;
  s.m_group = 8;   // 2048 rows of A stay in the L2 while the column tiles of B sweep over them
  return launch_big2s<false, LinearEpi, kFmtHH>(mah, mal, mbh, mbl, s, 1, kLinearSub, kLinearSub, epi, st);
}

int tc_linear_fwd(const float* x, const float* W, const float* bias, float* y, int rows, int in_f, int out_f, int act,
                  float slope, void* work, size_t work_bytes, cudaStream_t st) {
  int rc;
  if ((rc = linear_check(rows, in_f, out_f, slope, work, work_bytes))) return rc;
  uint8_t* wk = align1k(work);
  float* amax = reinterpret_cast<float*>(wk);      // [0] x, [1] W
  wk += 1024;
  const size_t nx = linear_planes(rows, in_f), nw = linear_planes(out_f, in_f);
  uint8_t *x_hi = wk, *x_lo = wk + nx, *w_hi = wk + 2 * nx, *w_lo = wk + 2 * nx + nw;
  if ((rc = linear_absmax(x, (long)rows * in_f, W, (long)out_f * in_f, nullptr, 0, amax, st))) return rc;
  if ((rc = launch_split_scaled(x, amax, BF(x_hi), BF(x_lo), (long)rows * in_f, st))) return rc;
  if ((rc = launch_split_scaled(W, amax + 1, BF(w_hi), BF(w_lo), (long)out_f * in_f, st))) return rc;
  LinearEpi epi{y, bias, amax, amax + 1, rows, out_f, (long)out_f, 0, act, slope};
  return linear_gemm_kmajor(x_hi, x_lo, w_hi, w_lo, rows, out_f, in_f, epi, st);
}

int tc_linear_bwd(const float* x, const float* W, const float* y, const float* dy, int rows, int in_f, int out_f, int act,
                  float slope, float* dx, float* dW, float* db, void* work, size_t work_bytes, cudaStream_t st) {
  int rc;
  if ((rc = linear_check(rows, in_f, out_f, slope, work, work_bytes))) return rc;
  NSVD_CHECK_ARG(act == 0 || y != nullptr, "dense layer backward: the activation needs the layer output y");
  uint8_t* wk = align1k(work);
  float* amax = reinterpret_cast<float*>(wk);      // [0] dy (bounds dz), [1] W, [2] x
  wk += 1024;
  const size_t nz = linear_planes(rows, out_f), nw = linear_planes(out_f, in_f), nx = linear_planes(rows, in_f);
  uint8_t *z_hi = wk, *z_lo = wk + nz, *wT_hi = wk + 2 * nz, *wT_lo = wT_hi + nw, *x_hi = wT_lo + nw, *x_lo = x_hi + nx;
  if ((rc = linear_absmax(dy, (long)rows * out_f, W, (long)out_f * in_f, x, (long)rows * in_f, amax, st))) return rc;
  if (db) NSVD_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * out_f, st));
  linear_dz_kernel<<<dim3(cdiv(out_f, 256), cdiv(rows, 64)), 256, 0, st>>>(dy, y, act, slope, amax, BF(z_hi), BF(z_lo), db,
                                                                         rows, out_f);
  NSVD_LAUNCH_CHECK();
  if (dx) {   // dx (rows, in) = dz (rows, out) . W (out, in): B operand = W^T planes (in, out), K-major over out
    split_scaled_transposed_kernel<<<dim3(cdiv(in_f, 32), cdiv(out_f, 32)), 256, 0, st>>>(W, amax + 1, BF(wT_hi), BF(wT_lo),
                                                                                        out_f, in_f);
    NSVD_LAUNCH_CHECK();
    LinearEpi epi{dx, nullptr, amax, amax + 1, rows, in_f, (long)in_f, 0, 0, 0.f};
    if ((rc = linear_gemm_kmajor(z_hi, z_lo, wT_hi, wT_lo, rows, in_f, out_f, epi, st))) return rc;
  }
  if (dW) {   // dW (out, in) = dz^T . x: both operands MN-major (K = rows), out features in column blocks of 128
    if ((rc = launch_split_scaled(x, amax + 2, BF(x_hi), BF(x_lo), (long)rows * in_f, st))) return rc;
    CUtensorMap mah, mal, mbh, mbl;
    if ((rc = make_tmap_bf16_3d(&mah, z_hi, out_f, rows, 1, (uint64_t)out_f * 2, (uint64_t)rows * out_f * 2, 64, 64))) return rc;
    if ((rc = make_tmap_bf16_3d(&mal, z_lo, out_f, rows, 1, (uint64_t)out_f * 2, (uint64_t)rows * out_f * 2, 64, 64))) return rc;
    if ((rc = make_tmap_bf16_3d(&mbh, x_hi, in_f, rows, 1, (uint64_t)in_f * 2, (uint64_t)rows * in_f * 2, 64, 64))) return rc;
    if ((rc = make_tmap_bf16_3d(&mbl, x_lo, in_f, rows, 1, (uint64_t)in_f * 2, (uint64_t)rows * in_f * 2, 64, 64))) return rc;
    const int nb = cdiv(out_f, big::BM);
    BigShape s{};
    s.m_tiles = 1;
    s.n_tiles = cdiv(in_f, big::BN);
    s.batches = cdiv(nb, 2);
    s.k_slices = 1;
    s.k_chunks_total = cdiv(rows, big::BK);
    s.k_chunks_per_slice = s.k_chunks_total;
    s.a_batched = 1;
    s.b_batched = 0;
    s.a_xbatch = big::BM;
    LinearEpi epi{dW, nullptr, amax, amax + 2, out_f, in_f, (long)in_f, big::BM, 0, 0.f};
    if ((rc = launch_big2s<true, LinearEpi, kFmtHH>(mah, mal, mbh, mbl, s, nb, kLinearSub, kLinearSub, epi, st))) return rc;
  }
  return 0;
}

}  // namespace nsvd
