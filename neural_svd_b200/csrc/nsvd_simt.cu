// nsvd_simt.cu — fp32 CUDA-core engine of the NestedLoRA step + the engine-independent
// HBM-bound kernels (K2 gram_reduce, K3 loss_dF, CDK element-wise pieces).
//
// The fp32 engine is the validation-grade path (reference-level accuracy, used for tiny batches
// and to cross-check the tcgen05 engine on the device).  It follows the forward-mode restatement
// of SURVEY.md §8(a); reference file:line citations are in include/nsvd.h next to each entry.
#include "nsvd_common.cuh"
#include "nsvd_simt.cuh"

namespace nsvd {

// ------------------------------------------------------------------------------------------
// generic strided batched SGEMM:  C[b](m,n) (+)= sum_k A[b](m,k) * B[b](k,n)
// ------------------------------------------------------------------------------------------
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256) sgemm_strided_kernel(SGemm g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int bz = blockIdx.z;
  const float* __restrict__ A = g.A + (long)bz * g.a_bs;
  const float* __restrict__ B = g.B + (long)bz * g.b_bs;
  float* __restrict__ C = g.C + (long)bz * g.c_bs;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4] = {};
  const bool a_kc = (g.a_cs == 1), b_kc = (g.b_rs == 1);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; ++i) {
      int e = tid + i * 256;
      int m = a_kc ? e / BK : e % BM;
      int k = a_kc ? e % BK : e / BM;
      float v = 0.f;
      if (m0 + m < g.M && k0 + k < g.K) v = A[(long)(m0 + m) * g.a_rs + (long)(k0 + k) * g.a_cs];
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / 256; ++i) {
      int e = tid + i * 256;
      int n = b_kc ? e / BK : e % BN;
      int k = b_kc ? e % BK : e / BN;
      float v = 0.f;
      if (n0 + n < g.N && k0 + k < g.K) v = B[(long)(k0 + k) * g.b_rs + (long)(n0 + n) * g.b_cs];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][tm + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tn + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + tm + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tn + j;
      if (n >= g.N) continue;
      long idx = (long)m * g.c_rs + n;
      float v = g.alpha * acc[i][j];
      C[idx] = g.accumulate ? C[idx] + v : v;
    }
  }
}

int sgemm_strided(const SGemm& g, int batch, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || batch <= 0) return 0;
  dim3 grid(cdiv(g.N, 64), cdiv(g.M, 64), batch);
  sgemm_strided_kernel<64, 64, 16><<<grid, 256, 0, st>>>(g);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// column sums of a batched row-major matrix: out[b][n] (+)= sum_m X[b][m][n]   (deterministic)
__global__ void colsum_kernel(const float* __restrict__ X, float* __restrict__ out, int M, int N,
                              long x_bs, int accumulate) {
  // block = 32 columns x 8 row-lanes
  __shared__ float sm[8][33];
  const int b = blockIdx.y;
  const int n = blockIdx.x * 32 + threadIdx.x;
  const float* Xb = X + (long)b * x_bs;
  float s = 0.f;
  if (n < N)
    for (int m = threadIdx.y; m < M; m += 8) s += Xb[(long)m * N + n];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    float* o = out + (long)b * N + n;
    *o = accumulate ? *o + t : t;
  }
}

// sum of ONE vector (the CDK operator term: n_rows row dots), one block, fixed order: the general kernel above would
// leave 8 threads adding 512 values each with dependent loads (56 us at 4096 rows)
__global__ void __launch_bounds__(1024) vecsum_kernel(const float* __restrict__ X, float* __restrict__ out, int M,
                                                      int accumulate) {
  __shared__ float sw[32];
  float s = 0.f;
  for (int m = threadIdx.x; m < M; m += 1024) s += X[m];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sw[i];
    *out = accumulate ? *out + t : t;
  }
}

int colsum(const float* X, float* out, int M, int N, int batch, long x_bs, int accumulate,
           cudaStream_t st) {
  if (N == 1 && batch == 1) {
    vecsum_kernel<<<1, 1024, 0, st>>>(X, out, M, accumulate);
    NSVD_LAUNCH_CHECK();
    return 0;
  }
  dim3 grid(cdiv(N, 32), batch);
  colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(X, out, M, N, x_bs, accumulate);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Fourier features, all S = D + 2 streams (examples/utils.py:126-143 + derivatives)
//   phis[s][p][k], k in [0,2M): [sin | cos] blocks; s = 0 value, 1 .. D gradient, D + 1 Laplacian
// ------------------------------------------------------------------------------------------
__global__ void features_f32_kernel(const float* __restrict__ x, const float* __restrict__ Bff,
                                    float* __restrict__ phis, float* __restrict__ phi_saved, int P,
                                    int M, int D) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)P * M) return;
  int p = (int)(i / M), j = (int)(i % M);
  const float b[kMaxDim] = {Bff[j], Bff[M + j], D == 3 ? Bff[2L * M + j] : 0.f};
  const float b2 = fmaf(b[2], b[2], fmaf(b[1], b[1], b[0] * b[0]));
  float ph;
  if (D == 2) ph = fmaf(x[2 * p + 1], b[1], x[2 * p] * b[0]);
  else ph = fmaf(x[3L * p + 2], b[2], fmaf(x[3L * p + 1], b[1], x[3L * p] * b[0]));
  float s, c;
  sincosf(ph, &s, &c);
  const long K0 = 2L * M, SP = (long)P * K0;
  float* o = phis + (long)p * K0;
  o[j] = s;
  o[M + j] = c;
#pragma unroll
  for (int d = 0; d < kMaxDim; ++d) {
    if (d < D) {
      o[(1 + d) * SP + j] = c * b[d];
      o[(1 + d) * SP + M + j] = -s * b[d];
    }
  }
  o[(D + 1) * SP + j] = -b2 * s;
  o[(D + 1) * SP + M + j] = -b2 * c;
  if (phi_saved) {
    phi_saved[(long)p * K0 + j] = s;
    phi_saved[(long)p * K0 + M + j] = c;
  }
}

// softplus on the S-stream pre-activations Z[l][s][p][h] (in place) ; saves the value stream.
__global__ void softplus_streams_kernel(float* __restrict__ Z, const float* __restrict__ bias,
                                        float* __restrict__ a_saved, int L, int P, long Btot,
                                        long p_off, int D) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n = (long)L * P * kHidden;
  if (i >= n) return;
  int h = (int)(i % kHidden);
  int p = (int)((i / kHidden) % P);
  int l = (int)(i / ((long)kHidden * P));
  long SP = (long)P * kHidden;
  float* z = Z + (long)l * (D + 2) * SP + (long)p * kHidden + h;
  float z0 = z[0] + bias[l * kHidden + h];
  float a, sg;
  softplus_sig(z0, a, sg);
  z[0] = a;
  float q = 0.f;
  for (int d = 0; d < D; ++d) {
    const float zd = z[(1 + d) * SP];
    q = fmaf(zd, zd, q);
    z[(1 + d) * SP] = sg * zd;
  }
  z[(D + 1) * SP] = sg * z[(D + 1) * SP] + sg * (1.f - sg) * q;
  a_saved[((long)l * Btot + p_off + p) * kHidden + h] = a;
}

// last layer (128 -> 1) on the S streams + operator epilogue.  One warp per (point, copy).
__global__ void head_operator_kernel(const float* __restrict__ A2, const float* __restrict__ W3,
                                     const float* __restrict__ b3, const float* __restrict__ x,
                                     const float* __restrict__ mscales, nsvd_problem_t pb,
                                     float* __restrict__ F, float* __restrict__ TF,
                                     float* __restrict__ U0, int P, long p_off) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int L = pb.n_copies;
  if (w >= P * L) return;
  const int D = problem_ndim(pb), S = D + 2;
  int p = w / L, l = w % L;
  long SP = (long)P * kHidden;
  const float* a = A2 + (long)l * S * SP + (long)p * kHidden;
  float u[kMaxDim + 2] = {0, 0, 0, 0, 0};       // every index below is a compile-time constant (registers)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int h = lane + 32 * q;
    float wv = W3[l * kHidden + h];
#pragma unroll
    for (int s = 0; s < kMaxDim + 2; ++s)
      if (s < S) u[s] = fmaf(a[s * SP + h], wv, u[s]);
  }
#pragma unroll
  for (int s = 0; s < kMaxDim + 2; ++s) u[s] = warp_sum(u[s]);
  if (lane == 0) {
    u[0] += b3[l];
    long pg = p_off + p;
    PointGeomN g = point_geom_n(x + (long)D * pg, pb);
    float f, tf;
    operator_epilogue_n(g, pb, pb.has_exp_mask != 0, pb.has_exp_mask ? mscales[l] : 1.f, u, f, tf);
    F[pg * L + l] = f;
    TF[pg * L + l] = tf;
    U0[pg * L + l] = u[0];
  }
}

// ---- finite-difference mode (pb.fd_eps > 0): 2 D stream slots carry the 2 D SHIFTED point sets
// x + eps e_0, x - eps e_0, x + eps e_1, ... through the same contractions, value stream only.
__global__ void features_shift_f32_kernel(const float* __restrict__ x, const float* __restrict__ Bff,
                                          float* __restrict__ phis, int P, int M, float eps, int D) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)P * M) return;
  int p = (int)(i / M), j = (int)(i % M);
  const float b[kMaxDim] = {Bff[j], Bff[M + j], D == 3 ? Bff[2L * M + j] : 0.f};
  const float y[kMaxDim] = {x[(long)D * p], x[(long)D * p + 1], D == 3 ? x[3L * p + 2] : 0.f};
  const long K0 = 2L * M, SP = (long)P * K0;
#pragma unroll
  for (int s = 0; s < 2 * kMaxDim; ++s) {
    if (s >= 2 * D) break;
    float ph;
    if (D == 2) {
      float y0 = y[0], y1 = y[1];
      fd_shift(s, eps, y0, y1);
      ph = fmaf(y1, b[1], y0 * b[0]);
    } else {
      const float dlt = (s & 1) ? -eps : eps;
      const float y0 = (s >> 1) == 0 ? y[0] + dlt : y[0], y1 = (s >> 1) == 1 ? y[1] + dlt : y[1],
                  y2 = (s >> 1) == 2 ? y[2] + dlt : y[2];
      ph = fmaf(y2, b[2], fmaf(y1, b[1], y0 * b[0]));
    }
    float sn, cs;
    sincosf(ph, &sn, &cs);
    phis[s * SP + (long)p * K0 + j] = sn;
    phis[s * SP + (long)p * K0 + M + j] = cs;
  }
}
// plain softplus on all `slots` slots of Z[l][s][p][h] (in place)
__global__ void softplus_values_kernel(float* __restrict__ Z, const float* __restrict__ bias, int L, int P, int slots) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n = (long)L * slots * P * kHidden;
  if (i >= n) return;
  int h = (int)(i % kHidden);
  int l = (int)(i / ((long)kHidden * P * slots));
  float a, sg;
  softplus_sig(Z[i] + bias[l * kHidden + h], a, sg);
  Z[i] = a;
}
// last layer on the 2 D shifted slots + finite-difference operator; the central value comes from U0 (exact pass)
__global__ void head_fd_kernel(const float* __restrict__ A2, const float* __restrict__ W3,
                               const float* __restrict__ b3, const float* __restrict__ x,
                               const float* __restrict__ mscales, nsvd_problem_t pb,
                               const float* __restrict__ U0, float* __restrict__ TF, int P, long p_off) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int L = pb.n_copies;
  if (w >= P * L) return;
  const int D = problem_ndim(pb), S = 2 * D;
  int p = w / L, l = w % L;
  long SP = (long)P * kHidden;
  const float* a = A2 + (long)l * S * SP + (long)p * kHidden;
  float u[2 * kMaxDim] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int h = lane + 32 * q;
    float wv = W3[l * kHidden + h];
#pragma unroll
    for (int s = 0; s < 2 * kMaxDim; ++s)
      if (s < S) u[s] = fmaf(a[s * SP + h], wv, u[s]);
  }
#pragma unroll
  for (int s = 0; s < 2 * kMaxDim; ++s) u[s] = warp_sum(u[s]) + b3[l];
  if (lane == 0) {
    long pg = p_off + p;
    TF[pg * L + l] = fd_operator_n(x + (long)D * pg, pb, pb.has_exp_mask != 0,
                                   pb.has_exp_mask ? mscales[l] : 1.f, U0[pg * L + l], u);
  }
}

// backward of the head: du = dF * c * m * rho ; dZ2 = du W3 (.) sigma(a2) ; T3 = du * a2 (for dW3)
__global__ void head_bwd_kernel(const float* __restrict__ dF, const float* __restrict__ U0,
                                const float* __restrict__ a2, const float* __restrict__ W3,
                                const float* __restrict__ x, const float* __restrict__ mscales,
                                nsvd_problem_t pb, float* __restrict__ dZ2, float* __restrict__ T3,
                                float* __restrict__ du_out, float* __restrict__ ds_out, int P,
                                long Btot, long p_off) {
  int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int L = pb.n_copies;
  if (w >= P * L) return;
  int p = w / L, l = w % L;
  long pg = p_off + p;
  PointGeomN g = point_geom_n(x + (long)problem_ndim(pb) * pg, pb);
  float m = 1.f, sc = 1.f;
  if (pb.has_exp_mask) {
    sc = mscales[l];
    m = expf(-g.r / sc);
  }
  float cm = head_factor_n(g, pb, m);
  float df = dF[pg * L + l];
  float du = df * cm;
  if (lane == 0) {
    du_out[(long)p * L + l] = du;
    ds_out[(long)p * L + l] = pb.has_exp_mask ? du * U0[pg * L + l] * g.r / (sc * sc) : 0.f;
  }
  const float* a = a2 + ((long)l * Btot + pg) * kHidden;
  long o = ((long)l * P + p) * kHidden;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int h = lane + 32 * q;
    float av = a[h];
    dZ2[o + h] = du * W3[l * kHidden + h] * sig_from_softplus(av);
    T3[o + h] = du * av;
  }
}

// dZ = dA (.) sigma(a)   with a from the saved value stream (layout [L][Btot][H])
__global__ void dact_kernel(float* __restrict__ dA, const float* __restrict__ a_saved, int L, int P,
                            long Btot, long p_off) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n = (long)L * P * kHidden;
  if (i >= n) return;
  int h = (int)(i % kHidden);
  int p = (int)((i / kHidden) % P);
  int l = (int)(i / ((long)kHidden * P));
  float a = a_saved[((long)l * Btot + p_off + p) * kHidden + h];
  dA[i] *= sig_from_softplus(a);
}

// ------------------------------------------------------------------------------------------
// fp32 engine: forward and backward drivers.  `slots` = stream slots of the work buffers: S = D + 2 in the exact
// pass, 2 D shifted point sets in the finite-difference pass (D = 3: 5 and 6).
// ------------------------------------------------------------------------------------------
static const int kSimtMicroBatch = 2048;
static inline int simt_slots(const nsvd_problem_t& pb) {
  const int D = problem_ndim(pb);
  return (pb.fd_eps > 0.f && 2 * D > D + 2) ? 2 * D : D + 2;
}

void simt_scratch_bytes(const nsvd_problem_t& pb, size_t* saved, size_t* work) {
  long B = pb.n_points, L = pb.n_copies, K0 = 2L * pb.n_fourier;
  long P = B < kSimtMicroBatch ? B : kSimtMicroBatch;
  const long NS = simt_slots(pb);
  *saved = sizeof(float) * (size_t)(B * K0 + 3 * L * B * kHidden + B * L) + 256;
  size_t fwd = (size_t)(NS * P * K0 + 2 * NS * L * P * kHidden);
  size_t bwd = (size_t)(3 * L * P * kHidden + 2 * P * L);
  *work = sizeof(float) * (fwd > bwd ? fwd : bwd) + 256;
}

struct SimtSaved {
  float *phi, *a[3], *u0;
};
static SimtSaved carve_saved(const nsvd_problem_t& pb, void* saved) {
  long B = pb.n_points, L = pb.n_copies, K0 = 2L * pb.n_fourier;
  SimtSaved s;
  float* p = (float*)saved;
  s.phi = p;
  p += B * K0;
  for (int i = 0; i < 3; ++i) {
    s.a[i] = p;
    p += L * B * kHidden;
  }
  s.u0 = p;
  return s;
}

// the three dense layers on `slots` stacked (slots x P)-row operands; `streams` selects the activation
static int simt_layers(const nsvd_params_t& pr, const SimtSaved& sv, float* phis, float* bufA, float* bufB, int slots,
                       bool streams, int D, long L, int P, long B, long p0, long K0, float** out, cudaStream_t st) {
  SGemm g{};
  g.A = phis; g.a_rs = K0; g.a_cs = 1; g.a_bs = 0;
  g.B = pr.W[0]; g.b_rs = 1; g.b_cs = K0; g.b_bs = (long)kHidden * K0;
  g.C = bufA; g.c_rs = kHidden; g.c_bs = (long)slots * P * kHidden;
  g.M = slots * P; g.N = kHidden; g.K = (int)K0; g.alpha = 1.f; g.accumulate = 0;
  int rc = sgemm_strided(g, (int)L, st);
  if (rc) return rc;
  const long ne = L * P * kHidden, nall = ne * slots;
  if (streams) softplus_streams_kernel<<<cdiv(ne, 256), 256, 0, st>>>(bufA, pr.b[0], sv.a[0], (int)L, P, B, p0, D);
  else softplus_values_kernel<<<cdiv(nall, 256), 256, 0, st>>>(bufA, pr.b[0], (int)L, P, slots);
  NSVD_LAUNCH_CHECK();
  float* cur = bufA;
  float* nxt = bufB;
  for (int i = 1; i <= 2; ++i) {
    SGemm h{};
    h.A = cur; h.a_rs = kHidden; h.a_cs = 1; h.a_bs = (long)slots * P * kHidden;
    h.B = pr.W[i]; h.b_rs = 1; h.b_cs = kHidden; h.b_bs = (long)kHidden * kHidden;
    h.C = nxt; h.c_rs = kHidden; h.c_bs = (long)slots * P * kHidden;
    h.M = slots * P; h.N = kHidden; h.K = kHidden; h.alpha = 1.f; h.accumulate = 0;
    if ((rc = sgemm_strided(h, (int)L, st))) return rc;
    if (streams) softplus_streams_kernel<<<cdiv(ne, 256), 256, 0, st>>>(nxt, pr.b[i], sv.a[i], (int)L, P, B, p0, D);
    else softplus_values_kernel<<<cdiv(nall, 256), 256, 0, st>>>(nxt, pr.b[i], (int)L, P, slots);
    NSVD_LAUNCH_CHECK();
    float* t = cur; cur = nxt; nxt = t;
  }
  *out = cur;
  return 0;
}

int simt_forward(const nsvd_problem_t& pb, const nsvd_params_t& pr, const float* x, float* F,
                 float* TF, void* saved, void* work, cudaStream_t st) {
  const long B = pb.n_points, L = pb.n_copies, M = pb.n_fourier, K0 = 2 * M;
  const int D = problem_ndim(pb), S = D + 2, NS = simt_slots(pb);
  SimtSaved sv = carve_saved(pb, saved);
  for (long p0 = 0; p0 < B; p0 += kSimtMicroBatch) {
    int P = (int)((B - p0) < kSimtMicroBatch ? (B - p0) : kSimtMicroBatch);
    float* phis = (float*)work;
    float* bufA = phis + (long)NS * P * K0;
    float* bufB = bufA + (long)NS * L * P * kHidden;
    long n = (long)P * M;
    features_f32_kernel<<<cdiv(n, 256), 256, 0, st>>>(x + D * p0, pr.Bff, phis, sv.phi + p0 * K0, P, (int)M, D);
    NSVD_LAUNCH_CHECK();
    float* cur = nullptr;
    int rc = simt_layers(pr, sv, phis, bufA, bufB, S, true, D, L, P, B, p0, K0, &cur, st);
    if (rc) return rc;
    long nw = (long)P * L;
    head_operator_kernel<<<cdiv(nw * 32, 256), 256, 0, st>>>(cur, pr.W[3], pr.b[3], x, pr.mask_scales,
                                                             pb, F, TF, sv.u0, P, p0);
    NSVD_LAUNCH_CHECK();
    if (pb.fd_eps > 0.f) {
      // finite-difference Laplacian: second pass, the 2 D slots = the shifted point sets, values only;
      // TF is overwritten, F / U0 / the saved activations of the central pass stay (the backward uses them)
      features_shift_f32_kernel<<<cdiv(n, 256), 256, 0, st>>>(x + D * p0, pr.Bff, phis, P, (int)M, pb.fd_eps, D);
      NSVD_LAUNCH_CHECK();
      if ((rc = simt_layers(pr, sv, phis, bufA, bufB, 2 * D, false, D, L, P, B, p0, K0, &cur, st))) return rc;
      head_fd_kernel<<<cdiv(nw * 32, 256), 256, 0, st>>>(cur, pr.W[3], pr.b[3], x, pr.mask_scales, pb, sv.u0, TF, P, p0);
      NSVD_LAUNCH_CHECK();
    }
  }
  return 0;
}

int simt_backward(const nsvd_problem_t& pb, const nsvd_params_t& pr, const float* x, const float* dF,
                  const void* saved, nsvd_grads_t& gr, void* work, cudaStream_t st) {
  const long B = pb.n_points, L = pb.n_copies, M = pb.n_fourier, K0 = 2 * M;
  SimtSaved sv = carve_saved(pb, const_cast<void*>(saved));
  for (long p0 = 0; p0 < B; p0 += kSimtMicroBatch) {
    int P = (int)((B - p0) < kSimtMicroBatch ? (B - p0) : kSimtMicroBatch);
    int acc = p0 > 0;
    float* dZ = (float*)work;
    float* dN = dZ + L * P * kHidden;
    float* T3 = dN + L * P * kHidden;
    float* du = T3 + L * P * kHidden;
    float* ds = du + (long)P * L;
    long nw = (long)P * L;
    head_bwd_kernel<<<cdiv(nw * 32, 256), 256, 0, st>>>(dF, sv.u0, sv.a[2], pr.W[3], x,
                                                        pr.mask_scales, pb, dZ, T3, du, ds, P, B, p0);
    NSVD_LAUNCH_CHECK();
    int rc;
    // dW3[l][h] = sum_p T3[l][p][h];  db3[l] = sum_p du[p][l];  dscales[l] = sum_p ds[p][l]
    if ((rc = colsum(T3, gr.dW[3], P, kHidden, (int)L, (long)P * kHidden, acc, st))) return rc;
    if ((rc = colsum(du, gr.db[3], P, (int)L, 1, 0, acc, st))) return rc;
    if (pb.has_exp_mask && gr.dmask_scales)
      if ((rc = colsum(ds, gr.dmask_scales, P, (int)L, 1, 0, acc, st))) return rc;
    for (int i = 2; i >= 1; --i) {
      // dW_i[l] (H x H) (+)= dZ[l]^T (H x P) . a_{i-1}[l] (P x H)
      SGemm w{};
      w.A = dZ; w.a_rs = 1; w.a_cs = kHidden; w.a_bs = (long)P * kHidden;
      w.B = sv.a[i - 1] + p0 * kHidden; w.b_rs = kHidden; w.b_cs = 1; w.b_bs = B * kHidden;
      w.C = gr.dW[i]; w.c_rs = kHidden; w.c_bs = (long)kHidden * kHidden;
      w.M = kHidden; w.N = kHidden; w.K = P; w.alpha = 1.f; w.accumulate = acc;
      if ((rc = sgemm_strided(w, (int)L, st))) return rc;
      if ((rc = colsum(dZ, gr.db[i], P, kHidden, (int)L, (long)P * kHidden, acc, st))) return rc;
      // dA[l] (P x H) = dZ[l] (P x H) . W_i[l] (H x H)
      SGemm d{};
      d.A = dZ; d.a_rs = kHidden; d.a_cs = 1; d.a_bs = (long)P * kHidden;
      d.B = pr.W[i]; d.b_rs = kHidden; d.b_cs = 1; d.b_bs = (long)kHidden * kHidden;
      d.C = dN; d.c_rs = kHidden; d.c_bs = (long)P * kHidden;
      d.M = P; d.N = kHidden; d.K = kHidden; d.alpha = 1.f; d.accumulate = 0;
      if ((rc = sgemm_strided(d, (int)L, st))) return rc;
      long ne = L * P * kHidden;
      dact_kernel<<<cdiv(ne, 256), 256, 0, st>>>(dN, sv.a[i - 1], (int)L, P, B, p0);
      NSVD_LAUNCH_CHECK();
      float* t = dZ; dZ = dN; dN = t;
    }
    // dW0[l] (H x K0) (+)= dZ0[l]^T (H x P) . Phi (P x K0)
    SGemm w{};
    w.A = dZ; w.a_rs = 1; w.a_cs = kHidden; w.a_bs = (long)P * kHidden;
    w.B = sv.phi + p0 * K0; w.b_rs = K0; w.b_cs = 1; w.b_bs = 0;
    w.C = gr.dW[0]; w.c_rs = K0; w.c_bs = (long)kHidden * K0;
    w.M = kHidden; w.N = (int)K0; w.K = P; w.alpha = 1.f; w.accumulate = acc;
    if ((rc = sgemm_strided(w, (int)L, st))) return rc;
    if ((rc = colsum(dZ, gr.db[0], P, kHidden, (int)L, (long)P * kHidden, acc, st))) return rc;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// K2 gram_reduce  (HBM-bound): two-stage deterministic reduction.
//   stage 1: each block owns a contiguous row range inside ONE half and accumulates the LxL
//            Gram of F (and optionally the cross Gram F^T TF) in registers: thread (i,j) holds a
//            TI x TJ sub-block; rows are staged through shared memory with float4 loads.
//   stage 2: one block per output element group sums the per-block partials in fixed order.
// ------------------------------------------------------------------------------------------
constexpr int kGramRows = 64;  // rows staged per iteration

__device__ __forceinline__ float nan_to_num_f(float v) {
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return v;
}

template <int LT>  // LT = L rounded up to a multiple of 16 (16, 32, 48, 64) ; threads = 256
__global__ void __launch_bounds__(256)
gram_stage1_kernel(const float* __restrict__ F, const float* __restrict__ TF,
                   const float* __restrict__ vmask, const float* __restrict__ roww,
                   const float* __restrict__ xrow, int L,
                   long row_begin, long row_end, int rows_per_block, int cross,
                   float* __restrict__ partials, int partial_stride, int block_off) {
  // thread tile: (LT/16) x (LT/16) outputs; 16x16 threads
  constexpr int T = LT / 16;
  __shared__ float sF[kGramRows][LT + 1];
  __shared__ float sT[kGramRows][LT + 1];
  __shared__ float sred[8];
  const int tid = threadIdx.x, ti = tid / 16, tj = tid % 16;
  long r0 = row_begin + (long)blockIdx.x * rows_per_block;
  long r1 = r0 + rows_per_block < row_end ? r0 + rows_per_block : row_end;
  float acc[T][T] = {};
  float accx[T][T] = {};
  float ops = 0.f;
  for (long rb = r0; rb < r1; rb += kGramRows) {
    int nr = (int)((r1 - rb) < kGramRows ? (r1 - rb) : kGramRows);
    for (int e = tid; e < kGramRows * LT; e += 256) {
      int rr = e / LT, c = e % LT;
      float f = 0.f, t = 0.f;
      if (rr < nr && c < L) {
        long idx = (rb + rr) * L + c;
        float w = roww ? roww[rb + rr] : 1.f;
        f = F[idx] * w;
        t = TF[idx] * w;
        if (cross) {  // torch.nan_to_num + zero the T-row at the origin (methods/spectrum.py:71-73)
          f = nan_to_num_f(f);
          t = nan_to_num_f(t);
          if (xrow && fabsf(xrow[2 * (rb + rr)]) <= 1e-8f && fabsf(xrow[2 * (rb + rr) + 1]) <= 1e-8f) t = 0.f;
        }
        if (!cross) ops += vmask[c] * f * t;
      }
      sF[rr][c] = f;
      sT[rr][c] = t;
    }
    __syncthreads();
    for (int rr = 0; rr < nr; ++rr) {
      float a[T], b[T], bx[T];
#pragma unroll
      for (int i = 0; i < T; ++i) a[i] = sF[rr][ti * T + i];
#pragma unroll
      for (int j = 0; j < T; ++j) {
        b[j] = sF[rr][tj * T + j];
        bx[j] = sT[rr][tj * T + j];
      }
#pragma unroll
      for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) {
          acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
          if (cross) accx[i][j] = fmaf(a[i], bx[j], accx[i][j]);
        }
    }
    __syncthreads();
  }
  float* out = partials + (long)(block_off + blockIdx.x) * partial_stride;
#pragma unroll
  for (int i = 0; i < T; ++i)
#pragma unroll
    for (int j = 0; j < T; ++j) {
      int gi = ti * T + i, gj = tj * T + j;
      if (gi < L && gj < L) {
        out[gi * L + gj] = acc[i][j];
        if (cross) out[L * L + gi * L + gj] = accx[i][j];
      }
    }
  if (!cross) {
    ops = warp_sum(ops);
    if ((tid & 31) == 0) sred[tid >> 5] = ops;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += sred[i];
      out[L * L] = t;
    }
  }
}

// L == 16 specialisation (the benchmark shape): ONE launch for the whole K2.
//   warp-cooperative register tiling: lane = (ib, jb) owns the 4 x 2 block G[4 ib + a][2 jb + c].  A warp loads row
//   pairs with one coalesced 128-byte load per array (8 pairs per iteration, the next iteration's loads in flight while
//   the current one is consumed), parks the F pairs in its private 1 KB of shared memory and reads its Gram operands
//   back as ONE 16-byte and ONE 8-byte broadcast load per row (4 + 8 distinct addresses inside 64 contiguous bytes:
//   a single wavefront each) - the first version fetched them with 6 shuffles per row and was bound by the shuffle
//   rate (profiles/README.md).  The operator sum is taken on the coalesced registers.  Blocks [0, nb1) cover the first
//   half of the rows, the rest the second half.  Each block writes a 257-float partial; the last block to finish
//   (ticket counter) adds the partials in block order, so the result is deterministic.
__global__ void __launch_bounds__(1024)
gram16_fused_kernel(const float* __restrict__ F, const float* __restrict__ TF, const float* __restrict__ vmask,
                    long B, long b1, int rows_per_block, int nb1, float* __restrict__ partials,
                    unsigned int* __restrict__ counter, float* __restrict__ terms) {
  __shared__ __align__(16) float sacc[32][257];
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = (lane >> 3) * 4, j0 = (lane & 7) * 2;
  const bool second = (int)blockIdx.x >= nb1;
  const long hb = second ? b1 : 0, he = second ? B : b1;
  long r0 = hb + (long)(second ? blockIdx.x - nb1 : blockIdx.x) * rows_per_block;
  long r1 = r0 + rows_per_block < he ? r0 + rows_per_block : he;
  float acc[4][2] = {};
  float ops = 0.f;
  const float vm = vmask[lane & 15];
  constexpr int U = 8;   // row pairs per iteration: 2 x 8 x 128 B per warp and array, ~128 KB per SM in flight
  float* stage = &sacc[0][0] + warp * (U * 32);   // this warp's 8 row pairs (16-byte aligned; sacc proper is used at the end)
  float fo[U], to[U], fn[U], tn[U];
  auto load8 = [&](long r, float* f8, float* t8) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long rr = r + 64 * u + (lane >> 4);
      const bool okr = rr < r1;
      f8[u] = okr ? F[rr * 16 + (lane & 15)] : 0.f;
      t8[u] = okr ? TF[rr * 16 + (lane & 15)] : 0.f;
    }
  };
  long r = r0 + 2 * warp;
  load8(r, fo, to);
  for (; r < r1; r += 64 * U) {
    load8(r + 64 * U, fn, tn);   // rows beyond r1 load as zeros
#pragma unroll
    for (int u = 0; u < U; ++u) {
      stage[u * 32 + lane] = fo[u];
      ops = fmaf(vm * fo[u], to[u], ops);
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(stage + u * 32 + 16 * k + i0);
        const float2 c = *reinterpret_cast<const float2*>(stage + u * 32 + 16 * k + j0);
        acc[0][0] = fmaf(a.x, c.x, acc[0][0]);
        acc[0][1] = fmaf(a.x, c.y, acc[0][1]);
        acc[1][0] = fmaf(a.y, c.x, acc[1][0]);
        acc[1][1] = fmaf(a.y, c.y, acc[1][1]);
        acc[2][0] = fmaf(a.z, c.x, acc[2][0]);
        acc[2][1] = fmaf(a.z, c.y, acc[2][1]);
        acc[3][0] = fmaf(a.w, c.x, acc[3][0]);
        acc[3][1] = fmaf(a.w, c.y, acc[3][1]);
      }
    }
    __syncwarp();                // the next iteration overwrites the staged pairs
#pragma unroll
    for (int u = 0; u < U; ++u) {
      fo[u] = fn[u];
      to[u] = tn[u];
    }
  }
  __syncthreads();               // every warp is done with its staging area before sacc is reused for the block sum
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) sacc[warp][(i0 + a) * 16 + j0 + c] = acc[a][c];
  ops = warp_sum(ops);
  if (lane == 0) sacc[warp][256] = ops;
  __syncthreads();
  float* out = partials + (long)blockIdx.x * 257;
  if (tid < 257) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += sacc[w][tid];
    out[tid] = t;
  }
  // ---- the last block to finish adds all partials (fixed order => deterministic result)
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int nb = gridDim.x;
  // both halves at once: warps 0-15 add the partials of the first half, warps 16-31 those of the second, warp w taking
  // rows w, w + 16, ... of its half in block order with 5 rows (40 loads per lane) in flight; lane handles columns
  // lane + 32 c.  The 16 warp sums of a half are then added in warp order - a fixed order, so the result is deterministic.
  {
    const int halfsel = warp >> 4, w16 = warp & 15;
    const int pb = halfsel ? nb1 : 0, pe = halfsel ? nb : nb1;
    float a8[8] = {}, a_ops = 0.f;
    for (int bb = pb + w16; bb < pe; bb += 16 * 5) {
      float v[5][8], vo[5];
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const int b2 = bb + 16 * u;
        const bool ok = b2 < pe;
        const float* prow = partials + (long)(ok ? b2 : pb) * 257;
#pragma unroll
        for (int c = 0; c < 8; ++c) v[u][c] = ok ? __ldcg(prow + lane + 32 * c) : 0.f;
        vo[u] = (ok && lane == 0) ? __ldcg(prow + 256) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 5; ++u) {
#pragma unroll
        for (int c = 0; c < 8; ++c) a8[c] += v[u][c];
        a_ops += vo[u];
      }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) sacc[warp][lane + 32 * c] = a8[c];
    if (lane == 0) sacc[warp][256] = a_ops;
    __syncthreads();
    if (tid < 512) {
      const int hs = tid >> 8, col = tid & 255;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 16; ++w) t += sacc[hs * 16 + w][col];
      terms[hs * 256 + col] = t;
    }
    if (tid == 512) {
      float o = 0.f;
#pragma unroll
      for (int w = 0; w < 32; ++w) o += sacc[w][256];
      terms[512] = o;
    }
  }
  if (tid == 0) *counter = 0u;
}

// ------------------------------------------------------------------------------------------
// K2 for 16 < L <= 64 on the warp-level tensor cores (mma.sync m16n8k8, 3xTF32): at L = 64 the Gram is 32 FLOP/B,
// past the CUDA-core ridge.  A block of 4 warps walks its rows in chunks of 32: the chunk (a contiguous
// run of 32 L floats) is loaded coalesced into shared memory (row stride LP + 8 floats: fragment reads hit 32
// different banks), every warp owns one 16-row band of the LP x LP Gram (LP = 64) or half of one (LP = 32).
// v = hi + lo with hi = tf32(v), lo = tf32(v - hi): lo*hi + hi*lo + hi*hi is exact to 2^-22.  A chunk is accumulated
// from zero (12 MMAs) and then added to the running sums with ordinary fp32 additions, which keeps the tensor-core
// accumulate (it truncates) out of the long sum.  sum_l v_l f Tf rides on the load loop.  Partials have the layout of
// gram_stage1_kernel (cross = 0), so stage 2 is shared and the result is deterministic.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int LP>  // 32 or 64
__global__ void __launch_bounds__(128)
gram_mma_kernel(const float* __restrict__ F, const float* __restrict__ TF, const float* __restrict__ vmask, int L,
                long row_begin, long row_end, int rows_per_block, float* __restrict__ partials, int partial_stride,
                int block_off) {
  constexpr int TM = LP / 16, TN = LP / 8;      // Gram tiles: 16 rows x 8 columns each
  constexpr int WPM = 4 / TM;                   // warps sharing one 16-row band
  constexpr int NT = TN / WPM;                  // column tiles per warp
  constexpr int SS = LP + 8, CH = 32;
  __shared__ float S[CH][SS];
  __shared__ float s_ops[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int mt = warp % TM, n0 = (warp / TM) * NT;
  const long r0 = row_begin + (long)blockIdx.x * rows_per_block;
  const long r1 = r0 + rows_per_block < row_end ? r0 + rows_per_block : row_end;
  float tot[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) tot[i][j] = 0.f;
  float ops = 0.f;
  for (int i = tid; i < CH * SS; i += 128) (&S[0][0])[i] = 0.f;   // columns >= L stay zero
  __syncthreads();
  const int dq = 128 / L, dr = 128 % L;         // (row, col) advance of a thread between its elements
  for (long c0 = r0; c0 < r1; c0 += CH) {
    const int nrows = (int)((r1 - c0) < CH ? (r1 - c0) : CH);
    const long base = c0 * L;
    const int nel = nrows * L;
    int row = tid / L, col = tid % L;
    for (int e = tid; e < CH * L; e += 128) {
      float f = 0.f;
      if (e < nel) {
        f = F[base + e];
        ops = fmaf(vmask[col] * f, TF[base + e], ops);
      }
      S[row][col] = f;
      row += dq;
      col += dr;
      if (col >= L) {
        col -= L;
        ++row;
      }
    }
    __syncthreads();
    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < CH / 8; ++ks) {
      const float* s0 = &S[ks * 8 + t][0];
      const float* s1 = &S[ks * 8 + t + 4][0];
      // A = F^T (M = Gram row i, K = data row): a0 (g, t), a1 (g + 8, t), a2 (g, t + 4), a3 (g + 8, t + 4)
      const float av[4] = {s0[16 * mt + g], s0[16 * mt + g + 8], s1[16 * mt + g], s1[16 * mt + g + 8]};
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ah[i] = to_tf32(av[i]);
        al[i] = to_tf32(av[i] - __uint_as_float(ah[i]));
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        // B (K = data row, N = Gram column j): b0 (k = t, n = g), b1 (k = t + 4, n = g)
        const float bv[2] = {s0[8 * (n0 + nt) + g], s1[8 * (n0 + nt) + g]};
        uint32_t bh[2], bl[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          bh[i] = to_tf32(bv[i]);
          bl[i] = to_tf32(bv[i] - __uint_as_float(bh[i]));
        }
        mma_tf32_16x8x8(acc[nt], al, bh);
        mma_tf32_16x8x8(acc[nt], ah, bl);
        mma_tf32_16x8x8(acc[nt], ah, bh);
      }
    }
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) tot[i][j] += acc[i][j];
    __syncthreads();
  }
  // C fragment: c0 (g, 2t), c1 (g, 2t + 1), c2 (g + 8, 2t), c3 (g + 8, 2t + 1) inside tile (mt, n0 + nt)
  float* out = partials + (long)(block_off + blockIdx.x) * partial_stride;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i_ = 16 * mt + g + (j >> 1) * 8, j_ = 8 * (n0 + nt) + 2 * t + (j & 1);
      if (i_ < L && j_ < L) out[i_ * L + j_] = tot[nt][j];
    }
  }
  ops = warp_sum(ops);
  if (lane == 0) s_ops[warp] = ops;
  __syncthreads();
  if (tid == 0) out[L * L] = (s_ops[0] + s_ops[1]) + (s_ops[2] + s_ops[3]);   // slot read by stage 2 (cross = 0)
}

// Second version of the tensor-core K2 for 16 < L <= 64, L a multiple of 4 (the config-4 shape L = 64): the kernel above
// loads a 32-row chunk, synchronises, multiplies, synchronises - nothing is in flight while it multiplies - and every
// fragment element is split into its two TF32 parts again by every warp that uses it: 1.4 TB/s.  Here a block reads a
// chunk with 16-byte loads into REGISTERS one chunk ahead (the loads of chunk c + 1 are in flight while chunk c is
// multiplied) and splits every value ONCE on its way into two shared-memory planes (tf32 hi / lo).  The m16n8k8 shape
// needs 12 fragment words per 3 products, so the multiply is bound by shared-memory loads unless fragments are reused in
// registers: a warp owns WM x WN Gram tiles (2 x 4 at LP = 64: 32 fragment loads per 24 MMAs; one tile per warp, the
// first attempt, needs 36 per 12 and ran at 1.9 TB/s).  Chunk sums leave the tensor-core accumulator after 32 rows and
// are added in fp32 registers, as before.  Blocks of 4 warps, several per SM.
template <int LP>
__global__ void __launch_bounds__(128)
gram_mma2_kernel(const float* __restrict__ F, const float* __restrict__ TF, const float* __restrict__ vmask, int L,
                 long row_begin, long row_end, int rows_per_block, float* __restrict__ partials, int partial_stride,
                 int block_off) {
  constexpr int TM = LP / 16, TN = LP / 8;      // Gram tiles: 16 rows x 8 columns each
  constexpr int WM = TM / 2, WN = TN / 2;       // tiles per warp; the 4 warps form a 2 x 2 grid
  constexpr int SS = LP + 8, CH = 32;           // row stride of the planes: fragment loads are bank-conflict free
  constexpr int NV = (CH * LP / 4 + 127) / 128; // float4 per thread and chunk
  __shared__ __align__(16) uint32_t Sh[CH][SS];
  __shared__ __align__(16) uint32_t Sl[CH][SS];
  __shared__ float s_ops[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp >> 1) * WM, n0 = (warp & 1) * WN;
  const long r0 = row_begin + (long)blockIdx.x * rows_per_block;
  const long r1 = r0 + rows_per_block < row_end ? r0 + rows_per_block : row_end;
  float tot[WM][WN][4];
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) tot[i][j][k] = 0.f;
  float ops = 0.f;
  for (int i = tid; i < CH * SS; i += 128) {    // columns >= L stay zero
    (&Sh[0][0])[i] = 0u;
    (&Sl[0][0])[i] = 0u;
  }
  // a chunk is CH * L contiguous floats = 8 L float4: thread owns float4 tid + 128 u of it
  const int q4 = 8 * L;
  float4 fr[NV], tr[NV];
  auto load_chunk = [&](long c0) {
    const long nel4 = ((r1 - c0) < CH ? (r1 - c0) : CH) * (long)(L >> 2);
    const float4* f4 = reinterpret_cast<const float4*>(F + c0 * L);
    const float4* t4 = reinterpret_cast<const float4*>(TF + c0 * L);
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int e = tid + 128 * u;
      const bool ok = e < q4 && e < nel4;
      fr[u] = ok ? __ldcs(f4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      tr[u] = ok ? __ldcs(t4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (r0 < r1) load_chunk(r0);
  __syncthreads();
  for (long c0 = r0; c0 < r1; c0 += CH) {
    // registers -> operator sum + the two tf32 planes
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      const int e = tid + 128 * u;
      if (e < q4) {
        const int row = (4 * e) / L, col = (4 * e) % L;
        const float4 v4 = *reinterpret_cast<const float4*>(vmask + col);
        ops = fmaf(v4.x * fr[u].x, tr[u].x, ops);
        ops = fmaf(v4.y * fr[u].y, tr[u].y, ops);
        ops = fmaf(v4.z * fr[u].z, tr[u].z, ops);
        ops = fmaf(v4.w * fr[u].w, tr[u].w, ops);
        uint4 h, l;
        h.x = to_tf32(fr[u].x); l.x = to_tf32(fr[u].x - __uint_as_float(h.x));
        h.y = to_tf32(fr[u].y); l.y = to_tf32(fr[u].y - __uint_as_float(h.y));
        h.z = to_tf32(fr[u].z); l.z = to_tf32(fr[u].z - __uint_as_float(h.z));
        h.w = to_tf32(fr[u].w); l.w = to_tf32(fr[u].w - __uint_as_float(h.w));
        *reinterpret_cast<uint4*>(&Sh[row][col]) = h;
        *reinterpret_cast<uint4*>(&Sl[row][col]) = l;
      }
    }
    __syncthreads();
    if (c0 + CH < r1) load_chunk(c0 + CH);      // in flight while this chunk is multiplied
    float acc[WM][WN][4];
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
#pragma unroll
    for (int ks = 0; ks < CH / 8; ++ks) {
      const int k0 = ks * 8 + t;
      // B (K = data row, N = Gram column): b0 (k = t, n = g), b1 (k = t + 4, n = g)
      uint32_t bh[WN][2], bl[WN][2];
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        const int c = 8 * (n0 + j) + g;
        bh[j][0] = Sh[k0][c]; bh[j][1] = Sh[k0 + 4][c];
        bl[j][0] = Sl[k0][c]; bl[j][1] = Sl[k0 + 4][c];
      }
#pragma unroll
      for (int i = 0; i < WM; ++i) {
        // A = F^T (M = Gram row, K = data row): a0 (g, t), a1 (g + 8, t), a2 (g, t + 4), a3 (g + 8, t + 4)
        const int c = 16 * (m0 + i) + g;
        const uint32_t ah[4] = {Sh[k0][c], Sh[k0][c + 8], Sh[k0 + 4][c], Sh[k0 + 4][c + 8]};
        const uint32_t al[4] = {Sl[k0][c], Sl[k0][c + 8], Sl[k0 + 4][c], Sl[k0 + 4][c + 8]};
#pragma unroll
        for (int j = 0; j < WN; ++j) {
          mma_tf32_16x8x8(acc[i][j], al, bh[j]);
          mma_tf32_16x8x8(acc[i][j], ah, bl[j]);
          mma_tf32_16x8x8(acc[i][j], ah, bh[j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) tot[i][j][k] += acc[i][j][k];
    __syncthreads();                            // the planes are rewritten by the next chunk
  }
  // C fragment: c0 (g, 2t), c1 (g, 2t + 1), c2 (g + 8, 2t), c3 (g + 8, 2t + 1) inside tile (mt, nt)
  float* out = partials + (long)(block_off + blockIdx.x) * partial_stride;
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i_ = 16 * (m0 + i) + g + (k >> 1) * 8, j_ = 8 * (n0 + j) + 2 * t + (k & 1);
        if (i_ < L && j_ < L) out[i_ * L + j_] = tot[i][j][k];
      }
  ops = warp_sum(ops);
  if (lane == 0) s_ops[warp] = ops;
  __syncthreads();
  if (tid == 0) out[L * L] = (s_ops[0] + s_ops[1]) + (s_ops[2] + s_ops[3]);   // slot read by stage 2 (cross = 0)
}

// stage 2: out[e] (+)= sum over blocks [b0, b1) of partials[b][src_off + e], in a fixed order (deterministic).
// A block of 8 warps owns 32 consecutive elements; warp w adds the partials b0 + w, b0 + w + 8, ... with four loads in
// flight, the 8 warp sums are then added in warp order.  (The first version gave one thread the whole loop over up to
// 592 partials - a chain of dependent-latency loads, ~70 us per launch at L = 64: half of the measured K2 time.)
__global__ void __launch_bounds__(256)
gram_stage2_kernel(const float* __restrict__ partials, int partial_stride, int b0, int b1, int src_off, int n,
                   float* __restrict__ out, int accumulate) {
  __shared__ float sw[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (e < n) {
    const float* p = partials + src_off + e;
    int b = b0 + warp;
    for (; b + 24 < b1; b += 32) {
      const float v0 = __ldcg(p + (long)b * partial_stride), v1 = __ldcg(p + (long)(b + 8) * partial_stride),
                  v2 = __ldcg(p + (long)(b + 16) * partial_stride), v3 = __ldcg(p + (long)(b + 24) * partial_stride);
      s += v0;
      s += v1;
      s += v2;
      s += v3;
    }
    for (; b < b1; b += 8) s += __ldcg(p + (long)b * partial_stride);
  }
  sw[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && e < n) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sw[w][lane];
    out[e] = accumulate ? out[e] + t : t;
  }
}
static inline int stage2_blocks(int n) { return (n + 31) / 32; }

static int gram_blocks_for(long rows) {
  // aim for >= 2 waves of 148 SMs x 2 CTAs when there is enough work; >= 256 rows per block
  long nb = (rows + 255) / 256;
  if (nb > 592) nb = 592;
  if (nb < 1) nb = 1;
  return (int)nb;
}

size_t gram_partials_bytes(int B, int L) {
  long half = (B + 1) / 2;
  int nb = 2 * gram_blocks_for(half) + 2;
  return sizeof(float) * (size_t)nb * (size_t)(2 * L * L + 1) + 256;
}

template <int LT>
static int gram_launch(const float* F, const float* TF, const float* vmask, const float* roww,
                       const float* xrow, int L, long rb, long re, int cross, float* partials, int stride,
                       int block_off, int* nblocks, cudaStream_t st) {
  long rows = re - rb;
  if (rows <= 0) {
    *nblocks = 0;
    return 0;
  }
  int nb = gram_blocks_for(rows);
  int rpb = (int)((rows + nb - 1) / nb);
  nb = (int)((rows + rpb - 1) / rpb);
  gram_stage1_kernel<LT><<<nb, 256, 0, st>>>(F, TF, vmask, roww, xrow, L, rb, re, rpb, cross, partials,
                                             stride, block_off);
  NSVD_LAUNCH_CHECK();
  *nblocks = nb;
  return 0;
}

static int gram_dispatch(const float* F, const float* TF, const float* vmask, const float* roww,
                         const float* xrow, int L, long rb, long re, int cross, float* partials, int stride,
                         int block_off, int* nblocks, cudaStream_t st) {
  if (!cross && !roww && L > 16 && L <= 64) {   // tensor-core path of K2
    long rows = re - rb;
    if (rows <= 0) {
      *nblocks = 0;
      return 0;
    }
    int nb = gram_blocks_for(rows);
    const bool v2 = (L & 3) == 0 && ((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(TF) |
                                      reinterpret_cast<uintptr_t>(vmask)) & 15) == 0;
    if (v2 && nb > 444) nb = 444;               // 128-thread blocks of 154 registers: three per SM, one wave
    int rpb = (int)((rows + nb - 1) / nb);
    rpb = (rpb + 31) / 32 * 32;                 // whole 32-row chunks per block
    nb = (int)((rows + rpb - 1) / rpb);
    if (v2) {
      if (L <= 32) gram_mma2_kernel<32><<<nb, 128, 0, st>>>(F, TF, vmask, L, rb, re, rpb, partials, stride, block_off);
      else gram_mma2_kernel<64><<<nb, 128, 0, st>>>(F, TF, vmask, L, rb, re, rpb, partials, stride, block_off);
    } else if (L <= 32) gram_mma_kernel<32><<<nb, 128, 0, st>>>(F, TF, vmask, L, rb, re, rpb, partials, stride, block_off);
    else gram_mma_kernel<64><<<nb, 128, 0, st>>>(F, TF, vmask, L, rb, re, rpb, partials, stride, block_off);
    NSVD_LAUNCH_CHECK();
    *nblocks = nb;
    return 0;
  }
  if (L <= 16) return gram_launch<16>(F, TF, vmask, roww, xrow, L, rb, re, cross, partials, stride, block_off, nblocks, st);
  if (L <= 32) return gram_launch<32>(F, TF, vmask, roww, xrow, L, rb, re, cross, partials, stride, block_off, nblocks, st);
  if (L <= 48) return gram_launch<48>(F, TF, vmask, roww, xrow, L, rb, re, cross, partials, stride, block_off, nblocks, st);
  if (L <= 64) return gram_launch<64>(F, TF, vmask, roww, xrow, L, rb, re, cross, partials, stride, block_off, nblocks, st);
  set_error("gram_reduce: n_copies %d > 64 is handled by the CDK path", L);
  return NSVD_E_BADARG;
}

int gram_reduce(const float* F, const float* TF, const float* vmask, int B, int L, int b1,
                float* terms, void* partials_v, cudaStream_t st) {
  float* partials = (float*)partials_v;
  if (L == 16) {
    // blocks sized for >= 512 rows each, at most 4 x 148 blocks, split between the halves in proportion
    long rows_max = b1 > B - b1 ? b1 : B - b1;
    int per_half = (int)((rows_max + 1023) / 1024);
    if (per_half > 74) per_half = 74;   // 148 blocks of 1024 threads: one per SM
    if (per_half < 1) per_half = 1;
    int rpb = (int)((rows_max + per_half - 1) / per_half);
    rpb = (rpb + 511) / 512 * 512;
    int nb1 = (int)((b1 + rpb - 1) / rpb), nb2 = (int)((B - b1 + rpb - 1) / rpb);
    unsigned int* counter = reinterpret_cast<unsigned int*>(partials + (long)(nb1 + nb2) * 257);
    NSVD_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st));
    if (nb1 + nb2 == 0) return 0;
    gram16_fused_kernel<<<nb1 + nb2, 1024, 0, st>>>(F, TF, vmask, B, b1, rpb, nb1, partials, counter, terms);
    NSVD_LAUNCH_CHECK();
    return 0;
  }
  int stride = 2 * L * L + 1, n1 = 0, n2 = 0, rc;
  if ((rc = gram_dispatch(F, TF, vmask, nullptr, nullptr, L, 0, b1, 0, partials, stride, 0, &n1, st))) return rc;
  if ((rc = gram_dispatch(F, TF, vmask, nullptr, nullptr, L, b1, B, 0, partials, stride, n1, &n2, st))) return rc;
  int LL = L * L;
  gram_stage2_kernel<<<stage2_blocks(LL), 256, 0, st>>>(partials, stride, 0, n1, 0, LL, terms, 0);
  NSVD_LAUNCH_CHECK();
  gram_stage2_kernel<<<stage2_blocks(LL), 256, 0, st>>>(partials, stride, n1, n1 + n2, 0, LL, terms + LL, 0);
  NSVD_LAUNCH_CHECK();
  gram_stage2_kernel<<<1, 256, 0, st>>>(partials, stride, 0, n1 + n2, LL, 1, terms + 2 * LL, 0);
  NSVD_LAUNCH_CHECK();
  return 0;
}

int cross_gram(const float* F, const float* TF, const float* roww, const float* xrow, int B, int L, float* cov,
               float* quad, void* partials_v, cudaStream_t st) {
  float* partials = (float*)partials_v;
  int stride = 2 * L * L + 1, n1 = 0, rc;
  if ((rc = gram_dispatch(F, TF, nullptr, roww, xrow, L, 0, B, 1, partials, stride, 0, &n1, st))) return rc;
  int LL = L * L;
  gram_stage2_kernel<<<stage2_blocks(LL), 256, 0, st>>>(partials, stride, 0, n1, 0, LL, cov, 1);
  NSVD_LAUNCH_CHECK();
  gram_stage2_kernel<<<stage2_blocks(LL), 256, 0, st>>>(partials, stride, 0, n1, LL, LL, quad, 1);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// loss + backward coefficient matrices from the (all-reduced) terms.  One block.
__global__ void loss_finalize_kernel(const float* __restrict__ terms, const float* __restrict__ Mm,
                                     int L, double Bg, double B1g, double B2g,
                                     float* __restrict__ loss, float* __restrict__ coef) {
  __shared__ double sred[32];
  int LL = L * L;
  if (Bg <= 0.0) {
    // global row counts travel WITH the all-reduced buffer (no second collective, no host copy): four floats after the
    // operator sum, [n mod 2^16, n / 2^16, b1 mod 2^16, b1 / 2^16] summed over the ranks (each sum exact in fp32)
    const float* c = terms + 2 * LL + 1;
    Bg = (double)c[0] + 65536.0 * (double)c[1];
    B1g = (double)c[2] + 65536.0 * (double)c[3];
    B2g = Bg - B1g;
    if (threadIdx.x == 0) coef[2 * LL] = (float)(4.0 / Bg);      // read by loss_dF (Bg <= 0 there too)
  }
  double s = 0.0;
  for (int e = threadIdx.x; e < LL; e += blockDim.x) {
    double lam1 = (double)terms[e] / B1g, lam2 = (double)terms[LL + e] / B2g;
    double m = Mm[e];
    s += m * lam1 * lam2;
    coef[e] = (float)(2.0 / B1g * m * lam2);
    coef[LL + e] = (float)(2.0 / B2g * m * lam1);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sred[i];
    loss[0] = (float)(-2.0 * (double)terms[2 * LL] / Bg + t);
  }
}

int loss_finalize(const float* terms, const float* Mm, int L, long Bg, long B1g, long B2g,
                  float* loss, float* coef, cudaStream_t st) {
  loss_finalize_kernel<<<1, 256, 0, st>>>(terms, Mm, L, (double)Bg, (double)B1g, (double)B2g, loss, coef);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// K3 loss_dF: dF[b][m] = gs * ( -(4/Bg) v_m TF[b][m] + sum_l F[b][l] coef_half[l][m] )
template <int LT>
__global__ void __launch_bounds__(256)
loss_dF_kernel(const float* __restrict__ F, const float* __restrict__ TF,
               const float* __restrict__ vmask, const float* __restrict__ coef,
               const float* __restrict__ gscale, int B, int L, int b1, float c4, float* __restrict__ dF) {
  if (c4 <= 0.f) c4 = coef[2 * L * L];          // 4 / B_global left by loss_finalize (device-side counts)
  extern __shared__ float sm[];
  float* sC = sm;                 // [2][L][LT]
  float* sV = sm + 2 * L * LT;    // [LT]
  for (int e = threadIdx.x; e < 2 * L * LT; e += blockDim.x) {
    int h = e / (L * LT), r = (e / LT) % L, c = e % LT;
    sC[e] = (coef && c < L) ? coef[h * L * L + r * L + c] : 0.f;
  }
  for (int e = threadIdx.x; e < LT; e += blockDim.x) sV[e] = e < L ? vmask[e] : 0.f;
  __syncthreads();
  const float gs = gscale ? gscale[0] : 1.f;
  // each thread produces 4 consecutive outputs of one row
  constexpr int Q = LT / 4;
  long total = (long)B * Q;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    long b = i / Q;
    int m0 = (int)(i % Q) * 4;
    if (m0 >= L) continue;
    const float* frow = F + b * L;
    const float* C = sC + (b < b1 ? 0 : L * LT);
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    for (int l = 0; l < L; ++l) {
      float fv = __ldg(frow + l);
      const float* cr = C + l * LT + m0;
      o0 = fmaf(fv, cr[0], o0);
      o1 = fmaf(fv, cr[1], o1);
      o2 = fmaf(fv, cr[2], o2);
      o3 = fmaf(fv, cr[3], o3);
    }
    float o[4] = {o0, o1, o2, o3};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int m = m0 + q;
      if (m < L) dF[b * L + m] = gs * (o[q] - (TF ? c4 * sV[m] * TF[b * L + m] : 0.f));
    }
  }
}

// 16 < L <= 64: register-tiled version.  A block stages a tile of rows of F in shared memory (coalesced loads) next to
// both coefficient halves; thread (ty, tx) produces a 4-row x 4-column block of dF: per l four broadcast reads of F,
// one float4 of coefficients and 16 FMAs (the thread-per-4-outputs kernel above does one load per FMA: 1.53 ms for
// 2^20 x 64, 8 % of the HBM roofline).  A tile that straddles the half boundary b1 takes its coefficients per row.
template <int LT>   // 32 or 64: padded row length; LT / 4 column groups, 256 / (LT / 4) row groups of 4 rows
__global__ void __launch_bounds__(256)
loss_dF_tile_kernel(const float* __restrict__ F, const float* __restrict__ TF, const float* __restrict__ vmask,
                    const float* __restrict__ coef, const float* __restrict__ gscale, int B, int L, int b1, float c4,
                    float* __restrict__ dF) {
  if (c4 <= 0.f) c4 = coef[2 * L * L];          // 4 / B_global left by loss_finalize (device-side counts)
  constexpr int CG = LT / 4, RG = 256 / CG, ROWS = RG * 4;
  extern __shared__ float sm[];
  float* sC = sm;                           // [2][L][LT]
  float* sV = sC + 2 * L * LT;              // [LT]
  float* sF = sV + LT;                      // [ROWS][LT + 1]
  for (int e = threadIdx.x; e < 2 * L * LT; e += 256) {
    int h = e / (L * LT), r = (e / LT) % L, c = e % LT;
    sC[e] = (coef && c < L) ? coef[h * L * L + r * L + c] : 0.f;
  }
  for (int e = threadIdx.x; e < LT; e += 256) sV[e] = e < L ? vmask[e] : 0.f;
  const float gs = gscale ? gscale[0] : 1.f;
  const int tx = threadIdx.x % CG, ty = threadIdx.x / CG, m0 = tx * 4;
  for (long base = (long)blockIdx.x * ROWS; base < B; base += (long)gridDim.x * ROWS) {
    __syncthreads();
    const int nrow = (B - base) < ROWS ? (int)(B - base) : ROWS;
    for (int e = threadIdx.x; e < ROWS * L; e += 256) {        // rows are contiguous in F: fully coalesced
      const int r = e / L, c = e % L;
      sF[r * (LT + 1) + c] = r < nrow ? F[base * L + e] : 0.f;
    }
    __syncthreads();
    float o[4][4] = {};
    const long r0 = base + ty * 4;
    const bool uniform = (r0 < b1) == (r0 + 3 < b1);
    const float* f0 = sF + (ty * 4) * (LT + 1);
    if (uniform) {
      const float* C = sC + (r0 < b1 ? 0 : L * LT) + m0;
#pragma unroll 4
      for (int l = 0; l < L; ++l) {
        const float4 c = *reinterpret_cast<const float4*>(C + l * LT);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float fv = f0[r * (LT + 1) + l];
          o[r][0] = fmaf(fv, c.x, o[r][0]);
          o[r][1] = fmaf(fv, c.y, o[r][1]);
          o[r][2] = fmaf(fv, c.z, o[r][2]);
          o[r][3] = fmaf(fv, c.w, o[r][3]);
        }
      }
    } else {
      for (int l = 0; l < L; ++l) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 c = *reinterpret_cast<const float4*>(sC + ((r0 + r) < b1 ? 0 : L * LT) + m0 + l * LT);
          const float fv = f0[r * (LT + 1) + l];
          o[r][0] = fmaf(fv, c.x, o[r][0]);
          o[r][1] = fmaf(fv, c.y, o[r][1]);
          o[r][2] = fmaf(fv, c.z, o[r][2]);
          o[r][3] = fmaf(fv, c.w, o[r][3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long b = r0 + r;
      if (b >= B) continue;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = m0 + q;
        if (m < L) dF[b * L + m] = gs * (o[r][q] - (TF ? c4 * sV[m] * TF[b * L + m] : 0.f));
      }
    }
  }
}

// L == 16 specialisation: one thread per row.  Rows are staged through shared memory so that every
// global access is a fully coalesced 16-byte-per-lane transfer (256 rows x 64 B per block and array); the
// 16x16 coefficient block of the row's half is read from shared memory as broadcast float4.
__global__ void __launch_bounds__(256)
loss_dF16_kernel(const float* __restrict__ F, const float* __restrict__ TF, const float* __restrict__ vmask,
                 const float* __restrict__ coef, const float* __restrict__ gscale, int B, int b1, float c4,
                 float* __restrict__ dF) {
  if (c4 <= 0.f) c4 = coef[512];                // 4 / B_global left by loss_finalize (device-side counts)
  __shared__ float4 sC[2][16][4];
  __shared__ float sV[16];
  __shared__ float sF[256][17];
  __shared__ float sT[256][17];
  const int tid = threadIdx.x;
  for (int e = tid; e < 512; e += 256) reinterpret_cast<float*>(sC)[e] = coef ? coef[e] : 0.f;
  if (tid < 16) sV[tid] = vmask[tid];
  const float gs = gscale ? gscale[0] : 1.f;
  for (long base = (long)blockIdx.x * 256; base < B; base += (long)gridDim.x * 256) {
    __syncthreads();
    const long nrow = (B - base) < 256 ? (B - base) : 256;
    // coalesced loads: 256 rows x 4 float4
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;          // float4 index inside the 256 x 16 block
      const int r = e >> 2, q = e & 3;
      float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f), t4 = f4;
      if (r < nrow) {
        f4 = reinterpret_cast<const float4*>(F + base * 16)[e];
        if (TF) t4 = reinterpret_cast<const float4*>(TF + base * 16)[e];
      }
      sF[r][4 * q] = f4.x; sF[r][4 * q + 1] = f4.y; sF[r][4 * q + 2] = f4.z; sF[r][4 * q + 3] = f4.w;
      sT[r][4 * q] = t4.x; sT[r][4 * q + 1] = t4.y; sT[r][4 * q + 2] = t4.z; sT[r][4 * q + 3] = t4.w;
    }
    __syncthreads();
    float f[16], o[16] = {};
#pragma unroll
    for (int l = 0; l < 16; ++l) f[l] = sF[tid][l];
    const int h = (base + tid) < b1 ? 0 : 1;
#pragma unroll
    for (int l = 0; l < 16; ++l) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 c = sC[h][l][q];
        o[4 * q] = fmaf(f[l], c.x, o[4 * q]);
        o[4 * q + 1] = fmaf(f[l], c.y, o[4 * q + 1]);
        o[4 * q + 2] = fmaf(f[l], c.z, o[4 * q + 2]);
        o[4 * q + 3] = fmaf(f[l], c.w, o[4 * q + 3]);
      }
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) sF[tid][m] = gs * (o[m] - c4 * sV[m] * sT[tid][m]);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      const int r = e >> 2, q = e & 3;
      if (r < nrow)
        reinterpret_cast<float4*>(dF + base * 16)[e] =
            make_float4(sF[r][4 * q], sF[r][4 * q + 1], sF[r][4 * q + 2], sF[r][4 * q + 3]);
    }
  }
}

// K3 at L = 16, second version (fused-step shape: TF and coef both present): the kernel further down stages 256 rows
// through shared memory between two block-wide synchronisations - while a block computes, nothing of its next rows is in
// flight (3.9 TB/s).  Here, as in the L = 16 Gram kernel, a warp owns row pairs: one coalesced 128-byte load per array
// and pair, 8 pairs per iteration with the next iteration's loads already issued; lane = (row of the pair, column m)
// keeps column m of BOTH coefficient halves in registers, parks its F value in the warp's 128-byte shared-memory slot
// and reads the 16 values of its row back as four 16-byte broadcast loads; 16 FMAs, one coalesced 128-byte store.
__global__ void __launch_bounds__(512)
loss_dF16_pipe_kernel(const float* __restrict__ F, const float* __restrict__ TF, const float* __restrict__ vmask,
                      const float* __restrict__ coef, const float* __restrict__ gscale, long B, long b1, float c4,
                      int rows_per_block, float* __restrict__ dF) {
  constexpr int U = 8, W = 16;                  // row pairs per iteration and warp; warps per block
  __shared__ __align__(16) float stage[W][U * 32];
  if (c4 <= 0.f) c4 = coef[512];                // 4 / B_global left by loss_finalize (device-side counts)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, m = lane & 15, rr = lane >> 4;
  const float gs = gscale ? gscale[0] : 1.f;
  float c0[16], c1[16];                         // column m of the two coefficient halves
#pragma unroll
  for (int l = 0; l < 16; ++l) {
    c0[l] = coef[l * 16 + m];
    c1[l] = coef[256 + l * 16 + m];
  }
  const float cv = c4 * vmask[m];
  const long r0 = (long)blockIdx.x * rows_per_block;
  const long r1 = r0 + rows_per_block < B ? r0 + rows_per_block : B;
  float fo[U], to[U], fn[U], tn[U];
  auto load8 = [&](long r, float* f8, float* t8) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long row = r + 2 * W * u + rr;
      const bool ok = row < r1;
      f8[u] = ok ? __ldcs(F + row * 16 + m) : 0.f;
      t8[u] = ok ? __ldcs(TF + row * 16 + m) : 0.f;
    }
  };
  float* st = &stage[warp][0];
  long r = r0 + 2 * warp;
  load8(r, fo, to);
  for (; r < r1; r += 2 * W * U) {
    load8(r + 2 * W * U, fn, tn);               // rows beyond r1 load as zeros
#pragma unroll
    for (int u = 0; u < U; ++u) st[u * 32 + lane] = fo[u];
    __syncwarp();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long row = r + 2 * W * u + rr;
      const float4* f4 = reinterpret_cast<const float4*>(st + u * 32 + 16 * rr);
      const float4 a = f4[0], b = f4[1], c = f4[2], d = f4[3];
      const float fv[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
      float o = 0.f;
      if (row < b1) {
#pragma unroll
        for (int l = 0; l < 16; ++l) o = fmaf(fv[l], c0[l], o);
      } else {
#pragma unroll
        for (int l = 0; l < 16; ++l) o = fmaf(fv[l], c1[l], o);
      }
      if (row < r1) __stcs(dF + row * 16 + m, gs * (o - cv * to[u]));
    }
    __syncwarp();                               // the next iteration overwrites the staged pairs
#pragma unroll
    for (int u = 0; u < U; ++u) {
      fo[u] = fn[u];
      to[u] = tn[u];
    }
  }
}

// K3 for 16 < L <= 64, L a multiple of 4, on warp-level tensor cores (mma.sync m16n8k8, 3xTF32 - the arithmetic of K2):
//   dF[b][m] = gs * ( sum_l F[b][l] C_h[l][m] - c4 v_m TF[b][m] ),   h = (b >= b1)
// is a (rows x L) . (L x L) product per half.  A block walks a contiguous row range; per half it splits the coefficient
// block ONCE into tf32 hi / lo planes in shared memory, then per 64-row chunk: F is read with 16-byte loads into
// registers one chunk ahead, split once into planes, a warp multiplies 2 x WN tiles (fragments reused in registers),
// the epilogue reads TF and writes dF as 8-byte pairs straight from the accumulator layout.  (The CUDA-core kernel above
// is bound by shared-memory loads - 5 per 16 FMAs - with nothing in flight while it multiplies: 1.9 TB/s at L = 64.)
template <int LP>
__global__ void __launch_bounds__(128)
loss_dF_mma_kernel(const float* __restrict__ F, const float* __restrict__ TF, const float* __restrict__ vmask,
                   const float* __restrict__ coef, const float* __restrict__ gscale, int B, int L, int b1, float c4,
                   int rows_per_block, float* __restrict__ dF) {
  constexpr int TN = LP / 8, WN = TN / 2, WM = 2, CH = 64, KS = LP / 8;
  constexpr int SF = LP + 4;                    // F planes [row][k]: A fragments (8 rows x 4 k) hit 32 distinct banks
  constexpr int SC = LP + 8;                    // C planes [k][m]:  B fragments (4 k x 8 m) hit 32 distinct banks
  constexpr int NV = CH * LP / 4 / 128;         // float4 per thread and chunk
  extern __shared__ __align__(16) uint32_t smem_u[];
  uint32_t* Ch = smem_u;                        // [LP][SC]
  uint32_t* Cl = Ch + LP * SC;
  uint32_t* Fh = Cl + LP * SC;                  // [CH][SF]
  uint32_t* Fl = Fh + CH * SF;
  if (c4 <= 0.f) c4 = coef[2 * L * L];          // 4 / B_global left by loss_finalize (device-side counts)
  const float gs = gscale ? gscale[0] : 1.f;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp >> 1) * WM, n0 = (warp & 1) * WN;
  const long R0 = (long)blockIdx.x * rows_per_block;
  const long R1 = R0 + rows_per_block < B ? R0 + rows_per_block : B;
  for (int i = tid; i < CH * SF; i += 128) {    // columns >= L of the F planes stay zero
    Fh[i] = 0u;
    Fl[i] = 0u;
  }
  const int q4 = CH * (L >> 2);                 // float4 per full chunk
  float4 fr[NV];
  for (int h = 0; h < 2; ++h) {
    const long r0 = h == 0 ? R0 : (R0 > b1 ? R0 : (long)b1);
    const long r1 = h == 0 ? (R1 < b1 ? R1 : (long)b1) : R1;
    if (r0 >= r1) continue;
    __syncthreads();                            // the previous half's multiplies are done with the C planes
    for (int e = tid; e < LP * LP; e += 128) {  // coefficient block of this half, zero beyond L
      const int l = e / LP, m = e % LP;
      const float c = (l < L && m < L) ? coef[(long)h * L * L + l * L + m] : 0.f;
      const uint32_t hi = to_tf32(c);
      Ch[l * SC + m] = hi;
      Cl[l * SC + m] = to_tf32(c - __uint_as_float(hi));
    }
    auto load_chunk = [&](long c0) {
      const long nel4 = ((r1 - c0) < CH ? (r1 - c0) : CH) * (long)(L >> 2);
      const float4* f4 = reinterpret_cast<const float4*>(F + c0 * L);
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int e = tid + 128 * u;
        fr[u] = (e < q4 && e < nel4) ? __ldcs(f4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    load_chunk(r0);
    for (long c0 = r0; c0 < r1; c0 += CH) {
      __syncthreads();                          // the previous chunk's multiplies are done with the F planes
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        const int e = tid + 128 * u;
        if (e < q4) {
          const int row = (4 * e) / L, col = (4 * e) % L;
          uint4 hi, lo;
          hi.x = to_tf32(fr[u].x); lo.x = to_tf32(fr[u].x - __uint_as_float(hi.x));
          hi.y = to_tf32(fr[u].y); lo.y = to_tf32(fr[u].y - __uint_as_float(hi.y));
          hi.z = to_tf32(fr[u].z); lo.z = to_tf32(fr[u].z - __uint_as_float(hi.z));
          hi.w = to_tf32(fr[u].w); lo.w = to_tf32(fr[u].w - __uint_as_float(hi.w));
          *reinterpret_cast<uint4*>(Fh + row * SF + col) = hi;
          *reinterpret_cast<uint4*>(Fl + row * SF + col) = lo;
        }
      }
      __syncthreads();
      if (c0 + CH < r1) load_chunk(c0 + CH);    // in flight while this chunk is multiplied
      // the TF values of this thread's outputs, also in flight under the multiply
      float2 tfv[WM][WN][2];
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int col = 8 * (n0 + j) + 2 * t;
            const long row = c0 + 16 * (m0 + i) + g + 8 * hh;
            tfv[i][j][hh] = (col < L && row < r1) ? __ldcs(reinterpret_cast<const float2*>(TF + row * L + col))
                                                  : make_float2(0.f, 0.f);
          }
      float acc[WM][WN][4];
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int k0 = ks * 8 + t;
        // B = C_h (K = l, N = m): b0 (k = t, n = g), b1 (k = t + 4, n = g)
        uint32_t bh[WN][2], bl[WN][2];
#pragma unroll
        for (int j = 0; j < WN; ++j) {
          const int c = 8 * (n0 + j) + g;
          bh[j][0] = Ch[k0 * SC + c]; bh[j][1] = Ch[(k0 + 4) * SC + c];
          bl[j][0] = Cl[k0 * SC + c]; bl[j][1] = Cl[(k0 + 4) * SC + c];
        }
#pragma unroll
        for (int i = 0; i < WM; ++i) {
          // A = F chunk (M = row, K = l): a0 (g, t), a1 (g + 8, t), a2 (g, t + 4), a3 (g + 8, t + 4)
          const int r = 16 * (m0 + i) + g;
          const uint32_t ah[4] = {Fh[r * SF + k0], Fh[(r + 8) * SF + k0], Fh[r * SF + k0 + 4], Fh[(r + 8) * SF + k0 + 4]};
          const uint32_t al[4] = {Fl[r * SF + k0], Fl[(r + 8) * SF + k0], Fl[r * SF + k0 + 4], Fl[(r + 8) * SF + k0 + 4]};
#pragma unroll
          for (int j = 0; j < WN; ++j) {
            mma_tf32_16x8x8(acc[i][j], al, bh[j]);
            mma_tf32_16x8x8(acc[i][j], ah, bl[j]);
            mma_tf32_16x8x8(acc[i][j], ah, bh[j]);
          }
        }
      }
      // C fragment: c0 (g, 2t), c1 (g, 2t + 1), c2 (g + 8, 2t), c3 (g + 8, 2t + 1) inside tile (m0 + i, n0 + j)
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) {
          const int col = 8 * (n0 + j) + 2 * t;
          if (col < L) {
            const float2 vv = *reinterpret_cast<const float2*>(vmask + col);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const long row = c0 + 16 * (m0 + i) + g + 8 * hh;
              if (row < r1) {
                const float2 tf = tfv[i][j][hh];
                float2 o;
                o.x = gs * (acc[i][j][2 * hh] - c4 * vv.x * tf.x);
                o.y = gs * (acc[i][j][2 * hh + 1] - c4 * vv.y * tf.y);
                __stcs(reinterpret_cast<float2*>(dF + row * L + col), o);
              }
            }
          }
        }
    }
  }
}

int loss_dF(const float* F, const float* TF, const float* vmask, const float* coef,
            const float* gscale, int B, int L, int b1, long Bg, float* dF, cudaStream_t st) {
  float c4 = Bg > 0 ? (float)(4.0 / (double)Bg) : 0.f;   // Bg <= 0: the kernels read 4 / B_global from coef[2 L^2]
  if (Bg <= 0 && (!coef || !TF)) {
    set_error("loss_dF: device-side counts (Bg <= 0) need both TF and coef");
    return NSVD_E_BADARG;
  }
  if (L == 16 && coef && TF && B >= 4096) {     // software-pipelined warp-per-row-pair kernel
    int nbp = cdiv(B, 2048);                    // >= 8 iterations of 256 rows per block ...
    if (nbp > 148) nbp = 148;                   // ... one 512-thread block (97 registers per thread) per SM
    int rpb = cdiv(B, nbp);
    rpb = (rpb + 255) / 256 * 256;
    nbp = cdiv(B, rpb);
    loss_dF16_pipe_kernel<<<nbp, 512, 0, st>>>(F, TF, vmask, coef, gscale, B, b1, c4, rpb, dF);
    NSVD_LAUNCH_CHECK();
    return 0;
  }
  if (L == 16) {
    int nb16 = cdiv(B, 256);
    if (nb16 > 148 * 6) nb16 = 148 * 6;
    loss_dF16_kernel<<<nb16, 256, 0, st>>>(F, TF, vmask, coef, gscale, B, b1, c4, dF);
    NSVD_LAUNCH_CHECK();
    return 0;
  }
  int LT = L <= 16 ? 16 : (L <= 32 ? 32 : (L <= 48 ? 48 : 64));
  if (L > 64) {
    set_error("loss_dF: n_copies %d > 64", L);
    return NSVD_E_BADARG;
  }
  if (L > 16 && (L & 3) == 0 && coef && TF &&
      ((reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(TF) | reinterpret_cast<uintptr_t>(dF) |
        reinterpret_cast<uintptr_t>(vmask)) & 15) == 0) {   // tensor-core kernel
    const int LPP = L <= 32 ? 32 : 64;
    const size_t smem_m = sizeof(uint32_t) * (size_t)(2 * LPP * (LPP + 8) + 2 * 64 * (LPP + 4));
    int nbm = cdiv(B, 256);                    // >= 4 chunks per block, at most three blocks per SM
    if (nbm > 444) nbm = 444;
    int rpbm = cdiv(B, nbm);
    rpbm = (rpbm + 63) / 64 * 64;
    nbm = cdiv(B, rpbm);
    if (LPP == 32) {
      NSVD_SMEM_OPTIN(loss_dF_mma_kernel<32>, 100 * 1024);
      loss_dF_mma_kernel<32><<<nbm, 128, smem_m, st>>>(F, TF, vmask, coef, gscale, B, L, b1, c4, rpbm, dF);
    } else {
      NSVD_SMEM_OPTIN(loss_dF_mma_kernel<64>, 100 * 1024);
      loss_dF_mma_kernel<64><<<nbm, 128, smem_m, st>>>(F, TF, vmask, coef, gscale, B, L, b1, c4, rpbm, dF);
    }
    NSVD_LAUNCH_CHECK();
    return 0;
  }
  if (L > 16) {   // register-tiled kernel
    const int LTT = L <= 32 ? 32 : 64, rows = (256 / (LTT / 4)) * 4;
    const size_t smem_t = sizeof(float) * (size_t)(2 * L * LTT + LTT + rows * (LTT + 1));
    int nbt = cdiv(B, rows);
    if (nbt > 148 * 4) nbt = 148 * 4;
    if (LTT == 32) {
      NSVD_SMEM_OPTIN(loss_dF_tile_kernel<32>, 100 * 1024);
      loss_dF_tile_kernel<32><<<nbt, 256, smem_t, st>>>(F, TF, vmask, coef, gscale, B, L, b1, c4, dF);
    } else {
      NSVD_SMEM_OPTIN(loss_dF_tile_kernel<64>, 100 * 1024);
      loss_dF_tile_kernel<64><<<nbt, 256, smem_t, st>>>(F, TF, vmask, coef, gscale, B, L, b1, c4, dF);
    }
    NSVD_LAUNCH_CHECK();
    return 0;
  }
  // L < 16: four outputs per thread
  size_t smem = sizeof(float) * (2 * L * LT + LT);
  long total = (long)B * (LT / 4);
  int nb = cdiv(total, 256);
  if (nb > 148 * 8) nb = 148 * 8;
  if (nb < 1) nb = 1;
  loss_dF_kernel<16><<<nb, 256, smem, st>>>(F, TF, vmask, coef, gscale, B, L, b1, c4, dF);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// CDK (methods/nestedlora.py:270-332), v1 on the fp32 building blocks
// ------------------------------------------------------------------------------------------
// pad a leading constant-1 column: out (B, Lp) from in (B, L)
__global__ void cdk_pad_kernel(const float* __restrict__ in, float* __restrict__ out, long B, int L,
                               int first_const) {
  int Lp = L + first_const;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Lp) return;
  long b = i / Lp;
  int c = (int)(i % Lp);
  out[i] = (first_const && c == 0) ? 1.f : in[b * L + c - first_const];
}

// per-row  sum_l v_l f g  and diag(Fp Gp^T); one warp per row; rowsum partials reduced by colsum
__global__ void cdk_rowdots_kernel(const float* __restrict__ fp, const float* __restrict__ gp,
                                   const float* __restrict__ v, long B, int Lp,
                                   float* __restrict__ opdot, float* __restrict__ rs_joint) {
  long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (w >= B) return;
  float a = 0.f, d = 0.f;
  for (int c = lane; c < Lp; c += 32) {
    float p = fp[w * Lp + c] * gp[w * Lp + c];
    d += p;
    a = fmaf(v[c], p, a);
  }
  a = warp_sum(a);
  d = warp_sum(d);
  if (lane == 0) {
    opdot[w] = a;
    if (rs_joint) rs_joint[w] = d;
  }
}

size_t cdk_work_bytes(int B, int L, int fc) {
  long Lp = L + fc;
  return sizeof(float) * (size_t)(2L * B * Lp + B + 64) + 256;
}

int cdk_fwd(const float* f, const float* g, const float* v, int B, int L, int fc, float* terms,
            float* rs_joint, void* work, cudaStream_t st) {
  int Lp = L + fc;
  float* fp = (float*)work;
  float* gp = fp + (long)B * Lp;
  float* opdot = gp + (long)B * Lp;
  long n = (long)B * Lp;
  cdk_pad_kernel<<<cdiv(n, 256), 256, 0, st>>>(f, fp, B, L, fc);
  NSVD_LAUNCH_CHECK();
  cdk_pad_kernel<<<cdiv(n, 256), 256, 0, st>>>(g, gp, B, L, fc);
  NSVD_LAUNCH_CHECK();
  int rc;
  for (int which = 0; which < 2; ++which) {
    const float* X = which ? gp : fp;
    SGemm w{};
    w.A = X; w.a_rs = 1; w.a_cs = Lp; w.a_bs = 0;
    w.B = X; w.b_rs = Lp; w.b_cs = 1; w.b_bs = 0;
    w.C = terms + (long)which * Lp * Lp; w.c_rs = Lp; w.c_bs = 0;
    w.M = Lp; w.N = Lp; w.K = B; w.alpha = 1.f; w.accumulate = 0;
    if ((rc = sgemm_strided(w, 1, st))) return rc;
  }
  cdk_rowdots_kernel<<<cdiv((long)B * 32, 256), 256, 0, st>>>(fp, gp, v, B, Lp, opdot, rs_joint);
  NSVD_LAUNCH_CHECK();
  return colsum(opdot, terms + 2L * Lp * Lp, B, 1, 1, 0, 0, st);
}

// loss_metric = sum M (.) Lambda_f (.) Lambda_g and the two masked coefficient matrices of the backward, in fp64.
// Lp = 513 means 263 k entries and 3.2 MB of reads: on ONE block (the first version) that is a memory-latency-bound
// 364 us, 40 % of the whole CDK step.  Now <= 64 blocks each reduce a contiguous slice to one fp64 partial (fixed
// in-block order) and a second one-warp launch adds the partials in block order - deterministic, ~5 us.
constexpr int kCdkFinBlocks = 64;
__global__ void __launch_bounds__(1024)
cdk_finalize_kernel(const float* __restrict__ terms, const float* __restrict__ Mm, int Lp, double Bg,
                    float* __restrict__ coef, double* __restrict__ partial) {
  __shared__ double sred[32];
  const int LL = Lp * Lp;
  const int per = (LL + gridDim.x - 1) / gridDim.x;
  const int e0 = blockIdx.x * per, e1 = min(LL, e0 + per);
  double s = 0.0;
  for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    double lf = (double)terms[e] / Bg, lg = (double)terms[LL + e] / Bg;
    double m = Mm[e];
    s += m * lf * lg;
    coef[e] = (float)(2.0 / Bg * m * lg);        // multiplies rows of Fp
    coef[LL + e] = (float)(2.0 / Bg * m * lf);   // multiplies rows of Gp
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sred[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void cdk_finalize_sum_kernel(const float* __restrict__ terms, const double* __restrict__ partial, int nb,
                                        int Lp, double Bg, float* __restrict__ losses) {
  double p = (int)threadIdx.x < nb ? partial[threadIdx.x] : 0.0;   // nb <= 64: all partials in flight at once
  __shared__ double sp[kCdkFinBlocks];
  sp[threadIdx.x] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nb; ++i) t += sp[i];
    const int LL = Lp * Lp;
    double lop = -2.0 * (double)terms[2 * LL] / Bg;
    losses[0] = (float)(lop + t);
    losses[1] = (float)lop;
    losses[2] = (float)t;
  }
}

int cdk_finalize(const float* terms, const float* Mm, int Lp, long Bg, float* losses, float* coef, double* scratch,
                 cudaStream_t st) {
  const long LL = (long)Lp * Lp;
  int nb = (int)((LL + 4095) / 4096);            // >= 4 entries per thread before another block pays off
  if (nb > kCdkFinBlocks) nb = kCdkFinBlocks;
  static_assert(kCdkFinBlocks * sizeof(double) <= NSVD_CDK_FINALIZE_SCRATCH, "finalize scratch");
  cdk_finalize_kernel<<<nb, 1024, 0, st>>>(terms, Mm, Lp, (double)Bg, coef, scratch);
  NSVD_LAUNCH_CHECK();
  cdk_finalize_sum_kernel<<<1, kCdkFinBlocks, 0, st>>>(terms, scratch, nb, Lp, (double)Bg, losses);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// grad_f[b][c] = gs * ( -(2/Bg) v_{c+fc} g[b][c] + sum_i Fp[b][i] coefF[i][c+fc] ),  same for g.
// Written as a batched SGEMM over the un-padded inputs + a rank-1 term for the constant column.
__global__ void cdk_bwd_epilogue_kernel(float* __restrict__ grad, const float* __restrict__ other,
                                        const float* __restrict__ v, const float* __restrict__ coef,
                                        const float* __restrict__ gscale, long B, int L, int fc,
                                        float c2) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  int c = (int)(i % L);
  int Lp = L + fc;
  float gs = gscale ? gscale[0] : 1.f;
  float acc = grad[i];
  if (fc) acc += coef[c + fc];  // row 0 of coef (constant-1 input column)
  grad[i] = gs * (acc - c2 * v[c + fc] * other[i]);
  (void)Lp;
}

int cdk_bwd(const float* f, const float* g, const float* v, const float* coef, const float* gscale,
            int B, int L, int fc, long Bg, float* grad_f, float* grad_g, cudaStream_t st) {
  int Lp = L + fc, rc;
  float c2 = (float)(2.0 / (double)Bg);
  for (int which = 0; which < 2; ++which) {
    const float* X = which ? g : f;
    const float* O = which ? f : g;
    float* G = which ? grad_g : grad_f;
    const float* C = coef + (long)which * Lp * Lp;
    // G (B x L) = X (B x L) . C[fc:, fc:] (L x L)
    SGemm d{};
    d.A = X; d.a_rs = L; d.a_cs = 1; d.a_bs = 0;
    d.B = C + (long)fc * Lp + fc; d.b_rs = Lp; d.b_cs = 1; d.b_bs = 0;
    d.C = G; d.c_rs = L; d.c_bs = 0;
    d.M = B; d.N = L; d.K = L; d.alpha = 1.f; d.accumulate = 0;
    if ((rc = sgemm_strided(d, 1, st))) return rc;
    long n = (long)B * L;
    cdk_bwd_epilogue_kernel<<<cdiv(n, 256), 256, 0, st>>>(G, O, v, C, gscale, B, L, fc, c2);
    NSVD_LAUNCH_CHECK();
  }
  return 0;
}

// off_diagonal(Fp Gp^T): out[i*(B-1) + (j - (j>i))] = fc + f_i . g_j, i != j   (methods/utils.py:16-22)
__global__ void cdk_offdiag_kernel(const float* __restrict__ f, const float* __restrict__ g, int B,
                                   int L, int fc, float* __restrict__ out) {
  __shared__ float sf[32][33], sg[32][33];
  int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  int tx = threadIdx.x, ty = threadIdx.y;
  float acc = fc ? 1.f : 0.f;
  for (int k0 = 0; k0 < L; k0 += 32) {
    sf[ty][tx] = (i0 + ty < B && k0 + tx < L) ? f[(long)(i0 + ty) * L + k0 + tx] : 0.f;
    sg[ty][tx] = (j0 + ty < B && k0 + tx < L) ? g[(long)(j0 + ty) * L + k0 + tx] : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = fmaf(sf[ty][k], sg[tx][k], acc);
    __syncthreads();
  }
  int i = i0 + ty, j = j0 + tx;
  if (i < B && j < B && i != j) out[(long)i * (B - 1) + (j - (j > i ? 1 : 0))] = acc;
}

int cdk_offdiag(const float* f, const float* g, int B, int L, int fc, float* out, cudaStream_t st) {
  dim3 grid(cdiv(B, 32), cdiv(B, 32));
  cdk_offdiag_kernel<<<grid, dim3(32, 32), 0, st>>>(f, g, B, L, fc, out);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// "next" rows: fused optimizer step and on-device sampler
// ------------------------------------------------------------------------------------------
// RMSprop (momentum 0, weight decay 0: examples/utils.py:48-57) + EMA shadow update (torch_ema) for up to
// 16 tensors in one launch:  sq = alpha sq + (1-alpha) g^2 ; p -= lr g / (sqrt(sq) + eps) ;
//                            ema -= ema_w (ema - p)        (ema_w = 1 - decay_t ; skipped when ema == NULL)
__global__ void rmsprop_ema_kernel(OptTensors t, float lr, float alpha, float eps, float ema_w) {
  const int ti = blockIdx.y;
  if (ti >= t.n) return;
  float* __restrict__ p = t.p[ti];
  const float* __restrict__ g = t.g[ti];
  float* __restrict__ sq = t.sq[ti];
  float* __restrict__ em = t.ema[ti];
  const long n = t.size[ti];
  const float oma = 1.f - alpha;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float gi = g[i];
    float s = sq[i] * alpha;
    s = s + oma * gi * gi;
    sq[i] = s;
    float pi = p[i] - lr * (gi / (sqrtf(s) + eps));
    p[i] = pi;
    if (em) {
      float e = em[i];
      em[i] = e - ema_w * (e - pi);
    }
  }
}

int rmsprop_ema_step(const OptTensors& t, float lr, float alpha, float eps, float ema_w, cudaStream_t st) {
  long mx = 0;
  for (int i = 0; i < t.n; ++i) mx = t.size[i] > mx ? t.size[i] : mx;
  if (t.n <= 0 || mx <= 0) return 0;
  int gx = cdiv(mx, 1024);
  if (gx > 148 * 4) gx = 148 * 4;
  rmsprop_ema_kernel<<<dim3(gx, t.n), 256, 0, st>>>(t, lr, alpha, eps, ema_w);
  NSVD_LAUNCH_CHECK();
  return 0;
}

// x[b][d] = sigma * N(0,1): counter-based (Philox-style mixing of (seed, index)) + Box-Muller; one thread
// per point (2 coordinates).  Reproducible for a given (seed, offset); NOT the torch CPU stream
// (main_pde.py:92-93 stays the parity mode).
__device__ __forceinline__ uint32_t mix32(uint64_t z) {   // splitmix64 finaliser, upper bits
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
__global__ void sample_gaussian2_kernel(float* __restrict__ x, long n, float sigma, uint64_t seed, uint64_t offset) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t ctr = (offset + (uint64_t)i) * 2ull;
  uint32_t a = mix32(seed ^ (ctr * 0xD1342543DE82EF95ull));
  uint32_t b = mix32((seed + 0x632BE59BD9B4E019ull) ^ ((ctr + 1ull) * 0xD1342543DE82EF95ull));
  // u1 strictly inside (0, 1): u1 == 1 would give r == 0, i.e. a point exactly at the origin where the hydrogen
  // potential is infinite (a (float)a + 1 construction rounds to 1.0 with probability 3e-8 per sample; a soak run
  // of 2e7 samples hit it).  23-bit mid-point grid (k + 0.5 is exact in fp32 for k < 2^23): u1 <= 1 - 2^-24,
  // r in [3.5e-4, 5.8] sigma.
  float u1 = ((float)(a >> 9) + 0.5f) * 1.1920928955078125e-07f;
  float u2 = (float)(b >> 8) * 5.9604644775390625e-08f;     // [0, 1)
  float r = sigma * sqrtf(-2.f * logf(u1));
  float sn, cs;
  sincospif(2.f * u2, &sn, &cs);
  x[2 * i] = r * cs;
  x[2 * i + 1] = r * sn;
}
// Laplace (inverse CDF: x = -b sign(u) ln(1 - 2|u|), u in (-1/2, 1/2)) and uniform [-s, s) samplers
// (main_pde.py:101-118), one thread per point, same (seed, offset) counter scheme as the Gaussian one.
__global__ void sample_other2_kernel(float* __restrict__ x, long n, int laplace, float scale, uint64_t seed,
                                     uint64_t offset) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t ctr = (offset + (uint64_t)i) * 2ull;
  uint32_t r[2] = {mix32(seed ^ (ctr * 0xD1342543DE82EF95ull)),
                   mix32((seed + 0x632BE59BD9B4E019ull) ^ ((ctr + 1ull) * 0xD1342543DE82EF95ull))};
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    // 23-bit mid-point grid: u in [2^-24, 1 - 2^-24], never 0 or 1
    float u = ((float)(r[d] >> 9) + 0.5f) * 1.1920928955078125e-07f;
    float v;
    if (laplace) {
      float c = u - 0.5f;                                  // (-1/2, 1/2)
      v = -scale * copysignf(1.f, c) * log1pf(-2.f * fabsf(c));
    } else {
      v = scale * (2.f * u - 1.f);
    }
    x[2 * i + d] = v;
  }
}
int sample_other2(float* x, long n, int laplace, float scale, uint64_t seed, uint64_t offset, cudaStream_t st) {
  if (n <= 0) return 0;
  sample_other2_kernel<<<cdiv(n, 256), 256, 0, st>>>(x, n, laplace, scale, seed, offset);
  NSVD_LAUNCH_CHECK();
  return 0;
}

int sample_gaussian2(float* x, long n, float sigma, uint64_t seed, uint64_t offset, cudaStream_t st) {
  if (n <= 0) return 0;
  sample_gaussian2_kernel<<<cdiv(n, 256), 256, 0, st>>>(x, n, sigma, seed, offset);
  NSVD_LAUNCH_CHECK();
  return 0;
}

}  // namespace nsvd
