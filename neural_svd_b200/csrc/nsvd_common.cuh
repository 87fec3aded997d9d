// nsvd_common.cuh — shared helpers for the NestedLoRA sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nsvd.h"

namespace nsvd {

constexpr int kHidden = 128;   // hidden width of the eigenfunction MLPs (mlp_hidden_dims='128,128,128')
constexpr int kStreams = 4;    // value, d/dx1, d/dx2, Laplacian (D = 2)

void set_error(const char* fmt, ...);

#define NSVD_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      ::nsvd::set_error(__VA_ARGS__);             \
      return NSVD_E_BADARG;                       \
    }                                             \
  } while (0)

#define NSVD_CUDA(call)                                                      \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      ::nsvd::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,        \
                        cudaGetErrorString(e_));                             \
      return (int)e_;                                                        \
    }                                                                        \
  } while (0)

// every kernel launch of the library goes through this macro: it counts launches (nsvd_launch_count)
#define NSVD_LAUNCH_CHECK()               \
  do {                                    \
    ::nsvd::g_launches++;                 \
    NSVD_CUDA(cudaGetLastError());        \
  } while (0)

extern long g_launches;

// Opt a kernel into more than 48 KB of dynamic shared memory, once per DEVICE (the attribute is per device: a flag
// per process would leave the second GPU of a process without the opt-in).
constexpr int kMaxDevices = 64;
#define NSVD_SMEM_OPTIN(kern, bytes)                                                                   \
  do {                                                                                                 \
    static bool done_[::nsvd::kMaxDevices] = {};                                                       \
    int dev_ = 0;                                                                                      \
    NSVD_CUDA(cudaGetDevice(&dev_));                                                                   \
    if (dev_ < 0 || dev_ >= ::nsvd::kMaxDevices || !done_[dev_]) {                                     \
      NSVD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)));     \
      if (dev_ >= 0 && dev_ < ::nsvd::kMaxDevices) done_[dev_] = true;                                 \
    }                                                                                                  \
  } while (0)

// kernel classes timed by the optional profiler (nsvd_profile_*): CUDA events on the launch stream
enum KernelClass { KC_L0_FWD = 0, KC_HID_FWD, KC_HID_BWD, KC_L0_WGRAD, KC_GRAM, KC_DF, KC_PREP, KC_HEAD_BWD, KC_COUNT };
struct ProfScope {
  int cls;
  cudaStream_t st;
  bool on;
  ProfScope(int cls_, cudaStream_t st_);
  ~ProfScope();
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// softplus(beta=1, threshold=20) and its sigmoid, as torch.nn.Softplus (mlp.py:86).
__device__ __forceinline__ void softplus_sig(float z, float& a, float& sig) {
  if (z > 20.f) {
    a = z;
    sig = 1.f;
  } else {
    float e = expf(-fabsf(z));
    a = fmaxf(z, 0.f) + log1pf(e);
    float inv = 1.f / (1.f + e);
    sig = z >= 0.f ? inv : e * inv;
  }
}

// sigmoid(z) recovered from a = softplus(z):  1 - exp(-a)
__device__ __forceinline__ float sig_from_softplus(float a) { return -expm1f(-a); }

// Per-point quantities of the operator epilogue: shared by forward and backward.
struct PointGeom {
  float x0, x1, r, rho, V;
  float gq0, gq1, lapq;            // grad / Laplacian of ln sqrt(w)   (exp-mask part added per copy)
  float mb, gmb0, gmb1, lapmb;     // Dirichlet box mask and its derivatives (1, 0, 0, 0 without one)
};

__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) - (v < 0.f); }   // torch.sign / d|x|/dx

// one factor of DirichletBoundaryMaskBox (pde/boundary.py:16-37): value, first and second derivative as autograd
// sees them (clamp and maximum pass no gradient outside the box)
__device__ __forceinline__ void box_factor(float x, float lim, int mode, float& m, float& d1, float& d2) {
  const bool inside = x >= -lim && x <= lim;
  const float xc = fminf(fmaxf(x, -lim), lim);
  if (mode == NSVD_BOX_SQRT) {           // max((sqrt(2 lim^2 - x^2) - lim) / lim, 0)
    const float q = 2.f * lim * lim - xc * xc, sq = sqrtf(q);
    const float t = (sq - lim) / lim;
    const bool on = inside && t > 0.f;
    m = fmaxf(t, 0.f);
    d1 = on ? -xc / (lim * sq) : 0.f;
    d2 = on ? -2.f * lim / (q * sq) : 0.f;
  } else {                               // (1 - exp(-(lim - x))) (1 - exp(-(x + lim)))
    const float a = expf(xc - lim), b = expf(-xc - lim);
    m = (1.f - a) * (1.f - b);
    d1 = inside ? b - a : 0.f;
    d2 = inside ? -(a + b) : 0.f;
  }
}

__device__ __forceinline__ PointGeom point_geom(float x0, float x1, const nsvd_problem_t& pb) {
  PointGeom g;
  g.x0 = x0;
  g.x1 = x1;
  float r2 = x0 * x0 + x1 * x1;
  g.r = sqrtf(r2);
  const float sg = pb.sampling_sigma;
  if (pb.importance == NSVD_IMP_GAUSSIAN) {
    // w = N(x; 0, sigma^2 I_2) through log_prob().exp() as main_pde.py:97-100
    float s2 = sg * sg;
    float logw = -r2 / (2.f * s2) - logf(6.283185307179586f * s2);
    float sw = sqrtf(expf(logw));
    g.rho = sw / fmaxf(sw, 1e-5f);  // diff_ops.py:15-18
    float inv = -1.f / (2.f * s2);
    g.gq0 = x0 * inv;
    g.gq1 = x1 * inv;
    g.lapq = 2.f * inv;  // -D/(2 sigma^2), D = 2
  } else if (pb.importance == NSVD_IMP_LAPLACE) {
    // w = prod_i exp(-|x_i|/b) / (2b)  (main_pde.py:101-112); d|x|/dx = sign(x), second derivative 0
    float logw = -(fabsf(x0) + fabsf(x1)) / sg - 2.f * logf(2.f * sg);
    float sw = sqrtf(expf(logw));
    g.rho = sw / fmaxf(sw, 1e-5f);
    g.gq0 = -sgnf(x0) / (2.f * sg);
    g.gq1 = -sgnf(x1) / (2.f * sg);
    g.lapq = 0.f;
  } else {
    // uniform on [-s, s]^2: w = (2s)^-2 (main_pde.py:113-118); NSVD_IMP_NONE: no re-weighting at all
    float sw = pb.importance == NSVD_IMP_UNIFORM ? sqrtf(1.f / (4.f * sg * sg)) : 1.f;
    g.rho = sw / fmaxf(sw, pb.importance == NSVD_IMP_UNIFORM ? 1e-5f : 0.f);
    g.gq0 = g.gq1 = g.lapq = 0.f;
  }
  switch (pb.potential) {                                   // schrodinger/potentials.py
    case NSVD_POT_HYDROGEN: g.V = -(pb.pot_coef / g.r); break;                      // :5-8
    case NSVD_POT_HARMONIC: g.V = pb.pot_coef * (g.r * g.r); break;                 // :24-27
    case NSVD_POT_HYDROGEN_MOL_ION: {                                               // :11-17, nuclei at (0, +-R)
      float ym = x1 - pb.pot_coef2, yp = x1 + pb.pot_coef2;
      g.V = -(pb.pot_coef / sqrtf(x0 * x0 + ym * ym)) - (pb.pot_coef / sqrtf(x0 * x0 + yp * yp));
      break;
    }
    case NSVD_POT_COSINE: g.V = pb.pot_coef * cosf(x0) + pb.pot_coef2 * cosf(x1); break;   // :30-31
    default: g.V = 0.f; break;                                                      // infinite well, :20-21
  }
  g.mb = 1.f;
  g.gmb0 = g.gmb1 = g.lapmb = 0.f;
  if (pb.box_mask != NSVD_BOX_NONE) {
    float m0, m1, a0, a1, c0, c1;
    box_factor(x0, pb.box_lim, pb.box_mask, m0, a0, c0);
    box_factor(x1, pb.box_lim, pb.box_mask, m1, a1, c1);
    g.mb = m0 * m1;
    g.gmb0 = a0 * m1;
    g.gmb1 = m0 * a1;
    g.lapmb = c0 * m1 + m0 * c1;
  }
  return g;
}

// value-stream factor f = cm * u0 (and df/du0 in the backward): hard_mul_const * exp mask * box mask * rho
__device__ __forceinline__ float head_factor(const PointGeom& g, const nsvd_problem_t& pb, float mexp) {
  return pb.hard_mul_const * mexp * g.rho * g.mb;
}

// (F, TF) of one (point, copy) from the raw network streams u = (value, d1, d2, lap): product rule on
// q u with q = sqrt(w) * exp-mask_l * box-mask, written with Q = ln(sqrt(w) exp-mask) and the box mask explicit
// (it may vanish).
__device__ __forceinline__ void operator_epilogue(const PointGeom& g, const nsvd_problem_t& pb,
                                                  bool has_mask, float mscale, float u0, float u1,
                                                  float u2, float u3, float& f, float& tf) {
  float gq0 = g.gq0, gq1 = g.gq1, lapq = g.lapq, m = 1.f;
  if (has_mask) {
    m = expf(-g.r / mscale);          // boundary.py:48-49
    float irs = 1.f / (g.r * mscale);
    gq0 -= g.x0 * irs;
    gq1 -= g.x1 * irs;
    lapq -= irs;                      // (D-1)/(r s), D = 2
  }
  float ce = pb.hard_mul_const * m * g.rho;
  float inner = u3 + 2.f * (gq0 * u1 + gq1 * u2) + u0 * (lapq + gq0 * gq0 + gq1 * gq1);
  inner = g.mb * inner + 2.f * (g.gmb0 * (u1 + u0 * gq0) + g.gmb1 * (u2 + u0 * gq1)) + u0 * g.lapmb;
  float lap = ce * inner;
  f = ce * g.mb * u0;
  float negH = pb.scale_kinetic * lap - g.V * f;   // schrodinger/__init__.py:19-22
  tf = pb.op_scale * negH + pb.op_shift * f;       // examples/__init__.py:9
}

// ---- finite-difference Laplacian (laplacian_eps > 0: VectorizedLaplacian.approx_laplacian, pde/diff_ops.py:25-52) ----
// un-clamped sqrt(w(y)) of the sampler's density; 1 without importance (diff_ops.py:10-13)
__device__ __forceinline__ float sqrt_w(float y0, float y1, const nsvd_problem_t& pb) {
  const float sg = pb.sampling_sigma;
  if (pb.importance == NSVD_IMP_GAUSSIAN) {
    const float s2 = sg * sg;
    return sqrtf(expf(-(y0 * y0 + y1 * y1) / (2.f * s2) - logf(6.283185307179586f * s2)));
  }
  if (pb.importance == NSVD_IMP_LAPLACE) return sqrtf(expf(-(fabsf(y0) + fabsf(y1)) / sg - 2.f * logf(2.f * sg)));
  if (pb.importance == NSVD_IMP_UNIFORM) return sqrtf(1.f / (4.f * sg * sg));
  return 1.f;
}
// g(y) = sqrt(w(y)) * hard_mul_const * exp-mask_l(y) * box-mask(y) * u : the function whose second differences are taken
__device__ __forceinline__ float weighted_value(float y0, float y1, const nsvd_problem_t& pb, bool has_mask,
                                                float mscale, float u) {
  float m = 1.f;
  if (has_mask) m = expf(-sqrtf(y0 * y0 + y1 * y1) / mscale);
  if (pb.box_mask != NSVD_BOX_NONE) {
    float m0, m1, a, c;
    box_factor(y0, pb.box_lim, pb.box_mask, m0, a, c);
    box_factor(y1, pb.box_lim, pb.box_mask, m1, a, c);
    m *= m0 * m1;
  }
  return sqrt_w(y0, y1, pb) * (pb.hard_mul_const * m * u);
}
// TF of one (point, copy) from the raw network values at x (uc) and at x + eps e_0, x - eps e_0, x + eps e_1,
// x - eps e_1 (us[0..3]); the accumulation order is the reference's (diff_ops.py:40-48)
__device__ __forceinline__ float fd_operator(float x0, float x1, const nsvd_problem_t& pb, bool has_mask, float mscale,
                                             float uc, const float* us) {
  const float eps = pb.fd_eps;
  const float gc = weighted_value(x0, x1, pb, has_mask, mscale, uc);
  float lap = -4.f * gc;
  lap += weighted_value(x0 + eps, x1, pb, has_mask, mscale, us[0]) + weighted_value(x0 - eps, x1, pb, has_mask, mscale, us[1]);
  lap += weighted_value(x0, x1 + eps, pb, has_mask, mscale, us[2]) + weighted_value(x0, x1 - eps, pb, has_mask, mscale, us[3]);
  lap = lap / (float)((double)eps * (double)eps);
  float fs = gc;
  if (pb.importance != NSVD_IMP_NONE) {
    const float sw = fmaxf(sqrt_w(x0, x1, pb), 1e-5f);          // diff_ops.py:15
    lap = lap / sw;
    fs = gc / sw;
  }
  PointGeom g = point_geom(x0, x1, pb);
  const float negH = pb.scale_kinetic * lap - g.V * fs;          // schrodinger/__init__.py:19-22
  return pb.op_scale * negH + pb.op_shift * fs;                  // examples/__init__.py:9
}
// coordinates of shifted point set s (0: +e_0, 1: -e_0, 2: +e_1, 3: -e_1)
__device__ __forceinline__ void fd_shift(int s, float eps, float& y0, float& y1) {
  const float d = (s & 1) ? -eps : eps;
  if (s < 2) y0 += d;
  else y1 += d;
}

// ------------------------------------------------------------------------------------------
// D-dimensional versions (D = 2 or 3; pb.ndim, 0 meaning 2) of the per-point operator math above, used by the fp32
// CUDA-core engine: S = D + 2 forward-mode streams (value, d/dx_1 .. d/dx_D, Laplacian).  ndim = 3 is what
// pde/problems.py:62-71 runs for hydrogen and the H2+ ion.
// ------------------------------------------------------------------------------------------
constexpr int kMaxDim = 3;
__host__ __device__ __forceinline__ int problem_ndim(const nsvd_problem_t& pb) { return pb.ndim == 3 ? 3 : 2; }

struct PointGeomN {
  int D;
  float x[kMaxDim], r, rho, V;
  float gq[kMaxDim], lapq;          // grad / Laplacian of ln sqrt(w)   (exp-mask part added per copy)
  float mb, gmb[kMaxDim], lapmb;    // Dirichlet box mask and its derivatives (1, 0, 0 without one)
};

__device__ __forceinline__ float sqrt_w_n(const float* y, int D, const nsvd_problem_t& pb) {
  const float sg = pb.sampling_sigma;
  float r2 = 0.f, a1 = 0.f;
#pragma unroll
  for (int d = 0; d < kMaxDim; ++d) {
    const float yd = d < D ? y[d] : 0.f;
    r2 = fmaf(yd, yd, r2);
    a1 += fabsf(yd);
  }
  if (pb.importance == NSVD_IMP_GAUSSIAN) {
    const float s2 = sg * sg;
    return sqrtf(expf(-r2 / (2.f * s2) - 0.5f * D * logf(6.283185307179586f * s2)));
  }
  if (pb.importance == NSVD_IMP_LAPLACE) return sqrtf(expf(-a1 / sg - D * logf(2.f * sg)));
  if (pb.importance == NSVD_IMP_UNIFORM) return sqrtf(D == 3 ? 1.f / (8.f * sg * sg * sg) : 1.f / (4.f * sg * sg));
  return 1.f;
}

// y must have kMaxDim entries (zero beyond D)
__device__ __forceinline__ float potential_n(const float* x, int D, float r, const nsvd_problem_t& pb) {
  switch (pb.potential) {                                   // schrodinger/potentials.py
    case NSVD_POT_HYDROGEN: return -(pb.pot_coef / r);                              // :5-8
    case NSVD_POT_HARMONIC: return pb.pot_coef * (r * r);                           // :24-27
    case NSVD_POT_HYDROGEN_MOL_ION: {                                               // :11-17, nuclei at (0, .., +-R)
      const float last = D == 3 ? x[2] : x[1];
      const float q = D == 3 ? fmaf(x[1], x[1], x[0] * x[0]) : x[0] * x[0];
      const float ym = last - pb.pot_coef2, yp = last + pb.pot_coef2;
      return -(pb.pot_coef / sqrtf(q + ym * ym)) - (pb.pot_coef / sqrtf(q + yp * yp));
    }
    case NSVD_POT_COSINE: return pb.pot_coef * cosf(x[0]) + pb.pot_coef2 * cosf(x[1]);   // :30-31 (2D only)
    default: return 0.f;                                                            // infinite well, :20-21
  }
}

__device__ __forceinline__ PointGeomN point_geom_n(const float* xs, const nsvd_problem_t& pb) {
  PointGeomN g;
  const int D = problem_ndim(pb);
  g.D = D;
  g.x[0] = xs[0];
  g.x[1] = xs[1];
  g.x[2] = D == 3 ? xs[2] : 0.f;
  g.r = sqrtf(fmaf(g.x[2], g.x[2], fmaf(g.x[1], g.x[1], g.x[0] * g.x[0])));
  const float sg = pb.sampling_sigma;
  const float sw = sqrt_w_n(g.x, D, pb);
  g.lapq = 0.f;
  g.gq[0] = g.gq[1] = g.gq[2] = 0.f;
  if (pb.importance == NSVD_IMP_GAUSSIAN) {
    g.rho = sw / fmaxf(sw, 1e-5f);  // diff_ops.py:15-18
    const float inv = -1.f / (2.f * sg * sg);
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) g.gq[d] = g.x[d] * inv;     // x[2] = 0 in 2D
    g.lapq = D * inv;
  } else if (pb.importance == NSVD_IMP_LAPLACE) {
    g.rho = sw / fmaxf(sw, 1e-5f);
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) g.gq[d] = d < D ? -sgnf(g.x[d]) / (2.f * sg) : 0.f;
  } else {
    g.rho = sw / fmaxf(sw, pb.importance == NSVD_IMP_UNIFORM ? 1e-5f : 0.f);
  }
  g.V = potential_n(g.x, D, g.r, pb);
  g.mb = 1.f;
  g.lapmb = 0.f;
  g.gmb[0] = g.gmb[1] = g.gmb[2] = 0.f;
  if (pb.box_mask != NSVD_BOX_NONE) {
    float m[kMaxDim] = {1.f, 1.f, 1.f}, a[kMaxDim] = {0.f, 0.f, 0.f}, c[kMaxDim] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d)
      if (d < D) box_factor(g.x[d], pb.box_lim, pb.box_mask, m[d], a[d], c[d]);
    g.mb = m[0] * m[1] * m[2];
    g.gmb[0] = a[0] * (m[1] * m[2]);
    g.gmb[1] = a[1] * (m[0] * m[2]);
    g.gmb[2] = a[2] * (m[0] * m[1]);
    g.lapmb = c[0] * (m[1] * m[2]) + c[1] * (m[0] * m[2]) + c[2] * (m[0] * m[1]);
  }
  return g;
}

__device__ __forceinline__ float head_factor_n(const PointGeomN& g, const nsvd_problem_t& pb, float mexp) {
  return pb.hard_mul_const * mexp * g.rho * g.mb;
}

// (F, TF) of one (point, copy) from the raw network streams u[0] = value, u[1 .. D] = gradient, u[D + 1] = Laplacian.
// All loops run over kMaxDim with a predicate, so every array index is a compile-time constant (registers, no local
// memory): the version with run-time trip counts was mis-evaluated by the device compiler in the exp-mask branch
// (caught by the golden tests; host build of the same source was right).
__device__ __forceinline__ void operator_epilogue_n(const PointGeomN& g, const nsvd_problem_t& pb, bool has_mask,
                                                    float mscale, const float* u, float& f, float& tf) {
  const int D = g.D;
  const float u0 = u[0], ulap = D == 3 ? u[4] : u[3];
  const float ud[kMaxDim] = {u[1], u[2], D == 3 ? u[3] : 0.f};
  float m = 1.f, irs = 0.f;
  if (has_mask) {
    m = expf(-g.r / mscale);          // boundary.py:48-49
    irs = 1.f / (g.r * mscale);
  }
  const float lapq = g.lapq - (float)(D - 1) * irs;
  float dot = 0.f, gq2 = 0.f, box = 0.f;
#pragma unroll
  for (int d = 0; d < kMaxDim; ++d) {
    const float gqd = d < D ? g.gq[d] - g.x[d] * irs : 0.f;
    const float udd = d < D ? ud[d] : 0.f;
    dot = fmaf(gqd, udd, dot);
    gq2 = fmaf(gqd, gqd, gq2);
    box = fmaf(d < D ? g.gmb[d] : 0.f, udd + u0 * gqd, box);
  }
  const float ce = pb.hard_mul_const * m * g.rho;
  float inner = ulap + 2.f * dot + u0 * (lapq + gq2);
  inner = g.mb * inner + 2.f * box + u0 * g.lapmb;
  const float lap = ce * inner;
  f = ce * g.mb * u0;
  const float negH = pb.scale_kinetic * lap - g.V * f;   // schrodinger/__init__.py:19-22
  tf = pb.op_scale * negH + pb.op_shift * f;             // examples/__init__.py:9
}

// g(y) = sqrt(w(y)) * hard_mul_const * exp-mask_l(y) * box-mask(y) * u   (y has kMaxDim entries, zero beyond D)
__device__ __forceinline__ float weighted_value_n(const float* y, int D, const nsvd_problem_t& pb, bool has_mask,
                                                  float mscale, float u) {
  float m = 1.f;
  if (has_mask) m = expf(-sqrtf(fmaf(y[2], y[2], fmaf(y[1], y[1], y[0] * y[0]))) / mscale);
  if (pb.box_mask != NSVD_BOX_NONE) {
#pragma unroll
    for (int d = 0; d < kMaxDim; ++d) {
      if (d < D) {
        float md, a, c;
        box_factor(y[d], pb.box_lim, pb.box_mask, md, a, c);
        m *= md;
      }
    }
  }
  return sqrt_w_n(y, D, pb) * (pb.hard_mul_const * m * u);
}
// TF of one (point, copy) from the raw network values at x (uc) and at the 2 D shifted points x + eps e_d (us[2 d]),
// x - eps e_d (us[2 d + 1]); the accumulation order is the reference's (diff_ops.py:40-48)
__device__ __forceinline__ float fd_operator_n(const float* xs, const nsvd_problem_t& pb, bool has_mask, float mscale,
                                               float uc, const float* us) {
  const int D = problem_ndim(pb);
  const float eps = pb.fd_eps;
  const float x[kMaxDim] = {xs[0], xs[1], D == 3 ? xs[2] : 0.f};
  const float gc = weighted_value_n(x, D, pb, has_mask, mscale, uc);
  float lap = -2.f * D * gc;
#pragma unroll
  for (int d = 0; d < kMaxDim; ++d) {
    if (d < D) {
      const float yp[kMaxDim] = {d == 0 ? x[0] + eps : x[0], d == 1 ? x[1] + eps : x[1], d == 2 ? x[2] + eps : x[2]};
      const float ym[kMaxDim] = {d == 0 ? x[0] - eps : x[0], d == 1 ? x[1] - eps : x[1], d == 2 ? x[2] - eps : x[2]};
      lap += weighted_value_n(yp, D, pb, has_mask, mscale, us[2 * d]) +
             weighted_value_n(ym, D, pb, has_mask, mscale, us[2 * d + 1]);
    }
  }
  lap = lap / (float)((double)eps * (double)eps);
  float fs = gc;
  if (pb.importance != NSVD_IMP_NONE) {
    const float sw = fmaxf(sqrt_w_n(x, D, pb), 1e-5f);          // diff_ops.py:15
    lap = lap / sw;
    fs = gc / sw;
  }
  const float r = sqrtf(fmaf(x[2], x[2], fmaf(x[1], x[1], x[0] * x[0])));
  const float V = potential_n(x, D, r, pb);
  const float negH = pb.scale_kinetic * lap - V * fs;            // schrodinger/__init__.py:19-22
  return pb.op_scale * negH + pb.op_shift * fs;                  // examples/__init__.py:9
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace nsvd
