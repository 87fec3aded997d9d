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
  float gq0, gq1, lapq;  // grad/laplacian of ln sqrt(w)   (mask part added per copy)
};

__device__ __forceinline__ PointGeom point_geom(float x0, float x1, const nsvd_problem_t& pb) {
  PointGeom g;
  g.x0 = x0;
  g.x1 = x1;
  float r2 = x0 * x0 + x1 * x1;
  g.r = sqrtf(r2);
  float s2 = pb.sampling_sigma * pb.sampling_sigma;
  // w = N(x; 0, sigma^2 I_2) through log_prob().exp() as main_pde.py:97-100
  float logw = -r2 / (2.f * s2) - logf(6.283185307179586f * s2);
  float sw = sqrtf(expf(logw));
  g.rho = sw / fmaxf(sw, 1e-5f);  // diff_ops.py:15-18
  float inv = -1.f / (2.f * s2);
  g.gq0 = x0 * inv;
  g.gq1 = x1 * inv;
  g.lapq = 2.f * inv;  // -D/(2 sigma^2), D = 2
  g.V = pb.potential == NSVD_POT_HYDROGEN ? -(pb.pot_coef / g.r) : pb.pot_coef * (g.r * g.r);
  return g;
}

// (F, TF) of one (point, copy) from the raw network streams u = (value, d1, d2, lap).
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
  float cm = pb.hard_mul_const * m * g.rho;
  float lap = cm * (u3 + 2.f * (gq0 * u1 + gq1 * u2) + u0 * (lapq + gq0 * gq0 + gq1 * gq1));
  f = cm * u0;
  float negH = pb.scale_kinetic * lap - g.V * f;   // schrodinger/__init__.py:19-22
  tf = pb.op_scale * negH + pb.op_shift * f;       // examples/__init__.py:9
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace nsvd
