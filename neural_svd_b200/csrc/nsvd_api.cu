// nsvd_api.cu — extern "C" entry points of libnsvd.so (see include/nsvd.h).
#include <stdarg.h>
#include <string.h>
#include "nsvd_simt.cuh"

#include <vector>
namespace nsvd {
long g_launches = 0;
static bool g_prof_on = false;
struct ProfRec { int cls; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;
ProfScope::ProfScope(int cls_, cudaStream_t st_) : cls(cls_), st(st_), on(g_prof_on) {
  if (!on) return;
  ProfRec r;
  r.cls = cls;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (on) cudaEventRecord(g_prof.back().b, st);
}
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
// Every entry point runs on the device that owns its buffers, whichever device is current in the calling thread
// (a model on cuda:1 while cuda:0 is current must not launch on cuda:0); the previous device is restored on return.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const void* p) {
    cudaPointerAttributes a;
    if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice) {
      if (cudaGetDevice(&prev) == cudaSuccess && prev != a.device) switched = cudaSetDevice(a.device) == cudaSuccess;
    } else {
      cudaGetLastError();   // not a device pointer known to this process: leave the current device alone
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

static int check_problem(const nsvd_problem_t* pb) {
  NSVD_CHECK_ARG(pb != nullptr, "problem is NULL");
  NSVD_CHECK_ARG(pb->n_points >= 2, "n_points must be >= 2 (got %d)", pb->n_points);
  NSVD_CHECK_ARG(pb->n_copies >= 1 && pb->n_copies <= 64, "n_copies must be in [1,64] (got %d)", pb->n_copies);
  NSVD_CHECK_ARG(pb->hidden == kHidden, "hidden must be %d (got %d)", kHidden, pb->hidden);
  NSVD_CHECK_ARG(pb->n_fourier >= 8 && pb->n_fourier % 8 == 0, "n_fourier must be a positive multiple of 8 (got %d)", pb->n_fourier);
  NSVD_CHECK_ARG(pb->potential >= NSVD_POT_HYDROGEN && pb->potential <= NSVD_POT_COSINE, "unknown potential %d", pb->potential);
  NSVD_CHECK_ARG(pb->importance >= NSVD_IMP_GAUSSIAN && pb->importance <= NSVD_IMP_NONE, "unknown importance %d", pb->importance);
  NSVD_CHECK_ARG(pb->importance == NSVD_IMP_NONE || pb->sampling_sigma > 0.f, "sampling_sigma must be > 0");
  NSVD_CHECK_ARG(pb->box_mask >= NSVD_BOX_NONE && pb->box_mask <= NSVD_BOX_EXP, "unknown box mask mode %d", pb->box_mask);
  NSVD_CHECK_ARG(pb->box_mask == NSVD_BOX_NONE || pb->box_lim > 0.f, "box_lim must be > 0");
  NSVD_CHECK_ARG(pb->fd_eps == pb->fd_eps && pb->fd_eps < 1e30f, "fd_eps must be finite (got %f)", (double)pb->fd_eps);
  NSVD_CHECK_ARG(pb->ndim == 0 || pb->ndim == 2 || pb->ndim == 3, "ndim must be 2 or 3 (got %d)", pb->ndim);
  NSVD_CHECK_ARG(pb->ndim != 3 || pb->potential == NSVD_POT_HYDROGEN || pb->potential == NSVD_POT_HYDROGEN_MOL_ION ||
                     pb->potential == NSVD_POT_HARMONIC,
                 "ndim = 3 is defined for the hydrogen, H2+ ion and harmonic potentials (got potential %d)", pb->potential);
  return 0;
}
static int check_params(const nsvd_problem_t* pb, const nsvd_params_t* pr) {
  NSVD_CHECK_ARG(pr != nullptr && pr->Bff != nullptr, "params / Bff is NULL");
  for (int i = 0; i < 4; ++i) NSVD_CHECK_ARG(pr->W[i] && pr->b[i], "W[%d] or b[%d] is NULL", i, i);
  NSVD_CHECK_ARG(!pb->has_exp_mask || pr->mask_scales, "has_exp_mask set but mask_scales is NULL");
  return 0;
}
}  // namespace nsvd

using namespace nsvd;

extern "C" {

int nsvd_abi_version(void) { return NSVD_ABI_VERSION; }
#ifndef NSVD_SRC_HASH
#define NSVD_SRC_HASH "unknown"
#endif
const char* nsvd_build_hash(void) { return NSVD_SRC_HASH; }
size_t nsvd_struct_size(int32_t which) {
  return which == 0 ? sizeof(nsvd_problem_t) : which == 1 ? sizeof(nsvd_params_t) : which == 2 ? sizeof(nsvd_grads_t) : 0;
}
long nsvd_launch_count(void) { return g_launches; }
void nsvd_set_tc_microbatch(int32_t points) { tc_set_micro_batch(points); }
void nsvd_profile_enable(int on) { g_prof_on = on != 0; }
int nsvd_profile_read(double* ms_per_class, long* launches_per_class, int n_classes, int reset) {
  for (int i = 0; i < n_classes; ++i) {
    ms_per_class[i] = 0.0;
    launches_per_class[i] = 0;
  }
  for (auto& r : g_prof) {
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e != cudaSuccess) {
      set_error("profile: %s", cudaGetErrorString(e));
      return (int)e;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (r.cls < n_classes) {
      ms_per_class[r.cls] += ms;
      launches_per_class[r.cls] += 1;
    }
  }
  if (reset) {
    for (auto& r : g_prof) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    g_prof.clear();
  }
  return 0;
}
const char* nsvd_last_error(void) { return g_err; }

int nsvd_device_ok(int dev) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    return NSVD_E_NODEVICE;
  }
  if (p.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", dev, p.major, p.minor);
    return NSVD_E_NODEVICE;
  }
  return 0;
}

int nsvd_scratch_bytes(const nsvd_problem_t* pb, int engine, size_t* saved_bytes, size_t* work_bytes) {
  int rc = check_problem(pb);
  if (rc) return rc;
  NSVD_CHECK_ARG(saved_bytes && work_bytes, "output pointers are NULL");
  if (engine == NSVD_ENGINE_FP32_SIMT) simt_scratch_bytes(*pb, saved_bytes, work_bytes);
  else if (engine == NSVD_ENGINE_BF16X3_TC) {
    NSVD_CHECK_ARG(pb->ndim != 3, "ndim = 3 (five forward-mode streams) runs on NSVD_ENGINE_FP32_SIMT only");
    tc_scratch_bytes(*pb, saved_bytes, work_bytes);
  } else NSVD_CHECK_ARG(false, "unknown engine %d", engine);
  return 0;
}

int nsvd_fwd_streams(const nsvd_problem_t* pb, const nsvd_params_t* pr, int engine, const float* x,
                     float* F, float* TF, void* saved, size_t saved_bytes, void* work,
                     size_t work_bytes, void* stream) {
  DeviceGuard dg_(x);
  int rc = check_problem(pb);
  if (rc) return rc;
  if ((rc = check_params(pb, pr))) return rc;
  NSVD_CHECK_ARG(x && F && TF && saved && work, "NULL buffer");
  size_t ns, nw;
  if ((rc = nsvd_scratch_bytes(pb, engine, &ns, &nw))) return rc;
  if (saved_bytes < ns || work_bytes < nw) {
    set_error("scratch too small: saved %zu < %zu or work %zu < %zu", saved_bytes, ns, work_bytes, nw);
    return NSVD_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (engine == NSVD_ENGINE_FP32_SIMT) return simt_forward(*pb, *pr, x, F, TF, saved, work, st);
  return tc_forward(*pb, *pr, x, F, TF, saved, work, work_bytes, st);
}

int nsvd_mlp_bwd(const nsvd_problem_t* pb, const nsvd_params_t* pr, int engine, const float* x,
                 const float* dF, const void* saved, size_t saved_bytes, nsvd_grads_t* gr, void* work,
                 size_t work_bytes, void* stream) {
  DeviceGuard dg_(dF);
  int rc = check_problem(pb);
  if (rc) return rc;
  if ((rc = check_params(pb, pr))) return rc;
  NSVD_CHECK_ARG(x && dF && saved && work && gr, "NULL buffer");
  for (int i = 0; i < 4; ++i) NSVD_CHECK_ARG(gr->dW[i] && gr->db[i], "dW[%d] or db[%d] is NULL", i, i);
  size_t ns, nw;
  if ((rc = nsvd_scratch_bytes(pb, engine, &ns, &nw))) return rc;
  if (saved_bytes < ns || work_bytes < nw) {
    set_error("scratch too small: saved %zu < %zu or work %zu < %zu", saved_bytes, ns, work_bytes, nw);
    return NSVD_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (engine == NSVD_ENGINE_FP32_SIMT) return simt_backward(*pb, *pr, x, dF, saved, *gr, work, st);
  return tc_backward(*pb, *pr, x, dF, saved, *gr, work, work_bytes, st);
}

size_t nsvd_gram_partials_bytes(int32_t n_points, int32_t n_copies) {
  return gram_partials_bytes(n_points, n_copies);
}

int nsvd_gram_reduce(const float* F, const float* TF, const float* vector_mask, int32_t n_points,
                     int32_t n_copies, int32_t b1, float* terms, void* partials, void* stream) {
  DeviceGuard dg_(F);
  NSVD_CHECK_ARG(F && TF && vector_mask && terms && partials, "NULL buffer");
  NSVD_CHECK_ARG(n_points >= 1 && n_copies >= 1 && n_copies <= 64, "bad shape B=%d L=%d", n_points, n_copies);
  NSVD_CHECK_ARG(b1 >= 0 && b1 <= n_points, "b1=%d out of range", b1);
  ProfScope ps(KC_GRAM, (cudaStream_t)stream);
  return gram_reduce(F, TF, vector_mask, n_points, n_copies, b1, terms, partials, (cudaStream_t)stream);
}

int nsvd_cross_gram(const float* F, const float* TF, const float* roww, const float* x, int32_t n_points,
                    int32_t n_copies, float* cov, float* quad, void* partials, void* stream) {
  DeviceGuard dg_(F);
  NSVD_CHECK_ARG(F && TF && cov && quad && partials, "NULL buffer");
  NSVD_CHECK_ARG(n_points >= 1 && n_copies >= 1 && n_copies <= 64, "bad shape B=%d L=%d", n_points, n_copies);
  return cross_gram(F, TF, roww, x, n_points, n_copies, cov, quad, partials, (cudaStream_t)stream);
}

int nsvd_loss_finalize(const float* terms, const float* matrix_mask, int32_t n_copies, int64_t Bg,
                       int64_t B1g, int64_t B2g, float* loss, float* coef, void* stream) {
  DeviceGuard dg_(terms);
  NSVD_CHECK_ARG(terms && matrix_mask && loss && coef, "NULL buffer");
  NSVD_CHECK_ARG(Bg <= 0 || (B1g > 0 && B2g > 0 && B1g + B2g == Bg), "bad counts B=%ld B1=%ld B2=%ld", (long)Bg, (long)B1g, (long)B2g);
  return loss_finalize(terms, matrix_mask, n_copies, Bg, B1g, B2g, loss, coef, (cudaStream_t)stream);
}

int nsvd_loss_dF(const float* F, const float* TF, const float* vector_mask, const float* coef,
                 const float* grad_scale, int32_t n_points, int32_t n_copies, int32_t b1, int64_t Bg,
                 float* dF, void* stream) {
  DeviceGuard dg_(F);
  NSVD_CHECK_ARG(F && vector_mask && dF && (TF || coef), "NULL buffer");
  NSVD_CHECK_ARG(b1 >= 0 && b1 <= n_points, "bad b1");
  NSVD_CHECK_ARG(n_copies >= 1 && n_copies <= 64, "n_copies %d out of [1,64]", n_copies);
  ProfScope ps(KC_DF, (cudaStream_t)stream);
  return loss_dF(F, TF, vector_mask, coef, grad_scale, n_points, n_copies, b1, Bg, dF, (cudaStream_t)stream);
}

size_t nsvd_cdk_work_bytes(int32_t n_rows, int32_t n_feat, int32_t first_const, int engine) {
  // + the tail reserved for nsvd_cdk_finalize (and slack to align it)
  return (engine == NSVD_ENGINE_BF16X3_TC ? tc_cdk_work_bytes(n_rows, n_feat, first_const)
                                          : cdk_work_bytes(n_rows, n_feat, first_const)) +
         NSVD_CDK_FINALIZE_SCRATCH + 16;
}
static int cdk_check(int engine, const void* work, size_t work_bytes, int n_rows, int n_feat, int fc) {
  NSVD_CHECK_ARG(engine == NSVD_ENGINE_FP32_SIMT || engine == NSVD_ENGINE_BF16X3_TC, "unknown engine %d", engine);
  NSVD_CHECK_ARG(n_rows >= 2 && n_feat >= 1 && (fc == 0 || fc == 1), "bad shape");
  NSVD_CHECK_ARG(work != nullptr, "work is NULL");
  if (work_bytes < nsvd_cdk_work_bytes(n_rows, n_feat, fc, engine)) {
    set_error("cdk work too small: %zu < %zu", work_bytes, nsvd_cdk_work_bytes(n_rows, n_feat, fc, engine));
    return NSVD_E_WORKSPACE;
  }
  return 0;
}
int nsvd_cdk_fwd(const float* f, const float* g, const float* vector_mask, int32_t n_rows, int32_t n_feat,
                 int32_t first_const, int engine, float* terms, float* rs_joint, void* work, size_t work_bytes,
                 void* stream) {
  DeviceGuard dg_(f);
  NSVD_CHECK_ARG(f && g && vector_mask && terms, "NULL buffer");
  int rc = cdk_check(engine, work, work_bytes, n_rows, n_feat, first_const);
  if (rc) return rc;
  if (engine == NSVD_ENGINE_BF16X3_TC)
    return tc_cdk_fwd(f, g, vector_mask, n_rows, n_feat, first_const, terms, rs_joint, work, (cudaStream_t)stream);
  return cdk_fwd(f, g, vector_mask, n_rows, n_feat, first_const, terms, rs_joint, work, (cudaStream_t)stream);
}
int nsvd_cdk_finalize(const float* terms, const float* matrix_mask, int32_t Lp, int64_t Bg, float* losses,
                      float* coef, void* scratch, void* stream) {
  DeviceGuard dg_(terms);
  NSVD_CHECK_ARG(terms && matrix_mask && losses && coef && Lp >= 1 && Bg >= 1, "bad args");
  NSVD_CHECK_ARG(scratch && ((uintptr_t)scratch & 7) == 0, "scratch must be 8-byte aligned device memory");
  return cdk_finalize(terms, matrix_mask, Lp, Bg, losses, coef, (double*)scratch, (cudaStream_t)stream);
}
int nsvd_cdk_bwd(const float* f, const float* g, const float* vector_mask, const float* coef,
                 const float* grad_scale, int32_t n_rows, int32_t n_feat, int32_t first_const, int64_t Bg, int engine,
                 float* grad_f, float* grad_g, void* work, size_t work_bytes, int32_t planes_ready, void* stream) {
  DeviceGuard dg_(f);
  NSVD_CHECK_ARG(f && g && vector_mask && coef && grad_f && grad_g, "NULL buffer");
  int rc = cdk_check(engine, work, work_bytes, n_rows, n_feat, first_const);
  if (rc) return rc;
  if (engine == NSVD_ENGINE_BF16X3_TC)
    return tc_cdk_bwd(f, g, vector_mask, coef, grad_scale, n_rows, n_feat, first_const, Bg, grad_f, grad_g, work,
                      planes_ready, (cudaStream_t)stream);
  return cdk_bwd(f, g, vector_mask, coef, grad_scale, n_rows, n_feat, first_const, Bg, grad_f, grad_g,
                 (cudaStream_t)stream);
}
int nsvd_cdk_offdiag(const float* f, const float* g, int32_t n_rows, int32_t n_feat, int32_t first_const, int engine,
                     float* rs_indep, void* work, size_t work_bytes, int32_t planes_ready, void* stream) {
  DeviceGuard dg_(f);
  NSVD_CHECK_ARG(f && g && rs_indep, "NULL buffer");
  int rc = cdk_check(engine, work, work_bytes, n_rows, n_feat, first_const);
  if (rc) return rc;
  if (engine == NSVD_ENGINE_BF16X3_TC)
    return tc_cdk_offdiag(f, g, n_rows, n_feat, first_const, rs_indep, work, planes_ready, (cudaStream_t)stream);
  return cdk_offdiag(f, g, n_rows, n_feat, first_const, rs_indep, (cudaStream_t)stream);
}

int nsvd_rmsprop_ema_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* square_avg,
                          float* const* ema, const int64_t* sizes, float lr, float alpha, float eps,
                          float ema_one_minus_decay, void* stream) {
  DeviceGuard dg_((n_tensors >= 1 && params) ? params[0] : nullptr);
  NSVD_CHECK_ARG(n_tensors >= 1 && n_tensors <= 16, "n_tensors must be in [1,16] (got %d)", n_tensors);
  NSVD_CHECK_ARG(params && grads && square_avg && sizes, "NULL table");
  OptTensors t{};
  t.n = n_tensors;
  for (int i = 0; i < n_tensors; ++i) {
    NSVD_CHECK_ARG(params[i] && grads[i] && square_avg[i] && sizes[i] >= 0, "tensor %d: NULL pointer or negative size", i);
    t.p[i] = params[i];
    t.g[i] = grads[i];
    t.sq[i] = square_avg[i];
    t.ema[i] = ema ? ema[i] : nullptr;
    t.size[i] = sizes[i];
  }
  return rmsprop_ema_step(t, lr, alpha, eps, ema_one_minus_decay, (cudaStream_t)stream);
}

int nsvd_sample_gaussian(float* x, int64_t n_points, float sigma, uint64_t seed, uint64_t offset, void* stream) {
  DeviceGuard dg_(x);
  NSVD_CHECK_ARG(x && n_points >= 0 && sigma > 0.f, "bad args");
  return sample_gaussian2(x, n_points, sigma, seed, offset, (cudaStream_t)stream);
}

int nsvd_sample_points(float* x, int64_t n_points, int32_t importance, float scale, uint64_t seed, uint64_t offset,
                       void* stream) {
  DeviceGuard dg_(x);
  NSVD_CHECK_ARG(x && n_points >= 0 && scale > 0.f, "bad args");
  if (importance == NSVD_IMP_GAUSSIAN) return sample_gaussian2(x, n_points, scale, seed, offset, (cudaStream_t)stream);
  NSVD_CHECK_ARG(importance == NSVD_IMP_LAPLACE || importance == NSVD_IMP_UNIFORM, "no sampler for importance %d",
                 importance);
  return sample_other2(x, n_points, importance == NSVD_IMP_LAPLACE, scale, seed, offset, (cudaStream_t)stream);
}

size_t nsvd_linear_work_bytes(int32_t rows, int32_t in_features, int32_t out_features) {
  return tc_linear_work_bytes(rows, in_features, out_features);
}
int nsvd_linear_fwd(const float* x, const float* W, const float* bias, float* y, int32_t rows, int32_t in_features,
                    int32_t out_features, int32_t act, float slope, void* work, size_t work_bytes, void* stream) {
  DeviceGuard dg_(x);
  NSVD_CHECK_ARG(x && W && y, "NULL buffer");
  NSVD_CHECK_ARG(act == 0 || act == 1, "unknown activation %d", act);
  return tc_linear_fwd(x, W, bias, y, rows, in_features, out_features, act, slope, work, work_bytes, (cudaStream_t)stream);
}
int nsvd_linear_bwd(const float* x, const float* W, const float* y, const float* dy, int32_t rows, int32_t in_features,
                    int32_t out_features, int32_t act, float slope, float* dx, float* dW, float* db, void* work,
                    size_t work_bytes, void* stream) {
  DeviceGuard dg_(dy);
  NSVD_CHECK_ARG(x && W && dy, "NULL buffer");
  NSVD_CHECK_ARG(act == 0 || act == 1, "unknown activation %d", act);
  return tc_linear_bwd(x, W, y, dy, rows, in_features, out_features, act, slope, dx, dW, db, work, work_bytes,
                       (cudaStream_t)stream);
}

int nsvd_tc_gemm_selftest(const float* A, const float* B, float* D, int32_t M, int32_t N, int32_t K,
                          int32_t a_kmajor, int32_t b_kmajor, void* work, size_t work_bytes, void* stream) {
  DeviceGuard dg_(A);
  NSVD_CHECK_ARG(A && B && D && work, "NULL buffer");
  return tc_gemm_selftest(A, B, D, M, N, K, a_kmajor, b_kmajor, work, work_bytes, (cudaStream_t)stream);
}

}  // extern "C"
