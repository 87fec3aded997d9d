// nsvd_simt.cuh — declarations of the fp32 engine and the engine-independent kernels.
#pragma once
#include "nsvd_common.cuh"

namespace nsvd {

// C[b](m,n) (+)= alpha * sum_k A[b](m,k) B[b](k,n); element strides: A(m,k) = A[m*a_rs + k*a_cs],
// B(k,n) = B[k*b_rs + n*b_cs], C(m,n) = C[m*c_rs + n]; *_bs = batch strides.
struct SGemm {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  long a_rs, a_cs, a_bs, b_rs, b_cs, b_bs, c_rs, c_bs;
  float alpha;
  int accumulate;
};
int sgemm_strided(const SGemm& g, int batch, cudaStream_t st);
int colsum(const float* X, float* out, int M, int N, int batch, long x_bs, int accumulate, cudaStream_t st);

void simt_scratch_bytes(const nsvd_problem_t& pb, size_t* saved, size_t* work);
int simt_forward(const nsvd_problem_t& pb, const nsvd_params_t& pr, const float* x, float* F, float* TF,
                 void* saved, void* work, cudaStream_t st);
int simt_backward(const nsvd_problem_t& pb, const nsvd_params_t& pr, const float* x, const float* dF,
                  const void* saved, nsvd_grads_t& gr, void* work, cudaStream_t st);

size_t gram_partials_bytes(int B, int L);
int gram_reduce(const float* F, const float* TF, const float* vmask, int B, int L, int b1, float* terms,
                void* partials, cudaStream_t st);
int cross_gram(const float* F, const float* TF, const float* roww, const float* xrow, int B, int L, float* cov,
               float* quad, void* partials, cudaStream_t st);
int loss_finalize(const float* terms, const float* Mm, int L, long Bg, long B1g, long B2g, float* loss,
                  float* coef, cudaStream_t st);
int loss_dF(const float* F, const float* TF, const float* vmask, const float* coef, const float* gscale,
            int B, int L, int b1, long Bg, float* dF, cudaStream_t st);

size_t cdk_work_bytes(int B, int L, int fc);
int cdk_fwd(const float* f, const float* g, const float* v, int B, int L, int fc, float* terms,
            float* rs_joint, void* work, cudaStream_t st);
int cdk_finalize(const float* terms, const float* Mm, int Lp, long Bg, float* losses, float* coef, double* scratch,
                 cudaStream_t st);
int cdk_bwd(const float* f, const float* g, const float* v, const float* coef, const float* gscale, int B,
            int L, int fc, long Bg, float* grad_f, float* grad_g, cudaStream_t st);
int cdk_offdiag(const float* f, const float* g, int B, int L, int fc, float* out, cudaStream_t st);

struct OptTensors {
  int n;
  float* p[16];
  const float* g[16];
  float* sq[16];
  float* ema[16];
  long size[16];
};
int rmsprop_ema_step(const OptTensors& t, float lr, float alpha, float eps, float ema_w, cudaStream_t st);
int sample_gaussian2(float* x, long n, float sigma, uint64_t seed, uint64_t offset, cudaStream_t st);
int sample_other2(float* x, long n, int laplace, float scale, uint64_t seed, uint64_t offset, cudaStream_t st);

// tcgen05 engine (nsvd_tc.cu)
void tc_set_micro_batch(int points);
void tc_scratch_bytes(const nsvd_problem_t& pb, size_t* saved, size_t* work);
int tc_forward(const nsvd_problem_t& pb, const nsvd_params_t& pr, const float* x, float* F, float* TF,
               void* saved, void* work, size_t work_bytes, cudaStream_t st);
int tc_backward(const nsvd_problem_t& pb, const nsvd_params_t& pr, const float* x, const float* dF,
                const void* saved, nsvd_grads_t& gr, void* work, size_t work_bytes, cudaStream_t st);
size_t tc_cdk_work_bytes(int B, int L, int fc);
int tc_cdk_fwd(const float* f, const float* g, const float* v, int B, int L, int fc, float* terms, float* rs_joint,
               void* work, cudaStream_t st);
int tc_cdk_bwd(const float* f, const float* g, const float* v, const float* coef, const float* gscale, int B, int L,
               int fc, long Bg, float* grad_f, float* grad_g, void* work, int planes_ready, cudaStream_t st);
int tc_cdk_offdiag(const float* f, const float* g, int B, int L, int fc, float* out, void* work, int planes_ready,
                   cudaStream_t st);
size_t tc_linear_work_bytes(int rows, int in_f, int out_f);
int tc_linear_fwd(const float* x, const float* W, const float* bias, float* y, int rows, int in_f, int out_f, int act,
                  float slope, void* work, size_t work_bytes, cudaStream_t st);
int tc_linear_bwd(const float* x, const float* W, const float* y, const float* dy, int rows, int in_f, int out_f, int act,
                  float slope, float* dx, float* dW, float* db, void* work, size_t work_bytes, cudaStream_t st);
int tc_gemm_selftest(const float* A, const float* B, float* D, int M, int N, int K, int a_kmajor,
                     int b_kmajor, void* work, size_t work_bytes, cudaStream_t st);

}  // namespace nsvd
