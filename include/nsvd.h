/*
 * nsvd.h — C-ABI of the B200-native NestedLoRA training-step library (libnsvd.so).
 *
 * The reference (jongharyu/neural-svd) is pure Python/PyTorch and has NO FFI; the drop-in
 * boundary is its Python API (methods/nestedlora.py:254-267 `NestedLoRA.compute_loss_operator`,
 * :365-378 `NestedLoRAForCDK.compute_loss`).  This header is the C-ABI a maintainer would bind
 * underneath that API (ctypes stub shown in INTEGRATION.md): plain pointers and sizes, no torch
 * types.  Every entry point names the reference code it replaces (paths relative to the
 * reference root).
 *
 * Conventions
 *  - all pointers are DEVICE pointers unless the name ends in `_host`;
 *  - everything is fp32, row-major, contiguous; tensors are owned by the caller (PyTorch's
 *    caching allocator in the shipped host code); the library never allocates device memory:
 *    scratch is passed in (`nsvd_*_bytes` tell how much);
 *  - `stream` is a `cudaStream_t` passed as `void*`; all work is enqueued on it, nothing
 *    synchronises the host;
 *  - return value: 0 on success, otherwise a `cudaError_t` value or NSVD_E_* below;
 *    `nsvd_last_error()` returns a static message for the calling thread;
 *  - there is no CPU fallback: without an sm_100 device every compute call fails.
 */
#ifndef NSVD_H_
#define NSVD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSVD_ABI_VERSION 5

enum {
  NSVD_E_BADARG = 10001,   /* shape / enum / alignment violation */
  NSVD_E_NODEVICE = 10002, /* no sm_100 device */
  NSVD_E_WORKSPACE = 10003 /* workspace too small */
};

/* potentials (examples/operator/pde/schrodinger/potentials.py): hydrogen :5-8 (pot_coef = charge Z), harmonic
 * oscillator :24-27 (pot_coef = k), H2+ ion :11-17 (pot_coef = nuclear charge, pot_coef2 = R, nuclei at (0, +-R)),
 * infinite well :20-21 (V = 0), cosine :30-31 (pot_coef, pot_coef2 = cs[0], cs[1]). */
enum { NSVD_POT_HYDROGEN = 0, NSVD_POT_HARMONIC = 1, NSVD_POT_HYDROGEN_MOL_ION = 2, NSVD_POT_INFINITE_WELL = 3,
       NSVD_POT_COSINE = 4 };

/* importance density w(x) of the sampler (examples/operator/pde/main_pde.py:89-118); sampling_sigma is its scale.
 * NONE = the operator is applied without re-weighting (importance=None, diff_ops.py:10-11). */
enum { NSVD_IMP_GAUSSIAN = 0, NSVD_IMP_LAPLACE = 1, NSVD_IMP_UNIFORM = 2, NSVD_IMP_NONE = 3 };

/* DirichletBoundaryMaskBox modes (examples/operator/pde/boundary.py:16-37) */
enum { NSVD_BOX_NONE = 0, NSVD_BOX_SQRT = 1, NSVD_BOX_EXP = 2 };

/* arithmetic engines for the dense contractions.
 *   FP32_SIMT   : CUDA-core fp32 FMA. Reference-grade accuracy; validation and tiny batches.
 *   F16X3_TC    : tcgen05 tensor cores.  Operator path: every fp32 operand v is stored as two fp16 planes of s*v
 *                 (s = power of two from a rigorous bound of |v|, 22 significant bits) and a product is formed as
 *                 hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM in chains of <= 72 MMAs, each chain
 *                 corrected for the accumulator's truncation bias (1e-6 on loss / gradients: DESIGN.md §3).
 *                 CDK loss: bf16 hi/lo planes (inputs of unknown range), 1e-5.
 *   NSVD_ENGINE_BF16X3_TC is the round-1 name of the same value. */
enum { NSVD_ENGINE_FP32_SIMT = 0, NSVD_ENGINE_F16X3_TC = 1, NSVD_ENGINE_BF16X3_TC = 1 };

/* One problem instance = what the reference spreads over get_problem (pde/problems.py:23-130),
 * get_wavefunctions (pde/__init__.py:19-55) and the Gaussian sampler (pde/main_pde.py:89-100). */
typedef struct nsvd_problem {
  int32_t n_points;        /* B: collocation points held by this rank                          */
  int32_t n_copies;        /* L: number of eigenfunctions = parallel MLP copies (mlp.py:167)    */
  int32_t n_fourier;       /* M_ff: Fourier mapping size; layer-0 fan-in is 2*M_ff (utils.py:102)*/
  int32_t hidden;          /* hidden width of the 3 hidden layers; must be 128                  */
  int32_t potential;       /* NSVD_POT_*                                                        */
  int32_t has_exp_mask;    /* ExponentialMask present (pde/boundary.py:39-53)                   */
  float pot_coef;          /* first potential coefficient (see NSVD_POT_*)                      */
  float scale_kinetic;     /* kappa, problems.py:28 (1.0 for single-particle problems)          */
  float op_scale;          /* OperatorWrapper.scale  (examples/__init__.py:2-9)                 */
  float op_shift;          /* OperatorWrapper.shift                                             */
  float sampling_sigma;    /* scale of the importance density: sigma / b / half-width           */
  float hard_mul_const;    /* WaveFunctions.hard_mul_const (pde/__init__.py:9-16)               */
  int32_t importance;      /* NSVD_IMP_*  (ABI 2)                                               */
  int32_t box_mask;        /* NSVD_BOX_*: Dirichlet box mask multiplying every eigenfunction    */
  float pot_coef2;         /* second potential coefficient                                      */
  float box_lim;           /* half-width `lim` of the Dirichlet box                             */
  float fd_eps;            /* laplacian_eps (ABI 3): <= 0 exact Laplacian (forward-mode streams); > 0 the
                            * finite-difference Laplacian of the shipped scripts, pde/diff_ops.py:25-52      */
  int32_t ndim;            /* spatial dimension D (ABI 4; 0 = 2): x is (n_points, D), Bff (D, n_fourier).  D = 3
                            * (problems.py:62-71: hydrogen, H2+ ion) runs on NSVD_ENGINE_FP32_SIMT only         */
} nsvd_problem_t;

/* Parameters in the reference's own layout (ParallelMLP, mlp.py:181-199):
 *   Bff (2, M_ff); W[0] (L,128,2*M_ff); W[1],W[2] (L,128,128); W[3] (L,1,128);
 *   b[0..2] (L,128,1); b[3] (L,1,1); mask_scales (L) or NULL.                                 */
typedef struct nsvd_params {
  const float* Bff;
  const float* W[4];
  const float* b[4];
  const float* mask_scales;
} nsvd_params_t;

typedef struct nsvd_grads {
  float* dW[4];
  float* db[4];
  float* dmask_scales; /* may be NULL when has_exp_mask == 0 */
} nsvd_grads_t;

int nsvd_abi_version(void);
/* sha256 over the sources, this header and the compiler flags the library was built from (neural_svd_b200/build.py): a
 * host layer that cannot rebuild refuses a library whose hash differs from its source tree.                     */
const char* nsvd_build_hash(void);
/* sizeof(nsvd_problem_t) / sizeof(nsvd_params_t) / sizeof(nsvd_grads_t) as compiled (which = 0, 1, 2): lets a
 * binding written in another language check its struct layout at load time.                            */
size_t nsvd_struct_size(int32_t which);
/* number of kernels this library has launched so far in this process (bench.py: gpu_launches) */
long nsvd_launch_count(void);
/* Optional per-kernel-class timing with CUDA events on the launch stream (bench.py roofline).
 * classes: 0 l0_fwd GEMM, 1 hidden_fwd, 2 hidden_bwd, 3 l0_wgrad GEMM, 4 gram_reduce, 5 loss_dF,
 *          6 prep (features / weight folding), 7 head_bwd.  read() synchronises on the events.  */
void nsvd_profile_enable(int on);
int nsvd_profile_read(double* ms_per_class, long* launches_per_class, int n_classes, int reset);
const char* nsvd_last_error(void);
/* 0 when device `dev` is an sm_100 part this library can run on. */
int nsvd_device_ok(int dev);

/* Points per micro-batch of the tcgen05 engine (rounded up to a multiple of 128; default 65536 or the
 * NSVD_TC_MICROBATCH environment variable).  Sets the size of `work`; call before nsvd_scratch_bytes.  */
void nsvd_set_tc_microbatch(int32_t points);

/* Scratch sizes for nsvd_fwd_streams / nsvd_mlp_bwd.  `saved_bytes`: value-stream activations
 * kept from forward to backward (whole batch).  `work_bytes`: micro-batch scratch, reusable. */
int nsvd_scratch_bytes(const nsvd_problem_t* pb, int engine, size_t* saved_bytes, size_t* work_bytes);

/* K1  fwd_streams.  Replaces, in one pass and in forward mode (no autograd double backward):
 *   GaussianFourierFeatureTransform.forward  examples/utils.py:126-143
 *   ParallelMLP.forward                      examples/models/mlp.py:204-221
 *   WaveFunctions.forward / ExponentialMask  pde/__init__.py:15-16, pde/boundary.py:46-53
 *   VectorizedLaplacian (exact, importance)  pde/diff_ops.py:9-23,54-93
 *   NegativeHamiltonian.__call__             pde/schrodinger/__init__.py:16-22
 *   OperatorWrapper.__call__                 examples/__init__.py:7-9
 * in : x (B,2);  out: F (B,L), TF (B,L); `saved` is filled for nsvd_mlp_bwd.                    */
int nsvd_fwd_streams(const nsvd_problem_t* pb, const nsvd_params_t* pr, int engine, const float* x,
                     float* F, float* TF, void* saved, size_t saved_bytes, void* work,
                     size_t work_bytes, void* stream);

/* K2  gram_reduce.  Replaces compute_lambda / compute_loss_metric / the operator term
 * (methods/nestedlora.py:10-11,57-64,92) up to normalisation: writes the UN-normalised sums
 *   terms = [ F1^T F1 (L*L) | F2^T F2 (L*L) | sum_b sum_l v_l F_bl TF_bl (1) ]
 * where F1 = rows [0,b1), F2 = rows [b1,B).  This is the buffer that is all-reduced across
 * ranks.  `partials` must hold nsvd_gram_partials_bytes(B, L) bytes.  Deterministic.           */
size_t nsvd_gram_partials_bytes(int32_t n_points, int32_t n_copies);
int nsvd_gram_reduce(const float* F, const float* TF, const float* vector_mask, int32_t n_points,
                     int32_t n_copies, int32_t b1, float* terms, void* partials, void* stream);

/* Full cross Grams for evaluation (methods/spectrum.py:62-75): with phi = nan_to_num(w F),
 * Tphi = nan_to_num(w TF) and Tphi rows zeroed where x is the origin (x may be NULL):
 *   cov += phi^T phi, quad += phi^T Tphi.  `roww` = per-row sqrt weight ratio (may be NULL).
 * Accumulates into cov, quad (L*L each, caller zero-initialises).                              */
int nsvd_cross_gram(const float* F, const float* TF, const float* roww, const float* x, int32_t n_points,
                    int32_t n_copies, float* cov, float* quad, void* partials, void* stream);

/* Loss value from (all-reduced) terms: NestedLoRALossFunctionEVD.forward, nestedlora.py:70-94.
 * Bg, B1g, B2g are the GLOBAL row counts.  Writes loss[0] and coef (2*L*L):
 *   coef[0:L*L]   = (2/B1g) * M * Lambda2   (applied to rows of F1)
 *   coef[L*L:]    = (2/B2g) * M * Lambda1   (applied to rows of F2)
 * Data-parallel form without a host copy of the counts: pass Bg <= 0 and let the counts travel in the all-reduced
 * buffer itself, terms[2 L^2 + 1 .. +4] = sum over ranks of [n mod 2^16, n / 2^16, b1 mod 2^16, b1 / 2^16]; `coef`
 * must then hold 2 L^2 + 1 floats (coef[2 L^2] = 4 / B_global, read by nsvd_loss_dF called with Bg <= 0).          */
int nsvd_loss_finalize(const float* terms, const float* matrix_mask, int32_t n_copies, int64_t Bg,
                       int64_t B1g, int64_t B2g, float* loss, float* coef, void* stream);

/* K3  loss_dF.  The reference's hand-written backward (nestedlora.py:98-111), summed over the
 * views f, f1, f2:  dF = gscale * ( -(4/Bg) v (.) TF + F_half . coef_half ).
 * TF == NULL drops the operator term, coef == NULL drops the metric term (used by the
 * stand-alone NestedLoRALossFunctionEVD, whose f, f1, f2 need not alias).                      */
int nsvd_loss_dF(const float* F, const float* TF, const float* vector_mask, const float* coef,
                 const float* grad_scale /*device scalar or NULL (=1)*/, int32_t n_points,
                 int32_t n_copies, int32_t b1, int64_t Bg, float* dF, void* stream);

/* K4  mlp_bwd.  Replaces the autograd backward through the central model evaluation
 * (value stream only).  Gradients are WRITTEN (not accumulated) into `gr`.   */
int nsvd_mlp_bwd(const nsvd_problem_t* pb, const nsvd_params_t* pr, int engine, const float* x,
                 const float* dF, const void* saved, size_t saved_bytes, nsvd_grads_t* gr,
                 void* work, size_t work_bytes, void* stream);

/* K5/K6  CDK loss (methods/nestedlora.py:270-332).  f, g: (B, L) WITHOUT the constant column;
 * vector_mask (Lp), matrix_mask (Lp,Lp) with Lp = L + first_const.
 *   fwd: terms = [Fp^T Fp | Gp^T Gp | sum v f g] un-normalised (all-reducible), then
 *   nsvd_cdk_finalize -> losses[3] = {loss, loss_operator, loss_metric}, coef (2*Lp*Lp);
 *   bwd: grad_f, grad_g (B, L).  rs_joint (B) = diag(Fp Gp^T) is produced by fwd when non-NULL;
 *   nsvd_cdk_offdiag writes off_diagonal(Fp Gp^T) (B*B-B) on request (methods/utils.py:16-22).
 *   With NSVD_ENGINE_BF16X3_TC the Grams (MN-major operands, K = rows), the backward GEMMs and
 *   Fp Gp^T run on the tcgen05 GEMM block; the row dots stay exact fp32.
 *   ABI 5: nsvd_cdk_finalize runs on many blocks and takes `scratch`, NSVD_CDK_FINALIZE_SCRATCH bytes of
 *   device memory (8-byte aligned, contents irrelevant; the LAST that many bytes of the cdk work buffer are
 *   reserved for it, so no extra allocation is needed).  `planes_ready` != 0 tells bwd / offdiag that `work`
 *   still holds what nsvd_cdk_fwd built from these same f, g (the operand planes are then not rebuilt).  */
#define NSVD_CDK_FINALIZE_SCRATCH 1024
size_t nsvd_cdk_work_bytes(int32_t n_rows, int32_t n_feat, int32_t first_const, int engine);
int nsvd_cdk_fwd(const float* f, const float* g, const float* vector_mask, int32_t n_rows,
                 int32_t n_feat, int32_t first_const, int engine, float* terms, float* rs_joint,
                 void* work, size_t work_bytes, void* stream);
int nsvd_cdk_finalize(const float* terms, const float* matrix_mask, int32_t Lp, int64_t Bg,
                      float* losses, float* coef, void* scratch, void* stream);
int nsvd_cdk_bwd(const float* f, const float* g, const float* vector_mask, const float* coef,
                 const float* grad_scale, int32_t n_rows, int32_t n_feat, int32_t first_const,
                 int64_t Bg, int engine, float* grad_f, float* grad_g, void* work, size_t work_bytes,
                 int32_t planes_ready, void* stream);
int nsvd_cdk_offdiag(const float* f, const float* g, int32_t n_rows, int32_t n_feat,
                     int32_t first_const, int engine, float* rs_indep, void* work, size_t work_bytes,
                     int32_t planes_ready, void* stream);

/* ---- "next" rows of the scope table ------------------------------------------
 * Fused optimizer step for up to 16 tensors in one launch: RMSprop with momentum 0 / weight decay 0
 * (examples/utils.py:48-57: lr, alpha = rmsprop_decay, eps = 1e-10) followed by the torch_ema shadow
 * update used by the loop (examples/operator/__init__.py:36,73):
 *   sq = alpha sq + (1-alpha) g^2 ; p -= lr g / (sqrt(sq) + eps) ; ema -= w (ema - p), w = 1 - decay_t.
 * The pointer tables are HOST arrays of device pointers; `ema` may be NULL.  The cosine learning-rate
 * schedule (operator/__init__.py:35,71-72) is evaluated by the caller and passed as `lr`.        */
int nsvd_rmsprop_ema_step(int32_t n_tensors, float* const* params, const float* const* grads,
                          float* const* square_avg, float* const* ema, const int64_t* sizes, float lr,
                          float alpha, float eps, float ema_one_minus_decay, void* stream);
/* On-device Gaussian sampler x (n_points, 2) = sigma * N(0, I) (main_pde.py:92-93 draws on the CPU;
 * that stays the RNG-parity mode).  Counter-based: reproducible for (seed, offset).               */
int nsvd_sample_gaussian(float* x, int64_t n_points, float sigma, uint64_t seed, uint64_t offset,
                         void* stream);
/* The other samplers of main_pde.py:101-118 on the device, same counter scheme: `importance` = NSVD_IMP_LAPLACE
 * (x_i ~ Laplace(0, scale), inverse CDF) or NSVD_IMP_UNIFORM (x_i ~ U[-scale, scale)); NSVD_IMP_GAUSSIAN forwards
 * to nsvd_sample_gaussian.                                                                          */
int nsvd_sample_points(float* x, int64_t n_points, int32_t importance, float scale, uint64_t seed, uint64_t offset,
                       void* stream);

/* Dense layers of the CDK encoder on the tcgen05 GEMM block: replaces `nn.Linear` (+ LeakyReLU / ReLU)
 * inside `get_mlp` (/root/reference/examples/models/mlp.py:129-164) as `HeteroNetwork` uses it
 * (/root/reference/examples/models/siam.py:132-165; main_sketchy.py:107-115: two 512 -> 8192 -> 512 towers).
 *   forward : y (rows, out) = act(x (rows, in) . W (out, in)^T + bias),  act = 0 none | 1 leaky ReLU(slope), slope 0 = ReLU
 *   backward: dz = dy * act'(y);  dx = dz . W;  dW = dz^T . x;  db = column sums of dz   (dx / dW / db may be NULL)
 * fp32 tensors, row-major; operands go through bf16 hi/lo planes (3 products, fp32 accumulation in bounded chains).
 * in / out features must be multiples of 8.  `work` >= nsvd_linear_work_bytes(rows, in, out) bytes.              */
size_t nsvd_linear_work_bytes(int32_t rows, int32_t in_features, int32_t out_features);
int nsvd_linear_fwd(const float* x, const float* W, const float* bias, float* y, int32_t rows, int32_t in_features,
                    int32_t out_features, int32_t act, float slope, void* work, size_t work_bytes, void* stream);
int nsvd_linear_bwd(const float* x, const float* W, const float* y, const float* dy, int32_t rows, int32_t in_features,
                    int32_t out_features, int32_t act, float slope, float* dx, float* dW, float* db, void* work,
                    size_t work_bytes, void* stream);

/* Self-test hooks for the tcgen05 building block (tests/test_gpu_tc_gemm.py):
 *   D (M,N) fp32 = A . B^T with bf16x3 splitting; A (M,K), B (N,K) fp32 when *_kmajor = 1,
 *   A (K,M) / B (K,N) when 0 (MN-major operands, as the weight-gradient GEMMs use them).       */
int nsvd_tc_gemm_selftest(const float* A, const float* B, float* D, int32_t M, int32_t N, int32_t K,
                          int32_t a_kmajor, int32_t b_kmajor, void* work, size_t work_bytes,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NSVD_H_ */
