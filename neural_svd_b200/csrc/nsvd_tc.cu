// nsvd_tc.cu — tcgen05 / TMEM / TMA engine (bf16x3).  Placeholder until the kernels land.
#include "nsvd_simt.cuh"
namespace nsvd {
void tc_scratch_bytes(const nsvd_problem_t& pb, size_t* saved, size_t* work) { *saved = 256; *work = 256; (void)pb; }
int tc_forward(const nsvd_problem_t&, const nsvd_params_t&, const float*, float*, float*, void*, void*, size_t, cudaStream_t) {
  set_error("tcgen05 engine not built yet"); return NSVD_E_BADARG; }
int tc_backward(const nsvd_problem_t&, const nsvd_params_t&, const float*, const float*, const void*, nsvd_grads_t&, void*, size_t, cudaStream_t) {
  set_error("tcgen05 engine not built yet"); return NSVD_E_BADARG; }
int tc_gemm_selftest(const float*, const float*, float*, int, int, int, int, int, void*, size_t, cudaStream_t) {
  set_error("tcgen05 engine not built yet"); return NSVD_E_BADARG; }
}
