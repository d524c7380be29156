// tcgen05 / TMEM / TMA 3xTF32 GEMM backend (placeholder until the tensor-core
// path lands): reports "not eligible" so immtsf_gemm routes to the FFMA kernel.
#include "common.cuh"

int immtsf_gemm_tc_eligible(int, int, int, int, int, const float*, int, const float*, int, const float*, int) { return 0; }
int immtsf_gemm_tc(int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int,
                   const float*, const int32_t*, int, cudaStream_t) {
  immtsf_set_error("gemm_tc: not built");
  return IMMTSF_ERR_UNSUPPORTED;
}
