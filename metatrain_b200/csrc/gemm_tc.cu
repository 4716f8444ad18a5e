// tcgen05 tensor-core GEMM (placeholder until the UMMA kernel lands).
#include "common.cuh"
#include "kernels.cuh"

namespace petb200 {
bool gemm_tc_supports(const GemmArgs&) { return false; }
int launch_gemm_tc(const GemmArgs&, int, cudaStream_t) {
  set_error("gemm: tcgen05 path not built");
  return PETB200_ERR_UNSUPPORTED;
}
}  // namespace petb200
