// Internal launcher declarations (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace petb200 {

struct GemmArgs {
  const float* A;
  int64_t lda;
  const float* W;
  int64_t ldw;
  float* C;
  int64_t ldc;
  int64_t M;
  int N;  // number of weight rows (= GEMM N)
  int K;
  const float* bias;       // [N] or null
  const float* row_scale;  // [M] or null
  const float* residual;   // [M, N] or null
  int64_t ldr;
  const float* aux_in;  // MUL_DSILU: pre-activation [M, N]; SWIGLU_BWD: [u | g] [M, 2N]
  float* aux_out;       // SILU: pre-activation [M, N]; SWIGLU: [u | g] [M, N]
  int64_t ld_aux;
  int epilogue;
  int accumulate;
};

int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream);
int launch_gemm_tc(const GemmArgs& g, int precision, cudaStream_t stream);
bool gemm_tc_supports(const GemmArgs& g);

}  // namespace petb200
