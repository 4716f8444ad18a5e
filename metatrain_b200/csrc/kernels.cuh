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
  int64_t residual_rows = -1;  // >= 0: the residual applies to rows [0, residual_rows) only (tensor-core path)
  const float* aux_in;  // MUL_DSILU: pre-activation [M, N]; SWIGLU_BWD: [u | g] [M, 2N]
  float* aux_out;       // SILU: pre-activation [M, N]; SWIGLU: [u | g] [M, N]
  int64_t ld_aux;
  int epilogue;
  int accumulate;
  // PETB200_EPI_SILU_GEO: per-row additive term of the pre-activation,
  //   geo_w[c] . (edge_vec[m], edge_dist[m]) + row_table[row_index[m]][c]
  const float* geo_vec = nullptr;     // [M, 3]
  const float* geo_dist = nullptr;    // [M]
  const float* geo_w = nullptr;       // [N, 4]
  const float* row_table = nullptr;   // [S, N] or null
  const int32_t* row_index = nullptr; // [M]
};

int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream);
int launch_gemm_tc(const GemmArgs& g, int precision, cudaStream_t stream);
bool gemm_tc_supports(const GemmArgs& g);

// tensor-core attention (attention_tc.cu): mma.sync bf16 hi/lo split, T <= 64 tokens per atom
bool attention_tc_supports(int num_heads, int head_dim, int max_row);
int launch_attention_fwd_tc(const float* qkv, const int32_t* row_ptr, const float* fc,
                            int64_t n_atoms, int64_t n_edges, float scale, int max_row, float* out,
                            float* lse, cudaStream_t stream);
int launch_attention_bwd_tc(const float* qkv, const float* out, const float* lse, const float* d_out,
                            const int32_t* row_ptr, const float* fc, int64_t n_atoms,
                            int64_t n_edges, float scale, int max_row, float* d_qkv, float* d_fc,
                            float* scratch /* [num_heads, E + N] */, cudaStream_t stream);

}  // namespace petb200
