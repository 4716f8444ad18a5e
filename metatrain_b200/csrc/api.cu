// C-ABI glue: error reporting and the GEMM entry point (argument validation + dispatch).
#include <stdarg.h>

#include "common.cuh"
#include "kernels.cuh"

namespace petb200 {

static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

}  // namespace petb200

using namespace petb200;

extern "C" PETB200_API const char* petb200_last_error(void) { return g_last_error; }
extern "C" PETB200_API int petb200_version(void) { return 1; }

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" PETB200_API int petb200_gemm(const float* A, int64_t lda, const float* W, int64_t ldw, float* C,
                            int64_t ldc, int64_t M, int N, int K, const float* bias,
                            const float* row_scale, const float* residual, int64_t ldr,
                            const float* aux_in, float* aux_out, int64_t ld_aux, int epilogue,
                            int accumulate, int precision, cudaStream_t stream) {
  PETB200_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm: bad shape M=%lld N=%d K=%d", (long long)M, N, K);
  PETB200_REQUIRE(N % 128 == 0, "gemm: N=%d must be a multiple of 128", N);
  PETB200_REQUIRE(K % 16 == 0, "gemm: K=%d must be a multiple of 16", K);
  if (M == 0) return PETB200_OK;  // empty operands carry null pointers
  PETB200_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && ldc % 4 == 0, "gemm: leading dims must be x4");
  PETB200_REQUIRE(aligned16(A) && aligned16(W) && aligned16(C), "gemm: pointers must be 16 B aligned");
  PETB200_REQUIRE(!residual || (ldr % 4 == 0 && aligned16(residual)), "gemm: residual alignment");
  PETB200_REQUIRE((!aux_in && !aux_out) || ld_aux % 4 == 0, "gemm: ld_aux must be x4");
  PETB200_REQUIRE(epilogue != PETB200_EPI_MUL_DSILU || aux_in, "gemm: MUL_DSILU needs aux_in");
  PETB200_REQUIRE(epilogue != PETB200_EPI_SWIGLU_BWD || aux_in, "gemm: SWIGLU_BWD needs aux_in");
  PETB200_REQUIRE(!(accumulate && (epilogue == PETB200_EPI_SWIGLU || epilogue == PETB200_EPI_SWIGLU_BWD)),
                  "gemm: accumulate is not available with SwiGLU epilogues");
  PETB200_REQUIRE(!(residual && (epilogue == PETB200_EPI_SWIGLU || epilogue == PETB200_EPI_SWIGLU_BWD)),
                  "gemm: residual is not available with SwiGLU epilogues");
  GemmArgs g;
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.C = C; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K;
  g.bias = bias; g.row_scale = row_scale; g.residual = residual; g.ldr = ldr;
  g.aux_in = aux_in; g.aux_out = aux_out; g.ld_aux = ld_aux;
  g.epilogue = epilogue; g.accumulate = accumulate;
  if (precision == PETB200_PREC_FP32) return launch_gemm_simt(g, stream);
  if (precision == PETB200_PREC_BF16X3 || precision == PETB200_PREC_BF16) {
    if (!gemm_tc_supports(g)) {
      set_error("gemm: shape M=%lld N=%d K=%d epilogue=%d is not covered by the tcgen05 kernel",
                (long long)M, N, K, epilogue);
      return PETB200_ERR_UNSUPPORTED;
    }
    return launch_gemm_tc(g, precision, stream);
  }
  set_error("gemm: unknown precision %d", precision);
  return PETB200_ERR_INVALID_ARGUMENT;
}

extern "C" PETB200_API int petb200_compress_gemm(const float* messages, int64_t ld_m, const float* w1m_split,
                                     const float* bias, const float* geo_w, const float* nbr_table,
                                     const int32_t* z_neighbor, const float* edge_vec,
                                     const float* edge_dist, int64_t n_edges, int d, float* pre,
                                     float* out, int precision, cudaStream_t stream) {
  PETB200_REQUIRE(d == 128, "compress_gemm: only d_pet = 128 is built (got %d)", d);
  PETB200_REQUIRE(precision == PETB200_PREC_BF16X3 || precision == PETB200_PREC_BF16,
                  "compress_gemm: tensor-core precisions only (the fp32 path uses compress_input + gemm)");
  if (n_edges == 0) return PETB200_OK;
  PETB200_REQUIRE(messages && w1m_split && geo_w && edge_vec && edge_dist && pre && out,
                  "compress_gemm: null argument");
  PETB200_REQUIRE(!nbr_table || z_neighbor, "compress_gemm: a table needs its row index");
  PETB200_REQUIRE(ld_m % 4 == 0, "compress_gemm: ld_m must be a multiple of 4");
  GemmArgs g;
  g.A = messages; g.lda = ld_m; g.W = w1m_split; g.ldw = d; g.C = out; g.ldc = d;
  g.M = n_edges; g.N = d; g.K = d;
  g.bias = bias; g.row_scale = nullptr; g.residual = nullptr; g.ldr = 0;
  g.aux_in = nullptr; g.aux_out = pre; g.ld_aux = d;
  g.epilogue = PETB200_EPI_SILU_GEO; g.accumulate = 0;
  g.geo_vec = edge_vec; g.geo_dist = edge_dist; g.geo_w = geo_w;
  g.row_table = nbr_table; g.row_index = z_neighbor;
  return launch_gemm_tc(g, precision, stream);
}
