// Stage-level schedule of the default PET layer, issued from C++ (see include/petb200.h,
// "stage-level schedule").  This file contains no kernels: it sequences the entry points of the
// other translation units on the caller's stream, with buffers carved out of the caller's
// `saved` / `scratch` allocations.  It is the C++ mirror of metatrain_b200/engine.py
// (_gnn_forward / _gnn_backward, default path), which itself restates
// CartesianTransformer.forward (src/metatrain/pet/modules/transformer.py:463-562) and
// TransformerLayer._forward_pre_ln_impl (:203-234) on the CSR token layout.
#include "common.cuh"
#include "kernels.cuh"

namespace petb200 {
namespace {

constexpr size_t kAlign = 256;
inline size_t round_up(size_t n) { return (n + kAlign - 1) / kAlign * kAlign; }

// bump allocator over a caller-owned buffer (dry run with base == nullptr measures)
struct Arena {
  char* base;
  size_t used = 0;
  explicit Arena(void* p) : base(static_cast<char*>(p)) {}
  float* f32(int64_t rows, int64_t cols = 1) {
    const size_t off = used;
    used += round_up((size_t)(rows > 0 ? rows : 0) * (size_t)cols * sizeof(float));
    return base ? reinterpret_cast<float*>(base + off) : nullptr;
  }
};

// what the forward keeps per attention layer for the backward
struct SavedLayer {
  float *x, *rstd1, *qkv, *o, *lse, *tp, *h1, *rstd3, *ugc;
};
struct Saved {
  float* c1;
  SavedLayer tl[16];
};

Saved carve_saved(Arena& a, const petb200_gnn_weights& w, const petb200_dims& g) {
  Saved s{};
  const int64_t E = g.n_edges, N = g.n_atoms, T = E + N;
  s.c1 = a.f32((E + 127) / 128 * 128, g.d);   // (whole 128-edge tiles: private layout of the fused token builder)
  for (int k = 0; k < w.n_tl; ++k) {
    SavedLayer& L = s.tl[k];
    L.x = a.f32(T, g.d);
    L.rstd1 = a.f32(T);
    L.qkv = a.f32(T, 3 * g.d);
    L.o = a.f32(T, g.d);
    L.lse = a.f32(T, g.num_heads);
    L.tp = a.f32(T, g.d);   // rows [0, E): t' (edge tokens after attention); rows [E, T): y_c (centre rows)
    L.h1 = a.f32(N, g.d_node);
    L.rstd3 = a.f32(N);
    L.ugc = a.f32(N, 4 * g.d_node);
  }
  return s;
}

#define CHECK(call)            \
  do {                         \
    if (int rc_ = (call)) return rc_; \
  } while (0)

int gemm(const float* A, int64_t lda, const petb200_mat& W, float* C, int64_t ldc, int64_t M, int N, int K,
         const float* bias, const float* row_scale, const float* residual, int64_t ldr, const float* aux_in,
         float* aux_out, int64_t ld_aux, int epilogue, int accumulate, int precision, cudaStream_t stream) {
  return petb200_gemm(A, lda, W.w, W.ld, C, ldc, M, N, K, bias, row_scale, residual, ldr, aux_in, aux_out, ld_aux,
                      epilogue, accumulate, precision, stream);
}

// One contraction over ALL token rows [edges | atoms] whose residual applies to the edge rows only:
// the edge-row and the centre-row GEMM of the per-op schedule share their weights, so they are one launch
int gemm_tokens(const float* A, int64_t lda, const petb200_mat& W, float* C, int64_t ldc, int64_t rows, int64_t edge_rows,
                int N, int K, const float* bias, const float* row_scale, const float* residual, int64_t ldr,
                const float* aux_in, int64_t ld_aux, int epilogue, int precision, cudaStream_t stream) {
  if (rows == 0) return PETB200_OK;
  GemmArgs g;
  g.A = A; g.lda = lda; g.W = W.w; g.ldw = W.ld; g.C = C; g.ldc = ldc; g.M = rows; g.N = N; g.K = K;
  g.bias = bias; g.row_scale = row_scale; g.residual = residual; g.ldr = ldr; g.residual_rows = edge_rows;
  g.aux_in = aux_in; g.aux_out = nullptr; g.ld_aux = ld_aux; g.epilogue = epilogue; g.accumulate = 0;
  if (!gemm_tc_supports(g)) {
    set_error("gnn schedule: contraction shape outside the tensor-core kernel (N %% 128, K %% 64)");
    return PETB200_ERR_UNSUPPORTED;
  }
  return launch_gemm_tc(g, precision, stream);
}

int check(const petb200_gnn_weights* w, const petb200_dims* g, const char* what) {
  PETB200_REQUIRE(w && g && w->tl, "%s: null weights / dims", what);
  PETB200_REQUIRE(g->precision != PETB200_PREC_FP32 && g->d == 128 && g->d_node % 4 == 0 && w->n_tl >= 1 && w->n_tl <= 16,
                  "%s: built for the tensor-core precisions, d_pet = 128 and 1..16 attention layers", what);
  return PETB200_OK;
}

}  // namespace
}  // namespace petb200

using namespace petb200;

extern "C" PETB200_API size_t petb200_gnn_saved_bytes(const petb200_gnn_weights* w, const petb200_dims* g) {
  if (!w || !g) return 0;
  Arena a(nullptr);
  carve_saved(a, *w, *g);
  return a.used + kAlign;
}

extern "C" PETB200_API size_t petb200_gnn_scratch_bytes(const petb200_gnn_weights* w, const petb200_dims* g) {
  if (!w || !g) return 0;
  const int64_t E = g->n_edges, N = g->n_atoms, T = E + N;
  Arena f(nullptr), b(nullptr);
  // forward: a1, (spare), sc, two node-feature buffers
  f.f32(E, g->d); f.f32(N, g->d); f.f32(N, 2 * g->d_node); f.f32(N, g->d_node); f.f32(N, g->d_node);
  // backward: d_tp | d_yc, two (d_t | d_c) buffers, d_ugc, d_xhc, d_h1, d_o, d_qkv, dsum, two d_h buffers, d_c1
  b.f32(T, g->d); b.f32(T, g->d); b.f32(T, g->d); b.f32(N, 4 * g->d_node); b.f32(N, g->d_node);
  b.f32(N, g->d_node); b.f32(T, g->d); b.f32(T, 3 * g->d); b.f32(T, g->num_heads);
  b.f32(N, g->d_node); b.f32(N, g->d_node); b.f32(E, g->d);
  return (f.used > b.used ? f.used : b.used) + kAlign;
}

extern "C" PETB200_API int petb200_gnn_fwd(const petb200_gnn_weights* w, const petb200_dims* g, const int32_t* row_ptr,
                                           const int32_t* z_neighbors, const float* edge_vec, const float* edge_dist,
                                           const float* fc, const float* h_in, const float* m_in, int64_t ld_m,
                                           float* x_out, float* h_out, void* saved, size_t saved_bytes, void* scratch,
                                           size_t scratch_bytes, cudaStream_t stream) {
  CHECK(check(w, g, "gnn_fwd"));
  if (saved_bytes < petb200_gnn_saved_bytes(w, g) - kAlign || scratch_bytes < petb200_gnn_scratch_bytes(w, g) - kAlign) {
    set_error("gnn_fwd: saved / scratch buffer too small");
    return PETB200_ERR_WORKSPACE;
  }
  const int64_t E = g->n_edges, N = g->n_atoms, T = E + N;
  const int d = g->d, dn = g->d_node, nh = g->num_heads, prec = g->precision;
  Arena sa(saved), sc(scratch);
  Saved S = carve_saved(sa, *w, *g);
  float* a1 = sc.f32(E, d);
  sc.f32(N, d);
  float* swi = sc.f32(N, 2 * dn);
  float* hbuf[2] = {sc.f32(N, dn), sc.f32(N, dn)};

  // token builder: t = W_2 silu(W_1 cat[geo, nbr, m] + b_1) + b_2, concatenation folded (compress_gemm)
  float* x = w->n_tl > 0 ? S.tl[0].x : x_out;
  if (w->compress_image_fwd != nullptr) {
    // both Linears and the SiLU between them in one kernel (chain_fused.cu)
    CHECK(petb200_compress_fwd(m_in, ld_m, w->compress_image_fwd, w->b_fold, w->geo_fold, w->nbr_fold, z_neighbors,
                               edge_vec, edge_dist, w->b2, E, d, S.c1, x, d, stream));
  } else {
    CHECK(petb200_compress_gemm(m_in, ld_m, w->w1m.w, w->b_fold, w->geo_fold, w->nbr_fold, z_neighbors, edge_vec,
                                edge_dist, E, d, S.c1, a1, prec, stream));
    CHECK(gemm(a1, d, w->w2, x, d, E, d, d, w->b2, nullptr, nullptr, 0, nullptr, nullptr, 0, PETB200_EPI_NONE, 0, prec,
               stream));
  }
  const float* h = h_in;
  for (int k = 0; k < w->n_tl; ++k) {
    const petb200_tl_weights& t = w->tl[k];
    SavedLayer& L = S.tl[k];
    float* xn = k + 1 < w->n_tl ? S.tl[k + 1].x : x_out;     // token matrix after this layer
    float* h2 = k + 1 < w->n_tl ? hbuf[k & 1] : h_out;
    // centre token = contraction of the node features
    CHECK(gemm(h, dn, t.w_con, L.x + E * d, d, N, d, dn, t.b_con, nullptr, nullptr, 0, nullptr, nullptr, 0,
               PETB200_EPI_NONE, 0, prec, stream));
    // RMSNorm + QKV projection, attention
    CHECK(petb200_norm_linear(L.x, d, t.qkv_image, t.b_qkv, T, d, 3 * d, L.qkv, 3 * d, L.rstd1, stream));
    CHECK(petb200_attention_fwd(L.qkv, row_ptr, fc, N, E, nh, d / nh, g->scale, g->max_row, prec, L.o, L.lse, stream));
    // output projection of all token rows in one launch: edge rows with the residual, centre rows without
    float* yc = L.tp + E * d;
    CHECK(gemm_tokens(L.o, d, t.w_o, L.tp, d, T, E, d, d, t.b_o, nullptr, L.x, d, nullptr, 0, PETB200_EPI_NONE, prec,
                      stream));
    // node update: h1 = h + W_exp y_c ; h2 = h1 + W_out swiglu(W_in rms(h1))
    CHECK(gemm(yc, d, t.w_exp, L.h1, dn, N, dn, d, t.b_exp, nullptr, h, dn, nullptr, nullptr, 0, PETB200_EPI_NONE, 0,
               prec, stream));
    CHECK(petb200_rms_rstd(L.h1, N, dn, L.rstd3, stream));
    CHECK(gemm(L.h1, dn, t.wc_in, swi, 2 * dn, N, 4 * dn, dn, t.bc_in, L.rstd3, nullptr, 0, nullptr, L.ugc, 4 * dn,
               PETB200_EPI_SWIGLU, 0, prec, stream));
    CHECK(gemm(swi, 2 * dn, t.wc_out, h2, dn, N, dn, 2 * dn, t.bc_out, nullptr, L.h1, dn, nullptr, nullptr, 0,
               PETB200_EPI_NONE, 0, prec, stream));
    // edge feed-forward (one fused kernel; the backward recomputes its hidden activations)
    CHECK(petb200_mlp_fwd(L.tp, d, t.mlp_image_fwd, t.b_in, t.b_out, E, d, t.d_ff, xn, d, stream));
    h = h2;
  }
  return PETB200_OK;
}

extern "C" PETB200_API int petb200_gnn_bwd(const petb200_gnn_weights* w, const petb200_dims* g, const int32_t* row_ptr,
                                           const float* fc, const void* saved, const float* d_h_out, const float* d_t_out,
                                           float* d_m, int64_t ld_dm, float* d_vec, float* d_dist, float* d_fc,
                                           float* d_h_in, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
  CHECK(check(w, g, "gnn_bwd"));
  if (scratch_bytes < petb200_gnn_scratch_bytes(w, g) - kAlign) {
    set_error("gnn_bwd: scratch buffer too small");
    return PETB200_ERR_WORKSPACE;
  }
  const int64_t E = g->n_edges, N = g->n_atoms, T = E + N;
  const int d = g->d, dn = g->d_node, nh = g->num_heads, prec = g->precision;
  Arena sa(const_cast<void*>(saved)), sc(scratch);
  const Saved S = carve_saved(sa, *w, *g);
  float* d_tp = sc.f32(T, d);          // rows [0, E): d t' ; rows [E, T): d y_c
  float* d_tbuf[2] = {sc.f32(T, d), sc.f32(T, d)};   // rows [0, E): d t ; rows [E, T): d (centre token)
  float* d_ugc = sc.f32(N, 4 * dn);
  float* d_xhc = sc.f32(N, dn);
  float* d_h1 = sc.f32(N, dn);
  float* d_yc = d_tp + E * d;
  float* d_o = sc.f32(T, d);
  float* d_qkv = sc.f32(T, 3 * d);
  float* dsum = sc.f32(T, nh);
  float* d_hbuf[2] = {sc.f32(N, dn), sc.f32(N, dn)};
  float* d_c1 = sc.f32(E, d);

  const float* d_h = d_h_out;
  const float* d_t = d_t_out;
  for (int k = w->n_tl - 1; k >= 0; --k) {
    const petb200_tl_weights& t = w->tl[k];
    const SavedLayer& L = S.tl[k];
    // edge feed-forward: t'' = t' + W_out swiglu(W_in rms(t'))
    CHECK(petb200_mlp_bwd(L.tp, d, d_t, d, t.mlp_image_bwd, t.b_in, E, d, t.d_ff, d_tp, d, stream));
    // centre feed-forward: h2 = h1 + Wc_out swiglu(Wc_in rms(h1))
    CHECK(gemm(d_h, dn, t.wc_out_t, d_ugc, 4 * dn, N, 2 * dn, dn, nullptr, nullptr, nullptr, 0, L.ugc, nullptr, 4 * dn,
               PETB200_EPI_SWIGLU_BWD, 0, prec, stream));
    CHECK(gemm(d_ugc, 4 * dn, t.wc_in_t, d_xhc, dn, N, dn, 4 * dn, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0,
               PETB200_EPI_NONE, 0, prec, stream));
    CHECK(petb200_rms_bwd(d_xhc, L.h1, L.rstd3, d_h, N, dn, d_h1, stream));
    // h1 = h + W_exp y_c ; t' = t + y_e ; y = W_o o
    CHECK(gemm(d_h1, dn, t.w_exp_t, d_yc, d, N, d, dn, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0,
               PETB200_EPI_NONE, 0, prec, stream));
    CHECK(gemm_tokens(d_tp, d, t.w_o_t, d_o, d, T, T, d, d, nullptr, nullptr, nullptr, 0, nullptr, 0, PETB200_EPI_NONE,
                      prec, stream));
    CHECK(petb200_attention_bwd(L.qkv, L.o, L.lse, d_o, row_ptr, fc, N, E, nh, d / nh, g->scale, g->max_row, prec, d_qkv,
                                d_fc, dsum, stream));
    // dgrad through the QKV projection and the RMSNorm in front of it (RMSNorm backward in the epilogue)
    float* d_t_new = d_tbuf[k & 1];
    float* d_c = d_t_new + E * d;
    CHECK(gemm_tokens(d_qkv, 3 * d, t.w_qkv_t, d_t_new, d, T, E, d, 3 * d, nullptr, L.rstd1, d_tp, d, L.x, d,
                      PETB200_EPI_RMS_BWD, prec, stream));
    if (k > 0 || d_h_in != nullptr) {
      float* d_h_new = k > 0 ? d_hbuf[k & 1] : d_h_in;
      CHECK(gemm(d_c, d, t.w_con_t, d_h_new, dn, N, dn, d, nullptr, nullptr, d_h1, dn, nullptr, nullptr, 0,
                 PETB200_EPI_NONE, 0, prec, stream));
      d_h = d_h_new;
    }
    d_t = d_t_new;
  }
  // token builder: t = W_2 silu(c_1) + b_2 ; c_1 = W_1m m + G (r, d) + Tbl[z_j] + b'
  if (w->compress_image_bwd != nullptr) {
    CHECK(petb200_compress_bwd(d_t, d, S.c1, w->compress_image_bwd, w->geo_fold, E, d, d_m, ld_dm, 1, d_vec, d_dist,
                               stream));
    return PETB200_OK;
  }
  CHECK(gemm(d_t, d, w->w2_t, d_c1, d, E, d, d, nullptr, nullptr, nullptr, 0, S.c1, nullptr, d, PETB200_EPI_MUL_DSILU, 0,
             prec, stream));
  CHECK(petb200_geom_embed_bwd(d_c1, d, w->geo_fold, E, d, 1, d_vec, d_dist, stream));
  if (d_m != nullptr)
    CHECK(gemm(d_c1, d, w->w1m_t, d_m, ld_dm, E, d, d, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0,
               PETB200_EPI_NONE, 1, prec, stream));
  return PETB200_OK;
}
