/*
 * libpetb200 — C ABI of the B200-native PET forward/backward engine.
 *
 * Drop-in boundary: these entry points are what a binding of the
 * reference's tensor backend `PETBackend` (metatensor/metatrain @ e2af0672,
 * src/metatrain/pet/modules/backend.py:12) would call instead of the torch ops it
 * executes today.  Every function
 *   - takes plain device pointers + sizes (no torch types), row-major, 16-byte aligned
 *     rows, fp32 features and int32 indices;
 *   - is asynchronous on the given CUDA stream, never synchronises, allocates, frees
 *     or retains a pointer past the call (the caller owns every buffer and workspace);
 *   - returns PETB200_OK (0) or a negative error code; the message is available from
 *     petb200_last_error() (thread local).  The Python host turns it into RuntimeError,
 *     which `mtt` wraps as ArchitectureError (src/metatrain/utils/errors.py:1-19).
 *
 * Edge layout: CSR over centre atoms, unpadded.  Edge e = (i -> j, S) lives at
 * row_ptr[i] <= e < row_ptr[i+1] in neighbor-list order (the stable order the
 * reference's NEF grid uses, src/metatrain/pet/modules/nef.py:34-85); col[e] = j,
 * ctr[e] = i, rev[e] = index of (j -> i, -S) (nef.py:88-166).
 *
 * Token layout inside a GNN layer: one [E + N, d_pet] matrix; rows [0, E) are the edge
 * tokens in CSR order, row E + i is the centre token of atom i
 * (transformer.py:214 concatenates [centre; edges] per atom).
 */
#ifndef PETB200_H
#define PETB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* petb200_stream_t; /* == cudaStream_t */

#if defined(__GNUC__)
#define PETB200_API __attribute__((visibility("default")))
#else
#define PETB200_API
#endif

enum {
  PETB200_OK = 0,
  PETB200_ERR_INVALID_ARGUMENT = -1,
  PETB200_ERR_CUDA = -2,
  PETB200_ERR_WORKSPACE = -3,
  PETB200_ERR_UNSUPPORTED = -4
};

/* GEMM epilogues */
enum {
  PETB200_EPI_NONE = 0,       /* C = rs*acc + bias (+ residual)                         */
  PETB200_EPI_SILU = 1,       /* pre = rs*acc + bias -> aux_out ; C = silu(pre) (+res)  */
  PETB200_EPI_SWIGLU = 2,     /* [u|g] = rs*acc + bias -> aux_out ; C = u * sigmoid(g)  */
  PETB200_EPI_MUL_DSILU = 3,  /* C = acc * silu'(aux_in)          (dgrad through SiLU)  */
  PETB200_EPI_SWIGLU_BWD = 4, /* C = [acc*sig(g) | acc*u*sig'(g)] (dgrad through SwiGLU)*/
  PETB200_EPI_RMS_BWD = 5,    /* C = residual + rs*acc - x*rs^3*(acc.x)/N : dgrad through the
                                 RMSNorm in front of the Linear (x = aux_in [M,N], rs = row_scale;
                                 N = 128; tensor-core precisions only)                         */
  PETB200_EPI_SILU_GEO = 6    /* SILU with a per-row geometry / table term in the pre-activation
                                 (petb200_compress_gemm only)                                   */
};

/* GEMM arithmetic */
enum {
  PETB200_PREC_FP32 = 0,   /* FFMA, exact fp32 (parity reference)                       */
  PETB200_PREC_BF16X3 = 1, /* tcgen05 kind::f16, 2-term bf16 split, 3 MMAs, fp32 accum  */
  PETB200_PREC_BF16 = 2    /* tcgen05 kind::f16, single bf16 pass (fast, ~1e-2 eV/A)    */
};

/* cutoff functions, src/metatrain/pet/modules/utilities.py:4-39 */
enum { PETB200_CUTOFF_BUMP = 0, PETB200_CUTOFF_COSINE = 1 };

PETB200_API const char* petb200_last_error(void);
PETB200_API int petb200_version(void);

/* ------------------------------------------------------------------ topology (a4-a6)
 * Replaces the integer half of compute_batch_tensors (structures.py:265-363):
 * drop pairs beyond the cutoff (non-strict list), count neighbours per centre, order
 * edges by centre (stable), find the reversed edge of every edge.                     */

/* Step 1: keep[e] = |r_j - r_i + S.cell| + 1e-15 <= cutoff (structures.py:220-221,267);
 * counts[i] += keep.  counts must be zeroed by the caller (N+1 entries, last unused). */
PETB200_API int petb200_nl_filter_count(const float* positions, const float* cells,
                            const int32_t* system_of_atom, const int32_t* centers,
                            const int32_t* neighbors, const int32_t* shifts,
                            int64_t n_pairs, int64_t n_atoms, float cutoff,
                            int32_t* keep, int32_t* counts, petb200_stream_t stream);

/* Bytes of scratch needed by petb200_csr_build for n_pairs / n_atoms. */
PETB200_API size_t petb200_csr_build_workspace(int64_t n_pairs, int64_t n_atoms);

/* Step 2: row_ptr = exclusive scan(counts) (N+1 entries; row_ptr[N] = E_kept),
 * stats[0] = E_kept, stats[1] = max_i counts[i]; perm[k] = index in the input list of
 * CSR edge k (stable by centre; only the first E_kept entries are meaningful).         */
PETB200_API int petb200_csr_build(const int32_t* centers, const int32_t* keep, const int32_t* counts,
                      int64_t n_pairs, int64_t n_atoms, int32_t* row_ptr, int32_t* perm,
                      int32_t* stats, void* workspace, size_t workspace_bytes,
                      petb200_stream_t stream);

/* Step 3: gather the kept pairs into CSR order.                                        */
PETB200_API int petb200_csr_gather(const int32_t* perm, const int32_t* centers,
                       const int32_t* neighbors, const int32_t* shifts, int64_t n_edges,
                       int32_t* ctr, int32_t* col, int32_t* shift_csr,
                       petb200_stream_t stream);

/* Step 4: rev[e] = index of (col[e] -> ctr[e], -S_e) (nef.py:88-166).  *n_missing is
 * incremented for every edge whose reverse does not exist (non-symmetric list, or — in an
 * atom-sharded run — a neighbour col[e] >= n_rows that has no CSR row on this rank).     */
PETB200_API int petb200_reverse_map(const int32_t* row_ptr, const int32_t* ctr, const int32_t* col,
                        const int32_t* shift_csr, int64_t n_edges, int64_t n_rows, int32_t* rev,
                        int32_t* n_missing, petb200_stream_t stream);

/* ------------------------------------------------------------ neighbor list (a1)
 * GPU replacement of vesin.ase_neighbor_list("ijS", ...) (src/metatrain/utils/neighbor_lists.py:131):
 * all ordered pairs (i, j, S) with |r_j + S.cell - r_i| <= cutoff (+2e-6 relative margin; the
 * model's own filter decides), no (i, i, 0), grouped by centre.  `cell_host` (9 floats, rows =
 * lattice vectors) and `origin_host` (3 floats, may be NULL) are HOST pointers; periodic = 1
 * (all three directions) or 0 (open: pass the bounding box as cell + origin).  Two phases
 * around one device->host read of offsets[n_atoms] (= number of pairs):
 *   nl_count: offsets[N+1] = exclusive scan of the per-atom neighbour counts;
 *   nl_fill : centers / neighbors [P], shifts [P,3].  The workspace must persist in between. */
PETB200_API int64_t petb200_nl_num_bins(const float* cell_host, int periodic, float cutoff,
                        int64_t n_atoms);
PETB200_API size_t petb200_nl_workspace(int64_t n_atoms, int64_t n_bins);
PETB200_API int petb200_nl_count(const float* positions, int64_t n_atoms, const float* cell_host,
                        const float* origin_host, int periodic, float cutoff, void* workspace,
                        size_t workspace_bytes, int32_t* offsets, petb200_stream_t stream);
PETB200_API int petb200_nl_fill(int64_t n_atoms, const float* cell_host, const float* origin_host,
                        int periodic, float cutoff, const void* workspace, size_t workspace_bytes,
                        const int32_t* offsets, int32_t* centers, int32_t* neighbors,
                        int32_t* shifts, petb200_stream_t stream);

/* CSR <-> padded NEF ([N, M, D], zero padded; nef.py:169-218).                         */
PETB200_API int petb200_csr_to_nef(const float* x_csr, const int32_t* row_ptr, int64_t n_atoms,
                       int64_t n_edges, int width_m, int d, float* x_nef,
                       petb200_stream_t stream);
PETB200_API int petb200_nef_to_csr(const float* x_nef, const int32_t* row_ptr, const int32_t* ctr,
                       int64_t n_atoms, int64_t n_edges, int width_m, int d, float* x_csr,
                       petb200_stream_t stream);

/* ------------------------------------------------------------------ geometry (a4, a7)
 * r_e = x_j - x_i + S.cell (structures.py:212-220); d_e = sqrt(r.r + 1e-15) (:330);
 * f_e = cutoff_function(|r_e| + 1e-15) (:221,306-316; utilities.py:4-39).              */
PETB200_API int petb200_edges_fwd(const float* positions, const float* cells,
                      const int32_t* system_of_atom, const int32_t* ctr, const int32_t* col,
                      const int32_t* shift_csr, int64_t n_edges, float cutoff, float width,
                      int cutoff_function, float* edge_vec, float* edge_dist,
                      float* cutoff_factor, petb200_stream_t stream);

/* Backward of petb200_edges_fwd + the force scatter: G_e = d_vec + d_dist*r/d +
 * d_fc*f'(|r|)*r/|r|;  d_pos[i] = sum_{e in row i} (G_rev(e) - G_e) (segmented reduce, no
 * atomics);  d_cells[b] += sum_e S_e^T G_e.  edge_grad is [E,3] scratch.               */
PETB200_API int petb200_edges_bwd(const float* d_vec, const float* d_dist, const float* d_fc,
                      const float* edge_vec, const float* edge_dist,
                      const int32_t* row_ptr, const int32_t* ctr, const int32_t* rev,
                      const int32_t* shift_csr, const int32_t* system_of_atom,
                      int64_t n_atoms, int64_t n_edges, float cutoff, float width,
                      int cutoff_function, float* edge_grad, float* d_pos, float* d_cells,
                      petb200_stream_t stream);

/* The same backward in two calls, for the atom-sharded path: edge gradients of halo edges
 * are exchanged between ranks in between (edge_grad has n_edges + n_ghost rows then, and
 * rev may point into the ghost rows).                                                   */
PETB200_API int petb200_edge_grad(const float* d_vec, const float* d_dist, const float* d_fc,
                      const float* edge_vec, const float* edge_dist, int64_t n_edges,
                      float cutoff, float width, int cutoff_function, float* edge_grad,
                      petb200_stream_t stream);
PETB200_API int petb200_force_scatter(const float* edge_grad, const int32_t* row_ptr,
                      const int32_t* ctr, const int32_t* rev, const int32_t* shift_csr,
                      const int32_t* system_of_atom, int64_t n_atoms, int64_t n_edges,
                      float* d_pos, float* d_cells, petb200_stream_t stream);

/* ------------------------------------------------------------- adaptive cutoff (a4)
 * get_adaptive_cutoffs_solver (src/metatrain/pet/modules/adaptive_cutoff.py:110-229) and its use in
 * compute_batch_tensors (structures.py:222-262), on the CSR rows of the pairs within the maximum
 * cutoff ("all" edges below); the pairs that survive the pair-cutoff mask form a second CSR
 * ("kept" edges, built with petb200_csr_build from the keep mask returned here).
 *   solve     : per atom r_root (root of n_total(r) = num_neighbors), dn_root = max(dn/dr, 1e-6)
 *               there, atomic_cutoff = clamp(r_root - residual/dn_root, R/16, R), pass = 1 where
 *               the clamp is inactive;
 *   pair_mask : pair_cutoff[e] = mean of the two atoms' cutoffs, keep[e] = |r_e| + 1e-15 <=
 *               pair_cutoff[e], counts[ctr] += keep (counts zeroed by the caller);
 *   edges_fwd_rc / edges_bwd_rc : petb200_edges_fwd / _bwd with one cutoff per edge; the backward
 *               also returns d_edge_cutoff[e] = -d_fc[e] * f'(|r_e|);
 *   bwd       : coef[i] = pass * 0.5 * sum_{e in kept row i} (d_pair_cutoff[e] + d_pair_cutoff[rev e])
 *               / dn_root[i];  d_dist_all[k] = coef[ctr k] * (d bump / d r)(d_k; r_root[ctr k]) —
 *               feed it to petb200_edges_bwd on the "all" topology as d_dist.                   */
PETB200_API int petb200_adaptive_cutoff_solve(const int32_t* row_ptr, const float* edge_dist, int64_t n_atoms,
                                  float num_neighbors, float max_cutoff, float width, float* r_root,
                                  float* dn_root, float* atomic_cutoff, float* pass,
                                  petb200_stream_t stream);
PETB200_API int petb200_adaptive_pair_mask(const int32_t* ctr, const int32_t* col, const float* edge_vec,
                               const float* atomic_cutoff, int64_t n_edges, float* pair_cutoff,
                               int32_t* keep, int32_t* counts, petb200_stream_t stream);
PETB200_API int petb200_edges_fwd_rc(const float* positions, const float* cells,
                         const int32_t* system_of_atom, const int32_t* ctr, const int32_t* col,
                         const int32_t* shift_csr, int64_t n_edges, const float* edge_cutoff,
                         float width, int cutoff_function, float* edge_vec, float* edge_dist,
                         float* cutoff_factor, petb200_stream_t stream);
PETB200_API int petb200_edges_bwd_rc(const float* d_vec, const float* d_dist, const float* d_fc,
                         const float* edge_vec, const float* edge_dist, const int32_t* row_ptr,
                         const int32_t* ctr, const int32_t* rev, const int32_t* shift_csr,
                         const int32_t* system_of_atom, int64_t n_atoms, int64_t n_edges,
                         const float* edge_cutoff, float width, int cutoff_function,
                         float* edge_grad, float* d_pos, float* d_cells, float* d_edge_cutoff,
                         petb200_stream_t stream);
PETB200_API int petb200_adaptive_cutoff_bwd(const int32_t* row_ptr_kept, const int32_t* rev_kept,
                                const float* d_pair_cutoff, const float* dn_root, const float* pass,
                                int64_t n_atoms, const int32_t* ctr_all, const float* dist_all,
                                const float* r_root, int64_t n_edges_all, float width, float* coef,
                                float* d_dist_all, petb200_stream_t stream);

/* The legacy "grid" method (adaptive_cutoff.py:232-395): cutoff_i = sum_p c_p w_ip with Gaussian
 * weights on how far the smoothed neighbour count at probe c_p = min_cutoff + p * spacing is from
 * the target.  solve also returns grad_d[i][q] = d cutoff_i / d D_q; bwd turns the gradient of the
 * pair cutoffs into d_dist_all on the "all" topology (`ones`: n_atoms floats equal to 1).          */
PETB200_API int petb200_adaptive_grid_solve(const int32_t* row_ptr, const float* edge_dist, int64_t n_atoms,
                                float num_neighbors, float width, float min_cutoff, float spacing,
                                int n_probes, float* atomic_cutoff, float* grad_d,
                                petb200_stream_t stream);
PETB200_API int petb200_adaptive_grid_bwd(const int32_t* row_ptr_kept, const int32_t* rev_kept,
                              const float* d_pair_cutoff, const float* ones, int64_t n_atoms,
                              const int32_t* ctr_all, const float* dist_all, const float* grad_d,
                              int64_t n_edges_all, float width, float min_cutoff, float spacing,
                              int n_probes, float* coef, float* d_dist_all, petb200_stream_t stream);

/* --------------------------------------------------------------- dense contractions
 * C[M,N] = epilogue(row_scale * (A[M,K] . W[N,K]^T) + bias) (+ residual) — every
 * torch.nn.Linear of transformer.py / backend.py, and its dgrad with W^T.             */
PETB200_API int petb200_gemm(const float* A, int64_t lda, const float* W, int64_t ldw, float* C,
                 int64_t ldc, int64_t M, int N, int K, const float* bias,
                 const float* row_scale, const float* residual, int64_t ldr,
                 const float* aux_in, float* aux_out, int64_t ld_aux, int epilogue,
                 int accumulate, int precision, petb200_stream_t stream);

/* Weight preparation for PETB200_PREC_BF16X3 / BF16: row r of `out` (same byte size as
 * the fp32 row) = [cols bf16 "hi" | cols bf16 "lo"], w = hi + lo + O(2^-17 w).  With those
 * precisions petb200_gemm expects its W argument in this format.                        */
PETB200_API int petb200_split_bf16(const float* w, int64_t rows, int cols, float* out,
                       petb200_stream_t stream);

/* ---------------------------------------------------- fused edge feed-forward block
 * One persistent tcgen05 kernel per direction for the PreLN feed-forward of a transformer
 * layer (transformer.py:229-232 with FeedForward :21-50, SwiGLU):
 *   fwd:  y   = x + W_out . swiglu(W_in . rmsnorm(x) + b_in) + b_out
 *   bwd:  d_x = d_y + (d rmsnorm/dx)^T W_in^T swiglu'(.) W_out^T d_y   (pre-activations are
 *         recomputed from x; nothing but x has to be kept from the forward pass)
 * w_in is [2*d_ff, d] with the RMSNorm weight folded into its columns (value rows first, gate
 * rows second: v, g = chunk(2)), w_out is [d, d_ff]; both fp32.  petb200_mlp_pack turns them
 * into the bf16 hi/lo operand-tile images (petb200_mlp_image_bytes bytes each, 16-byte
 * aligned, either pointer may be NULL) that the kernels stream from L2.  Built for d = 128,
 * d_ff a multiple of 64 up to 512; PETB200_ERR_UNSUPPORTED otherwise (callers then use
 * petb200_gemm + petb200_rms_*).  Arithmetic: bf16 2-term split, fp32 accumulation.        */
PETB200_API size_t petb200_mlp_image_bytes(int d_ff, int backward);
PETB200_API int petb200_mlp_pack(const float* w_in, const float* w_out, int d, int d_ff,
                     void* image_fwd, void* image_bwd, petb200_stream_t stream);
PETB200_API int petb200_mlp_fwd(const float* x, int64_t ldx, const void* image_fwd, const float* b_in,
                    const float* b_out, int64_t n_rows, int d, int d_ff, float* y, int64_t ldy,
                    petb200_stream_t stream);
PETB200_API int petb200_mlp_bwd(const float* x, int64_t ldx, const float* d_y, int64_t ld_dy,
                    const void* image_bwd, const float* b_in, int64_t n_rows, int d, int d_ff,
                    float* d_x, int64_t ld_dx, petb200_stream_t stream);

/* ------------------------------------------------ fused RMSNorm + Linear (QKV projection)
 * out = rmsnorm(x) . W^T + bias for d = 128 and n_out a multiple of 64 up to 1024
 * (AttentionBlock.input_linear after norm_attention, transformer.py:105-108, :218).  w is
 * [n_out, d] with the norm weight folded into its columns; petb200_norm_linear_pack builds the
 * operand-tile image (petb200_norm_linear_image_bytes bytes).  rstd_out (nullable, [n_rows])
 * receives rsqrt(mean(x^2) + eps) for the backward.  One persistent tcgen05 kernel: every
 * activation tile is normalised and converted once and produces all n_out columns.  The output leaves
 * through TMA tile stores: `out` must be 16-byte aligned and ldo a multiple of 4 floats
 * (PETB200_ERR_CUDA if the tensor map cannot be encoded).                                   */
PETB200_API size_t petb200_norm_linear_image_bytes(int n_out);
PETB200_API int petb200_norm_linear_pack(const float* w, int d, int n_out, void* image,
                             petb200_stream_t stream);
PETB200_API int petb200_norm_linear(const float* x, int64_t ldx, const void* image, const float* bias,
                        int64_t n_rows, int d, int n_out, float* out, int64_t ldo,
                        float* rstd_out, petb200_stream_t stream);

/* out[m,:] = table[idx[m],:] (torch.nn.Embedding, backend.py:515-516).                 */
PETB200_API int petb200_embedding(const float* table, const int32_t* idx, int64_t n_rows, int d,
                      float* out, int64_t ld_out, petb200_stream_t stream);

/* x[m,:] += table[idx[m],:]: per-system conditioning embedding added to the node features of the
 * atoms of each system (backend.py:551-552; conditioning.py:97-100).                              */
PETB200_API int petb200_add_gathered_rows(const float* table, const int32_t* idx, int64_t n_rows, int d,
                              float* x, int64_t ld_x, petb200_stream_t stream);

/* out[k][n] = in[n][k] * (col_scale ? col_scale[k] : 1): weight preparation
 * (dgrad operand W^T; RMSNorm weight folded into the following Linear).                */
PETB200_API int petb200_transpose_scale(const float* in, int rows, int cols, const float* col_scale,
                            float* out_transposed, float* out_scaled,
                            petb200_stream_t stream);

/* GNN-layer input (transformer.py:500-519): cat[e] = [W_geo.(r_e,d_e)+b | NbrEmb[z_j] |
 * m_e]; the neighbour-embedding block is absent when nbr_table is null (first layer).  */
PETB200_API int petb200_compress_input(const float* edge_vec, const float* edge_dist,
                           const float* w_geo, const float* b_geo, const float* nbr_table,
                           const int32_t* z_neighbor, const float* messages,
                           int64_t n_edges, int d, float* cat, petb200_stream_t stream);

/* The first Linear of the GNN-layer token builder with the concatenation folded away
 * (transformer.py:500-521):  W_1 . cat[geo | nbr | m] + b_1  =  W_1m . m  +  G . (r_e, d_e)
 * + Tbl[z_j] + b'  with G = W_1geo . W_geo [d, 4], Tbl = NbrEmb . W_1nbr^T [species, d] (nullable)
 * and b' = b_1 + W_1geo . b_geo prepared by the caller.  pre[e] (the pre-activation, kept for
 * the backward) and out[e] = silu(pre[e]) are [E, d]; w1m is the bf16 hi/lo split of W_1m
 * ([d, d], petb200_split_bf16).  One tcgen05 GEMM with K = d instead of a gather kernel plus a
 * GEMM with K = 2d / 3d.  precision: PETB200_PREC_BF16X3 or PETB200_PREC_BF16.               */
PETB200_API int petb200_compress_gemm(const float* messages, int64_t ld_m, const float* w1m_split,
                          const float* bias, const float* geo_w, const float* nbr_table,
                          const int32_t* z_neighbor, const float* edge_vec, const float* edge_dist,
                          int64_t n_edges, int d, float* pre, float* out, int precision,
                          petb200_stream_t stream);

/* d_(r,d)[e] (+)= W_geo^T . d_geo[e]  (backward of the 4 -> d geometry embedder).      */
PETB200_API int petb200_geom_embed_bwd(const float* d_geo, int64_t ld, const float* w_geo,
                           int64_t n_edges, int d, int accumulate, float* d_vec,
                           float* d_dist, petb200_stream_t stream);

/* ------------------------------------------------------------------------ RMSNorm
 * rstd[m] = rsqrt(mean(x[m,:]^2) + eps), eps = FLT_EPSILON (torch.nn.RMSNorm default;
 * transformer.py:184-186,193).  The norm weight is folded into the next Linear.        */
PETB200_API int petb200_rms_rstd(const float* x, int64_t n_rows, int d, float* rstd,
                     petb200_stream_t stream);
/* out = base + rstd*(d_xhat - xhat*mean(d_xhat*xhat)), xhat = x*rstd.                  */
PETB200_API int petb200_rms_bwd(const float* d_xhat, const float* x, const float* rstd,
                    const float* base, int64_t n_rows, int d, float* out,
                    petb200_stream_t stream);

/* RMSNorm as a standalone op (PostLN transformer, transformer.py:236-262): y = x * rstd * gamma
 * with rstd kept; backward out = base + rstd * (g - xhat * mean(g * xhat)), g = d_y * gamma
 * (base nullable).                                                                        */
PETB200_API int petb200_rms_norm_fwd(const float* x, const float* gamma, int64_t n_rows, int d, float* y,
                         float* rstd, petb200_stream_t stream);
PETB200_API int petb200_rms_norm_bwd(const float* d_y, const float* x, const float* rstd, const float* gamma,
                         const float* base, int64_t n_rows, int d, float* out,
                         petb200_stream_t stream);

/* torch.nn.LayerNorm(d) (eps 1e-5, affine) as a standalone op for normalization = "LayerNorm"
 * (transformer.py:181-186): y = (x - mean) * rstd * gamma + beta; backward
 * out = base + rstd * (g - mean(g) - xhat * mean(g * xhat)), g = d_y * gamma (base nullable).  */
PETB200_API int petb200_layer_norm_fwd(const float* x, const float* gamma, const float* beta, int64_t n_rows,
                           int d, float* y, float* mean, float* rstd, petb200_stream_t stream);
PETB200_API int petb200_layer_norm_bwd(const float* d_y, const float* x, const float* mean, const float* rstd,
                           const float* gamma, const float* base, int64_t n_rows, int d, float* out,
                           petb200_stream_t stream);

/* ---------------------------------------------------------------------- attention
 * Per-atom multi-head attention over tokens {centre i} U {edges of row i} with the
 * key-only additive bias log(max(w_q, 1e-15)), w = 1 for the centre token and f_e for
 * edge tokens (transformer.py:86-152, 524-540).  qkv is [E+N, 3*d] = [q | k | v], each
 * split into num_heads heads of 16; out is [E+N, d]; lse is [E+N, num_heads] (base-2
 * units).  precision = PETB200_PREC_FP32: packed-fp32 CUDA-core kernels (any row length that
 * fits shared memory); otherwise, for rows of at most 63 neighbours, warp-level tensor-core
 * kernels (mma.sync bf16, 2-term operand split, fp32 accumulation), falling back to the
 * CUDA-core kernels for longer rows.                                                     */
PETB200_API int petb200_attention_fwd(const float* qkv, const int32_t* row_ptr,
                          const float* cutoff_factor, int64_t n_atoms, int64_t n_edges,
                          int num_heads, int head_dim, float scale, int max_row,
                          int precision, float* out, float* lse, petb200_stream_t stream);
/* d_qkv from d_out; d_fc[e] += (sum_{heads,queries} dS[.,e]) / f_e  (f_e > 1e-15).
 * dsum: (E+N) * num_heads floats of scratch (fp32 path: row sums dO.O passed between its two
 * kernels; tensor-core path: per-head key-bias gradients [num_heads, E+N], reduced over the heads
 * by a second small kernel so that the heads of an atom never synchronise).                     */
PETB200_API int petb200_attention_bwd(const float* qkv, const float* out, const float* lse,
                          const float* d_out, const int32_t* row_ptr,
                          const float* cutoff_factor, int64_t n_atoms, int64_t n_edges,
                          int num_heads, int head_dim, float scale, int max_row,
                          int precision, float* d_qkv, float* d_fc, float* dsum,
                          petb200_stream_t stream);

/* ------------------------------------------------ message reversal + combine (a8)
 * backend.py:559-575: cc[e] = LayerNorm_{2d}(cat[t_e, t_rev(e)]) (eps 1e-5, affine).
 * This is the HBM-bound "edge scatter" kernel: 2 KB of algorithmic traffic per edge.   */
PETB200_API int petb200_combine_ln_fwd(const float* t, const int32_t* rev, const float* gamma,
                           const float* beta, int64_t n_edges, int d, float* cc,
                           float* mean, float* rstd, petb200_stream_t stream);
PETB200_API int petb200_combine_ln_bwd(const float* d_cc, const float* t, const int32_t* rev,
                           const float* gamma, const float* mean, const float* rstd,
                           int64_t n_edges, int d, float* d_cat, petb200_stream_t stream);
/* The same block as ONE persistent tcgen05 kernel per direction (combine_fused.cu): the message
 * reversal gather ("edge scatter"), the LayerNorm, both Linears of the combine MLP and the
 * residual update
 *   fwd:  m_io[e] += t[e] + W_b . silu(W_a . LN(cat[t[e], t[rev[e]]]) + b_a) + b_b
 * with the LayerNorm folded into the first contraction:  W_a . LN(c) + b_a = r (W' c - mu s) + b',
 * W' = W_a diag(gamma) (`w_a_folded`, [2d, 2d]), s = row sums of W' (`s_vec`), b' = W_a beta + b_a
 * (`b_fold`).  Side outputs for the backward: the pre-activations `p` and (mu, r) per edge (`stats`,
 * [E, 2]).  `p` is an opaque buffer of ceil(E / 128) * 128 * 2d floats in the kernels' private layout
 * [128-edge tile][32-unit chunk][unit][edge of the tile] (the row-per-thread epilogues of both kernels
 * then read / write whole 128 B lines); element (e, n) lives at
 * ((e / 128 * 8 + n / 32) * 32 + n % 32) * 128 + e % 128.  `t` may hold ghost rows behind its first n_edges rows (atom-sharded runs):
 * rev indexes rows of `t`.
 *   bwd:  d_cat[e] = LN'(cat_e)^T W_a^T silu'(p_e) W_b^T g[e]   ([E, 2d]; finish with
 *         petb200_combine_scatter_bwd).
 * petb200_combine_pack builds the bf16 hi/lo operand-tile images (petb200_combine_image_bytes bytes
 * each) from w_a_folded and w_b [d, 2d].  Built for d = 128; PETB200_ERR_UNSUPPORTED otherwise.
 * Replaces, per GNN layer: backend.py:559-575 (reference), and the unfused sequence
 * combine_ln_fwd + 2 x petb200_gemm (forward) / 2 x petb200_gemm + combine_ln_bwd (backward).      */
PETB200_API size_t petb200_combine_image_bytes(int d, int backward);
PETB200_API int petb200_combine_pack(const float* w_a_folded, const float* w_b, int d, void* image_fwd,
                         void* image_bwd, petb200_stream_t stream);
PETB200_API int petb200_combine_fwd(const float* t, int64_t ld_t, const int32_t* rev, const void* image_fwd,
                        const float* s_vec, const float* b_fold, const float* b_out, int64_t n_edges,
                        int d, float* m_io, int64_t ld_m, float* p_out, float* stats_out,
                        petb200_stream_t stream);
PETB200_API int petb200_combine_bwd(const float* g, int64_t ld_g, const float* p, const float* t, int64_t ld_t,
                        const int32_t* rev, const float* stats, const void* image_bwd,
                        const float* s_vec, const float* b_fold, int64_t n_edges, int d, float* d_cat,
                        petb200_stream_t stream);
/* out[e] = base[e] + d_cat[e, :d] + d_cat[rev[e], d:]  (rev is an involution, so the
 * scatter of the reversed half is a gather: no atomics).                               */
PETB200_API int petb200_combine_scatter_bwd(const float* d_cat, const float* base, const int32_t* rev,
                                int64_t n_edges, int d, float* out,
                                petb200_stream_t stream);

/* ------------------------------------------------ residual featurizer (a8, second variant)
 * backend.py:589-649: the input messages of GNN layer l+1 are the average of layer l's input
 * messages and its reversed output tokens, out[e] = 0.5 (m[e] + t[rev[e]]).  Backward:
 * d_t[e] += 0.5 d_next[rev[e]] (rev is an involution: a gather), d_m[e] = 0.5 d_next[e].     */
PETB200_API int petb200_avg_reverse_fwd(const float* m, const float* t, const int32_t* rev,
                            int64_t n_edges, int d, float* out, petb200_stream_t stream);
PETB200_API int petb200_avg_reverse_bwd(const float* d_next, const int32_t* rev, int64_t n_edges, int d,
                            float* d_t, float* d_m, petb200_stream_t stream);

/* ------------------------------------------------------------------- readout (a12)
 * backend.py:195-217, 762-772: atomic[i,p] = w_n[p].n2[i] + b_n[p] +
 * sum_{e in row i} f_e * (w_e[p].e2[e] + b_e[p]);  edge_pred[e,p] is kept for backward. */
/* (edge_feat may be NULL: edge_pred then holds precomputed edge predictions, petb200_edge_head_fwd) */
PETB200_API int petb200_readout_fwd(const float* node_feat, const float* edge_feat, const float* w_node,
                        const float* b_node, const float* w_edge, const float* b_edge,
                        const float* cutoff_factor, const int32_t* row_ptr, int64_t n_atoms,
                        int64_t n_edges, int d, int n_out, float* atomic, float* edge_pred,
                        petb200_stream_t stream);
/* node_pre / edge_pre (nullable): pre-activations of the SiLU that produced node_feat /
 * edge_feat; when given, the returned gradients are w.r.t. those pre-activations.
 * d_fc[e] += sum_p edge_pred[e,p] * d_atomic[ctr[e],p].                                 */
PETB200_API int petb200_readout_bwd(const float* d_atomic, const float* edge_pred, const float* w_node,
                        const float* w_edge, const float* cutoff_factor, const int32_t* ctr,
                        const float* node_pre, const float* edge_pre,
                        int64_t n_atoms, int64_t n_edges, int d, int n_out, float* d_node_feat,
                        float* d_edge_feat, float* d_fc, petb200_stream_t stream);

/* ------------------------------------------- fused 128 -> 128 -> 128 chains (chain_fused.cu)
 * Two Linears with a SiLU between them as ONE persistent tcgen05 kernel per direction, for the two
 * places PET has that shape on its edge rows.  petb200_chain_pack builds the operand-tile images
 * (petb200_chain_image_bytes bytes each) from w1 [128, 128] and w2 [128, 128]; d = 128 only.
 * `e1p` / `c1` (pre-activation of the first Linear) is an opaque buffer of ceil(E / 128) * 128 * 128
 * floats in the kernels' private tile layout (see petb200_combine_fwd).
 *
 * Edge head (backend.py:171-217, 762-772; single-property targets):
 *   fwd:  e2p = W_2 silu(W_1 m + b_1) + b_2  [E, 128],  edge_pred[e] = w_e . silu(e2p[e]) + b_e
 *         (finish with petb200_readout_fwd, edge_feat = NULL: atomic = node part + sum_j f_ij pe_ij)
 *   bwd:  d_m = W_1^T silu'(e1p) W_2^T g,  g[e] = d_atomic[ctr e] f_e w_e silu'(e2p[e]) (formed on the
 *         fly), d_fc[e] += d_atomic[ctr e] edge_pred[e]
 * replacing 2 x petb200_gemm + readout (forward) and readout_bwd + 2 x petb200_gemm (backward).
 *
 * Token builder of a CartesianTransformer (transformer.py:500-521; arguments as petb200_compress_gemm):
 *   fwd:  c_1 = W_1m m + G (r, d) + Tbl[z_j] + b' ;  t = W_2 silu(c_1) + b_2
 *   bwd:  d_c1 = silu'(c_1) W_2^T d_t ;  d_m (+)= W_1m^T d_c1 (skipped when d_m is NULL) ;
 *         d_vec / d_dist += G^T d_c1
 * replacing petb200_compress_gemm + petb200_gemm (forward) and 2 x petb200_gemm +
 * petb200_geom_embed_bwd (backward).                                                              */
/* (the row-major outputs e2p / t_out / d_m of the four chain kernels are written by TMA tile stores —
 *  d_m by a TMA reduce-add when it accumulates: they must be 16-byte aligned with leading dimensions
 *  that are multiples of 4 floats) */
PETB200_API size_t petb200_chain_image_bytes(int d);
PETB200_API int petb200_chain_pack(const float* w1, const float* w2, int d, void* image_fwd, void* image_bwd,
                       petb200_stream_t stream);
PETB200_API int petb200_edge_head_fwd(const float* m, int64_t ld_m, const void* image_fwd, const float* b1,
                          const float* b2, const float* w_e, float b_e, int64_t n_edges, int d,
                          float* e1p, float* e2p, float* edge_pred, petb200_stream_t stream);
PETB200_API int petb200_edge_head_bwd(const float* d_atomic, const int32_t* ctr, const float* cutoff_factor,
                          const float* e1p, const float* e2p, const float* edge_pred,
                          const void* image_bwd, const float* w_e, int64_t n_edges, int d, float* d_m,
                          int64_t ld_dm, float* d_fc, petb200_stream_t stream);
PETB200_API int petb200_compress_fwd(const float* messages, int64_t ld_m, const void* image_fwd,
                         const float* b_fold, const float* geo_w, const float* nbr_table,
                         const int32_t* z_neighbor, const float* edge_vec, const float* edge_dist,
                         const float* b2, int64_t n_edges, int d, float* c1, float* t_out, int64_t ld_t,
                         petb200_stream_t stream);
PETB200_API int petb200_compress_bwd(const float* d_t, int64_t ld_dt, const float* c1, const void* image_bwd,
                         const float* geo_w, int64_t n_edges, int d, float* d_m, int64_t ld_dm,
                         int accumulate, float* d_vec, float* d_dist, petb200_stream_t stream);

/* ------------------------------------------------------ stage-level schedule (schedule.cu)
 * One call enqueues the whole kernel sequence of a stage on the stream, from C++, with caller-owned
 * buffers: what the Python host would otherwise issue as ~30 separate entry-point calls per GNN
 * layer and direction.  Built for the default layer (PreLN + RMSNorm + SwiGLU, d_pet = 128,
 * tensor-core precisions); every other variant keeps the per-op schedule of the host.
 *
 * petb200_gnn_fwd / _bwd = one CartesianTransformer (transformer.py:463-562 with its
 * TransformerLayers :203-234) on the [E + N] token layout: token builder, and per attention layer
 * centre contraction, RMSNorm + QKV, attention, output projection, centre expansion + centre
 * feed-forward, edge feed-forward.  Weight matrices are passed in the bf16 hi/lo split format of
 * petb200_split_bf16 (`w` = device pointer, `ld` = leading dimension in floats).                  */
typedef struct petb200_mat {
  const float* w;
  int64_t ld;
} petb200_mat;

typedef struct petb200_tl_weights {   /* one TransformerLayer (transformer.py:155-262) */
  const void* qkv_image;              /* petb200_norm_linear_pack of W_qkv diag(gamma_attention) */
  const float* b_qkv;
  petb200_mat w_qkv_t;                /* [d, 3d]: (W_qkv diag(gamma_attention))^T */
  petb200_mat w_o, w_o_t;             /* attention.output_linear [d, d] and its transpose */
  const float* b_o;
  const void* mlp_image_fwd;          /* petb200_mlp_pack images of the edge feed-forward */
  const void* mlp_image_bwd;
  const float* b_in;
  const float* b_out;
  int d_ff;                           /* hidden width of the edge feed-forward (after SwiGLU) */
  petb200_mat w_con, w_con_t;         /* center_contraction [d, d_node] */
  const float* b_con;
  petb200_mat w_exp, w_exp_t;         /* center_expansion [d_node, d] */
  const float* b_exp;
  petb200_mat wc_in, wc_in_t;         /* center_mlp.w_in diag(gamma_center) [4 d_node, d_node] */
  const float* bc_in;
  petb200_mat wc_out, wc_out_t;       /* center_mlp.w_out [d_node, 2 d_node] */
  const float* bc_out;
} petb200_tl_weights;

typedef struct petb200_gnn_weights {  /* one CartesianTransformer */
  petb200_mat w1m;                    /* compress[0][:, -d:] (message columns), split */
  petb200_mat w1m_t;                  /* its transpose [d, d], split */
  const float* b_fold;                /* b_1 + W_1geo b_geo */
  const float* geo_fold;              /* W_1geo W_geo [d, 4] */
  const float* nbr_fold;              /* NbrEmb W_1nbr^T [S, d] or NULL (first GNN layer) */
  petb200_mat w2, w2_t;               /* compress[2] [d, d] */
  const float* b2;
  const void* compress_image_fwd;     /* petb200_chain_pack(w1m, w2) images (NULL: unfused token builder) */
  const void* compress_image_bwd;
  int n_tl;
  const petb200_tl_weights* tl;       /* HOST array of n_tl entries */
} petb200_gnn_weights;

typedef struct petb200_dims {
  int64_t n_atoms, n_edges, n_ghost;  /* ghost rows follow the [E | N] token rows (atom-sharded runs) */
  int d, d_node, num_heads, max_row, precision;
  float scale;                        /* 1 / (sqrt(head_dim) * attention_temperature) */
} petb200_dims;

/* bytes of the forward's saved-for-backward buffer and of the scratch either direction needs */
PETB200_API size_t petb200_gnn_saved_bytes(const petb200_gnn_weights* w, const petb200_dims* dims);
PETB200_API size_t petb200_gnn_scratch_bytes(const petb200_gnn_weights* w, const petb200_dims* dims);
/* x_out: [(E + N + n_ghost), d] token matrix after the last attention layer (rows [0, E) = the
 * layer's output edge tokens); h_out [N, d_node].                                               */
PETB200_API int petb200_gnn_fwd(const petb200_gnn_weights* w, const petb200_dims* dims, const int32_t* row_ptr,
                    const int32_t* z_neighbors, const float* edge_vec, const float* edge_dist,
                    const float* cutoff_factor, const float* h_in, const float* m_in, int64_t ld_m,
                    float* x_out, float* h_out, void* saved, size_t saved_bytes, void* scratch,
                    size_t scratch_bytes, petb200_stream_t stream);
/* d_h [N, d_node] and d_t [E, d]: gradients of h_out and of the output edge tokens.  Accumulates
 * into d_vec [E, 3], d_dist [E], d_fc [E]; adds the gradient w.r.t. m_in to d_m when d_m is not
 * NULL; writes the gradient w.r.t. h_in to d_h_in when it is not NULL.                           */
PETB200_API int petb200_gnn_bwd(const petb200_gnn_weights* w, const petb200_dims* dims, const int32_t* row_ptr,
                    const float* cutoff_factor, const void* saved, const float* d_h, const float* d_t,
                    float* d_m, int64_t ld_dm, float* d_vec, float* d_dist, float* d_fc, float* d_h_in,
                    void* scratch, size_t scratch_bytes, petb200_stream_t stream);

/* per-structure sums, src/metatrain/utils/sum_over_atoms.py:31 (deterministic: atoms of a
 * structure are contiguous; one warp per structure).                                    */
PETB200_API int petb200_sum_over_atoms(const float* atomic, const int32_t* struct_ptr,
                           int64_t n_structures, int n_out, float* energies,
                           petb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PETB200_H */
