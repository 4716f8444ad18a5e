// TorchScript custom operators over libpetb200 (SURVEY.md 8(b) level B3).
//
// The reference ships models as TorchScript (`AtomisticModel(self.eval(), ...).save(...)`,
// src/metatrain/pet/model.py:990-1021; src/metatrain/cli/export.py:243-266 collects the shared
// libraries of the custom operators a scripted model uses into the `extensions/` directory, and
// src/metatrain/utils/io.py:183-184 loads them back).  A scripted module cannot call Python
// autograd.Functions or ctypes, so the hot path is registered here as operators:
//
//   petb200::topology(...)   -> CSR topology of a batch (the integer half of preprocess, a4-a6)
//   petb200::pet_atomic(...) -> per-atom predictions [N, P], differentiable w.r.t. positions and
//                               cells through a C++ torch::autograd::Function
//
// Both only sequence entry points of libpetb200.so (include/petb200.h) on torch's current CUDA
// stream with torch-allocated buffers; no kernel lives here.  The packed weights travel as a flat
// list of tensors (metatrain_b200/export.py writes it, order documented there and mirrored by
// `Weights` below).  Built for the default PET configuration (PreLN + RMSNorm + SwiGLU, feedforward
// featurizer, fixed cutoff, one readout layer, tensor-core precision).
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/script.h>
#include <torch/torch.h>

#include <vector>

#include "../../include/petb200.h"

namespace {

using torch::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

cudaStream_t current_stream() {
  return reinterpret_cast<cudaStream_t>(c10::cuda::getCurrentCUDAStream().stream());
}
petb200_stream_t S() { return reinterpret_cast<petb200_stream_t>(current_stream()); }

void ok(int status, const char* what) {
  TORCH_CHECK(status == 0, "petb200_", what, " failed (", status, "): ", petb200_last_error());
}
const float* F(const Tensor& t) { return t.defined() && t.numel() > 0 ? t.data_ptr<float>() : nullptr; }
float* Fm(Tensor& t) { return t.defined() && t.numel() > 0 ? t.data_ptr<float>() : nullptr; }
const int32_t* I(const Tensor& t) { return t.numel() > 0 ? t.data_ptr<int32_t>() : nullptr; }
const void* B(const Tensor& t) { return t.data_ptr(); }
petb200_mat M(const Tensor& t) { return petb200_mat{t.data_ptr<float>(), t.stride(0)}; }
Tensor f32(at::IntArrayRef shape, const Tensor& like) {
  return torch::empty(shape, like.options().dtype(torch::kFloat32).requires_grad(false));
}
Tensor bytes(size_t n, const Tensor& like) {
  return torch::empty({(int64_t)(n > 16 ? n : 16)}, like.options().dtype(torch::kUInt8).requires_grad(false));
}

// meta ints / floats (metatrain_b200/export.py)
enum { M_NGNN, M_NTL, M_D, M_DN, M_NH, M_DFF, M_DH, M_NOUT, M_PREC, M_CUTFN, M_COUNT };
enum { F_CUTOFF, F_WIDTH, F_TEMP, F_BE, F_COUNT };   // F_BE: bias of the edge last layer (fused edge head)
constexpr int GNN_T = 10, TL_T = 22, COMB_T = 5, HEAD_T = 18;

// views into the flat weight list
struct Weights {
  const std::vector<Tensor>& w;
  int n_gnn, n_tl;
  int gnn0(int l) const { return l * (GNN_T + n_tl * TL_T + COMB_T); }
  int tl0(int l, int k) const { return gnn0(l) + GNN_T + k * TL_T; }
  int comb0(int l) const { return gnn0(l) + GNN_T + n_tl * TL_T; }
  int tail0() const { return n_gnn * (GNN_T + n_tl * TL_T + COMB_T); }
  const Tensor& node_emb() const { return w[tail0()]; }
  const Tensor& edge_emb() const { return w[tail0() + 1]; }
  const Tensor& head(int i) const { return w[tail0() + 2 + i]; }
  void fill(int l, int d_ff, petb200_gnn_weights& g, std::vector<petb200_tl_weights>& tls) const {
    const int o = gnn0(l);
    g.w1m = M(w[o]); g.w1m_t = M(w[o + 1]); g.b_fold = F(w[o + 2]); g.geo_fold = F(w[o + 3]);
    g.nbr_fold = F(w[o + 4]); g.w2 = M(w[o + 5]); g.w2_t = M(w[o + 6]); g.b2 = F(w[o + 7]);
    g.compress_image_fwd = w[o + 8].numel() > 0 ? B(w[o + 8]) : nullptr;
    g.compress_image_bwd = w[o + 9].numel() > 0 ? B(w[o + 9]) : nullptr;
    tls.resize(n_tl);
    for (int k = 0; k < n_tl; ++k) {
      const int q = tl0(l, k);
      petb200_tl_weights& t = tls[k];
      t.qkv_image = B(w[q]); t.b_qkv = F(w[q + 1]); t.w_qkv_t = M(w[q + 2]);
      t.w_o = M(w[q + 3]); t.w_o_t = M(w[q + 4]); t.b_o = F(w[q + 5]);
      t.mlp_image_fwd = B(w[q + 6]); t.mlp_image_bwd = B(w[q + 7]); t.b_in = F(w[q + 8]); t.b_out = F(w[q + 9]);
      t.d_ff = d_ff;
      t.w_con = M(w[q + 10]); t.w_con_t = M(w[q + 11]); t.b_con = F(w[q + 12]);
      t.w_exp = M(w[q + 13]); t.w_exp_t = M(w[q + 14]); t.b_exp = F(w[q + 15]);
      t.wc_in = M(w[q + 16]); t.wc_in_t = M(w[q + 17]); t.bc_in = F(w[q + 18]);
      t.wc_out = M(w[q + 19]); t.wc_out_t = M(w[q + 20]); t.bc_out = F(w[q + 21]);
    }
    g.n_tl = n_tl;
    g.tl = tls.data();
  }
};

void gemm(const Tensor& a, const Tensor& w, Tensor& out, const Tensor* bias, int epilogue, const Tensor* aux_in,
          Tensor* aux_out, int prec) {
  const Tensor* aux = aux_in ? aux_in : aux_out;
  ok(petb200_gemm(F(a), a.stride(0), F(w), w.stride(0), Fm(out), out.stride(0), a.size(0), (int)w.size(0),
                  (int)a.size(1), bias ? F(*bias) : nullptr, nullptr, nullptr, 0, aux_in ? F(*aux_in) : nullptr,
                  aux_out ? Fm(*aux_out) : nullptr, aux ? aux->stride(0) : 0, epilogue, 0, prec, S()),
     "gemm");
}

// --------------------------------------------------------------------------- topology
// inputs as PETBackend.preprocess receives them (backend.py:238); z_nodes = species index per atom.
// Returns [row_ptr, ctr, col, rev, shift, z_neighbors, system_of_atom, z_nodes32, max_row (CPU int64 [1])].
std::vector<Tensor> topology(const Tensor& positions, const Tensor& cells, const Tensor& centers,
                             const Tensor& neighbors, const Tensor& cell_shifts, const Tensor& system_indices,
                             const Tensor& z_nodes, double cutoff) {
  TORCH_CHECK(positions.is_cuda(), "petb200::topology: expected CUDA tensors; this backend has no CPU path");
  c10::cuda::CUDAGuard guard(positions.device());
  const auto i32 = positions.options().dtype(torch::kInt32).requires_grad(false);
  const int64_t n_atoms = positions.size(0), n_pairs = centers.size(0);
  Tensor pos = positions.detach().to(torch::kFloat32).contiguous();
  Tensor cel = cells.detach().to(torch::kFloat32).contiguous();
  Tensor cen = centers.to(torch::kInt32).contiguous(), nei = neighbors.to(torch::kInt32).contiguous();
  Tensor shf = cell_shifts.to(torch::kInt32).contiguous(), sys = system_indices.to(torch::kInt32).contiguous();
  Tensor z32 = z_nodes.to(torch::kInt32).contiguous();
  Tensor keep = torch::empty({n_pairs > 0 ? n_pairs : 1}, i32), counts = torch::zeros({n_atoms + 1}, i32);
  ok(petb200_nl_filter_count(F(pos), F(cel), I(sys), I(cen), I(nei), I(shf), n_pairs, n_atoms, (float)cutoff,
                             keep.data_ptr<int32_t>(), counts.data_ptr<int32_t>(), S()),
     "nl_filter_count");
  const size_t ws = petb200_csr_build_workspace(n_pairs, n_atoms);
  Tensor work = bytes(ws, positions), row_ptr = torch::empty({n_atoms + 1}, i32);
  Tensor perm = torch::empty({n_pairs > 0 ? n_pairs : 1}, i32), stats = torch::zeros({4}, i32);
  ok(petb200_csr_build(I(cen), keep.data_ptr<int32_t>(), counts.data_ptr<int32_t>(), n_pairs, n_atoms,
                       row_ptr.data_ptr<int32_t>(), perm.data_ptr<int32_t>(), stats.data_ptr<int32_t>(),
                       work.data_ptr(), ws, S()),
     "csr_build");
  stats.select(0, 3).copy_((z32 < 0).any());
  Tensor host = stats.cpu();   // the stage's one device -> host read (the reference syncs here too)
  const int64_t n_edges = host[0].item<int32_t>(), max_row = host[1].item<int32_t>();
  TORCH_CHECK(host[3].item<int32_t>() == 0, "atomic types outside the model's atomic_types in the input");
  Tensor ctr = torch::empty({n_edges}, i32), col = torch::empty({n_edges}, i32);
  Tensor shift = torch::empty({n_edges, 3}, i32), rev = torch::empty({n_edges}, i32);
  if (n_edges > 0) {
    ok(petb200_csr_gather(perm.data_ptr<int32_t>(), I(cen), I(nei), I(shf), n_edges, ctr.data_ptr<int32_t>(),
                          col.data_ptr<int32_t>(), shift.data_ptr<int32_t>(), S()),
       "csr_gather");
    ok(petb200_reverse_map(row_ptr.data_ptr<int32_t>(), I(ctr), I(col), I(shift), n_edges, n_atoms,
                           rev.data_ptr<int32_t>(), stats.data_ptr<int32_t>() + 2, S()),
       "reverse_map");
    const int64_t missing = stats.cpu()[2].item<int32_t>();
    if (missing != 0) {
      const std::string msg = "neighbor list is not symmetric: " + std::to_string(missing) +
                              " edges have no reversed edge (PET needs a full list)";
      TORCH_CHECK(false, msg);
    }
  }
  Tensor z_nb = n_edges > 0 ? z32.index_select(0, col.to(torch::kInt64)).contiguous() : torch::empty({0}, i32);
  return {row_ptr, ctr, col, rev, shift, z_nb, sys, z32, torch::full({1}, max_row, torch::kInt64)};
}

// ------------------------------------------------------------------------- pet_atomic
struct PetAtomic : public torch::autograd::Function<PetAtomic> {
  static Tensor forward(AutogradContext* ctx, const Tensor& positions, const Tensor& cells,
                        std::vector<Tensor> topo, std::vector<Tensor> weights, std::vector<int64_t> meta,
                        std::vector<double> fmeta) {
    TORCH_CHECK(positions.is_cuda(), "petb200::pet_atomic: expected CUDA tensors; this backend has no CPU path");
    TORCH_CHECK((int)meta.size() == M_COUNT && (int)fmeta.size() == F_COUNT && topo.size() == 9, "petb200::pet_atomic: bad metadata");
    c10::cuda::CUDAGuard guard(positions.device());
    const int n_gnn = (int)meta[M_NGNN], n_tl = (int)meta[M_NTL], d = (int)meta[M_D], dn = (int)meta[M_DN];
    const int nh = (int)meta[M_NH], dh = (int)meta[M_DH], n_out = (int)meta[M_NOUT], prec = (int)meta[M_PREC];
    const Weights W{weights, n_gnn, n_tl};
    TORCH_CHECK((int)weights.size() == W.tail0() + 2 + HEAD_T, "petb200::pet_atomic: weight list has the wrong length");
    const Tensor &row_ptr = topo[0], &ctr = topo[1], &col = topo[2], &rev = topo[3], &shift = topo[4];
    const Tensor &z_nb = topo[5], &sys = topo[6], &z_nodes = topo[7];
    const int64_t N = positions.size(0), E = ctr.size(0);
    const int max_row = (int)topo[8].item<int64_t>();
    TORCH_CHECK(max_row + 1 <= 64, "petb200::pet_atomic: rows of more than 63 neighbours need the eager backend");
    Tensor pos = positions.detach().to(torch::kFloat32).contiguous();
    Tensor cel = cells.detach().to(torch::kFloat32).contiguous();

    // a4 / a7: edge vectors, distances, cutoff factors
    Tensor vec = f32({E, 3}, pos), dist = f32({E}, pos), fc = f32({E}, pos);
    ok(petb200_edges_fwd(F(pos), F(cel), I(sys), I(ctr), I(col), I(shift), E, (float)fmeta[F_CUTOFF],
                         (float)fmeta[F_WIDTH], (int)meta[M_CUTFN], Fm(vec), Fm(dist), Fm(fc), S()),
       "edges_fwd");
    // a8: embeddings, GNN layers, message updates
    Tensor h = f32({N, dn}, pos), m = f32({E, d}, pos);
    ok(petb200_embedding(F(W.node_emb()), I(z_nodes), N, dn, Fm(h), dn, S()), "embedding");
    ok(petb200_embedding(F(W.edge_emb()), I(z_nb), E, d, Fm(m), d, S()), "embedding");
    petb200_dims dims{N, E, 0, d, dn, nh, max_row, prec,
                      (float)(1.0 / (std::sqrt((double)(d / nh)) * fmeta[F_TEMP]))};
    std::vector<Tensor> keep;   // per layer: saved, x (token matrix), p1, cstats
    for (int l = 0; l < n_gnn; ++l) {
      petb200_gnn_weights g{};
      std::vector<petb200_tl_weights> tls;
      W.fill(l, (int)meta[M_DFF], g, tls);
      Tensor saved = bytes(petb200_gnn_saved_bytes(&g, &dims), pos);
      Tensor scratch = bytes(petb200_gnn_scratch_bytes(&g, &dims), pos);
      Tensor x = f32({E + N, d}, pos), h_out = f32({N, dn}, pos);
      if (E > 0 || N > 0)
        ok(petb200_gnn_fwd(&g, &dims, I(row_ptr), I(z_nb), F(vec), F(dist), F(fc), F(h), F(m), d, Fm(x), Fm(h_out),
                           saved.data_ptr(), saved.numel(), scratch.data_ptr(), scratch.numel(), S()),
           "gnn_fwd");
      h = h_out;
      const int c0 = W.comb0(l);
      Tensor p1 = f32({(E + 127) / 128 * 128, 2 * d}, pos), cstats = f32({E, 2}, pos);
      ok(petb200_combine_fwd(F(x), d, I(rev), B(weights[c0]), F(weights[c0 + 2]), F(weights[c0 + 3]),
                             F(weights[c0 + 4]), E, d, Fm(m), d, Fm(p1), Fm(cstats), S()),
         "combine_fwd");
      keep.insert(keep.end(), {saved, x, p1, cstats});
    }
    // a12: heads, last layers, sum_j f_ij e_ij
    Tensor n1 = f32({N, dh}, pos), n1p = f32({N, dh}, pos), n2 = f32({N, dh}, pos), n2p = f32({N, dh}, pos);
    const bool fused_head = W.head(16).numel() > 0 && n_out == 1;   // petb200_edge_head_fwd / _bwd
    Tensor e1p = f32({fused_head ? (E + 127) / 128 * 128 : E, dh}, pos), e2p = f32({E, dh}, pos);
    gemm(h, W.head(0), n1, &W.head(1), PETB200_EPI_SILU, nullptr, &n1p, prec);
    gemm(n1, W.head(2), n2, &W.head(3), PETB200_EPI_SILU, nullptr, &n2p, prec);
    Tensor atomic = f32({N, n_out}, pos), pe = f32({E, n_out}, pos);
    if (fused_head) {
      ok(petb200_edge_head_fwd(F(m), d, B(W.head(16)), F(W.head(5)), F(W.head(7)), F(W.head(14)), (float)fmeta[F_BE], E,
                               dh, Fm(e1p), Fm(e2p), Fm(pe), S()),
         "edge_head_fwd");
      ok(petb200_readout_fwd(F(n2), nullptr, F(W.head(12)), F(W.head(13)), F(W.head(14)), F(W.head(15)), F(fc),
                             I(row_ptr), N, E, dh, n_out, Fm(atomic), Fm(pe), S()),
         "readout_fwd");
    } else {
      Tensor e1 = f32({E, dh}, pos), e2 = f32({E, dh}, pos);
      gemm(m, W.head(4), e1, &W.head(5), PETB200_EPI_SILU, nullptr, &e1p, prec);
      gemm(e1, W.head(6), e2, &W.head(7), PETB200_EPI_SILU, nullptr, &e2p, prec);
      ok(petb200_readout_fwd(F(n2), F(e2), F(W.head(12)), F(W.head(13)), F(W.head(14)), F(W.head(15)), F(fc),
                             I(row_ptr), N, E, dh, n_out, Fm(atomic), Fm(pe), S()),
         "readout_fwd");
    }

    std::vector<Tensor> to_save = {vec, dist, fc, n1p, n2p, e1p, e2p, pe};
    to_save.insert(to_save.end(), keep.begin(), keep.end());
    to_save.insert(to_save.end(), topo.begin(), topo.end());
    to_save.insert(to_save.end(), weights.begin(), weights.end());
    ctx->save_for_backward(to_save);
    ctx->saved_data["meta"] = meta;
    ctx->saved_data["fmeta"] = fmeta;
    ctx->saved_data["n_structures"] = cells.size(0);
    ctx->saved_data["pos_dtype"] = (int64_t)positions.scalar_type();
    ctx->saved_data["cell_dtype"] = (int64_t)cells.scalar_type();
    ctx->saved_data["need_cells"] = cells.requires_grad();
    return atomic;
  }

  static variable_list backward(AutogradContext* ctx, variable_list grads) {
    const auto meta = ctx->saved_data["meta"].toIntVector();
    const auto fmeta = ctx->saved_data["fmeta"].toDoubleVector();
    const int n_gnn = (int)meta[M_NGNN], n_tl = (int)meta[M_NTL], d = (int)meta[M_D], dn = (int)meta[M_DN];
    const int nh = (int)meta[M_NH], dh = (int)meta[M_DH], n_out = (int)meta[M_NOUT], prec = (int)meta[M_PREC];
    const auto sv = ctx->get_saved_variables();
    const Tensor &vec = sv[0], &dist = sv[1], &fc = sv[2], &n1p = sv[3], &n2p = sv[4], &e1p = sv[5], &e2p = sv[6],
                 &pe = sv[7];
    const int keep0 = 8, topo0 = keep0 + 4 * n_gnn, w0 = topo0 + 9;
    const Tensor &row_ptr = sv[topo0], &ctr = sv[topo0 + 1], &rev = sv[topo0 + 3], &shift = sv[topo0 + 4];
    const Tensor& sys = sv[topo0 + 6];
    const std::vector<Tensor> weights(sv.begin() + w0, sv.end());
    const Weights W{weights, n_gnn, n_tl};
    c10::cuda::CUDAGuard guard(vec.device());
    const int64_t N = row_ptr.size(0) - 1, E = ctr.size(0);
    const int max_row = (int)sv[topo0 + 8].item<int64_t>();
    Tensor d_atomic = grads[0].to(torch::kFloat32).contiguous();

    // readout and heads
    const bool fused_head = W.head(16).numel() > 0 && n_out == 1;
    Tensor d_n2p = f32({N, dh}, vec), d_fc = torch::zeros({E}, vec.options());
    Tensor d_n1p = f32({N, dh}, vec), d_h = f32({N, dn}, vec), d_m = f32({E, d}, vec);
    if (fused_head) {
      ok(petb200_readout_bwd(F(d_atomic), nullptr, F(W.head(12)), F(W.head(14)), F(fc), I(ctr), F(n2p), nullptr, N, 0, dh,
                             n_out, Fm(d_n2p), nullptr, nullptr, S()),
         "readout_bwd");
      gemm(d_n2p, W.head(9), d_n1p, nullptr, PETB200_EPI_MUL_DSILU, &n1p, nullptr, prec);
      gemm(d_n1p, W.head(8), d_h, nullptr, PETB200_EPI_NONE, nullptr, nullptr, prec);
      ok(petb200_edge_head_bwd(F(d_atomic), I(ctr), F(fc), F(e1p), F(e2p), F(pe), B(W.head(17)), F(W.head(14)), E, dh,
                               Fm(d_m), d, Fm(d_fc), S()),
         "edge_head_bwd");
    } else {
      Tensor d_e2p = f32({E, dh}, vec), d_e1p = f32({E, dh}, vec);
      ok(petb200_readout_bwd(F(d_atomic), F(pe), F(W.head(12)), F(W.head(14)), F(fc), I(ctr), F(n2p), F(e2p), N, E, dh,
                             n_out, Fm(d_n2p), Fm(d_e2p), Fm(d_fc), S()),
         "readout_bwd");
      gemm(d_n2p, W.head(9), d_n1p, nullptr, PETB200_EPI_MUL_DSILU, &n1p, nullptr, prec);
      gemm(d_n1p, W.head(8), d_h, nullptr, PETB200_EPI_NONE, nullptr, nullptr, prec);
      gemm(d_e2p, W.head(11), d_e1p, nullptr, PETB200_EPI_MUL_DSILU, &e1p, nullptr, prec);
      gemm(d_e1p, W.head(10), d_m, nullptr, PETB200_EPI_NONE, nullptr, nullptr, prec);
    }

    // GNN layers in reverse
    Tensor d_vec = torch::zeros({E, 3}, vec.options()), d_dist = torch::zeros({E}, vec.options());
    petb200_dims dims{N, E, 0, d, dn, nh, max_row, prec,
                      (float)(1.0 / (std::sqrt((double)(d / nh)) * fmeta[F_TEMP]))};
    for (int l = n_gnn - 1; l >= 0; --l) {
      const Tensor &saved = sv[keep0 + 4 * l], &x = sv[keep0 + 4 * l + 1], &p1 = sv[keep0 + 4 * l + 2],
                   &cstats = sv[keep0 + 4 * l + 3];
      const int c0 = W.comb0(l);
      Tensor d_cat = f32({E, 2 * d}, vec), d_t = f32({E, d}, vec);
      ok(petb200_combine_bwd(F(d_m), d, F(p1), F(x), d, I(rev), F(cstats), B(weights[c0 + 1]), F(weights[c0 + 2]),
                             F(weights[c0 + 3]), E, d, Fm(d_cat), S()),
         "combine_bwd");
      ok(petb200_combine_scatter_bwd(F(d_cat), F(d_m), I(rev), E, d, Fm(d_t), S()), "combine_scatter_bwd");
      petb200_gnn_weights g{};
      std::vector<petb200_tl_weights> tls;
      W.fill(l, (int)meta[M_DFF], g, tls);
      Tensor scratch = bytes(petb200_gnn_scratch_bytes(&g, &dims), vec);
      Tensor d_h_in = l > 0 ? f32({N, dn}, vec) : Tensor();
      if (E > 0 || N > 0)
        ok(petb200_gnn_bwd(&g, &dims, I(row_ptr), F(fc), saved.data_ptr(), F(d_h), F(d_t), l > 0 ? Fm(d_m) : nullptr, d,
                           Fm(d_vec), Fm(d_dist), Fm(d_fc), l > 0 ? Fm(d_h_in) : nullptr, scratch.data_ptr(),
                           scratch.numel(), S()),
           "gnn_bwd");
      if (l > 0) d_h = d_h_in;
    }
    // geometry + force scatter
    const bool need_cells = ctx->saved_data["need_cells"].toBool();
    const int64_t n_struct = ctx->saved_data["n_structures"].toInt();
    Tensor scratch = f32({E > 0 ? E : 1, 3}, vec), d_pos = f32({N, 3}, vec);
    Tensor d_cells = need_cells ? torch::zeros({n_struct, 3, 3}, vec.options()) : Tensor();
    ok(petb200_edges_bwd(F(d_vec), F(d_dist), F(d_fc), F(vec), F(dist), I(row_ptr), I(ctr), I(rev), I(shift), I(sys), N, E,
                         (float)fmeta[F_CUTOFF], (float)fmeta[F_WIDTH], (int)meta[M_CUTFN], Fm(scratch), Fm(d_pos),
                         need_cells ? Fm(d_cells) : nullptr, S()),
       "edges_bwd");
    d_pos = d_pos.to((c10::ScalarType)ctx->saved_data["pos_dtype"].toInt());
    if (need_cells) d_cells = d_cells.to((c10::ScalarType)ctx->saved_data["cell_dtype"].toInt());
    return {d_pos, d_cells, Tensor(), Tensor(), Tensor(), Tensor()};
  }
};

Tensor pet_atomic(const Tensor& positions, const Tensor& cells, std::vector<Tensor> topo, std::vector<Tensor> weights,
                  std::vector<int64_t> meta, std::vector<double> fmeta) {
  return PetAtomic::apply(positions, cells, std::move(topo), std::move(weights), std::move(meta), std::move(fmeta));
}

}  // namespace

TORCH_LIBRARY(petb200, m) {
  m.def("topology(Tensor positions, Tensor cells, Tensor centers, Tensor neighbors, Tensor cell_shifts, "
        "Tensor system_indices, Tensor z_nodes, float cutoff) -> Tensor[]");
  m.def("pet_atomic(Tensor positions, Tensor cells, Tensor[] topology, Tensor[] weights, int[] meta, "
        "float[] fmeta) -> Tensor");
}

TORCH_LIBRARY_IMPL(petb200, CUDA, m) { m.impl("topology", topology); }
TORCH_LIBRARY_IMPL(petb200, Autograd, m) { m.impl("pet_atomic", pet_atomic); }
