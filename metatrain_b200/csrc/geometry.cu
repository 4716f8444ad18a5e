// Edge geometry: displacement, distance, cutoff factor; backward = force scatter.
//
// Forward restates src/metatrain/pet/modules/structures.py:212-221,306-316,330 and
// utilities.py:4-39.  Backward restates what torch.autograd does to those lines when
// src/metatrain/utils/output_gradient.py:34-40 asks for dE/dpositions: the two
// index_select backward scatters of structures.py:220 become ONE segmented reduction over
// CSR rows (the edges pointing *to* atom i are exactly rev(e) for e in row i), so there are
// no atomics and the result is bit-reproducible.
#include "common.cuh"

namespace petb200 {
namespace {

constexpr float kPi = 3.14159265358979323846f;

// cutoff factor and its derivative w.r.t. the distance
__device__ __forceinline__ void cutoff_eval(float dist, float cutoff, float width, int func,
                                            float& f, float& df) {
  float s = (dist - (cutoff - width)) / width;
  if (func == PETB200_CUTOFF_BUMP) {
    // utilities.py:19-22: clamp(s, 1e-6, 1-1e-6); 0.5*(1+tanh(1/tan(pi*s)))
    const float lo = 1e-6f, hi = 1.0f - 1e-6f;
    bool inside = (s >= lo) && (s <= hi);  // torch.clamp passes gradient on the closed range
    float sc = fminf(fmaxf(s, lo), hi);
    float sn, cs;
    sincosf(kPi * sc, &sn, &cs);
    float u = cs / sn;  // cot(pi s)
    float th = tanhf(u);
    f = 0.5f * (1.0f + th);
    // d/ds = 0.5 * sech^2(u) * (-pi / sin^2)
    df = inside ? (0.5f * (1.0f - th * th) * (-kPi / (sn * sn)) / width) : 0.f;
  } else {
    // utilities.py:37-39: 0.5*(1+cos(pi*clamp(s,0,1)))
    bool inside = (s >= 0.f) && (s <= 1.f);
    float sc = fminf(fmaxf(s, 0.f), 1.f);
    f = 0.5f * (1.0f + cosf(kPi * sc));
    df = inside ? (-0.5f * kPi * sinf(kPi * sc) / width) : 0.f;
  }
}

__global__ void edges_fwd_kernel(const float* __restrict__ pos, const float* __restrict__ cells,
                                 const int32_t* __restrict__ sys_of_atom,
                                 const int32_t* __restrict__ ctr, const int32_t* __restrict__ col,
                                 const int32_t* __restrict__ shift, int64_t n_edges, float cutoff,
                                 const float* __restrict__ edge_cutoff /* nullable: per-edge cutoff */,
                                 float width, int func, float* __restrict__ vec,
                                 float* __restrict__ dist_out, float* __restrict__ fc_out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  int i = ctr[e], j = col[e];
  const float* c = cells + (int64_t)sys_of_atom[i] * 9;
  float sa = (float)shift[3 * e], sb = (float)shift[3 * e + 1], sc = (float)shift[3 * e + 2];
  float rx = pos[3 * (int64_t)j + 0] - pos[3 * (int64_t)i + 0] + (sa * c[0] + sb * c[3] + sc * c[6]);
  float ry = pos[3 * (int64_t)j + 1] - pos[3 * (int64_t)i + 1] + (sa * c[1] + sb * c[4] + sc * c[7]);
  float rz = pos[3 * (int64_t)j + 2] - pos[3 * (int64_t)i + 2] + (sa * c[2] + sb * c[5] + sc * c[8]);
  float r2 = rx * rx + ry * ry + rz * rz;
  vec[3 * e + 0] = rx;
  vec[3 * e + 1] = ry;
  vec[3 * e + 2] = rz;
  dist_out[e] = sqrtf(r2 + 1e-15f);  // structures.py:330 (embedder input)
  float f, df;
  cutoff_eval(sqrtf(r2) + 1e-15f, edge_cutoff ? edge_cutoff[e] : cutoff, width, func, f, df);  // structures.py:221
  fc_out[e] = f;
}

__global__ void edge_grad_kernel(const float* __restrict__ d_vec, const float* __restrict__ d_dist,
                                 const float* __restrict__ d_fc, const float* __restrict__ vec,
                                 const float* __restrict__ dist, int64_t n_edges, float cutoff,
                                 const float* __restrict__ edge_cutoff /* nullable */,
                                 float* __restrict__ d_edge_cutoff /* nullable: -d_fc * f'(d) */,
                                 float width, int func, float* __restrict__ G) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  float rx = vec[3 * e], ry = vec[3 * e + 1], rz = vec[3 * e + 2];
  float r2 = rx * rx + ry * ry + rz * rz;
  float nrm = sqrtf(r2);
  float gx = d_vec ? d_vec[3 * e] : 0.f, gy = d_vec ? d_vec[3 * e + 1] : 0.f,
        gz = d_vec ? d_vec[3 * e + 2] : 0.f;
  float coef = 0.f;
  if (d_dist) coef += d_dist[e] / dist[e];  // d sqrt(r.r+eps) / dr = r / sqrt(r.r+eps)
  if (d_edge_cutoff) d_edge_cutoff[e] = 0.f;
  if (d_fc) {
    float f, df;
    cutoff_eval(nrm + 1e-15f, edge_cutoff ? edge_cutoff[e] : cutoff, width, func, f, df);
    if (nrm > 0.f) coef += d_fc[e] * df / nrm;  // d|r|/dr = r/|r| (0 at r = 0, as torch.norm)
    // the cutoff functions depend on (d - cutoff) only: df/dcutoff = -df/dd
    if (d_edge_cutoff) d_edge_cutoff[e] = -d_fc[e] * df;
  }
  G[3 * e + 0] = gx + coef * rx;
  G[3 * e + 1] = gy + coef * ry;
  G[3 * e + 2] = gz + coef * rz;
}

// ----------------------------------------------------------------- adaptive cutoff
// Restates get_adaptive_cutoffs_solver (src/metatrain/pet/modules/adaptive_cutoff.py:110-229) and
// its use in compute_batch_tensors (structures.py:222-262) on the CSR rows of the pairs within
// the maximum cutoff R: per atom, solve  n_total(r) = sum_j bump(d_j; r, w) + n* (r/R)^3 = n*
// with a bracketed Newton iteration, then one implicit-function step that carries the gradient.

// smoothed step of a neighbour at distance d for probe cutoff r, and its derivative w.r.t. r
// (adaptive_cutoff.py:74-95)
__device__ __forceinline__ void probe_count(float d, float r, float w, float& f, float& df_dr) {
  const float s = (d - (r - w)) / w;
  const bool active = s > 0.f && s < 1.f;
  const float safe = fminf(fmaxf(s, 1e-6f), 1.0f - 1e-6f);
  float sn, cs;
  sincosf(kPi * safe, &sn, &cs);
  const float th = tanhf(cs / sn);
  f = active ? 0.5f * (1.0f + th) : (s <= 0.f ? 1.f : 0.f);
  df_dr = active ? (0.5f * kPi / w) * (1.0f - th * th) / (sn * sn) : 0.f;
}

// one warp per atom
__global__ void adaptive_solve_kernel(const int32_t* __restrict__ row_ptr, const float* __restrict__ dist,
                                      int64_t n_atoms, float n_target, float r_max, float width,
                                      float* __restrict__ r_root, float* __restrict__ dn_root,
                                      float* __restrict__ r_atom, float* __restrict__ pass) {
  const int64_t atom = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (atom >= n_atoms) return;
  const int lo = row_ptr[atom], hi = row_ptr[atom + 1];
  const float inv_max = 1.0f / r_max;
  auto totals = [&](float r, float& n, float& dn) {
    float a = 0.f, b = 0.f;
    for (int e = lo + lane; e < hi; e += 32) {
      float f, df;
      probe_count(dist[e], r, width, f, df);
      a += f;
      b += df;
    }
    const float x = r * inv_max;
    n = warp_sum(a) + n_target * x * x * x;
    dn = warp_sum(b) + 3.0f * n_target * x * x * inv_max;
  };
  float r_lo = 0.f, r_hi = r_max, r = 0.5f * r_max, n, dn;
  for (int it = 0; it < 10; ++it) {  // adaptive_cutoff.py:171-188
    totals(r, n, dn);
    const float f = n - n_target;
    if (f <= 0.f) r_lo = r; else r_hi = r;
    const float r_newton = r - f / fmaxf(dn, 1e-6f);
    r = (r_newton >= r_lo && r_newton <= r_hi) ? r_newton : 0.5f * (r_lo + r_hi);
  }
  totals(r, n, dn);
  // implicit-function step (:203-227): r and dn are constants for the gradient
  const float raw = r - (n - n_target) / fmaxf(dn, 1e-6f);
  const float lo_c = r_max * (1.0f / 16.0f);
  if (lane == 0) {
    r_root[atom] = r;
    dn_root[atom] = fmaxf(dn, 1e-6f);
    r_atom[atom] = fminf(fmaxf(raw, lo_c), r_max);
    pass[atom] = (raw >= lo_c && raw <= r_max) ? 1.f : 0.f;  // torch.clamp passes gradient on the closed range
  }
}

// pair cutoffs (mean of the two atoms', structures.py:253-255) and the keep mask (:256-262)
__global__ void adaptive_pair_kernel(const int32_t* __restrict__ ctr, const int32_t* __restrict__ col,
                                     const float* __restrict__ vec, const float* __restrict__ r_atom,
                                     int64_t n_edges, float* __restrict__ pair_cutoff,
                                     int32_t* __restrict__ keep, int32_t* __restrict__ counts) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const float rc = 0.5f * (r_atom[ctr[e]] + r_atom[col[e]]);
  const float x = vec[3 * e], y = vec[3 * e + 1], z = vec[3 * e + 2];
  const int k = sqrtf(x * x + y * y + z * z) + 1e-15f <= rc;
  pair_cutoff[e] = rc;
  keep[e] = k;
  if (k) atomicAdd(&counts[ctr[e]], 1);
}

// backward, step 1 (one warp per atom, rows of the masked topology): the gradient of the atom's
// cutoff is half the sum over its edges and their reverses (pair cutoffs are symmetric means);
// coef[i] = pass * d_r_atom / dn_root is what every pre-mask neighbour distance of i receives
__global__ void adaptive_atom_grad_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ rev,
                                          const float* __restrict__ d_pair_cutoff,
                                          const float* __restrict__ dn_root, const float* __restrict__ pass,
                                          int64_t n_atoms, float* __restrict__ coef) {
  const int64_t atom = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (atom >= n_atoms) return;
  float acc = 0.f;
  for (int e = row_ptr[atom] + lane; e < row_ptr[atom + 1]; e += 32)
    acc += d_pair_cutoff[e] + d_pair_cutoff[rev[e]];
  acc = warp_sum(acc);
  if (lane == 0) coef[atom] = pass[atom] * 0.5f * acc / dn_root[atom];
}

// backward, step 2 (pre-mask edges): d r_atom / d d_k = (df/dr)(d_k; r_root) / dn_root
__global__ void adaptive_dist_grad_kernel(const int32_t* __restrict__ ctr, const float* __restrict__ dist,
                                          const float* __restrict__ r_root, const float* __restrict__ coef,
                                          int64_t n_edges, float width, float* __restrict__ d_dist) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int i = ctr[e];
  float f, df;
  probe_count(dist[e], r_root[i], width, f, df);
  d_dist[e] = coef[i] * df;
}

// ---- the legacy "grid" method (adaptive_cutoff.py:232-395): smoothed neighbour counts n_p at P
// probe cutoffs c_p = min_cutoff + p * spacing, D_p = n_p - n* + n* x_p^3 (x_p = p / (P - 1)),
// widths W_p = max(|torch.gradient(D)_p|, 1e-12), weights e_p = exp(-(D_p / W_p)^2 / 2) normalised
// over p, cutoff = sum_p c_p e_p.  One warp per atom, lane p owns probe p (P <= 32).  Besides the
// cutoff the kernel stores G[i][q] = d cutoff_i / d D_q, which is all the backward needs.
// (The reference subtracts the batch-wide max of the log-weights before exp(); it cancels in
// the normalisation — the per-atom max is used here.)
constexpr int kMaxProbes = 32;

__global__ void adaptive_grid_kernel(const int32_t* __restrict__ row_ptr, const float* __restrict__ dist,
                                     int64_t n_atoms, float n_target, float width, float min_cutoff,
                                     float spacing, int n_probes, float* __restrict__ r_atom,
                                     float* __restrict__ grad_d /* [n_atoms, n_probes] */) {
  const int64_t atom = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (atom >= n_atoms) return;
  const int lo = row_ptr[atom], hi = row_ptr[atom + 1];
  const int P = n_probes;
  // n_p for every probe: lanes stride over the row, then P warp reductions; lane p keeps n_p
  float mine = 0.f;
  for (int p = 0; p < P; ++p) {
    const float c = min_cutoff + p * spacing;
    float acc = 0.f;
    for (int e = lo + lane; e < hi; e += 32) {
      float f, df;
      probe_count(dist[e], c, width, f, df);
      acc += f;
    }
    acc = warp_sum(acc);
    if (lane == p) mine = acc;
  }
  const bool on = lane < P;
  const float c_p = min_cutoff + lane * spacing;
  if (P == 1) {  // a single probe: the weighted mean is that probe, whatever the counts
    if (lane == 0) {
      r_atom[atom] = c_p;
      grad_d[atom] = 0.f;
    }
    return;
  }
  const float x = (float)lane / (float)(P - 1);
  const float D = on ? mine - n_target + n_target * x * x * x : 0.f;
  // torch.gradient, unit spacing: centred differences inside, one-sided at the two ends
  const float D_up = __shfl_down_sync(0xffffffffu, D, 1), D_dn = __shfl_up_sync(0xffffffffu, D, 1);
  float g;
  if (lane == 0) g = D_up - D;
  else if (lane == P - 1) g = D - D_dn;
  else g = 0.5f * (D_up - D_dn);
  const float W = fmaxf(fabsf(g), 1e-12f);
  const float L = on ? -0.5f * (D / W) * (D / W) : -INFINITY;
  float Lmax = L;
  for (int o = 16; o > 0; o >>= 1) Lmax = fmaxf(Lmax, __shfl_xor_sync(0xffffffffu, Lmax, o));
  const float e = on ? expf(L - Lmax) : 0.f;
  const float S = warp_sum(e);
  const float rbar = warp_sum(c_p * e) / S;
  // backward coefficients
  const float A = on ? e * (c_p - rbar) / S : 0.f;                        // d rbar / d L_p
  const float dL_dD = -D / (W * W);
  const float sgn = g > 0.f ? 1.f : (g < 0.f ? -1.f : 0.f);
  const float B = (on && fabsf(g) > 1e-12f) ? A * (D * D / (W * W * W)) * sgn : 0.f;  // d rbar / d g_p
  const float B_dn = __shfl_up_sync(0xffffffffu, B, 1);    // B_{q-1}
  const float B_up = __shfl_down_sync(0xffffffffu, B, 1);  // B_{q+1}
  float G = A * dL_dD;
  // g_{q-1} depends on D_q with +1/2 (interior q-1) or +1 (q-1 = 0)
  if (lane >= 1) G += (lane - 1 == 0 ? 1.f : 0.5f) * B_dn;
  // g_{q+1} depends on D_q with -1/2 (interior q+1) or -1 (q+1 = P-1)
  if (lane + 1 <= P - 1) G -= (lane + 1 == P - 1 ? 1.f : 0.5f) * B_up;
  if (lane == 0) G -= B;
  if (lane == P - 1) G += B;
  if (on) grad_d[atom * P + lane] = G;
  if (lane == 0) r_atom[atom] = rbar;
}

// backward through the grid method: d cutoff_i / d d_k = sum_q G[i][q] * d bump(d_k; c_q) / d d_k
__global__ void adaptive_grid_dist_grad_kernel(const int32_t* __restrict__ ctr, const float* __restrict__ dist,
                                               const float* __restrict__ grad_d, const float* __restrict__ coef,
                                               int64_t n_edges, float width, float min_cutoff, float spacing,
                                               int n_probes, float* __restrict__ d_dist) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int i = ctr[e];
  const float d = dist[e];
  float acc = 0.f;
  for (int q = 0; q < n_probes; ++q) {
    float f, df_dr;
    probe_count(d, min_cutoff + q * spacing, width, f, df_dr);
    acc -= grad_d[(int64_t)i * n_probes + q] * df_dr;  // d bump / d d = - d bump / d r
  }
  d_dist[e] = coef[i] * acc;
}

// one warp per atom: d_pos[i] = sum_{e in row i} (G[rev e] - G[e])
__global__ void force_scatter_kernel(const float* __restrict__ G, const int32_t* __restrict__ row_ptr,
                                     const int32_t* __restrict__ rev, int64_t n_atoms,
                                     float* __restrict__ d_pos) {
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n_atoms) return;
  int lo = row_ptr[warp], hi = row_ptr[warp + 1];
  float ax = 0.f, ay = 0.f, az = 0.f;
  for (int e = lo + lane; e < hi; e += 32) {
    int64_t r = rev[e];
    ax += G[3 * r + 0] - G[3 * (int64_t)e + 0];
    ay += G[3 * r + 1] - G[3 * (int64_t)e + 1];
    az += G[3 * r + 2] - G[3 * (int64_t)e + 2];
  }
  ax = warp_sum(ax);
  ay = warp_sum(ay);
  az = warp_sum(az);
  if (lane == 0) {
    d_pos[3 * warp + 0] = ax;
    d_pos[3 * warp + 1] = ay;
    d_pos[3 * warp + 2] = az;
  }
}

// d_cells[b][a][c] += sum_e S_e[a] * G_e[c]   (strain / virial path, evaluate_model.py:310-321)
__global__ void cell_grad_kernel(const float* __restrict__ G, const int32_t* __restrict__ shift,
                                 const int32_t* __restrict__ ctr,
                                 const int32_t* __restrict__ sys_of_atom, int64_t n_edges,
                                 float* __restrict__ d_cells) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float v[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) v[k] = 0.f;
  int sys = -1;
  if (e < n_edges) {
    int sa = shift[3 * e], sb = shift[3 * e + 1], sc = shift[3 * e + 2];
    if (sa | sb | sc) {
      sys = sys_of_atom[ctr[e]];
      float gx = G[3 * e], gy = G[3 * e + 1], gz = G[3 * e + 2];
      v[0] = sa * gx; v[1] = sa * gy; v[2] = sa * gz;
      v[3] = sb * gx; v[4] = sb * gy; v[5] = sb * gz;
      v[6] = sc * gx; v[7] = sc * gy; v[8] = sc * gz;
    }
  }
  // warp-aggregate when every contributing lane belongs to one structure (the common case)
  unsigned contributing = __ballot_sync(0xffffffffu, sys >= 0);
  if (contributing == 0u) return;  // warp-uniform
  int target = __shfl_sync(0xffffffffu, sys, __ffs(contributing) - 1);
  unsigned agree = __ballot_sync(0xffffffffu, sys < 0 || sys == target);
  if (agree == 0xffffffffu) {
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = warp_sum(v[k]);
    if ((threadIdx.x & 31) == 0)
      for (int k = 0; k < 9; ++k) atomicAdd(&d_cells[(int64_t)target * 9 + k], v[k]);
  } else if (sys >= 0) {
    for (int k = 0; k < 9; ++k) atomicAdd(&d_cells[(int64_t)sys * 9 + k], v[k]);
  }
}

}  // namespace
}  // namespace petb200

using namespace petb200;

extern "C" PETB200_API int petb200_edges_fwd(const float* positions, const float* cells,
                                 const int32_t* system_of_atom, const int32_t* ctr,
                                 const int32_t* col, const int32_t* shift_csr, int64_t n_edges,
                                 float cutoff, float width, int cutoff_function, float* edge_vec,
                                 float* edge_dist, float* cutoff_factor, cudaStream_t stream) {
  PETB200_REQUIRE(cutoff_function == PETB200_CUTOFF_BUMP || cutoff_function == PETB200_CUTOFF_COSINE,
                  "edges_fwd: unknown cutoff function %d", cutoff_function);
  if (n_edges == 0) return PETB200_OK;
  edges_fwd_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(
      positions, cells, system_of_atom, ctr, col, shift_csr, n_edges, cutoff, nullptr, width,
      cutoff_function, edge_vec, edge_dist, cutoff_factor);
  return check_launch("edges_fwd");
}

extern "C" PETB200_API int petb200_edges_bwd(const float* d_vec, const float* d_dist, const float* d_fc,
                                 const float* edge_vec, const float* edge_dist,
                                 const int32_t* row_ptr, const int32_t* ctr, const int32_t* rev,
                                 const int32_t* shift_csr, const int32_t* system_of_atom,
                                 int64_t n_atoms, int64_t n_edges, float cutoff, float width,
                                 int cutoff_function, float* edge_grad, float* d_pos,
                                 float* d_cells, cudaStream_t stream) {
  PETB200_REQUIRE(cutoff_function == PETB200_CUTOFF_BUMP || cutoff_function == PETB200_CUTOFF_COSINE,
                  "edges_bwd: unknown cutoff function %d", cutoff_function);
  if (n_edges > 0) {
    edge_grad_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(
        d_vec, d_dist, d_fc, edge_vec, edge_dist, n_edges, cutoff, nullptr, nullptr, width,
        cutoff_function, edge_grad);
  }
  if (n_atoms > 0 && d_pos) {
    force_scatter_kernel<<<(unsigned)ceil_div(n_atoms * 32, 256), 256, 0, stream>>>(
        edge_grad, row_ptr, rev, n_atoms, d_pos);
  }
  if (n_edges > 0 && d_cells) {
    cell_grad_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(
        edge_grad, shift_csr, ctr, system_of_atom, n_edges, d_cells);
  }
  return check_launch("edges_bwd");
}

// --- split form of petb200_edges_bwd for the atom-sharded path: the edge gradients of halo
// edges are exchanged between the two calls (metatrain_b200/sharded.py).
extern "C" PETB200_API int petb200_edge_grad(const float* d_vec, const float* d_dist,
                                             const float* d_fc, const float* edge_vec,
                                             const float* edge_dist, int64_t n_edges, float cutoff,
                                             float width, int cutoff_function, float* edge_grad,
                                             cudaStream_t stream) {
  PETB200_REQUIRE(cutoff_function == PETB200_CUTOFF_BUMP || cutoff_function == PETB200_CUTOFF_COSINE,
                  "edge_grad: unknown cutoff function %d", cutoff_function);
  if (n_edges == 0) return PETB200_OK;
  edge_grad_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(
      d_vec, d_dist, d_fc, edge_vec, edge_dist, n_edges, cutoff, nullptr, nullptr, width, cutoff_function,
      edge_grad);
  return check_launch("edge_grad");
}

extern "C" PETB200_API int petb200_force_scatter(const float* edge_grad, const int32_t* row_ptr,
                                                 const int32_t* ctr, const int32_t* rev,
                                                 const int32_t* shift_csr,
                                                 const int32_t* system_of_atom, int64_t n_atoms,
                                                 int64_t n_edges, float* d_pos, float* d_cells,
                                                 cudaStream_t stream) {
  if (n_atoms > 0 && d_pos) {
    force_scatter_kernel<<<(unsigned)ceil_div(n_atoms * 32, 256), 256, 0, stream>>>(
        edge_grad, row_ptr, rev, n_atoms, d_pos);
  }
  if (n_edges > 0 && d_cells) {
    cell_grad_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(
        edge_grad, shift_csr, ctr, system_of_atom, n_edges, d_cells);
  }
  return check_launch("force_scatter");
}

// ---- per-edge cutoffs (adaptive cutoff): the same kernels with a cutoff array
extern "C" PETB200_API int petb200_edges_fwd_rc(const float* positions, const float* cells,
                                    const int32_t* system_of_atom, const int32_t* ctr,
                                    const int32_t* col, const int32_t* shift_csr, int64_t n_edges,
                                    const float* edge_cutoff, float width, int cutoff_function,
                                    float* edge_vec, float* edge_dist, float* cutoff_factor,
                                    cudaStream_t stream) {
  PETB200_REQUIRE(cutoff_function == PETB200_CUTOFF_BUMP || cutoff_function == PETB200_CUTOFF_COSINE,
                  "edges_fwd_rc: unknown cutoff function %d", cutoff_function);
  if (n_edges == 0) return PETB200_OK;
  PETB200_REQUIRE(edge_cutoff != nullptr, "edges_fwd_rc: edge_cutoff is null");
  edges_fwd_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(
      positions, cells, system_of_atom, ctr, col, shift_csr, n_edges, 0.f, edge_cutoff, width,
      cutoff_function, edge_vec, edge_dist, cutoff_factor);
  return check_launch("edges_fwd_rc");
}

extern "C" PETB200_API int petb200_edges_bwd_rc(const float* d_vec, const float* d_dist, const float* d_fc,
                                    const float* edge_vec, const float* edge_dist,
                                    const int32_t* row_ptr, const int32_t* ctr, const int32_t* rev,
                                    const int32_t* shift_csr, const int32_t* system_of_atom,
                                    int64_t n_atoms, int64_t n_edges, const float* edge_cutoff,
                                    float width, int cutoff_function, float* edge_grad, float* d_pos,
                                    float* d_cells, float* d_edge_cutoff, cudaStream_t stream) {
  PETB200_REQUIRE(cutoff_function == PETB200_CUTOFF_BUMP || cutoff_function == PETB200_CUTOFF_COSINE,
                  "edges_bwd_rc: unknown cutoff function %d", cutoff_function);
  PETB200_REQUIRE(n_edges == 0 || edge_cutoff != nullptr, "edges_bwd_rc: edge_cutoff is null");
  if (n_edges > 0) {
    edge_grad_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(
        d_vec, d_dist, d_fc, edge_vec, edge_dist, n_edges, 0.f, edge_cutoff, d_edge_cutoff, width,
        cutoff_function, edge_grad);
  }
  if (n_atoms > 0 && d_pos) {
    force_scatter_kernel<<<(unsigned)ceil_div(n_atoms * 32, 256), 256, 0, stream>>>(
        edge_grad, row_ptr, rev, n_atoms, d_pos);
  }
  if (n_edges > 0 && d_cells) {
    cell_grad_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(
        edge_grad, shift_csr, ctr, system_of_atom, n_edges, d_cells);
  }
  return check_launch("edges_bwd_rc");
}

extern "C" PETB200_API int petb200_adaptive_cutoff_solve(const int32_t* row_ptr, const float* edge_dist,
                                             int64_t n_atoms, float num_neighbors, float max_cutoff,
                                             float width, float* r_root, float* dn_root,
                                             float* atomic_cutoff, float* pass, cudaStream_t stream) {
  PETB200_REQUIRE(max_cutoff > 0.f && width > 0.f && num_neighbors > 0.f,
                  "adaptive_cutoff_solve: cutoff, width and neighbour target must be positive");
  if (n_atoms == 0) return PETB200_OK;
  adaptive_solve_kernel<<<(unsigned)ceil_div(n_atoms * 32, 256), 256, 0, stream>>>(
      row_ptr, edge_dist, n_atoms, num_neighbors, max_cutoff, width, r_root, dn_root, atomic_cutoff, pass);
  return check_launch("adaptive_cutoff_solve");
}

extern "C" PETB200_API int petb200_adaptive_pair_mask(const int32_t* ctr, const int32_t* col,
                                          const float* edge_vec, const float* atomic_cutoff,
                                          int64_t n_edges, float* pair_cutoff, int32_t* keep,
                                          int32_t* counts, cudaStream_t stream) {
  if (n_edges == 0) return PETB200_OK;
  adaptive_pair_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(
      ctr, col, edge_vec, atomic_cutoff, n_edges, pair_cutoff, keep, counts);
  return check_launch("adaptive_pair_mask");
}

extern "C" PETB200_API int petb200_adaptive_cutoff_bwd(const int32_t* row_ptr_kept, const int32_t* rev_kept,
                                           const float* d_pair_cutoff, const float* dn_root,
                                           const float* pass, int64_t n_atoms, const int32_t* ctr_all,
                                           const float* dist_all, const float* r_root,
                                           int64_t n_edges_all, float width, float* coef,
                                           float* d_dist_all, cudaStream_t stream) {
  if (n_atoms > 0) {
    adaptive_atom_grad_kernel<<<(unsigned)ceil_div(n_atoms * 32, 256), 256, 0, stream>>>(
        row_ptr_kept, rev_kept, d_pair_cutoff, dn_root, pass, n_atoms, coef);
  }
  if (n_edges_all > 0) {
    adaptive_dist_grad_kernel<<<(unsigned)ceil_div(n_edges_all, 256), 256, 0, stream>>>(
        ctr_all, dist_all, r_root, coef, n_edges_all, width, d_dist_all);
  }
  return check_launch("adaptive_cutoff_bwd");
}

extern "C" PETB200_API int petb200_adaptive_grid_solve(const int32_t* row_ptr, const float* edge_dist,
                                           int64_t n_atoms, float num_neighbors, float width,
                                           float min_cutoff, float spacing, int n_probes,
                                           float* atomic_cutoff, float* grad_d, cudaStream_t stream) {
  PETB200_REQUIRE(width > 0.f && spacing > 0.f && num_neighbors > 0.f,
                  "adaptive_grid_solve: width, spacing and neighbour target must be positive");
  if (n_probes < 1 || n_probes > kMaxProbes) {
    set_error("adaptive_grid_solve: %d probe cutoffs; 1..%d are built", n_probes, kMaxProbes);
    return PETB200_ERR_UNSUPPORTED;
  }
  if (n_atoms == 0) return PETB200_OK;
  adaptive_grid_kernel<<<(unsigned)ceil_div(n_atoms * 32, 256), 256, 0, stream>>>(
      row_ptr, edge_dist, n_atoms, num_neighbors, width, min_cutoff, spacing, n_probes, atomic_cutoff,
      grad_d);
  return check_launch("adaptive_grid_solve");
}

extern "C" PETB200_API int petb200_adaptive_grid_bwd(const int32_t* row_ptr_kept, const int32_t* rev_kept,
                                         const float* d_pair_cutoff, const float* ones,
                                         int64_t n_atoms, const int32_t* ctr_all, const float* dist_all,
                                         const float* grad_d, int64_t n_edges_all, float width,
                                         float min_cutoff, float spacing, int n_probes, float* coef,
                                         float* d_dist_all, cudaStream_t stream) {
  if (n_atoms > 0) {
    adaptive_atom_grad_kernel<<<(unsigned)ceil_div(n_atoms * 32, 256), 256, 0, stream>>>(
        row_ptr_kept, rev_kept, d_pair_cutoff, ones, ones, n_atoms, coef);
  }
  if (n_edges_all > 0) {
    adaptive_grid_dist_grad_kernel<<<(unsigned)ceil_div(n_edges_all, 256), 256, 0, stream>>>(
        ctr_all, dist_all, grad_d, coef, n_edges_all, width, min_cutoff, spacing, n_probes, d_dist_all);
  }
  return check_launch("adaptive_grid_bwd");
}
