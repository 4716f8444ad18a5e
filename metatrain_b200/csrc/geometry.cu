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
  cutoff_eval(sqrtf(r2) + 1e-15f, cutoff, width, func, f, df);  // structures.py:221
  fc_out[e] = f;
}

__global__ void edge_grad_kernel(const float* __restrict__ d_vec, const float* __restrict__ d_dist,
                                 const float* __restrict__ d_fc, const float* __restrict__ vec,
                                 const float* __restrict__ dist, int64_t n_edges, float cutoff,
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
  if (d_fc) {
    float f, df;
    cutoff_eval(nrm + 1e-15f, cutoff, width, func, f, df);
    if (nrm > 0.f) coef += d_fc[e] * df / nrm;  // d|r|/dr = r/|r| (0 at r = 0, as torch.norm)
  }
  G[3 * e + 0] = gx + coef * rx;
  G[3 * e + 1] = gy + coef * ry;
  G[3 * e + 2] = gz + coef * rz;
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
      positions, cells, system_of_atom, ctr, col, shift_csr, n_edges, cutoff, width,
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
        d_vec, d_dist, d_fc, edge_vec, edge_dist, n_edges, cutoff, width, cutoff_function,
        edge_grad);
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
      d_vec, d_dist, d_fc, edge_vec, edge_dist, n_edges, cutoff, width, cutoff_function, edge_grad);
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
