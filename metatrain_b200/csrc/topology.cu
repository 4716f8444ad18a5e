// Edge topology: non-strict filter, CSR build, reverse-edge map, CSR<->NEF.
//
// Replaces the integer half of compute_batch_tensors
// (src/metatrain/pet/modules/structures.py:265-363) and nef.py:34-251.  The reference
// runs three argsorts over E plus ~40 small launches and a host sync; here it is one
// filter/count kernel, a CUB scan + stable radix sort (toolkit plumbing, not a hot op),
// one gather, and a row-scan reverse lookup that needs no sort at all because the rows
// are already grouped by centre.
#include <cub/cub.cuh>

#include "common.cuh"

namespace petb200 {
namespace {

__device__ __forceinline__ void edge_vector(const float* __restrict__ pos,
                                            const float* __restrict__ cells, int sys, int i,
                                            int j, int sa, int sb, int sc, float& rx, float& ry,
                                            float& rz) {
  const float* c = cells + (int64_t)sys * 9;
  // cell_shifts @ cell  (row vector times matrix), structures.py:212-219
  float ox = sa * c[0] + sb * c[3] + sc * c[6];
  float oy = sa * c[1] + sb * c[4] + sc * c[7];
  float oz = sa * c[2] + sb * c[5] + sc * c[8];
  rx = pos[3 * (int64_t)j + 0] - pos[3 * (int64_t)i + 0] + ox;
  ry = pos[3 * (int64_t)j + 1] - pos[3 * (int64_t)i + 1] + oy;
  rz = pos[3 * (int64_t)j + 2] - pos[3 * (int64_t)i + 2] + oz;
}

__global__ void nl_filter_count_kernel(const float* __restrict__ pos,
                                       const float* __restrict__ cells,
                                       const int32_t* __restrict__ sys_of_atom,
                                       const int32_t* __restrict__ centers,
                                       const int32_t* __restrict__ neighbors,
                                       const int32_t* __restrict__ shifts, int64_t n_pairs,
                                       float cutoff, int32_t* __restrict__ keep,
                                       int32_t* __restrict__ counts) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_pairs) return;
  int i = centers[e], j = neighbors[e];
  float rx, ry, rz;
  edge_vector(pos, cells, sys_of_atom[i], i, j, shifts[3 * e], shifts[3 * e + 1],
              shifts[3 * e + 2], rx, ry, rz);
  // torch.norm(edge_vectors) + 1e-15 <= cutoff   (structures.py:221,267)
  float dist = sqrtf(rx * rx + ry * ry + rz * rz) + 1e-15f;
  int k = dist <= cutoff ? 1 : 0;
  keep[e] = k;
  if (k) atomicAdd(&counts[i], 1);
}

__global__ void make_sort_keys_kernel(const int32_t* __restrict__ centers,
                                      const int32_t* __restrict__ keep, int64_t n_pairs,
                                      int32_t n_atoms, int32_t* __restrict__ keys,
                                      int32_t* __restrict__ vals) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_pairs) return;
  keys[e] = keep[e] ? centers[e] : n_atoms;  // dropped pairs sort to the end
  vals[e] = (int32_t)e;
}

__global__ void finish_row_ptr_kernel(const int32_t* __restrict__ counts, int64_t n_atoms,
                                      int32_t* __restrict__ row_ptr, int32_t* __restrict__ stats,
                                      const int32_t* __restrict__ max_count) {
  // row_ptr[0..N) already holds the exclusive scan; close it and publish the stats
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int32_t total = n_atoms > 0 ? row_ptr[n_atoms - 1] + counts[n_atoms - 1] : 0;
    row_ptr[n_atoms] = total;
    stats[0] = total;
    stats[1] = n_atoms > 0 ? *max_count : 0;
  }
}

__global__ void csr_gather_kernel(const int32_t* __restrict__ perm,
                                  const int32_t* __restrict__ centers,
                                  const int32_t* __restrict__ neighbors,
                                  const int32_t* __restrict__ shifts, int64_t n_edges,
                                  int32_t* __restrict__ ctr, int32_t* __restrict__ col,
                                  int32_t* __restrict__ shift_csr) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_edges) return;
  int64_t e = perm[k];
  ctr[k] = centers[e];
  col[k] = neighbors[e];
  shift_csr[3 * k + 0] = shifts[3 * e + 0];
  shift_csr[3 * k + 1] = shifts[3 * e + 1];
  shift_csr[3 * k + 2] = shifts[3 * e + 2];
}

__global__ void reverse_map_kernel(const int32_t* __restrict__ row_ptr,
                                   const int32_t* __restrict__ ctr,
                                   const int32_t* __restrict__ col,
                                   const int32_t* __restrict__ shift, int64_t n_edges,
                                   int64_t n_rows, int32_t* __restrict__ rev,
                                   int32_t* __restrict__ n_missing) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int i = ctr[e], j = col[e];
  const int sa = -shift[3 * e], sb = -shift[3 * e + 1], sc = -shift[3 * e + 2];
  int found = -1;
  // a neighbour without a CSR row is a ghost atom of an atom-sharded run: its edges (and the
  // reverse of this one) live on another rank
  const bool has_row = j < n_rows;
  const int lo = has_row ? row_ptr[j] : 0, hi = has_row ? row_ptr[j + 1] : 0;
  for (int k = lo; k < hi; ++k) {
    if (col[k] == i && shift[3 * (int64_t)k] == sa && shift[3 * (int64_t)k + 1] == sb &&
        shift[3 * (int64_t)k + 2] == sc) {
      found = k;
      break;
    }
  }
  if (found < 0) {
    atomicAdd(n_missing, 1);
    found = (int)e;  // keep indices in range; caller raises on n_missing != 0
  }
  rev[e] = found;
}

__global__ void csr_to_nef_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_ptr,
                                  int64_t n_atoms, int width_m, int d, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n_atoms * width_m * d;
  if (idx >= total) return;
  int c = (int)(idx % d);
  int64_t slot_flat = idx / d;
  int slot = (int)(slot_flat % width_m);
  int64_t i = slot_flat / width_m;
  int lo = row_ptr[i], n = row_ptr[i + 1] - lo;
  out[idx] = slot < n ? x[(int64_t)(lo + slot) * d + c] : 0.f;
}

__global__ void nef_to_csr_kernel(const float* __restrict__ x_nef,
                                  const int32_t* __restrict__ row_ptr,
                                  const int32_t* __restrict__ ctr, int64_t n_edges, int width_m,
                                  int d, float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_edges * d) return;
  int c = (int)(idx % d);
  int64_t e = idx / d;
  int i = ctr[e];
  int slot = (int)(e - row_ptr[i]);
  out[idx] = x_nef[((int64_t)i * width_m + slot) * d + c];
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct CsrWorkspace {
  size_t keys_in, vals_in, keys_out, max_count, cub, total;
};

CsrWorkspace csr_layout(int64_t n_pairs, int64_t n_atoms) {
  CsrWorkspace w;
  size_t off = 0;
  w.keys_in = off;
  off += align256(sizeof(int32_t) * (size_t)(n_pairs > 0 ? n_pairs : 1));
  w.vals_in = off;
  off += align256(sizeof(int32_t) * (size_t)(n_pairs > 0 ? n_pairs : 1));
  w.keys_out = off;
  off += align256(sizeof(int32_t) * (size_t)(n_pairs > 0 ? n_pairs : 1));
  w.max_count = off;
  off += 256;
  size_t scan_bytes = 0, sort_bytes = 0, max_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                (int)(n_atoms > 0 ? n_atoms : 1));
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (int)(n_pairs > 0 ? n_pairs : 1));
  cub::DeviceReduce::Max(nullptr, max_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                         (int)(n_atoms > 0 ? n_atoms : 1));
  size_t m = scan_bytes > sort_bytes ? scan_bytes : sort_bytes;
  m = m > max_bytes ? m : max_bytes;
  w.cub = off;
  off += align256(m + 256);
  w.total = off;
  return w;
}

}  // namespace
}  // namespace petb200

using namespace petb200;

extern "C" PETB200_API int petb200_nl_filter_count(const float* positions, const float* cells,
                                       const int32_t* system_of_atom, const int32_t* centers,
                                       const int32_t* neighbors, const int32_t* shifts,
                                       int64_t n_pairs, int64_t n_atoms, float cutoff,
                                       int32_t* keep, int32_t* counts, cudaStream_t stream) {
  (void)n_atoms;
  if (n_pairs == 0) return PETB200_OK;
  PETB200_REQUIRE(n_pairs < (1ll << 31), "nl_filter_count: more than 2^31 pairs");
  nl_filter_count_kernel<<<(unsigned)ceil_div(n_pairs, 256), 256, 0, stream>>>(
      positions, cells, system_of_atom, centers, neighbors, shifts, n_pairs, cutoff, keep, counts);
  return check_launch("nl_filter_count");
}

extern "C" PETB200_API size_t petb200_csr_build_workspace(int64_t n_pairs, int64_t n_atoms) {
  return csr_layout(n_pairs, n_atoms).total;
}

extern "C" PETB200_API int petb200_csr_build(const int32_t* centers, const int32_t* keep,
                                 const int32_t* counts, int64_t n_pairs, int64_t n_atoms,
                                 int32_t* row_ptr, int32_t* perm, int32_t* stats, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream) {
  CsrWorkspace w = csr_layout(n_pairs, n_atoms);
  if (workspace_bytes < w.total) {
    set_error("csr_build: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return PETB200_ERR_WORKSPACE;
  }
  PETB200_REQUIRE(n_pairs < (1ll << 31) && n_atoms < (1ll << 31) - 1, "csr_build: index overflow");
  char* base = static_cast<char*>(workspace);
  int32_t* keys_in = reinterpret_cast<int32_t*>(base + w.keys_in);
  int32_t* vals_in = reinterpret_cast<int32_t*>(base + w.vals_in);
  int32_t* keys_out = reinterpret_cast<int32_t*>(base + w.keys_out);
  int32_t* max_count = reinterpret_cast<int32_t*>(base + w.max_count);
  void* cub_tmp = base + w.cub;
  size_t cub_bytes = w.total - w.cub;
  if (n_atoms > 0) {
    cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, counts, row_ptr, (int)n_atoms, stream);
    cub::DeviceReduce::Max(cub_tmp, cub_bytes, counts, max_count, (int)n_atoms, stream);
  }
  finish_row_ptr_kernel<<<1, 32, 0, stream>>>(counts, n_atoms, row_ptr, stats, max_count);
  if (n_pairs > 0) {
    make_sort_keys_kernel<<<(unsigned)ceil_div(n_pairs, 256), 256, 0, stream>>>(
        centers, keep, n_pairs, (int32_t)n_atoms, keys_in, vals_in);
    int end_bit = 1;
    while ((1ll << end_bit) <= n_atoms) ++end_bit;  // keys are in [0, n_atoms]
    cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, keys_in, keys_out, vals_in, perm,
                                    (int)n_pairs, 0, end_bit, stream);
  }
  return check_launch("csr_build");
}

extern "C" PETB200_API int petb200_csr_gather(const int32_t* perm, const int32_t* centers,
                                  const int32_t* neighbors, const int32_t* shifts,
                                  int64_t n_edges, int32_t* ctr, int32_t* col, int32_t* shift_csr,
                                  cudaStream_t stream) {
  if (n_edges == 0) return PETB200_OK;
  csr_gather_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(
      perm, centers, neighbors, shifts, n_edges, ctr, col, shift_csr);
  return check_launch("csr_gather");
}

extern "C" PETB200_API int petb200_reverse_map(const int32_t* row_ptr, const int32_t* ctr, const int32_t* col,
                                   const int32_t* shift_csr, int64_t n_edges, int64_t n_rows,
                                   int32_t* rev, int32_t* n_missing, cudaStream_t stream) {
  if (n_edges == 0) return PETB200_OK;
  reverse_map_kernel<<<(unsigned)ceil_div(n_edges, 128), 128, 0, stream>>>(
      row_ptr, ctr, col, shift_csr, n_edges, n_rows, rev, n_missing);
  return check_launch("reverse_map");
}

extern "C" PETB200_API int petb200_csr_to_nef(const float* x_csr, const int32_t* row_ptr, int64_t n_atoms,
                                  int64_t n_edges, int width_m, int d, float* x_nef,
                                  cudaStream_t stream) {
  (void)n_edges;
  int64_t total = n_atoms * width_m * d;
  if (total == 0) return PETB200_OK;
  csr_to_nef_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(x_csr, row_ptr, n_atoms,
                                                                        width_m, d, x_nef);
  return check_launch("csr_to_nef");
}

extern "C" PETB200_API int petb200_nef_to_csr(const float* x_nef, const int32_t* row_ptr, const int32_t* ctr,
                                  int64_t n_atoms, int64_t n_edges, int width_m, int d,
                                  float* x_csr, cudaStream_t stream) {
  (void)n_atoms;
  int64_t total = n_edges * d;
  if (total == 0) return PETB200_OK;
  nef_to_csr_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(x_nef, row_ptr, ctr,
                                                                        n_edges, width_m, d, x_csr);
  return check_launch("nef_to_csr");
}
