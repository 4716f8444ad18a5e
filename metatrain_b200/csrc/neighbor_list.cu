// Cell-list neighbor list on the GPU.
//
// The reference builds its neighbor list on the CPU with the C library vesin
// (src/metatrain/utils/neighbor_lists.py:125-201, call :131) inside the DataLoader; inside an
// MD step that host round trip dominates once the model step takes milliseconds.  This file
// produces the same object — every ordered pair (i, j, S) with |r_j + S.cell - r_i| <= cutoff,
// excluding (i, i, 0), grouped by centre i — directly in HBM, so it can feed
// petb200_nl_filter_count / csr_build without touching the host.
//
// Algorithm: wrap atoms into the cell, bin them on a grid whose cells are at least `cutoff`
// wide along each lattice direction (fewer, larger bins — and several periodic images — when
// the cell is smaller than the cutoff), sort atoms by bin (CUB radix sort, stable), then one
// thread per atom walks the (2 reach + 1)^3 neighbouring bins twice: count, scan, fill.
// Pairs are emitted in a deterministic order (bin offsets, then atom index within a bin).
// Candidates are accepted with a relative margin of 2e-6 on the cutoff: the model's own
// symmetric fp32 filter (petb200_nl_filter_count) takes the final decision.
#include <cub/cub.cuh>

#include "common.cuh"

namespace petb200 {
namespace {

struct NlGrid {
  float cell[9];   // rows = lattice vectors
  float inv[9];    // inverse: frac = pos * inv (row vector convention)
  float origin[3];
  int nb[3];
  int reach[3];
  int periodic;
  float cutoff2;
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct NlWorkspace {
  size_t wrapped, wrap, keys_in, keys_out, idx_in, idx_out, bin_start, cub, total;
  size_t cub_bytes;
};

NlWorkspace nl_layout(int64_t n, int64_t n_bins) {
  NlWorkspace w;
  size_t off = 0;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  w.wrapped = off; off += align256(sizeof(float) * 3 * nn);
  w.wrap = off; off += align256(sizeof(int32_t) * 3 * nn);
  w.keys_in = off; off += align256(sizeof(int32_t) * nn);
  w.keys_out = off; off += align256(sizeof(int32_t) * nn);
  w.idx_in = off; off += align256(sizeof(int32_t) * nn);
  w.idx_out = off; off += align256(sizeof(int32_t) * nn);
  w.bin_start = off; off += align256(sizeof(int32_t) * (size_t)(n_bins + 1));
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)nn);
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                (int)nn + 1);
  w.cub_bytes = (sort_bytes > scan_bytes ? sort_bytes : scan_bytes) + 256;
  w.cub = off; off += align256(w.cub_bytes);
  w.total = off;
  return w;
}

bool make_grid(const float* cell_host, const float* origin_host, int periodic, float cutoff,
               int64_t n_atoms, NlGrid& g) {
  double c[9];
  for (int k = 0; k < 9; ++k) { c[k] = cell_host[k]; g.cell[k] = cell_host[k]; }
  for (int k = 0; k < 3; ++k) g.origin[k] = origin_host ? origin_host[k] : 0.f;
  const double det = c[0] * (c[4] * c[8] - c[5] * c[7]) - c[1] * (c[3] * c[8] - c[5] * c[6]) +
                     c[2] * (c[3] * c[7] - c[4] * c[6]);
  if (!(fabs(det) > 1e-12)) return false;
  const double inv[9] = {(c[4] * c[8] - c[5] * c[7]) / det, (c[2] * c[7] - c[1] * c[8]) / det,
                         (c[1] * c[5] - c[2] * c[4]) / det, (c[5] * c[6] - c[3] * c[8]) / det,
                         (c[0] * c[8] - c[2] * c[6]) / det, (c[2] * c[3] - c[0] * c[5]) / det,
                         (c[3] * c[7] - c[4] * c[6]) / det, (c[1] * c[6] - c[0] * c[7]) / det,
                         (c[0] * c[4] - c[1] * c[3]) / det};
  for (int k = 0; k < 9; ++k) g.inv[k] = (float)inv[k];
  // perpendicular height along lattice direction k = 1 / |column k of inv|
  int64_t budget = 4 * n_atoms + 64;
  for (int k = 0; k < 3; ++k) {
    const double col = sqrt(inv[k] * inv[k] + inv[3 + k] * inv[3 + k] + inv[6 + k] * inv[6 + k]);
    const double height = 1.0 / col;
    int nb = (int)floor(height / cutoff);
    if (nb < 1) nb = 1;
    g.nb[k] = nb;
  }
  while ((int64_t)g.nb[0] * g.nb[1] * g.nb[2] > budget) {  // sparse boxes: fewer, larger bins
    int k = g.nb[0] >= g.nb[1] ? (g.nb[0] >= g.nb[2] ? 0 : 2) : (g.nb[1] >= g.nb[2] ? 1 : 2);
    g.nb[k] = (g.nb[k] + 1) / 2;
  }
  for (int k = 0; k < 3; ++k) {
    const double col = sqrt(inv[k] * inv[k] + inv[3 + k] * inv[3 + k] + inv[6 + k] * inv[6 + k]);
    const double width = (1.0 / col) / g.nb[k];
    int r = (int)ceil(cutoff / width - 1e-9);
    g.reach[k] = r < 1 ? 1 : r;
  }
  g.periodic = periodic;
  const float rc = cutoff * (1.0f + 2e-6f) + 1e-7f;
  g.cutoff2 = rc * rc;
  return true;
}

__global__ void nl_bin_kernel(const float* __restrict__ pos, int64_t n, NlGrid g,
                              float* __restrict__ wrapped, int32_t* __restrict__ wrap,
                              int32_t* __restrict__ keys, int32_t* __restrict__ idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pos[3 * i] - g.origin[0], y = pos[3 * i + 1] - g.origin[1],
              z = pos[3 * i + 2] - g.origin[2];
  float f[3] = {x * g.inv[0] + y * g.inv[3] + z * g.inv[6], x * g.inv[1] + y * g.inv[4] + z * g.inv[7],
                x * g.inv[2] + y * g.inv[5] + z * g.inv[8]};
  int b[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float w = g.periodic ? floorf(f[k]) : 0.f;
    f[k] -= w;
    if (g.periodic && f[k] >= 1.0f) { f[k] -= 1.0f; w += 1.0f; }  // fp32 edge case
    wrap[3 * i + k] = (int32_t)w;
    int bk = (int)(f[k] * g.nb[k]);
    b[k] = bk < 0 ? 0 : (bk >= g.nb[k] ? g.nb[k] - 1 : bk);
  }
  wrapped[3 * i + 0] = f[0] * g.cell[0] + f[1] * g.cell[3] + f[2] * g.cell[6];
  wrapped[3 * i + 1] = f[0] * g.cell[1] + f[1] * g.cell[4] + f[2] * g.cell[7];
  wrapped[3 * i + 2] = f[0] * g.cell[2] + f[1] * g.cell[5] + f[2] * g.cell[8];
  keys[i] = (b[0] * g.nb[1] + b[1]) * g.nb[2] + b[2];
  idx[i] = (int32_t)i;
}

__global__ void nl_bin_bounds_kernel(const int32_t* __restrict__ sorted_keys, int64_t n,
                                     int n_bins, int32_t* __restrict__ bin_start) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int cur = i < n ? sorted_keys[i] : n_bins;
  const int prev = i > 0 ? sorted_keys[i - 1] : -1;
  for (int b = prev + 1; b <= cur; ++b) bin_start[b] = (int32_t)i;
}

// FILL = false: counts[i] = number of neighbours; FILL = true: write the pairs at offsets[i]
template <bool FILL>
__global__ void nl_pairs_kernel(const float* __restrict__ wrapped, const int32_t* __restrict__ wrap,
                                const int32_t* __restrict__ bin_of_atom,
                                const int32_t* __restrict__ sorted_idx,
                                const int32_t* __restrict__ bin_start, int64_t n, NlGrid g,
                                int32_t* __restrict__ counts, const int32_t* __restrict__ offsets,
                                int32_t* __restrict__ centers, int32_t* __restrict__ neighbors,
                                int32_t* __restrict__ shifts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xi = wrapped[3 * i], yi = wrapped[3 * i + 1], zi = wrapped[3 * i + 2];
  const int key = bin_of_atom[i];  // the bin nl_bin_kernel put this atom in
  const int bi[3] = {key / (g.nb[1] * g.nb[2]), (key / g.nb[2]) % g.nb[1], key % g.nb[2]};
  const int wi0 = wrap[3 * i], wi1 = wrap[3 * i + 1], wi2 = wrap[3 * i + 2];
  int count = 0;
  int64_t out = FILL ? offsets[i] : 0;
  for (int da = -g.reach[0]; da <= g.reach[0]; ++da) {
    int ta = bi[0] + da, sa = 0;
    if (g.periodic) { sa = (ta >= 0) ? ta / g.nb[0] : -((-ta + g.nb[0] - 1) / g.nb[0]); ta -= sa * g.nb[0]; }
    else if (ta < 0 || ta >= g.nb[0]) continue;
    for (int db = -g.reach[1]; db <= g.reach[1]; ++db) {
      int tb = bi[1] + db, sb = 0;
      if (g.periodic) { sb = (tb >= 0) ? tb / g.nb[1] : -((-tb + g.nb[1] - 1) / g.nb[1]); tb -= sb * g.nb[1]; }
      else if (tb < 0 || tb >= g.nb[1]) continue;
      for (int dc = -g.reach[2]; dc <= g.reach[2]; ++dc) {
        int tc = bi[2] + dc, sc = 0;
        if (g.periodic) { sc = (tc >= 0) ? tc / g.nb[2] : -((-tc + g.nb[2] - 1) / g.nb[2]); tc -= sc * g.nb[2]; }
        else if (tc < 0 || tc >= g.nb[2]) continue;
        const float ox = sa * g.cell[0] + sb * g.cell[3] + sc * g.cell[6] - xi;
        const float oy = sa * g.cell[1] + sb * g.cell[4] + sc * g.cell[7] - yi;
        const float oz = sa * g.cell[2] + sb * g.cell[5] + sc * g.cell[8] - zi;
        const int bin = (ta * g.nb[1] + tb) * g.nb[2] + tc;
        const int lo = bin_start[bin], hi = bin_start[bin + 1];
        for (int k = lo; k < hi; ++k) {
          const int j = sorted_idx[k];
          const float dx = wrapped[3 * (int64_t)j] + ox, dy = wrapped[3 * (int64_t)j + 1] + oy,
                      dz = wrapped[3 * (int64_t)j + 2] + oz;
          const float d2 = dx * dx + dy * dy + dz * dz;
          if (d2 <= g.cutoff2 && !(j == i && sa == 0 && sb == 0 && sc == 0)) {
            if (FILL) {
              centers[out] = (int32_t)i;
              neighbors[out] = j;
              shifts[3 * out + 0] = sa - wrap[3 * (int64_t)j] + wi0;
              shifts[3 * out + 1] = sb - wrap[3 * (int64_t)j + 1] + wi1;
              shifts[3 * out + 2] = sc - wrap[3 * (int64_t)j + 2] + wi2;
              ++out;
            } else {
              ++count;
            }
          }
        }
      }
    }
  }
  if (!FILL) counts[i] = count;
}

}  // namespace
}  // namespace petb200

using namespace petb200;

extern "C" PETB200_API int64_t petb200_nl_num_bins(const float* cell_host, int periodic,
                                                   float cutoff, int64_t n_atoms) {
  NlGrid g;
  if (!make_grid(cell_host, nullptr, periodic, cutoff, n_atoms, g)) return -1;
  return (int64_t)g.nb[0] * g.nb[1] * g.nb[2];
}

extern "C" PETB200_API size_t petb200_nl_workspace(int64_t n_atoms, int64_t n_bins) {
  return nl_layout(n_atoms, n_bins).total;
}

extern "C" PETB200_API int petb200_nl_count(const float* positions, int64_t n_atoms,
                                            const float* cell_host, const float* origin_host,
                                            int periodic, float cutoff, void* workspace,
                                            size_t workspace_bytes, int32_t* offsets,
                                            cudaStream_t stream) {
  NlGrid g;
  PETB200_REQUIRE(make_grid(cell_host, origin_host, periodic, cutoff, n_atoms, g),
                  "nl_count: singular cell");
  PETB200_REQUIRE(n_atoms < (1ll << 31) - 1, "nl_count: too many atoms");
  const int n_bins = g.nb[0] * g.nb[1] * g.nb[2];
  NlWorkspace w = nl_layout(n_atoms, n_bins);
  if (workspace_bytes < w.total) {
    set_error("nl_count: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return PETB200_ERR_WORKSPACE;
  }
  if (n_atoms == 0) return PETB200_OK;
  char* base = static_cast<char*>(workspace);
  float* wrapped = reinterpret_cast<float*>(base + w.wrapped);
  int32_t* wrap = reinterpret_cast<int32_t*>(base + w.wrap);
  int32_t* keys_in = reinterpret_cast<int32_t*>(base + w.keys_in);
  int32_t* keys_out = reinterpret_cast<int32_t*>(base + w.keys_out);
  int32_t* idx_in = reinterpret_cast<int32_t*>(base + w.idx_in);
  int32_t* idx_out = reinterpret_cast<int32_t*>(base + w.idx_out);
  int32_t* bin_start = reinterpret_cast<int32_t*>(base + w.bin_start);
  size_t cub_bytes = w.cub_bytes;
  const unsigned grid = (unsigned)ceil_div(n_atoms, 128);
  nl_bin_kernel<<<grid, 128, 0, stream>>>(positions, n_atoms, g, wrapped, wrap, keys_in, idx_in);
  int end_bit = 1;
  while ((1ll << end_bit) <= n_bins) ++end_bit;
  cub::DeviceRadixSort::SortPairs(base + w.cub, cub_bytes, keys_in, keys_out, idx_in, idx_out,
                                  (int)n_atoms, 0, end_bit, stream);
  nl_bin_bounds_kernel<<<(unsigned)ceil_div(n_atoms + 1, 128), 128, 0, stream>>>(keys_out, n_atoms,
                                                                                 n_bins, bin_start);
  // counts land in offsets[0..N), then an exclusive scan over N+1 entries closes the row table
  cudaMemsetAsync(offsets + n_atoms, 0, sizeof(int32_t), stream);
  nl_pairs_kernel<false><<<grid, 128, 0, stream>>>(wrapped, wrap, keys_in, idx_out, bin_start, n_atoms, g,
                                                   offsets, nullptr, nullptr, nullptr, nullptr);
  cub_bytes = w.cub_bytes;
  cub::DeviceScan::ExclusiveSum(base + w.cub, cub_bytes, offsets, offsets, (int)n_atoms + 1, stream);
  return check_launch("nl_count");
}

extern "C" PETB200_API int petb200_nl_fill(int64_t n_atoms, const float* cell_host,
                                           const float* origin_host, int periodic, float cutoff,
                                           const void* workspace, size_t workspace_bytes,
                                           const int32_t* offsets, int32_t* centers,
                                           int32_t* neighbors, int32_t* shifts,
                                           cudaStream_t stream) {
  NlGrid g;
  PETB200_REQUIRE(make_grid(cell_host, origin_host, periodic, cutoff, n_atoms, g),
                  "nl_fill: singular cell");
  const int n_bins = g.nb[0] * g.nb[1] * g.nb[2];
  NlWorkspace w = nl_layout(n_atoms, n_bins);
  if (workspace_bytes < w.total) {
    set_error("nl_fill: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return PETB200_ERR_WORKSPACE;
  }
  if (n_atoms == 0) return PETB200_OK;
  const char* base = static_cast<const char*>(workspace);
  nl_pairs_kernel<true><<<(unsigned)ceil_div(n_atoms, 128), 128, 0, stream>>>(
      reinterpret_cast<const float*>(base + w.wrapped), reinterpret_cast<const int32_t*>(base + w.wrap),
      reinterpret_cast<const int32_t*>(base + w.keys_in),
      reinterpret_cast<const int32_t*>(base + w.idx_out),
      reinterpret_cast<const int32_t*>(base + w.bin_start), n_atoms, g, nullptr, offsets, centers,
      neighbors, shifts);
  return check_launch("nl_fill");
}
