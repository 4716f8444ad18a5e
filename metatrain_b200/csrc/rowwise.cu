// Row-wise (HBM-bound) kernels: embeddings, RMSNorm statistics, the message-reversal +
// LayerNorm "edge scatter" kernel, the geometry embedder and the readout reductions.
//
// All of them use one warp per row with float4 lanes (d = 128 -> one float4 per lane,
// d = 256 -> two), i.e. fully coalesced 512 B / 1 KiB row accesses, no shared memory and
// no atomics.  Reference lines are cited at each kernel.
#include <float.h>

#include "common.cuh"

namespace petb200 {
namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ int64_t global_warp() {
  return ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
}
inline unsigned warp_grid(int64_t rows) { return (unsigned)ceil_div(rows, kWarpsPerBlock); }

// ------------------------------------------------------------------ embedding lookup
// torch.nn.Embedding (backend.py:515-516, transformer.py:509-511)
__global__ void embedding_kernel(const float* __restrict__ table, const int32_t* __restrict__ idx,
                                 int64_t n_rows, int d4, float4* __restrict__ out, int64_t ld4) {
  int64_t row = global_warp();
  if (row >= n_rows) return;
  const float4* src = reinterpret_cast<const float4*>(table) + (int64_t)idx[row] * d4;
  for (int c = threadIdx.x & 31; c < d4; c += 32) out[row * ld4 + c] = __ldg(src + c);
}

// x[row,:] += table[idx[row],:]: the per-system conditioning embedding broadcast to the atoms of each
// system and added to the node features (backend.py:551-552, conditioning.py:97-100)
__global__ void add_gathered_rows_kernel(const float* __restrict__ table, const int32_t* __restrict__ idx,
                                         int64_t n_rows, int d4, float4* __restrict__ x, int64_t ld4) {
  int64_t row = global_warp();
  if (row >= n_rows) return;
  const float4* src = reinterpret_cast<const float4*>(table) + (int64_t)idx[row] * d4;
  for (int c = threadIdx.x & 31; c < d4; c += 32) {
    const float4 a = x[row * ld4 + c], b = __ldg(src + c);
    x[row * ld4 + c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// ------------------------------------------------------------------ weight preparation
__global__ void transpose_scale_kernel(const float* __restrict__ in, int rows, int cols,
                                       const float* __restrict__ col_scale,
                                       float* __restrict__ out_t, float* __restrict__ out_s) {
  __shared__ float tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int r0 = 0; r0 < 32; r0 += 8) {
    int r = blockIdx.y * 32 + threadIdx.y + r0;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = in[(int64_t)r * cols + c] * (col_scale ? col_scale[c] : 1.f);
      if (out_s) out_s[(int64_t)r * cols + c] = v;
    }
    tile[threadIdx.y + r0][threadIdx.x] = v;
  }
  __syncthreads();
  if (!out_t) return;
  int r = blockIdx.y * 32 + threadIdx.x;  // transposed: fastest index walks input rows
  for (int c0 = 0; c0 < 32; c0 += 8) {
    int cc = blockIdx.x * 32 + threadIdx.y + c0;
    if (r < rows && cc < cols) out_t[(int64_t)cc * rows + r] = tile[threadIdx.x][threadIdx.y + c0];
  }
}

// ------------------------------------------------------------------ GNN-layer input
// transformer.py:500-519: cat[e] = [Linear(4->d)([r, d]) | NbrEmb[z_j] | m_e]
__global__ void compress_input_kernel(const float* __restrict__ vec, const float* __restrict__ dist,
                                      const float* __restrict__ w_geo,
                                      const float* __restrict__ b_geo,
                                      const float* __restrict__ nbr_table,
                                      const int32_t* __restrict__ z_nbr,
                                      const float* __restrict__ msg, int64_t n_edges,
                                      float* __restrict__ cat) {
  constexpr int D = 128;
  int64_t e = global_warp();
  if (e >= n_edges) return;
  const int lane = threadIdx.x & 31;
  const int width = nbr_table ? 3 * D : 2 * D;
  const float x = vec[3 * e], y = vec[3 * e + 1], z = vec[3 * e + 2], dd = dist[e];
  float g[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = lane * 4 + j;
    float4 w = __ldg(reinterpret_cast<const float4*>(w_geo) + c);  // W_geo[c, 0..3]
    g[j] = fmaf(w.x, x, fmaf(w.y, y, fmaf(w.z, z, fmaf(w.w, dd, b_geo[c]))));
  }
  float4* dst = reinterpret_cast<float4*>(cat + e * width);
  dst[lane] = make_float4(g[0], g[1], g[2], g[3]);
  int off = D / 4;
  if (nbr_table) {
    dst[off + lane] = __ldg(reinterpret_cast<const float4*>(nbr_table) + (int64_t)z_nbr[e] * (D / 4) + lane);
    off += D / 4;
  }
  dst[off + lane] = __ldg(reinterpret_cast<const float4*>(msg) + e * (D / 4) + lane);
}

// backward of the geometry embedder: d_(r,d)[e] (+)= W_geo^T d_geo[e]
// Eight lanes per edge (four edges per warp): 3 shuffle steps on 4 values per 4 rows instead of
// 5 steps per row; every load is a coalesced 128-byte segment.  Grid-stride over edge groups with
// the geometry weights of this lane's 16 columns in registers (they were one L1 load per 4 FMAs).
// Requires d = 128.
__global__ void geom_embed_bwd_kernel(const float* __restrict__ d_geo, int64_t ld,
                                      const float* __restrict__ w_geo, int64_t n_edges,
                                      int accumulate, float* __restrict__ d_vec,
                                      float* __restrict__ d_dist) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  float4 w[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) w[j][c] = __ldg(reinterpret_cast<const float4*>(w_geo) + (sub + 8 * j) * 4 + c);
  const int64_t n_groups = ceil_div(n_edges, 4);
  const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t grp = global_warp(); grp < n_groups; grp += stride) {
    const int64_t e = grp * 4 + (lane >> 3);
    const bool ok = e < n_edges;
    float4 gv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      gv[j] = ok ? __ldg(reinterpret_cast<const float4*>(d_geo + e * ld) + sub + 8 * j)
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    float ax = 0.f, ay = 0.f, az = 0.f, ad = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gg[4] = {gv[j].x, gv[j].y, gv[j].z, gv[j].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ax = fmaf(gg[c], w[j][c].x, ax);
        ay = fmaf(gg[c], w[j][c].y, ay);
        az = fmaf(gg[c], w[j][c].z, az);
        ad = fmaf(gg[c], w[j][c].w, ad);
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      ax += __shfl_xor_sync(0xffffffffu, ax, o);
      ay += __shfl_xor_sync(0xffffffffu, ay, o);
      az += __shfl_xor_sync(0xffffffffu, az, o);
      ad += __shfl_xor_sync(0xffffffffu, ad, o);
    }
    if (ok && sub == 0) {
      if (accumulate) {
        d_vec[3 * e] += ax; d_vec[3 * e + 1] += ay; d_vec[3 * e + 2] += az; d_dist[e] += ad;
      } else {
        d_vec[3 * e] = ax; d_vec[3 * e + 1] = ay; d_vec[3 * e + 2] = az; d_dist[e] = ad;
      }
    }
  }
}

// ------------------------------------------ residual featurizer: message averaging
// backend.py:640-647: m_next[e] = 0.5 (m[e] + t[rev[e]])  (one warp per edge, d = 128 V)
__global__ void avg_reverse_fwd_kernel(const float* __restrict__ m, const float* __restrict__ t,
                                       const int32_t* __restrict__ rev, int64_t n_edges, int v4,
                                       float* __restrict__ out) {
  const int64_t e = global_warp();
  if (e >= n_edges) return;
  const int lane = threadIdx.x & 31;
  const int64_t r = rev[e];
  for (int c = lane; c < v4; c += 32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(m) + e * v4 + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(t) + r * v4 + c);
    reinterpret_cast<float4*>(out)[e * v4 + c] =
        make_float4(0.5f * (a.x + b.x), 0.5f * (a.y + b.y), 0.5f * (a.z + b.z), 0.5f * (a.w + b.w));
  }
}
// its backward: d_t[e] += 0.5 d_next[rev[e]] (rev is an involution: a gather, no atomics),
// d_m[e] = 0.5 d_next[e]
__global__ void avg_reverse_bwd_kernel(const float* __restrict__ d_next, const int32_t* __restrict__ rev,
                                       int64_t n_edges, int v4, float* __restrict__ d_t,
                                       float* __restrict__ d_m) {
  const int64_t e = global_warp();
  if (e >= n_edges) return;
  const int lane = threadIdx.x & 31;
  const int64_t r = rev[e];
  for (int c = lane; c < v4; c += 32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(d_next) + e * v4 + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(d_next) + r * v4 + c);
    float4 g = reinterpret_cast<float4*>(d_t)[e * v4 + c];
    g.x += 0.5f * b.x; g.y += 0.5f * b.y; g.z += 0.5f * b.z; g.w += 0.5f * b.w;
    reinterpret_cast<float4*>(d_t)[e * v4 + c] = g;
    reinterpret_cast<float4*>(d_m)[e * v4 + c] = make_float4(0.5f * a.x, 0.5f * a.y, 0.5f * a.z, 0.5f * a.w);
  }
}

// ------------------------------------------------------------------------- RMSNorm
// torch.nn.RMSNorm(d) with eps=None -> finfo(fp32).eps (transformer.py:184-186,193)
template <int V>  // V float4 per lane: d = 128*V
__global__ void rms_rstd_kernel(const float* __restrict__ x, int64_t n_rows,
                                float* __restrict__ rstd) {
  int64_t row = global_warp();
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const float4* src = reinterpret_cast<const float4*>(x) + row * (32 * V);
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    float4 t = __ldg(src + v * 32 + lane);
    ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  ss = warp_sum(ss);
  if (lane == 0) rstd[row] = rsqrtf(ss / (float)(128 * V) + FLT_EPSILON);
}

// y = x * rstd * gamma (torch.nn.RMSNorm as a standalone op: the PostLN transformer normalises
// the residual stream itself, transformer.py:236-262), rstd kept for the backward
template <int V>
__global__ void rms_norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                    int64_t n_rows, float* __restrict__ y, float* __restrict__ rstd) {
  int64_t row = global_warp();
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  float4 t[V];
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    t[v] = __ldg(reinterpret_cast<const float4*>(x) + row * (32 * V) + v * 32 + lane);
    ss += t[v].x * t[v].x + t[v].y * t[v].y + t[v].z * t[v].z + t[v].w * t[v].w;
  }
  const float rs = rsqrtf(warp_sum(ss) / (float)(128 * V) + FLT_EPSILON);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + v * 32 + lane);
    reinterpret_cast<float4*>(y)[row * (32 * V) + v * 32 + lane] =
        make_float4(t[v].x * rs * g.x, t[v].y * rs * g.y, t[v].z * rs * g.z, t[v].w * rs * g.w);
  }
  if (lane == 0) rstd[row] = rs;
}

// out = base + rstd * (d_xhat - xhat * mean(d_xhat * xhat)),  xhat = x * rstd
template <int V>
__global__ void rms_bwd_kernel(const float* __restrict__ d_xhat, const float* __restrict__ x,
                               const float* __restrict__ rstd, const float* __restrict__ base,
                               const float* __restrict__ gamma /* nullable: d_xhat = dy * gamma */,
                               int64_t n_rows, float* __restrict__ out) {
  int64_t row = global_warp();
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const float rs = rstd[row];
  float4 g[V], xh[V];
  float dot = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    g[v] = __ldg(reinterpret_cast<const float4*>(d_xhat) + row * (32 * V) + v * 32 + lane);
    if (gamma) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(gamma) + v * 32 + lane);
      g[v] = make_float4(g[v].x * w.x, g[v].y * w.y, g[v].z * w.z, g[v].w * w.w);
    }
    float4 t = __ldg(reinterpret_cast<const float4*>(x) + row * (32 * V) + v * 32 + lane);
    xh[v] = make_float4(t.x * rs, t.y * rs, t.z * rs, t.w * rs);
    dot += g[v].x * xh[v].x + g[v].y * xh[v].y + g[v].z * xh[v].z + g[v].w * xh[v].w;
  }
  dot = warp_sum(dot) / (float)(128 * V);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    float4 b = base ? __ldg(reinterpret_cast<const float4*>(base) + row * (32 * V) + v * 32 + lane)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    o.x = b.x + rs * (g[v].x - xh[v].x * dot);
    o.y = b.y + rs * (g[v].y - xh[v].y * dot);
    o.z = b.z + rs * (g[v].z - xh[v].z * dot);
    o.w = b.w + rs * (g[v].w - xh[v].w * dot);
    reinterpret_cast<float4*>(out)[row * (32 * V) + v * 32 + lane] = o;
  }
}

// ------------------------------------------------- LayerNorm as a standalone op
// torch.nn.LayerNorm(d) (eps = 1e-5, affine) for the normalization = "LayerNorm" variant
// (transformer.py:181-186): y = (x - mean) * rstd * gamma + beta; mean and rstd kept.
template <int V>
__global__ void layer_norm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int64_t n_rows,
                                      float* __restrict__ y, float* __restrict__ mean,
                                      float* __restrict__ rstd) {
  int64_t row = global_warp();
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  float4 t[V];
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    t[v] = __ldg(reinterpret_cast<const float4*>(x) + row * (32 * V) + v * 32 + lane);
    s += t[v].x + t[v].y + t[v].z + t[v].w;
  }
  const float mu = warp_sum(s) / (float)(128 * V);
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    t[v] = make_float4(t[v].x - mu, t[v].y - mu, t[v].z - mu, t[v].w - mu);
    ss += t[v].x * t[v].x + t[v].y * t[v].y + t[v].z * t[v].z + t[v].w * t[v].w;
  }
  const float rs = rsqrtf(warp_sum(ss) / (float)(128 * V) + 1e-5f);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + v * 32 + lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + v * 32 + lane);
    reinterpret_cast<float4*>(y)[row * (32 * V) + v * 32 + lane] =
        make_float4(t[v].x * rs * g.x + b.x, t[v].y * rs * g.y + b.y, t[v].z * rs * g.z + b.z,
                    t[v].w * rs * g.w + b.w);
  }
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}
// out = base + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = d_y * gamma, xhat = (x - mean) * rstd
template <int V>
__global__ void layer_norm_bwd_kernel(const float* __restrict__ d_y, const float* __restrict__ x,
                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                      const float* __restrict__ gamma, const float* __restrict__ base,
                                      int64_t n_rows, float* __restrict__ out) {
  int64_t row = global_warp();
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const float mu = mean[row], rs = rstd[row];
  float4 g[V], xh[V];
  float sg = 0.f, dot = 0.f;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    g[v] = __ldg(reinterpret_cast<const float4*>(d_y) + row * (32 * V) + v * 32 + lane);
    const float4 w = __ldg(reinterpret_cast<const float4*>(gamma) + v * 32 + lane);
    g[v] = make_float4(g[v].x * w.x, g[v].y * w.y, g[v].z * w.z, g[v].w * w.w);
    const float4 t = __ldg(reinterpret_cast<const float4*>(x) + row * (32 * V) + v * 32 + lane);
    xh[v] = make_float4((t.x - mu) * rs, (t.y - mu) * rs, (t.z - mu) * rs, (t.w - mu) * rs);
    sg += g[v].x + g[v].y + g[v].z + g[v].w;
    dot += g[v].x * xh[v].x + g[v].y * xh[v].y + g[v].z * xh[v].z + g[v].w * xh[v].w;
  }
  sg = warp_sum(sg) / (float)(128 * V);
  dot = warp_sum(dot) / (float)(128 * V);
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const float4 b = base ? __ldg(reinterpret_cast<const float4*>(base) + row * (32 * V) + v * 32 + lane)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(out)[row * (32 * V) + v * 32 + lane] =
        make_float4(b.x + rs * (g[v].x - sg - xh[v].x * dot), b.y + rs * (g[v].y - sg - xh[v].y * dot),
                    b.z + rs * (g[v].z - sg - xh[v].z * dot), b.w + rs * (g[v].w - sg - xh[v].w * dot));
  }
}

// ----------------------------------------------- message reversal + LayerNorm (d = 128)
// backend.py:559-575: cc[e] = LayerNorm_256(cat[t_e, t_rev(e)]), eps = 1e-5, affine.
// Algorithmic traffic per edge: read 2 x 512 B (t_e, gathered t_rev(e)) + 4 B (rev),
// write 1024 B (+ 8 B statistics).
__global__ void combine_ln_fwd_kernel(const float* __restrict__ t, const int32_t* __restrict__ rev,
                                      const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int64_t n_edges,
                                      float* __restrict__ cc, float* __restrict__ mean_out,
                                      float* __restrict__ rstd_out) {
  int64_t e = global_warp();
  if (e >= n_edges) return;
  const int lane = threadIdx.x & 31;
  const float4 a = __ldg(reinterpret_cast<const float4*>(t) + e * 32 + lane);
  const float4 b = __ldg(reinterpret_cast<const float4*>(t) + (int64_t)rev[e] * 32 + lane);
  float s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  const float mean = warp_sum(s) * (1.0f / 256.0f);
  float4 da = make_float4(a.x - mean, a.y - mean, a.z - mean, a.w - mean);
  float4 db = make_float4(b.x - mean, b.y - mean, b.z - mean, b.w - mean);
  float v = da.x * da.x + da.y * da.y + da.z * da.z + da.w * da.w + db.x * db.x + db.y * db.y +
            db.z * db.z + db.w * db.w;
  const float rstd = rsqrtf(warp_sum(v) * (1.0f / 256.0f) + 1e-5f);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 32 + lane);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 32 + lane);
  float4* dst = reinterpret_cast<float4*>(cc) + e * 64;
  dst[lane] = make_float4(da.x * rstd * g0.x + b0.x, da.y * rstd * g0.y + b0.y,
                          da.z * rstd * g0.z + b0.z, da.w * rstd * g0.w + b0.w);
  dst[32 + lane] = make_float4(db.x * rstd * g1.x + b1.x, db.y * rstd * g1.y + b1.y,
                               db.z * rstd * g1.z + b1.z, db.w * rstd * g1.w + b1.w);
  if (lane == 0) {
    mean_out[e] = mean;
    rstd_out[e] = rstd;
  }
}

// d_cat[e] = LayerNorm backward of d_cc[e] w.r.t. cat[t_e, t_rev(e)]
__global__ void combine_ln_bwd_kernel(const float* __restrict__ d_cc, const float* __restrict__ t,
                                      const int32_t* __restrict__ rev,
                                      const float* __restrict__ gamma,
                                      const float* __restrict__ mean_in,
                                      const float* __restrict__ rstd_in, int64_t n_edges,
                                      float* __restrict__ d_cat) {
  int64_t e = global_warp();
  if (e >= n_edges) return;
  const int lane = threadIdx.x & 31;
  const float mean = mean_in[e], rstd = rstd_in[e];
  const float4 a = __ldg(reinterpret_cast<const float4*>(t) + e * 32 + lane);
  const float4 b = __ldg(reinterpret_cast<const float4*>(t) + (int64_t)rev[e] * 32 + lane);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 32 + lane);
  const float4 d0 = __ldg(reinterpret_cast<const float4*>(d_cc) + e * 64 + lane);
  const float4 d1 = __ldg(reinterpret_cast<const float4*>(d_cc) + e * 64 + 32 + lane);
  float xh[8] = {(a.x - mean) * rstd, (a.y - mean) * rstd, (a.z - mean) * rstd, (a.w - mean) * rstd,
                 (b.x - mean) * rstd, (b.y - mean) * rstd, (b.z - mean) * rstd, (b.w - mean) * rstd};
  float dx[8] = {d0.x * g0.x, d0.y * g0.y, d0.z * g0.z, d0.w * g0.w,
                 d1.x * g1.x, d1.y * g1.y, d1.z * g1.z, d1.w * g1.w};
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s1 += dx[k];
    s2 += dx[k] * xh[k];
  }
  s1 = warp_sum(s1) * (1.0f / 256.0f);
  s2 = warp_sum(s2) * (1.0f / 256.0f);
  float o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = rstd * (dx[k] - s1 - xh[k] * s2);
  float4* dst = reinterpret_cast<float4*>(d_cat) + e * 64;
  dst[lane] = make_float4(o[0], o[1], o[2], o[3]);
  dst[32 + lane] = make_float4(o[4], o[5], o[6], o[7]);
}

// out[e] = base[e] + d_cat[e, :128] + d_cat[rev[e], 128:]
__global__ void combine_scatter_bwd_kernel(const float* __restrict__ d_cat,
                                           const float* __restrict__ base,
                                           const int32_t* __restrict__ rev, int64_t n_edges,
                                           float* __restrict__ out) {
  int64_t e = global_warp();
  if (e >= n_edges) return;
  const int lane = threadIdx.x & 31;
  const float4 a = __ldg(reinterpret_cast<const float4*>(d_cat) + e * 64 + lane);
  const float4 b = __ldg(reinterpret_cast<const float4*>(d_cat) + (int64_t)rev[e] * 64 + 32 + lane);
  float4 c = base ? __ldg(reinterpret_cast<const float4*>(base) + e * 32 + lane)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
  reinterpret_cast<float4*>(out)[e * 32 + lane] =
      make_float4(a.x + b.x + c.x, a.y + b.y + c.y, a.z + b.z + c.z, a.w + b.w + c.w);
}

// ------------------------------------------------------------------------- readout
// backend.py:195-217 (last layers) and :762-772 (masked sum_j f_ij * e_ij); d = 128.
__global__ void readout_fwd_kernel(const float* __restrict__ node_feat,
                                   const float* __restrict__ edge_feat,
                                   const float* __restrict__ w_node, const float* __restrict__ b_node,
                                   const float* __restrict__ w_edge, const float* __restrict__ b_edge,
                                   const float* __restrict__ fc, const int32_t* __restrict__ row_ptr,
                                   int64_t n_atoms, int n_out, float* __restrict__ atomic,
                                   float* __restrict__ edge_pred) {
  int64_t i = global_warp();
  if (i >= n_atoms) return;
  const int lane = threadIdx.x & 31;
  const int lo = row_ptr[i], hi = row_ptr[i + 1];
  const float4 nf = __ldg(reinterpret_cast<const float4*>(node_feat) + i * 32 + lane);
  for (int p = 0; p < n_out; ++p) {
    const float4 wn = __ldg(reinterpret_cast<const float4*>(w_node) + p * 32 + lane);
    const float4 we = __ldg(reinterpret_cast<const float4*>(w_edge) + p * 32 + lane);
    float acc = warp_sum(nf.x * wn.x + nf.y * wn.y + nf.z * wn.z + nf.w * wn.w) + b_node[p];
    const float be = b_edge[p];
    if (edge_feat == nullptr) {
      // precomputed edge predictions (petb200_edge_head_fwd): a plain segmented sum, same order
      for (int e = lo; e < hi; ++e) acc = fmaf(fc[e], edge_pred[(int64_t)e * n_out + p], acc);
    } else {
      for (int e = lo; e < hi; ++e) {
        const float4 ef = __ldg(reinterpret_cast<const float4*>(edge_feat) + (int64_t)e * 32 + lane);
        float pe = warp_sum(ef.x * we.x + ef.y * we.y + ef.z * we.z + ef.w * we.w) + be;
        if (lane == 0) edge_pred[(int64_t)e * n_out + p] = pe;
        acc = fmaf(fc[e], pe, acc);
      }
    }
    if (lane == 0) atomic[i * n_out + p] = acc;
  }
}

__global__ void readout_bwd_node_kernel(const float* __restrict__ d_atomic,
                                        const float* __restrict__ w_node,
                                        const float* __restrict__ pre, int64_t n_atoms,
                                        int n_out, float* __restrict__ d_node_feat) {
  int64_t i = global_warp();
  if (i >= n_atoms) return;
  const int lane = threadIdx.x & 31;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = 0; p < n_out; ++p) {
    const float g = d_atomic[i * n_out + p];
    const float4 w = __ldg(reinterpret_cast<const float4*>(w_node) + p * 32 + lane);
    acc.x = fmaf(g, w.x, acc.x); acc.y = fmaf(g, w.y, acc.y);
    acc.z = fmaf(g, w.z, acc.z); acc.w = fmaf(g, w.w, acc.w);
  }
  if (pre) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(pre) + i * 32 + lane);
    acc.x *= dsiluf_(q.x); acc.y *= dsiluf_(q.y); acc.z *= dsiluf_(q.z); acc.w *= dsiluf_(q.w);
  }
  reinterpret_cast<float4*>(d_node_feat)[i * 32 + lane] = acc;
}

__global__ void readout_bwd_edge_kernel(const float* __restrict__ d_atomic,
                                        const float* __restrict__ edge_pred,
                                        const float* __restrict__ w_edge,
                                        const float* __restrict__ fc, const int32_t* __restrict__ ctr,
                                        const float* __restrict__ pre, int64_t n_edges, int n_out,
                                        float* __restrict__ d_edge_feat, float* __restrict__ d_fc) {
  int64_t e = global_warp();
  if (e >= n_edges) return;
  const int lane = threadIdx.x & 31;
  const int64_t i = ctr[e];
  const float f = fc[e];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float dfc = 0.f;
  for (int p = 0; p < n_out; ++p) {
    const float g = d_atomic[i * n_out + p];
    const float4 w = __ldg(reinterpret_cast<const float4*>(w_edge) + p * 32 + lane);
    acc.x = fmaf(g * f, w.x, acc.x); acc.y = fmaf(g * f, w.y, acc.y);
    acc.z = fmaf(g * f, w.z, acc.z); acc.w = fmaf(g * f, w.w, acc.w);
    dfc = fmaf(g, edge_pred[e * n_out + p], dfc);
  }
  if (pre) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(pre) + e * 32 + lane);
    acc.x *= dsiluf_(q.x); acc.y *= dsiluf_(q.y); acc.z *= dsiluf_(q.z); acc.w *= dsiluf_(q.w);
  }
  reinterpret_cast<float4*>(d_edge_feat)[e * 32 + lane] = acc;
  if (lane == 0 && d_fc) d_fc[e] += dfc;
}

// sum_over_atoms.py:31 — atoms of one structure are contiguous: one CTA per structure, a fixed
// (thread-strided, then warp-tree, then warp-ordered) summation order, so the result is deterministic.
constexpr int kSumThreads = 256;
__global__ void __launch_bounds__(kSumThreads) sum_over_atoms_kernel(const float* __restrict__ atomic,
                                                                     const int32_t* __restrict__ struct_ptr,
                                                                     int n_out, float* __restrict__ energies) {
  __shared__ float part[kSumThreads / 32];
  const int64_t b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lo = struct_ptr[b], hi = struct_ptr[b + 1];
  for (int p = 0; p < n_out; ++p) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four loads in flight per thread
    int i = lo + threadIdx.x;
    for (; i + 3 * kSumThreads < hi; i += 4 * kSumThreads) {
      a0 += atomic[(int64_t)i * n_out + p];
      a1 += atomic[(int64_t)(i + kSumThreads) * n_out + p];
      a2 += atomic[(int64_t)(i + 2 * kSumThreads) * n_out + p];
      a3 += atomic[(int64_t)(i + 3 * kSumThreads) * n_out + p];
    }
    for (; i < hi; i += kSumThreads) a0 += atomic[(int64_t)i * n_out + p];
    const float acc = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kSumThreads / 32; ++w) tot += part[w];
      energies[b * n_out + p] = tot;
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace petb200

using namespace petb200;

#define LAUNCH_ROWS(kernel, rows, ...)                                                  \
  do {                                                                                  \
    if ((rows) > 0)                                                                     \
      kernel<<<warp_grid(rows), kWarpsPerBlock * 32, 0, stream>>>(__VA_ARGS__);         \
  } while (0)

extern "C" PETB200_API int petb200_embedding(const float* table, const int32_t* idx, int64_t n_rows, int d,
                                 float* out, int64_t ld_out, cudaStream_t stream) {
  PETB200_REQUIRE(d % 4 == 0 && ld_out % 4 == 0, "embedding: d and ld_out must be multiples of 4");
  LAUNCH_ROWS(embedding_kernel, n_rows, table, idx, n_rows, d / 4, reinterpret_cast<float4*>(out),
              ld_out / 4);
  return check_launch("embedding");
}

extern "C" PETB200_API int petb200_add_gathered_rows(const float* table, const int32_t* idx, int64_t n_rows,
                                         int d, float* x, int64_t ld_x, cudaStream_t stream) {
  PETB200_REQUIRE(d % 4 == 0 && ld_x % 4 == 0, "add_gathered_rows: d and ld_x must be multiples of 4");
  LAUNCH_ROWS(add_gathered_rows_kernel, n_rows, table, idx, n_rows, d / 4, reinterpret_cast<float4*>(x),
              ld_x / 4);
  return check_launch("add_gathered_rows");
}

extern "C" PETB200_API int petb200_transpose_scale(const float* in, int rows, int cols,
                                       const float* col_scale, float* out_transposed,
                                       float* out_scaled, cudaStream_t stream) {
  if (rows == 0 || cols == 0) return PETB200_OK;
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
  transpose_scale_kernel<<<grid, dim3(32, 8), 0, stream>>>(in, rows, cols, col_scale,
                                                           out_transposed, out_scaled);
  return check_launch("transpose_scale");
}

extern "C" PETB200_API int petb200_compress_input(const float* edge_vec, const float* edge_dist,
                                      const float* w_geo, const float* b_geo,
                                      const float* nbr_table, const int32_t* z_neighbor,
                                      const float* messages, int64_t n_edges, int d, float* cat,
                                      cudaStream_t stream) {
  if (d != 128) {
    set_error("compress_input: only d_pet = 128 is built (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  LAUNCH_ROWS(compress_input_kernel, n_edges, edge_vec, edge_dist, w_geo, b_geo, nbr_table,
              z_neighbor, messages, n_edges, cat);
  return check_launch("compress_input");
}

extern "C" PETB200_API int petb200_geom_embed_bwd(const float* d_geo, int64_t ld, const float* w_geo,
                                      int64_t n_edges, int d, int accumulate, float* d_vec,
                                      float* d_dist, cudaStream_t stream) {
  if (d != 128) {
    set_error("geom_embed_bwd: only d_pet = 128 is built (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  PETB200_REQUIRE(ld % 4 == 0, "geom_embed_bwd: ld must be a multiple of 4");
  if (n_edges > 0) {
    const int64_t blocks = ceil_div(ceil_div(n_edges, 4), kWarpsPerBlock);
    const int64_t cap = 16 * kNumSMs;   // persistent: the weights are loaded once per thread
    geom_embed_bwd_kernel<<<(unsigned)(blocks < cap ? blocks : cap), kWarpsPerBlock * 32, 0, stream>>>(
        d_geo, ld, w_geo, n_edges, accumulate, d_vec, d_dist);
  }
  return check_launch("geom_embed_bwd");
}

extern "C" PETB200_API int petb200_avg_reverse_fwd(const float* m, const float* t, const int32_t* rev,
                                       int64_t n_edges, int d, float* out, cudaStream_t stream) {
  PETB200_REQUIRE(d % 4 == 0, "avg_reverse_fwd: d must be a multiple of 4");
  LAUNCH_ROWS(avg_reverse_fwd_kernel, n_edges, m, t, rev, n_edges, d / 4, out);
  return check_launch("avg_reverse_fwd");
}

extern "C" PETB200_API int petb200_avg_reverse_bwd(const float* d_next, const int32_t* rev, int64_t n_edges,
                                       int d, float* d_t, float* d_m, cudaStream_t stream) {
  PETB200_REQUIRE(d % 4 == 0, "avg_reverse_bwd: d must be a multiple of 4");
  LAUNCH_ROWS(avg_reverse_bwd_kernel, n_edges, d_next, rev, n_edges, d / 4, d_t, d_m);
  return check_launch("avg_reverse_bwd");
}

extern "C" PETB200_API int petb200_rms_rstd(const float* x, int64_t n_rows, int d, float* rstd,
                                cudaStream_t stream) {
  if (d == 128) {
    LAUNCH_ROWS(rms_rstd_kernel<1>, n_rows, x, n_rows, rstd);
  } else if (d == 256) {
    LAUNCH_ROWS(rms_rstd_kernel<2>, n_rows, x, n_rows, rstd);
  } else {
    set_error("rms_rstd: d must be 128 or 256 (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  return check_launch("rms_rstd");
}

extern "C" PETB200_API int petb200_rms_bwd(const float* d_xhat, const float* x, const float* rstd,
                               const float* base, int64_t n_rows, int d, float* out,
                               cudaStream_t stream) {
  if (d == 128) {
    LAUNCH_ROWS(rms_bwd_kernel<1>, n_rows, d_xhat, x, rstd, base, nullptr, n_rows, out);
  } else if (d == 256) {
    LAUNCH_ROWS(rms_bwd_kernel<2>, n_rows, d_xhat, x, rstd, base, nullptr, n_rows, out);
  } else {
    set_error("rms_bwd: d must be 128 or 256 (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  return check_launch("rms_bwd");
}

extern "C" PETB200_API int petb200_rms_norm_fwd(const float* x, const float* gamma, int64_t n_rows, int d,
                                    float* y, float* rstd, cudaStream_t stream) {
  if (d == 128) {
    LAUNCH_ROWS(rms_norm_fwd_kernel<1>, n_rows, x, gamma, n_rows, y, rstd);
  } else if (d == 256) {
    LAUNCH_ROWS(rms_norm_fwd_kernel<2>, n_rows, x, gamma, n_rows, y, rstd);
  } else {
    set_error("rms_norm_fwd: d must be 128 or 256 (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  return check_launch("rms_norm_fwd");
}

extern "C" PETB200_API int petb200_rms_norm_bwd(const float* d_y, const float* x, const float* rstd,
                                    const float* gamma, const float* base, int64_t n_rows, int d,
                                    float* out, cudaStream_t stream) {
  if (d == 128) {
    LAUNCH_ROWS(rms_bwd_kernel<1>, n_rows, d_y, x, rstd, base, gamma, n_rows, out);
  } else if (d == 256) {
    LAUNCH_ROWS(rms_bwd_kernel<2>, n_rows, d_y, x, rstd, base, gamma, n_rows, out);
  } else {
    set_error("rms_norm_bwd: d must be 128 or 256 (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  return check_launch("rms_norm_bwd");
}

extern "C" PETB200_API int petb200_layer_norm_fwd(const float* x, const float* gamma, const float* beta,
                                      int64_t n_rows, int d, float* y, float* mean, float* rstd,
                                      cudaStream_t stream) {
  if (d == 128) {
    LAUNCH_ROWS(layer_norm_fwd_kernel<1>, n_rows, x, gamma, beta, n_rows, y, mean, rstd);
  } else if (d == 256) {
    LAUNCH_ROWS(layer_norm_fwd_kernel<2>, n_rows, x, gamma, beta, n_rows, y, mean, rstd);
  } else {
    set_error("layer_norm_fwd: d must be 128 or 256 (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  return check_launch("layer_norm_fwd");
}

extern "C" PETB200_API int petb200_layer_norm_bwd(const float* d_y, const float* x, const float* mean,
                                      const float* rstd, const float* gamma, const float* base,
                                      int64_t n_rows, int d, float* out, cudaStream_t stream) {
  if (d == 128) {
    LAUNCH_ROWS(layer_norm_bwd_kernel<1>, n_rows, d_y, x, mean, rstd, gamma, base, n_rows, out);
  } else if (d == 256) {
    LAUNCH_ROWS(layer_norm_bwd_kernel<2>, n_rows, d_y, x, mean, rstd, gamma, base, n_rows, out);
  } else {
    set_error("layer_norm_bwd: d must be 128 or 256 (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  return check_launch("layer_norm_bwd");
}

extern "C" PETB200_API int petb200_combine_ln_fwd(const float* t, const int32_t* rev, const float* gamma,
                                      const float* beta, int64_t n_edges, int d, float* cc,
                                      float* mean, float* rstd, cudaStream_t stream) {
  if (d != 128) {
    set_error("combine_ln_fwd: only d_pet = 128 is built (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  LAUNCH_ROWS(combine_ln_fwd_kernel, n_edges, t, rev, gamma, beta, n_edges, cc, mean, rstd);
  return check_launch("combine_ln_fwd");
}

extern "C" PETB200_API int petb200_combine_ln_bwd(const float* d_cc, const float* t, const int32_t* rev,
                                      const float* gamma, const float* mean, const float* rstd,
                                      int64_t n_edges, int d, float* d_cat, cudaStream_t stream) {
  if (d != 128) {
    set_error("combine_ln_bwd: only d_pet = 128 is built (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  LAUNCH_ROWS(combine_ln_bwd_kernel, n_edges, d_cc, t, rev, gamma, mean, rstd, n_edges, d_cat);
  return check_launch("combine_ln_bwd");
}

extern "C" PETB200_API int petb200_combine_scatter_bwd(const float* d_cat, const float* base,
                                           const int32_t* rev, int64_t n_edges, int d, float* out,
                                           cudaStream_t stream) {
  if (d != 128) {
    set_error("combine_scatter_bwd: only d_pet = 128 is built (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  LAUNCH_ROWS(combine_scatter_bwd_kernel, n_edges, d_cat, base, rev, n_edges, out);
  return check_launch("combine_scatter_bwd");
}

extern "C" PETB200_API int petb200_readout_fwd(const float* node_feat, const float* edge_feat,
                                   const float* w_node, const float* b_node, const float* w_edge,
                                   const float* b_edge, const float* cutoff_factor,
                                   const int32_t* row_ptr, int64_t n_atoms, int64_t n_edges, int d,
                                   int n_out, float* atomic, float* edge_pred,
                                   cudaStream_t stream) {
  (void)n_edges;
  if (d != 128) {
    set_error("readout_fwd: only d_head = 128 is built (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  LAUNCH_ROWS(readout_fwd_kernel, n_atoms, node_feat, edge_feat, w_node, b_node, w_edge, b_edge,
              cutoff_factor, row_ptr, n_atoms, n_out, atomic, edge_pred);
  return check_launch("readout_fwd");
}

extern "C" PETB200_API int petb200_readout_bwd(const float* d_atomic, const float* edge_pred,
                                   const float* w_node, const float* w_edge,
                                   const float* cutoff_factor, const int32_t* ctr,
                                   const float* node_pre, const float* edge_pre, int64_t n_atoms,
                                   int64_t n_edges, int d, int n_out, float* d_node_feat,
                                   float* d_edge_feat, float* d_fc, cudaStream_t stream) {
  if (d != 128) {
    set_error("readout_bwd: only d_head = 128 is built (got %d)", d);
    return PETB200_ERR_UNSUPPORTED;
  }
  LAUNCH_ROWS(readout_bwd_node_kernel, n_atoms, d_atomic, w_node, node_pre, n_atoms, n_out,
              d_node_feat);
  LAUNCH_ROWS(readout_bwd_edge_kernel, n_edges, d_atomic, edge_pred, w_edge, cutoff_factor, ctr,
              edge_pre, n_edges, n_out, d_edge_feat, d_fc);
  return check_launch("readout_bwd");
}

extern "C" PETB200_API int petb200_sum_over_atoms(const float* atomic, const int32_t* struct_ptr,
                                      int64_t n_structures, int n_out, float* energies,
                                      cudaStream_t stream) {
  if (n_structures > 0)
    sum_over_atoms_kernel<<<(unsigned)n_structures, kSumThreads, 0, stream>>>(atomic, struct_ptr, n_out, energies);
  return check_launch("sum_over_atoms");
}
