// Per-atom multi-head attention over ragged token sets, forward and backward.
//
// Restates AttentionBlock.forward (src/metatrain/pet/modules/transformer.py:86-152) and
// manual_attention (:565-589) on the unpadded CSR token layout.  Tokens of atom i are
// {centre token (row E+i)} U {edge tokens row_ptr[i] .. row_ptr[i+1]}; the additive key
// bias is log(max(w, 1e-15)) with w = 1 for the centre token and the cutoff factor for an
// edge token (transformer.py:109-110, 524-540).  Padded NEF slots of the reference carry
// weight 1e-15 and are dropped here (SURVEY.md 8(a) note P).
//
// T = tokens per atom is small (26..49 for water at 4.5 A) and head_dim = 16, so this is a
// CUDA-core kernel: one CTA per atom, one warp per head, one lane per query (forward) or
// per query then per key (backward); K/V (and Q/dO in backward) are staged in shared
// memory once and read as warp-wide broadcasts.  No atomics: every output row has exactly
// one writer, so results are bit-reproducible.
#include "common.cuh"

namespace petb200 {
namespace {

constexpr int HD = 16;  // head dim (d_pet / num_heads = 128 / 8 with the default hypers)

__device__ __forceinline__ int64_t token_row(int p, int row_lo, int64_t n_edges, int64_t atom) {
  return p == 0 ? n_edges + atom : (int64_t)row_lo + (p - 1);
}

// qkv: [E+N, 3*D] (q | k | v), D = H*HD.  smem: Ks[T][D], Vs[T][D], lb[T]
template <int H>
__global__ void __launch_bounds__(H * 32) attention_fwd_kernel(
    const float* __restrict__ qkv, const int32_t* __restrict__ row_ptr,
    const float* __restrict__ fc, int64_t n_edges, float scale, float* __restrict__ out,
    float* __restrict__ lse) {
  constexpr int D = H * HD;
  extern __shared__ __align__(16) float smem[];
  const int64_t atom = blockIdx.x;
  const int lo = row_ptr[atom];
  const int T = row_ptr[atom + 1] - lo + 1;
  float* Ks = smem;
  float* Vs = Ks + (size_t)T * D;
  float* lb = Vs + (size_t)T * D;

  // stage K, V (coalesced float4) and the key bias
  for (int idx = threadIdx.x; idx < T * (D / 4); idx += blockDim.x) {
    int p = idx / (D / 4), c4 = idx % (D / 4);
    const float* row = qkv + token_row(p, lo, n_edges, atom) * (3 * D);
    reinterpret_cast<float4*>(Ks)[idx] = __ldg(reinterpret_cast<const float4*>(row + D) + c4);
    reinterpret_cast<float4*>(Vs)[idx] = __ldg(reinterpret_cast<const float4*>(row + 2 * D) + c4);
  }
  for (int p = threadIdx.x; p < T; p += blockDim.x)
    lb[p] = p == 0 ? 0.f : logf(fmaxf(fc[lo + p - 1], 1e-15f));
  __syncthreads();

  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int p = lane; p < T; p += 32) {
    const int64_t row = token_row(p, lo, n_edges, atom);
    float q[HD];
    {
      const float4* src = reinterpret_cast<const float4*>(qkv + row * (3 * D) + h * HD);
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 t = __ldg(src + c);
        q[4 * c] = t.x * scale; q[4 * c + 1] = t.y * scale;
        q[4 * c + 2] = t.z * scale; q[4 * c + 3] = t.w * scale;
      }
    }
    float m = -INFINITY, l = 0.f, o[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) o[c] = 0.f;
    for (int k = 0; k < T; ++k) {
      const float4* kk = reinterpret_cast<const float4*>(Ks + (size_t)k * D + h * HD);
      float s = lb[k];
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 t = kk[c];
        s = fmaf(q[4 * c], t.x, s); s = fmaf(q[4 * c + 1], t.y, s);
        s = fmaf(q[4 * c + 2], t.z, s); s = fmaf(q[4 * c + 3], t.w, s);
      }
      float m_new = fmaxf(m, s);
      float corr = expf(m - m_new);  // exp(-inf) = 0 on the first key
      float pexp = expf(s - m_new);
      l = l * corr + pexp;
      const float4* vv = reinterpret_cast<const float4*>(Vs + (size_t)k * D + h * HD);
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 t = vv[c];
        o[4 * c] = fmaf(pexp, t.x, o[4 * c] * corr);
        o[4 * c + 1] = fmaf(pexp, t.y, o[4 * c + 1] * corr);
        o[4 * c + 2] = fmaf(pexp, t.z, o[4 * c + 2] * corr);
        o[4 * c + 3] = fmaf(pexp, t.w, o[4 * c + 3] * corr);
      }
      m = m_new;
    }
    const float inv = 1.0f / l;
    float4* dst = reinterpret_cast<float4*>(out + row * D + h * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c)
      dst[c] = make_float4(o[4 * c] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv);
    lse[row * H + h] = m + logf(l);
  }
}

// smem: Qs, Ks, Vs, dOs [T][D]; Ls, Ds [T][H]; lb[T]; dlb[H][T]
template <int H>
__global__ void __launch_bounds__(H * 32) attention_bwd_kernel(
    const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ lse,
    const float* __restrict__ d_out, const int32_t* __restrict__ row_ptr,
    const float* __restrict__ fc, int64_t n_edges, float scale, float* __restrict__ d_qkv,
    float* __restrict__ d_fc) {
  constexpr int D = H * HD;
  extern __shared__ __align__(16) float smem[];
  const int64_t atom = blockIdx.x;
  const int lo = row_ptr[atom];
  const int T = row_ptr[atom + 1] - lo + 1;
  float* Qs = smem;
  float* Ks = Qs + (size_t)T * D;
  float* Vs = Ks + (size_t)T * D;
  float* dOs = Vs + (size_t)T * D;
  float* Ls = dOs + (size_t)T * D;
  float* Ds = Ls + (size_t)T * H;
  float* lb = Ds + (size_t)T * H;
  float* dlb = lb + T;  // [H][T]

  for (int idx = threadIdx.x; idx < T * (D / 4); idx += blockDim.x) {
    int p = idx / (D / 4), c4 = idx % (D / 4);
    const int64_t row = token_row(p, lo, n_edges, atom);
    const float* src = qkv + row * (3 * D);
    reinterpret_cast<float4*>(Qs)[idx] = __ldg(reinterpret_cast<const float4*>(src) + c4);
    reinterpret_cast<float4*>(Ks)[idx] = __ldg(reinterpret_cast<const float4*>(src + D) + c4);
    reinterpret_cast<float4*>(Vs)[idx] = __ldg(reinterpret_cast<const float4*>(src + 2 * D) + c4);
    reinterpret_cast<float4*>(dOs)[idx] =
        __ldg(reinterpret_cast<const float4*>(d_out + row * D) + c4);
  }
  for (int p = threadIdx.x; p < T; p += blockDim.x)
    lb[p] = p == 0 ? 0.f : logf(fmaxf(fc[lo + p - 1], 1e-15f));
  for (int idx = threadIdx.x; idx < T * H; idx += blockDim.x) {
    int p = idx / H;
    Ls[idx] = lse[token_row(p, lo, n_edges, atom) * H + (idx % H)];
  }
  __syncthreads();

  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- phase A: lane <-> query.  D_p = dO_p . O_p ; dQ_p = scale * sum_k dS[p,k] K_k
  for (int p = lane; p < T; p += 32) {
    const int64_t row = token_row(p, lo, n_edges, atom);
    float q[HD], dO[HD], dq[HD];
    float Dp = 0.f;
    {
      const float4* qs = reinterpret_cast<const float4*>(Qs + (size_t)p * D + h * HD);
      const float4* ds = reinterpret_cast<const float4*>(dOs + (size_t)p * D + h * HD);
      const float4* os = reinterpret_cast<const float4*>(out + row * D + h * HD);
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 a = qs[c], b = ds[c], o = __ldg(os + c);
        q[4 * c] = a.x * scale; q[4 * c + 1] = a.y * scale;
        q[4 * c + 2] = a.z * scale; q[4 * c + 3] = a.w * scale;
        dO[4 * c] = b.x; dO[4 * c + 1] = b.y; dO[4 * c + 2] = b.z; dO[4 * c + 3] = b.w;
        Dp += b.x * o.x + b.y * o.y + b.z * o.z + b.w * o.w;
      }
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) dq[c] = 0.f;
    const float L = Ls[p * H + h];
    for (int k = 0; k < T; ++k) {
      const float4* kk = reinterpret_cast<const float4*>(Ks + (size_t)k * D + h * HD);
      const float4* vv = reinterpret_cast<const float4*>(Vs + (size_t)k * D + h * HD);
      float s = lb[k], dP = 0.f;
      float kr[HD];
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 t = kk[c], v = vv[c];
        kr[4 * c] = t.x; kr[4 * c + 1] = t.y; kr[4 * c + 2] = t.z; kr[4 * c + 3] = t.w;
        s = fmaf(q[4 * c], t.x, s); s = fmaf(q[4 * c + 1], t.y, s);
        s = fmaf(q[4 * c + 2], t.z, s); s = fmaf(q[4 * c + 3], t.w, s);
        dP = fmaf(dO[4 * c], v.x, dP); dP = fmaf(dO[4 * c + 1], v.y, dP);
        dP = fmaf(dO[4 * c + 2], v.z, dP); dP = fmaf(dO[4 * c + 3], v.w, dP);
      }
      float dS = expf(s - L) * (dP - Dp);
#pragma unroll
      for (int c = 0; c < HD; ++c) dq[c] = fmaf(dS, kr[c], dq[c]);
    }
    Ds[p * H + h] = Dp;
    float4* dst = reinterpret_cast<float4*>(d_qkv + row * (3 * D) + h * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c)
      dst[c] = make_float4(dq[4 * c] * scale, dq[4 * c + 1] * scale, dq[4 * c + 2] * scale,
                           dq[4 * c + 3] * scale);
  }
  __syncthreads();  // Ds complete (each warp only reads its own head, but T may exceed 32)

  // ---- phase B: lane <-> key.  dK_k = scale * sum_p dS[p,k] Q_p ; dV_k = sum_p P[p,k] dO_p
  for (int k = lane; k < T; k += 32) {
    const int64_t row = token_row(k, lo, n_edges, atom);
    float kr[HD], vr[HD], dk[HD], dv[HD];
    {
      const float4* ks = reinterpret_cast<const float4*>(Ks + (size_t)k * D + h * HD);
      const float4* vs = reinterpret_cast<const float4*>(Vs + (size_t)k * D + h * HD);
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 a = ks[c], b = vs[c];
        kr[4 * c] = a.x * scale; kr[4 * c + 1] = a.y * scale;
        kr[4 * c + 2] = a.z * scale; kr[4 * c + 3] = a.w * scale;
        vr[4 * c] = b.x; vr[4 * c + 1] = b.y; vr[4 * c + 2] = b.z; vr[4 * c + 3] = b.w;
      }
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
    const float bias = lb[k];
    float dbias = 0.f;
    for (int p = 0; p < T; ++p) {
      const float4* qq = reinterpret_cast<const float4*>(Qs + (size_t)p * D + h * HD);
      const float4* dd = reinterpret_cast<const float4*>(dOs + (size_t)p * D + h * HD);
      float s = bias, dP = 0.f;
      float qr[HD], dr[HD];
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 a = qq[c], b = dd[c];
        qr[4 * c] = a.x; qr[4 * c + 1] = a.y; qr[4 * c + 2] = a.z; qr[4 * c + 3] = a.w;
        dr[4 * c] = b.x; dr[4 * c + 1] = b.y; dr[4 * c + 2] = b.z; dr[4 * c + 3] = b.w;
        s = fmaf(a.x, kr[4 * c], s); s = fmaf(a.y, kr[4 * c + 1], s);
        s = fmaf(a.z, kr[4 * c + 2], s); s = fmaf(a.w, kr[4 * c + 3], s);
        dP = fmaf(b.x, vr[4 * c], dP); dP = fmaf(b.y, vr[4 * c + 1], dP);
        dP = fmaf(b.z, vr[4 * c + 2], dP); dP = fmaf(b.w, vr[4 * c + 3], dP);
      }
      float P = expf(s - Ls[p * H + h]);
      float dS = P * (dP - Ds[p * H + h]);
      dbias += dS;
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        dk[c] = fmaf(dS, qr[c], dk[c]);
        dv[c] = fmaf(P, dr[c], dv[c]);
      }
    }
    dlb[h * T + k] = dbias;
    float4* dstk = reinterpret_cast<float4*>(d_qkv + row * (3 * D) + D + h * HD);
    float4* dstv = reinterpret_cast<float4*>(d_qkv + row * (3 * D) + 2 * D + h * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      dstk[c] = make_float4(dk[4 * c] * scale, dk[4 * c + 1] * scale, dk[4 * c + 2] * scale,
                            dk[4 * c + 3] * scale);
      dstv[c] = make_float4(dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
    }
  }
  __syncthreads();

  // ---- d_fc[e] += (sum_h dlb[h][k]) * d log(max(f,1e-15))/df
  if (d_fc) {
    for (int k = 1 + threadIdx.x; k < T; k += blockDim.x) {
      float acc = 0.f;
#pragma unroll
      for (int hh = 0; hh < H; ++hh) acc += dlb[hh * T + k];
      float f = fc[lo + k - 1];
      if (f >= 1e-15f) d_fc[lo + k - 1] += acc / f;
    }
  }
}

// =====================================================================================
// v2 kernels (T <= 64 tokens per atom, the common case: 26..49 for water at 4.5 A).
//  * work items (head, query) resp. (head, key) are flattened over the 128 threads of the CTA,
//    so lanes are ~83 % busy for T ~ 39 instead of 61 % with one warp per head;
//  * logits live in registers (fully unrolled key loop), one ex2.approx per (query, key): the
//    softmax runs in base 2 (q pre-scaled by scale*log2(e)), lse is stored in base-2 units;
//  * K/V (and Q/dO) rows are read from shared memory as multi-address broadcasts.
// =====================================================================================
constexpr int MAXT = 64;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int H>
__global__ void __launch_bounds__(128) attention_fwd_v2_kernel(
    const float* __restrict__ qkv, const int32_t* __restrict__ row_ptr,
    const float* __restrict__ fc, int64_t n_edges, float scale, float* __restrict__ out,
    float* __restrict__ lse) {
  constexpr int D = H * HD;
  extern __shared__ __align__(16) float smem[];
  const int64_t atom = blockIdx.x;
  const int lo = row_ptr[atom];
  const int T = row_ptr[atom + 1] - lo + 1;
  float* Ks = smem;
  float* Vs = Ks + (size_t)T * D;
  float* lb = Vs + (size_t)T * D;
  for (int idx = threadIdx.x; idx < T * (D / 4); idx += blockDim.x) {
    int p = idx / (D / 4), c4 = idx % (D / 4);
    const float* row = qkv + token_row(p, lo, n_edges, atom) * (3 * D);
    reinterpret_cast<float4*>(Ks)[idx] = __ldg(reinterpret_cast<const float4*>(row + D) + c4);
    reinterpret_cast<float4*>(Vs)[idx] = __ldg(reinterpret_cast<const float4*>(row + 2 * D) + c4);
  }
  for (int p = threadIdx.x; p < T; p += blockDim.x)
    lb[p] = p == 0 ? 0.f : kLog2e * logf(fmaxf(fc[lo + p - 1], 1e-15f));
  __syncthreads();

  const float qs = scale * kLog2e;
  for (int item = threadIdx.x; item < H * T; item += blockDim.x) {
    const int h = item / T, p = item - h * T;
    const int64_t row = token_row(p, lo, n_edges, atom);
    float q[HD];
    {
      const float4* src = reinterpret_cast<const float4*>(qkv + row * (3 * D) + h * HD);
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 t = __ldg(src + c);
        q[4 * c] = t.x * qs; q[4 * c + 1] = t.y * qs; q[4 * c + 2] = t.z * qs; q[4 * c + 3] = t.w * qs;
      }
    }
    float s[MAXT];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < MAXT; ++k) {
      if (k < T) {
        const float4* kk = reinterpret_cast<const float4*>(Ks + (size_t)k * D + h * HD);
        float a = lb[k];
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
          float4 t = kk[c];
          a = fmaf(q[4 * c], t.x, a); a = fmaf(q[4 * c + 1], t.y, a);
          a = fmaf(q[4 * c + 2], t.z, a); a = fmaf(q[4 * c + 3], t.w, a);
        }
        s[k] = a;
        m = fmaxf(m, a);
      }
    }
    float l = 0.f, o[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) o[c] = 0.f;
#pragma unroll
    for (int k = 0; k < MAXT; ++k) {
      if (k < T) {
        const float pexp = ex2_approx(s[k] - m);
        l += pexp;
        const float4* vv = reinterpret_cast<const float4*>(Vs + (size_t)k * D + h * HD);
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
          float4 t = vv[c];
          o[4 * c] = fmaf(pexp, t.x, o[4 * c]); o[4 * c + 1] = fmaf(pexp, t.y, o[4 * c + 1]);
          o[4 * c + 2] = fmaf(pexp, t.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(pexp, t.w, o[4 * c + 3]);
        }
      }
    }
    const float inv = 1.0f / l;
    float4* dst = reinterpret_cast<float4*>(out + row * D + h * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c)
      dst[c] = make_float4(o[4 * c] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv);
    lse[row * H + h] = m + log2f(l);  // base-2 units (only attention_bwd_v2 reads it)
  }
}

template <int H>
__global__ void __launch_bounds__(128) attention_bwd_v2_kernel(
    const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ lse,
    const float* __restrict__ d_out, const int32_t* __restrict__ row_ptr,
    const float* __restrict__ fc, int64_t n_edges, float scale, float* __restrict__ d_qkv,
    float* __restrict__ d_fc) {
  constexpr int D = H * HD;
  extern __shared__ __align__(16) float smem[];
  const int64_t atom = blockIdx.x;
  const int lo = row_ptr[atom];
  const int T = row_ptr[atom + 1] - lo + 1;
  float* Qs = smem;
  float* Ks = Qs + (size_t)T * D;
  float* Vs = Ks + (size_t)T * D;
  float* dOs = Vs + (size_t)T * D;
  float* Ls = dOs + (size_t)T * D;
  float* Ds = Ls + (size_t)T * H;
  float* lb = Ds + (size_t)T * H;
  float* dlb = lb + T;  // [H][T]
  for (int idx = threadIdx.x; idx < T * (D / 4); idx += blockDim.x) {
    int p = idx / (D / 4), c4 = idx % (D / 4);
    const int64_t row = token_row(p, lo, n_edges, atom);
    const float* src = qkv + row * (3 * D);
    reinterpret_cast<float4*>(Qs)[idx] = __ldg(reinterpret_cast<const float4*>(src) + c4);
    reinterpret_cast<float4*>(Ks)[idx] = __ldg(reinterpret_cast<const float4*>(src + D) + c4);
    reinterpret_cast<float4*>(Vs)[idx] = __ldg(reinterpret_cast<const float4*>(src + 2 * D) + c4);
    reinterpret_cast<float4*>(dOs)[idx] = __ldg(reinterpret_cast<const float4*>(d_out + row * D) + c4);
  }
  for (int p = threadIdx.x; p < T; p += blockDim.x)
    lb[p] = p == 0 ? 0.f : kLog2e * logf(fmaxf(fc[lo + p - 1], 1e-15f));
  for (int idx = threadIdx.x; idx < T * H; idx += blockDim.x)
    Ls[idx] = lse[token_row(idx / H, lo, n_edges, atom) * H + (idx % H)];
  __syncthreads();
  const float qs = scale * kLog2e;

  // ---- phase A: item = (head, query).  D_p = dO_p . O_p ; dQ_p = scale * sum_k dS[p,k] K_k
  for (int item = threadIdx.x; item < H * T; item += blockDim.x) {
    const int h = item / T, p = item - h * T;
    const int64_t row = token_row(p, lo, n_edges, atom);
    float q[HD], dO[HD], dq[HD];
    float Dp = 0.f;
    {
      const float4* qsrc = reinterpret_cast<const float4*>(Qs + (size_t)p * D + h * HD);
      const float4* ds = reinterpret_cast<const float4*>(dOs + (size_t)p * D + h * HD);
      const float4* os = reinterpret_cast<const float4*>(out + row * D + h * HD);
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 a = qsrc[c], b = ds[c], o = __ldg(os + c);
        q[4 * c] = a.x * qs; q[4 * c + 1] = a.y * qs; q[4 * c + 2] = a.z * qs; q[4 * c + 3] = a.w * qs;
        dO[4 * c] = b.x; dO[4 * c + 1] = b.y; dO[4 * c + 2] = b.z; dO[4 * c + 3] = b.w;
        Dp += b.x * o.x + b.y * o.y + b.z * o.z + b.w * o.w;
      }
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) dq[c] = 0.f;
    const float L = Ls[p * H + h];
    for (int k = 0; k < T; ++k) {
      const float4* kk = reinterpret_cast<const float4*>(Ks + (size_t)k * D + h * HD);
      const float4* vv = reinterpret_cast<const float4*>(Vs + (size_t)k * D + h * HD);
      float a = lb[k], dP = 0.f;
      float kr[HD];
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 t = kk[c], v = vv[c];
        kr[4 * c] = t.x; kr[4 * c + 1] = t.y; kr[4 * c + 2] = t.z; kr[4 * c + 3] = t.w;
        a = fmaf(q[4 * c], t.x, a); a = fmaf(q[4 * c + 1], t.y, a);
        a = fmaf(q[4 * c + 2], t.z, a); a = fmaf(q[4 * c + 3], t.w, a);
        dP = fmaf(dO[4 * c], v.x, dP); dP = fmaf(dO[4 * c + 1], v.y, dP);
        dP = fmaf(dO[4 * c + 2], v.z, dP); dP = fmaf(dO[4 * c + 3], v.w, dP);
      }
      const float dS = ex2_approx(a - L) * (dP - Dp);
#pragma unroll
      for (int c = 0; c < HD; ++c) dq[c] = fmaf(dS, kr[c], dq[c]);
    }
    Ds[p * H + h] = Dp;
    float4* dst = reinterpret_cast<float4*>(d_qkv + row * (3 * D) + h * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c)
      dst[c] = make_float4(dq[4 * c] * scale, dq[4 * c + 1] * scale, dq[4 * c + 2] * scale,
                           dq[4 * c + 3] * scale);
  }
  __syncthreads();

  // ---- phase B: item = (head, key).  dK_k = scale * sum_p dS[p,k] Q_p ; dV_k = sum_p P[p,k] dO_p
  for (int item = threadIdx.x; item < H * T; item += blockDim.x) {
    const int h = item / T, k = item - h * T;
    const int64_t row = token_row(k, lo, n_edges, atom);
    float kr[HD], vr[HD], dk[HD], dv[HD];
    {
      const float4* ks = reinterpret_cast<const float4*>(Ks + (size_t)k * D + h * HD);
      const float4* vs = reinterpret_cast<const float4*>(Vs + (size_t)k * D + h * HD);
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 a = ks[c], b = vs[c];
        kr[4 * c] = a.x * qs; kr[4 * c + 1] = a.y * qs; kr[4 * c + 2] = a.z * qs; kr[4 * c + 3] = a.w * qs;
        vr[4 * c] = b.x; vr[4 * c + 1] = b.y; vr[4 * c + 2] = b.z; vr[4 * c + 3] = b.w;
      }
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
    const float bias = lb[k];
    float dbias = 0.f;
    for (int p = 0; p < T; ++p) {
      const float4* qq = reinterpret_cast<const float4*>(Qs + (size_t)p * D + h * HD);
      const float4* dd = reinterpret_cast<const float4*>(dOs + (size_t)p * D + h * HD);
      float a = bias, dP = 0.f;
      float qr[HD], dr[HD];
#pragma unroll
      for (int c = 0; c < HD / 4; ++c) {
        float4 x = qq[c], b = dd[c];
        qr[4 * c] = x.x; qr[4 * c + 1] = x.y; qr[4 * c + 2] = x.z; qr[4 * c + 3] = x.w;
        dr[4 * c] = b.x; dr[4 * c + 1] = b.y; dr[4 * c + 2] = b.z; dr[4 * c + 3] = b.w;
        a = fmaf(x.x, kr[4 * c], a); a = fmaf(x.y, kr[4 * c + 1], a);
        a = fmaf(x.z, kr[4 * c + 2], a); a = fmaf(x.w, kr[4 * c + 3], a);
        dP = fmaf(b.x, vr[4 * c], dP); dP = fmaf(b.y, vr[4 * c + 1], dP);
        dP = fmaf(b.z, vr[4 * c + 2], dP); dP = fmaf(b.w, vr[4 * c + 3], dP);
      }
      const float P = ex2_approx(a - Ls[p * H + h]);
      const float dS = P * (dP - Ds[p * H + h]);
      dbias += dS;
#pragma unroll
      for (int c = 0; c < HD; ++c) {
        dk[c] = fmaf(dS, qr[c], dk[c]);
        dv[c] = fmaf(P, dr[c], dv[c]);
      }
    }
    dlb[h * T + k] = dbias;
    float4* dstk = reinterpret_cast<float4*>(d_qkv + row * (3 * D) + D + h * HD);
    float4* dstv = reinterpret_cast<float4*>(d_qkv + row * (3 * D) + 2 * D + h * HD);
#pragma unroll
    for (int c = 0; c < HD / 4; ++c) {
      dstk[c] = make_float4(dk[4 * c] * scale, dk[4 * c + 1] * scale, dk[4 * c + 2] * scale,
                            dk[4 * c + 3] * scale);
      dstv[c] = make_float4(dv[4 * c], dv[4 * c + 1], dv[4 * c + 2], dv[4 * c + 3]);
    }
  }
  __syncthreads();
  if (d_fc) {
    for (int k = 1 + threadIdx.x; k < T; k += blockDim.x) {
      float acc = 0.f;
#pragma unroll
      for (int hh = 0; hh < H; ++hh) acc += dlb[hh * T + k];
      float f = fc[lo + k - 1];
      if (f >= 1e-15f) d_fc[lo + k - 1] += acc / f;
    }
  }
}

size_t fwd_smem_bytes(int T, int H) { return sizeof(float) * ((size_t)2 * T * H * HD + T); }
size_t bwd_smem_bytes(int T, int H) {
  return sizeof(float) * ((size_t)4 * T * H * HD + (size_t)2 * T * H + T + (size_t)H * T);
}

}  // namespace
}  // namespace petb200

using namespace petb200;

extern "C" PETB200_API int petb200_attention_fwd(const float* qkv, const int32_t* row_ptr,
                                     const float* cutoff_factor, int64_t n_atoms, int64_t n_edges,
                                     int num_heads, int head_dim, float scale, int max_row,
                                     float* out, float* lse, cudaStream_t stream) {
  if (num_heads != 8 || head_dim != HD) {
    set_error("attention_fwd: only num_heads=8, head_dim=16 is built (got %d x %d)", num_heads,
              head_dim);
    return PETB200_ERR_UNSUPPORTED;
  }
  if (n_atoms == 0) return PETB200_OK;
  size_t smem = fwd_smem_bytes(max_row + 1, num_heads);
  if (smem > 227 * 1024) {
    set_error("attention_fwd: %d neighbours per atom exceed the shared-memory tile", max_row);
    return PETB200_ERR_UNSUPPORTED;
  }
  if (max_row + 1 <= MAXT) {  // fwd and bwd use the same rule: lse units must match
    auto kern = attention_fwd_v2_kernel<8>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)n_atoms, 128, smem, stream>>>(qkv, row_ptr, cutoff_factor, n_edges, scale, out,
                                                   lse);
    return check_launch("attention_fwd_v2");
  }
  auto kern = attention_fwd_kernel<8>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)n_atoms, 256, smem, stream>>>(qkv, row_ptr, cutoff_factor, n_edges, scale, out,
                                                 lse);
  return check_launch("attention_fwd");
}

extern "C" PETB200_API int petb200_attention_bwd(const float* qkv, const float* out, const float* lse,
                                     const float* d_out, const int32_t* row_ptr,
                                     const float* cutoff_factor, int64_t n_atoms, int64_t n_edges,
                                     int num_heads, int head_dim, float scale, int max_row,
                                     float* d_qkv, float* d_fc, cudaStream_t stream) {
  if (num_heads != 8 || head_dim != HD) {
    set_error("attention_bwd: only num_heads=8, head_dim=16 is built (got %d x %d)", num_heads,
              head_dim);
    return PETB200_ERR_UNSUPPORTED;
  }
  if (n_atoms == 0) return PETB200_OK;
  size_t smem = bwd_smem_bytes(max_row + 1, num_heads);
  if (smem > 227 * 1024) {
    set_error("attention_bwd: %d neighbours per atom exceed the shared-memory tile", max_row);
    return PETB200_ERR_UNSUPPORTED;
  }
  if (max_row + 1 <= MAXT) {
    auto kern = attention_bwd_v2_kernel<8>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)n_atoms, 128, smem, stream>>>(qkv, out, lse, d_out, row_ptr, cutoff_factor,
                                                   n_edges, scale, d_qkv, d_fc);
    return check_launch("attention_bwd_v2");
  }
  auto kern = attention_bwd_kernel<8>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)n_atoms, 256, smem, stream>>>(qkv, out, lse, d_out, row_ptr, cutoff_factor,
                                                 n_edges, scale, d_qkv, d_fc);
  return check_launch("attention_bwd");
}
