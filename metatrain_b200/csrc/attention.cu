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
// CUDA-core kernel, written for Blackwell's packed fp32 pipe: every 16-long dot product /
// axpy is 8 FFMA2 (fma.rn.f32x2) instead of 16 FFMA.  One CTA (128 threads) per atom; work
// items (head, query) resp. (head, key) are flattened over the threads so lanes stay ~80 %
// busy for ragged T; K/V (resp. Q/dO) rows are staged once in shared memory and read as
// multi-address broadcasts; softmax runs online in base 2 (ex2.approx, q pre-scaled by
// scale*log2(e)) over chunks of 8 keys whose logits live in registers.  The backward is two
// kernels (dQ; then dK, dV and the key-bias gradient) so each needs only half the shared
// memory -> 4 CTAs per SM.  No atomics: every output row has exactly one writer, results
// are bit-reproducible.  The log-sum-exp is stored in base-2 units.
#include "common.cuh"
#include "kernels.cuh"

namespace petb200 {
namespace {

constexpr int HD = 16;  // head dim (d_pet / num_heads = 128 / 8 with the default hypers)
constexpr float kLog2e = 1.4426950408889634f;
constexpr int KC = 8;   // keys per online-softmax chunk

__device__ __forceinline__ int64_t token_row(int p, int row_lo, int64_t n_edges, int64_t atom) {
  return p == 0 ? n_edges + atom : (int64_t)row_lo + (p - 1);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32x2 fused multiply-add (Blackwell FFMA2)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
// dot of two 16-vectors held as 8 float2 each, plus an initial value
__device__ __forceinline__ float dot16(const float2* a, const float4* b, float init) {
  float2 acc = make_float2(init, 0.f);
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    const float4 t = b[c];
    acc = ffma2(a[2 * c], make_float2(t.x, t.y), acc);
    acc = ffma2(a[2 * c + 1], make_float2(t.z, t.w), acc);
  }
  return acc.x + acc.y;
}
// y[0..15] += s * b[0..15]
__device__ __forceinline__ void axpy16(float2* y, float s, const float4* b) {
  const float2 s2 = make_float2(s, s);
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    const float4 t = b[c];
    y[2 * c] = ffma2(s2, make_float2(t.x, t.y), y[2 * c]);
    y[2 * c + 1] = ffma2(s2, make_float2(t.z, t.w), y[2 * c + 1]);
  }
}
__device__ __forceinline__ void load16(float2* dst, const float4* src, float mul) {
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    const float4 t = src[c];
    dst[2 * c] = make_float2(t.x * mul, t.y * mul);
    dst[2 * c + 1] = make_float2(t.z * mul, t.w * mul);
  }
}
__device__ __forceinline__ void store16(float4* dst, const float2* src, float mul) {
#pragma unroll
  for (int c = 0; c < HD / 4; ++c)
    dst[c] = make_float4(src[2 * c].x * mul, src[2 * c].y * mul, src[2 * c + 1].x * mul,
                         src[2 * c + 1].y * mul);
}

// stage D columns (from col0) of the atom's token rows into smem [T][D] with cp.async: all
// copies of a thread are in flight at once (no register round trip per 16 bytes)
template <int D>
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int64_t ld,
                                           int col0, int T, int lo, int64_t n_edges, int64_t atom) {
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  for (int idx = threadIdx.x; idx < T * (D / 4); idx += blockDim.x) {
    const int p = idx / (D / 4), c4 = idx % (D / 4);
    const float* row = src + token_row(p, lo, n_edges, atom) * ld + col0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + 16u * idx),
                 "l"(reinterpret_cast<const float4*>(row) + c4)
                 : "memory");
  }
}
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
}

// ------------------------------------------------------------------------- forward
// smem: Ks[T][D], Vs[T][D], lb[T]
template <int H>
__global__ void __launch_bounds__(128) attention_fwd_kernel(
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
  stage_rows<D>(Ks, qkv, 3 * D, D, T, lo, n_edges, atom);
  stage_rows<D>(Vs, qkv, 3 * D, 2 * D, T, lo, n_edges, atom);
  for (int p = threadIdx.x; p < T; p += blockDim.x)
    lb[p] = p == 0 ? 0.f : kLog2e * logf(fmaxf(fc[lo + p - 1], 1e-15f));

  const float qs = scale * kLog2e;
  // the query of the NEXT item is fetched while the current one is processed
  float2 q_next[HD / 2];
  auto fetch = [&](int item) {
    const int h = item / T, p = item - h * T;
    load16(q_next, reinterpret_cast<const float4*>(
                       qkv + token_row(p, lo, n_edges, atom) * (3 * D) + h * HD), qs);
  };
  if ((int)threadIdx.x < H * T) fetch(threadIdx.x);
  stage_wait();
  for (int item = threadIdx.x; item < H * T; item += blockDim.x) {
    const int h = item / T, p = item - h * T;
    const int64_t row = token_row(p, lo, n_edges, atom);
    float2 q[HD / 2], o[HD / 2];
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) q[c] = q_next[c];
    if (item + (int)blockDim.x < H * T) fetch(item + blockDim.x);
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) o[c] = make_float2(0.f, 0.f);
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < T; k0 += KC) {
      float s[KC];
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < KC; ++j) {
        const int k = min(k0 + j, T - 1);
        const float v = dot16(q, reinterpret_cast<const float4*>(Ks + (size_t)k * D + h * HD), lb[k]);
        s[j] = (k0 + j < T) ? v : -INFINITY;
        cm = fmaxf(cm, s[j]);
      }
      const float m_new = fmaxf(m, cm);
      const float corr = ex2_approx(m - m_new);  // 0 on the first chunk
      l *= corr;
      const float2 c2 = make_float2(corr, corr), z2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < HD / 2; ++c) o[c] = ffma2(o[c], c2, z2);
#pragma unroll
      for (int j = 0; j < KC; ++j) {
        const int k = min(k0 + j, T - 1);
        const float pexp = ex2_approx(s[j] - m_new);  // exactly 0 for masked keys
        l += pexp;
        axpy16(o, pexp, reinterpret_cast<const float4*>(Vs + (size_t)k * D + h * HD));
      }
      m = m_new;
    }
    store16(reinterpret_cast<float4*>(out + row * D + h * HD), o, 1.0f / l);
    lse[row * H + h] = m + log2f(l);  // base-2 units (read only by the backward kernels)
  }
}

// ------------------------------------------------------------------------ backward
// kernel A: item = (head, query).  D_p = dO_p . O_p ; dQ_p = scale * sum_k dS[p,k] K_k
// smem: Ks[T][D], Vs[T][D], lb[T]
template <int H>
__global__ void __launch_bounds__(128) attention_bwd_dq_kernel(
    const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ lse,
    const float* __restrict__ d_out, const int32_t* __restrict__ row_ptr,
    const float* __restrict__ fc, int64_t n_edges, float scale, float* __restrict__ d_qkv,
    float* __restrict__ dsum) {
  constexpr int D = H * HD;
  extern __shared__ __align__(16) float smem[];
  const int64_t atom = blockIdx.x;
  const int lo = row_ptr[atom];
  const int T = row_ptr[atom + 1] - lo + 1;
  float* Ks = smem;
  float* Vs = Ks + (size_t)T * D;
  float* lb = Vs + (size_t)T * D;
  stage_rows<D>(Ks, qkv, 3 * D, D, T, lo, n_edges, atom);
  stage_rows<D>(Vs, qkv, 3 * D, 2 * D, T, lo, n_edges, atom);
  for (int p = threadIdx.x; p < T; p += blockDim.x)
    lb[p] = p == 0 ? 0.f : kLog2e * logf(fmaxf(fc[lo + p - 1], 1e-15f));
  const float qs = scale * kLog2e;
  float2 q_next[HD / 2], dO_next[HD / 2];
  float Dp_next = 0.f, L_next = 0.f;
  auto fetch = [&](int item) {
    const int h = item / T, p = item - h * T;
    const int64_t row = token_row(p, lo, n_edges, atom);
    load16(q_next, reinterpret_cast<const float4*>(qkv + row * (3 * D) + h * HD), qs);
    load16(dO_next, reinterpret_cast<const float4*>(d_out + row * D + h * HD), 1.0f);
    Dp_next = dot16(dO_next, reinterpret_cast<const float4*>(out + row * D + h * HD), 0.f);
    L_next = lse[row * H + h];
  };
  if ((int)threadIdx.x < H * T) fetch(threadIdx.x);
  stage_wait();
  for (int item = threadIdx.x; item < H * T; item += blockDim.x) {
    const int h = item / T, p = item - h * T;
    const int64_t row = token_row(p, lo, n_edges, atom);
    float2 q[HD / 2], dO[HD / 2], dq[HD / 2];
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) {
      q[c] = q_next[c];
      dO[c] = dO_next[c];
    }
    const float Dp = Dp_next, L = L_next;
    if (item + (int)blockDim.x < H * T) fetch(item + blockDim.x);
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) dq[c] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < T; ++k) {
      const float4* kk = reinterpret_cast<const float4*>(Ks + (size_t)k * D + h * HD);
      const float4* vv = reinterpret_cast<const float4*>(Vs + (size_t)k * D + h * HD);
      const float a = dot16(q, kk, lb[k]);
      const float dP = dot16(dO, vv, 0.f);
      const float dS = ex2_approx(a - L) * (dP - Dp);
      axpy16(dq, dS, kk);
    }
    dsum[row * H + h] = Dp;
    store16(reinterpret_cast<float4*>(d_qkv + row * (3 * D) + h * HD), dq, scale);
  }
}

// kernel B: item = (head, key).  dK_k = scale * sum_p dS[p,k] Q_p ; dV_k = sum_p P[p,k] dO_p ;
// d_fc[e] += (sum_{heads,queries} dS[.,e]) / f_e.   smem: Qs, dOs [T][D]; Ls, Ds [T][H]; dlb [H][T]
template <int H>
__global__ void __launch_bounds__(128) attention_bwd_dkv_kernel(
    const float* __restrict__ qkv, const float* __restrict__ lse, const float* __restrict__ dsum,
    const float* __restrict__ d_out, const int32_t* __restrict__ row_ptr,
    const float* __restrict__ fc, int64_t n_edges, float scale, float* __restrict__ d_qkv,
    float* __restrict__ d_fc) {
  constexpr int D = H * HD;
  extern __shared__ __align__(16) float smem[];
  const int64_t atom = blockIdx.x;
  const int lo = row_ptr[atom];
  const int T = row_ptr[atom + 1] - lo + 1;
  float* Qs = smem;
  float* dOs = Qs + (size_t)T * D;
  float* Ls = dOs + (size_t)T * D;
  float* Ds = Ls + (size_t)T * H;
  float* dlb = Ds + (size_t)T * H;  // [H][T]
  stage_rows<D>(Qs, qkv, 3 * D, 0, T, lo, n_edges, atom);
  stage_rows<D>(dOs, d_out, D, 0, T, lo, n_edges, atom);
  for (int idx = threadIdx.x; idx < T * H; idx += blockDim.x) {
    const int64_t row = token_row(idx / H, lo, n_edges, atom);
    Ls[idx] = lse[row * H + (idx % H)];
    Ds[idx] = dsum[row * H + (idx % H)];
  }
  const float qs = scale * kLog2e;
  float2 k_next[HD / 2], v_next[HD / 2];
  auto fetch = [&](int item) {
    const int h = item / T, k = item - h * T;
    const int64_t row = token_row(k, lo, n_edges, atom);
    load16(k_next, reinterpret_cast<const float4*>(qkv + row * (3 * D) + D + h * HD), qs);
    load16(v_next, reinterpret_cast<const float4*>(qkv + row * (3 * D) + 2 * D + h * HD), 1.0f);
  };
  if ((int)threadIdx.x < H * T) fetch(threadIdx.x);
  stage_wait();
  for (int item = threadIdx.x; item < H * T; item += blockDim.x) {
    const int h = item / T, k = item - h * T;
    const int64_t row = token_row(k, lo, n_edges, atom);
    float2 kr[HD / 2], vr[HD / 2], dk[HD / 2], dv[HD / 2];
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) {
      kr[c] = k_next[c];
      vr[c] = v_next[c];
    }
    if (item + (int)blockDim.x < H * T) fetch(item + blockDim.x);
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) {
      dk[c] = make_float2(0.f, 0.f);
      dv[c] = make_float2(0.f, 0.f);
    }
    const float bias = k == 0 ? 0.f : kLog2e * logf(fmaxf(fc[lo + k - 1], 1e-15f));
    float dbias = 0.f;
#pragma unroll 4
    for (int p = 0; p < T; ++p) {
      const float4* qq = reinterpret_cast<const float4*>(Qs + (size_t)p * D + h * HD);
      const float4* dd = reinterpret_cast<const float4*>(dOs + (size_t)p * D + h * HD);
      const float a = dot16(kr, qq, bias);
      const float dP = dot16(vr, dd, 0.f);
      const float P = ex2_approx(a - Ls[p * H + h]);
      const float dS = P * (dP - Ds[p * H + h]);
      dbias += dS;
      axpy16(dk, dS, qq);
      axpy16(dv, P, dd);
    }
    dlb[h * T + k] = dbias;
    store16(reinterpret_cast<float4*>(d_qkv + row * (3 * D) + D + h * HD), dk, scale);
    store16(reinterpret_cast<float4*>(d_qkv + row * (3 * D) + 2 * D + h * HD), dv, 1.0f);
  }
  __syncthreads();
  if (d_fc) {
    for (int k = 1 + threadIdx.x; k < T; k += blockDim.x) {
      float acc = 0.f;
#pragma unroll
      for (int hh = 0; hh < H; ++hh) acc += dlb[hh * T + k];
      const float f = fc[lo + k - 1];
      if (f >= 1e-15f) d_fc[lo + k - 1] += acc / f;
    }
  }
}

size_t fwd_smem_bytes(int T, int H) { return sizeof(float) * ((size_t)2 * T * H * HD + T); }
size_t dkv_smem_bytes(int T, int H) {
  return sizeof(float) * ((size_t)2 * T * H * HD + (size_t)2 * T * H + (size_t)H * T);
}

int check_shape(const char* what, int num_heads, int head_dim, size_t smem, int max_row) {
  if (num_heads != 8 || head_dim != HD) {
    set_error("%s: only num_heads=8, head_dim=16 is built (got %d x %d)", what, num_heads, head_dim);
    return PETB200_ERR_UNSUPPORTED;
  }
  if (smem > 227 * 1024) {
    set_error("%s: %d neighbours per atom exceed the shared-memory tile", what, max_row);
    return PETB200_ERR_UNSUPPORTED;
  }
  return PETB200_OK;
}

}  // namespace
}  // namespace petb200

using namespace petb200;

extern "C" PETB200_API int petb200_attention_fwd(const float* qkv, const int32_t* row_ptr,
                                     const float* cutoff_factor, int64_t n_atoms, int64_t n_edges,
                                     int num_heads, int head_dim, float scale, int max_row,
                                     int precision, float* out, float* lse, cudaStream_t stream) {
  if (precision != PETB200_PREC_FP32 && n_atoms > 0 && attention_tc_supports(num_heads, head_dim, max_row))
    return launch_attention_fwd_tc(qkv, row_ptr, cutoff_factor, n_atoms, n_edges, scale, max_row, out,
                                   lse, stream);
  const size_t smem = fwd_smem_bytes(max_row + 1, 8);
  if (int rc = check_shape("attention_fwd", num_heads, head_dim, smem, max_row)) return rc;
  if (n_atoms == 0) return PETB200_OK;
  auto kern = attention_fwd_kernel<8>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)n_atoms, 128, smem, stream>>>(qkv, row_ptr, cutoff_factor, n_edges, scale, out, lse);
  return check_launch("attention_fwd");
}

extern "C" PETB200_API int petb200_attention_bwd(const float* qkv, const float* out, const float* lse,
                                     const float* d_out, const int32_t* row_ptr,
                                     const float* cutoff_factor, int64_t n_atoms, int64_t n_edges,
                                     int num_heads, int head_dim, float scale, int max_row,
                                     int precision, float* d_qkv, float* d_fc, float* dsum,
                                     cudaStream_t stream) {
  if (precision != PETB200_PREC_FP32 && n_atoms > 0 && attention_tc_supports(num_heads, head_dim, max_row))
    return launch_attention_bwd_tc(qkv, out, lse, d_out, row_ptr, cutoff_factor, n_atoms, n_edges, scale,
                                   max_row, d_qkv, d_fc, dsum, stream);
  const size_t smem_a = fwd_smem_bytes(max_row + 1, 8), smem_b = dkv_smem_bytes(max_row + 1, 8);
  if (int rc = check_shape("attention_bwd", num_heads, head_dim, smem_b > smem_a ? smem_b : smem_a, max_row))
    return rc;
  if (n_atoms == 0) return PETB200_OK;
  auto ka = attention_bwd_dq_kernel<8>;
  auto kb = attention_bwd_dkv_kernel<8>;
  cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
  cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
  ka<<<(unsigned)n_atoms, 128, smem_a, stream>>>(qkv, out, lse, d_out, row_ptr, cutoff_factor, n_edges,
                                                 scale, d_qkv, dsum);
  kb<<<(unsigned)n_atoms, 128, smem_b, stream>>>(qkv, lse, dsum, d_out, row_ptr, cutoff_factor, n_edges,
                                                 scale, d_qkv, d_fc);
  return check_launch("attention_bwd");
}
