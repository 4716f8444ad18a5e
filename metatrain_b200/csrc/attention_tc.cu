// Tensor-core per-atom attention (forward and backward) for ragged token sets with
// T = n_i + 1 <= 64 tokens per atom and head_dim = 16.
//
// Same mathematics as attention.cu (AttentionBlock.forward / manual_attention,
// src/metatrain/pet/modules/transformer.py:86-152, 565-589), but every product
// (Q K^T, P V; in the backward also dO V^T, P^T dO, dS^T Q, dS K) runs on the tensor cores
// with warp-level mma.sync.m16n8k16 bf16 instructions and fp32 accumulation.  The problem
// per (atom, head) is tiny (<= 64 x 64 x 16), far below a tcgen05 128-row tile, so the
// warp-level MMA is the right granularity: one warp owns one (atom, head), one CTA (8
// warps) owns one atom.  Operands are split on the fly into bf16 hi + lo (x = hi + lo +
// O(2^-17 x)) and every product is formed as lo*hi + hi*lo + hi*hi, the same 2-term scheme
// as the tcgen05 GEMM (gemm_tc.cu), so the attention stays within the 1e-4 eV/A force bar.
//
// Fragment conventions (PTX ISA, mma.m16n8k16; g = lane / 4, t = lane % 4):
//   A (16 x 16, row):  a0 = (row g,   k 2t..2t+1)   a1 = (row g+8, k 2t..2t+1)
//                      a2 = (row g,   k 2t+8..2t+9) a3 = (row g+8, k 2t+8..2t+9)
//   B (16 x 8, col):   b0 = (k 2t..2t+1, n g)       b1 = (k 2t+8..2t+9, n g)
//   C (16 x 8):        c0,c1 = (row g, n 2t..2t+1)  c2,c3 = (row g+8, n 2t..2t+1)
// The head dimension is contracted in a permuted order so that one float4 global load
// (dims 4t..4t+3 of a token) is exactly one thread's share of a fragment: MMA k position
// 2t+u <-> dim 4t+u and 2t+8+u <-> dim 4t+2+u (u = 0, 1).  Output tiles come out with
// n 2t+u <-> dim 4t+u (first n-tile) and 4t+2+u (second), i.e. again one float4 per row.
// Transposed operands are produced with movmatrix (registers) or ldmatrix.trans (smem).
//
// No atomics: every output row has one writer; results are bit-reproducible.
#include <cuda_bf16.h>

#include "common.cuh"

namespace petb200 {
namespace {

constexpr int H = 8, HD = 16, D = H * HD;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ int64_t token_row(int p, int row_lo, int64_t n_edges, int64_t atom) {
  return p == 0 ? n_edges + atom : (int64_t)row_lo + (p - 1);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);  // .x = a in the low half
  return *reinterpret_cast<uint32_t*>(&v);
}
// (a, b) -> packed bf16 hi pair and packed bf16 lo pair (residuals)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);
  lo = pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  // (not volatile: independent accumulator chains may be interleaved by the scheduler)
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// c += (A_hi + A_lo) . (B_hi + B_lo) without the lo*lo term; small terms first
__device__ __forceinline__ void mma_x3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                       uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma16816(c, al[0], al[1], al[2], al[3], bh0, bh1);
  mma16816(c, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
  mma16816(c, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}
// transpose of an 8x8 b16 matrix held one 32-bit register per lane (row g, cols 2t..2t+1)
__device__ __forceinline__ uint32_t movm(uint32_t x) {
  uint32_t y;
  asm("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float key_bias(int k, int T, int lo, const float* __restrict__ fc) {
  if (k == 0) return 0.f;
  if (k >= T) return -INFINITY;
  return log2f(fmaxf(__ldg(fc + lo + k - 1), 1e-15f));
}

// L2 prefetch of the token rows of a later atom (one whose CTA starts when the CTAs resident
// now retire): the edge-token rows of an atom are contiguous, so one bulk prefetch per array
// covers them.  Decouples DRAM traffic from the (register-limited) occupancy of these kernels.
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// ------------------------------------------------------------------------- forward
// Per warp (= head) shared memory: K fragments [2*NKB][32] uint4 (hi0, hi1, lo0, lo1 per
// 8-key tile), V^T fragments [2*NKB][32] uint4, key bias [16*NKB] floats.  These are the
// warp's own registers parked in smem (each lane reads back what it wrote).
template <int NKB>
__global__ void __launch_bounds__(256, 3) attention_fwd_tc_kernel(
    const float* __restrict__ qkv, const int32_t* __restrict__ row_ptr,
    const float* __restrict__ fc, int64_t n_atoms, int64_t n_edges, float scale, float* __restrict__ out,
    float* __restrict__ lse) {
  extern __shared__ uint4 smem4[];
  constexpr int WARP_U4 = 2 * (2 * NKB * 32) + 4 * NKB;  // uint4 units per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int h = warp;
  uint4* kfrag = smem4 + (size_t)warp * WARP_U4;
  uint4* vfrag = kfrag + 2 * NKB * 32;
  float* lbias = reinterpret_cast<float*>(vfrag + 2 * NKB * 32);
  // Persistent CTAs: warp h of CTA b walks the atoms b, b + grid, ... (the eight heads of an atom run
  // side by side and share its rows in L1 / L2, but never synchronise).  The CSR bounds of the next
  // atom are loaded, and its rows pulled into L2, one atom ahead.
  int64_t atom = blockIdx.x;
  int lo_next = atom < n_atoms ? __ldg(row_ptr + atom) : 0;
  int hi_next = atom < n_atoms ? __ldg(row_ptr + atom + 1) : 0;
  for (; atom < n_atoms; atom += gridDim.x) {
  const int lo = lo_next;
  const int T = hi_next - lo + 1;
  const int nkb = (T + 15) >> 4;
  {
    const int64_t nxt = atom + gridDim.x;
    if (nxt < n_atoms) {
      lo_next = __ldg(row_ptr + nxt);
      hi_next = __ldg(row_ptr + nxt + 1);
      if (threadIdx.x == 0) {
        if (hi_next > lo_next) prefetch_l2(qkv + (int64_t)lo_next * (3 * D), (uint32_t)(hi_next - lo_next) * 3 * D * 4u);
        prefetch_l2(qkv + (n_edges + nxt) * (3 * D), 3 * D * 4u);
      }
    }
  }

  // all K / V rows of the atom are requested before the first conversion (one exposed memory
  // latency per atom instead of one per 16-key tile)
  float4 k4[NKB][2], v4[NKB][2];
#pragma unroll
  for (int jb = 0; jb < NKB; ++jb) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int key = 16 * jb + 8 * u + g;
      const bool ok = jb < nkb && key < T;
      const float* row = qkv + token_row(ok ? key : 0, lo, n_edges, atom) * (3 * D) + h * HD + 4 * t;
      k4[jb][u] = ok ? ldg4(row + D) : make_float4(0.f, 0.f, 0.f, 0.f);
      v4[jb][u] = ok ? ldg4(row + 2 * D) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int jb = 0; jb < NKB; ++jb) {
    if (jb < nkb) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = 2 * jb + u;
        uint4 kf, vf;
        split2(k4[jb][u].x, k4[jb][u].y, kf.x, kf.z);
        split2(k4[jb][u].z, k4[jb][u].w, kf.y, kf.w);
        split2(v4[jb][u].x, v4[jb][u].y, vf.x, vf.z);
        split2(v4[jb][u].z, v4[jb][u].w, vf.y, vf.w);
        vf.x = movm(vf.x);
        vf.y = movm(vf.y);
        vf.z = movm(vf.z);
        vf.w = movm(vf.w);
        kfrag[j * 32 + lane] = kf;
        vfrag[j * 32 + lane] = vf;
      }
    }
  }
  for (int k = lane; k < 16 * nkb; k += 32) lbias[k] = key_bias(k, T, lo, fc);
  __syncwarp();

  const float qs = scale * kLog2e;
  // the queries of the next tile are fetched while the current one is processed
  float4 qn0, qn1;
  auto fetch_q = [&](int qt) {
    const int p0 = 16 * qt + g, p1 = p0 + 8;
    qn0 = ldg4(qkv + token_row(p0 < T ? p0 : 0, lo, n_edges, atom) * (3 * D) + h * HD + 4 * t);
    qn1 = ldg4(qkv + token_row(p1 < T ? p1 : 0, lo, n_edges, atom) * (3 * D) + h * HD + 4 * t);
  };
  fetch_q(0);
  for (int qt = 0; qt < nkb; ++qt) {
    const int p0 = 16 * qt + g, p1 = p0 + 8;
    const bool ok0 = p0 < T, ok1 = p1 < T;
    const int64_t r0 = token_row(ok0 ? p0 : 0, lo, n_edges, atom);
    const int64_t r1 = token_row(ok1 ? p1 : 0, lo, n_edges, atom);
    const float4 q0 = qn0, q1 = qn1;
    if (qt + 1 < nkb) fetch_q(qt + 1);
    uint32_t ah[4], al[4];
    split2(q0.x * qs, q0.y * qs, ah[0], al[0]);
    split2(q1.x * qs, q1.y * qs, ah[1], al[1]);
    split2(q0.z * qs, q0.w * qs, ah[2], al[2]);
    split2(q1.z * qs, q1.w * qs, ah[3], al[3]);

    float s[2 * NKB][4];
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2 * NKB; ++j) {
      if (j < 2 * nkb) {
        const float2 b = *reinterpret_cast<const float2*>(lbias + 8 * j + 2 * t);
        s[j][0] = b.x; s[j][1] = b.y; s[j][2] = b.x; s[j][3] = b.y;
        const uint4 kf = kfrag[j * 32 + lane];
        mma_x3(s[j], ah, al, kf.x, kf.y, kf.z, kf.w);
      } else {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = -INFINITY;
      }
      m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
      m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
    }
    m0 = quad_max(m0);
    m1 = quad_max(m1);
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < 2 * NKB; ++j) {
      s[j][0] = ex2_approx(s[j][0] - m0);
      s[j][1] = ex2_approx(s[j][1] - m0);
      s[j][2] = ex2_approx(s[j][2] - m1);
      s[j][3] = ex2_approx(s[j][3] - m1);
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
      if (kb < nkb) {
        uint32_t ph[4], pl[4];
        split2(s[2 * kb][0], s[2 * kb][1], ph[0], pl[0]);
        split2(s[2 * kb][2], s[2 * kb][3], ph[1], pl[1]);
        split2(s[2 * kb + 1][0], s[2 * kb + 1][1], ph[2], pl[2]);
        split2(s[2 * kb + 1][2], s[2 * kb + 1][3], ph[3], pl[3]);
        const uint4 v0 = vfrag[(2 * kb) * 32 + lane], v1 = vfrag[(2 * kb + 1) * 32 + lane];
        mma_x3(o[0], ph, pl, v0.x, v1.x, v0.z, v1.z);
        mma_x3(o[1], ph, pl, v0.y, v1.y, v0.w, v1.w);
      }
    }
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    if (ok0) {
      *reinterpret_cast<float4*>(out + r0 * D + h * HD + 4 * t) =
          make_float4(o[0][0] * i0, o[0][1] * i0, o[1][0] * i0, o[1][1] * i0);
      if (t == 0) lse[r0 * H + h] = m0 + log2f(l0);
    }
    if (ok1) {
      *reinterpret_cast<float4*>(out + r1 * D + h * HD + 4 * t) =
          make_float4(o[0][2] * i1, o[0][3] * i1, o[1][2] * i1, o[1][3] * i1);
      if (t == 0) lse[r1 * H + h] = m1 + log2f(l1);
    }
  }
    __syncwarp();   // the warp's fragments are rewritten by its next atom
  }
}

// ------------------------------------------------------------------------ backward
// One warp per (atom, head).  Outer loop: 16-key tiles (K, V fragments and the dK, dV
// accumulators live in registers); inner loop: 16-query blocks whose Q (pre-scaled) and dO
// rows were staged once per warp in shared memory as bf16 hi / lo in two 8-column halves
// (half 0 = dims {4t, 4t+1}, half 1 = dims {4t+2, 4t+3}), so that ldmatrix returns the
// B fragments of S^T = K Q^T and dP^T = V dO^T, and ldmatrix.trans the B fragments of
// dV = P^T dO and dK = dS^T Q.  dQ = dS K accumulates in registers over the key tiles (dS
// is transposed in registers with movmatrix).
//   P^T[k,q] = 2^(S^T[k,q] - L_q),  dS^T = P^T (dP^T - D_q),  D_q = dO_q . O_q,
//   d_fc[e] += sum_{h,q} dS[q,e] / f_e.
// Per-warp smem: 4 arrays (Qh, Ql, dOh, dOl) x 2 halves x Tp rows x 16 B; L, Dq [Tp] floats.
template <int NKB>
__global__ void __launch_bounds__(256, 2) attention_bwd_tc_kernel(
    const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ lse,
    const float* __restrict__ d_out, const int32_t* __restrict__ row_ptr,
    const float* __restrict__ fc, int64_t n_atoms, int64_t n_edges, float scale, float* __restrict__ d_qkv,
    float* __restrict__ d_bias_part) {
  constexpr int Tp = 16 * NKB;
  constexpr int ARR = 2 * Tp * 16;                   // bytes of one staged array
  constexpr int WARP_BYTES = 4 * ARR + 3 * Tp * 4;   // + L, Dq, key bias
  extern __shared__ uint4 smem4[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(smem4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int h = warp;
  uint8_t* wbase = smem + (size_t)warp * WARP_BYTES;
  const uint32_t w_u32 = static_cast<uint32_t>(__cvta_generic_to_shared(wbase));
  float* Ls = reinterpret_cast<float*>(wbase + 4 * ARR);
  float* Dq = Ls + Tp;
  float* Kb = Dq + Tp;
  const float qs = scale * kLog2e;
  const int64_t n_tokens = n_edges + n_atoms;
  // Persistent CTAs, warp-independent items (see the forward kernel).  The key-bias gradient is
  // written per head ([H][tokens] scratch) and reduced over the heads by attention_dfc_kernel, so
  // the heads of an atom never synchronise.
  int64_t atom = blockIdx.x;
  int lo_next = atom < n_atoms ? __ldg(row_ptr + atom) : 0;
  int hi_next = atom < n_atoms ? __ldg(row_ptr + atom + 1) : 0;
  for (; atom < n_atoms; atom += gridDim.x) {
  const int lo = lo_next;
  const int T = hi_next - lo + 1;
  const int nkb = (T + 15) >> 4;
  {
    const int64_t nxt = atom + gridDim.x;
    if (nxt < n_atoms) {
      lo_next = __ldg(row_ptr + nxt);
      hi_next = __ldg(row_ptr + nxt + 1);
      if (threadIdx.x < 3) {
        const float* base = threadIdx.x == 0 ? qkv : (threadIdx.x == 1 ? d_out : out);
        const int ld = threadIdx.x == 0 ? 3 * D : D;
        if (hi_next > lo_next) prefetch_l2(base + (int64_t)lo_next * ld, (uint32_t)(hi_next - lo_next) * ld * 4u);
        prefetch_l2(base + (n_edges + nxt) * ld, (uint32_t)ld * 4u);
      }
    }
  }

  // ---- stage Q (scaled) and dO, L and D for every token of this (atom, head).  All global rows of
  // the atom are requested before the first conversion: one exposed memory latency per atom instead
  // of one per 16-token block (nothing else is live in registers yet).
  {
    float4 q4[NKB][2], g4[NKB][2], o4[NKB][2];
    float lq[NKB][2];
#pragma unroll
    for (int ib = 0; ib < NKB; ++ib) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int p = 16 * ib + 8 * u + g;
        const bool ok = ib < nkb && p < T;
        const int64_t row = token_row(ok ? p : 0, lo, n_edges, atom);
        q4[ib][u] = g4[ib][u] = o4[ib][u] = make_float4(0.f, 0.f, 0.f, 0.f);
        lq[ib][u] = INFINITY;
        if (ok) {
          q4[ib][u] = ldg4(qkv + row * (3 * D) + h * HD + 4 * t);
          g4[ib][u] = ldg4(d_out + row * D + h * HD + 4 * t);
          o4[ib][u] = ldg4(out + row * D + h * HD + 4 * t);
          if (t == 0) lq[ib][u] = __ldg(lse + row * H + h);
        }
      }
    }
    // key bias of every key of the atom (the cutoff factors are shared by the heads; each warp keeps
    // its own copy so that no block-level barrier is needed)
    for (int k = lane; k < 16 * nkb; k += 32) Kb[k] = key_bias(k, T, lo, fc);
#pragma unroll
    for (int ib = 0; ib < NKB; ++ib) {
      if (ib < nkb) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int p = 16 * ib + 8 * u + g;
          uint32_t h0, l0, h1, l1;
          split2(q4[ib][u].x * qs, q4[ib][u].y * qs, h0, l0);
          split2(q4[ib][u].z * qs, q4[ib][u].w * qs, h1, l1);
          uint32_t* dst = reinterpret_cast<uint32_t*>(wbase) + p * 4 + t;  // 16 B rows, word t
          dst[0] = h0;
          dst[Tp * 4] = h1;                      // half 1
          dst[ARR / 4] = l0;
          dst[ARR / 4 + Tp * 4] = l1;
          split2(g4[ib][u].x, g4[ib][u].y, h0, l0);
          split2(g4[ib][u].z, g4[ib][u].w, h1, l1);
          dst[2 * (ARR / 4)] = h0;
          dst[2 * (ARR / 4) + Tp * 4] = h1;
          dst[3 * (ARR / 4)] = l0;
          dst[3 * (ARR / 4) + Tp * 4] = l1;
          const float dsum = quad_sum(g4[ib][u].x * o4[ib][u].x + g4[ib][u].y * o4[ib][u].y +
                                      g4[ib][u].z * o4[ib][u].z + g4[ib][u].w * o4[ib][u].w);
          if (t == 0) {
            Dq[p] = dsum;
            Ls[p] = lq[ib][u];
          }
        }
      }
    }
  }
  __syncwarp();

  float dq[NKB][2][4];
#pragma unroll
  for (int qb = 0; qb < NKB; ++qb)
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) dq[qb][n][c] = 0.f;

  // ldmatrix lane addressing: lanes 8i..8i+7 supply the rows of matrix i
  const int mi = lane >> 3, mr = lane & 7;
  // non-transposed: matrices (tokens 0-7, half 0), (0-7, half 1), (8-15, half 0), (8-15, half 1)
  const uint32_t off_n = (uint32_t)(((mi & 1) * Tp + (mi >> 1) * 8 + mr) * 16);
  // transposed:     matrices (tokens 0-7, half 0), (8-15, half 0), (0-7, half 1), (8-15, half 1)
  const uint32_t off_t = (uint32_t)(((mi >> 1) * Tp + (mi & 1) * 8 + mr) * 16);

  // K / V rows of the next key tile are fetched while the current tile is processed
  float4 ka, kb4, va, vb;
  auto fetch_kv = [&](int kt) {
    const int k0 = 16 * kt + g, k1 = k0 + 8;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* r0 = qkv + token_row(k0 < T ? k0 : 0, lo, n_edges, atom) * (3 * D) + h * HD + 4 * t;
    const float* r1 = qkv + token_row(k1 < T ? k1 : 0, lo, n_edges, atom) * (3 * D) + h * HD + 4 * t;
    ka = k0 < T ? ldg4(r0 + D) : z4;
    va = k0 < T ? ldg4(r0 + 2 * D) : z4;
    kb4 = k1 < T ? ldg4(r1 + D) : z4;
    vb = k1 < T ? ldg4(r1 + 2 * D) : z4;
  };
  fetch_kv(0);
#pragma unroll 1
  for (int kt = 0; kt < nkb; ++kt) {
    const int k0 = 16 * kt + g, k1 = k0 + 8;
    const bool ok0 = k0 < T, ok1 = k1 < T;
    const int64_t r0 = token_row(ok0 ? k0 : 0, lo, n_edges, atom);
    const int64_t r1 = token_row(ok1 ? k1 : 0, lo, n_edges, atom);
    uint32_t kh[4], kl[4], vh[4], vl[4];
    split2(ka.x, ka.y, kh[0], kl[0]);
    split2(kb4.x, kb4.y, kh[1], kl[1]);
    split2(ka.z, ka.w, kh[2], kl[2]);
    split2(kb4.z, kb4.w, kh[3], kl[3]);
    split2(va.x, va.y, vh[0], vl[0]);
    split2(vb.x, vb.y, vh[1], vl[1]);
    split2(va.z, va.w, vh[2], vl[2]);
    split2(vb.z, vb.w, vh[3], vl[3]);
    // K^T fragments (B operand of dQ = dS K): n-tile 0 <- halves 0 of keys 0-7 / 8-15, ...
    uint32_t kth[4], ktl[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      kth[c] = movm(kh[c]);
      ktl[c] = movm(kl[c]);
    }
    if (kt + 1 < nkb) fetch_kv(kt + 1);
    const float lb0 = Kb[k0], lb1 = Kb[k1];
    float dk[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float dv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float db0 = 0.f, db1 = 0.f;

#pragma unroll
    for (int qb = 0; qb < NKB; ++qb) {
      if (qb < nkb) {
        const uint32_t blk = w_u32 + (uint32_t)(qb * 16 * 16);
        uint32_t qh[4], ql[4], gh[4], gl[4];
        ldsm4(qh, blk + off_n);
        ldsm4(ql, blk + ARR + off_n);
        ldsm4(gh, blk + 2 * ARR + off_n);
        ldsm4(gl, blk + 3 * ARR + off_n);
        float st[2][4], dp[2][4];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          st[n][0] = lb0; st[n][1] = lb0; st[n][2] = lb1; st[n][3] = lb1;
          dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
          mma_x3(st[n], kh, kl, qh[2 * n], qh[2 * n + 1], ql[2 * n], ql[2 * n + 1]);
          mma_x3(dp[n], vh, vl, gh[2 * n], gh[2 * n + 1], gl[2 * n], gl[2 * n + 1]);
        }
        uint32_t pth[4], ptl[4], sth[4], stl[4];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int q = 16 * qb + 8 * n + 2 * t;
          const float2 L2 = *reinterpret_cast<const float2*>(Ls + q);
          const float2 D2 = *reinterpret_cast<const float2*>(Dq + q);
          const float p0 = ex2_approx(st[n][0] - L2.x), p1 = ex2_approx(st[n][1] - L2.y);
          const float p2 = ex2_approx(st[n][2] - L2.x), p3 = ex2_approx(st[n][3] - L2.y);
          const float s0 = p0 * (dp[n][0] - D2.x), s1 = p1 * (dp[n][1] - D2.y);
          const float s2 = p2 * (dp[n][2] - D2.x), s3 = p3 * (dp[n][3] - D2.y);
          db0 += s0 + s1;
          db1 += s2 + s3;
          split2(p0, p1, pth[2 * n], ptl[2 * n]);
          split2(p2, p3, pth[2 * n + 1], ptl[2 * n + 1]);
          split2(s0, s1, sth[2 * n], stl[2 * n]);
          split2(s2, s3, sth[2 * n + 1], stl[2 * n + 1]);
        }
        // dV += P^T dO ; dK += dS^T Q   (k = the 16 queries of this block)
        uint32_t th[4], tl[4];
        ldsm4_t(th, blk + 2 * ARR + off_t);
        ldsm4_t(tl, blk + 3 * ARR + off_t);
        mma_x3(dv[0], pth, ptl, th[0], th[1], tl[0], tl[1]);
        mma_x3(dv[1], pth, ptl, th[2], th[3], tl[2], tl[3]);
        ldsm4_t(th, blk + off_t);
        ldsm4_t(tl, blk + ARR + off_t);
        mma_x3(dk[0], sth, stl, th[0], th[1], tl[0], tl[1]);
        mma_x3(dk[1], sth, stl, th[2], th[3], tl[2], tl[3]);
        // dQ += dS K: A = dS (rows = queries) is the register transpose of the dS^T tiles
        uint32_t ah[4], al[4];
        ah[0] = movm(sth[0]); ah[1] = movm(sth[2]); ah[2] = movm(sth[1]); ah[3] = movm(sth[3]);
        al[0] = movm(stl[0]); al[1] = movm(stl[2]); al[2] = movm(stl[1]); al[3] = movm(stl[3]);
        mma_x3(dq[qb][0], ah, al, kth[0], kth[1], ktl[0], ktl[1]);
        mma_x3(dq[qb][1], ah, al, kth[2], kth[3], ktl[2], ktl[3]);
      }
    }
    // dK = scale * sum dS^T Q  (Q was staged pre-scaled by scale*log2(e))
    if (ok0) {
      float* dst = d_qkv + r0 * (3 * D) + h * HD + 4 * t;
      *reinterpret_cast<float4*>(dst + D) =
          make_float4(dk[0][0] * kLn2, dk[0][1] * kLn2, dk[1][0] * kLn2, dk[1][1] * kLn2);
      *reinterpret_cast<float4*>(dst + 2 * D) = make_float4(dv[0][0], dv[0][1], dv[1][0], dv[1][1]);
    }
    if (ok1) {
      float* dst = d_qkv + r1 * (3 * D) + h * HD + 4 * t;
      *reinterpret_cast<float4*>(dst + D) =
          make_float4(dk[0][2] * kLn2, dk[0][3] * kLn2, dk[1][2] * kLn2, dk[1][3] * kLn2);
      *reinterpret_cast<float4*>(dst + 2 * D) = make_float4(dv[0][2], dv[0][3], dv[1][2], dv[1][3]);
    }
    db0 = quad_sum(db0);
    db1 = quad_sum(db1);
    if (t == 0) {   // key k >= 1 is edge row lo + k - 1 (the centre token, k = 0, has no cutoff factor)
      if (ok0 && k0 > 0) d_bias_part[h * n_tokens + r0] = db0;
      if (ok1 && k1 > 0) d_bias_part[h * n_tokens + r1] = db1;
    }
  }
#pragma unroll
  for (int qb = 0; qb < NKB; ++qb) {
    if (qb < nkb) {
      const int p0 = 16 * qb + g, p1 = p0 + 8;
      if (p0 < T) {
        *reinterpret_cast<float4*>(d_qkv + token_row(p0, lo, n_edges, atom) * (3 * D) + h * HD + 4 * t) =
            make_float4(dq[qb][0][0] * scale, dq[qb][0][1] * scale, dq[qb][1][0] * scale,
                        dq[qb][1][1] * scale);
      }
      if (p1 < T) {
        *reinterpret_cast<float4*>(d_qkv + token_row(p1, lo, n_edges, atom) * (3 * D) + h * HD + 4 * t) =
            make_float4(dq[qb][0][2] * scale, dq[qb][0][3] * scale, dq[qb][1][2] * scale,
                        dq[qb][1][3] * scale);
      }
    }
  }
  __syncwarp();   // the warp's staged rows are rewritten by its next atom
  }
}

// d_fc[e] += (sum over heads of the key-bias gradient of edge e) / f_e   (f_e > 1e-15: the clamp of
// transformer.py:109-110 has zero slope below)
__global__ void attention_dfc_kernel(const float* __restrict__ d_bias_part, const float* __restrict__ fc,
                                     int64_t n_edges, int64_t n_tokens, float* __restrict__ d_fc) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  float acc = 0.f;
#pragma unroll
  for (int hh = 0; hh < H; ++hh) acc += d_bias_part[hh * n_tokens + e];
  const float f = __ldg(fc + e);
  if (f >= 1e-15f) d_fc[e] += acc / f;
}

template <int NKB>
size_t fwd_smem() { return (size_t)H * (2 * (2 * NKB * 32) + 4 * NKB) * 16; }
template <int NKB>
size_t bwd_smem() {
  constexpr int Tp = 16 * NKB;
  return (size_t)H * (4 * 2 * Tp * 16 + 3 * Tp * 4);
}

template <int NKB>
int launch_fwd(const float* qkv, const int32_t* row_ptr, const float* fc, int64_t n_atoms,
               int64_t n_edges, float scale, float* out, float* lse, cudaStream_t stream) {
  auto kern = attention_fwd_tc_kernel<NKB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<NKB>());
  const int64_t grid = n_atoms < 3 * kNumSMs ? n_atoms : 3 * kNumSMs;   // 3 resident CTAs per SM
  kern<<<(unsigned)grid, 256, fwd_smem<NKB>(), stream>>>(qkv, row_ptr, fc, n_atoms, n_edges, scale, out, lse);
  return check_launch("attention_fwd_tc");
}
template <int NKB>
int launch_bwd(const float* qkv, const float* out, const float* lse, const float* d_out,
               const int32_t* row_ptr, const float* fc, int64_t n_atoms, int64_t n_edges,
               float scale, float* d_qkv, float* d_fc, float* scratch, cudaStream_t stream) {
  auto kern = attention_bwd_tc_kernel<NKB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem<NKB>());
  const int64_t grid = n_atoms < 2 * kNumSMs ? n_atoms : 2 * kNumSMs;   // 2 resident CTAs per SM
  kern<<<(unsigned)grid, 256, bwd_smem<NKB>(), stream>>>(qkv, out, lse, d_out, row_ptr, fc, n_atoms, n_edges,
                                                        scale, d_qkv, scratch);
  if (d_fc && n_edges > 0)
    attention_dfc_kernel<<<(unsigned)ceil_div(n_edges, 256), 256, 0, stream>>>(scratch, fc, n_edges,
                                                                              n_edges + n_atoms, d_fc);
  return check_launch("attention_bwd_tc");
}

}  // namespace

bool attention_tc_supports(int num_heads, int head_dim, int max_row) {
  return num_heads == H && head_dim == HD && max_row + 1 <= 64;
}

int launch_attention_fwd_tc(const float* qkv, const int32_t* row_ptr, const float* fc,
                            int64_t n_atoms, int64_t n_edges, float scale, int max_row, float* out,
                            float* lse, cudaStream_t stream) {
  switch ((max_row + 1 + 15) / 16) {
    case 1: return launch_fwd<1>(qkv, row_ptr, fc, n_atoms, n_edges, scale, out, lse, stream);
    case 2: return launch_fwd<2>(qkv, row_ptr, fc, n_atoms, n_edges, scale, out, lse, stream);
    case 3: return launch_fwd<3>(qkv, row_ptr, fc, n_atoms, n_edges, scale, out, lse, stream);
    default: return launch_fwd<4>(qkv, row_ptr, fc, n_atoms, n_edges, scale, out, lse, stream);
  }
}

int launch_attention_bwd_tc(const float* qkv, const float* out, const float* lse, const float* d_out,
                            const int32_t* row_ptr, const float* fc, int64_t n_atoms,
                            int64_t n_edges, float scale, int max_row, float* d_qkv, float* d_fc,
                            float* scratch, cudaStream_t stream) {
  switch ((max_row + 1 + 15) / 16) {
    case 1: return launch_bwd<1>(qkv, out, lse, d_out, row_ptr, fc, n_atoms, n_edges, scale, d_qkv, d_fc, scratch, stream);
    case 2: return launch_bwd<2>(qkv, out, lse, d_out, row_ptr, fc, n_atoms, n_edges, scale, d_qkv, d_fc, scratch, stream);
    case 3: return launch_bwd<3>(qkv, out, lse, d_out, row_ptr, fc, n_atoms, n_edges, scale, d_qkv, d_fc, scratch, stream);
    default: return launch_bwd<4>(qkv, out, lse, d_out, row_ptr, fc, n_atoms, n_edges, scale, d_qkv, d_fc, scratch, stream);
  }
}

}  // namespace petb200
