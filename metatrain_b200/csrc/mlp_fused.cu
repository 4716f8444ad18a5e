// Fused edge-token feed-forward block of a PET transformer layer, forward and backward, as
// single persistent tcgen05 kernels (sm_100a).
//
//   forward :  y  = x + W_out . swiglu(W_in . rmsnorm(x) + b_in) + b_out
//   backward:  dx = dy + rmsnorm'(x)^T [ W_in^T . swiglu'(u, g) (W_out^T . dy) ]
// (TransformerLayer._forward_pre_ln_impl, src/metatrain/pet/modules/transformer.py:229-232;
//  FeedForward with SwiGLU :21-50 — v, g = w_in(x).chunk(2); w_out(v * sigmoid(g)).)
//
// The unfused path (gemm_tc.cu) moves ~25 KB per edge and layer through HBM for this block
// (the [E, 2F] pre-activations are written, read, and their gradient written and read
// again).  Here the hidden activations never leave the SM: per 128-row tile the kernel
// walks the hidden dimension in chunks of 32 units,
//   GEMM1  acc1[128 x 64]   = X[128 x 128] . W_in[chunk]^T        (tcgen05.mma, A and B in smem)
//   epi1   a[128 x 32]      = (rs*v + b_v) * sigmoid(rs*g + b_g)  (TMEM -> regs -> TMEM, bf16 hi/lo)
//   GEMM2  acc2[128 x 128] += a . W_out[:, chunk]^T               (tcgen05.mma, A in TMEM, B in smem)
// and the backward recomputes the pre-activations instead of loading them,
//   GEMM1a ug = X . W_in[chunk]^T, GEMM1b ds = dY . W_out[:, chunk], epi1 d_ug = swiglu'(ug) ds,
//   GEMM2  d_xhat += d_ug . W_in[chunk]      followed by the RMSNorm backward in the epilogue.
// HBM traffic drops to the algorithmic minimum: forward 1 KB per row (read x, write y),
// backward 1.5 KB per row (read x, dy, write dx).  All products use the bf16 hi/lo 2-term split
// of gemm_tc.cu (lo*hi + hi*lo + hi*hi, fp32 accumulation in TMEM).
//
// Weights are streamed from L2 as 16 KB "stages": a pre-swizzled image of the exact
// shared-memory operand tiles in the order the MMA warp consumes them (petb200_mlp_pack),
// moved by one thread with 1-D bulk async copies (cp.async.bulk, TMA engine) into a ring.
//
// Warp roles (448 threads): warps 0-7 epilogues, warp 8 MMA issue + TMEM owner, warps 9-12
// activation producers (cp.async fp32 rows -> in-place bf16 hi/lo swizzled operand tiles +
// per-row RMS statistics), warp 13 weight-stage producer.
#include <cuda_bf16.h>

#include "fused_common.cuh"

namespace petb200 {
namespace {

using namespace tc;
using namespace fused;


// forward image: per pair of chunks: W1(2p,k0) W1(2p,k1) W1(2p+1,k0) W1(2p+1,k1) W2hi(p) W2lo(p)
__host__ __device__ constexpr int fwd_stages(int F) { return (F / 64) * 6; }
// backward image: G1 stages of chunk 0; then for c >= 1: G1 stages of c, WT(c-1); finally WT(last)
__host__ __device__ constexpr int bwd_stages(int F) { return (F / CH) * 5; }

// Optional in-kernel timeline (tools/mlp_trace.py): CTA 0 records clock64() at pipeline events of
// its first tiles into a caller-provided buffer.  Off (null pointer) in normal operation.
#ifdef PETB200_MLP_TRACE
__device__ long long* g_trace = nullptr;
constexpr int TRACE_SLOTS = 16;    // events per (role, tile, chunk)
__device__ __forceinline__ void trace(int role, int tile, int chunk, int ev) {
  if (g_trace != nullptr && blockIdx.x == 0 && tile < 4)
    g_trace[((role * 4 + tile) * 16 + chunk) * TRACE_SLOTS + ev] = clock64();
}
#else   // compiled out: even a disabled run-time check costs a global load per event
__device__ __forceinline__ void trace(int, int, int, int) {}
#endif


// --------------------------------------------------------------------- weight images
// source row of w_in ([2F, D], RMSNorm weight folded in) behind GEMM1 column n (0..63) of
// chunk c: columns [0,16) = value units 0..15 of the chunk's first half, [16,32) the matching
// gate units, [32,64) the same for the second half — so each epilogue warp sees both members
// of every (value, gate) pair.
__device__ __forceinline__ int w1_row(int c, int n, int F) {
  const int hf = n >> 5, s = (n >> 4) & 1, i = n & 15;
  return s * F + c * CH + hf * 16 + i;
}


// one thread per 16-byte chunk (8 bf16) of the image
__global__ void mlp_pack_kernel(const float* __restrict__ w_in, const float* __restrict__ w_out, int F,
                                int backward, uint4* __restrict__ image) {
  const int n_stage = backward ? bwd_stages(F) : fwd_stages(F);
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n_stage * (STAGE / 16)) return;
  const int s = (int)(idx / (STAGE / 16));
  const int o = (int)(idx % (STAGE / 16)) * 16;  // byte offset inside the stage
  // stage kind: 0 = W1(c, kh) ; 1 = W2 hi/lo (pair p) ; 2 = WO(c) ; 3 = WT hi/lo (c)
  int kind, c = 0, sub = 0;
  if (!backward) {
    const int p = s / 6, r = s % 6;
    if (r < 4) { kind = 0; c = 2 * p + r / 2; sub = r % 2; }
    else { kind = 1; c = p; sub = r - 4; }
  } else {
    const int last = F / CH - 1;
    int r;
    if (s < 3) { c = 0; r = s; }
    else if (s < 3 + 5 * last) { c = 1 + (s - 3) / 5; r = (s - 3) % 5; }
    else { c = last + 1; r = 3 + (s - 3 - 5 * last); }
    if (r < 2) { kind = 0; sub = r; }
    else if (r == 2) { kind = 2; }
    else { kind = 3; c -= 1; sub = r - 3; }
  }
  float v[8];
  bool lo;
  if (kind == 0) {          // [64 rows x 64 k] hi (8 KB) | lo (8 KB)
    lo = o >= 8192;
    const int t = o & 8191, n = (t >> 10) * 8 + ((t >> 7) & 7), j = ((t >> 4) & 7) ^ (n & 7);
    const float* src = w_in + (int64_t)w1_row(c, n, F) * D + sub * 64 + j * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = src[e];
  } else if (kind == 1) {   // [128 rows (out dim) x 64 k (hidden units of the pair)]
    lo = sub == 1;
    const int n = (o >> 10) * 8 + ((o >> 7) & 7), j = ((o >> 4) & 7) ^ (n & 7);
    const float* src = w_out + (int64_t)n * F + c * 64 + j * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = src[e];
  } else if (kind == 2) {   // per k-half: [32 rows (hidden unit) x 64 k (out dim)] hi (4 KB) | lo (4 KB)
    const int kh = o >> 13, t = o & 8191;
    lo = t >= 4096;
    const int u = t & 4095, n = (u >> 10) * 8 + ((u >> 7) & 7), j = ((u >> 4) & 7) ^ (n & 7);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = w_out[(int64_t)(kh * 64 + j * 8 + e) * F + c * CH + n];
  } else {                  // [128 rows (in dim) x 64 k (d_v / d_g order of epi1)]
    lo = sub == 1;
    const int n = (o >> 10) * 8 + ((o >> 7) & 7), j = ((o >> 4) & 7) ^ (n & 7);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = w_in[(int64_t)w1_row(c, j * 8 + e, F) * D + n];
  }
  image[idx] = pack8(v, lo);
}

// ------------------------------------------------------------------ shared pieces

// Producer side of the backward kernel: a 128 x 128 fp32 tile lands (TMA boxes) as the four bf16 operand
// tiles' bytes ([k-chunk][hi|lo], 16 KB each): floats 0..31 of a k-chunk in the bytes of the hi row and
// floats 32..63 in the lo row; warp `pw` then converts its rows 32 pw .. 32 pw + 31 in place (same scheme
// as gemm_tc.cu).
// In-place conversion of 16 rows (row_base ..) of one k-chunk: fp32 landed as [hi tile row | lo
// tile row] -> bf16 hi / lo swizzled rows.  Lane l works on row 2 jj + (l >> 4) of the batch and
// owns k = 4 (l & 15) .. + 3.  ss[jj] accumulates this lane's partial sum of squares.
template <bool STATS>
__device__ __forceinline__ void convert_batch(uint8_t* st, int row_base, int lane, float (&ss)[8]) {
  const int quad = lane & 15, rsub = lane >> 4;
  const int chunk = quad >> 1, within = (quad & 1) * 8;
  const uint8_t* src_tile = st + (quad < 8 ? 0 : TILE) + (quad & 7) * 16;
  float4 x[8];
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int r = row_base + jj * 2 + rsub;
    x[jj] = *reinterpret_cast<const float4*>(src_tile + r * 128);
  }
  __syncwarp();  // whole rows are in registers before they are overwritten
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const int r = row_base + jj * 2 + rsub;
    if (STATS) ss[jj] += x[jj].x * x[jj].x + x[jj].y * x[jj].y + x[jj].z * x[jj].z + x[jj].w * x[jj].w;
    const uint32_t h01 = pack_bf16(x[jj].x, x[jj].y), h23 = pack_bf16(x[jj].z, x[jj].w);
    const uint32_t off = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4) + within);
    *reinterpret_cast<uint2*>(st + off) = make_uint2(h01, h23);
    const float l0 = x[jj].x - __uint_as_float(h01 << 16);
    const float l1 = x[jj].y - __uint_as_float(h01 & 0xffff0000u);
    const float l2 = x[jj].z - __uint_as_float(h23 << 16);
    const float l3 = x[jj].w - __uint_as_float(h23 & 0xffff0000u);
    *reinterpret_cast<uint2*>(st + TILE + off) = make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
  }
}
template <bool STATS>
__device__ __forceinline__ void convert_tile(uint8_t* dst, int pw, int lane, float (&ss)[2][8]) {
#pragma unroll
  for (int kc = 0; kc < 2; ++kc)
#pragma unroll
    for (int b = 0; b < 2; ++b) convert_batch<STATS>(dst + kc * 2 * TILE, pw * 32 + b * 16, lane, ss[b]);
}
__device__ __forceinline__ void store_rstd(float* rstd, int pw, int lane, float (&ss)[2][8]) {
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      float v = ss[b][jj];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      if ((lane & 15) == 0) rstd[pw * 32 + b * 16 + jj * 2 + (lane >> 4)] = rsqrtf(v * (1.0f / D) + kRmsEps);
    }
}


__device__ __forceinline__ void weight_producer(const uint8_t* __restrict__ image, int n_stage, int count,
                                                uint32_t ring_u32, const Barriers& bar) {
  Ring ring;
  for (int i = 0; i < count; ++i)
    for (int s = 0; s < n_stage; ++s) {
      mbar_wait(bar.w_empty(ring.stage), ring.phase ^ 1);
      mbar_expect_tx(bar.w_full(ring.stage), STAGE);
      bulk_g2s(ring_u32 + (uint32_t)ring.stage * STAGE, image + (size_t)s * STAGE, STAGE, bar.w_full(ring.stage));
      ring.advance(bar.ring);
    }
}


// ======================================================================== forward
// Forward-kernel layout.  The activation tile never sits in shared memory as an MMA operand:
// tcgen05.mma with A in shared memory is bound by the 128 B/clk shared-memory read port for
// N <= 64 (measured 48 clk per M128 N64 K16 instruction instead of 32, tools/umma_rate.cu),
// and that port also carries the weight stages.  So the producers stage the raw fp32 rows
// with cp.async (double buffered, pitch 132 floats: conflict-free row-per-thread reads),
// convert them to bf16 hi / lo in registers and park them in TENSOR memory (lane = row),
// where GEMM1 reads its A operand at full rate.  19 warps: 0-7 SwiGLU epilogue groups,
// 8 GEMM1 issue, 9-12 activation producers, 13 weight stages, 14-17 output store, 18 GEMM2
// issue (two issuing threads: the per-chunk barrier hops of one GEMM no longer delay the other).
constexpr int FWD_NUM_THREADS = 32 * 15;   // the four producer warps also run the output store (see below)
constexpr int FWD_MMA2_WARP = 14;
constexpr int FWD_RING_A = 5, FWD_RING_B = 4;   // W1 stages / W2 stages
constexpr int FWD_RING = FWD_RING_A + FWD_RING_B;
constexpr int XPITCH = D + 4;                                // floats per staged row
constexpr int FWD_STAGING_BYTES = BM * XPITCH * 4;           // 67 584
constexpr int FWD_XS_OFF = 0;                                // raw fp32 staging buffer
constexpr int FWD_RING_OFF = ((FWD_XS_OFF + FWD_STAGING_BYTES + 1023) / 1024) * 1024;
constexpr int FWD_EPI_OFF = FWD_RING_OFF + FWD_RING * STAGE; // 4 store warps x 32 x STAGE_LD floats
constexpr int FWD_BIAS_OFF = FWD_EPI_OFF + 4 * 32 * STAGE_LD * 4;  // b_in [2 MAX_F], b_out [D]
constexpr int FWD_RSTD_OFF = FWD_BIAS_OFF + (2 * MAX_F + D) * 4;
constexpr int FWD_BAR_OFF = FWD_RSTD_OFF + 2 * BM * 4;
constexpr int FWD_SMEM = FWD_BAR_OFF + 8 * (16 + 2 * FWD_RING) + 16 + 1024;
// TMEM columns: X hi 0..63, X lo 64..127 ; acc1 at 128 (single buffer: GEMM1 of the next chunk
// only waits for the tcgen05.ld of the previous one) ; A2[b] hi at 192 + 32 b, lo 16 further ;
// acc2[t] at 256 + 128 t (double buffer: the output store overlaps the next tile)
constexpr int FWD_XLO_COL = 64, FWD_ACC1_COL = 128, FWD_A2_COL = 192, FWD_ACC2_COL = 256;


__global__ void __launch_bounds__(FWD_NUM_THREADS, 1)
mlp_fwd_kernel(const float* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ image,
               const float* __restrict__ b_in, const float* __restrict__ b_out, int64_t M, int F,
               float* __restrict__ y, int64_t ldy) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const Barriers bar{smem_base + FWD_BAR_OFF, FWD_RING};
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + FWD_BAR_OFF + 8 * (16 + 2 * FWD_RING));
  float* bias_s = reinterpret_cast<float*>(smem + FWD_BIAS_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;  // TMEM lane quarter this warp may access
  const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
  const int nch = F / CH;
  const TileSchedule sched(M);

  if (threadIdx.x == 0) {
    bar.init_all();
    // only the four store warps drain acc2
    mbar_init(bar.acc2_empty(0), 4 * 32);
    mbar_init(bar.acc2_empty(1), 4 * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(const_cast<uint32_t*>(tmem_slot))));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < 2 * F; i += FWD_NUM_THREADS) bias_s[i] = b_in[i];
  for (int i = threadIdx.x; i < D; i += FWD_NUM_THREADS) bias_s[2 * MAX_F + i] = b_out[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ============================================================ output store (run by the producer warps:
  // they are idle once their tile is converted, and without four dedicated store warps the CTA has 480
  // threads = 128 registers per thread)
    // y = acc2 + b_out + x, overlapped with the chunk loop of the next tile.  Warp -> 32 rows
    // (its TMEM lane quarter) x 128 columns in 8 slices of 16; the residual of the next slice is
    // in flight while the current one is transposed and stored.
    const int sw = warp - FIRST_PROD_WARP;
    const EpiStage es{reinterpret_cast<float*>(smem + FWD_EPI_OFF) + (sw & 3) * (32 * STAGE_LD), lane, lane & 3,
                      (lane >> 3) + 4 * ((lane >> 2) & 1)};
  auto store_tile = [&](int i) {
    {
      const int64_t m_base = sched.m0(i) + quarter * 32;
      const int t = i & 1;
      float4 res[3][4];
      auto fetch = [&](int sl) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int64_t m = m_base + it * 8 + es.rsel;
          res[sl % 3][it] = m < M ? ld4(x + m * ldx + 16 * sl + 4 * es.c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      fetch(0);
      fetch(1);
      if (lane == 0 && sw == 0) trace(1, i, 15, 0);
      mbar_wait(bar.acc2_full(t), (i >> 1) & 1);
      tc_fence_after();
      if (lane == 0 && sw == 0) trace(1, i, 15, 1);
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {
        if (sl + 2 < 8) fetch(sl + 2);
        const int c0 = 16 * sl + 4 * es.c4;
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + 2 * MAX_F + c0);
        es.fill(tmem_base + lane_base + FWD_ACC2_COL + t * D + 16 * sl);
        if (sl == 7) {   // the accumulator is in registers / smem now: release it early
          tc_fence_before();
          mbar_arrive(bar.acc2_empty(t));
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int64_t m = m_base + it * 8 + es.rsel;
          if (m >= M) continue;
          const float4 a = es.get(it), r = res[sl % 3][it];
          *reinterpret_cast<float4*>(y + m * ldy + c0) =
              make_float4(a.x + b4.x + r.x, a.y + b4.y + r.y, a.z + b4.z + r.z, a.w + b4.w + r.w);
        }
      }
      if (lane == 0 && sw == 0) trace(1, i, 15, 2);
    }
  };

  if (warp >= FIRST_PROD_WARP && warp < TMA_WARP) {
    // ============================================================ activation producers
    // warp -> rows 32 * quarter .. + 31 (the TMEM lanes it may write); lane -> one row
    auto issue = [&](int i) {
      const uint32_t dst = smem_base + FWD_XS_OFF;
      const int64_t m0 = sched.m0(i);
#pragma unroll 8
      for (int it = 0; it < 32; ++it) {  // one row (32 x 16 B) per instruction
        const int row = quarter * 32 + it;
        const int64_t m = m0 + row;
        const bool ok = m < M;
        cp_async16(dst + (uint32_t)(row * XPITCH * 4 + lane * 16), x + (ok ? m : 0) * ldx + 4 * lane,
                   ok ? 16u : 0u);
      }
      cp_async_commit();
    };
    if (sched.count > 0) issue(0);
    for (int i = 0; i < sched.count; ++i) {
      cp_async_wait_group<0>();
      __syncwarp();
      if (lane == 0 && quarter == 1) trace(3, i, 0, 0);
      mbar_wait(bar.x_empty(0), (i & 1) ^ 1);   // GEMM1 of the previous tile has consumed X
      tc_fence_after();
      if (lane == 0 && quarter == 1) trace(3, i, 0, 1);
      const float* row = reinterpret_cast<const float*>(smem + FWD_XS_OFF) + (quarter * 32 + lane) * XPITCH;
      // pass 1: RMS statistic of the row; pass 2: x_hat = x * rstd -> bf16 hi / lo -> TMEM (the
      // norm weight is folded into W_in, so GEMM1 consumes the normalised row directly)
      float ss = 0.f;
#pragma unroll 8
      for (int q = 0; q < 32; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(row + 4 * q);
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      const float rs = rsqrtf(ss * (1.0f / D) + kRmsEps);
#pragma unroll
      for (int part = 0; part < 4; ++part) {   // 32 floats -> 16 hi + 16 lo columns
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v = *reinterpret_cast<const float4*>(row + part * 32 + 4 * q);
          v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
          hi[2 * q] = pack_bf16(v.x, v.y);
          hi[2 * q + 1] = pack_bf16(v.z, v.w);
          lo[2 * q] = pack_bf16(v.x - __uint_as_float(hi[2 * q] << 16), v.y - __uint_as_float(hi[2 * q] & 0xffff0000u));
          lo[2 * q + 1] = pack_bf16(v.z - __uint_as_float(hi[2 * q + 1] << 16),
                                    v.w - __uint_as_float(hi[2 * q + 1] & 0xffff0000u));
        }
        tmem_st16(tmem_base + lane_base + part * 16, hi);
        tmem_st16(tmem_base + lane_base + FWD_XLO_COL + part * 16, lo);
      }
      // this warp's rows of the staging buffer are consumed: fetch the next tile behind the
      // whole chunk loop of this one
      __syncwarp();
      if (i + 1 < sched.count) issue(i + 1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar.x_full(0));
      if (lane == 0 && quarter == 1) trace(3, i, 0, 3);
      if (i > 0) store_tile(i - 1);   // overlaps the chunk loop of tile i
    }
    if (sched.count > 0) store_tile(sched.count - 1);
  } else if (warp == TMA_WARP) {
    // ============================================================ weight-stage producer
    // one thread walks the image in order and routes W1 stages to ring A (slots 0..RA-1) and
    // W2 stages to ring B (slots RA..RA+RB-1); a stage only ever waits for EARLIER stages to be
    // consumed, so the two rings cannot deadlock each other
    if (elect_one()) {
      Ring ra, rb;
      const int n_stage = fwd_stages(F);
      for (int i = 0; i < sched.count; ++i)
        for (int st = 0; st < n_stage; ++st) {
          const bool is_b = (st % 6) >= 4;
          Ring& r = is_b ? rb : ra;
          const int slot = is_b ? FWD_RING_A + r.stage : r.stage;
          mbar_wait(bar.w_empty(slot), r.phase ^ 1);
          mbar_expect_tx(bar.w_full(slot), STAGE);
          bulk_g2s(smem_base + FWD_RING_OFF + (uint32_t)slot * STAGE, image + (size_t)st * STAGE, STAGE,
                   bar.w_full(slot));
          r.advance(is_b ? FWD_RING_B : FWD_RING_A);
        }
    }
  } else if (warp == MMA_WARP) {
    // ============================================================ GEMM1 issuer (one thread)
    if (elect_one()) {
      constexpr uint32_t idesc1 = make_idesc(BM, 64);
      const uint32_t ring_u32 = smem_base + FWD_RING_OFF;
      Ring ring;
      for (int i = 0; i < sched.count; ++i) {
        mbar_wait(bar.x_full(0), i & 1);
        for (int c = 0; c < nch; ++c) {
          const int b = c & 1;
          trace(0, i, c, 0);
          // single acc1 buffer: the epilogue group of the PREVIOUS chunk (barrier pair b ^ 1)
          // must have pulled it into registers
          const int n = i * nch + c;
          if (n > 0) mbar_wait(bar.acc1_empty(b ^ 1), ((uint32_t)(n - 1) >> 1) & 1);
          tc_fence_after();
          trace(0, i, c, 1);
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            mbar_wait(bar.w_full(ring.stage), ring.phase);
            const uint32_t st = ring_u32 + ring.stage * STAGE;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t a_hi = tmem_base + (kh * 4 + kk) * 8;
              mma3_ts(tmem_base + FWD_ACC1_COL, a_hi, a_hi + FWD_XLO_COL, st + kk * 32,
                      st + 8192 + kk * 32, idesc1, (kh | kk) != 0);
            }
            tc_commit(bar.w_empty(ring.stage));
            ring.advance(FWD_RING_A);
          }
          tc_commit(bar.acc1_full(b));
          if (c == nch - 1) tc_commit(bar.x_empty(0));
          trace(0, i, c, 7);
        }
      }
    }
  } else if (warp == FWD_MMA2_WARP) {
    // ============================================================ GEMM2 issuer (one thread)
    if (elect_one()) {
      constexpr uint32_t idesc2 = make_idesc(BM, D);
      const uint32_t ring_u32 = smem_base + FWD_RING_OFF + FWD_RING_A * STAGE;
      Ring ring;
      int w2_hi = 0, w2_lo = 0;
      for (int i = 0; i < sched.count; ++i) {
        for (int c = 0; c < nch; ++c) {
          const int b = c & 1;
          const uint32_t u = (uint32_t)(i * nch + c) >> 1;
          if (b == 0) {
            mbar_wait(bar.w_full(FWD_RING_A + ring.stage), ring.phase);
            w2_hi = ring.stage;
            ring.advance(FWD_RING_B);
            mbar_wait(bar.w_full(FWD_RING_A + ring.stage), ring.phase);
            w2_lo = ring.stage;
            ring.advance(FWD_RING_B);
          }
          trace(0, i, c, 3);
          mbar_wait(bar.a2_full(b), u & 1);
          if (c == 0) mbar_wait(bar.acc2_empty(i & 1), ((i >> 1) & 1) ^ 1);
          tc_fence_after();
          trace(0, i, c, 4);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t koff = (uint32_t)(b * 64 + kk * 32);
            const uint32_t a_hi = tmem_base + FWD_A2_COL + b * 32 + kk * 8;
            mma3_ts(tmem_base + FWD_ACC2_COL + (i & 1) * D, a_hi, a_hi + 16, ring_u32 + w2_hi * STAGE + koff,
                    ring_u32 + w2_lo * STAGE + koff, idesc2, (c | kk) != 0);
          }
          tc_commit(bar.a2_empty(b));
          if (b == 1) {
            tc_commit(bar.w_empty(FWD_RING_A + w2_hi));
            tc_commit(bar.w_empty(FWD_RING_A + w2_lo));
          }
          trace(0, i, c, 8);
        }
        tc_commit(bar.acc2_full(i & 1));
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ============================================================ SwiGLU epilogues
    const int half = warp >> 2;
    for (int i = 0; i < sched.count; ++i) {
      // the two groups of four warps take alternate chunks (group = chunk parity = buffer), so
      // the epilogue of chunk c overlaps GEMM1 of c+1, GEMM2 of c-1 AND the epilogue of c+1;
      // a thread owns one row: 64 accumulator columns in, 32 activations out
      for (int c = half; c < nch; c += 2) {
        const int b = half;
        const uint32_t u = (uint32_t)(i * nch + c) >> 1;
        if (lane == 0 && quarter == 0) trace(1 + half, i, c, 0);
        mbar_wait(bar.acc1_full(b), u & 1);
        tc_fence_after();
        if (lane == 0 && quarter == 0) trace(1 + half, i, c, 1);
        float v[64];
        tmem_ld32(tmem_base + lane_base + FWD_ACC1_COL, v);
        tmem_ld32(tmem_base + lane_base + FWD_ACC1_COL + 32, v + 32);
        tc_fence_before();
        mbar_arrive(bar.acc1_empty(b));
        if (lane == 0 && quarter == 0) trace(1 + half, i, c, 2);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const float* bv = bias_s + c * CH + hf * 16;
          const float* vv = v + hf * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float a0 = (vv[2 * q] + bv[2 * q]) * fsigmoid(vv[16 + 2 * q] + bv[F + 2 * q]);
            const float a1 = (vv[2 * q + 1] + bv[2 * q + 1]) * fsigmoid(vv[17 + 2 * q] + bv[F + 2 * q + 1]);
            const uint32_t h2 = pack_bf16(a0, a1);
            hi[hf * 8 + q] = h2;
            lo[hf * 8 + q] = pack_bf16(a0 - __uint_as_float(h2 << 16), a1 - __uint_as_float(h2 & 0xffff0000u));
          }
        }
        if (lane == 0 && quarter == 0) trace(1 + half, i, c, 3);
        mbar_wait(bar.a2_empty(b), (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t a2 = tmem_base + lane_base + FWD_A2_COL + b * 32;
        tmem_st16(a2, hi);
        tmem_st16(a2 + 16, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar.a2_full(b));
        if (lane == 0 && quarter == 0) trace(1 + half, i, c, 4);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// ================================================================ RMSNorm + Linear
// out = rmsnorm(x) . W^T + bias for K = 128 and N a multiple of 64 (the QKV projection,
// transformer.py:105-108 after norm_attention :218), built from the forward kernel's pieces:
// the producers normalise each row once (statistics per thread = per row, rstd also written
// out for the backward), park x_hat as bf16 hi / lo in tensor memory, and ONE pass over the
// tile produces all N output columns in chunks of 64 — the generic GEMM (gemm_tc.cu) lets a
// CTA own one 128-column chunk, so a 384-column projection re-reads and re-converts every
// activation tile three times and needs a separate RMS-statistics kernel in front.
// 14 warps: 0-7 epilogue (two groups on alternate chunks; TMEM -> registers (+ bias) -> a swizzled
// [32 rows x 32 columns] box in shared memory -> ONE TMA tile store per box: per 32 columns a warp issues
// 1 tcgen05.ld + 8 STS.128 + 1 bulk tensor store instead of 2 tcgen05.ld + 8 STS.128 + 8 LDS.128 +
// 8 STG.128 — the memory-instruction queue (MIO) of the SM was what this kernel waited for), 8 MMA
// issue, 9-12 activation producers, 13 weight stages.
constexpr int NL_RING = 5;
constexpr int NL_BOX_BYTES = 32 * 128;                        // [32 rows x 32 floats], SWIZZLE_128B
constexpr int NL_XS_OFF = 0;
constexpr int NL_RING_OFF = ((NL_XS_OFF + FWD_STAGING_BYTES + 1023) / 1024) * 1024;
constexpr int NL_EPI_OFF = NL_RING_OFF + NL_RING * STAGE;    // 8 warps x 2 boxes
constexpr int NL_BIAS_OFF = NL_EPI_OFF + NUM_EPI_WARPS * 2 * NL_BOX_BYTES;    // bias [NL_MAX_N]
constexpr int NL_MAX_N = 1024;
constexpr int NL_BAR_OFF = NL_BIAS_OFF + NL_MAX_N * 4;
constexpr int NL_SMEM = NL_BAR_OFF + 8 * (16 + 2 * NL_RING) + 16 + 1024;
// TMEM columns: x_hat[i & 1] hi at 128 (i & 1), lo 64 further ; acc[b] at 256 + 64 b
constexpr int NL_ACC_COL = 256;

__global__ void __launch_bounds__(NUM_THREADS, 1)
norm_linear_kernel(const __grid_constant__ CUtensorMap map_out, const float* __restrict__ x, int64_t ldx,
                   const uint8_t* __restrict__ image, const float* __restrict__ bias, int64_t M, int N,
                   float* __restrict__ rstd_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const Barriers bar{smem_base + NL_BAR_OFF, NL_RING};
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + NL_BAR_OFF + 8 * (16 + 2 * NL_RING));
  float* bias_s = reinterpret_cast<float*>(smem + NL_BIAS_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;
  const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
  const int nch = N / 64;
  const TileSchedule sched(M);

  if (threadIdx.x == 0) bar.init_all();   // acc1_empty: one epilogue group (128 threads) per chunk
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(const_cast<uint32_t*>(tmem_slot))));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < N; i += NUM_THREADS) bias_s[i] = bias ? bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= FIRST_PROD_WARP && warp < TMA_WARP) {
    // ============================================================ activation producers
    auto issue = [&](int i) {
      const uint32_t dst = smem_base + NL_XS_OFF;
      const int64_t m0 = sched.m0(i);
#pragma unroll 8
      for (int it = 0; it < 32; ++it) {
        const int row = quarter * 32 + it;
        const int64_t m = m0 + row;
        const bool ok = m < M;
        cp_async16(dst + (uint32_t)(row * XPITCH * 4 + lane * 16), x + (ok ? m : 0) * ldx + 4 * lane,
                   ok ? 16u : 0u);
      }
      cp_async_commit();
    };
    if (sched.count > 0) issue(0);
    for (int i = 0; i < sched.count; ++i) {
      cp_async_wait_group<0>();
      __syncwarp();
      const int xb = i & 1;                      // x_hat is double buffered in tensor memory
      mbar_wait(bar.x_empty(xb), ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      const float* row = reinterpret_cast<const float*>(smem + NL_XS_OFF) + (quarter * 32 + lane) * XPITCH;
      float ss = 0.f;
#pragma unroll 8
      for (int q = 0; q < 32; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(row + 4 * q);
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      const float rs = rsqrtf(ss * (1.0f / D) + kRmsEps);
      const int64_t m = sched.m0(i) + quarter * 32 + lane;
      if (rstd_out != nullptr && m < M) rstd_out[m] = rs;
#pragma unroll
      for (int part = 0; part < 4; ++part) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v = *reinterpret_cast<const float4*>(row + part * 32 + 4 * q);
          v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
          hi[2 * q] = pack_bf16(v.x, v.y);
          hi[2 * q + 1] = pack_bf16(v.z, v.w);
          lo[2 * q] = pack_bf16(v.x - __uint_as_float(hi[2 * q] << 16), v.y - __uint_as_float(hi[2 * q] & 0xffff0000u));
          lo[2 * q + 1] = pack_bf16(v.z - __uint_as_float(hi[2 * q + 1] << 16),
                                    v.w - __uint_as_float(hi[2 * q + 1] & 0xffff0000u));
        }
        tmem_st16(tmem_base + lane_base + xb * 128 + part * 16, hi);
        tmem_st16(tmem_base + lane_base + xb * 128 + FWD_XLO_COL + part * 16, lo);
      }
      __syncwarp();
      if (i + 1 < sched.count) issue(i + 1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar.x_full(xb));
    }
  } else if (warp == TMA_WARP) {
    if (elect_one()) weight_producer(image, 2 * nch, sched.count, smem_base + NL_RING_OFF, bar);
  } else if (warp == MMA_WARP) {
    // ============================================================ MMA issuer (one thread)
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(BM, 64);
      const uint32_t ring_u32 = smem_base + NL_RING_OFF;
      Ring ring;
      for (int i = 0; i < sched.count; ++i) {
        const int xb = i & 1;
        mbar_wait(bar.x_full(xb), (i >> 1) & 1);
        tc_fence_after();
        for (int c = 0; c < nch; ++c) {
          const uint32_t n = (uint32_t)(i * nch + c);
          const int b = n & 1;
          mbar_wait(bar.acc1_empty(b), ((n >> 1) & 1) ^ 1);
          tc_fence_after();
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            mbar_wait(bar.w_full(ring.stage), ring.phase);
            const uint32_t st = ring_u32 + ring.stage * STAGE;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t a_hi = tmem_base + xb * 128 + (kh * 4 + kk) * 8;
              mma3_ts(tmem_base + NL_ACC_COL + b * 64, a_hi, a_hi + FWD_XLO_COL, st + kk * 32,
                      st + 8192 + kk * 32, idesc, (kh | kk) != 0);
            }
            tc_commit(bar.w_empty(ring.stage));
            ring.advance(NL_RING);
          }
          tc_commit(bar.acc1_full(b));
          if (c == nch - 1) tc_commit(bar.x_empty(xb));
        }
      }
    }
  } else {
    // ============================================================ epilogue: out = acc + bias
    const int grp = warp >> 2;
    const uint32_t box_u32 = smem_base + NL_EPI_OFF + (uint32_t)warp * 2 * NL_BOX_BYTES;
    uint8_t* box = smem + NL_EPI_OFF + warp * 2 * NL_BOX_BYTES;
    uint32_t nbox = 0;
    for (int i = 0; i < sched.count; ++i) {
      const int m_row = (int)(sched.m0(i) + quarter * 32);
      for (int c = 0; c < nch; ++c) {
        const uint32_t n = (uint32_t)(i * nch + c);
        const int b = n & 1;
        if (b != grp) continue;           // chunk n belongs to group n & 1 (its accumulator buffer)
        mbar_wait(bar.acc1_full(b), (n >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int sl = 0; sl < 2; ++sl, ++nbox) {
          float v[32];
          tmem_ld32(tmem_base + lane_base + NL_ACC_COL + b * 64 + 32 * sl, v);
          if (sl == 1) {
            tc_fence_before();
            mbar_arrive(bar.acc1_empty(b));
          }
          // the tile store that used this box two slices ago has read it
          if (lane == 0) bulk_wait_group_read<1>();
          __syncwarp();
          const float* bs = bias_s + c * 64 + 32 * sl;
          uint8_t* dst = box + (nbox & 1) * NL_BOX_BYTES + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {   // row `lane`, 16-byte chunk j at its SWIZZLE_128B position
            const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * j);
            *reinterpret_cast<float4*>(dst + ((j ^ (lane & 7)) << 4)) =
                make_float4(v[4 * j] + b4.x, v[4 * j + 1] + b4.y, v[4 * j + 2] + b4.z, v[4 * j + 3] + b4.w);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_out, c * 64 + 32 * sl, m_row, box_u32 + (nbox & 1) * NL_BOX_BYTES);
            bulk_commit_group();
          }
        }
      }
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// image of a plain [N, 128] weight for norm_linear_kernel: per 64-row chunk two stages (k-halves),
// each [64 rows x 64 k] bf16 hi (8 KB) | lo (8 KB), K-major SWIZZLE_128B
__global__ void linear_pack_kernel(const float* __restrict__ w, int N, uint4* __restrict__ image) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)(N / 64) * 2 * (STAGE / 16)) return;
  const int s = (int)(idx / (STAGE / 16)), o = (int)(idx % (STAGE / 16)) * 16;
  const int c = s >> 1, kh = s & 1;
  const bool lo = o >= 8192;
  const int t = o & 8191, n = (t >> 10) * 8 + ((t >> 7) & 7), j = ((t >> 4) & 7) ^ (n & 7);
  const float* src = w + (int64_t)(c * 64 + n) * D + kh * 64 + j * 8;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = src[e];
  image[idx] = pack8(v, lo);
}

// ======================================================================= backward
// 15 warps: 0-7 swiglu' epilogue groups, 8 GEMM1 issue (ug and ds), 9-12 activation producers (X and
// dY tiles as bf16 hi / lo operand tiles in shared memory) which ALSO run the output store of the
// previous tile (RMSNorm backward) once their tile is converted — they are idle for the rest of the
// chunk loop, and without four dedicated store warps the CTA has 480 threads: 128 registers per
// thread instead of 96, no spills in the epilogue —, 13 weight stages, 14 GEMM2 issue.
constexpr int BWD_NUM_THREADS = 32 * 15;
constexpr int BWD_MMA2_WARP = 14;
constexpr int BWD_RING_A = 3, BWD_RING_B = 2;                 // G1 stages / WT stages
constexpr int BWD_RING = BWD_RING_A + BWD_RING_B;
constexpr int BWD_X_OFF = 0;                                  // X tiles (4), then dY tiles (4)
constexpr int BWD_RING_OFF = BWD_X_OFF + 8 * TILE;
constexpr int BWD_EPI_OFF = BWD_RING_OFF + BWD_RING * STAGE;  // 4 store warps x 32 x STAGE_LD floats
constexpr int BWD_BIAS_OFF = BWD_EPI_OFF + 4 * 32 * STAGE_LD * 4;   // b_in [2 MAX_F]
constexpr int BWD_RSTD_OFF = BWD_BIAS_OFF + 2 * MAX_F * 4;    // rstd [2][BM]
constexpr int BWD_BAR_OFF = BWD_RSTD_OFF + 2 * BM * 4;
constexpr int BWD_SMEM = BWD_BAR_OFF + 8 * (16 + 2 * BWD_RING) + 8 + 16 + 1024;
// TMEM columns: acc1 at 0 (ug: 0..63, ds: 64..95; single buffer) ; A2[b] hi at 128 + 64 b, lo 32
// further ; acc2[t] at 256 + 128 t
constexpr int BWD_A2_COL = 128, BWD_ACC2_COL = 256;

__global__ void __launch_bounds__(BWD_NUM_THREADS, 1)
mlp_bwd_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
               const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, int64_t ld_dy,
               const uint8_t* __restrict__ image, const float* __restrict__ b_in, int64_t M, int F,
               float* __restrict__ dx, int64_t ld_dx) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const Barriers bar{smem_base + BWD_BAR_OFF, BWD_RING};
  const uint32_t landed_bar = bar.at(16 + 2 * BWD_RING);   // TMA of the X / dY tile has landed
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem + BWD_BAR_OFF + 8 * (16 + 2 * BWD_RING) + 8);
  float* bias_s = reinterpret_cast<float*>(smem + BWD_BIAS_OFF);
  float* rstd_s = reinterpret_cast<float*>(smem + BWD_RSTD_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;
  const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
  const int nch = F / CH;
  const TileSchedule sched(M);

  if (threadIdx.x == 0) {
    bar.init_all();
    mbar_init(bar.acc2_empty(0), 4 * 32);   // drained by the four store warps
    mbar_init(bar.acc2_empty(1), 4 * 32);
    for (int b = 0; b < 2; ++b) {           // every epilogue warp takes part in every chunk
      mbar_init(bar.acc1_empty(b), NUM_EPI_WARPS * 32);
      mbar_init(bar.a2_full(b), NUM_EPI_WARPS * 32);
    }
    // X rows are converted by the producer warps, dY rows by the (then idle) epilogue warps
    mbar_init(bar.x_full(0), NUM_PROD_THREADS + NUM_EPI_WARPS * 32);
    mbar_init(landed_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(const_cast<uint32_t*>(tmem_slot))));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < 2 * F; i += BWD_NUM_THREADS) bias_s[i] = b_in[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ============================================================ output store (run by the producer warps)
    // dx = dy + rs * d - x * rs^3 * (d . x) / D   (d = acc2 = gradient w.r.t. x_hat), overlapped
    // with the chunk loop of the next tile.  Warp -> 32 rows x 128 columns in 8 slices of 16:
    // pass 1 accumulates d . x per row, pass 2 forms the output; x (and dy) of the next slice
    // are in flight while the current one is transposed.
    const int sw = warp - FIRST_PROD_WARP;
    const EpiStage es{reinterpret_cast<float*>(smem + BWD_EPI_OFF) + (sw & 3) * (32 * STAGE_LD), lane, lane & 3,
                      (lane >> 3) + 4 * ((lane >> 2) & 1)};
  auto store_tile = [&](int i) {
    {
      const int t = i & 1;
      const int64_t m_base = sched.m0(i) + quarter * 32;
      float4 xr[2][4], gr[2][4];
      auto fetch_x = [&](int sl) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int64_t m = m_base + it * 8 + es.rsel;
          xr[sl & 1][it] = m < M ? ld4(x + m * ldx + 16 * sl + 4 * es.c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      auto fetch_g = [&](int sl) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int64_t m = m_base + it * 8 + es.rsel;
          gr[sl & 1][it] = m < M ? ld4(dy + m * ld_dy + 16 * sl + 4 * es.c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      fetch_x(0);
      if (lane == 0 && sw == 0) trace(1, i, 15, 0);
      mbar_wait(bar.acc2_full(t), (i >> 1) & 1);
      tc_fence_after();
      if (lane == 0 && sw == 0) trace(1, i, 15, 1);
      // The row's RMS statistic is re-derived here from the x values this pass reads anyway: the
      // producers' rstd_s buffer of this parity is rewritten for tile i + 2 as soon as GEMM1 of tile
      // i + 1 is done, and nothing orders that against this (overlapped, possibly late) store phase.
      float dot[4] = {0.f, 0.f, 0.f, 0.f}, ssx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {
        if (sl + 1 < 8) fetch_x(sl + 1);
        es.fill(tmem_base + lane_base + BWD_ACC2_COL + t * D + 16 * sl);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const float4 a = es.get(it), xv = xr[sl & 1][it];
          dot[it] += a.x * xv.x + a.y * xv.y + a.z * xv.z + a.w * xv.w;
          ssx[it] += xv.x * xv.x + xv.y * xv.y + xv.z * xv.z + xv.w * xv.w;
        }
      }
      fetch_x(0);
      fetch_g(0);
      float rr[4], kap[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        dot[it] += __shfl_xor_sync(0xffffffffu, dot[it], 1);
        dot[it] += __shfl_xor_sync(0xffffffffu, dot[it], 2);
        ssx[it] += __shfl_xor_sync(0xffffffffu, ssx[it], 1);
        ssx[it] += __shfl_xor_sync(0xffffffffu, ssx[it], 2);
        rr[it] = rsqrtf(ssx[it] * (1.0f / D) + kRmsEps);
        kap[it] = rr[it] * rr[it] * rr[it] * dot[it] * (1.0f / D);
      }
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {
        if (sl + 1 < 8) {
          fetch_x(sl + 1);
          fetch_g(sl + 1);
        }
        const int c0 = 16 * sl + 4 * es.c4;
        es.fill(tmem_base + lane_base + BWD_ACC2_COL + t * D + 16 * sl);
        if (sl == 7) {
          tc_fence_before();
          mbar_arrive(bar.acc2_empty(t));
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int64_t m = m_base + it * 8 + es.rsel;
          if (m >= M) continue;
          const float4 a = es.get(it), xv = xr[sl & 1][it], g = gr[sl & 1][it];
          *reinterpret_cast<float4*>(dx + m * ld_dx + c0) =
              make_float4(g.x + rr[it] * a.x - xv.x * kap[it], g.y + rr[it] * a.y - xv.y * kap[it],
                          g.z + rr[it] * a.z - xv.z * kap[it], g.w + rr[it] * a.w - xv.w * kap[it]);
        }
      }
      if (lane == 0 && sw == 0) trace(1, i, 15, 2);
    }
  };

  if (warp >= FIRST_PROD_WARP && warp < TMA_WARP) {
    // ============================================================ activation producers
    const int pw = warp - FIRST_PROD_WARP;
    for (int i = 0; i < sched.count; ++i) {
      if (lane == 0 && pw == 0) trace(3, i, 0, 0);
      mbar_wait(bar.x_empty(0), (i & 1) ^ 1);
      if (lane == 0 && pw == 0) trace(3, i, 0, 1);
      // eight TMA boxes of [128 rows x 32 floats] land the fp32 rows as [hi tile row | lo tile
      // row] per k-chunk (rows beyond M arrive as zeros); one thread issues them
      if (pw == 0 && elect_one()) {
        mbar_expect_tx(landed_bar, 8 * TILE);
        const int m0 = (int)sched.m0(i);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          tma_load_2d(smem_base + BWD_X_OFF + q * TILE, &map_x, 32 * q, m0, landed_bar);
          tma_load_2d(smem_base + BWD_X_OFF + (4 + q) * TILE, &map_dy, 32 * q, m0, landed_bar);
        }
        // the next tile's boxes are pulled into L2 while this one is processed: all CTAs reach
        // their tile boundary together, and a 19 MB burst straight from DRAM would cost ~3 us
        if (i + 1 < sched.count) {
          const int m1 = (int)sched.m0(i + 1);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            tma_prefetch_2d(&map_x, 32 * q, m1);
            tma_prefetch_2d(&map_dy, 32 * q, m1);
          }
        }
      }
      __syncwarp();
      mbar_wait(landed_bar, i & 1);
      if (lane == 0 && pw == 0) trace(3, i, 0, 2);
      float ss[2][8] = {};
      convert_tile<true>(smem + BWD_X_OFF, pw, lane, ss);
      store_rstd(rstd_s + (i & 1) * BM, pw, lane, ss);
      fence_proxy_async();
      mbar_arrive(bar.x_full(0));
      if (lane == 0 && pw == 0) trace(3, i, 0, 3);
      if (i > 0) store_tile(i - 1);   // overlaps the chunk loop of tile i
    }
    if (sched.count > 0) store_tile(sched.count - 1);
  } else if (warp == TMA_WARP) {
    // ============================================================ weight-stage producer
    // image order: G1 stages of chunk 0; for c >= 1: G1 stages of c, then WT(c-1); WT(last).
    // G1 stages (W_in k-halves, W_out^T chunk) -> ring A, WT stages -> ring B.
    if (elect_one()) {
      Ring ra, rb;
      const int n_stage = bwd_stages(F), tail = n_stage - 2;
      for (int i = 0; i < sched.count; ++i)
        for (int st = 0; st < n_stage; ++st) {
          const bool is_b = st >= tail || (st >= 3 && (st - 3) % 5 >= 3);
          Ring& r = is_b ? rb : ra;
          const int slot = is_b ? BWD_RING_A + r.stage : r.stage;
          mbar_wait(bar.w_empty(slot), r.phase ^ 1);
          mbar_expect_tx(bar.w_full(slot), STAGE);
          bulk_g2s(smem_base + BWD_RING_OFF + (uint32_t)slot * STAGE, image + (size_t)st * STAGE, STAGE,
                   bar.w_full(slot));
          r.advance(is_b ? BWD_RING_B : BWD_RING_A);
        }
    }
  } else if (warp == MMA_WARP) {
    // ============================================================ GEMM1 issuer (one thread)
    if (elect_one()) {
      constexpr uint32_t idesc_ug = make_idesc(BM, 64), idesc_ds = make_idesc(BM, 32);
      const uint32_t ring_u32 = smem_base + BWD_RING_OFF;
      const uint32_t xt = smem_base + BWD_X_OFF, dyt = xt + 4 * TILE;
      Ring ring;
      for (int i = 0; i < sched.count; ++i) {
        mbar_wait(bar.x_full(0), i & 1);
        for (int c = 0; c < nch; ++c) {
          const int b = c & 1;
          const int n = i * nch + c;
          trace(0, i, c, 0);
          if (n > 0) mbar_wait(bar.acc1_empty(b ^ 1), ((uint32_t)(n - 1) >> 1) & 1);
          tc_fence_after();
          trace(0, i, c, 1);
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {   // ug = X . W_in[chunk]^T
            mbar_wait(bar.w_full(ring.stage), ring.phase);
            const uint32_t st = ring_u32 + ring.stage * STAGE;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              mma3_ss(tmem_base, xt + (kh * 2) * TILE + kk * 32, xt + (kh * 2 + 1) * TILE + kk * 32, st + kk * 32,
                      st + 8192 + kk * 32, idesc_ug, (kh | kk) != 0);
            tc_commit(bar.w_empty(ring.stage));
            ring.advance(BWD_RING_A);
          }
          trace(0, i, c, 6);
          mbar_wait(bar.w_full(ring.stage), ring.phase);   // ds = dY . W_out[:, chunk]
          trace(0, i, c, 2);
          {
            const uint32_t st = ring_u32 + ring.stage * STAGE;
#pragma unroll
            for (int kh = 0; kh < 2; ++kh)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma3_ss(tmem_base + 64, dyt + (kh * 2) * TILE + kk * 32, dyt + (kh * 2 + 1) * TILE + kk * 32,
                        st + kh * 8192 + kk * 32, st + kh * 8192 + 4096 + kk * 32, idesc_ds, (kh | kk) != 0);
            tc_commit(bar.w_empty(ring.stage));
            ring.advance(BWD_RING_A);
          }
          tc_commit(bar.acc1_full(b));
          if (c == nch - 1) tc_commit(bar.x_empty(0));
          trace(0, i, c, 7);
        }
      }
    }
  } else if (warp == BWD_MMA2_WARP) {
    // ============================================================ GEMM2 issuer (one thread)
    if (elect_one()) {
      constexpr uint32_t idesc2 = make_idesc(BM, D);
      const uint32_t ring_u32 = smem_base + BWD_RING_OFF + BWD_RING_A * STAGE;
      Ring ring;
      for (int i = 0; i < sched.count; ++i) {
        for (int c = 0; c < nch; ++c) {
          const int b = c & 1;
          const uint32_t u = (uint32_t)(i * nch + c) >> 1;
          mbar_wait(bar.w_full(BWD_RING_A + ring.stage), ring.phase);
          const int s_hi = ring.stage;
          ring.advance(BWD_RING_B);
          mbar_wait(bar.w_full(BWD_RING_A + ring.stage), ring.phase);
          const int s_lo = ring.stage;
          ring.advance(BWD_RING_B);
          trace(0, i, c, 3);
          mbar_wait(bar.a2_full(b), u & 1);
          if (c == 0) mbar_wait(bar.acc2_empty(i & 1), ((i >> 1) & 1) ^ 1);
          tc_fence_after();
          trace(0, i, c, 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t a_hi = tmem_base + BWD_A2_COL + b * 64 + kk * 8;
            mma3_ts(tmem_base + BWD_ACC2_COL + (i & 1) * D, a_hi, a_hi + 32, ring_u32 + s_hi * STAGE + kk * 32,
                    ring_u32 + s_lo * STAGE + kk * 32, idesc2, (c | kk) != 0);
          }
          tc_commit(bar.a2_empty(b));
          tc_commit(bar.w_empty(BWD_RING_A + s_hi));
          tc_commit(bar.w_empty(BWD_RING_A + s_lo));
          trace(0, i, c, 8);
        }
        tc_commit(bar.acc2_full(i & 1));
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ============================================================ swiglu' epilogues
    // all eight warps work on every chunk: group `half` (four warps = 128 rows) takes the
    // 16-unit half `half` of the chunk, so the single acc1 buffer is released as soon as each
    // thread has pulled its 48 accumulator columns into registers
    const int half = warp >> 2;
    for (int i = 0; i < sched.count; ++i) {
      {  // dY rows 16 warp .. + 15 -> bf16 hi / lo operand tiles (both k-chunks)
        mbar_wait(landed_bar, i & 1);
        float unused[8];
        convert_batch<false>(smem + BWD_X_OFF + 4 * TILE, warp * 16, lane, unused);
        convert_batch<false>(smem + BWD_X_OFF + 6 * TILE, warp * 16, lane, unused);
        fence_proxy_async();
        mbar_arrive(bar.x_full(0));
      }
      mbar_wait(bar.x_full(0), i & 1);
      const float rs = rstd_s[(i & 1) * BM + quarter * 32 + lane];
      for (int c = 0; c < nch; ++c) {
        const int b = c & 1;
        const uint32_t u = (uint32_t)(i * nch + c) >> 1;
        if (lane == 0 && quarter == 0) trace(1 + half, i, c, 0);
        mbar_wait(bar.acc1_full(b), u & 1);
        tc_fence_after();
        if (lane == 0 && quarter == 0) trace(1 + half, i, c, 1);
        float v[32], ds[16];
        tmem_ld32(tmem_base + lane_base + half * 32, v);
        tmem_ld16(tmem_base + lane_base + 64 + half * 16, ds);
        tc_fence_before();
        mbar_arrive(bar.acc1_empty(b));
        const float* bv = bias_s + c * CH + half * 16;
        // K order of GEMM2: [d_v 0..15 | d_g 0..15] of this half
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float dv[2], dg[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int k = 2 * q + e;
            const float val = rs * v[k] + bv[k];
            const float sg = fsigmoid(rs * v[16 + k] + bv[F + k]);
            dv[e] = ds[k] * sg;
            dg[e] = ds[k] * val * sg * (1.0f - sg);
          }
          hi[q] = pack_bf16(dv[0], dv[1]);
          lo[q] = pack_bf16(dv[0] - __uint_as_float(hi[q] << 16), dv[1] - __uint_as_float(hi[q] & 0xffff0000u));
          hi[8 + q] = pack_bf16(dg[0], dg[1]);
          lo[8 + q] = pack_bf16(dg[0] - __uint_as_float(hi[8 + q] << 16),
                                dg[1] - __uint_as_float(hi[8 + q] & 0xffff0000u));
        }
        mbar_wait(bar.a2_empty(b), (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t a2 = tmem_base + lane_base + BWD_A2_COL + b * 64 + half * 16;
        tmem_st16(a2, hi);
        tmem_st16(a2 + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar.a2_full(b));
        if (lane == 0 && quarter == 0) trace(1 + half, i, c, 4);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

int check_dims(const char* what, int d, int d_ff) {
  if (d != D || d_ff % 64 != 0 || d_ff < 64 || d_ff > MAX_F) {
    set_error("%s: built for d_pet = %d and a hidden width that is a multiple of 64 up to %d (got %d, %d)",
              what, D, MAX_F, d, d_ff);
    return PETB200_ERR_UNSUPPORTED;
  }
  return PETB200_OK;
}

}  // namespace
}  // namespace petb200

using namespace petb200;

// debugging aid, not part of the documented ABI: buffer of 4 roles x 4 tiles x 16 chunks x 16
// int64 slots that CTA 0 of petb200_mlp_fwd fills with clock64() stamps (NULL switches it off)
extern "C" PETB200_API int petb200_debug_mlp_trace(void* buffer) {
#ifdef PETB200_MLP_TRACE
  cudaMemcpyToSymbol(g_trace, &buffer, sizeof(void*));
  return check_launch("debug_mlp_trace");
#else
  (void)buffer;
  set_error("debug_mlp_trace: rebuild with -DPETB200_MLP_TRACE");
  return PETB200_ERR_UNSUPPORTED;
#endif
}

extern "C" PETB200_API size_t petb200_mlp_image_bytes(int d_ff, int backward) {
  return (size_t)(backward ? bwd_stages(d_ff) : fwd_stages(d_ff)) * STAGE;
}

extern "C" PETB200_API int petb200_mlp_pack(const float* w_in, const float* w_out, int d, int d_ff,
                                            void* image_fwd, void* image_bwd, cudaStream_t stream) {
  if (int rc = check_dims("mlp_pack", d, d_ff)) return rc;
  for (int backward = 0; backward < 2; ++backward) {
    void* image = backward ? image_bwd : image_fwd;
    if (!image) continue;
    const int64_t chunks = (int64_t)petb200_mlp_image_bytes(d_ff, backward) / 16;
    mlp_pack_kernel<<<(unsigned)ceil_div(chunks, 256), 256, 0, stream>>>(w_in, w_out, d_ff, backward,
                                                                        reinterpret_cast<uint4*>(image));
  }
  return check_launch("mlp_pack");
}

extern "C" PETB200_API int petb200_mlp_fwd(const float* x, int64_t ldx, const void* image_fwd,
                                           const float* b_in, const float* b_out, int64_t n_rows, int d,
                                           int d_ff, float* y, int64_t ldy, cudaStream_t stream) {
  if (int rc = check_dims("mlp_fwd", d, d_ff)) return rc;
  PETB200_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "mlp_fwd: leading dimensions must be multiples of 4");
  if (n_rows == 0) return PETB200_OK;
  const int tiles = (int)ceil_div(n_rows, BM);
  cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM);
  mlp_fwd_kernel<<<tiles < kNumSMs ? tiles : kNumSMs, FWD_NUM_THREADS, FWD_SMEM, stream>>>(
      x, ldx, reinterpret_cast<const uint8_t*>(image_fwd), b_in, b_out, n_rows, d_ff, y, ldy);
  return check_launch("mlp_fwd");
}

extern "C" PETB200_API int petb200_mlp_bwd(const float* x, int64_t ldx, const float* d_y, int64_t ld_dy,
                                           const void* image_bwd, const float* b_in, int64_t n_rows, int d,
                                           int d_ff, float* d_x, int64_t ld_dx, cudaStream_t stream) {
  if (int rc = check_dims("mlp_bwd", d, d_ff)) return rc;
  PETB200_REQUIRE(ldx % 4 == 0 && ld_dy % 4 == 0 && ld_dx % 4 == 0,
                  "mlp_bwd: leading dimensions must be multiples of 4");
  if (n_rows == 0) return PETB200_OK;
  const int tiles = (int)ceil_div(n_rows, BM);
  CUtensorMap map_x, map_dy;
  if (make_tma_map_f32(&map_x, x, n_rows, d, ldx, 32, BM) || make_tma_map_f32(&map_dy, d_y, n_rows, d, ld_dy, 32, BM)) {
    set_error("mlp_bwd: cuTensorMapEncodeTiled failed (x / d_y must be 16-byte aligned)");
    return PETB200_ERR_CUDA;
  }
  cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
  mlp_bwd_kernel<<<tiles < kNumSMs ? tiles : kNumSMs, BWD_NUM_THREADS, BWD_SMEM, stream>>>(
      map_x, map_dy, x, ldx, d_y, ld_dy, reinterpret_cast<const uint8_t*>(image_bwd), b_in, n_rows, d_ff, d_x, ld_dx);
  return check_launch("mlp_bwd");
}

extern "C" PETB200_API size_t petb200_norm_linear_image_bytes(int n_out) {
  return (size_t)(n_out / 64) * 2 * STAGE;
}

extern "C" PETB200_API int petb200_norm_linear_pack(const float* w, int d, int n_out, void* image,
                                                    cudaStream_t stream) {
  PETB200_REQUIRE(d == D && n_out % 64 == 0 && n_out >= 64 && n_out <= NL_MAX_N,
                  "norm_linear_pack: built for d = %d and n_out a multiple of 64 up to %d (got %d, %d)", D,
                  NL_MAX_N, d, n_out);
  const int64_t chunks = (int64_t)petb200_norm_linear_image_bytes(n_out) / 16;
  linear_pack_kernel<<<(unsigned)ceil_div(chunks, 256), 256, 0, stream>>>(w, n_out, reinterpret_cast<uint4*>(image));
  return check_launch("norm_linear_pack");
}

extern "C" PETB200_API int petb200_norm_linear(const float* x, int64_t ldx, const void* image, const float* bias,
                                               int64_t n_rows, int d, int n_out, float* out, int64_t ldo,
                                               float* rstd_out, cudaStream_t stream) {
  PETB200_REQUIRE(d == D && n_out % 64 == 0 && n_out >= 64 && n_out <= NL_MAX_N,
                  "norm_linear: built for d = %d and n_out a multiple of 64 up to %d (got %d, %d)", D, NL_MAX_N,
                  d, n_out);
  PETB200_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0, "norm_linear: leading dimensions must be multiples of 4");
  if (n_rows == 0) return PETB200_OK;
  const int tiles = (int)ceil_div(n_rows, BM);
  CUtensorMap map_out;
  if (make_tma_map_f32(&map_out, out, n_rows, n_out, ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B) != 0) {
    set_error("norm_linear: cuTensorMapEncodeTiled failed (output pointer must be 16-byte aligned)");
    return PETB200_ERR_CUDA;
  }
  cudaFuncSetAttribute(norm_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NL_SMEM);
  norm_linear_kernel<<<tiles < kNumSMs ? tiles : kNumSMs, NUM_THREADS, NL_SMEM, stream>>>(
      map_out, x, ldx, reinterpret_cast<const uint8_t*>(image), bias, n_rows, n_out, rstd_out);
  return check_launch("norm_linear");
}
