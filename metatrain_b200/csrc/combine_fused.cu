// Fused message update of a PET GNN layer ("combine block"), forward and backward, as single
// persistent tcgen05 kernels (sm_100a).
//
//   forward :  m'_e = m_e + t_e + W_b . silu(W_a . LayerNorm_256(cat[t_e, t_rev(e)]) + b_a) + b_b
//   backward:  d_cat_e = LayerNorm'(cat_e)^T [ W_a^T . silu'(p_e) (W_b^T . g_e) ]        (g = d m')
// (PETBackend._feedforward_featurization_impl, src/metatrain/pet/modules/backend.py:559-575;
//  modules :95-107: combination_norms = LayerNorm(2 d_pet), combination_mlps = Linear(2d, 2d) ->
//  SiLU -> Linear(2d, d).)
//
// The unfused path (combine_ln_fwd + two GEMMs) moved 7.7 KB per edge and GNN layer through HBM
// for this block: the [E, 256] concatenation, the pre-activations and the activations were each
// written and read back.  Here a 128-edge tile is gathered once (own rows + the rows of the
// reversed edges, the "edge scatter" of the message passing), the LayerNorm is folded into the
// first contraction,
//     W_a . LN(c) + b_a = r (W' c - mu s) + b',   W' = W_a diag(gamma), s = W' 1, b' = W_a beta + b_a,
// (mu, r = mean and 1/std of the 256 gathered values, computed by the producer threads while they
// convert the rows), the hidden activations never leave the SM, and only the pre-activations p
// (1 KB per edge, needed by the backward) and the (mu, r) pair are written besides m'.  Forward
// traffic: 2 KB read + 1.5 KB written per edge.  The backward reads g, p and the gathered rows and
// writes d_cat; the two LayerNorm-backward row statistics come out of the first epilogue for free:
//     mean(z) = d_p . s / 256,   mean(z x_hat) = d_p . (p - b') / 256     (z = W'^T d_p).
//
// Both kernels reuse the structure of mlp_fused.cu: A operands in tensor memory (lane = row),
// weights streamed from L2 as 16 KB stages of a pre-swizzled image (petb200_combine_pack), hidden
// dimension walked in chunks of 32 units.  15 warps: 0-7 activation epilogues (two groups on
// alternate chunks; they also run the output store of the tile, which falls into the row-producer
// phase of the next tile, when they would idle — 480 threads leave them 126 registers, no spills),
// 8 GEMM1 issue, 9-12 row producers, 13 weight stages, 14 GEMM2 issue.  All products use the bf16
// hi/lo 2-term split (fp32 accumulation in TMEM).
#include "fused_common.cuh"

namespace petb200 {
namespace {

using namespace tc;
using namespace fused;

constexpr int HID = 256;            // hidden width of the combine MLP (= 2 d_pet) and LayerNorm width
constexpr int NCH = HID / CH;       // 8 chunks of 32 hidden units
constexpr int XPITCH = D + 4;                                // floats per staged row
constexpr int CB_STAGING_BYTES = BM * XPITCH * 4;            // 67 584
constexpr float kLnEps = 1e-5f;                              // torch.nn.LayerNorm default

// Optional in-kernel timeline (tools/combine_trace.py; build with -DPETB200_COMBINE_TRACE): CTA 0
// records clock64() at pipeline events of its first tiles.  Compiled out otherwise.
#ifdef PETB200_COMBINE_TRACE
__device__ long long* g_ctrace = nullptr;
__device__ __forceinline__ void ctrace(int role, int tile, int chunk, int ev) {
  if (g_ctrace != nullptr && blockIdx.x == 0 && tile >= 2 && tile < 6)
    g_ctrace[((role * 4 + (tile - 2)) * 16 + chunk) * 8 + ev] = clock64();
}
#else
__device__ __forceinline__ void ctrace(int, int, int, int) {}
#endif

__host__ __device__ constexpr int cf_stages() { return (HID / 64) * 6; }   // forward image
__host__ __device__ constexpr int cb_stages() { return (HID / 64) * 6; }   // backward image

// ----------------------------------------------------------------------- weight images
// forward, per pair of chunks p (hidden units 64 p .. 64 p + 63):
//   W1(2p, 0) W1(2p, 1) W1(2p+1, 0) W1(2p+1, 1) W2hi(p) W2lo(p)
//   W1(c, j): k-quarters 2j, 2j+1 of W'[32 c .. 32 c + 31, :]: per quarter [32 rows x 64 k] hi 4 KB | lo 4 KB
//   W2(p)   : W_b[:, 64 p .. 64 p + 63] as [128 rows x 64 k]
// backward, per pair of chunks p:
//   G1(2p) G1(2p+1) WT(p, hi, 0) WT(p, hi, 1) WT(p, lo, 0) WT(p, lo, 1)
//   G1(c)   : W_b^T[32 c .. 32 c + 31, :]: per k-half [32 rows (hidden) x 64 k (out dim)] hi 4 KB | lo 4 KB
//   WT(p, ., h): W'^T[128 h .. 128 h + 127, 64 p .. 64 p + 63] as [128 rows (input dim) x 64 k (hidden)]
__global__ void combine_pack_kernel(const float* __restrict__ w_a, const float* __restrict__ w_b, int backward,
                                    uint4* __restrict__ image) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)cf_stages() * (STAGE / 16)) return;
  const int s = (int)(idx / (STAGE / 16));
  const int o = (int)(idx % (STAGE / 16)) * 16;  // byte offset inside the stage
  const int p = s / 6, r = s % 6;
  float v[8];
  bool lo;
  if (!backward) {
    if (r < 4) {
      const int c = 2 * p + r / 2, kq = 2 * (r % 2) + (o >> 13), t = o & 8191;
      lo = t >= 4096;
      const int u = t & 4095, n = (u >> 10) * 8 + ((u >> 7) & 7), j = ((u >> 4) & 7) ^ (n & 7);
      const float* src = w_a + (int64_t)(c * CH + n) * HID + kq * 64 + j * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = src[e];
    } else {
      lo = r == 5;
      const int n = (o >> 10) * 8 + ((o >> 7) & 7), j = ((o >> 4) & 7) ^ (n & 7);
      const float* src = w_b + (int64_t)n * HID + p * 64 + j * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = src[e];
    }
  } else {
    if (r < 2) {
      const int c = 2 * p + r, kh = o >> 13, t = o & 8191;
      lo = t >= 4096;
      const int u = t & 4095, n = (u >> 10) * 8 + ((u >> 7) & 7), j = ((u >> 4) & 7) ^ (n & 7);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = w_b[(int64_t)(kh * 64 + j * 8 + e) * HID + c * CH + n];
    } else {
      lo = r >= 4;
      const int half = r & 1;
      const int n = (o >> 10) * 8 + ((o >> 7) & 7), j = ((o >> 4) & 7) ^ (n & 7);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = w_a[(int64_t)(p * 64 + j * 8 + e) * HID + half * 128 + n];
    }
  }
  image[idx] = pack8(v, lo);
}

// fp32 row (32 floats at `row`) -> 16 packed bf16 hi columns + 16 lo columns; s1 / s2 accumulate the
// sum and the sum of squares of (x - shift) (shifted data: no cancellation for rows with a large mean)
template <bool STATS>
__device__ __forceinline__ void split_part(const float* row, uint32_t (&hi)[16], uint32_t (&lo)[16],
                                           float shift, float& s1, float& s2) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(row + 4 * q);
    if (STATS) {
      const float a = v.x - shift, b = v.y - shift, c = v.z - shift, d = v.w - shift;
      s1 += (a + b) + (c + d);
      s2 = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, s2))));
    }
    hi[2 * q] = pack_bf16(v.x, v.y);
    hi[2 * q + 1] = pack_bf16(v.z, v.w);
    lo[2 * q] = pack_bf16(v.x - __uint_as_float(hi[2 * q] << 16), v.y - __uint_as_float(hi[2 * q] & 0xffff0000u));
    lo[2 * q + 1] = pack_bf16(v.z - __uint_as_float(hi[2 * q + 1] << 16),
                              v.w - __uint_as_float(hi[2 * q + 1] & 0xffff0000u));
  }
}

// ======================================================================== forward
// 15 warps: 0-7 epilogue groups (LayerNorm fold + SiLU per chunk, then the output store of the tile:
// the store falls into the row-producer phase of the NEXT tile, when these warps would idle), 8 GEMM1
// issue, 9-12 row producers, 13 weight stages, 14 GEMM2 issue
constexpr int CF_NUM_THREADS = 32 * 15;
constexpr int CF_PROD_WARPS = 4, CF_TMA_WARP = FIRST_PROD_WARP + CF_PROD_WARPS, CF_MMA2_WARP = CF_TMA_WARP + 1;
constexpr int CF_RING_A = 5, CF_RING_B = 3;
constexpr int CF_RING = CF_RING_A + CF_RING_B;
constexpr int CF_XS_OFF = 0;
constexpr int CF_RING_OFF = ((CF_XS_OFF + CB_STAGING_BYTES + 1023) / 1024) * 1024;
constexpr int CF_EPI_OFF = CF_RING_OFF + CF_RING * STAGE;            // 8 epilogue warps x 32 x STAGE_LD floats
constexpr int CF_CONST_OFF = CF_EPI_OFF + EPI_STAGE_BYTES;           // s [256], b' [256], b_b [128]
constexpr int CF_STAT_OFF = CF_CONST_OFF + (2 * HID + D) * 4;        // (mu, r) [2 tile parities][128]
constexpr int CF_PART_OFF = CF_STAT_OFF + 2 * BM * 8;                // partial row sums exchanged by producer pairs
constexpr int CF_BAR_OFF = CF_PART_OFF + BM * 8;
constexpr int CF_SMEM = CF_BAR_OFF + 8 * (16 + 2 * CF_RING) + 16 + 1024;
static_assert(CF_SMEM <= 232448, "combine_fwd: shared memory budget");
// TMEM columns: c hi 0..127 (own half 0..63, reversed half 64..127), c lo 128..255 ; acc1[b] at
// 256 + 32 b ; A2[b] at 320 + 32 b (hi 16 | lo 16) ; acc2 at 384 (single buffer)
constexpr int CF_XLO_COL = 128, CF_ACC1_COL = 256, CF_A2_COL = 320, CF_ACC2_COL = 384;

__global__ void __launch_bounds__(CF_NUM_THREADS, 1)
combine_fwd_kernel(const float* __restrict__ t, int64_t ld_t, const int32_t* __restrict__ rev,
                   const uint8_t* __restrict__ image, const float* __restrict__ s_vec,
                   const float* __restrict__ b_fold, const float* __restrict__ b_out, int64_t M,
                   float* __restrict__ m_io, int64_t ld_m, float* __restrict__ p_out,
                   float2* __restrict__ stats_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const Barriers bar{smem_base + CF_BAR_OFF, CF_RING};
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + CF_BAR_OFF + 8 * (16 + 2 * CF_RING));
  float* const_s = reinterpret_cast<float*>(smem + CF_CONST_OFF);
  float2* stat_s = reinterpret_cast<float2*>(smem + CF_STAT_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;  // TMEM lane quarter this warp may access
  const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
  const TileSchedule sched(M);

  if (threadIdx.x == 0) {
    bar.init_all();   // acc2_empty: drained by all eight epilogue warps
    mbar_init(bar.x_full(0), CF_PROD_WARPS * 32);
    mbar_init(bar.x_full(1), CF_PROD_WARPS * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(const_cast<uint32_t*>(tmem_slot))));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < HID; i += CF_NUM_THREADS) {
    const_s[i] = s_vec[i];
    const_s[HID + i] = b_fold[i];
  }
  for (int i = threadIdx.x; i < D; i += CF_NUM_THREADS) const_s[2 * HID + i] = b_out[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= FIRST_PROD_WARP && warp < CF_TMA_WARP) {
    // ============================================================ row producers
    // warp -> rows 32 quarter .. + 31 (the TMEM lanes it may write); lane -> one row.  Two phases
    // per tile through one staging buffer: own rows, then the rows of the reversed edges.
    // (Eight producer warps — two per lane quarter, half a row each — were tried: the register cap
    // of a 608-thread CTA made the epilogue warps spill and the chunk loop lost more than the
    // conversion gained, profiles/r2_combine_fwd_timeline.txt.)
    const uint32_t dst = smem_base + CF_XS_OFF;
    auto issue_own = [&](int i) {
      const int64_t m0 = sched.m0(i);
#pragma unroll 8
      for (int it = 0; it < 32; ++it) {  // one row (32 x 16 B) per instruction
        const int row = quarter * 32 + it;
        const int64_t m = m0 + row;
        const bool ok = m < M;
        cp_async16(dst + (uint32_t)(row * XPITCH * 4 + lane * 16), t + (ok ? m : 0) * ld_t + 4 * lane,
                   ok ? 16u : 0u);
      }
      cp_async_commit();
    };
    auto issue_rev = [&](int i) {
      const int64_t mine = sched.m0(i) + quarter * 32 + lane;
      const int my_rev = mine < M ? __ldg(rev + mine) : -1;
#pragma unroll 8
      for (int it = 0; it < 32; ++it) {
        const int row = quarter * 32 + it;
        const int r = __shfl_sync(0xffffffffu, my_rev, it);
        cp_async16(dst + (uint32_t)(row * XPITCH * 4 + lane * 16), t + (int64_t)(r >= 0 ? r : 0) * ld_t + 4 * lane,
                   r >= 0 ? 16u : 0u);
      }
      cp_async_commit();
    };
    const float* row = reinterpret_cast<const float*>(smem + CF_XS_OFF) + (quarter * 32 + lane) * XPITCH;
    // staged fp32 row -> bf16 hi / lo columns of the half in TMEM; returns the row's mean and sum of
    // squared deviations (one pass, data shifted by the row's first element)
    auto park = [&](int half, float& mean, float& m2) {
      const float shift = row[0];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int part = 0; part < 4; ++part) {
        uint32_t hi[16], lo[16];
        split_part<true>(row + part * 32, hi, lo, shift, s1, s2);
        tmem_st16(tmem_base + lane_base + half * 64 + part * 16, hi);
        tmem_st16(tmem_base + lane_base + CF_XLO_COL + half * 64 + part * 16, lo);
      }
      mean = shift + s1 * (1.0f / D);
      m2 = fmaxf(s2 - s1 * s1 * (1.0f / D), 0.f);
    };
    // the gathered rows of a tile are pulled into L2 one tile ahead (one 512 B row per lane), so the
    // gather in the middle of the tile's critical path is served from L2, not DRAM
    auto prefetch_rev = [&](int i) {
      const int64_t mine = sched.m0(i) + quarter * 32 + lane;
      if (mine < M) prefetch_l2_bulk(t + (int64_t)__ldg(rev + mine) * ld_t, D * 4);
    };
    if (sched.count > 0) {
      issue_own(0);
      prefetch_rev(0);
    }
    for (int i = 0; i < sched.count; ++i) {
      if (i + 1 < sched.count) prefetch_rev(i + 1);
      cp_async_wait_group<0>();
      __syncwarp();
      if (lane == 0 && quarter == 0) ctrace(3, i, 0, 0);
      mbar_wait(bar.x_empty(0), (i & 1) ^ 1);   // GEMM1 of the previous tile has consumed c
      tc_fence_after();
      if (lane == 0 && quarter == 0) ctrace(3, i, 0, 1);
      float mean1, m2_1;
      park(0, mean1, m2_1);
      __syncwarp();            // the warp's rows of the staging buffer are consumed
      issue_rev(i);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar.x_full(0));
      if (lane == 0 && quarter == 0) ctrace(3, i, 0, 2);
      cp_async_wait_group<0>();
      __syncwarp();
      if (lane == 0 && quarter == 0) ctrace(3, i, 0, 3);
      float mean2, m2_2;
      park(1, mean2, m2_2);
      // LayerNorm statistics of the 256 gathered values (pairwise combination of the two halves)
      const float mu = 0.5f * (mean1 + mean2), dm = mean1 - mean2;
      const float var = (m2_1 + m2_2 + dm * dm * (0.25f * HID)) * (1.0f / HID);
      const float2 st = make_float2(mu, rsqrtf(var + kLnEps));
      stat_s[(i & 1) * BM + quarter * 32 + lane] = st;
      const int64_t m = sched.m0(i) + quarter * 32 + lane;
      if (m < M) stats_out[m] = st;
      __syncwarp();
      if (i + 1 < sched.count) issue_own(i + 1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar.x_full(1));
      if (lane == 0 && quarter == 0) ctrace(3, i, 0, 4);
    }
  } else if (warp == CF_TMA_WARP) {
    // ============================================================ weight-stage producer
    if (elect_one()) {
      Ring ra, rb;
      for (int i = 0; i < sched.count; ++i)
        for (int st = 0; st < cf_stages(); ++st) {
          const bool is_b = (st % 6) >= 4;
          Ring& r = is_b ? rb : ra;
          const int slot = is_b ? CF_RING_A + r.stage : r.stage;
          mbar_wait(bar.w_empty(slot), r.phase ^ 1);
          mbar_expect_tx(bar.w_full(slot), STAGE);
          bulk_g2s(smem_base + CF_RING_OFF + (uint32_t)slot * STAGE, image + (size_t)st * STAGE, STAGE,
                   bar.w_full(slot));
          r.advance(is_b ? CF_RING_B : CF_RING_A);
        }
    }
  } else if (warp == MMA_WARP) {
    // ============================================================ GEMM1 issuer (one thread)
    // acc1[b] [128 x 32] = c [128 x 256] . W'[chunk]^T : two stages per chunk (own half, reversed half)
    if (elect_one()) {
      constexpr uint32_t idesc1 = make_idesc(BM, CH);
      const uint32_t ring_u32 = smem_base + CF_RING_OFF;
      Ring ring;
      for (int i = 0; i < sched.count; ++i) {
        mbar_wait(bar.x_full(0), i & 1);
        for (int c = 0; c < NCH; ++c) {
          const uint32_t n = (uint32_t)(i * NCH + c);
          const int b = c & 1;
          ctrace(0, i, c, 0);
          mbar_wait(bar.acc1_empty(b), ((n >> 1) & 1) ^ 1);
          tc_fence_after();
          ctrace(0, i, c, 1);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (c == 0 && j == 1) {
              mbar_wait(bar.x_full(1), i & 1);
              tc_fence_after();
            }
            mbar_wait(bar.w_full(ring.stage), ring.phase);
            ctrace(0, i, c, 2 + j);
            const uint32_t st = ring_u32 + ring.stage * STAGE;
#pragma unroll
            for (int kq = 0; kq < 2; ++kq)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint32_t a_hi = tmem_base + (j * 8 + kq * 4 + kk) * 8;
                mma3_ts(tmem_base + CF_ACC1_COL + b * CH, a_hi, a_hi + CF_XLO_COL, st + kq * 8192 + kk * 32,
                        st + kq * 8192 + 4096 + kk * 32, idesc1, (j | kq | kk) != 0);
              }
            tc_commit(bar.w_empty(ring.stage));
            ring.advance(CF_RING_A);
          }
          tc_commit(bar.acc1_full(b));
          ctrace(0, i, c, 4);
          if (c == NCH - 1) tc_commit(bar.x_empty(0));
        }
      }
    }
  } else if (warp == CF_MMA2_WARP) {
    // ============================================================ GEMM2 issuer (one thread)
    if (elect_one()) {
      constexpr uint32_t idesc2 = make_idesc(BM, D);
      const uint32_t ring_u32 = smem_base + CF_RING_OFF + CF_RING_A * STAGE;
      Ring ring;
      int w2_hi = 0, w2_lo = 0;
      for (int i = 0; i < sched.count; ++i) {
        for (int c = 0; c < NCH; ++c) {
          const int b = c & 1;
          const uint32_t u = (uint32_t)(i * NCH + c) >> 1;
          if (b == 0) {
            mbar_wait(bar.w_full(CF_RING_A + ring.stage), ring.phase);
            w2_hi = ring.stage;
            ring.advance(CF_RING_B);
            mbar_wait(bar.w_full(CF_RING_A + ring.stage), ring.phase);
            w2_lo = ring.stage;
            ring.advance(CF_RING_B);
          }
          ctrace(4, i, c, 0);
          mbar_wait(bar.a2_full(b), u & 1);
          ctrace(4, i, c, 1);
          if (c == 0) mbar_wait(bar.acc2_empty(0), (i & 1) ^ 1);
          tc_fence_after();
          ctrace(4, i, c, 2);
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t koff = (uint32_t)(b * 64 + kk * 32);
            const uint32_t a_hi = tmem_base + CF_A2_COL + b * 32 + kk * 8;
            mma3_ts(tmem_base + CF_ACC2_COL, a_hi, a_hi + 16, ring_u32 + w2_hi * STAGE + koff,
                    ring_u32 + w2_lo * STAGE + koff, idesc2, (c | kk) != 0);
          }
          tc_commit(bar.a2_empty(b));
          ctrace(4, i, c, 3);
          if (b == 1) {
            tc_commit(bar.w_empty(CF_RING_A + w2_hi));
            tc_commit(bar.w_empty(CF_RING_A + w2_lo));
          }
        }
        tc_commit(bar.acc2_full(0));
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ============================================================ LayerNorm fold + SiLU epilogues
    // group `half` (four warps = 128 rows) takes the chunks of parity `half`; a thread owns one row:
    // 32 accumulator columns in, 32 pre-activations out (global) and 32 activations out (TMEM)
    const int half = warp >> 2;
    const float* s_s = const_s;
    const float* bf_s = const_s + HID;
    const float* bo_s = const_s + 2 * HID;
    const EpiStage es{reinterpret_cast<float*>(smem + CF_EPI_OFF) + warp * (32 * STAGE_LD), lane, lane & 3,
                      (lane >> 3) + 4 * ((lane >> 2) & 1)};
    for (int i = 0; i < sched.count; ++i) {
      const int64_t m_base = sched.m0(i) + quarter * 32;
      // the tile's rows of m are first touched by the store below: pull them into L2 now
      if (half == 0) {
        if (ld_m == D) {
          if (lane == 0 && m_base < M)
            prefetch_l2_bulk(m_io + m_base * ld_m, (uint32_t)((M - m_base < 32 ? M - m_base : 32) * D * 4));
        } else if (m_base + lane < M) {
          prefetch_l2_bulk(m_io + (m_base + lane) * ld_m, D * 4);
        }
      }
      mbar_wait(bar.x_full(1), i & 1);          // the tile's (mu, r) are in shared memory
      const float2 st = stat_s[(i & 1) * BM + quarter * 32 + lane];
      const float mu = st.x, rs = st.y;
      const int64_t m = sched.m0(i) + quarter * 32 + lane;
      for (int c = half; c < NCH; c += 2) {
        const int b = half;
        const uint32_t u = (uint32_t)(i * NCH + c) >> 1;
        if (lane == 0 && quarter == 0) ctrace(1 + half, i, c, 0);
        mbar_wait(bar.acc1_full(b), u & 1);
        tc_fence_after();
        if (lane == 0 && quarter == 0) ctrace(1 + half, i, c, 1);
        float v[32];
        tmem_ld32(tmem_base + lane_base + CF_ACC1_COL + b * CH, v);
        tc_fence_before();
        mbar_arrive(bar.acc1_empty(b));
        if (lane == 0 && quarter == 0) ctrace(1 + half, i, c, 2);
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = rs * (v[k] - mu * s_s[c * CH + k]) + bf_s[c * CH + k];
        {   // saved for the backward in the kernels' private layout [tile][chunk][unit][row]: the 32
            // lanes (= rows) of a warp write one 128 B line per unit
          float* dst = p_out + ((int64_t)(sched.first + i * sched.stride) * NCH + c) * (CH * BM) + quarter * 32 + lane;
#pragma unroll
          for (int k = 0; k < 32; ++k) dst[k * BM] = v[k];
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float a0 = v[2 * q] * fsigmoid(v[2 * q]), a1 = v[2 * q + 1] * fsigmoid(v[2 * q + 1]);
          hi[q] = pack_bf16(a0, a1);
          lo[q] = pack_bf16(a0 - __uint_as_float(hi[q] << 16), a1 - __uint_as_float(hi[q] & 0xffff0000u));
        }
        if (lane == 0 && quarter == 0) ctrace(1 + half, i, c, 3);
        mbar_wait(bar.a2_empty(b), (u & 1) ^ 1);
        tc_fence_after();
        if (lane == 0 && quarter == 0) ctrace(1 + half, i, c, 4);
        const uint32_t a2 = tmem_base + lane_base + CF_A2_COL + b * 32;
        tmem_st16(a2, hi);
        tmem_st16(a2 + 16, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar.a2_full(b));
        if (lane == 0 && quarter == 0) ctrace(1 + half, i, c, 5);
      }
      // ---- output store: m' = acc2 + b_b + t + m, in place on m.  Warp -> its 32 rows x the 64
      // columns of its group, in 4 slices of 16; overlaps the row-producer phase of the next tile.
      {
        float4 rt[2][4], rm[2][4];
        auto fetch = [&](int sl) {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int64_t mm = m_base + it * 8 + es.rsel;
            const bool ok = mm < M;
            rt[sl & 1][it] = ok ? ld4(t + mm * ld_t + 16 * sl + 4 * es.c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            rm[sl & 1][it] = ok ? *reinterpret_cast<const float4*>(m_io + mm * ld_m + 16 * sl + 4 * es.c4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        fetch(4 * half);
        if (lane == 0 && quarter == 0) ctrace(5, i, half, 0);
        mbar_wait(bar.acc2_full(0), i & 1);
        tc_fence_after();
        if (lane == 0 && quarter == 0) ctrace(5, i, half, 1);
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int sl = 4 * half + s4;
          if (s4 + 1 < 4) fetch(sl + 1);
          const int c0 = 16 * sl + 4 * es.c4;
          const float4 b4 = *reinterpret_cast<const float4*>(bo_s + c0);
          es.fill(tmem_base + lane_base + CF_ACC2_COL + 16 * sl);
          if (s4 == 3) {   // the accumulator is in registers / smem now: release it early
            tc_fence_before();
            mbar_arrive(bar.acc2_empty(0));
          }
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int64_t mm = m_base + it * 8 + es.rsel;
            if (mm >= M) continue;
            const float4 a = es.get(it), x = rt[sl & 1][it], y = rm[sl & 1][it];
            *reinterpret_cast<float4*>(m_io + mm * ld_m + c0) =
                make_float4(a.x + b4.x + x.x + y.x, a.y + b4.y + x.y + y.y, a.z + b4.z + x.z + y.z,
                            a.w + b4.w + x.w + y.w);
          }
        }
        if (lane == 0 && quarter == 0) ctrace(5, i, half, 2);
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

// ======================================================================= backward
// 15 warps, as in the forward kernel: the eight epilogue warps also run the LayerNorm-backward store of
// the tile (while the row producers convert the next tile), no dedicated store warps
constexpr int CBW_NUM_THREADS = 32 * 15;
constexpr int CBW_MMA2_WARP = 14;
constexpr int CBW_RING_A = 3, CBW_RING_B = 5;                 // G1 stages / WT stages
constexpr int CBW_RING = CBW_RING_A + CBW_RING_B;
constexpr int CBW_XS_OFF = 0;
constexpr int CBW_RING_OFF = ((CBW_XS_OFF + CB_STAGING_BYTES + 1023) / 1024) * 1024;
constexpr int CBW_EPI_OFF = CBW_RING_OFF + CBW_RING * STAGE;
constexpr int CBW_CONST_OFF = CBW_EPI_OFF + EPI_STAGE_BYTES;         // s [256], b' [256]
constexpr int CBW_ROWSTAT_OFF = CBW_CONST_OFF + 2 * HID * 4;         // partial (S1, S2) [2 groups][128]
constexpr int CBW_BAR_OFF = CBW_ROWSTAT_OFF + 2 * BM * 8;
constexpr int CBW_SMEM = CBW_BAR_OFF + 8 * (16 + 2 * CBW_RING) + 8 + 16 + 1024;
static_assert(CBW_SMEM <= 232448, "combine_bwd: shared memory budget");
// TMEM columns: g hi 0..63, g lo 64..127 ; acc1[b] at 128 + 32 b ; A2[b] at 192 + 32 b (hi 16 | lo 16) ;
// acc2 at 256 (256 columns: z for the own half, then for the reversed half)
constexpr int CBW_GLO_COL = 64, CBW_ACC1_COL = 128, CBW_A2_COL = 192, CBW_ACC2_COL = 256;

__global__ void __launch_bounds__(CBW_NUM_THREADS, 1)
combine_bwd_kernel(const float* __restrict__ g, int64_t ld_g, const float* __restrict__ p,
                   const float* __restrict__ t, int64_t ld_t, const int32_t* __restrict__ rev,
                   const float2* __restrict__ stats, const uint8_t* __restrict__ image,
                   const float* __restrict__ s_vec, const float* __restrict__ b_fold, int64_t M,
                   float* __restrict__ d_cat) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const Barriers bar{smem_base + CBW_BAR_OFF, CBW_RING};
  const uint32_t rowstat_bar = bar.at(16 + 2 * CBW_RING);   // partial row statistics of the tile are written
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem + CBW_BAR_OFF + 8 * (16 + 2 * CBW_RING) + 8);
  float* const_s = reinterpret_cast<float*>(smem + CBW_CONST_OFF);
  float2* rowstat_s = reinterpret_cast<float2*>(smem + CBW_ROWSTAT_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;
  const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
  const TileSchedule sched(M);

  if (threadIdx.x == 0) {
    bar.init_all();   // acc2_empty: drained by all eight epilogue warps
    mbar_init(rowstat_bar, NUM_EPI_WARPS * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(const_cast<uint32_t*>(tmem_slot))));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < HID; i += CBW_NUM_THREADS) {
    const_s[i] = s_vec[i];
    const_s[HID + i] = b_fold[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= FIRST_PROD_WARP && warp < TMA_WARP) {
    // ============================================================ row producers: g -> TMEM
    const uint32_t dst = smem_base + CBW_XS_OFF;
    auto issue = [&](int i) {
      const int64_t m0 = sched.m0(i);
#pragma unroll 8
      for (int it = 0; it < 32; ++it) {
        const int row = quarter * 32 + it;
        const int64_t m = m0 + row;
        const bool ok = m < M;
        cp_async16(dst + (uint32_t)(row * XPITCH * 4 + lane * 16), g + (ok ? m : 0) * ld_g + 4 * lane,
                   ok ? 16u : 0u);
      }
      cp_async_commit();
    };
    const float* row = reinterpret_cast<const float*>(smem + CBW_XS_OFF) + (quarter * 32 + lane) * XPITCH;
    if (sched.count > 0) issue(0);
    for (int i = 0; i < sched.count; ++i) {
      cp_async_wait_group<0>();
      __syncwarp();
      mbar_wait(bar.x_empty(0), (i & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int part = 0; part < 4; ++part) {
        uint32_t hi[16], lo[16];
        float unused1 = 0.f, unused2 = 0.f;
        split_part<false>(row + part * 32, hi, lo, 0.f, unused1, unused2);
        tmem_st16(tmem_base + lane_base + part * 16, hi);
        tmem_st16(tmem_base + lane_base + CBW_GLO_COL + part * 16, lo);
      }
      __syncwarp();
      if (i + 1 < sched.count) issue(i + 1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar.x_full(0));
    }
  } else if (warp == TMA_WARP) {
    // ============================================================ weight-stage producer
    if (elect_one()) {
      Ring ra, rb;
      for (int i = 0; i < sched.count; ++i)
        for (int st = 0; st < cb_stages(); ++st) {
          const bool is_b = (st % 6) >= 2;
          Ring& r = is_b ? rb : ra;
          const int slot = is_b ? CBW_RING_A + r.stage : r.stage;
          mbar_wait(bar.w_empty(slot), r.phase ^ 1);
          mbar_expect_tx(bar.w_full(slot), STAGE);
          bulk_g2s(smem_base + CBW_RING_OFF + (uint32_t)slot * STAGE, image + (size_t)st * STAGE, STAGE,
                   bar.w_full(slot));
          r.advance(is_b ? CBW_RING_B : CBW_RING_A);
        }
    }
  } else if (warp == MMA_WARP) {
    // ============================================================ GEMM1 issuer: d_q = g . W_b[:, chunk]
    if (elect_one()) {
      constexpr uint32_t idesc1 = make_idesc(BM, CH);
      const uint32_t ring_u32 = smem_base + CBW_RING_OFF;
      Ring ring;
      for (int i = 0; i < sched.count; ++i) {
        mbar_wait(bar.x_full(0), i & 1);
        for (int c = 0; c < NCH; ++c) {
          const uint32_t n = (uint32_t)(i * NCH + c);
          const int b = c & 1;
          mbar_wait(bar.acc1_empty(b), ((n >> 1) & 1) ^ 1);
          tc_fence_after();
          mbar_wait(bar.w_full(ring.stage), ring.phase);
          const uint32_t st = ring_u32 + ring.stage * STAGE;
#pragma unroll
          for (int kh = 0; kh < 2; ++kh)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t a_hi = tmem_base + (kh * 4 + kk) * 8;
              mma3_ts(tmem_base + CBW_ACC1_COL + b * CH, a_hi, a_hi + CBW_GLO_COL, st + kh * 8192 + kk * 32,
                      st + kh * 8192 + 4096 + kk * 32, idesc1, (kh | kk) != 0);
            }
          tc_commit(bar.w_empty(ring.stage));
          ring.advance(CBW_RING_A);
          tc_commit(bar.acc1_full(b));
          if (c == NCH - 1) tc_commit(bar.x_empty(0));
        }
      }
    }
  } else if (warp == CBW_MMA2_WARP) {
    // ============================================================ GEMM2 issuer: z += d_p[chunk] . W'[chunk, :]
    if (elect_one()) {
      constexpr uint32_t idesc2 = make_idesc(BM, D);
      const uint32_t ring_u32 = smem_base + CBW_RING_OFF + CBW_RING_A * STAGE;
      Ring ring;
      int wt[4] = {0, 0, 0, 0};   // slots of (hi, half 0), (hi, half 1), (lo, half 0), (lo, half 1)
      for (int i = 0; i < sched.count; ++i) {
        for (int c = 0; c < NCH; ++c) {
          const int b = c & 1;
          const uint32_t u = (uint32_t)(i * NCH + c) >> 1;
          if (b == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              mbar_wait(bar.w_full(CBW_RING_A + ring.stage), ring.phase);
              wt[q] = ring.stage;
              ring.advance(CBW_RING_B);
            }
          }
          mbar_wait(bar.a2_full(b), u & 1);
          if (c == 0) mbar_wait(bar.acc2_empty(0), (i & 1) ^ 1);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint32_t koff = (uint32_t)(b * 64 + kk * 32);
              const uint32_t a_hi = tmem_base + CBW_A2_COL + b * 32 + kk * 8;
              mma3_ts(tmem_base + CBW_ACC2_COL + h * D, a_hi, a_hi + 16, ring_u32 + wt[h] * STAGE + koff,
                      ring_u32 + wt[2 + h] * STAGE + koff, idesc2, (c | kk) != 0);
            }
          tc_commit(bar.a2_empty(b));
          if (b == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q) tc_commit(bar.w_empty(CBW_RING_A + wt[q]));
          }
        }
        tc_commit(bar.acc2_full(0));
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ============================================================ silu' epilogues + row statistics
    const int half = warp >> 2;
    const float* s_s = const_s;
    const float* bf_s = const_s + HID;
    // the pre-activation block of a tile (128 KB, contiguous) is pulled into L2 one tile ahead
    auto prefetch_p = [&](int i) {
      if (warp == 0 && lane < 8)
        prefetch_l2_bulk(p + (int64_t)(sched.first + i * sched.stride) * (HID * BM) + lane * (HID * BM / 8),
                         HID * BM * 4 / 8);
    };
    if (sched.count > 0) prefetch_p(0);
    const EpiStage es{reinterpret_cast<float*>(smem + CBW_EPI_OFF) + warp * (32 * STAGE_LD), lane, lane & 3,
                      (lane >> 3) + 4 * ((lane >> 2) & 1)};
    for (int i = 0; i < sched.count; ++i) {
      const int64_t m_base = sched.m0(i) + quarter * 32;
      {   // rows of t this warp needs at the END of the tile (group 0: own rows, group 1: the rows of
          // the reversed edges): into L2 now, one row per lane
        const int64_t mp = m_base + lane;
        if (mp < M) prefetch_l2_bulk(t + (half == 0 ? mp : (int64_t)__ldg(rev + mp)) * ld_t, D * 4);
      }
      const int64_t m = sched.m0(i) + quarter * 32 + lane;
      const bool ok = m < M;
      if (i + 1 < sched.count) prefetch_p(i + 1);
      const float* ptile = p + (int64_t)(sched.first + i * sched.stride) * (HID * BM) + quarter * 32 + lane;
      float s1 = 0.f, s2 = 0.f;
      float pv[32];
      auto fetch_p = [&](int c) {   // [tile][chunk][unit][row]: one 128 B line per warp and unit
#pragma unroll
        for (int k = 0; k < 32; ++k) pv[k] = __ldg(ptile + (c * CH + k) * BM);
      };
      fetch_p(half);
      for (int c = half; c < NCH; c += 2) {
        const int b = half;
        const uint32_t u = (uint32_t)(i * NCH + c) >> 1;
        mbar_wait(bar.acc1_full(b), u & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem_base + lane_base + CBW_ACC1_COL + b * CH, v);
        tc_fence_before();
        mbar_arrive(bar.acc1_empty(b));
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float pe[4] = {pv[4 * q], pv[4 * q + 1], pv[4 * q + 2], pv[4 * q + 3]};
          float dp[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = 4 * q + e;
            const float sg = fsigmoid(pe[e]);
            dp[e] = ok ? v[k] * sg * (1.0f + pe[e] * (1.0f - sg)) : 0.f;
            s1 = fmaf(dp[e], s_s[c * CH + k], s1);
            s2 = fmaf(dp[e], pe[e] - bf_s[c * CH + k], s2);
          }
          hi[2 * q] = pack_bf16(dp[0], dp[1]);
          hi[2 * q + 1] = pack_bf16(dp[2], dp[3]);
          lo[2 * q] = pack_bf16(dp[0] - __uint_as_float(hi[2 * q] << 16), dp[1] - __uint_as_float(hi[2 * q] & 0xffff0000u));
          lo[2 * q + 1] = pack_bf16(dp[2] - __uint_as_float(hi[2 * q + 1] << 16),
                                    dp[3] - __uint_as_float(hi[2 * q + 1] & 0xffff0000u));
        }
        if (c + 2 < NCH) fetch_p(c + 2);
        mbar_wait(bar.a2_empty(b), (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t a2 = tmem_base + lane_base + CBW_A2_COL + b * 32;
        tmem_st16(a2, hi);
        tmem_st16(a2 + 16, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar.a2_full(b));
      }
      // this group's share of mean(z) and mean(z x_hat) of the row; both groups' shares are added below
      rowstat_s[half * BM + quarter * 32 + lane] = make_float2(s1 * (1.0f / HID), s2 * (1.0f / HID));
      mbar_arrive(rowstat_bar);
      // ---- output store: LayerNorm backward, d_cat = r (z - S1 - x_hat S2), x_hat = (c - mu) r,
      // c = [t_e | t_rev(e)].  Warp -> its 32 rows x the 128 columns of its group (group 0: the own
      // half, group 1: the reversed half) in 8 slices of 16; the rows of t of the next slice are in
      // flight while the current one is transposed.  Overlaps the row-producer phase of the next tile.
      {
        int64_t own[4], src[4];
        float mu[4], rs[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int64_t mm = m_base + it * 8 + es.rsel;
          const bool okr = mm < M;
          own[it] = okr ? mm : -1;
          src[it] = okr ? (half == 0 ? mm : (int64_t)__ldg(rev + mm)) : -1;
          const float2 st = okr ? __ldg(stats + mm) : make_float2(0.f, 0.f);
          mu[it] = st.x;
          rs[it] = st.y;
        }
        float4 xr[4][4];   // three slices of t in flight ahead of the one being processed
        auto fetch = [&](int s8) {
          const int col = 16 * s8 + 4 * es.c4;
#pragma unroll
          for (int it = 0; it < 4; ++it)
            xr[s8 & 3][it] = src[it] >= 0 ? ld4(t + src[it] * ld_t + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        fetch(0);
        fetch(1);
        fetch(2);
        mbar_wait(rowstat_bar, i & 1);
        float a1[4], a2[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = quarter * 32 + it * 8 + es.rsel;
          const float2 a = rowstat_s[r], b = rowstat_s[BM + r];
          a1[it] = a.x + b.x;
          a2[it] = a.y + b.y;
        }
        mbar_wait(bar.acc2_full(0), i & 1);
        tc_fence_after();
#pragma unroll
        for (int s8 = 0; s8 < 8; ++s8) {
          if (s8 + 3 < 8) fetch(s8 + 3);
          const int c0 = D * half + 16 * s8 + 4 * es.c4;
          es.fill(tmem_base + lane_base + CBW_ACC2_COL + D * half + 16 * s8);
          if (s8 == 7) {
            tc_fence_before();
            mbar_arrive(bar.acc2_empty(0));
          }
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (own[it] < 0) continue;
            const float4 z = es.get(it), x = xr[s8 & 3][it];
            const float r = rs[it], a = a1[it], k2 = a2[it] * r, mo = mu[it];
            *reinterpret_cast<float4*>(d_cat + own[it] * HID + c0) =
                make_float4(r * (z.x - a - (x.x - mo) * k2), r * (z.y - a - (x.y - mo) * k2),
                            r * (z.z - a - (x.z - mo) * k2), r * (z.w - a - (x.w - mo) * k2));
          }
        }
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

int check_dims(const char* what, int d) {
  if (d != D) {
    set_error("%s: built for d_pet = %d (LayerNorm / hidden width %d), got %d", what, D, HID, d);
    return PETB200_ERR_UNSUPPORTED;
  }
  return PETB200_OK;
}

}  // namespace
}  // namespace petb200

using namespace petb200;

// debugging aid, not part of the documented ABI: buffer of 6 roles x 4 tiles x 16 chunks x 8 int64
// slots that CTA 0 of petb200_combine_fwd fills with clock64() stamps (NULL switches it off)
extern "C" PETB200_API int petb200_debug_combine_trace(void* buffer) {
#ifdef PETB200_COMBINE_TRACE
  cudaMemcpyToSymbol(g_ctrace, &buffer, sizeof(void*));
  return check_launch("debug_combine_trace");
#else
  (void)buffer;
  set_error("debug_combine_trace: rebuild with -DPETB200_COMBINE_TRACE");
  return PETB200_ERR_UNSUPPORTED;
#endif
}

extern "C" PETB200_API size_t petb200_combine_image_bytes(int d, int backward) {
  (void)backward;
  return d == D ? (size_t)cf_stages() * STAGE : 0;
}

extern "C" PETB200_API int petb200_combine_pack(const float* w_a_folded, const float* w_b, int d, void* image_fwd,
                                                void* image_bwd, cudaStream_t stream) {
  if (int rc = check_dims("combine_pack", d)) return rc;
  for (int backward = 0; backward < 2; ++backward) {
    void* image = backward ? image_bwd : image_fwd;
    if (!image) continue;
    const int64_t chunks = (int64_t)cf_stages() * STAGE / 16;
    combine_pack_kernel<<<(unsigned)ceil_div(chunks, 256), 256, 0, stream>>>(w_a_folded, w_b, backward,
                                                                            reinterpret_cast<uint4*>(image));
  }
  return check_launch("combine_pack");
}

extern "C" PETB200_API int petb200_combine_fwd(const float* t, int64_t ld_t, const int32_t* rev, const void* image_fwd,
                                               const float* s_vec, const float* b_fold, const float* b_out,
                                               int64_t n_edges, int d, float* m_io, int64_t ld_m, float* p_out,
                                               float* stats_out, cudaStream_t stream) {
  if (int rc = check_dims("combine_fwd", d)) return rc;
  PETB200_REQUIRE(ld_t % 4 == 0 && ld_m % 4 == 0, "combine_fwd: leading dimensions must be multiples of 4");
  if (n_edges == 0) return PETB200_OK;
  const int tiles = (int)ceil_div(n_edges, BM);
  cudaFuncSetAttribute(combine_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM);
  combine_fwd_kernel<<<tiles < kNumSMs ? tiles : kNumSMs, CF_NUM_THREADS, CF_SMEM, stream>>>(
      t, ld_t, rev, reinterpret_cast<const uint8_t*>(image_fwd), s_vec, b_fold, b_out, n_edges, m_io, ld_m, p_out,
      reinterpret_cast<float2*>(stats_out));
  return check_launch("combine_fwd");
}

extern "C" PETB200_API int petb200_combine_bwd(const float* g, int64_t ld_g, const float* p, const float* t,
                                               int64_t ld_t, const int32_t* rev, const float* stats,
                                               const void* image_bwd, const float* s_vec, const float* b_fold,
                                               int64_t n_edges, int d, float* d_cat, cudaStream_t stream) {
  if (int rc = check_dims("combine_bwd", d)) return rc;
  PETB200_REQUIRE(ld_g % 4 == 0 && ld_t % 4 == 0, "combine_bwd: leading dimensions must be multiples of 4");
  if (n_edges == 0) return PETB200_OK;
  const int tiles = (int)ceil_div(n_edges, BM);
  cudaFuncSetAttribute(combine_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CBW_SMEM);
  combine_bwd_kernel<<<tiles < kNumSMs ? tiles : kNumSMs, CBW_NUM_THREADS, CBW_SMEM, stream>>>(
      g, ld_g, p, t, ld_t, rev, reinterpret_cast<const float2*>(stats), reinterpret_cast<const uint8_t*>(image_bwd),
      s_vec, b_fold, n_edges, d_cat);
  return check_launch("combine_bwd");
}
