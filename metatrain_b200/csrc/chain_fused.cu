// Two chained 128 -> 128 Linears with a SiLU between them, forward and backward, as single persistent
// tcgen05 kernels (sm_100a), for the two places PET has this shape on its edge rows:
//
//   HEADS    edge head + last layer + its gradient seed (backend.py:171-217, 651-777):
//              fwd:  e2p = W_2 silu(W_1 m + b_1) + b_2 ;  pe = w_e . silu(e2p) + b_e
//              bwd:  d_m = W_1^T [ silu'(e1p) (W_2^T g) ],  g = (d_atomic[ctr] f_c) w_e silu'(e2p),
//                    d_fc += d_atomic[ctr] pe            (g is formed on the fly by the row producers)
//   COMPRESS token builder of a CartesianTransformer (transformer.py:500-521, concatenation folded:
//            c_1 = W_1m m + G (r, d) + Tbl[z_j] + b', see petb200_compress_gemm):
//              fwd:  t = W_2 silu(c_1) + b_2
//              bwd:  d_c1 = silu'(c_1) (W_2^T d_t) ;  d_m += W_1m^T d_c1 ;  d_(r, d) += G^T d_c1
//
// The unfused schedules moved the [E, 128] pre-activations and activations of both Linears through
// HBM (heads: 8.1 KB per edge forward + backward incl. the readout kernels; compress: 5.5 KB).  Here
// the hidden activations never leave the SM; the forward keeps only the pre-activation of the first
// Linear (kernels' private tile layout, see combine_fused.cu) and — for the heads — that of the second.
//
// Structure as in mlp_fused.cu: A operands in tensor memory (lane = row), weights streamed from L2 as
// 16 KB stages of a pre-swizzled image, hidden dimension walked in 4 chunks of 32 units, double-buffered
// output accumulator, 19 warps: 0-7 activation epilogues (two groups on alternate chunks), 8 GEMM1
// issue, 9-12 row producers, 13 weight stages, 14-17 output store, 18 GEMM2 issue.  All products use
// the bf16 hi/lo 2-term split (fp32 accumulation in TMEM).
#include "fused_common.cuh"

namespace petb200 {
namespace {

using namespace tc;
using namespace fused;

constexpr int MODE_HEADS = 0, MODE_COMPRESS = 1;
constexpr int HID = 128;            // hidden width (= d_pet = d_head)
constexpr int NCH = HID / CH;       // 4 chunks of 32 hidden units
// 15 warps: 0-7 epilogue groups, 8 GEMM1 issue, 9-12 row producers (which also run the output store of the
// previous tile once their tile is converted: no dedicated store warps, 480 threads, 128 registers), 13
// weight stages, 14 GEMM2 issue
constexpr int CN_NUM_THREADS = 32 * 15;
constexpr int CN_MMA2_WARP = 14;
constexpr int XPITCH = D + 4;
constexpr int CN_STAGING_BYTES = BM * XPITCH * 4;
constexpr int CN_RING_A = 3, CN_RING_B = 4;
constexpr int CN_BOX_BYTES = 32 * 128;   // [32 rows x 32 floats] output box, SWIZZLE_128B (TMA tile store)
constexpr int CN_RING = CN_RING_A + CN_RING_B;
constexpr int CN_XS_OFF = 0;
constexpr int CN_RING_OFF = ((CN_XS_OFF + CN_STAGING_BYTES + 1023) / 1024) * 1024;
constexpr int CN_EPI_OFF = CN_RING_OFF + CN_RING * STAGE;           // 4 store warps x 2 output boxes
constexpr int CN_CONST_OFF = CN_EPI_OFF + 4 * 2 * CN_BOX_BYTES;     // b1 [128], b2 [128], w_e [128], G [128 x 4]
constexpr int CN_CONST_FLOATS = 3 * HID + 4 * HID;
constexpr int CN_PART_OFF = CN_CONST_OFF + CN_CONST_FLOATS * 4;     // geometry partials [2 groups][128] float4
constexpr int CN_BAR_OFF = CN_PART_OFF + 2 * BM * 16;
constexpr int CN_SMEM = CN_BAR_OFF + 8 * (16 + 2 * CN_RING) + 8 + 16 + 1024;
static_assert(CN_SMEM <= 232448, "chain kernels: shared memory budget");
// TMEM columns: A hi 0..63, lo 64..127 ; acc1[b] at 128 + 32 b ; A2[b] at 192 + 32 b (hi 16 | lo 16) ;
// acc2[t] at 256 + 128 t
constexpr int CN_ALO_COL = 64, CN_ACC1_COL = 128, CN_A2_COL = 192, CN_ACC2_COL = 256;

__host__ __device__ constexpr int cn_stages() { return (HID / 64) * 4; }   // per direction

// ----------------------------------------------------------------------- weight images
// forward, per pair of chunks p: W1(2p) W1(2p+1) W2hi(p) W2lo(p)
//   W1(c): w1[32 c .. 32 c + 31, :] per k-half [32 rows x 64 k] hi 4 KB | lo 4 KB ; W2(p): w2[:, 64 p ..] [128 x 64]
// backward, per pair of chunks p: G1(2p) G1(2p+1) WT(p, hi) WT(p, lo)
//   G1(c): w2^T[32 c .. 32 c + 31, :] per k-half [32 rows (hidden) x 64 k (out dim)] ; WT(p): w1^T[:, 64 p ..] [128 x 64]
__global__ void chain_pack_kernel(const float* __restrict__ w1, const float* __restrict__ w2, int backward,
                                  uint4* __restrict__ image) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)cn_stages() * (STAGE / 16)) return;
  const int s = (int)(idx / (STAGE / 16));
  const int o = (int)(idx % (STAGE / 16)) * 16;
  const int p = s / 4, r = s % 4;
  float v[8];
  bool lo;
  if (r < 2) {
    const int c = 2 * p + r, kh = o >> 13, t = o & 8191;
    lo = t >= 4096;
    const int u = t & 4095, n = (u >> 10) * 8 + ((u >> 7) & 7), j = ((u >> 4) & 7) ^ (n & 7);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      v[e] = backward ? w2[(int64_t)(kh * 64 + j * 8 + e) * HID + c * CH + n]
                      : w1[(int64_t)(c * CH + n) * HID + kh * 64 + j * 8 + e];
  } else {
    lo = r == 3;
    const int n = (o >> 10) * 8 + ((o >> 7) & 7), j = ((o >> 4) & 7) ^ (n & 7);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      v[e] = backward ? w1[(int64_t)(p * 64 + j * 8 + e) * HID + n] : w2[(int64_t)n * HID + p * 64 + j * 8 + e];
  }
  image[idx] = pack8(v, lo);
}

__device__ __forceinline__ void split16(const float (&x)[32], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    hi[q] = pack_bf16(x[2 * q], x[2 * q + 1]);
    lo[q] = pack_bf16(x[2 * q] - __uint_as_float(hi[q] << 16), x[2 * q + 1] - __uint_as_float(hi[q] & 0xffff0000u));
  }
}
__device__ __forceinline__ float dsilu_fast(float x) {
  const float sg = fsigmoid(x);
  return sg * (1.0f + x * (1.0f - sg));
}

struct ChainArgs {
  const float* a;            // fwd: m [E, ld_a] ; bwd HEADS: e2p [E, 128] ; bwd COMPRESS: d_t [E, ld_a]
  int64_t ld_a;
  const uint8_t* image;
  const float* b1;           // fwd: bias of the first Linear (COMPRESS: b' = b_1 + W_1geo b_geo)
  const float* b2;           // fwd: bias of the second Linear
  const float* w_e;          // HEADS: last-layer weight [128]
  float b_e;                 // HEADS: last-layer bias
  const float* geo_w;        // COMPRESS: G [128, 4]
  const float* table;        // COMPRESS: Tbl [S, 128] or null
  const int32_t* z;          // COMPRESS: species index per edge
  const float* vec;          // COMPRESS: edge vectors [E, 3]
  const float* dist;         // COMPRESS: edge distances [E]
  float* p1;                 // pre-activation of the first Linear, private tile layout (fwd: out, bwd: in)
  float* out;                // fwd HEADS: e2p [E, 128] ; fwd COMPRESS: t [E, ld_out] ; bwd: d_m [E, ld_out]
  int64_t ld_out;
  float* pe;                 // HEADS fwd: edge predictions [E] (out) ; bwd: in
  const float* d_atomic;     // HEADS bwd: [N]
  const int32_t* ctr;        // HEADS bwd: centre atom of each edge
  const float* fc;           // HEADS bwd: cutoff factors
  float* d_fc;               // HEADS bwd: += d_atomic[ctr] pe
  float* d_vec;              // COMPRESS bwd: += G^T d_c1 (first 3 components) [E, 3]
  float* d_dist;             // COMPRESS bwd: += (4th component) [E]
  int accumulate;            // bwd: d_m += result instead of d_m = result
  int64_t M;
};

// ======================================================================== forward
template <int MODE>
__global__ void __launch_bounds__(CN_NUM_THREADS, 1) chain_fwd_kernel(const __grid_constant__ CUtensorMap map_out, const ChainArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const Barriers bar{smem_base + CN_BAR_OFF, CN_RING};
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + CN_BAR_OFF + 8 * (16 + 2 * CN_RING) + 8);
  float* const_s = reinterpret_cast<float*>(smem + CN_CONST_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;
  const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
  const int64_t M = g.M;
  const TileSchedule sched(M);

  if (threadIdx.x == 0) {
    bar.init_all();
    mbar_init(bar.acc2_empty(0), 4 * 32);   // drained by the four store warps
    mbar_init(bar.acc2_empty(1), 4 * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(const_cast<uint32_t*>(tmem_slot))));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < HID; i += CN_NUM_THREADS) {
    const_s[i] = g.b1[i];
    const_s[HID + i] = g.b2[i];
    const_s[2 * HID + i] = MODE == MODE_HEADS ? g.w_e[i] : 0.f;
  }
  if (MODE == MODE_COMPRESS)
    for (int i = threadIdx.x; i < 4 * HID; i += CN_NUM_THREADS) const_s[3 * HID + i] = g.geo_w[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

    // ============================================================ output store (run by the row-producer warps)
    // out = acc2 + b_2 ; HEADS: also pe = w_e . silu(out) + b_e (a store warp holds all 128 columns of
    // its 32 rows: the row dot is complete inside the thread).
    // A thread owns one row (its TMEM lane): per 32-column slice it adds the bias, writes the row's 128 bytes
    // into a SWIZZLE_128B box in shared memory and one lane hands the box to the TMA (tile store; rows beyond
    // M are clipped by the tensor map) — 1 tcgen05.ld + 8 STS.128 + 1 bulk store per slice and warp.
    const int sw = (warp - FIRST_PROD_WARP) & 3;
    const uint32_t box_u32 = smem_base + CN_EPI_OFF + (uint32_t)sw * 2 * CN_BOX_BYTES;
    uint8_t* box = smem + CN_EPI_OFF + sw * 2 * CN_BOX_BYTES;
    const float* b2_s = const_s + HID;
    const float* we_s = const_s + 2 * HID;
    uint32_t nbox = 0;
    auto store_tile = [&](int i) {
      const int64_t m = sched.m0(i) + quarter * 32 + lane;
      const int m_row = (int)(sched.m0(i) + quarter * 32);
      const int t = i & 1;
      mbar_wait(bar.acc2_full(t), (i >> 1) & 1);
      tc_fence_after();
      float dot = 0.f;
#pragma unroll
      for (int sl = 0; sl < 4; ++sl, ++nbox) {
        float v[32];
        tmem_ld32(tmem_base + lane_base + CN_ACC2_COL + t * D + 32 * sl, v);
        if (sl == 3) {   // the accumulator is in registers now: release it early
          tc_fence_before();
          mbar_arrive(bar.acc2_empty(t));
        }
        if (lane == 0) bulk_wait_group_read<1>();   // the store that used this box two slices ago has read it
        __syncwarp();
        uint8_t* dst = box + (nbox & 1) * CN_BOX_BYTES + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = *reinterpret_cast<const float4*>(b2_s + 32 * sl + 4 * j);
          const float4 o = make_float4(v[4 * j] + b4.x, v[4 * j + 1] + b4.y, v[4 * j + 2] + b4.z, v[4 * j + 3] + b4.w);
          *reinterpret_cast<float4*>(dst + ((j ^ (lane & 7)) << 4)) = o;
          if (MODE == MODE_HEADS) {
            const float4 w4 = *reinterpret_cast<const float4*>(we_s + 32 * sl + 4 * j);
            dot += w4.x * o.x * fsigmoid(o.x) + w4.y * o.y * fsigmoid(o.y) + w4.z * o.z * fsigmoid(o.z) +
                   w4.w * o.w * fsigmoid(o.w);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&map_out, 32 * sl, m_row, box_u32 + (nbox & 1) * CN_BOX_BYTES);
          bulk_commit_group();
        }
      }
      if (MODE == MODE_HEADS && m < M) g.pe[m] = dot + g.b_e;
    };
    auto store_drain = [&]() {
      if (lane == 0) bulk_wait_group<0>();
    };

  if (warp >= FIRST_PROD_WARP && warp < TMA_WARP) {
    // ============================================================ row producers: m -> TMEM (raw rows)
    const uint32_t dst = smem_base + CN_XS_OFF;
    auto issue = [&](int i) {
      const int64_t m0 = sched.m0(i);
#pragma unroll 8
      for (int it = 0; it < 32; ++it) {
        const int row = quarter * 32 + it;
        const int64_t m = m0 + row;
        const bool ok = m < M;
        cp_async16(dst + (uint32_t)(row * XPITCH * 4 + lane * 16), g.a + (ok ? m : 0) * g.ld_a + 4 * lane,
                   ok ? 16u : 0u);
      }
      cp_async_commit();
    };
    const float* row = reinterpret_cast<const float*>(smem + CN_XS_OFF) + (quarter * 32 + lane) * XPITCH;
    if (sched.count > 0) issue(0);
    for (int i = 0; i < sched.count; ++i) {
      cp_async_wait_group<0>();
      __syncwarp();
      mbar_wait(bar.x_empty(0), (i & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int part = 0; part < 4; ++part) {
        float x[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(row + part * 32 + 4 * q);
          x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
        uint32_t hi[16], lo[16];
        split16(x, hi, lo);
        tmem_st16(tmem_base + lane_base + part * 16, hi);
        tmem_st16(tmem_base + lane_base + CN_ALO_COL + part * 16, lo);
      }
      __syncwarp();
      if (i + 1 < sched.count) issue(i + 1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar.x_full(0));
      if (i > 0) store_tile(i - 1);   // overlaps the chunk loop of tile i
    }
    if (sched.count > 0) store_tile(sched.count - 1);
    store_drain();
  } else if (warp == TMA_WARP) {
    // ============================================================ weight-stage producer
    if (elect_one()) {
      Ring ra, rb;
      for (int i = 0; i < sched.count; ++i)
        for (int st = 0; st < cn_stages(); ++st) {
          const bool is_b = (st % 4) >= 2;
          Ring& r = is_b ? rb : ra;
          const int slot = is_b ? CN_RING_A + r.stage : r.stage;
          mbar_wait(bar.w_empty(slot), r.phase ^ 1);
          mbar_expect_tx(bar.w_full(slot), STAGE);
          bulk_g2s(smem_base + CN_RING_OFF + (uint32_t)slot * STAGE, g.image + (size_t)st * STAGE, STAGE,
                   bar.w_full(slot));
          r.advance(is_b ? CN_RING_B : CN_RING_A);
        }
    }
  } else if (warp == MMA_WARP) {
    // ============================================================ GEMM1 issuer: acc1[b] = A . W1[chunk]^T
    if (elect_one()) {
      constexpr uint32_t idesc1 = make_idesc(BM, CH);
      const uint32_t ring_u32 = smem_base + CN_RING_OFF;
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
              mma3_ts(tmem_base + CN_ACC1_COL + b * CH, a_hi, a_hi + CN_ALO_COL, st + kh * 8192 + kk * 32,
                      st + kh * 8192 + 4096 + kk * 32, idesc1, (kh | kk) != 0);
            }
          tc_commit(bar.w_empty(ring.stage));
          ring.advance(CN_RING_A);
          tc_commit(bar.acc1_full(b));
          if (c == NCH - 1) tc_commit(bar.x_empty(0));
        }
      }
    }
  } else if (warp == CN_MMA2_WARP) {
    // ============================================================ GEMM2 issuer: acc2[t] += a[chunk] . W2[:, chunk]^T
    if (elect_one()) {
      constexpr uint32_t idesc2 = make_idesc(BM, D);
      const uint32_t ring_u32 = smem_base + CN_RING_OFF + CN_RING_A * STAGE;
      Ring ring;
      int w2_hi = 0, w2_lo = 0;
      for (int i = 0; i < sched.count; ++i) {
        for (int c = 0; c < NCH; ++c) {
          const int b = c & 1;
          const uint32_t u = (uint32_t)(i * NCH + c) >> 1;
          if (b == 0) {
            mbar_wait(bar.w_full(CN_RING_A + ring.stage), ring.phase);
            w2_hi = ring.stage;
            ring.advance(CN_RING_B);
            mbar_wait(bar.w_full(CN_RING_A + ring.stage), ring.phase);
            w2_lo = ring.stage;
            ring.advance(CN_RING_B);
          }
          mbar_wait(bar.a2_full(b), u & 1);
          if (c == 0) mbar_wait(bar.acc2_empty(i & 1), ((i >> 1) & 1) ^ 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t koff = (uint32_t)(b * 64 + kk * 32);
            const uint32_t a_hi = tmem_base + CN_A2_COL + b * 32 + kk * 8;
            mma3_ts(tmem_base + CN_ACC2_COL + (i & 1) * D, a_hi, a_hi + 16, ring_u32 + w2_hi * STAGE + koff,
                    ring_u32 + w2_lo * STAGE + koff, idesc2, (c | kk) != 0);
          }
          tc_commit(bar.a2_empty(b));
          if (b == 1) {
            tc_commit(bar.w_empty(CN_RING_A + w2_hi));
            tc_commit(bar.w_empty(CN_RING_A + w2_lo));
          }
        }
        tc_commit(bar.acc2_full(i & 1));
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ============================================================ first-Linear epilogues: bias (+ geometry) + SiLU
    const int half = warp >> 2;
    const float* b1_s = const_s;
    const float4* geo_s = reinterpret_cast<const float4*>(const_s + 3 * HID);
    for (int i = 0; i < sched.count; ++i) {
      const int64_t m = sched.m0(i) + quarter * 32 + lane;
      const bool ok = m < M;
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* trow = nullptr;
      if (MODE == MODE_COMPRESS && ok) {
        gv = make_float4(__ldg(g.vec + 3 * m), __ldg(g.vec + 3 * m + 1), __ldg(g.vec + 3 * m + 2), __ldg(g.dist + m));
        if (g.table) trow = reinterpret_cast<const float4*>(g.table + (int64_t)__ldg(g.z + m) * HID);
      }
      for (int c = half; c < NCH; c += 2) {
        const int b = half;
        const uint32_t u = (uint32_t)(i * NCH + c) >> 1;
        mbar_wait(bar.acc1_full(b), u & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem_base + lane_base + CN_ACC1_COL + b * CH, v);
        tc_fence_before();
        mbar_arrive(bar.acc1_empty(b));
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] += b1_s[c * CH + k];
        if (MODE == MODE_COMPRESS) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float4 w = geo_s[c * CH + k];
            v[k] += w.x * gv.x + w.y * gv.y + w.z * gv.z + w.w * gv.w;
          }
          if (trow != nullptr) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 tb = __ldg(trow + c * 8 + q);
              v[4 * q] += tb.x; v[4 * q + 1] += tb.y; v[4 * q + 2] += tb.z; v[4 * q + 3] += tb.w;
            }
          }
        }
        {   // saved for the backward: private layout [tile][chunk][unit][row]
          float* dst = g.p1 + ((int64_t)(sched.first + i * sched.stride) * NCH + c) * (CH * BM) + quarter * 32 + lane;
#pragma unroll
          for (int k = 0; k < 32; ++k) dst[k * BM] = v[k];
        }
        float a[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) a[k] = v[k] * fsigmoid(v[k]);
        uint32_t hi[16], lo[16];
        split16(a, hi, lo);
        mbar_wait(bar.a2_empty(b), (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t a2 = tmem_base + lane_base + CN_A2_COL + b * 32;
        tmem_st16(a2, hi);
        tmem_st16(a2 + 16, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar.a2_full(b));
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
template <int MODE>
__global__ void __launch_bounds__(CN_NUM_THREADS, 1) chain_bwd_kernel(const __grid_constant__ CUtensorMap map_out, const ChainArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const Barriers bar{smem_base + CN_BAR_OFF, CN_RING};
  const uint32_t part_bar = bar.at(16 + 2 * CN_RING);   // COMPRESS: geometry partials of the tile are written
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + CN_BAR_OFF + 8 * (16 + 2 * CN_RING) + 8);
  float* const_s = reinterpret_cast<float*>(smem + CN_CONST_OFF);
  float4* part_s = reinterpret_cast<float4*>(smem + CN_PART_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;
  const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
  const int64_t M = g.M;
  const TileSchedule sched(M);
  // COMPRESS, first GNN layer: the gradient w.r.t. the input messages is not needed (they are an
  // embedding of the neighbour species): no second contraction, no output store
  const bool need_dm = g.out != nullptr;

  if (threadIdx.x == 0) {
    bar.init_all();
    mbar_init(bar.acc2_empty(0), 4 * 32);
    mbar_init(bar.acc2_empty(1), 4 * 32);
    mbar_init(part_bar, NUM_EPI_WARPS * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(const_cast<uint32_t*>(tmem_slot))));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (MODE == MODE_HEADS)
    for (int i = threadIdx.x; i < HID; i += CN_NUM_THREADS) const_s[2 * HID + i] = g.w_e[i];
  if (MODE == MODE_COMPRESS)
    for (int i = threadIdx.x; i < 4 * HID; i += CN_NUM_THREADS) const_s[3 * HID + i] = g.geo_w[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

    // ============================================================ output store: d_m (+)= acc2
    // as in the forward kernel: rows go through SWIZZLE_128B boxes and the TMA; with `accumulate` the tile
    // operation is a reduction (d_m += box, added at the L2 — one contribution per element and launch), so
    // the old values are never loaded by the SM
    const int sw = (warp - FIRST_PROD_WARP) & 3;
    const uint32_t box_u32 = smem_base + CN_EPI_OFF + (uint32_t)sw * 2 * CN_BOX_BYTES;
    uint8_t* box = smem + CN_EPI_OFF + sw * 2 * CN_BOX_BYTES;
    uint32_t nbox = 0;
    auto store_tile = [&](int i) {
      if (!need_dm) return;
      const int m_row = (int)(sched.m0(i) + quarter * 32);
      const int t = i & 1;
      mbar_wait(bar.acc2_full(t), (i >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int sl = 0; sl < 4; ++sl, ++nbox) {
        float v[32];
        tmem_ld32(tmem_base + lane_base + CN_ACC2_COL + t * D + 32 * sl, v);
        if (sl == 3) {
          tc_fence_before();
          mbar_arrive(bar.acc2_empty(t));
        }
        if (lane == 0) bulk_wait_group_read<1>();
        __syncwarp();
        uint8_t* dst = box + (nbox & 1) * CN_BOX_BYTES + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(dst + ((j ^ (lane & 7)) << 4)) =
              make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (g.accumulate) tma_reduce_add_2d(&map_out, 32 * sl, m_row, box_u32 + (nbox & 1) * CN_BOX_BYTES);
          else tma_store_2d(&map_out, 32 * sl, m_row, box_u32 + (nbox & 1) * CN_BOX_BYTES);
          bulk_commit_group();
        }
      }
    };
    auto store_drain = [&]() {
      if (lane == 0) bulk_wait_group<0>();
    };

  if (warp >= FIRST_PROD_WARP && warp < TMA_WARP) {
    // ============================================================ row producers
    // COMPRESS: d_t rows -> TMEM.  HEADS: the gradient seed of the second Linear is formed here,
    // g = (d_atomic[ctr] f_c) w_e silu'(e2p), from the staged e2p rows; d_fc += d_atomic[ctr] pe.
    const uint32_t dst = smem_base + CN_XS_OFF;
    auto issue = [&](int i) {
      const int64_t m0 = sched.m0(i);
#pragma unroll 8
      for (int it = 0; it < 32; ++it) {
        const int row = quarter * 32 + it;
        const int64_t m = m0 + row;
        const bool ok = m < M;
        cp_async16(dst + (uint32_t)(row * XPITCH * 4 + lane * 16), g.a + (ok ? m : 0) * g.ld_a + 4 * lane,
                   ok ? 16u : 0u);
      }
      cp_async_commit();
    };
    const float* row = reinterpret_cast<const float*>(smem + CN_XS_OFF) + (quarter * 32 + lane) * XPITCH;
    const float* we_s = const_s + 2 * HID;
    if (sched.count > 0) issue(0);
    for (int i = 0; i < sched.count; ++i) {
      float coef = 0.f;
      if (MODE == MODE_HEADS) {
        const int64_t m = sched.m0(i) + quarter * 32 + lane;
        if (m < M) {
          const float da = __ldg(g.d_atomic + __ldg(g.ctr + m));
          coef = da * __ldg(g.fc + m);
          if (g.d_fc) g.d_fc[m] += da * __ldg(g.pe + m);
        }
      }
      cp_async_wait_group<0>();
      __syncwarp();
      mbar_wait(bar.x_empty(0), (i & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int part = 0; part < 4; ++part) {
        float x[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(row + part * 32 + 4 * q);
          x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
        if (MODE == MODE_HEADS) {
#pragma unroll
          for (int k = 0; k < 32; ++k) x[k] = coef * we_s[part * 32 + k] * dsilu_fast(x[k]);
        }
        uint32_t hi[16], lo[16];
        split16(x, hi, lo);
        tmem_st16(tmem_base + lane_base + part * 16, hi);
        tmem_st16(tmem_base + lane_base + CN_ALO_COL + part * 16, lo);
      }
      __syncwarp();
      if (i + 1 < sched.count) issue(i + 1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar.x_full(0));
      if (i > 0) store_tile(i - 1);   // overlaps the chunk loop of tile i
    }
    if (sched.count > 0) store_tile(sched.count - 1);
    store_drain();
  } else if (warp == TMA_WARP) {
    if (elect_one()) {
      Ring ra, rb;
      for (int i = 0; i < sched.count; ++i)
        for (int st = 0; st < cn_stages(); ++st) {
          const bool is_b = (st % 4) >= 2;
          if (is_b && !need_dm) continue;
          Ring& r = is_b ? rb : ra;
          const int slot = is_b ? CN_RING_A + r.stage : r.stage;
          mbar_wait(bar.w_empty(slot), r.phase ^ 1);
          mbar_expect_tx(bar.w_full(slot), STAGE);
          bulk_g2s(smem_base + CN_RING_OFF + (uint32_t)slot * STAGE, g.image + (size_t)st * STAGE, STAGE,
                   bar.w_full(slot));
          r.advance(is_b ? CN_RING_B : CN_RING_A);
        }
    }
  } else if (warp == MMA_WARP) {
    // ============================================================ GEMM1 issuer: d_a[chunk] = g . W2[:, chunk]
    if (elect_one()) {
      constexpr uint32_t idesc1 = make_idesc(BM, CH);
      const uint32_t ring_u32 = smem_base + CN_RING_OFF;
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
              mma3_ts(tmem_base + CN_ACC1_COL + b * CH, a_hi, a_hi + CN_ALO_COL, st + kh * 8192 + kk * 32,
                      st + kh * 8192 + 4096 + kk * 32, idesc1, (kh | kk) != 0);
            }
          tc_commit(bar.w_empty(ring.stage));
          ring.advance(CN_RING_A);
          tc_commit(bar.acc1_full(b));
          if (c == NCH - 1) tc_commit(bar.x_empty(0));
        }
      }
    }
  } else if (warp == CN_MMA2_WARP) {
    // ============================================================ GEMM2 issuer: d_m += d_p[chunk] . W1[chunk, :]
    if (need_dm && elect_one()) {
      constexpr uint32_t idesc2 = make_idesc(BM, D);
      const uint32_t ring_u32 = smem_base + CN_RING_OFF + CN_RING_A * STAGE;
      Ring ring;
      int w_hi = 0, w_lo = 0;
      for (int i = 0; i < sched.count; ++i) {
        for (int c = 0; c < NCH; ++c) {
          const int b = c & 1;
          const uint32_t u = (uint32_t)(i * NCH + c) >> 1;
          if (b == 0) {
            mbar_wait(bar.w_full(CN_RING_A + ring.stage), ring.phase);
            w_hi = ring.stage;
            ring.advance(CN_RING_B);
            mbar_wait(bar.w_full(CN_RING_A + ring.stage), ring.phase);
            w_lo = ring.stage;
            ring.advance(CN_RING_B);
          }
          mbar_wait(bar.a2_full(b), u & 1);
          if (c == 0) mbar_wait(bar.acc2_empty(i & 1), ((i >> 1) & 1) ^ 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t koff = (uint32_t)(b * 64 + kk * 32);
            const uint32_t a_hi = tmem_base + CN_A2_COL + b * 32 + kk * 8;
            mma3_ts(tmem_base + CN_ACC2_COL + (i & 1) * D, a_hi, a_hi + 16, ring_u32 + w_hi * STAGE + koff,
                    ring_u32 + w_lo * STAGE + koff, idesc2, (c | kk) != 0);
          }
          tc_commit(bar.a2_empty(b));
          if (b == 1) {
            tc_commit(bar.w_empty(CN_RING_A + w_hi));
            tc_commit(bar.w_empty(CN_RING_A + w_lo));
          }
        }
        tc_commit(bar.acc2_full(i & 1));
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ============================================================ silu' epilogues (+ geometry gradient)
    const int half = warp >> 2;
    const float4* geo_s = reinterpret_cast<const float4*>(const_s + 3 * HID);
    for (int i = 0; i < sched.count; ++i) {
      const int64_t m = sched.m0(i) + quarter * 32 + lane;
      const bool ok = m < M;
      const float* ptile = g.p1 + (int64_t)(sched.first + i * sched.stride) * (HID * BM) + quarter * 32 + lane;
      float pv[32];
      auto fetch_p = [&](int c) {
#pragma unroll
        for (int k = 0; k < 32; ++k) pv[k] = __ldg(ptile + (c * CH + k) * BM);
      };
      fetch_p(half);
      float4 geo = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = half; c < NCH; c += 2) {
        const int b = half;
        const uint32_t u = (uint32_t)(i * NCH + c) >> 1;
        mbar_wait(bar.acc1_full(b), u & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem_base + lane_base + CN_ACC1_COL + b * CH, v);
        tc_fence_before();
        mbar_arrive(bar.acc1_empty(b));
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          v[k] = ok ? v[k] * dsilu_fast(pv[k]) : 0.f;
          if (MODE == MODE_COMPRESS) {
            const float4 w = geo_s[c * CH + k];
            geo.x = fmaf(v[k], w.x, geo.x); geo.y = fmaf(v[k], w.y, geo.y);
            geo.z = fmaf(v[k], w.z, geo.z); geo.w = fmaf(v[k], w.w, geo.w);
          }
        }
        if (c + 2 < NCH) fetch_p(c + 2);
        if (need_dm) {
          uint32_t hi[16], lo[16];
          split16(v, hi, lo);
          mbar_wait(bar.a2_empty(b), (u & 1) ^ 1);
          tc_fence_after();
          const uint32_t a2 = tmem_base + lane_base + CN_A2_COL + b * 32;
          tmem_st16(a2, hi);
          tmem_st16(a2 + 16, lo);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(bar.a2_full(b));
        }
      }
      if (MODE == MODE_COMPRESS) {
        // d_(r, d) += G^T d_c1: this group's share; group 0 adds both after the barrier (fixed order:
        // deterministic)
        part_s[half * BM + quarter * 32 + lane] = geo;
        mbar_arrive(part_bar);
        if (half == 0) {
          mbar_wait(part_bar, i & 1);
          const float4 o = part_s[BM + quarter * 32 + lane];
          if (ok) {
            g.d_vec[3 * m] += geo.x + o.x;
            g.d_vec[3 * m + 1] += geo.y + o.y;
            g.d_vec[3 * m + 2] += geo.z + o.z;
            g.d_dist[m] += geo.w + o.w;
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

template <int MODE>
int launch_fwd(const ChainArgs& g, cudaStream_t stream) {
  const int tiles = (int)ceil_div(g.M, BM);
  CUtensorMap map_out;
  if (make_tma_map_f32(&map_out, g.out, g.M, D, g.ld_out, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B) != 0) {
    set_error("chain_fwd: cuTensorMapEncodeTiled failed (output must be 16-byte aligned, ld a multiple of 4)");
    return PETB200_ERR_CUDA;
  }
  cudaFuncSetAttribute(chain_fwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, CN_SMEM);
  chain_fwd_kernel<MODE><<<tiles < kNumSMs ? tiles : kNumSMs, CN_NUM_THREADS, CN_SMEM, stream>>>(map_out, g);
  return check_launch("chain_fwd");
}
template <int MODE>
int launch_bwd(const ChainArgs& g, cudaStream_t stream) {
  const int tiles = (int)ceil_div(g.M, BM);
  CUtensorMap map_out;
  if (g.out != nullptr) {
    if (make_tma_map_f32(&map_out, g.out, g.M, D, g.ld_out, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B) != 0) {
      set_error("chain_bwd: cuTensorMapEncodeTiled failed (output must be 16-byte aligned, ld a multiple of 4)");
      return PETB200_ERR_CUDA;
    }
  } else {
    memset(&map_out, 0, sizeof(map_out));   // no d_m requested: the store warps exit without touching it
  }
  cudaFuncSetAttribute(chain_bwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, CN_SMEM);
  chain_bwd_kernel<MODE><<<tiles < kNumSMs ? tiles : kNumSMs, CN_NUM_THREADS, CN_SMEM, stream>>>(map_out, g);
  return check_launch("chain_bwd");
}
}  // namespace
}  // namespace petb200

using namespace petb200;

extern "C" PETB200_API size_t petb200_chain_image_bytes(int d) { return d == D ? (size_t)cn_stages() * STAGE : 0; }

extern "C" PETB200_API int petb200_chain_pack(const float* w1, const float* w2, int d, void* image_fwd, void* image_bwd,
                                              cudaStream_t stream) {
  PETB200_REQUIRE(d == D, "chain_pack: built for 128 -> 128 -> 128 (got d = %d)", d);
  for (int backward = 0; backward < 2; ++backward) {
    void* image = backward ? image_bwd : image_fwd;
    if (!image) continue;
    const int64_t chunks = (int64_t)cn_stages() * STAGE / 16;
    chain_pack_kernel<<<(unsigned)ceil_div(chunks, 256), 256, 0, stream>>>(w1, w2, backward,
                                                                          reinterpret_cast<uint4*>(image));
  }
  return check_launch("chain_pack");
}

extern "C" PETB200_API int petb200_edge_head_fwd(const float* m, int64_t ld_m, const void* image_fwd, const float* b1,
                                                 const float* b2, const float* w_e, float b_e, int64_t n_edges, int d,
                                                 float* e1p, float* e2p, float* edge_pred, cudaStream_t stream) {
  PETB200_REQUIRE(d == D && ld_m % 4 == 0, "edge_head_fwd: built for d = 128, ld_m %% 4 == 0");
  if (n_edges == 0) return PETB200_OK;
  ChainArgs g{};
  g.a = m; g.ld_a = ld_m; g.image = reinterpret_cast<const uint8_t*>(image_fwd); g.b1 = b1; g.b2 = b2; g.w_e = w_e;
  g.b_e = b_e; g.p1 = e1p; g.out = e2p; g.ld_out = D; g.pe = edge_pred; g.M = n_edges;
  return launch_fwd<MODE_HEADS>(g, stream);
}

extern "C" PETB200_API int petb200_edge_head_bwd(const float* d_atomic, const int32_t* ctr, const float* fc,
                                                 const float* e1p, const float* e2p, const float* edge_pred,
                                                 const void* image_bwd, const float* w_e, int64_t n_edges, int d,
                                                 float* d_m, int64_t ld_dm, float* d_fc, cudaStream_t stream) {
  PETB200_REQUIRE(d == D && ld_dm % 4 == 0, "edge_head_bwd: built for d = 128, ld_dm %% 4 == 0");
  if (n_edges == 0) return PETB200_OK;
  ChainArgs g{};
  g.a = e2p; g.ld_a = D; g.image = reinterpret_cast<const uint8_t*>(image_bwd); g.w_e = w_e;
  g.p1 = const_cast<float*>(e1p); g.out = d_m; g.ld_out = ld_dm; g.pe = const_cast<float*>(edge_pred);
  g.d_atomic = d_atomic; g.ctr = ctr; g.fc = fc; g.d_fc = d_fc; g.accumulate = 0; g.M = n_edges;
  return launch_bwd<MODE_HEADS>(g, stream);
}

extern "C" PETB200_API int petb200_compress_fwd(const float* messages, int64_t ld_m, const void* image_fwd,
                                                const float* b_fold, const float* geo_w, const float* nbr_table,
                                                const int32_t* z_neighbor, const float* edge_vec,
                                                const float* edge_dist, const float* b2, int64_t n_edges, int d,
                                                float* c1, float* t_out, int64_t ld_t, cudaStream_t stream) {
  PETB200_REQUIRE(d == D && ld_m % 4 == 0 && ld_t % 4 == 0, "compress_fwd: built for d = 128, leading dimensions %% 4 == 0");
  if (n_edges == 0) return PETB200_OK;
  PETB200_REQUIRE(!nbr_table || z_neighbor, "compress_fwd: a table needs its row index");
  ChainArgs g{};
  g.a = messages; g.ld_a = ld_m; g.image = reinterpret_cast<const uint8_t*>(image_fwd); g.b1 = b_fold; g.b2 = b2;
  g.geo_w = geo_w; g.table = nbr_table; g.z = z_neighbor; g.vec = edge_vec; g.dist = edge_dist; g.p1 = c1;
  g.out = t_out; g.ld_out = ld_t; g.M = n_edges;
  return launch_fwd<MODE_COMPRESS>(g, stream);
}

extern "C" PETB200_API int petb200_compress_bwd(const float* d_t, int64_t ld_dt, const float* c1, const void* image_bwd,
                                                const float* geo_w, int64_t n_edges, int d, float* d_m, int64_t ld_dm,
                                                int accumulate, float* d_vec, float* d_dist, cudaStream_t stream) {
  PETB200_REQUIRE(d == D && ld_dt % 4 == 0 && (d_m == nullptr || ld_dm % 4 == 0),
                  "compress_bwd: built for d = 128, leading dimensions %% 4 == 0");
  if (n_edges == 0) return PETB200_OK;
  ChainArgs g{};
  g.a = d_t; g.ld_a = ld_dt; g.image = reinterpret_cast<const uint8_t*>(image_bwd); g.geo_w = geo_w;
  g.p1 = const_cast<float*>(c1); g.out = d_m; g.ld_out = ld_dm; g.accumulate = accumulate; g.d_vec = d_vec;
  g.d_dist = d_dist; g.M = n_edges;
  return launch_bwd<MODE_COMPRESS>(g, stream);
}
