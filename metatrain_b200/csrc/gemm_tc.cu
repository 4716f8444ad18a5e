// tcgen05 tensor-core GEMM for sm_100a:  C[M,N] = epi(A[M,K] * W[N,K]^T), fp32 in/out.
//
// Every dense contraction of the PET step (the torch.nn.Linear calls of
// src/metatrain/pet/modules/transformer.py and backend.py, and their dgrads) runs here.
// Activations stay fp32 in HBM; operands are split on the fly into bf16 "hi" and "lo"
// parts (x = hi + lo + O(2^-17 x)) and the product is formed with three tcgen05.mma
// (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM): the cheapest tensor-core scheme that
// keeps forces within 1e-4 eV/A of the fp32 reference (precision
// ladder).  PETB200_PREC_BF16 issues only hi*hi.
//
// Persistent, warp-specialised CTA (one per SM), 416 threads:
//   warps 0-7   epilogue   TMEM -> registers (tcgen05.ld 32x32b.x16) -> per-warp smem
//                          transpose tile -> coalesced bias / row scale / activation /
//                          residual -> global.  Warps w and w+4 share TMEM lane quarter
//                          w%4 and split the 128 accumulator columns in halves.
//   warp  8     MMA issue  one lane issues tcgen05.mma (M=128, N=128, K=16) and
//                          tcgen05.commit; owns the TMEM allocation (2 x 128 columns)
//   warps 9-12  producers  global fp32 A tile -> bf16 hi/lo -> 128B-swizzled K-major smem
//                          (mbarrier ring).  Two operand-B modes:
//                          * stationary (K <= 256): the CTA owns one 128-column chunk for
//                            its whole life, W hi/lo sits in smem once, only A streams
//                            (register double buffering keeps loads in flight);
//                          * streaming  (K  > 256): W tiles ride the ring with A.
// Work item = (128-row tile, 128-column chunk); the accumulator is double buffered in
// TMEM so the epilogue of item i overlaps the main loop of item i+1.
//
// Weights arrive pre-split (petb200_split_bf16): row n of the "W" buffer holds K bf16 hi
// values followed by K bf16 lo values (same bytes per row as the fp32 row it replaces).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace petb200 {
namespace {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 64;  // BK bf16 = 128 B = one swizzle row
constexpr int TILE_BYTES = BM * BK * 2;      // one bf16 operand tile: 16 KiB
constexpr int OPERAND_BYTES = 192 * 1024;    // A ring (+ resident or streamed W tiles)
constexpr int NUM_EPI_WARPS = 8, NUM_PROD_WARPS = 4;
constexpr int NUM_PROD_THREADS = NUM_PROD_WARPS * 32;
constexpr int NUM_THREADS = 32 * (NUM_EPI_WARPS + 1 + NUM_PROD_WARPS);
constexpr int MAX_STAGES = 4;
constexpr int TMEM_COLS = 256;
constexpr int EPI_COLS = 16;   // accumulator columns per tcgen05.ld in the epilogue
constexpr int STAGE_LD = 20;   // floats per row of the epilogue transpose tile
constexpr int EPI_STAGE_BYTES = NUM_EPI_WARPS * 32 * STAGE_LD * 4;
constexpr size_t SMEM_BYTES =
    (size_t)OPERAND_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGE_BYTES +
    2 * 2 * BM * 4 /*row-dot exchange of the RMSNorm-backward epilogue*/;

// Tile column -> weight row.  Plain layouts: identity.  SwiGLU forward: within each
// 64-column half of the tile, columns [0,32) are "value" columns and [32,64) the matching
// "gate" columns F + ... (transformer.py:40-44: v, g = w_in(x).chunk(2)), so that one
// epilogue warp sees both members of every pair.
// epilogue activations with SFU intrinsics (ex2.approx / rcp.approx: ~1e-7 relative, far
// below the bf16x3 operand error); the fp32 FFMA kernel keeps the exact expf versions
__device__ __forceinline__ float fsilu(float x) { return x * fsigmoid(x); }
__device__ __forceinline__ float fdsilu(float x) {
  const float s = fsigmoid(x);
  return s * (1.0f + x * (1.0f - s));
}

template <int EPI>
__device__ __forceinline__ int weight_row(int n0, int c, int F) {
  if (EPI == PETB200_EPI_SWIGLU) {
    const int value_col = n0 / 2 + (c >> 6) * 32 + (c & 31);
    return (c & 32) ? F + value_col : value_col;
  }
  return n0 + c;
}

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int num_stages) {
    if (++stage == num_stages) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// Work distribution.  Stationary-B: CTA b owns column chunk b % n_chunks and walks row
// tiles (b / n_chunks) + j * (grid / n_chunks).  Streaming: items b, b + grid, ...
struct Schedule {
  int num_n_chunks, first, stride, count;
  bool stationary;
  __device__ __forceinline__ Schedule(const GemmArgs& g, bool stat) {
    const int num_m_tiles = (int)ceil_div(g.M, BM);
    num_n_chunks = g.N / BN;
    stationary = stat;
    if (stat) {
      const int per = gridDim.x / num_n_chunks;  // CTAs per column chunk
      first = blockIdx.x / num_n_chunks;
      stride = per;
      count = first < num_m_tiles ? (num_m_tiles - first + per - 1) / per : 0;
    } else {
      const int items = num_m_tiles * num_n_chunks;
      first = blockIdx.x;
      stride = gridDim.x;
      count = first < items ? (items - first + stride - 1) / stride : 0;
    }
  }
  __device__ __forceinline__ void item(int j, int64_t& m0, int& n0) const {
    if (stationary) {
      m0 = (int64_t)(first + j * stride) * BM;
      n0 = (blockIdx.x % num_n_chunks) * BN;
    } else {
      const int it = first + j * stride;
      m0 = (int64_t)(it / num_n_chunks) * BM;
      n0 = (it % num_n_chunks) * BN;
    }
  }
};

template <int EPI, int NPROD /*3 = bf16x3, 1 = bf16*/, bool BSTAT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + OPERAND_BYTES;
  // barriers: loaded / full / empty [MAX_STAGES], tmem_full[2], tmem_empty[2], TMEM base
  auto loaded_bar = [&](int s) { return bar_base + 8u * s; };
  auto full_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * MAX_STAGES + 2 + a); };
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem_gen + OPERAND_BYTES + 8 * (3 * MAX_STAGES + 4));
  float* stage_all = reinterpret_cast<float*>(smem_gen + OPERAND_BYTES + 256);
  float* dots_all = stage_all + EPI_STAGE_BYTES / 4;  // [item parity][half][BM]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = g.K / BK;
  const int F = g.N / 2;
  // operand memory map.  Stationary: [W: num_k x (hi, lo)] then the A ring of (hi, lo)
  // pairs; streaming: ring of (A_hi, A_lo, W_hi, W_lo).
  const int b_bytes = BSTAT ? num_k * 2 * TILE_BYTES : 0;
  const int stage_bytes = BSTAT ? 2 * TILE_BYTES : 4 * TILE_BYTES;
  int num_stages = (OPERAND_BYTES - b_bytes) / stage_bytes;
  if (num_stages > MAX_STAGES) num_stages = MAX_STAGES;
  const Schedule sched(g, BSTAT);

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      // one expect_tx arrival for the TMA boxes of A (+ the cp.async arrivals of streamed W)
      mbar_init(loaded_bar(s), BSTAT ? 1 : NUM_PROD_THREADS + 1);
      mbar_init(full_bar(s), NUM_PROD_THREADS);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), NUM_EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  if (warp == NUM_EPI_WARPS) {  // the MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_slot))),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (BSTAT && warp > NUM_EPI_WARPS) {
    // resident W: every producer thread copies its share of the (hi, lo) tiles once
    const int t = threadIdx.x - 32 * (NUM_EPI_WARPS + 1);
    const int n0 = (blockIdx.x % sched.num_n_chunks) * BN;
    for (int kc = 0; kc < num_k; ++kc) {
#pragma unroll
      for (int i = 0; i < 1024 / NUM_PROD_THREADS; ++i) {
        const int idx = t + NUM_PROD_THREADS * i, row = idx >> 3, c = idx & 7;
        const uint8_t* wrow =
            reinterpret_cast<const uint8_t*>(g.W + (int64_t)weight_row<EPI>(n0, row, F) * g.ldw);
        const uint4 hi = __ldg(reinterpret_cast<const uint4*>(wrow + (size_t)kc * BK * 2) + c);
        *reinterpret_cast<uint4*>(smem_gen + (size_t)kc * 2 * TILE_BYTES + swz(row, c)) = hi;
        if (NPROD == 3) {
          const uint4 lo = __ldg(
              reinterpret_cast<const uint4*>(wrow + (size_t)g.K * 2 + (size_t)kc * BK * 2) + c);
          *reinterpret_cast<uint4*>(smem_gen + (size_t)kc * 2 * TILE_BYTES + TILE_BYTES + swz(row, c)) = lo;
        }
      }
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp > NUM_EPI_WARPS) {
    // =============================================================== producers
    // Stage layout: [A_hi tile | A_lo tile | (streaming) W_hi tile | W_lo tile].  The fp32
    // A chunk (128 rows x 64 floats) is copied asynchronously (cp.async, no registers) so
    // that row r's floats 0..31 land in the 128 bytes of A_hi row r and floats 32..63 in
    // A_lo row r; a warp then converts the row IN PLACE to its bf16 hi / lo 128-byte
    // swizzled rows.  Copies run `num_stages - 1` chunks ahead of the conversion.
    const int t = threadIdx.x - 32 * (NUM_EPI_WARPS + 1);  // 0..NUM_PROD_THREADS-1
    const int pw = t >> 5;
    const uint32_t ring_u32 = smem_base + (uint32_t)b_bytes;
    uint8_t* ring_base = smem_gen + b_bytes;
    const int total = sched.count * num_k;  // (item, k-chunk) pairs of this CTA, flattened
    const int depth = num_stages - 1;
    Ring load_ring, conv_ring;
    for (int q = 0; q < total + depth; ++q) {
      if (q < total) {
        int64_t m0;
        int n0;
        sched.item(q / num_k, m0, n0);
        const int k0 = (q % num_k) * BK;
        mbar_wait(empty_bar(load_ring.stage), load_ring.phase ^ 1);
        const uint32_t st = ring_u32 + (uint32_t)load_ring.stage * stage_bytes;
        // the fp32 A chunk arrives as two TMA boxes of [128 rows x 32 floats] (rows beyond M as
        // zeros): floats 0..31 of a row into the bytes of its hi-tile row, 32..63 into the lo-tile row
        if (t == 0) {
          mbar_expect_tx(loaded_bar(load_ring.stage), 2 * TILE_BYTES);
          tma_load_2d(st, &map_a, k0, (int)m0, loaded_bar(load_ring.stage));
          tma_load_2d(st + TILE_BYTES, &map_a, k0 + 32, (int)m0, loaded_bar(load_ring.stage));
        }
        if (!BSTAT) {
#pragma unroll
          for (int i = 0; i < 1024 / NUM_PROD_THREADS; ++i) {
            const int idx = t + NUM_PROD_THREADS * i, row = idx >> 3, c = idx & 7;
            const uint8_t* wrow =
                reinterpret_cast<const uint8_t*>(g.W + (int64_t)weight_row<EPI>(n0, row, F) * g.ldw);
            cp_async16(st + 2 * TILE_BYTES + swz(row, c), wrow + (size_t)k0 * 2 + c * 16, 16u);
            if (NPROD == 3)
              cp_async16(st + 3 * TILE_BYTES + swz(row, c),
                         wrow + (size_t)g.K * 2 + (size_t)k0 * 2 + c * 16, 16u);
          }
        }
        if (!BSTAT) cp_async_arrive(loaded_bar(load_ring.stage));
        load_ring.advance(num_stages);
      }
      if (q >= depth) {
        mbar_wait(loaded_bar(conv_ring.stage), conv_ring.phase);
        uint8_t* st = ring_base + (size_t)conv_ring.stage * stage_bytes;
        // warp pw converts its ROWS_PER_WARP rows, two rows per instruction: lane l works on
        // row (l >> 4) of the pair and owns k = 4*quad .. 4*quad+3 (quad = l & 15).  One
        // cvt.rn.bf16x2 (F2FP) rounds and packs two values; hi is rebuilt with shifts.
        constexpr int ROWS_PER_WARP = BM / NUM_PROD_WARPS;
        const int quad = lane & 15, rsub = lane >> 4;
        const uint8_t* src_tile = st + (quad < 8 ? 0 : TILE_BYTES) + (quad & 7) * 16;
        const int chunk = quad >> 1, within = (quad & 1) * 8;
#pragma unroll
        for (int b = 0; b < ROWS_PER_WARP / 16; ++b) {
          float4 x[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int r = pw * ROWS_PER_WARP + b * 16 + jj * 2 + rsub;
            x[jj] = *reinterpret_cast<const float4*>(src_tile + r * 128);
          }
          __syncwarp();  // whole rows are in registers before they are overwritten
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int r = pw * ROWS_PER_WARP + b * 16 + jj * 2 + rsub;
            const uint32_t h01 = pack_bf16(x[jj].x, x[jj].y), h23 = pack_bf16(x[jj].z, x[jj].w);
            const uint32_t off = (uint32_t)(r * 128 + ((chunk ^ (r & 7)) << 4) + within);
            *reinterpret_cast<uint2*>(st + off) = make_uint2(h01, h23);
            if (NPROD == 3) {
              const float l0 = x[jj].x - __uint_as_float(h01 << 16);
              const float l1 = x[jj].y - __uint_as_float(h01 & 0xffff0000u);
              const float l2 = x[jj].z - __uint_as_float(h23 << 16);
              const float l3 = x[jj].w - __uint_as_float(h23 & 0xffff0000u);
              *reinterpret_cast<uint2*>(st + TILE_BYTES + off) =
                  make_uint2(pack_bf16(l0, l1), pack_bf16(l2, l3));
            }
          }
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core
        mbar_arrive(full_bar(conv_ring.stage));
        conv_ring.advance(num_stages);
      }
    }
  } else if (warp == NUM_EPI_WARPS) {
    // =============================================================== MMA issuer
    constexpr uint32_t idesc = make_idesc(BM, BN);
    Ring ring;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int j = 0; j < sched.count; ++j) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kc = 0; kc < num_k; ++kc) {
        mbar_wait(full_bar(ring.stage), ring.phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_base + (uint32_t)b_bytes + (uint32_t)ring.stage * stage_bytes;
          const uint32_t bt = BSTAT ? smem_base + (uint32_t)kc * 2 * TILE_BYTES : st + 2 * TILE_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t a_hi = make_smem_desc(st + kk * 32);
            const uint64_t b_hi = make_smem_desc(bt + kk * 32);
            if (NPROD == 3) {
              const uint64_t a_lo = make_smem_desc(st + TILE_BYTES + kk * 32);
              const uint64_t b_lo = make_smem_desc(bt + TILE_BYTES + kk * 32);
              // small terms first, then the leading term
              tc_mma(d_tmem, a_lo, b_hi, idesc, (kc | kk) != 0);
              tc_mma(d_tmem, a_hi, b_lo, idesc, 1);
              tc_mma(d_tmem, a_hi, b_hi, idesc, 1);
            } else {
              tc_mma(d_tmem, a_hi, b_hi, idesc, (kc | kk) != 0);
            }
          }
          tc_commit(empty_bar(ring.stage));  // frees the smem stage when these MMAs retire
          if (kc == num_k - 1) tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        ring.advance(num_stages);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // =============================================================== epilogue
    // TMEM gives each thread one accumulator row (16 columns per tcgen05.ld).  Rows are
    // transposed through a per-warp shared-memory tile (stride 20 floats: conflict-free
    // float4 both ways) so that every global access of the epilogue (C, residual, aux) is a
    // coalesced 64-byte row segment: lane -> (row = 8*it + rsel, float4 column lane%4).
    float* stage = stage_all + warp * (32 * STAGE_LD);
    const int quarter = warp & 3, half = warp >> 2;
    const int c4 = lane & 3, rsel = (lane >> 3) + 4 * ((lane >> 2) & 1);
    int acc = 0;
    uint32_t acc_phase = 0;
    auto ld4 = [&](const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int NCH = 64 / EPI_COLS;  // column chunks per warp
    // one additive operand is prefetched per item: the residual, or (C += ...) the old C;
    // the rare "residual AND accumulate" case adds the old C late, unprefetched
    const float* pre_src = g.residual ? g.residual : (g.accumulate ? g.C : nullptr);
    const int64_t pre_ld = g.residual ? g.ldr : g.ldc;
    // the residual applies to rows [0, residual_rows) only (edge rows of an [edges | atoms] token matrix)
    const int64_t res_rows = (g.residual && g.residual_rows >= 0) ? g.residual_rows : g.M;
    const bool late_acc = g.residual && g.accumulate;
    // SILU_GEO (petb200_compress_gemm, N = BN): the 4 geometry weights of each of this lane's 16 output
    // columns stay in registers for the CTA's life (they were 4 L1 loads per output float4 before)
    float4 gw[EPI == PETB200_EPI_SILU_GEO ? NCH : 1][4];
    if (EPI == PETB200_EPI_SILU_GEO) {
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
        for (int q = 0; q < 4; ++q) gw[ch][q] = ld4(g.geo_w + (64 * half + EPI_COLS * ch + 4 * c4 + q) * 4);
    }
    for (int j = 0; j < sched.count; ++j) {
      int64_t m0;
      int n0;
      sched.item(j, m0, n0);
      const int64_t m_base = m0 + quarter * 32;
      float rs[4];
      bool ok[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int64_t m = m_base + it * 8 + rsel;
        ok[it] = m < g.M;
        rs[it] = (g.row_scale && ok[it]) ? __ldg(g.row_scale + m) : 1.0f;
      }
      float4 geo[4];
      int zrow[4];
      if (EPI == PETB200_EPI_SILU_GEO) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int64_t m = m_base + it * 8 + rsel;
          geo[it] = zero4;
          zrow[it] = 0;
          if (!ok[it]) continue;
          geo[it] = make_float4(__ldg(g.geo_vec + 3 * m), __ldg(g.geo_vec + 3 * m + 1),
                                __ldg(g.geo_vec + 3 * m + 2), __ldg(g.geo_dist + m));
          if (g.row_table) zrow[it] = __ldg(g.row_index + m);
        }
      }
      // Global operands of the whole item are requested BEFORE waiting for the MMA, so
      // their latency hides behind the main loop of this item.
      float4 pre[EPI == PETB200_EPI_SILU_GEO ? 1 : NCH][4];   // (SILU_GEO has no additive operand)
      if (EPI == PETB200_EPI_NONE || EPI == PETB200_EPI_SILU || EPI == PETB200_EPI_MUL_DSILU ||
          EPI == PETB200_EPI_RMS_BWD) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int64_t m = m_base + it * 8 + rsel;
            const int c0 = n0 + 64 * half + EPI_COLS * ch + 4 * c4;
            pre[ch][it] = zero4;
            if (!ok[it]) continue;
            if (EPI == PETB200_EPI_MUL_DSILU || EPI == PETB200_EPI_RMS_BWD) {
              pre[ch][it] = ld4(g.aux_in + m * g.ld_aux + c0);
            } else if (pre_src && m < res_rows) {
              pre[ch][it] = ld4(pre_src + m * pre_ld + c0);
            }
          }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      auto stage_chunk = [&](int tile_col) {
        float v[EPI_COLS];
        tmem_ld16(taddr + tile_col, v);
        __syncwarp();  // readers of the previous chunk are done with the tile
#pragma unroll
        for (int q = 0; q < EPI_COLS / 4; ++q)
          *reinterpret_cast<float4*>(stage + lane * STAGE_LD + 4 * q) =
              make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
      };
      auto staged = [&](int it) {
        return *reinterpret_cast<const float4*>(stage + (it * 8 + rsel) * STAGE_LD + 4 * c4);
      };

      if (EPI == PETB200_EPI_SWIGLU) {
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          const int tile_u = 64 * half + EPI_COLS * ch;       // value columns of this warp
          const int cu = n0 / 2 + 32 * half + EPI_COLS * ch + 4 * c4;
          const float4 bu = g.bias ? ld4(g.bias + cu) : zero4;
          const float4 bg = g.bias ? ld4(g.bias + F + cu) : zero4;
          float4 u[4];
          stage_chunk(tile_u);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const float4 x = staged(it);
            u[it] = make_float4(rs[it] * x.x + bu.x, rs[it] * x.y + bu.y, rs[it] * x.z + bu.z,
                                rs[it] * x.w + bu.w);
          }
          stage_chunk(tile_u + 32);  // the matching gate columns
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (!ok[it]) continue;
            const int64_t m = m_base + it * 8 + rsel;
            const float4 x = staged(it);
            const float4 gt = make_float4(rs[it] * x.x + bg.x, rs[it] * x.y + bg.y,
                                          rs[it] * x.z + bg.z, rs[it] * x.w + bg.w);
            if (g.aux_out) {
              *reinterpret_cast<float4*>(g.aux_out + m * g.ld_aux + cu) = u[it];
              *reinterpret_cast<float4*>(g.aux_out + m * g.ld_aux + F + cu) = gt;
            }
            *reinterpret_cast<float4*>(g.C + m * g.ldc + cu) =
                make_float4(u[it].x * fsigmoid(gt.x), u[it].y * fsigmoid(gt.y),
                            u[it].z * fsigmoid(gt.z), u[it].w * fsigmoid(gt.w));
          }
        }
      } else if (EPI == PETB200_EPI_RMS_BWD) {
        // C = residual + rs * d - x * rs^3 * (d . x) / N with d = the accumulator row (gradient
        // w.r.t. the normalised row), x = aux_in, rs = row_scale; N = BN (one column chunk).
        // Pass 1: row dots (this warp's 64 columns, then exchanged with the partner warp of the
        // same lane quarter); pass 2: re-read the accumulator and form the output.
        float dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          stage_chunk(64 * half + EPI_COLS * ch);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const float4 v = staged(it), x4 = pre[ch][it];
            dot[it] += v.x * x4.x + v.y * x4.y + v.z * x4.z + v.w * x4.w;
          }
        }
        float* dots = dots_all + (j & 1) * 2 * BM;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          dot[it] += __shfl_xor_sync(0xffffffffu, dot[it], 1);
          dot[it] += __shfl_xor_sync(0xffffffffu, dot[it], 2);
          if (c4 == 0) dots[half * BM + quarter * 32 + it * 8 + rsel] = dot[it];
        }
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        float kap[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = quarter * 32 + it * 8 + rsel;
          kap[it] = rs[it] * rs[it] * rs[it] * (dots[r] + dots[BM + r]) * (1.0f / BN);
        }
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int tile_col = 64 * half + EPI_COLS * ch;
          const int c0 = n0 + tile_col + 4 * c4;
          stage_chunk(tile_col);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (!ok[it]) continue;
            const int64_t m = m_base + it * 8 + rsel;
            const float4 v = staged(it), x4 = pre[ch][it];
            float4 o = make_float4(rs[it] * v.x - x4.x * kap[it], rs[it] * v.y - x4.y * kap[it],
                                   rs[it] * v.z - x4.z * kap[it], rs[it] * v.w - x4.w * kap[it]);
            if (g.residual && m < res_rows) {
              const float4 r4 = ld4(g.residual + m * g.ldr + c0);
              o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
            }
            *reinterpret_cast<float4*>(g.C + m * g.ldc + c0) = o;
          }
        }
      } else if (EPI == PETB200_EPI_SWIGLU_BWD) {
        // two operands per output: fetch them two column chunks at a time
#pragma unroll 1
        for (int pair = 0; pair < NCH / 2; ++pair) {
          float4 uu[2][4], gg[2][4];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int64_t m = m_base + it * 8 + rsel;
              const int c0 = n0 + 64 * half + EPI_COLS * (2 * pair + h2) + 4 * c4;
              uu[h2][it] = ok[it] ? ld4(g.aux_in + m * g.ld_aux + c0) : zero4;
              gg[h2][it] = ok[it] ? ld4(g.aux_in + m * g.ld_aux + g.N + c0) : zero4;
            }
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int tile_col = 64 * half + EPI_COLS * (2 * pair + h2);
            const int c0 = n0 + tile_col + 4 * c4;
            stage_chunk(tile_col);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              if (!ok[it]) continue;
              const int64_t m = m_base + it * 8 + rsel;
              const float4 v = staged(it);
              const float4 u4 = uu[h2][it], g4 = gg[h2][it];
              const float s0 = fsigmoid(g4.x), s1 = fsigmoid(g4.y), s2 = fsigmoid(g4.z),
                          s3 = fsigmoid(g4.w);
              *reinterpret_cast<float4*>(g.C + m * g.ldc + c0) =
                  make_float4(v.x * s0, v.y * s1, v.z * s2, v.w * s3);
              *reinterpret_cast<float4*>(g.C + m * g.ldc + g.N + c0) =
                  make_float4(v.x * u4.x * s0 * (1.f - s0), v.y * u4.y * s1 * (1.f - s1),
                              v.z * u4.z * s2 * (1.f - s2), v.w * u4.w * s3 * (1.f - s3));
            }
          }
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const int tile_col = 64 * half + EPI_COLS * ch;
          const int c0 = n0 + tile_col + 4 * c4;  // this lane's 4 output columns
          const float4 b4 = g.bias ? ld4(g.bias + c0) : zero4;
          stage_chunk(tile_col);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (!ok[it]) continue;
            const int64_t m = m_base + it * 8 + rsel;
            float4 v = staged(it);
            v = make_float4(rs[it] * v.x + b4.x, rs[it] * v.y + b4.y, rs[it] * v.z + b4.z,
                            rs[it] * v.w + b4.w);
            if (EPI == PETB200_EPI_SILU_GEO) {
              // + G[c] . (r, d) of this row + table row of its neighbour species
              const float4 gv = geo[it];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 w = gw[EPI == PETB200_EPI_SILU_GEO ? ch : 0][q];
                (&v.x)[q] += w.x * gv.x + w.y * gv.y + w.z * gv.z + w.w * gv.w;
              }
              if (g.row_table) {
                const float4 tb = ld4(g.row_table + (int64_t)zrow[it] * g.N + c0);
                v.x += tb.x; v.y += tb.y; v.z += tb.z; v.w += tb.w;
              }
            }
            if (EPI == PETB200_EPI_SILU || EPI == PETB200_EPI_SILU_GEO) {
              if (g.aux_out) *reinterpret_cast<float4*>(g.aux_out + m * g.ld_aux + c0) = v;
              v = make_float4(fsilu(v.x), fsilu(v.y), fsilu(v.z), fsilu(v.w));
            }
            if (EPI == PETB200_EPI_MUL_DSILU) {
              v.x *= fdsilu(pre[ch][it].x);
              v.y *= fdsilu(pre[ch][it].y);
              v.z *= fdsilu(pre[ch][it].z);
              v.w *= fdsilu(pre[ch][it].w);
              if (pre_src) {  // rare with this epilogue: not prefetched
                const float4 r = ld4(pre_src + m * pre_ld + c0);
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
              }
            } else if (EPI != PETB200_EPI_SILU_GEO) {
              v.x += pre[ch][it].x;
              v.y += pre[ch][it].y;
              v.z += pre[ch][it].z;
              v.w += pre[ch][it].w;
            }
            if (late_acc) {
              const float4 o = *reinterpret_cast<const float4*>(g.C + m * g.ldc + c0);
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *reinterpret_cast<float4*>(g.C + m * g.ldc + c0) = v;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NUM_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS));
  }
}

__global__ void split_bf16_kernel(const float* __restrict__ w, int64_t rows, int cols,
                                  float* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  int64_t r = idx / cols;
  int c = (int)(idx % cols);
  float x = w[idx];
  __nv_bfloat16 hi = __float2bfloat16_rn(x);
  __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(out + r * cols);
  row[c] = hi;
  row[cols + c] = lo;
}

template <int EPI, int NPROD, bool BSTAT>
int launch_one(const GemmArgs& g, int grid, cudaStream_t stream) {
  auto kern = gemm_tc_kernel<EPI, NPROD, BSTAT>;
  CUtensorMap map_a;
  if (make_tma_map_f32(&map_a, g.A, g.M, g.K, g.lda, 32, BM)) {
    set_error("gemm: cuTensorMapEncodeTiled failed for A (pointer / leading dimension must be 16-byte aligned)");
    return PETB200_ERR_CUDA;
  }
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(map_a, g);
  return check_launch("gemm_tc");
}

template <int EPI>
int launch_epi(const GemmArgs& g, int precision, cudaStream_t stream) {
  const int m_tiles = (int)ceil_div(g.M, BM), n_chunks = g.N / BN;
  // stationary W needs K/64 x 32 KiB of smem and at least two A stages in 192 KiB
  const bool stationary = g.K <= 256 && n_chunks <= kNumSMs;
  int grid;
  if (stationary) {
    int per = kNumSMs / n_chunks;
    if (per > m_tiles) per = m_tiles;
    grid = per * n_chunks;
  } else {
    const int items = m_tiles * n_chunks;
    grid = items < kNumSMs ? items : kNumSMs;
  }
  const bool x3 = precision == PETB200_PREC_BF16X3;
  if (stationary) return x3 ? launch_one<EPI, 3, true>(g, grid, stream) : launch_one<EPI, 1, true>(g, grid, stream);
  return x3 ? launch_one<EPI, 3, false>(g, grid, stream) : launch_one<EPI, 1, false>(g, grid, stream);
}

}  // namespace

bool gemm_tc_supports(const GemmArgs& g) {
  return g.N % BN == 0 && g.K % BK == 0 && g.M < (1ll << 31);
}

int launch_gemm_tc(const GemmArgs& g, int precision, cudaStream_t stream) {
  if (g.M == 0) return PETB200_OK;
  switch (g.epilogue) {
    case PETB200_EPI_NONE: return launch_epi<PETB200_EPI_NONE>(g, precision, stream);
    case PETB200_EPI_SILU: return launch_epi<PETB200_EPI_SILU>(g, precision, stream);
    case PETB200_EPI_SWIGLU: return launch_epi<PETB200_EPI_SWIGLU>(g, precision, stream);
    case PETB200_EPI_MUL_DSILU: return launch_epi<PETB200_EPI_MUL_DSILU>(g, precision, stream);
    case PETB200_EPI_SWIGLU_BWD: return launch_epi<PETB200_EPI_SWIGLU_BWD>(g, precision, stream);
    case PETB200_EPI_SILU_GEO: return launch_epi<PETB200_EPI_SILU_GEO>(g, precision, stream);
    case PETB200_EPI_RMS_BWD:
      if (g.N != BN || !g.aux_in || !g.row_scale || g.accumulate) {
        set_error("gemm: the RMSNorm-backward epilogue needs N = %d, aux_in (x) and row_scale (rstd)", BN);
        return PETB200_ERR_INVALID_ARGUMENT;
      }
      return launch_epi<PETB200_EPI_RMS_BWD>(g, precision, stream);
    default:
      set_error("gemm: unknown epilogue %d", g.epilogue);
      return PETB200_ERR_INVALID_ARGUMENT;
  }
}

}  // namespace petb200

extern "C" PETB200_API int petb200_split_bf16(const float* w, int64_t rows, int cols, float* out,
                                              cudaStream_t stream) {
  using namespace petb200;
  PETB200_REQUIRE(cols % 8 == 0, "split_bf16: cols must be a multiple of 8");
  if (rows * cols == 0) return PETB200_OK;
  split_bf16_kernel<<<(unsigned)ceil_div(rows * cols, 256), 256, 0, stream>>>(w, rows, cols, out);
  return check_launch("split_bf16");
}
