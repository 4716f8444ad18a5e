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
// Persistent, warp-specialised CTA (one per SM), 288 threads:
//   warps 0-3  epilogue   TMEM -> registers (tcgen05.ld 32x32b) -> bias / row scale /
//                         activation / residual -> global
//   warp  4    MMA issue  one elected lane issues tcgen05.mma (M=128, N=128, K=16) and
//                         tcgen05.commit; owns the TMEM allocation (2 x 128 columns)
//   warps 5-8  producers  global fp32 A tile -> bf16 hi/lo -> 128B-swizzled K-major smem;
//                         pre-split bf16 W tile -> smem; 3-stage mbarrier ring
// Work item = (128-row tile, 128-column chunk); the accumulator is double buffered in
// TMEM so the epilogue of item i overlaps the main loop of item i+1.
//
// Weights arrive pre-split (petb200_split_bf16): row n of the "W" buffer holds K bf16 hi
// values followed by K bf16 lo values (same bytes per row as the fp32 row it replaces).
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace petb200 {
namespace {

constexpr int BM = 128, BN = 128, BK = 64;  // BK bf16 = 128 B = one swizzle row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 2;      // one bf16 operand tile: 16 KiB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;  // A_hi, A_lo, B_hi, B_lo
constexpr int NUM_EPI_WARPS = 4, NUM_PROD_WARPS = 4;
constexpr int NUM_THREADS = 32 * (NUM_EPI_WARPS + 1 + NUM_PROD_WARPS);
constexpr int TMEM_COLS = 256;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 | LBO (unused, 1) << 16 | SBO (8 rows x 128 B = 1024 B) >> 4 << 32 |
// version 1 << 46 | layout type SWIZZLE_128B (2) << 61
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, N, M
constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// byte offset of (row, 16-byte chunk) inside a [128 x 64] bf16 K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);  // .x = a in the low half
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

template <int EPI>
__device__ __forceinline__ int weight_row(int n0, int c, int F) {
  if (EPI == PETB200_EPI_SWIGLU) return (c < 64) ? (n0 / 2 + c) : (F + n0 / 2 + (c - 64));
  return n0 + c;
}

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance() {
    if (++stage == STAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
};

template <int EPI, int NPROD /*3 = bf16x3, 1 = bf16*/>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m_tiles = (int)ceil_div(g.M, BM);
  const int num_n_chunks = g.N / BN;
  const int num_items = num_m_tiles * num_n_chunks;
  const int num_k = g.K / BK;
  const int F = g.N / 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), NUM_PROD_WARPS * 32);
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp > NUM_EPI_WARPS) {
    // =============================================================== producers
    const int t = threadIdx.x - 32 * (NUM_EPI_WARPS + 1);  // 0..127
    PipeState ps;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int64_t m0 = (int64_t)(item / num_n_chunks) * BM;
      const int n0 = (item % num_n_chunks) * BN;
      for (int kc = 0; kc < num_k; ++kc) {
        mbar_wait(empty_bar(ps.stage), ps.phase ^ 1);
        uint8_t* st = smem_gen + (size_t)ps.stage * STAGE_BYTES;
        const int k0 = kc * BK;
        // ---- issue all global loads of this stage first (32 x 16 B in flight per thread)
        float4 av[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = t + 128 * i, row = idx >> 4, c4 = idx & 15;
          const int64_t m = m0 + row;
          av[i] = m < g.M ? __ldg(reinterpret_cast<const float4*>(g.A + m * g.lda + k0) + c4)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint4 bh[8], bl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = t + 128 * i, row = idx >> 3, c = idx & 7;
          // split weight row: [K bf16 hi | K bf16 lo] in the bytes of an fp32 row
          const uint8_t* wrow = reinterpret_cast<const uint8_t*>(
              g.W + (int64_t)weight_row<EPI>(n0, row, F) * g.ldw);
          bh[i] = __ldg(reinterpret_cast<const uint4*>(wrow + (size_t)k0 * 2) + c);
          if (NPROD == 3)
            bl[i] = __ldg(reinterpret_cast<const uint4*>(wrow + (size_t)g.K * 2 + (size_t)k0 * 2) + c);
        }
        // ---- A: fp32 -> bf16 hi / lo, swizzled K-major tiles
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int idx = t + 128 * i, row = idx >> 4, c4 = idx & 15;
          const float4 x = av[i];
          const float hx = bf16_round(x.x), hy = bf16_round(x.y), hz = bf16_round(x.z),
                      hw = bf16_round(x.w);
          const uint32_t off = swz(row, c4 >> 1) + ((c4 & 1) << 3);
          *reinterpret_cast<uint2*>(st + off) = make_uint2(pack_bf16(hx, hy), pack_bf16(hz, hw));
          if (NPROD == 3)
            *reinterpret_cast<uint2*>(st + TILE_BYTES + off) =
                make_uint2(pack_bf16(x.x - hx, x.y - hy), pack_bf16(x.z - hz, x.w - hw));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = t + 128 * i, row = idx >> 3, c = idx & 7;
          const uint32_t off = swz(row, c);
          *reinterpret_cast<uint4*>(st + 2 * TILE_BYTES + off) = bh[i];
          if (NPROD == 3) *reinterpret_cast<uint4*>(st + 3 * TILE_BYTES + off) = bl[i];
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core
        mbar_arrive(full_bar(ps.stage));
        ps.advance();
      }
    }
  } else if (warp == NUM_EPI_WARPS) {
    // =============================================================== MMA issuer
    constexpr uint32_t idesc = make_idesc(BM, BN);
    PipeState ps;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kc = 0; kc < num_k; ++kc) {
        mbar_wait(full_bar(ps.stage), ps.phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t st = smem_base + (uint32_t)ps.stage * STAGE_BYTES;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t a_hi = make_smem_desc(st + kk * 32);
            const uint64_t b_hi = make_smem_desc(st + 2 * TILE_BYTES + kk * 32);
            if (NPROD == 3) {
              const uint64_t a_lo = make_smem_desc(st + TILE_BYTES + kk * 32);
              const uint64_t b_lo = make_smem_desc(st + 3 * TILE_BYTES + kk * 32);
              // small terms first, then the leading term
              tc_mma(d_tmem, a_lo, b_hi, idesc, (kc | kk) != 0);
              tc_mma(d_tmem, a_hi, b_lo, idesc, 1);
              tc_mma(d_tmem, a_hi, b_hi, idesc, 1);
            } else {
              tc_mma(d_tmem, a_hi, b_hi, idesc, (kc | kk) != 0);
            }
          }
          tc_commit(empty_bar(ps.stage));  // frees the smem stage when these MMAs retire
          if (kc == num_k - 1) tc_commit(tfull_bar(acc));
        }
        __syncwarp();
        ps.advance();
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // =============================================================== epilogue
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int64_t m = (int64_t)(item / num_n_chunks) * BM + warp * 32 + lane;
      const int n0 = (item % num_n_chunks) * BN;
      const bool ok = m < g.M;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN);
      const float rs = (ok && g.row_scale) ? g.row_scale[m] : 1.0f;
      if (EPI == PETB200_EPI_SWIGLU) {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          float u[32], gt[32];
          tmem_ld32(taddr + 32 * h, u);
          tmem_ld32(taddr + 64 + 32 * h, gt);
          if (ok) {
            const int cu = n0 / 2 + 32 * h;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              u[j] = rs * u[j] + (g.bias ? __ldg(g.bias + cu + j) : 0.f);
              gt[j] = rs * gt[j] + (g.bias ? __ldg(g.bias + F + cu + j) : 0.f);
            }
            if (g.aux_out) {
              float4* pu = reinterpret_cast<float4*>(g.aux_out + m * g.ld_aux + cu);
              float4* pg = reinterpret_cast<float4*>(g.aux_out + m * g.ld_aux + F + cu);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                pu[j] = make_float4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
                pg[j] = make_float4(gt[4 * j], gt[4 * j + 1], gt[4 * j + 2], gt[4 * j + 3]);
              }
            }
            float4* po = reinterpret_cast<float4*>(g.C + m * g.ldc + cu);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              po[j] = make_float4(u[4 * j] * sigmoidf_(gt[4 * j]), u[4 * j + 1] * sigmoidf_(gt[4 * j + 1]),
                                  u[4 * j + 2] * sigmoidf_(gt[4 * j + 2]),
                                  u[4 * j + 3] * sigmoidf_(gt[4 * j + 3]));
          }
        }
      } else {
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          float v[32];
          tmem_ld32(taddr + 32 * ch, v);
          if (!ok) continue;
          const int c0 = n0 + 32 * ch;
          if (EPI == PETB200_EPI_SWIGLU_BWD) {
            const float4* pu = reinterpret_cast<const float4*>(g.aux_in + m * g.ld_aux + c0);
            const float4* pg = reinterpret_cast<const float4*>(g.aux_in + m * g.ld_aux + g.N + c0);
            float4* du = reinterpret_cast<float4*>(g.C + m * g.ldc + c0);
            float4* dg = reinterpret_cast<float4*>(g.C + m * g.ldc + g.N + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 uu = pu[j], gg = pg[j];
              const float s0 = sigmoidf_(gg.x), s1 = sigmoidf_(gg.y), s2 = sigmoidf_(gg.z),
                          s3 = sigmoidf_(gg.w);
              du[j] = make_float4(v[4 * j] * s0, v[4 * j + 1] * s1, v[4 * j + 2] * s2, v[4 * j + 3] * s3);
              dg[j] = make_float4(v[4 * j] * uu.x * s0 * (1.f - s0), v[4 * j + 1] * uu.y * s1 * (1.f - s1),
                                  v[4 * j + 2] * uu.z * s2 * (1.f - s2),
                                  v[4 * j + 3] * uu.w * s3 * (1.f - s3));
            }
            continue;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = rs * v[j] + (g.bias ? __ldg(g.bias + c0 + j) : 0.f);
          if (EPI == PETB200_EPI_SILU) {
            if (g.aux_out) {
              float4* pa = reinterpret_cast<float4*>(g.aux_out + m * g.ld_aux + c0);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                pa[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = siluf_(v[j]);
          }
          if (EPI == PETB200_EPI_MUL_DSILU) {
            const float4* pp = reinterpret_cast<const float4*>(g.aux_in + m * g.ld_aux + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 p = pp[j];
              v[4 * j] *= dsiluf_(p.x);
              v[4 * j + 1] *= dsiluf_(p.y);
              v[4 * j + 2] *= dsiluf_(p.z);
              v[4 * j + 3] *= dsiluf_(p.w);
            }
          }
          if (g.residual) {
            const float4* pr = reinterpret_cast<const float4*>(g.residual + m * g.ldr + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 r = pr[j];
              v[4 * j] += r.x;
              v[4 * j + 1] += r.y;
              v[4 * j + 2] += r.z;
              v[4 * j + 3] += r.w;
            }
          }
          float4* pc = reinterpret_cast<float4*>(g.C + m * g.ldc + c0);
          if (g.accumulate) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 o = pc[j];
              v[4 * j] += o.x;
              v[4 * j + 1] += o.y;
              v[4 * j + 2] += o.z;
              v[4 * j + 3] += o.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            pc[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
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

template <int EPI>
int launch_epi(const GemmArgs& g, int precision, cudaStream_t stream) {
  const int items = (int)ceil_div(g.M, BM) * (g.N / BN);
  const int grid = items < kNumSMs ? items : kNumSMs;
  if (precision == PETB200_PREC_BF16X3) {
    auto kern = gemm_tc_kernel<EPI, 3>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(g);
  } else {
    auto kern = gemm_tc_kernel<EPI, 1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(g);
  }
  return check_launch("gemm_tc");
}

}  // namespace

bool gemm_tc_supports(const GemmArgs& g) {
  return g.N % BN == 0 && g.K % BK == 0 && g.M < (1ll << 31) * 100;
}

int launch_gemm_tc(const GemmArgs& g, int precision, cudaStream_t stream) {
  if (g.M == 0) return PETB200_OK;
  switch (g.epilogue) {
    case PETB200_EPI_NONE: return launch_epi<PETB200_EPI_NONE>(g, precision, stream);
    case PETB200_EPI_SILU: return launch_epi<PETB200_EPI_SILU>(g, precision, stream);
    case PETB200_EPI_SWIGLU: return launch_epi<PETB200_EPI_SWIGLU>(g, precision, stream);
    case PETB200_EPI_MUL_DSILU: return launch_epi<PETB200_EPI_MUL_DSILU>(g, precision, stream);
    case PETB200_EPI_SWIGLU_BWD: return launch_epi<PETB200_EPI_SWIGLU_BWD>(g, precision, stream);
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
