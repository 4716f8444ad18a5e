// Pieces shared by the fused tcgen05 kernels (mlp_fused.cu, combine_fused.cu): tile schedule,
// weight-stage rings, the mbarrier table, the 2-term-split MMA triples and the TMEM -> smem epilogue
// transpose.  sm_100a only.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace petb200 {
namespace fused {

using namespace tc;

constexpr int D = 128;          // d_pet
constexpr int MAX_F = 512;      // largest hidden width (after SwiGLU) the bias buffer holds
constexpr int BM = 128;
constexpr int CH = 32;          // hidden units per chunk
constexpr int TILE = 16384;     // [128 x 64] bf16 operand tile, K-major SWIZZLE_128B
constexpr int STAGE = 16384;    // weight ring stage
constexpr int NUM_EPI_WARPS = 8, NUM_PROD_WARPS = 4;
constexpr int NUM_PROD_THREADS = NUM_PROD_WARPS * 32;
constexpr int MMA_WARP = NUM_EPI_WARPS, FIRST_PROD_WARP = NUM_EPI_WARPS + 1;
constexpr int TMA_WARP = FIRST_PROD_WARP + NUM_PROD_WARPS;
constexpr int NUM_THREADS = 32 * (TMA_WARP + 1);
constexpr int STAGE_LD = 20;    // floats per row of the epilogue transpose tile
constexpr int EPI_STAGE_BYTES = NUM_EPI_WARPS * 32 * STAGE_LD * 4;
constexpr float kRmsEps = 1.1920928955078125e-07f;  // torch.nn.RMSNorm default: finfo(fp32).eps

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++stage == n) {
      stage = 0;
      phase ^= 1;
    }
  }
};

__device__ __forceinline__ uint4 pack8(const float* v, bool lo) {
  uint32_t w[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float a = v[2 * q], b = v[2 * q + 1];
    if (lo) {
      a -= __bfloat162float(__float2bfloat16_rn(a));
      b -= __bfloat162float(__float2bfloat16_rn(b));
    }
    w[q] = pack_bf16(a, b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

struct TileSchedule {
  int first, stride, count;
  __device__ __forceinline__ TileSchedule(int64_t M) {
    const int tiles = (int)ceil_div(M, BM);
    first = blockIdx.x;
    stride = gridDim.x;
    count = first < tiles ? (tiles - first + stride - 1) / stride : 0;
  }
  __device__ __forceinline__ int64_t m0(int i) const { return (int64_t)(first + i * stride) * BM; }
};

// three MMAs of the 2-term split, A and B in shared memory (K-major SWIZZLE_128B tiles)
__device__ __forceinline__ void mma3_ss(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                        uint32_t b_lo, uint32_t idesc, bool accumulate) {
  const uint64_t ah = make_smem_desc(a_hi), al = make_smem_desc(a_lo);
  const uint64_t bh = make_smem_desc(b_hi), bl = make_smem_desc(b_lo);
  tc_mma(d_tmem, al, bh, idesc, accumulate);
  tc_mma(d_tmem, ah, bl, idesc, 1);
  tc_mma(d_tmem, ah, bh, idesc, 1);
}
// the same with A in tensor memory
__device__ __forceinline__ void mma3_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                        uint32_t b_lo, uint32_t idesc, bool accumulate) {
  const uint64_t bh = make_smem_desc(b_hi), bl = make_smem_desc(b_lo);
  tc_mma_ts(d_tmem, a_lo, bh, idesc, accumulate);
  tc_mma_ts(d_tmem, a_hi, bl, idesc, 1);
  tc_mma_ts(d_tmem, a_hi, bh, idesc, 1);
}


struct Barriers {
  uint32_t base;
  int ring;
  __device__ __forceinline__ uint32_t at(int i) const { return base + 8u * i; }
  __device__ __forceinline__ uint32_t x_full(int b) const { return at(b); }
  __device__ __forceinline__ uint32_t x_empty(int b) const { return at(2 + b); }
  __device__ __forceinline__ uint32_t acc1_full(int b) const { return at(4 + b); }
  __device__ __forceinline__ uint32_t acc1_empty(int b) const { return at(6 + b); }
  __device__ __forceinline__ uint32_t a2_full(int b) const { return at(8 + b); }
  __device__ __forceinline__ uint32_t a2_empty(int b) const { return at(10 + b); }
  __device__ __forceinline__ uint32_t acc2_full(int b) const { return at(12 + b); }
  __device__ __forceinline__ uint32_t acc2_empty(int b) const { return at(14 + b); }
  __device__ __forceinline__ uint32_t w_full(int s) const { return at(16 + s); }
  __device__ __forceinline__ uint32_t w_empty(int s) const { return at(16 + ring + s); }
  __device__ __forceinline__ void init_all() const {
    for (int b = 0; b < 2; ++b) {
      mbar_init(x_full(b), NUM_PROD_THREADS);
      mbar_init(x_empty(b), 1);
      mbar_init(acc1_full(b), 1);
      mbar_init(acc1_empty(b), NUM_EPI_WARPS * 16);  // one epilogue group (4 warps)
      mbar_init(a2_full(b), NUM_EPI_WARPS * 16);
      mbar_init(a2_empty(b), 1);
      mbar_init(acc2_full(b), 1);
      mbar_init(acc2_empty(b), NUM_EPI_WARPS * 32);
    }
    for (int s = 0; s < ring; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
};

// TMEM -> per-warp smem transpose tile: after the call lane l finds row (8 it + rsel), float4
// column c4 of the 32 x 16 block at staged(it)
struct EpiStage {
  float* stage;
  int lane, c4, rsel;
  __device__ __forceinline__ void fill(uint32_t taddr) const {
    float v[16];
    tmem_ld16(taddr, v);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(stage + lane * STAGE_LD + 4 * q) =
          make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    __syncwarp();
  }
  __device__ __forceinline__ float4 get(int it) const {
    return *reinterpret_cast<const float4*>(stage + (it * 8 + rsel) * STAGE_LD + 4 * c4);
  }
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace fused
}  // namespace petb200
