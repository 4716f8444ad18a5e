// Microbenchmark: sustained tcgen05.mma dispatch time (clk per instruction) for M=128, K=16, kind::f16 as a
// function of N, with the A operand in tensor memory (TS form, as the fused kernels use it) or in shared
// memory (SS form).  One thread per CTA issues `count` MMAs back to back, commits once and waits.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I metatrain_b200/csrc -o /tmp/mma_rate tools/mma_rate.cu
//   /tmp/mma_rate            (prints one table; used for DESIGN.md's "why the first GEMM runs at N=128")
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace petb200::tc;

// NOISE: 0 none; 1 = four warps stream tcgen05.ld over another TMEM region; 2 = four warps stream LDS/STS.128 over
// another shared-memory region; 3 = one thread streams 16 KB cp.async.bulk copies from global into shared memory
template <int N, bool TS, int NOISE>
__global__ void __launch_bounds__(256) rate_kernel(int count, long long* clk_out, const uint8_t* gsrc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_store;
  __shared__ __align__(8) uint64_t nbar_store;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  const uint32_t bar = smem_u32(&bar_store);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(smem_u32(&nbar_store), 1);
    done = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int k = threadIdx.x; k < 64 * 1024 / 4; k += blockDim.x) reinterpret_cast<uint32_t*>(smem)[k] = 0x3c003c00u;
  fence_proxy_async();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t b0 = smem_u32(smem);            // B: up to 256 rows x 64 k (32 KB), K-major SW128
    const uint32_t a0 = smem_u32(smem) + 32768;    // A (SS form): 128 rows x 64 k (16 KB)
    const long long t0 = clock64();
    for (int r = 0; r < count; r += 8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int ks = q & 3;
        const uint32_t d = tmem + 256 + (q >> 2) * 0;   // one accumulator (dependent chain, as in a real K loop)
        if (TS) tc_mma_ts(d, tmem + ks * 8, make_smem_desc(b0 + ks * 32), idesc, 1);
        else    tc_mma(d, make_smem_desc(a0 + ks * 32), make_smem_desc(b0 + ks * 32), idesc, 1);
      }
    }
    const long long t1 = clock64();
    tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { clk_out[0] = t1 - t0; clk_out[1] = t2 - t0; }
    done = 1;
  } else if (warp >= 4 && NOISE == 1) {
    float v[32];
    float acc = 0.f;
    while (!done) {
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 64, v);
      acc += v[0] + v[31];
    }
    if (acc == 1234.5f) clk_out[2] = 1;
  } else if (warp >= 4 && NOISE == 2) {
    float4* region = reinterpret_cast<float4*>(smem + 65536) + (warp - 4) * 1024;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    while (!done) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 x = region[k * 32 + (threadIdx.x & 31)];
        acc.x += x.x;
        region[(k + 8) * 32 + (threadIdx.x & 31)] = acc;
      }
    }
    if (acc.x == 1234.5f) clk_out[2] = 1;
  } else if (warp == 4 && NOISE == 3 && elect_one()) {
    const uint32_t nbar = smem_u32(&nbar_store);
    uint32_t phase = 0;
    while (!done) {
      mbar_expect_tx(nbar, 4 * 16384);
      for (int q = 0; q < 4; ++q) bulk_g2s(smem_u32(smem) + 65536 + q * 16384, gsrc + (size_t)blockIdx.x * 65536 + q * 16384, 16384, nbar);
      mbar_wait(nbar, phase);
      phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
  }
}

template <int N, bool TS, int NOISE = 0>
void run(int grid, long long* d_clk, const uint8_t* gsrc) {
  const int count = 4096;
  const int SMEM = 160 * 1024;
  cudaFuncSetAttribute(rate_kernel<N, TS, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  rate_kernel<N, TS, NOISE><<<grid, 256, SMEM>>>(count, d_clk, gsrc);
  rate_kernel<N, TS, NOISE><<<grid, 256, SMEM>>>(count, d_clk, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
  printf("| %s | %d | %3d | %4d | %6.1f | %6.1f | %5.1f | %s\n", TS ? "TS" : "SS", NOISE, N, grid, (double)h[0] / count,
         (double)h[1] / count, 128.0 * N / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_clk;
  cudaMalloc(&d_clk, 32);
  printf("| form | noise | N | CTAs | issue clk/MMA | complete clk/MMA | floor 128*N/256 |\n|---|---:|---:|---:|---:|---:|---:|\n");
  uint8_t* gsrc;
  cudaMalloc(&gsrc, (size_t)148 * 65536);
  cudaMemset(gsrc, 0, (size_t)148 * 65536);
  for (int grid : {148}) {
    run<32, true>(grid, d_clk, gsrc);
    run<128, true>(grid, d_clk, gsrc);
    run<32, false>(grid, d_clk, gsrc);
    run<32, true, 1>(grid, d_clk, gsrc);
    run<128, true, 1>(grid, d_clk, gsrc);
    run<32, true, 2>(grid, d_clk, gsrc);
    run<128, true, 2>(grid, d_clk, gsrc);
    run<32, true, 3>(grid, d_clk, gsrc);
    run<128, true, 3>(grid, d_clk, gsrc);
  }
  return 0;
}
