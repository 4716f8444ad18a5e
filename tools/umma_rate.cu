// Microbenchmark: tcgen05.mma (kind::f16, M=128, K=16) cycles per instruction on sm_100a for several N,
// with A in shared memory (SS) or tensor memory (TS); all 148 SMs busy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 --expt-relaxed-constexpr -o tools/build/umma_rate tools/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int N, bool TS, int DISTINCT_B, int NACC = 2>
__global__ void __launch_bounds__(128) rate(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + 160 * 1024, slot = bar + 8;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 160 * 1024 + 8);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, N);
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          // A tile at 0 (16 KB), B tiles from 32 KB on; rotate over DISTINCT_B different B tiles / k-steps
          const uint64_t adesc = make_smem_desc(base + (u & 3) * 32);
          const uint64_t bdesc = make_smem_desc(base + 32768 + ((u % DISTINCT_B) * 32768) + (u & 3) * 32);
          const uint32_t dt = tmem + (u % NACC) * 128, at = tmem + 496 + 0 * u;
          if (TS) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(dt), "r"(at),
                         "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
          } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dt), "l"(adesc),
                         "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(0u) : "memory");
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int N, bool TS, int DB, int NACC = 2>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int smem = 160 * 1024 + 1024 + 64, iters = 2000;
  cudaFuncSetAttribute(rate<N, TS, DB, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate<N, TS, DB, NACC><<<148, 128, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const double cyc = avg / (iters * 8.0);
  printf("%-28s N=%3d  %7.1f cycles/MMA  (floor N/2 = %3d)  -> %6.0f TFLOP/s chip @1.965 GHz  [%s]\n", name, N, cyc, N / 2,
         2.0 * 128 * N * 16 / cyc * 1.965e9 * 148 / 1e12, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<32, false, 1>("SS"); run<64, false, 1>("SS"); run<128, false, 1>("SS"); run<256, false, 1>("SS");
  run<64, false, 4>("SS 4 distinct B tiles"); run<128, false, 4>("SS 4 distinct B tiles");
  run<16, true, 1>("TS"); run<32, true, 1>("TS"); run<64, true, 1>("TS"); run<128, true, 1>("TS"); run<256, true, 1>("TS");
  run<32, true, 1, 1>("TS same accumulator"); run<64, true, 1, 1>("TS same accumulator"); run<128, true, 1, 1>("TS same accumulator");
  run<64, false, 1, 1>("SS same accumulator"); run<128, false, 1, 1>("SS same accumulator"); run<256, false, 1, 1>("SS same accumulator");
  run<64, true, 1, 4>("TS 4 accumulators");
  return 0;
}
