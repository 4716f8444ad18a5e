// Microbenchmark: aggregate L2 -> shared-memory bandwidth of 16 KB bulk async copies (cp.async.bulk) when every
// SM streams the SAME 384 KB region (the weight-streaming pattern of mlp_fused.cu), unicast vs cluster multicast.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/build/l2_bulk_bw tools/l2_bulk_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

constexpr int STAGE = 16384, RING = 8, REGION = 24 * STAGE;

// CS = cluster size; with multicast each CTA issues 1/CS of every stage and multicasts it to all CTAs of the cluster
template <int CS>
__global__ void bw_kernel(const uint8_t* src, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[RING];
  const uint32_t base = smem_u32(smem);
  uint32_t rank = 0;
  if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < RING; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CS > 1) cg::this_cluster().sync(); else __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    const int total = iters * 24;
    for (int q = 0; q < total + RING; ++q) {
      if (q >= RING) mbar_wait(smem_u32(&bars[(q - RING) % RING]), ((q - RING) / RING) & 1);   // consume
      if (CS > 1 && q >= RING && q < total) {
        // all CTAs of the cluster must have consumed the slot before anyone overwrites it
      }
      if (q < total) {
        const uint32_t bar = smem_u32(&bars[q % RING]);
        const uint32_t dst = base + (q % RING) * STAGE;
        const uint8_t* s = src + (size_t)(q % 24) * STAGE;
        mbar_expect_tx(bar, STAGE);
        if (CS == 1) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(s), "r"(STAGE), "r"(bar) : "memory");
        } else {
          const uint32_t part = STAGE / CS;
          const uint16_t mask = (1u << CS) - 1;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst + rank * part), "l"(s + rank * part), "r"(part), "r"(bar), "h"(mask) : "memory");
        }
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  if (CS > 1) cg::this_cluster().sync(); else __syncthreads();
}

template <int CS>
void run(const uint8_t* src, long long* d_cycles) {
  const int iters = 200, grid = 148 / CS * CS;
  cudaFuncSetAttribute(bw_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, RING * STAGE);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = RING * STAGE;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, bw_kernel<CS>, src, 10, d_cycles);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, bw_kernel<CS>, src, iters, d_cycles);
  cudaEventRecord(e1); cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)grid * iters * REGION;
  printf("cluster %d (%s): %.1f us, %.2f TB/s delivered to shared memory, %.2f TB/s read from L2  [%s]\n", CS,
         CS == 1 ? "unicast" : "multicast", ms * 1e3, bytes / ms / 1e9, bytes / CS / ms / 1e9, cudaGetErrorString(err));
}

int main() {
  uint8_t* src; cudaMalloc(&src, REGION); cudaMemset(src, 1, REGION);
  long long* d_cycles; cudaMalloc(&d_cycles, 148 * 8);
  run<1>(src, d_cycles); run<2>(src, d_cycles); run<4>(src, d_cycles);
  return 0;
}
