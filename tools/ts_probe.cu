// Probe: tcgen05.mma with the A operand in TMEM (".ts" form) on sm_100a.
// A [128 x 64] bf16 is written to TMEM with tcgen05.st (lane = row, 32-bit column = 2 consecutive k),
// B [64 x 64] bf16 K-major SWIZZLE_128B in smem, D [128 x 64] fp32 in TMEM.  Prints max |D - ref|.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/build/ts_probe tools/ts_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + 8192, slot = base + 8192 + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // B tile: 64 rows x 64 k
  for (int idx = threadIdx.x; idx < 64 * 8; idx += 128) {
    const int row = idx >> 3, c = idx & 7;
    *reinterpret_cast<uint4*>(gen + swz(row, c)) = *reinterpret_cast<const uint4*>(B + row * 64 + c * 8);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(slot));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 8192 + 8);
  // A: row r = 32*warp + lane, 32 columns (64 bf16)
  const int r = 32 * warp + lane;
  uint32_t a[32];
  for (int c = 0; c < 32; ++c) a[c] = *reinterpret_cast<const uint32_t*>(A + r * 64 + 2 * c);
  const uint32_t ta = tmem + ((uint32_t)(32 * warp) << 16);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(ta),
      "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]),
      "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]), "r"(a[16]),
      "r"(a[17]), "r"(a[18]), "r"(a[19]), "r"(a[20]), "r"(a[21]), "r"(a[22]), "r"(a[23]), "r"(a[24]),
      "r"(a[25]), "r"(a[26]), "r"(a[27]), "r"(a[28]), "r"(a[29]), "r"(a[30]), "r"(a[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc(128, 64);
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t bdesc = make_smem_desc(base + kk * 32);
      const uint32_t at = tmem + kk * 8, dt = tmem + 64;
      const uint32_t acc = kk != 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(dt),
          "r"(at), "l"(bdesc), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  }
  // wait
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  for (int part = 0; part < 2; ++part) {
    const uint32_t td = tmem + ((uint32_t)(32 * warp) << 16) + 64 + part * 32;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(td));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) D[r * 64 + part * 32 + c] = __uint_as_float(v[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

int main() {
  const int M = 128, N = 64, K = 64;
  __nv_bfloat16 *hA = new __nv_bfloat16[M * K], *hB = new __nv_bfloat16[N * K];
  srand(1);
  for (int i = 0; i < M * K; ++i) hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.0f);
  for (int i = 0; i < N * K; ++i) hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 500.0f);
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * K * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  probe<<<1, 128, 16384>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  float* hD = new float[M * N];
  cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)__bfloat162float(hA[m * K + k]) * __bfloat162float(hB[n * K + k]);
      maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("TS-form probe: max |D - ref| = %.3e (max |ref| = %.3e) -> %s\n", maxerr, maxref, maxerr < 1e-3 ? "OK" : "MISMATCH");
  return 0;
}
