// Microbenchmark: legacy mma.sync throughput on sm_100a (bf16 m16n8k16, tf32 m16n8k8) vs FFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/build/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_bf16(float* out, int iters) {
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x, 5u};
  float c[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_tf32(float* out, int iters) {
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x, 5u};
  float c[ILP][4];
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_ffma(float* out, int iters) {
  float c[ILP], a = threadIdx.x * 1e-3f, b = 1.0001f;
  for (int i = 0; i < ILP; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fmaf(c[i], b, a);
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    for (int ctas : {1, 2}) {
      if (warps * ctas > 64) continue;
      int grid = 148 * ctas, block = warps * 32;
      float ms = time_ms([&] { k_bf16<8><<<grid, block>>>(out, iters); });
      double fl = (double)grid * warps * iters * 8 * 2.0 * 16 * 8 * 16;
      printf("bf16 m16n8k16 warps/CTA=%2d CTAs/SM=%d: %8.1f TFLOP/s\n", warps, ctas, fl / ms * 1e-9);
      ms = time_ms([&] { k_tf32<8><<<grid, block>>>(out, iters); });
      fl = (double)grid * warps * iters * 8 * 2.0 * 16 * 8 * 8;
      printf("tf32 m16n8k8  warps/CTA=%2d CTAs/SM=%d: %8.1f TFLOP/s\n", warps, ctas, fl / ms * 1e-9);
    }
  }
  {
    int grid = 148 * 2, block = 1024;
    float ms = time_ms([&] { k_ffma<8><<<grid, block>>>(out, iters); });
    double fl = (double)grid * block * iters * 8 * 2.0;
    printf("ffma: %8.1f TFLOP/s\n", fl / ms * 1e-9);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
