// fp32 SIMT GEMM with fused epilogues:  C[M,N] = epi(A[M,K] * W[N,K]^T).
//
// This is the plain-FFMA kernel: the exact-fp32 parity reference for the tensor-core
// path (gemm_tc.cu) and the fallback for shapes the tcgen05 kernel does not cover.  Both
// operands are K-contiguous ("TN"), which is how torch.nn.Linear stores its weight
// ([out, in], reference: every F.linear in src/metatrain/pet/modules/transformer.py).
//
// Tiling: 128x128x16 per CTA, 256 threads, 8x8 register tile per thread split 4+4 in both
// directions (rows ty*4+{0..3}, 64+ty*4+{0..3}; cols tx*4+{0..3}, 64+tx*4+{0..3}) so the
// smem fragment reads are float4 and conflict free; register-prefetch double buffering.
#include "common.cuh"
#include "kernels.cuh"

namespace petb200 {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

struct Frag {
  float4 a[2];
  float4 b[2];
};

// Map a tile-local output column to the global weight row / output column.
// Plain layouts: identity.  SwiGLU forward: tile columns [0,64) are "value" columns
// n0/2 + c and [64,128) the matching "gate" columns F + n0/2 + (c-64)
// (transformer.py:40-44: v, g = w_in(x).chunk(2)).
template <int EPI>
__device__ __forceinline__ int weight_row(int n0, int c, int F) {
  if (EPI == PETB200_EPI_SWIGLU) {
    return (c < 64) ? (n0 / 2 + c) : (F + n0 / 2 + (c - 64));
  }
  return n0 + c;
}

template <int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int F = g.N / 2;  // only meaningful for SWIGLU (g.N = 2F weight rows)

  // global->smem load assignment: 512 float4 per operand tile, 2 per thread.
  int l_row[2], l_kq[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int idx = tid + i * 256;
    l_row[i] = idx >> 2;
    l_kq[i] = idx & 3;
  }
  const float* a_ptr[2];
  const float* b_ptr[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int64_t m = m0 + l_row[i];
    a_ok[i] = m < g.M;
    a_ptr[i] = g.A + (a_ok[i] ? m : 0) * g.lda + l_kq[i] * 4;
    b_ptr[i] = g.W + (int64_t)weight_row<EPI>(n0, l_row[i], F) * g.ldw + l_kq[i] * 4;
  }

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto load_global = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ra[i] = a_ok[i] ? __ldg(reinterpret_cast<const float4*>(a_ptr[i] + k0))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[i] = __ldg(reinterpret_cast<const float4*>(b_ptr[i] + k0));
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int k = l_kq[i] * 4, r = l_row[i];
      As[buf][k + 0][r] = ra[i].x;
      As[buf][k + 1][r] = ra[i].y;
      As[buf][k + 2][r] = ra[i].z;
      As[buf][k + 3][r] = ra[i].w;
      Bs[buf][k + 0][r] = rb[i].x;
      Bs[buf][k + 1][r] = rb[i].y;
      Bs[buf][k + 2][r] = rb[i].z;
      Bs[buf][k + 3][r] = rb[i].w;
    }
  };

  const int num_k = g.K / BK;
  load_global(0);
  store_smem(0);
  __syncthreads();
  for (int kt = 0; kt < num_k; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < num_k) load_global((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < num_k) {
      store_smem(buf ^ 1);
      __syncthreads();
    }
  }

  // ------------------------------------------------------------------ epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
    const float rs = g.row_scale ? g.row_scale[m] : 1.0f;
    if (EPI == PETB200_EPI_SWIGLU) {
      // columns j (value) and j+4 (gate) of this thread are a SwiGLU pair
      const int cu = n0 / 2 + tx * 4;  // first output column of this thread
      float u[4], gt[4], o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        u[j] = rs * acc[i][j] + (g.bias ? g.bias[cu + j] : 0.f);
        gt[j] = rs * acc[i][j + 4] + (g.bias ? g.bias[F + cu + j] : 0.f);
        o[j] = u[j] * sigmoidf_(gt[j]);
      }
      if (g.aux_out) {
        *reinterpret_cast<float4*>(g.aux_out + m * g.ld_aux + cu) = make_float4(u[0], u[1], u[2], u[3]);
        *reinterpret_cast<float4*>(g.aux_out + m * g.ld_aux + F + cu) =
            make_float4(gt[0], gt[1], gt[2], gt[3]);
      }
      *reinterpret_cast<float4*>(g.C + m * g.ldc + cu) = make_float4(o[0], o[1], o[2], o[3]);
    } else if (EPI == PETB200_EPI_SWIGLU_BWD) {
      // acc = d_s[m, c], c in [0, N); aux_in = [u | g] (ld_aux, 2N columns);
      // C[m, c] = d_s * sigma(g), C[m, N + c] = d_s * u * sigma(g) * (1 - sigma(g))
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c0 = n0 + h * 64 + tx * 4;
        float4 uu = *reinterpret_cast<const float4*>(g.aux_in + m * g.ld_aux + c0);
        float4 gg = *reinterpret_cast<const float4*>(g.aux_in + m * g.ld_aux + g.N + c0);
        float uv[4] = {uu.x, uu.y, uu.z, uu.w}, gv[4] = {gg.x, gg.y, gg.z, gg.w};
        float du[4], dg[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float ds = acc[i][h * 4 + j];
          float s = sigmoidf_(gv[j]);
          du[j] = ds * s;
          dg[j] = ds * uv[j] * s * (1.f - s);
        }
        *reinterpret_cast<float4*>(g.C + m * g.ldc + c0) = make_float4(du[0], du[1], du[2], du[3]);
        *reinterpret_cast<float4*>(g.C + m * g.ldc + g.N + c0) =
            make_float4(dg[0], dg[1], dg[2], dg[3]);
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c0 = n0 + h * 64 + tx * 4;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = rs * acc[i][h * 4 + j] + (g.bias ? g.bias[c0 + j] : 0.f);
        if (EPI == PETB200_EPI_SILU) {
          if (g.aux_out)
            *reinterpret_cast<float4*>(g.aux_out + m * g.ld_aux + c0) =
                make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = siluf_(v[j]);
        }
        if (EPI == PETB200_EPI_MUL_DSILU) {
          float4 p = *reinterpret_cast<const float4*>(g.aux_in + m * g.ld_aux + c0);
          v[0] *= dsiluf_(p.x);
          v[1] *= dsiluf_(p.y);
          v[2] *= dsiluf_(p.z);
          v[3] *= dsiluf_(p.w);
        }
        if (g.residual) {
          float4 r = *reinterpret_cast<const float4*>(g.residual + m * g.ldr + c0);
          v[0] += r.x;
          v[1] += r.y;
          v[2] += r.z;
          v[3] += r.w;
        }
        float4* dst = reinterpret_cast<float4*>(g.C + m * g.ldc + c0);
        if (g.accumulate) {
          float4 old = *dst;
          v[0] += old.x;
          v[1] += old.y;
          v[2] += old.z;
          v[3] += old.w;
        }
        *dst = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
}

}  // namespace

int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream) {
  if (g.M == 0) return PETB200_OK;
  dim3 grid((unsigned)(g.N / BN), (unsigned)ceil_div(g.M, BM));
  dim3 block(256);
  switch (g.epilogue) {
    case PETB200_EPI_NONE:
      gemm_simt_kernel<PETB200_EPI_NONE><<<grid, block, 0, stream>>>(g);
      break;
    case PETB200_EPI_SILU:
      gemm_simt_kernel<PETB200_EPI_SILU><<<grid, block, 0, stream>>>(g);
      break;
    case PETB200_EPI_SWIGLU:
      gemm_simt_kernel<PETB200_EPI_SWIGLU><<<grid, block, 0, stream>>>(g);
      break;
    case PETB200_EPI_MUL_DSILU:
      gemm_simt_kernel<PETB200_EPI_MUL_DSILU><<<grid, block, 0, stream>>>(g);
      break;
    case PETB200_EPI_SWIGLU_BWD:
      gemm_simt_kernel<PETB200_EPI_SWIGLU_BWD><<<grid, block, 0, stream>>>(g);
      break;
    case PETB200_EPI_RMS_BWD:
      set_error("gemm: the RMSNorm-backward epilogue is built for the tensor-core precisions only "
                "(use petb200_rms_bwd after a plain fp32 gemm)");
      return PETB200_ERR_UNSUPPORTED;
    default:
      set_error("gemm: unknown epilogue %d", g.epilogue);
      return PETB200_ERR_INVALID_ARGUMENT;
  }
  return check_launch("gemm_simt");
}

}  // namespace petb200
