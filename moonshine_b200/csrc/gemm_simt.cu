// Grouped fp32 "NT" GEMM on the SIMT pipes (exact fp32 accumulate).
//
// This is the reference-grade dense kernel of the runtime: every contraction of
// the encoder (conv2/conv3 as strided-view GEMMs, QKV/O/FC1/FC2, QK^T, PV) and
// the cross-K/V projection can run through it, and the tcgen05 bf16x3 kernel
// (gemm_tc.cu) is validated against it.  128x128x16 CTA tile, 256 threads,
// 8x8 register tile per thread, register-staged double buffering.
#include <cuda_fp16.h>

#include "kernels.h"

namespace msb {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int PAD = 4;

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ int64_t split_off(int i, int m1, int m2, int64_t s1, int64_t s2,
                                             int64_t s3) {
  if (m1 == 0) return (int64_t)i * s3;
  int r = i % m1;
  return (int64_t)(i / m1) * s1 + (int64_t)(r / m2) * s2 + (int64_t)(r % m2) * s3;
}

__global__ void __launch_bounds__(256, 2) gemm_nt_kernel(GemmParams p) {
  const int z = blockIdx.z;
  const int Mz = p.Mz ? p.Mz[z] : p.M;
  const int Nz = p.Nz ? p.Nz[z] : p.N;
  const int Kz = p.Kz ? p.Kz[z] : p.K;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  if (m0 >= Mz || n0 >= Nz) return;

  const float* __restrict__ A = p.A + (p.offA ? p.offA[z] : (int64_t)z * p.strideA);
  const float* __restrict__ W = p.W + (p.offW ? p.offW[z] : (int64_t)z * p.strideW);
  const int64_t c_base = p.offC ? p.offC[z] : (int64_t)z * p.strideC;

  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Ws[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // global -> register staging: 2 float4 per operand per thread
  float4 ra[2], rw[2];
  int lrow[2], lkq[2];
#pragma unroll
  for (int i = 0; i < 2; i++) {
    int f = tid + i * 256;
    lrow[i] = f >> 2;
    lkq[i] = (f & 3) * 4;
  }
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      int k = k0 + lkq[i];
      int am = m0 + lrow[i];
      int wn = n0 + lrow[i];
      ra[i] = (am < Mz && k < Kz) ? *reinterpret_cast<const float4*>(A + (int64_t)am * p.lda + k)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[i] = (wn < Nz && k < Kz) ? *reinterpret_cast<const float4*>(W + (int64_t)wn * p.ldw + k)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      // K tail inside a float4 (K % 4 != 0): elements beyond K are row
      // padding and must not contribute, whatever they hold.
      if (k + 3 >= Kz && k < Kz) {
        if (k + 1 >= Kz) { ra[i].y = 0.f; rw[i].y = 0.f; }
        if (k + 2 >= Kz) { ra[i].z = 0.f; rw[i].z = 0.f; }
        if (k + 3 >= Kz) { ra[i].w = 0.f; rw[i].w = 0.f; }
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
      As[buf][lkq[i] + 0][lrow[i]] = ra[i].x;
      As[buf][lkq[i] + 1][lrow[i]] = ra[i].y;
      As[buf][lkq[i] + 2][lrow[i]] = ra[i].z;
      As[buf][lkq[i] + 3][lrow[i]] = ra[i].w;
      Ws[buf][lkq[i] + 0][lrow[i]] = rw[i].x;
      Ws[buf][lkq[i] + 1][lrow[i]] = rw[i].y;
      Ws[buf][lkq[i] + 2][lrow[i]] = rw[i].z;
      Ws[buf][lkq[i] + 3][lrow[i]] = rw[i].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  const int nk = (Kz + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; kt++) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue ----
  const int half_rot = p.rot_dim >> 1;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= Mz) continue;
    const int64_t roff = c_base + split_off(m, p.rm1, p.rm2, p.rs1, p.rs2, p.rs);
    const float bias_m = (p.bias && p.bias_on_m) ? p.bias[m] : 0.f;
    const int pos = p.pos ? p.pos[m] : 0;
#pragma unroll
    for (int g = 0; g < 2; g++) {
      const int n = n0 + g * 64 + tx * 4;
      if (n >= Nz) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        float x = acc[i][g * 4 + j] * p.alpha;
        if (p.bias) x += p.bias_on_m ? bias_m : ((n + j < Nz) ? p.bias[n + j] : 0.f);
        if (p.act == 1) x = gelu_erf(x);
        v[j] = x;
      }
      if (p.pos && n < p.rope_cols) {
        const int d = n % p.head_dim;  // head offsets are multiples of 4, pairs never straddle
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int dd = d + 2 * q;
          if (dd < p.rot_dim) {
            const float c = p.rope_cos[(int64_t)pos * half_rot + (dd >> 1)];
            const float s = p.rope_sin[(int64_t)pos * half_rot + (dd >> 1)];
            const float x0 = v[2 * q], x1 = v[2 * q + 1];
            v[2 * q] = x0 * c - x1 * s;
            v[2 * q + 1] = x1 * c + x0 * s;
          }
        }
      }
      const int64_t c0 = split_off(n, p.cm1, p.cm2, p.cs1, p.cs2, 1);
      const bool full = (n + 3 < Nz);
      const bool contig = full && (p.cm1 == 0 || ((n % p.cm2) + 3 < p.cm2));
      const int64_t addr = roff + c0;
      if (p.out_half) {
        __half* C = reinterpret_cast<__half*>(p.C);
        if (contig && (addr & 3) == 0) {
          __half2 h0 = __floats2half2_rn(v[0], v[1]);
          __half2 h1 = __floats2half2_rn(v[2], v[3]);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0);
          u.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(C + addr) = u;
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (n + j < Nz)
              C[roff + split_off(n + j, p.cm1, p.cm2, p.cs1, p.cs2, 1)] = __float2half_rn(v[j]);
        }
      } else {
        float* C = reinterpret_cast<float*>(p.C);
        if (contig && (addr & 3) == 0) {
          float4 o = make_float4(v[0], v[1], v[2], v[3]);
          if (p.accumulate) {
            float4 old = *reinterpret_cast<const float4*>(C + addr);
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *reinterpret_cast<float4*>(C + addr) = o;
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (n + j < Nz) {
              const int64_t a2 = roff + split_off(n + j, p.cm1, p.cm2, p.cs1, p.cs2, 1);
              C[a2] = p.accumulate ? C[a2] + v[j] : v[j];
            }
        }
      }
    }
  }
}

}  // namespace

void launch_gemm(const GemmParams& p, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.groups <= 0) return;
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.groups);
  gemm_nt_kernel<<<grid, 256, 0, stream>>>(p);
}

}  // namespace msb
