// Grouped fp32 "NT" GEMM on the SIMT pipes (exact fp32 accumulate).
//
// This is the reference-grade dense kernel of the runtime: every contraction of
// the encoder (conv2/conv3 as strided-view GEMMs, QKV/O/FC1/FC2, QK^T, PV) and
// the cross-K/V projection can run through it, and the tcgen05 bf16x3 kernel
// (gemm_tc.cu) is validated against it.  128x128x16 CTA tile, 256 threads,
// 8x8 register tile per thread, register-staged double buffering.
#include <cuda_fp16.h>

#include "gemm_epilogue.cuh"
#include "kernels.h"

namespace msb {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int PAD = 4;

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
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= Mz) continue;
    const RowCtx rc = epilogue_row(p, c_base, m);
#pragma unroll
    for (int g = 0; g < 2; g++) epilogue_store4(p, rc, n0 + g * 64 + tx * 4, Nz, &acc[i][g * 4]);
  }
}

}  // namespace

void launch_gemm_simt(const GemmParams& p, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.groups <= 0) return;
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.groups);
  gemm_nt_kernel<<<grid, 256, 0, stream>>>(p);
}

}  // namespace msb
