// Shared GEMM epilogue: alpha, bias, exact-erf GELU, interleaved RoPE,
// residual accumulate, fp16 conversion and split (head-major) addressing for
// 4 consecutive output columns of one row.  Used by the SIMT kernel and the
// tcgen05 kernel so both produce bit-identical post-processing.
#pragma once
#include <cuda_fp16.h>

#include "kernels.h"

namespace msb {

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ int64_t split_off(int i, int m1, int m2, int64_t s1, int64_t s2, int64_t s3) {
  if (m1 == 0) return (int64_t)i * s3;
  int r = i % m1;
  return (int64_t)(i / m1) * s1 + (int64_t)(r / m2) * s2 + (int64_t)(r % m2) * s3;
}

struct RowCtx {
  int64_t roff;   // c_base + rowoff(m)
  float bias_m;
  int pos;
};

__device__ __forceinline__ RowCtx epilogue_row(const GemmParams& p, int64_t c_base, int m) {
  RowCtx r;
  r.roff = c_base + split_off(m, p.rm1, p.rm2, p.rs1, p.rs2, p.rs);
  r.bias_m = (p.bias && p.bias_on_m) ? p.bias[m] : 0.f;
  r.pos = p.pos ? p.pos[m] : 0;
  return r;
}

// acc[0..3] = raw accumulators of columns n .. n+3 (n % 4 == 0) of the row.
__device__ __forceinline__ void epilogue_store4(const GemmParams& p, const RowCtx& r, int n, int Nz,
                                                const float* acc) {
  if (n >= Nz) return;
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    float x = acc[j] * p.alpha;
    if (p.bias) x += p.bias_on_m ? r.bias_m : ((n + j < Nz) ? p.bias[n + j] : 0.f);
    if (p.act == 1) x = gelu_erf(x);
    else if (p.act == 2) x = x / (1.0f + expf(-x));  // SiLU
    v[j] = x;
  }
  if (p.pos && n < p.rope_cols) {
    const int half_rot = p.rot_dim >> 1;
    const int d = n % p.head_dim;  // head offsets are multiples of 4: pairs never straddle
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int dd = d + 2 * q;
      if (dd < p.rot_dim) {
        const float c = p.rope_cos[(int64_t)r.pos * half_rot + (dd >> 1)];
        const float s = p.rope_sin[(int64_t)r.pos * half_rot + (dd >> 1)];
        const float x0 = v[2 * q], x1 = v[2 * q + 1];
        v[2 * q] = x0 * c - x1 * s;
        v[2 * q + 1] = x1 * c + x0 * s;
      }
    }
  }
  const int64_t c0 = split_off(n, p.cm1, p.cm2, p.cs1, p.cs2, 1);
  const bool full = (n + 3 < Nz);
  const bool contig = full && (p.cm1 == 0 || ((n % p.cm2) + 3 < p.cm2));
  const int64_t addr = r.roff + c0;
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
        if (n + j < Nz) C[r.roff + split_off(n + j, p.cm1, p.cm2, p.cs1, p.cs2, 1)] = __float2half_rn(v[j]);
    }
  } else {
    float* C = reinterpret_cast<float*>(p.C);
    if (contig && (addr & 3) == 0) {
      float4 o = make_float4(v[0], v[1], v[2], v[3]);
      if (p.accumulate) {
        const float4 old = *reinterpret_cast<const float4*>(C + addr);
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *reinterpret_cast<float4*>(C + addr) = o;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (n + j < Nz) {
          const int64_t a2 = r.roff + split_off(n + j, p.cm1, p.cm2, p.cs1, p.cs2, 1);
          C[a2] = p.accumulate ? C[a2] + v[j] : v[j];
        }
    }
  }
}

}  // namespace msb
