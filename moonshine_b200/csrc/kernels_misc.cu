// Frontend conv1 / GroupNorm, LayerNorm, softmax and small element-wise kernels.
#include <cuda_fp16.h>

#include "kernels.h"

#include <algorithm>

namespace msb {

namespace {

constexpr int kConv1Frames = 16;  // output frames per CTA
constexpr int kConv1Taps = 127;
constexpr int kConv1Stride = 64;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// grid (blocks_per_utt, B), block = round_up(D, 32) threads (thread = channel).
__global__ void conv1_tanh_kernel(const float* __restrict__ pcm, int64_t pcm_stride,
                                  const int* __restrict__ n_samples, const int* __restrict__ t1,
                                  const int64_t* __restrict__ off1, const float* __restrict__ w1t,
                                  float* __restrict__ h1, int D, double* __restrict__ gn_partial) {
  constexpr int kWin = kConv1Stride * (kConv1Frames - 1) + kConv1Taps + 1;  // +1 -> multiple of 4
  __shared__ __align__(16) float xs[kWin];
  __shared__ float red[2][32];
  const int b = blockIdx.y;
  const int T1 = t1[b];
  const int t0 = blockIdx.x * kConv1Frames;
  const int c = threadIdx.x;
  float s1 = 0.f, s2 = 0.f;
  if (t0 < T1) {  // block-uniform
    const float* x = pcm + (int64_t)b * pcm_stride + (int64_t)t0 * kConv1Stride;
    const int navail = n_samples[b] - t0 * kConv1Stride;
    for (int i = threadIdx.x; i < kWin; i += blockDim.x) xs[i] = i < navail ? x[i] : 0.f;
    __syncthreads();
    if (c < D) {
      float acc[kConv1Frames];
#pragma unroll
      for (int t = 0; t < kConv1Frames; t++) acc[t] = 0.f;
      for (int j = 0; j < 124; j += 4) {
        const float w0 = w1t[(j + 0) * D + c], w1 = w1t[(j + 1) * D + c];
        const float w2 = w1t[(j + 2) * D + c], w3 = w1t[(j + 3) * D + c];
#pragma unroll
        for (int t = 0; t < kConv1Frames; t++) {
          const float4 v = *reinterpret_cast<const float4*>(&xs[t * kConv1Stride + j]);
          acc[t] = fmaf(w0, v.x, acc[t]);
          acc[t] = fmaf(w1, v.y, acc[t]);
          acc[t] = fmaf(w2, v.z, acc[t]);
          acc[t] = fmaf(w3, v.w, acc[t]);
        }
      }
      {
        const float w0 = w1t[124 * D + c], w1 = w1t[125 * D + c], w2 = w1t[126 * D + c];
#pragma unroll
        for (int t = 0; t < kConv1Frames; t++) {
          const float4 v = *reinterpret_cast<const float4*>(&xs[t * kConv1Stride + 124]);
          acc[t] = fmaf(w0, v.x, acc[t]);
          acc[t] = fmaf(w1, v.y, acc[t]);
          acc[t] = fmaf(w2, v.z, acc[t]);
        }
      }
      float* out = h1 + (off1[b] + t0) * D + c;
#pragma unroll
      for (int t = 0; t < kConv1Frames; t++) {
        if (t0 + t < T1) {
          const float y = tanhf(acc[t]);
          out[(int64_t)t * D] = y;
          s1 += y;
          s2 += y * y;
        }
      }
    }
  }
  // block reduce -> partial sums (double) for the GroupNorm statistics
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = s1; red[1][w] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, q = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; i++) { a += red[0][i]; q += red[1][i]; }
    double* dst = gn_partial + ((int64_t)b * gridDim.x + blockIdx.x) * 2;
    dst[0] = a;
    dst[1] = q;
  }
}

// grid (chunks, B), 256 threads; float4 over the contiguous [T1_b * D] block.
__global__ void groupnorm_apply_kernel(float* __restrict__ h1, const int* __restrict__ t1,
                                       const int64_t* __restrict__ off1,
                                       const double* __restrict__ gn_partial, int nblk,
                                       const float* __restrict__ gamma,
                                       const float* __restrict__ beta, int D) {
  __shared__ float stat[2];
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    double a = 0.0, q = 0.0;
    for (int i = 0; i < nblk; i++) {
      a += gn_partial[((int64_t)b * nblk + i) * 2];
      q += gn_partial[((int64_t)b * nblk + i) * 2 + 1];
    }
    const double n = (double)t1[b] * D;
    const double mean = n > 0 ? a / n : 0.0;
    double var = n > 0 ? q / n - mean * mean : 0.0;
    if (var < 0) var = 0;
    stat[0] = (float)mean;
    stat[1] = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1];
  const int64_t total4 = (int64_t)t1[b] * D / 4;  // D % 4 == 0
  float4* base = reinterpret_cast<float4*>(h1 + off1[b] * D);
  const int D4 = D / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % D4);
    float4 v = base[i];
    const float4 g = reinterpret_cast<const float4*>(gamma)[c4];
    const float4 be = reinterpret_cast<const float4*>(beta)[c4];
    v.x = (v.x - mean) * rstd * g.x + be.x;
    v.y = (v.y - mean) * rstd * g.y + be.y;
    v.z = (v.z - mean) * rstd * g.z + be.z;
    v.w = (v.w - mean) * rstd * g.w + be.w;
    base[i] = v;
  }
}

// one warp per row, D <= 32 * kMaxPerLane
constexpr int kLnMaxPerLane = 16;
__global__ void layernorm_kernel(const float* __restrict__ x, float* __restrict__ y,
                                 const float* __restrict__ gamma, int64_t rows, int D) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * D;
  float v[kLnMaxPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; i++) {
    const int c = lane + i * 32;
    v[i] = c < D ? xr[c] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; i++) {
    const int c = lane + i * 32;
    const float d = c < D ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / D + 1e-5f);
  float* yr = y + row * D;
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; i++) {
    const int c = lane + i * 32;
    if (c < D) yr[c] = (v[i] - mean) * rstd * gamma[c];
  }
}

// grid (ceil(max_m / warps), groups); one warp per score row.
__global__ void softmax_rows_kernel(float* __restrict__ S, const int64_t* __restrict__ offS,
                                    const int* __restrict__ Mz, const int* __restrict__ Nz,
                                    int ld, int win_past, int win_future) {
  const int z = blockIdx.y;
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= Mz[z]) return;
  const int n = Nz[z];
  const int lane = threadIdx.x & 31;
  float* row = S + offS[z] + (int64_t)m * ld;
  // allowed columns [lo, hi)
  int lo = 0, hi = n;
  if (win_past >= 0) {
    lo = max(0, m - win_past);
    hi = min(n, m + win_future + 1);
  }
  float mx = -INFINITY;
  for (int j = lo + lane; j < hi; j += 32) mx = fmaxf(mx, row[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lo + lane; j < hi; j += 32) {
    const float e = expf(row[j] - mx);
    row[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lo + lane; j < hi; j += 32) row[j] *= inv;
  for (int j = lane; j < lo; j += 32) row[j] = 0.f;
  for (int j = hi + lane; j < ld; j += 32) row[j] = 0.f;
}

// one warp per 80-sample frame
__global__ void stream_frames_kernel(const float* __restrict__ pcm, int64_t pcm_stride,
                                     const int* __restrict__ n_frames, const int64_t* __restrict__ row0,
                                     float k, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= n_frames[b]) return;
  const int lane = threadIdx.x & 31;
  const float* src = pcm + (int64_t)b * pcm_stride + (int64_t)f * 80;
  float v[3];
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 3; u++) {
    const int i = lane + 32 * u;
    v[u] = i < 80 ? src[i] : 0.f;
    s += v[u];
  }
  const float mean = warp_sum(s) * (1.0f / 80.0f);
  float q = 0.f;
#pragma unroll
  for (int u = 0; u < 3; u++) {
    const int i = lane + 32 * u;
    v[u] = i < 80 ? v[u] - mean : 0.f;
    q += v[u] * v[u];
  }
  const float rms = sqrtf(warp_sum(q) * (1.0f / 80.0f) + 1e-6f);
  float* dst = out + (row0[b] + f) * 80;
#pragma unroll
  for (int u = 0; u < 3; u++) {
    const int i = lane + 32 * u;
    if (i < 80) dst[i] = asinhf(k * (v[u] / rms));
  }
}

__global__ void add_rows_by_index_kernel(const float* __restrict__ x, const float* __restrict__ table,
                                         const int* __restrict__ pos, float* __restrict__ y, int64_t rows,
                                         int D4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * D4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D4;
    const int c = (int)(i - r * D4);
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    const float4 t = reinterpret_cast<const float4*>(table)[(int64_t)pos[r] * D4 + c];
    reinterpret_cast<float4*>(y)[i] = make_float4(a.x + t.x, a.y + t.y, a.z + t.z, a.w + t.w);
  }
}

__global__ void gelu_inplace_kernel(float* x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    x[i] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  }
}

}  // namespace

int conv1_blocks_per_utt(int max_t1) { return (max_t1 + kConv1Frames - 1) / kConv1Frames; }

void launch_conv1_tanh(const float* pcm, int64_t pcm_stride, const int* n_samples, const int* t1,
                       const int64_t* off1, const float* w1t, float* h1, int D, int B, int max_t1,
                       double* gn_partial, int* nblk_out, cudaStream_t stream) {
  const int nblk = conv1_blocks_per_utt(max_t1);
  if (nblk_out) *nblk_out = nblk;
  if (nblk == 0 || B == 0) return;
  dim3 grid(nblk, B);
  const int threads = (D + 31) / 32 * 32;
  conv1_tanh_kernel<<<grid, threads, 0, stream>>>(pcm, pcm_stride, n_samples, t1, off1, w1t, h1, D,
                                                  gn_partial);
}

void launch_groupnorm_apply(float* h1, const int* t1, const int64_t* off1, const double* gn_partial,
                            int nblk, const float* gamma, const float* beta, int D, int B,
                            int max_t1, cudaStream_t stream) {
  if (B == 0 || max_t1 == 0) return;
  int64_t total4 = (int64_t)max_t1 * D / 4;
  int chunks = (int)((total4 + 256 * 8 - 1) / (256 * 8));
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, B);
  groupnorm_apply_kernel<<<grid, 256, 0, stream>>>(h1, t1, off1, gn_partial, nblk, gamma, beta, D);
}

void launch_layernorm(const float* x, float* y, const float* gamma, int64_t rows, int D,
                      cudaStream_t stream) {
  if (rows == 0) return;
  const int warps = 8;
  const int64_t blocks = (rows + warps - 1) / warps;
  layernorm_kernel<<<(unsigned)blocks, warps * 32, 0, stream>>>(x, y, gamma, rows, D);
}

void launch_softmax_rows(float* S, const int64_t* offS, const int* Mz, const int* Nz, int ld,
                         int groups, int max_m, cudaStream_t stream, int win_past, int win_future) {
  if (groups == 0 || max_m == 0) return;
  const int warps = 8;
  dim3 grid((max_m + warps - 1) / warps, groups);
  softmax_rows_kernel<<<grid, warps * 32, 0, stream>>>(S, offS, Mz, Nz, ld, win_past, win_future);
}

void launch_stream_frames(const float* pcm, int64_t pcm_stride, const int* n_frames, const int64_t* row0,
                          float k, float* out, int B, int max_frames, cudaStream_t stream) {
  if (B == 0 || max_frames == 0) return;
  const int warps = 8;
  dim3 grid((max_frames + warps - 1) / warps, B);
  stream_frames_kernel<<<grid, warps * 32, 0, stream>>>(pcm, pcm_stride, n_frames, row0, k, out);
}

void launch_add_rows_by_index(const float* x, const float* table, const int* pos, float* y, int64_t rows,
                              int D, cudaStream_t stream) {
  if (rows == 0) return;
  const int64_t n = rows * (D / 4);
  int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  add_rows_by_index_kernel<<<blocks, 256, 0, stream>>>(x, table, pos, y, rows, D / 4);
}

void launch_gelu_inplace(float* x, int64_t n, cudaStream_t stream) {
  if (n == 0) return;
  int blocks = (int)((n + 256 * 4 - 1) / (256 * 4));
  if (blocks > 148 * 16) blocks = 148 * 16;
  gelu_inplace_kernel<<<blocks, 256, 0, stream>>>(x, n);
}

}  // namespace msb
