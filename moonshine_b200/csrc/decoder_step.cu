// Persistent decoder-step kernel: ONE launch per greedy token.
//
// Replaces one `decoder_session` Run of the reference's hot loop
// (core/moonshine-model.cpp:380-517: embed -> L x {self-attn with KV cache,
// cross-attn, gated-SiLU MLP} -> final LN -> tied-embedding logits -> argmax)
// for a whole batch of utterances.
//
// Structure (cooperative launch, one CTA per SM, software grid barrier):
//   per layer, three phases separated by a grid barrier
//     A  item (b-tile, head):  [resolve h] LN1, QKV_h, RoPE, KV append,
//                              causal self-attn, O-proj partial -> partA[h]
//     B  item (b-tile, head):  [h += sum partA] LN2, Qc_h, cross-attn over the
//                              fp16 cross K/V, O-proj partial -> partB[h]
//     C  item (b-tile, chunk): [h += sum partB] LN3, FC1 chunk, SiLU gate,
//                              FC2 partial -> partC[chunk]
//   then   F  item (b-tile):   [h += sum partC + b2] final LN -> xfin
//          G  item (vocab chunk): logits chunk + per-utterance argmax candidate
// Partial sums are combined in a fixed order by the consumer (deterministic).
// The argmax candidates of step t are resolved in the prologue of step t+1
// (and by decoder_finalize after the last step).
#include <cuda_fp16.h>

#include "common.h"
#include "kernels.h"

namespace msb {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kLogitsTile = 32;  // utterances per pass of the logits phase

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

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Monotonic-counter grid barrier: bar[0] counts arrivals for ever; `target` is
// the arrival count that releases this barrier.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while ((int)(ld_acquire(bar) - target) < 0) {
    }
    __threadfence();
  }
  __syncthreads();
}

struct Smem {
  float* hs;    // [NB][D]   resolved residual rows / GEMM output staging
  float* xs;    // [NB][D]   normalised rows
  float* act;   // [NB][actw] q|k|v, q, or fc1 output
  float* att;   // [NB][attw] attention output / gated activation
  float* red;   // split-K scratch [1024 * NB]; aliased by the logits x tile
  float* ps;    // [2][Tps]  cross-attention probabilities (one per half-block)
  float* sc;    // [kWarps][Smax] self-attention scores
  int* flags;   // [2 * 32]  done flags / tokens
  float* argv;  // [kWarps][kLogitsTile] per-warp argmax values (logits phase)
  int* argi;    // [kWarps][kLogitsTile] per-warp argmax indices
};

struct SmemLayout {
  int hs, xs, act, att, red, ps, sc, flags, argv, argi, total;  // float offsets
  int actw, attw, Tps;
};

__host__ __device__ inline SmemLayout smem_layout(int NBmax, int D, int hd, int IC, int Tpad,
                                                  int Smax) {
  SmemLayout L;
  L.actw = max(3 * hd, 2 * IC);
  L.attw = max(hd, IC);
  L.Tps = Tpad;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) / 4 * 4; return r; };
  L.hs = take(NBmax * D);
  L.xs = take(NBmax * D);
  L.act = take(NBmax * L.actw);
  L.att = take(NBmax * L.attw);
  int red = 1024 * NBmax;
  if (red < D * kLogitsTile) red = D * kLogitsTile;
  L.red = take(red);
  L.ps = take(2 * L.Tps);
  L.sc = take(kWarps * Smax);
  L.flags = take(64);
  L.argv = take(kWarps * kLogitsTile);
  L.argi = take(kWarps * kLogitsTile);
  L.total = o;
  return L;
}

// out[b][n] = sum_k x[b][k] * Wt[k][n] (+bias[n]);  x in smem (row stride ldx),
// Wt global, k-major, row stride N (N % 4 == 0).  Threads = (k-slice, float4 of
// features); slices are summed in a fixed order.  Result -> out (smem).
template <int NB>
__device__ __forceinline__ void tile_gemm(const float* x, int ldx, int K,
                                          const float* __restrict__ Wt, int N,
                                          const float* __restrict__ bias, float* red, float* out,
                                          int ldo) {
  const int N4 = N >> 2;            // N4 <= 256 for every shape of this model
  int S = kThreads / N4;
  if (S > K) S = K;
  const int Ks = (K + S - 1) / S;
  {
    const int t = threadIdx.x;
    const int n4 = t % N4;
    const int s = t / N4;
    if (s < S) {
      float acc[NB][4];
#pragma unroll
      for (int b = 0; b < NB; b++) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.f;
      const int k0 = s * Ks;
      const int k1 = min(K, k0 + Ks);
      const float4* wp = reinterpret_cast<const float4*>(Wt) + (int64_t)k0 * N4 + n4;
#pragma unroll 8
      for (int k = k0; k < k1; k++) {
        const float4 w = __ldg(wp);
        wp += N4;
#pragma unroll
        for (int b = 0; b < NB; b++) {
          const float xv = x[b * ldx + k];
          acc[b][0] = fmaf(xv, w.x, acc[b][0]);
          acc[b][1] = fmaf(xv, w.y, acc[b][1]);
          acc[b][2] = fmaf(xv, w.z, acc[b][2]);
          acc[b][3] = fmaf(xv, w.w, acc[b][3]);
        }
      }
#pragma unroll
      for (int b = 0; b < NB; b++)
        *reinterpret_cast<float4*>(&red[(s * NB + b) * N + n4 * 4]) =
            make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NB * N; i += kThreads) {
    const int b = i / N, n = i - b * N;
    float v = bias ? bias[n] : 0.f;
    for (int s = 0; s < S; s++) v += red[(s * NB + b) * N + n];
    out[b * ldo + n] = v;
  }
  __syncthreads();
}

// LayerNorm (no bias) of nb rows in smem: one warp per row, round-robin.
__device__ __forceinline__ void layernorm_rows(const float* hs, float* xs,
                                               const float* __restrict__ gamma, int nb, int D) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = w; b < nb; b += kWarps) {
    const float* h = hs + b * D;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += h[c];
    const float mean = warp_sum(s) / D;
    float q = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float d = h[c] - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / D + 1e-5f);
    for (int c = lane; c < D; c += 32) xs[b * D + c] = (h[c] - mean) * rstd * gamma[c];
  }
  __syncthreads();
}

// Argmax over the per-chunk candidates a step left behind, for utterance b.
// Lowest index wins ties (MoonshineTensorView::argmax uses strict '>',
// core/ort-utils/moonshine-tensor-view.cpp:222-236).  Whole warp; result in
// every lane.
__device__ __forceinline__ int resolve_token_warp(const DecoderParams& p, int b, int parity) {
  const int lane = threadIdx.x & 31;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  const float* cv = p.cand_val + (int64_t)parity * p.n_vchunk * p.B;
  const int* ci = p.cand_idx + (int64_t)parity * p.n_vchunk * p.B;
  for (int c = lane; c < p.n_vchunk; c += 32) {
    const float v = cv[(int64_t)c * p.B + b];
    const int i = ci[(int64_t)c * p.B + b];
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (bi == 0x7fffffff) bi = 0;  // nothing compared greater than -inf: index 0
  return bi;
}

// Step prologue for utterance b (whole warp).  Emitted token #step is the
// argmax of step-1's logits; it is always recorded.  The token consumed at
// this step is that token (or the teacher-forced one).  An utterance finishes
// after emitting EOS (id 2) or max_len tokens (moonshine-model.cpp:380,511-516).
__device__ __forceinline__ void step_prologue_warp(const DecoderParams& p, int b, bool writer,
                                                   int& tok_in, bool& finished) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)b * (p.Smax + 1);
  if (p.done[b]) {
    tok_in = 0;
    finished = true;
    return;
  }
  if (p.step == 0) {
    tok_in = p.forced ? p.forced[row] : p.tokens[row];
    finished = p.max_len[b] <= 0;
  } else {
    const int emitted = resolve_token_warp(p, b, (p.step - 1) & 1);
    tok_in = p.forced ? p.forced[row + p.step] : emitted;
    finished = (tok_in == 2) || (p.step >= p.max_len[b]);
    if (writer && lane == 0) {
      p.tokens[row + p.step] = emitted;
      p.n_tokens[b] = p.step + 1;
    }
  }
  if (finished && writer && lane == 0) p.done[b] = 1;
}

// hs[b][:] = hbuf_rd[b0+b][:] + sum_j part[j][b0+b][:] (+ bias), rows of
// finished / out-of-range utterances are zero.  Optionally stores the resolved
// rows to hbuf_wr.
__device__ __forceinline__ void resolve_rows(const DecoderParams& p, const Smem& sm, int NB, int b0,
                                             const float* hrd, float* hwr, const float* part,
                                             int nparts, const float* __restrict__ bias,
                                             bool store) {
  const int D = p.D;
  const int D4 = D >> 2;
  for (int i = threadIdx.x; i < NB * D4; i += kThreads) {
    const int b = i / D4, c4 = i - b * D4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!sm.flags[b]) {
      const int64_t r = (int64_t)(b0 + b) * D4 + c4;
      v = reinterpret_cast<const float4*>(hrd)[r];
      for (int j = 0; j < nparts; j++) {
        const float4 q = reinterpret_cast<const float4*>(part)[(int64_t)j * p.B * D4 + r];
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      if (bias) {
        const float4 q = reinterpret_cast<const float4*>(bias)[c4];
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      if (store) reinterpret_cast<float4*>(hwr)[r] = v;
    }
    reinterpret_cast<float4*>(sm.hs)[b * D4 + c4] = v;
  }
  __syncthreads();
}

__device__ __forceinline__ void load_flags(const DecoderParams& p, const Smem& sm, int NB, int b0) {
  if (threadIdx.x < NB) {
    const int b = b0 + threadIdx.x;
    sm.flags[threadIdx.x] = (b < p.B) ? p.done[b] : 1;
  }
  __syncthreads();
}

__device__ __forceinline__ void store_partial(const DecoderParams& p, const Smem& sm, int NB, int b0,
                                              float* part_slice) {
  const int D4 = p.D >> 2;
  for (int i = threadIdx.x; i < NB * D4; i += kThreads) {
    const int b = i / D4, c4 = i - b * D4;
    if (!sm.flags[b])
      reinterpret_cast<float4*>(part_slice)[(int64_t)(b0 + b) * D4 + c4] =
          reinterpret_cast<const float4*>(sm.hs)[b * D4 + c4];
  }
  __syncthreads();
}

// ------------------------------- phase A ---------------------------------
template <int NB>
__device__ void phase_self(const DecoderParams& p, int l, int item, const Smem& sm,
                           const float* hrd, float* hwr, const float* partC, float* partA) {
  const int D = p.D, hd = p.hd, H = p.H;
  const int h = item % H;
  const int b0 = (item / H) * NB;
  const DecLayerWeights& w = p.layers[l];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SmemLayout L = smem_layout(1, D, hd, p.IC, p.Tpad, p.Smax);  // widths only
  const int actw = L.actw, attw = L.attw;

  if (l == 0) {
    // token resolution + embedding gather
    for (int b = warp; b < NB; b += kWarps) {
      int tok = 0;
      bool fin = true;
      if (b0 + b < p.B) step_prologue_warp(p, b0 + b, h == 0, tok, fin);
      if (lane == 0) {
        sm.flags[b] = fin ? 1 : 0;
        sm.flags[32 + b] = tok;
      }
    }
    __syncthreads();
    const int D4 = D >> 2;
    for (int i = threadIdx.x; i < NB * D4; i += kThreads) {
      const int b = i / D4, c4 = i - b * D4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!sm.flags[b]) {
        v = reinterpret_cast<const float4*>(p.embed)[(int64_t)sm.flags[32 + b] * D4 + c4];
        if (h == 0) reinterpret_cast<float4*>(hwr)[(int64_t)(b0 + b) * D4 + c4] = v;
      }
      reinterpret_cast<float4*>(sm.hs)[b * D4 + c4] = v;
    }
    __syncthreads();
  } else {
    load_flags(p, sm, NB, b0);
    resolve_rows(p, sm, NB, b0, hrd, hwr, partC, p.n_chunk, p.layers[l - 1].b2, h == 0);
  }
  bool any = false;
  for (int b = 0; b < NB; b++) any |= (sm.flags[b] == 0);
  if (!any) { __syncthreads(); return; }

  layernorm_rows(sm.hs, sm.xs, w.ln1, NB, D);
  tile_gemm<NB>(sm.xs, D, D, w.wqkv + (int64_t)h * D * 3 * hd, 3 * hd, nullptr, sm.red, sm.act,
                actw);

  // RoPE (interleaved pairs) on q and k at position `step`
  {
    const int half_rot = p.rot_dim >> 1;
    for (int i = threadIdx.x; i < NB * 2 * half_rot; i += kThreads) {
      const int b = i / (2 * half_rot);
      const int r = i - b * 2 * half_rot;
      const int which = r / half_rot;  // 0 = q, 1 = k
      const int pr = r - which * half_rot;
      const float c = p.rope_cos[(int64_t)p.step * half_rot + pr];
      const float s = p.rope_sin[(int64_t)p.step * half_rot + pr];
      float* v = sm.act + b * actw + which * hd + 2 * pr;
      const float x0 = v[0], x1 = v[1];
      v[0] = x0 * c - x1 * s;
      v[1] = x1 * c + x0 * s;
    }
  }
  __syncthreads();
  // KV append
  for (int i = threadIdx.x; i < NB * hd; i += kThreads) {
    const int b = i / hd, d = i - b * hd;
    if (!sm.flags[b]) {
      const int64_t bh = ((int64_t)l * p.B + (b0 + b)) * H + h;
      p.ks[(bh * hd + d) * p.Smax + p.step] = sm.act[b * actw + hd + d];
      p.vs[(bh * p.Smax + p.step) * hd + d] = sm.act[b * actw + 2 * hd + d];
    }
  }
  // causal self-attention: one warp per utterance; position `step` comes
  // from smem (just computed), earlier positions from the cache.
  {
    const float scale = rsqrtf((float)hd);
    float* sc = sm.sc + warp * p.Smax;
    for (int b = warp; b < NB; b += kWarps) {
      if (sm.flags[b]) continue;
      const int64_t bh = ((int64_t)l * p.B + (b0 + b)) * H + h;
      const float* q = sm.act + b * actw;
      const float* kcur = q + hd;
      const float* vcur = q + 2 * hd;
      const float* kc = p.ks + bh * hd * p.Smax;
      const float* vc = p.vs + bh * p.Smax * hd;
      float mx = -INFINITY;
      for (int t = lane; t <= p.step; t += 32) {
        float s = 0.f;
        if (t == p.step) {
          for (int d = 0; d < hd; d++) s = fmaf(q[d], kcur[d], s);
        } else {
          for (int d = 0; d < hd; d++) s = fmaf(q[d], kc[(int64_t)d * p.Smax + t], s);
        }
        s *= scale;
        sc[t] = s;
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      float sum = 0.f;
      for (int t = lane; t <= p.step; t += 32) {
        const float e = expf(sc[t] - mx);
        sc[t] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      __syncwarp();
      const float inv = 1.0f / sum;
      for (int d = lane; d < hd; d += 32) {
        float o = 0.f;
        for (int t = 0; t < p.step; t++) o = fmaf(sc[t], vc[(int64_t)t * hd + d], o);
        o = fmaf(sc[p.step], vcur[d], o);
        sm.att[b * attw + d] = o * inv;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // O-projection partial of this head
  tile_gemm<NB>(sm.att, attw, hd, w.wo + (int64_t)h * hd * D, D, nullptr, sm.red, sm.hs, D);
  store_partial(p, sm, NB, b0, partA + (int64_t)h * p.B * D);
}

// ------------------------------- phase B ---------------------------------
template <int NB>
__device__ void phase_cross(const DecoderParams& p, int l, int item, const Smem& sm,
                            const float* hrd, float* hwr, const float* partA, float* partB) {
  const int D = p.D, hd = p.hd, H = p.H;
  const int h = item % H;
  const int b0 = (item / H) * NB;
  const DecLayerWeights& w = p.layers[l];
  const SmemLayout L = smem_layout(1, D, hd, p.IC, p.Tpad, p.Smax);
  const int actw = L.actw, attw = L.attw;

  load_flags(p, sm, NB, b0);
  bool any = false;
  for (int b = 0; b < NB; b++) any |= (sm.flags[b] == 0);
  if (!any) { __syncthreads(); return; }
  resolve_rows(p, sm, NB, b0, hrd, hwr, partA, H, nullptr, h == 0);
  layernorm_rows(sm.hs, sm.xs, w.ln2, NB, D);
  tile_gemm<NB>(sm.xs, D, D, w.wqc + (int64_t)h * D * hd, hd, nullptr, sm.red, sm.act, actw);

  // cross attention: each 128-thread half-block takes every other utterance.
  {
    const int half = threadIdx.x >> 7;
    const int ht = threadIdx.x & 127;
    const int hw = ht >> 5, lane = ht & 31;
    float* ps = sm.ps + half * L.Tps;
    float* redh = sm.red + half * 520;  // [4] warp partials + [G][hd] (<= 512) PV scratch
    const float scale = rsqrtf((float)hd);
    const int Tpad = p.Tpad;
    const int tpr = hd >> 2;             // threads per V row (4 halves each)
    const int G = 128 / tpr;             // V rows per pass
    for (int bb = half; bb < NB; bb += 2) {
      // `flags` and loop bounds are uniform across the half-block
      const bool skip = sm.flags[bb] != 0;
      if (!skip) {
        const int b = b0 + bb;
        const int T = p.enc_len[b];
        const int64_t bh = ((int64_t)l * p.B + b) * H + h;
        const __half* kc = p.kc + bh * hd * Tpad;
        const __half* vc = p.vc + bh * Tpad * hd;
        const float* q = sm.act + bb * actw;
        // scores: thread j covers t = 4j .. 4j+3 (and +512 for long clips)
        float lmax = -INFINITY;
        for (int t4 = ht * 4; t4 < Tpad; t4 += 512) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          const __half* kp = kc + t4;
#pragma unroll 4
          for (int d = 0; d < hd; d++) {
            const uint2 u = *reinterpret_cast<const uint2*>(kp + (int64_t)d * Tpad);
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            const float qd = q[d];
            s0 = fmaf(qd, f0.x, s0);
            s1 = fmaf(qd, f0.y, s1);
            s2 = fmaf(qd, f1.x, s2);
            s3 = fmaf(qd, f1.y, s3);
          }
          s0 = (t4 + 0 < T) ? s0 * scale : -INFINITY;
          s1 = (t4 + 1 < T) ? s1 * scale : -INFINITY;
          s2 = (t4 + 2 < T) ? s2 * scale : -INFINITY;
          s3 = (t4 + 3 < T) ? s3 * scale : -INFINITY;
          ps[t4 + 0] = s0; ps[t4 + 1] = s1; ps[t4 + 2] = s2; ps[t4 + 3] = s3;
          lmax = fmaxf(fmaxf(lmax, fmaxf(s0, s1)), fmaxf(s2, s3));
        }
        lmax = warp_max(lmax);
        if (lane == 0) redh[hw] = lmax;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + half));
        const float mx = fmaxf(fmaxf(redh[0], redh[1]), fmaxf(redh[2], redh[3]));
        float lsum = 0.f;
        for (int t = ht; t < Tpad; t += 128) {
          const float e = (t < T) ? expf(ps[t] - mx) : 0.f;
          ps[t] = e;
          lsum += e;
        }
        lsum = warp_sum(lsum);
        asm volatile("bar.sync %0, 128;" ::"r"(1 + half));  // redh max reads done
        if (lane == 0) redh[hw] = lsum;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + half));
        const float inv = 1.0f / (redh[0] + redh[1] + redh[2] + redh[3]);
        // PV: thread (g, dq) accumulates rows t = g, g+G, ... for 4 dims
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const int g = ht / tpr, dq = ht - g * tpr;
        if (g < G) {
          const __half* vp = vc + dq * 4;
#pragma unroll 4
          for (int t = g; t < T; t += G) {
            const uint2 u = *reinterpret_cast<const uint2*>(vp + (int64_t)t * hd);
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            const float pt = ps[t];
            a0 = fmaf(pt, f0.x, a0);
            a1 = fmaf(pt, f0.y, a1);
            a2 = fmaf(pt, f1.x, a2);
            a3 = fmaf(pt, f1.y, a3);
          }
        }
        float* pv = redh + 4;  // [G][hd]
        if (g < G) {
          pv[g * hd + dq * 4 + 0] = a0;
          pv[g * hd + dq * 4 + 1] = a1;
          pv[g * hd + dq * 4 + 2] = a2;
          pv[g * hd + dq * 4 + 3] = a3;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + half));
        if (ht < hd) {
          float o = 0.f;
          for (int gg = 0; gg < G; gg++) o += pv[gg * hd + ht];
          sm.att[bb * attw + ht] = o * inv;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + half));
      }
    }
  }
  __syncthreads();
  tile_gemm<NB>(sm.att, attw, hd, w.woc + (int64_t)h * hd * D, D, nullptr, sm.red, sm.hs, D);
  store_partial(p, sm, NB, b0, partB + (int64_t)h * p.B * D);
}

// ------------------------------- phase C ---------------------------------
template <int NB>
__device__ void phase_mlp(const DecoderParams& p, int l, int item, const Smem& sm,
                          const float* hrd, float* hwr, const float* partB, float* partC) {
  const int D = p.D, IC = p.IC;
  const int c = item % p.n_chunk;
  const int b0 = (item / p.n_chunk) * NB;
  const DecLayerWeights& w = p.layers[l];
  const SmemLayout L = smem_layout(1, D, p.hd, IC, p.Tpad, p.Smax);
  const int actw = L.actw, attw = L.attw;

  load_flags(p, sm, NB, b0);
  bool any = false;
  for (int b = 0; b < NB; b++) any |= (sm.flags[b] == 0);
  if (!any) { __syncthreads(); return; }
  resolve_rows(p, sm, NB, b0, hrd, hwr, partB, p.H, nullptr, c == 0);
  layernorm_rows(sm.hs, sm.xs, w.ln3, NB, D);
  tile_gemm<NB>(sm.xs, D, D, w.w1 + (int64_t)c * D * 2 * IC, 2 * IC, w.b1 + (int64_t)c * 2 * IC,
                sm.red, sm.act, actw);
  for (int i = threadIdx.x; i < NB * IC; i += kThreads) {
    const int b = i / IC, j = i - b * IC;
    const float up = sm.act[b * actw + j];
    const float gate = sm.act[b * actw + IC + j];
    sm.att[b * attw + j] = gate / (1.0f + expf(-gate)) * up;  // silu(gate) * up
  }
  __syncthreads();
  tile_gemm<NB>(sm.att, attw, IC, w.w2 + (int64_t)c * IC * D, D, nullptr, sm.red, sm.hs, D);
  store_partial(p, sm, NB, b0, partC + (int64_t)c * p.B * D);
}

// ------------------------------- phase F ---------------------------------
template <int NB>
__device__ void phase_final_ln(const DecoderParams& p, int item, const Smem& sm, const float* hrd,
                               const float* partC) {
  const int D = p.D;
  const int b0 = item * NB;
  load_flags(p, sm, NB, b0);
  resolve_rows(p, sm, NB, b0, hrd, nullptr, partC, p.n_chunk, p.layers[p.L - 1].b2, false);
  layernorm_rows(sm.hs, sm.xs, p.final_ln, NB, D);
  for (int i = threadIdx.x; i < NB * D; i += kThreads) {
    const int b = i / D;
    if (b0 + b < p.B) p.xfin[(int64_t)(b0 + b) * D + (i - b * D)] = sm.xs[i];
  }
  __syncthreads();
}

// ------------------------------- phase G ---------------------------------
// One vocab entry per thread, kLogitsTile utterances per pass.
__device__ void phase_logits(const DecoderParams& p, int item, const Smem& sm) {
  const int D = p.D, V = p.V;
  const int v = item * p.vchunk + threadIdx.x;
  const bool vok = threadIdx.x < p.vchunk && v < V;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xt = sm.red;                 // [D][kLogitsTile]  (k-major)
  float* wv = sm.argv;                // [kWarps][kLogitsTile] warp argmax values
  int* wi = sm.argi;
  const int parity = p.step & 1;
  for (int b0 = 0; b0 < p.B; b0 += kLogitsTile) {
    const int nb = min(kLogitsTile, p.B - b0);
    for (int i = threadIdx.x; i < D * kLogitsTile; i += kThreads) {
      const int b = i % kLogitsTile, k = i / kLogitsTile;
      xt[i] = (b < nb) ? p.xfin[(int64_t)(b0 + b) * D + k] : 0.f;
    }
    __syncthreads();
    float acc[kLogitsTile];
#pragma unroll
    for (int b = 0; b < kLogitsTile; b++) acc[b] = 0.f;
    if (vok) {
      const float* ep = p.embT + v;
#pragma unroll 4
      for (int k = 0; k < D; k++) {
        const float wk = __ldg(ep + (int64_t)k * V);
        const float4* xr = reinterpret_cast<const float4*>(xt + k * kLogitsTile);
#pragma unroll
        for (int b4 = 0; b4 < kLogitsTile / 4; b4++) {
          const float4 x = xr[b4];
          acc[b4 * 4 + 0] = fmaf(wk, x.x, acc[b4 * 4 + 0]);
          acc[b4 * 4 + 1] = fmaf(wk, x.y, acc[b4 * 4 + 1]);
          acc[b4 * 4 + 2] = fmaf(wk, x.z, acc[b4 * 4 + 2]);
          acc[b4 * 4 + 3] = fmaf(wk, x.w, acc[b4 * 4 + 3]);
        }
      }
      if (p.logits_out) {
#pragma unroll
        for (int b = 0; b < kLogitsTile; b++)
          if (b < nb) p.logits_out[(int64_t)(b0 + b) * V + v] = acc[b];
      }
    }
    // per-utterance argmax over this chunk (lowest index wins ties)
#pragma unroll
    for (int b = 0; b < kLogitsTile; b++) {
      float bv = vok ? acc[b] : -INFINITY;
      int bi = vok ? v : 0x7fffffff;
      if (bv != bv) { bv = -INFINITY; bi = 0x7fffffff; }  // NaN never wins a strict '>'
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { wv[warp * kLogitsTile + b] = bv; wi[warp * kLogitsTile + b] = bi; }
    }
    __syncthreads();
    if (threadIdx.x < nb) {
      const int b = threadIdx.x;
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int w2 = 0; w2 < kWarps; w2++) {
        const float ov = wv[w2 * kLogitsTile + b];
        const int oi = wi[w2 * kLogitsTile + b];
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      p.cand_val[((int64_t)parity * p.n_vchunk + item) * p.B + b0 + b] = bv;
      p.cand_idx[((int64_t)parity * p.n_vchunk + item) * p.B + b0 + b] = bi;
    }
    __syncthreads();
  }
}

template <int NB, int NBM>
__global__ void __launch_bounds__(kThreads, 1) decoder_step_kernel(const __grid_constant__ DecoderParams p) {
  if (*p.n_active == 0) return;  // written at the end of the previous launch
  extern __shared__ __align__(16) float smem_raw[];
  constexpr int NBmax = NB > NBM ? NB : NBM;
  const SmemLayout L = smem_layout(NBmax, p.D, p.hd, p.IC, p.Tpad, p.Smax);
  Smem sm;
  sm.hs = smem_raw + L.hs;
  sm.xs = smem_raw + L.xs;
  sm.act = smem_raw + L.act;
  sm.att = smem_raw + L.att;
  sm.red = smem_raw + L.red;
  sm.ps = smem_raw + L.ps;
  sm.sc = smem_raw + L.sc;
  sm.flags = reinterpret_cast<int*>(smem_raw + L.flags);
  sm.argv = smem_raw + L.argv;
  sm.argi = reinterpret_cast<int*>(smem_raw + L.argi);

  const unsigned G = gridDim.x;
  unsigned nbar = p.barrier[1];  // barriers completed by earlier launches
  const int n_bt = (p.B + NB - 1) / NB;
  const int n_btm = (p.B + NBM - 1) / NBM;
  const int64_t BD = (int64_t)p.B * p.D;
  float* partA = p.part;
  float* partB = p.part + (int64_t)p.H * BD;
  float* partC = p.part + (int64_t)2 * p.H * BD;
  int ph = 0;
  for (int l = 0; l < p.L; l++) {
    {
      const float* hrd = p.hbuf + (int64_t)(ph & 1) * BD;
      float* hwr = p.hbuf + (int64_t)((ph + 1) & 1) * BD;
      for (int it = blockIdx.x; it < n_bt * p.H; it += G)
        phase_self<NB>(p, l, it, sm, hrd, hwr, partC, partA);
      grid_barrier(p.barrier, (++nbar) * G);
      ph++;
    }
    {
      const float* hrd = p.hbuf + (int64_t)(ph & 1) * BD;
      float* hwr = p.hbuf + (int64_t)((ph + 1) & 1) * BD;
      for (int it = blockIdx.x; it < n_bt * p.H; it += G)
        phase_cross<NB>(p, l, it, sm, hrd, hwr, partA, partB);
      grid_barrier(p.barrier, (++nbar) * G);
      ph++;
    }
    {
      const float* hrd = p.hbuf + (int64_t)(ph & 1) * BD;
      float* hwr = p.hbuf + (int64_t)((ph + 1) & 1) * BD;
      for (int it = blockIdx.x; it < n_btm * p.n_chunk; it += G)
        phase_mlp<NBM>(p, l, it, sm, hrd, hwr, partB, partC);
      grid_barrier(p.barrier, (++nbar) * G);
      ph++;
    }
  }
  {
    const float* hrd = p.hbuf + (int64_t)(ph & 1) * BD;
    for (int it = blockIdx.x; it < n_btm; it += G) phase_final_ln<NBM>(p, it, sm, hrd, partC);
    grid_barrier(p.barrier, (++nbar) * G);
  }
  for (int it = blockIdx.x; it < p.n_vchunk; it += G) phase_logits(p, it, sm);
  grid_barrier(p.barrier, (++nbar) * G);
  if (blockIdx.x == 0) {
    // count utterances still decoding -> early exit of later launches
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int c = 0;
    for (int b = threadIdx.x; b < p.B; b += kThreads) c += p.done[b] ? 0 : 1;
    if (c) atomicAdd(&cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) {
      *p.n_active = cnt;
      p.barrier[1] = nbar;
    }
  }
}

__global__ void decoder_finalize_kernel(const __grid_constant__ DecoderParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= p.B) return;
  const int b = warp;
  if (p.done[b]) return;
  const int emitted = resolve_token_warp(p, b, (p.step - 1) & 1);
  if (lane == 0) {
    p.tokens[(int64_t)b * (p.Smax + 1) + p.step] = emitted;
    p.n_tokens[b] = p.step + 1;
    p.done[b] = 1;
  }
}

template <int NB, int NBM>
void launch_variant(const DecoderParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = decoder_step_kernel<NB, NBM>;
  static SmemAttrCache cache;  // one per template instantiation
  if (cache.needs(smem == 0 ? 1 : smem))
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {const_cast<DecoderParams*>(&p)};
  CUDA_CHECK(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kThreads), args, smem, stream));
}

}  // namespace

void decoder_tiles_for_batch(int B, int& nb, int& nbm) {
  int t = (B + 15) / 16;
  nb = 1;
  while (nb < t && nb < 16) nb <<= 1;
  nbm = nb * 2 > 16 ? 16 : nb * 2;
  if (B == 1) nbm = 1;
}

size_t decoder_step_smem_bytes(const DecoderParams& p) {
  int nb, nbm;
  decoder_tiles_for_batch(p.B, nb, nbm);
  const SmemLayout L = smem_layout(nb > nbm ? nb : nbm, p.D, p.hd, p.IC, p.Tpad, p.Smax);
  return (size_t)L.total * sizeof(float);
}

void launch_decoder_step(const DecoderParams& p, int grid, cudaStream_t stream) {
  int nb, nbm;
  decoder_tiles_for_batch(p.B, nb, nbm);
  const size_t smem = decoder_step_smem_bytes(p);
  if (nb == 1 && nbm == 1) launch_variant<1, 1>(p, grid, smem, stream);
  else if (nb == 1) launch_variant<1, 2>(p, grid, smem, stream);
  else if (nb == 2) launch_variant<2, 4>(p, grid, smem, stream);
  else if (nb == 4) launch_variant<4, 8>(p, grid, smem, stream);
  else if (nb == 8) launch_variant<8, 16>(p, grid, smem, stream);
  else launch_variant<16, 16>(p, grid, smem, stream);
}

void launch_decoder_finalize(const DecoderParams& p, cudaStream_t stream) {
  const int warps_per_block = 4;
  const int blocks = (p.B + warps_per_block - 1) / warps_per_block;
  decoder_finalize_kernel<<<blocks, warps_per_block * 32, 0, stream>>>(p);
}

}  // namespace msb
