// Persistent decoder-step kernel, v2: operand streaming.
//
// Same phase structure and arithmetic order as decoder_step.cu (3 grid barriers
// per layer, deterministic partial sums), but every operand that does not
// depend on this step's activations -- all weights, the fp16 cross K/V cache,
// the self K/V cache up to position step-1 -- is brought into shared memory
// by a dedicated PRODUCER WARP through a ring of TMA bulk copies
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx).  The producer
// walks the same (phase, item, operand, chunk) sequence as the 8 consumer warps
// but never waits for the grid barriers, so the HBM/L2 stream runs ahead of the
// dependency chain and the consumers only ever touch shared memory, plus the
// small activation / partial-sum exchanges through L2.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"
#include "kernels.h"

namespace msb {

namespace {

constexpr int kConsumers = 256;
constexpr int kThreads2 = kConsumers + 32;
constexpr int kWarpsC = kConsumers / 32;
constexpr int kLogitsTile = 32;
constexpr int kStageBytes = 32768;
constexpr long long kSpinLimit = 4000000000LL;  // ~2 s of SM cycles: trap instead of hanging

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
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Consumer-side grid barrier (the producer warp never takes part).
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  csync();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    while ((int)(ld_acquire(bar) - target) < 0) {
      if (clock64() - t0 > kSpinLimit) __trap();
    }
    __threadfence();
  }
  csync();
}

// ---- mbarrier / bulk-copy primitives (PTX) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- the operand ring ----
struct Ring {
  uint64_t* full;
  uint64_t* empty;
  char* data;
  int ns;
  int idx;  // chunks consumed / produced so far (thread-local, uniform)
  __device__ __forceinline__ int stage() const { return idx % ns; }
  __device__ __forceinline__ uint32_t parity() const { return (uint32_t)((idx / ns) & 1); }
  // consumers (all 256 threads call both)
  __device__ __forceinline__ const char* acquire() {
    mbar_wait(&full[stage()], parity());
    return data + (size_t)stage() * kStageBytes;
  }
  // every consumer WARP releases the stage once its lanes are done reading it
  // (empty barriers count kWarpsC arrivals): no CTA-wide sync per chunk.
  __device__ __forceinline__ void release() {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stage()]);
    idx++;
  }
  // producer (one lane)
  __device__ __forceinline__ void produce(const void* src, uint32_t bytes) {
    mbar_wait(&empty[stage()], parity() ^ 1u);
    mbar_expect_tx(&full[stage()], bytes);
    bulk_g2s(data + (size_t)stage() * kStageBytes, src, bytes, &full[stage()]);
    idx++;
  }
};

// rows of a [K][N] fp32 k-major block per ring chunk
__device__ __forceinline__ int rows_per_chunk_f32(int K, int N) {
  int r = kStageBytes / (N * 4);
  return r < 1 ? 1 : (r > K ? K : r);
}
// rows of a fp16 [rows][cols] block per chunk; even row count keeps 16-byte granularity
__device__ __forceinline__ int rows_per_chunk_f16(int rows, int cols) {
  int r = (kStageBytes / (cols * 2)) & ~1;
  if (r < 2) r = 2;
  return r > rows ? rows : r;
}

struct SmemLayout2 {
  int hs, xs, act, att, red, ps, sc, flags, argv, argi, active, bars, ring, rope, bias;  // byte offsets
  int actw, attw, total, ns, xpitch;
};

__host__ __device__ inline SmemLayout2 smem_layout2(int NBmax, int B, int D, int hd, int IC, int Tpad,
                                                    int Smax, int smem_limit) {
  SmemLayout2 L;
  L.actw = max(3 * hd, 2 * IC);
  L.attw = max(hd, IC);
  int o = 0;
  auto take = [&](int bytes) { int r = o; o += (bytes + 15) / 16 * 16; return r; };
  L.hs = take(NBmax * D * 4);
  L.xs = take(NBmax * D * 4);
  L.act = take(NBmax * L.actw * 4);
  L.att = take(NBmax * L.attw * 4);
  // split-K scratch only for small tiles (NB <= 4); larger tiles use the row-split mapping
  int red = (NBmax <= 4 ? 1024 * NBmax : 0) * 4;
  if (red < (16 + 1024) * 4) red = (16 + 1024) * 4;
  L.red = take(red);
  // the logits phase aliases [0, xg_bytes) with its 32 x xpitch x-tile (xpitch % 32 == 8)
  L.xpitch = D + ((8 - (D % 32)) + 32) % 32;
  const int xg_bytes = kLogitsTile * L.xpitch * 4;
  if (o < xg_bytes) o = (xg_bytes + 15) / 16 * 16;
  L.ps = take(Tpad * 4);
  L.sc = take(kWarpsC * (Smax + 4) * 4);
  L.flags = take(64 * 4);
  L.argv = take(kWarpsC * kLogitsTile * 4);
  L.argi = take(kWarpsC * kLogitsTile * 4);
  L.rope = take(128 * 4);
  L.bias = take(2 * IC * 4);
  L.active = take(B);
  L.bars = take(2 * 16 * 8);
  o = (o + 127) / 128 * 128;
  L.ring = o;
  int ns = (smem_limit - o) / kStageBytes;
  if (ns > 16) ns = 16;
  L.ns = ns;
  L.total = o + ns * kStageBytes;
  return L;
}

constexpr int kProfSlots = 512;
struct Ctx {
  const float* rope;         // smem: cos[0..64) | sin[64..128) of this step's position
  float* bias;               // smem: staged FC1 bias chunk (2 * IC floats)
  float* xg;                 // smem: [32][xpitch] final-LN rows for the logits phase (aliases hs..red)
  int xpitch;
  unsigned long long* prof;  // optional [grid][kProfSlots] globaltimer stamps (thread 0)
  int prof_n;
  float *hs, *xs, *act, *att, *red, *ps, *sc, *argv;
  int *flags, *argi;
  const unsigned char* active;  // [B] 1 = utterance still decoding at kernel start
  int actw, attw;
};

__device__ __forceinline__ void prof_mark(Ctx& c, int tag) {
  if (c.prof != nullptr && threadIdx.x == 0 && c.prof_n < kProfSlots) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    c.prof[(size_t)blockIdx.x * kProfSlots + c.prof_n] = (t << 8) | (unsigned)tag;
    c.prof_n++;
  }
}

// ------------------------------------------------------------------------
// GEMMs over ring-staged weights.  Weight block Wt[K][N] (k-major, fp32) arrives
// in chunks of whole rows.  out[b][n] = sum_k x[b][k] * Wt[k][n] (+ bias[n]).
// ------------------------------------------------------------------------
// (a) split-K mapping for small row tiles: threads = (k-slice s, float4 n4).
template <int NB>
__device__ __forceinline__ void gemm_ring_splitk(Ring& ring, const float* x, int ldx, int K, int N,
                                                 const float* __restrict__ bias, float* red,
                                                 float* out, int ldo) {
  const int N4 = N >> 2;
  int S = kConsumers / N4;
  if (S > K) S = K;
  const int t = threadIdx.x;
  const int n4 = t % N4, s = t / N4;
  const bool on = s < S;
  float acc[NB][4];
#pragma unroll
  for (int b = 0; b < NB; b++) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.f;
  const int rpc = rows_per_chunk_f32(K, N);
  for (int k0 = 0; k0 < K; k0 += rpc) {
    const int rows = min(rpc, K - k0);
    const float4* W = reinterpret_cast<const float4*>(ring.acquire());
    if (on) {
#pragma unroll 4
      for (int r = s; r < rows; r += S) {
        const float4 w = W[r * N4 + n4];
        const int k = k0 + r;
#pragma unroll
        for (int b = 0; b < NB; b++) {
          const float xv = x[b * ldx + k];
          acc[b][0] = fmaf(xv, w.x, acc[b][0]);
          acc[b][1] = fmaf(xv, w.y, acc[b][1]);
          acc[b][2] = fmaf(xv, w.z, acc[b][2]);
          acc[b][3] = fmaf(xv, w.w, acc[b][3]);
        }
      }
    }
    ring.release();
  }
  if (on) {
#pragma unroll
    for (int b = 0; b < NB; b++)
      *reinterpret_cast<float4*>(&red[(s * NB + b) * N + n4 * 4]) =
          make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
  }
  csync();
  for (int i = threadIdx.x; i < NB * N; i += kConsumers) {
    const int b = i / N, n = i - b * N;
    float v = bias ? bias[n] : 0.f;
    for (int s2 = 0; s2 < S; s2++) v += red[(s2 * NB + b) * N + n];
    out[b * ldo + n] = v;
  }
  csync();
}

// (b) row-split mapping for larger tiles: threads = (row group g, float4 n4);
// each thread owns up to MAXR rows of the tile for its 4 features, full K.
template <int NB, int MAXR>
__device__ __forceinline__ void gemm_ring_rows(Ring& ring, const float* x, int ldx, int K, int N,
                                               const float* __restrict__ bias, float* out, int ldo) {
  const int N4 = N >> 2;
  const int G = kConsumers / N4;
  const int t = threadIdx.x;
  const int n4 = t % N4, g = t / N4;
  const bool on = g < G && g < NB;
  float acc[MAXR][4];
#pragma unroll
  for (int r = 0; r < MAXR; r++) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  const int rpc = rows_per_chunk_f32(K, N);
  for (int k0 = 0; k0 < K; k0 += rpc) {
    const int rows = min(rpc, K - k0);
    const float4* W = reinterpret_cast<const float4*>(ring.acquire());
    if (on) {
#pragma unroll 2
      for (int r = 0; r < rows; r++) {
        const float4 w = W[r * N4 + n4];
        const int k = k0 + r;
#pragma unroll
        for (int j = 0; j < MAXR; j++) {
          const int b = g + j * G;
          if (b < NB) {
            const float xv = x[b * ldx + k];
            acc[j][0] = fmaf(xv, w.x, acc[j][0]);
            acc[j][1] = fmaf(xv, w.y, acc[j][1]);
            acc[j][2] = fmaf(xv, w.z, acc[j][2]);
            acc[j][3] = fmaf(xv, w.w, acc[j][3]);
          }
        }
      }
    }
    ring.release();
  }
  if (on) {
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = reinterpret_cast<const float4*>(bias)[n4];
#pragma unroll
    for (int j = 0; j < MAXR; j++) {
      const int b = g + j * G;
      if (b < NB)
        *reinterpret_cast<float4*>(&out[b * ldo + n4 * 4]) =
            make_float4(acc[j][0] + bv.x, acc[j][1] + bv.y, acc[j][2] + bv.z, acc[j][3] + bv.w);
    }
  }
  csync();
}

template <int NB>
__device__ __forceinline__ void gemm_ring(Ring& ring, const Ctx& c, const float* x, int ldx, int K, int N,
                                          const float* __restrict__ bias, float* out, int ldo) {
  if constexpr (NB <= 4) {
    gemm_ring_splitk<NB>(ring, x, ldx, K, N, bias, c.red, out, ldo);
  } else {
    // rows per thread = ceil(NB / G), G = 256 / (N/4) >= 2 for every N <= 512
    const int G = kConsumers / (N >> 2);
    if (NB <= G) gemm_ring_rows<NB, 1>(ring, x, ldx, K, N, bias, out, ldo);
    else if (NB <= 2 * G) gemm_ring_rows<NB, 2>(ring, x, ldx, K, N, bias, out, ldo);
    else if (NB <= 4 * G) gemm_ring_rows<NB, 4>(ring, x, ldx, K, N, bias, out, ldo);
    else gemm_ring_rows<NB, (NB + 1) / 2>(ring, x, ldx, K, N, bias, out, ldo);
  }
}

// producer side of a [K][N] fp32 block
__device__ __forceinline__ void produce_block_f32(Ring& ring, const float* Wt, int K, int N) {
  const int rpc = rows_per_chunk_f32(K, N);
  for (int k0 = 0; k0 < K; k0 += rpc) {
    const int rows = min(rpc, K - k0);
    ring.produce(Wt + (size_t)k0 * N, (uint32_t)rows * N * 4);
  }
}
__device__ __forceinline__ void produce_block_f16(Ring& ring, const __half* M, int rows, int cols) {
  const int rpc = rows_per_chunk_f16(rows, cols);
  for (int r0 = 0; r0 < rows; r0 += rpc) {
    const int n = min(rpc, rows - r0);
    ring.produce(M + (size_t)r0 * cols, (uint32_t)n * cols * 2);
  }
}

// LayerNorm WITHOUT the affine weight: gamma is folded into the rows of the
// following weight block at load time (Model::build_weights), so the dependency
// chain holds no global load here.
__device__ __forceinline__ void layernorm_rows(const float* hs, float* xs, int nb, int D) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = w; b < nb; b += kWarpsC) {
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
    for (int c = lane; c < D; c += 32) xs[b * D + c] = (h[c] - mean) * rstd;
  }
  csync();
}

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
  if (bi == 0x7fffffff) bi = 0;
  return bi;
}

// Step prologue for an utterance that was active at kernel start (whole warp).
// Records the token emitted by the previous step; an utterance that finishes
// now (EOS consumed / max_len reached) is flagged done for LATER launches but
// still flows through this launch (its results are ignored), so that the
// producer warp and the consumers agree on the work list without talking.
__device__ __forceinline__ int step_prologue_warp(const DecoderParams& p, int b, bool writer) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)b * (p.Smax + 1);
  int tok_in;
  bool finished;
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
  if (tok_in < 0 || tok_in >= p.V) tok_in = 0;
  return tok_in;
}

__device__ __forceinline__ void load_flags(const DecoderParams& p, const Ctx& c, int NB, int b0,
                                           bool with_enc_len = false) {
  if (threadIdx.x < NB) {
    const int b = b0 + threadIdx.x;
    c.flags[threadIdx.x] = (b < p.B) ? (c.active[b] ? 0 : 1) : 1;
    if (with_enc_len) c.flags[32 + threadIdx.x] = (b < p.B) ? p.enc_len[b] : 0;
  }
  csync();
}

__device__ __forceinline__ bool tile_active(const DecoderParams& p, const unsigned char* active, int NB,
                                            int b0) {
  bool any = false;
  for (int b = 0; b < NB; b++) any |= (b0 + b < p.B) && active[b0 + b];
  return any;
}

__device__ __forceinline__ void resolve_rows(const DecoderParams& p, const Ctx& c, int NB, int b0,
                                             const float* hrd, float* hwr, const float* part,
                                             int nparts, const float* __restrict__ bias, bool store) {
  const int D4 = p.D >> 2;
  for (int i = threadIdx.x; i < NB * D4; i += kConsumers) {
    const int b = i / D4, c4 = i - b * D4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!c.flags[b]) {
      const int64_t r = (int64_t)(b0 + b) * D4 + c4;
      v = reinterpret_cast<const float4*>(hrd)[r];
      const float4* pp = reinterpret_cast<const float4*>(part) + r;
      const int64_t pstride = (int64_t)p.B * D4;
      int j = 0;
      for (; j + 8 <= nparts; j += 8) {  // 8 independent loads in flight, summed in order
        float4 q[8];
#pragma unroll
        for (int u = 0; u < 8; u++) q[u] = pp[(int64_t)(j + u) * pstride];
#pragma unroll
        for (int u = 0; u < 8; u++) { v.x += q[u].x; v.y += q[u].y; v.z += q[u].z; v.w += q[u].w; }
      }
      for (; j < nparts; j++) {
        const float4 q = pp[(int64_t)j * pstride];
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      if (bias) {
        const float4 q = reinterpret_cast<const float4*>(bias)[c4];
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      if (store) reinterpret_cast<float4*>(hwr)[r] = v;
    }
    reinterpret_cast<float4*>(c.hs)[b * D4 + c4] = v;
  }
  csync();
}

__device__ __forceinline__ void store_partial(const DecoderParams& p, const Ctx& c, int NB, int b0,
                                              float* part_slice) {
  const int D4 = p.D >> 2;
  for (int i = threadIdx.x; i < NB * D4; i += kConsumers) {
    const int b = i / D4, c4 = i - b * D4;
    if (!c.flags[b])
      reinterpret_cast<float4*>(part_slice)[(int64_t)(b0 + b) * D4 + c4] =
          reinterpret_cast<const float4*>(c.hs)[b * D4 + c4];
  }
  csync();
}

// ============================== phase A =================================
__device__ __forceinline__ void produce_self(const DecoderParams& p, int l, int item, int NB, Ring& ring,
                                             const unsigned char* active) {
  const int H = p.H, hd = p.hd, D = p.D;
  const int h = item % H, b0 = (item / H) * NB;
  if (!tile_active(p, active, NB, b0)) return;
  const DecLayerWeights& w = p.layers[l];
  produce_block_f32(ring, w.wqkv + (int64_t)h * D * 3 * hd, D, 3 * hd);
  if (p.step > 0) {
    for (int b = 0; b < NB; b++) {
      if (b0 + b >= p.B || !active[b0 + b]) continue;
      const int64_t bh = ((int64_t)l * p.B + (b0 + b)) * H + h;
      produce_block_f32(ring, p.ks + bh * hd * p.Smax, hd, p.Smax);   // K^T [hd][Smax]
      produce_block_f32(ring, p.vs + bh * p.Smax * hd, p.step, hd);   // V rows [0, step)
    }
  }
  produce_block_f32(ring, w.wo + (int64_t)h * hd * D, hd, D);
}

template <int NB>
__device__ void phase_self(const DecoderParams& p, int l, int item, Ctx& c, Ring& ring,
                           const float* hrd, float* hwr, const float* partC, float* partA) {
  const int D = p.D, hd = p.hd, H = p.H;
  const int h = item % H;
  const int b0 = (item / H) * NB;
  if (!tile_active(p, c.active, NB, b0)) return;
  const DecLayerWeights& w = p.layers[l];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int actw = c.actw, attw = c.attw;

  load_flags(p, c, NB, b0);
  if (l == 0) {
    for (int b = warp; b < NB; b += kWarpsC) {
      int tok = 0;
      if (!c.flags[b]) tok = step_prologue_warp(p, b0 + b, h == 0);
      if (lane == 0) c.flags[32 + b] = tok;
    }
    csync();
    const int D4 = D >> 2;
    for (int i = threadIdx.x; i < NB * D4; i += kConsumers) {
      const int b = i / D4, c4 = i - b * D4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!c.flags[b]) {
        v = reinterpret_cast<const float4*>(p.embed)[(int64_t)c.flags[32 + b] * D4 + c4];
        if (h == 0) reinterpret_cast<float4*>(hwr)[(int64_t)(b0 + b) * D4 + c4] = v;
      }
      reinterpret_cast<float4*>(c.hs)[b * D4 + c4] = v;
    }
    csync();
  } else {
    resolve_rows(p, c, NB, b0, hrd, hwr, partC, p.n_chunk, p.layers[l - 1].b2, h == 0);
  }
  prof_mark(c, 1);
  layernorm_rows(c.hs, c.xs, NB, D);
  prof_mark(c, 2);
  gemm_ring<NB>(ring, c, c.xs, D, D, 3 * hd, nullptr, c.act, actw);
  prof_mark(c, 3);

  {  // RoPE (interleaved pairs) on q and k at position `step`
    const int half_rot = p.rot_dim >> 1;
    for (int i = threadIdx.x; i < NB * 2 * half_rot; i += kConsumers) {
      const int b = i / (2 * half_rot);
      const int r = i - b * 2 * half_rot;
      const int which = r / half_rot;
      const int pr = r - which * half_rot;
      const float cs = c.rope[pr];
      const float sn = c.rope[64 + pr];
      float* v = c.act + b * actw + which * hd + 2 * pr;
      const float x0 = v[0], x1 = v[1];
      v[0] = x0 * cs - x1 * sn;
      v[1] = x1 * cs + x0 * sn;
    }
  }
  csync();
  for (int i = threadIdx.x; i < NB * hd; i += kConsumers) {  // KV append
    const int b = i / hd, d = i - b * hd;
    if (!c.flags[b]) {
      const int64_t bh = ((int64_t)l * p.B + (b0 + b)) * H + h;
      p.ks[(bh * hd + d) * p.Smax + p.step] = c.act[b * actw + hd + d];
      p.vs[(bh * p.Smax + p.step) * hd + d] = c.act[b * actw + 2 * hd + d];
    }
  }
  // causal self-attention: ONE WARP per utterance (round-robin), warp-level syncs only.
  // Cache chunks arrive in utterance order; every warp acquires / releases every chunk (the
  // ring is CTA-wide) but only the owner warp computes on it, so utterances overlap.
  const float scale = rsqrtf((float)hd);
  {
    float* sc = c.sc + warp * (p.Smax + 4);
    int owner = 0;
    for (int b = 0; b < NB; b++) {
      if (c.flags[b]) continue;  // uniform
      const bool mine = (owner == warp);
      owner = (owner + 1) & (kWarpsC - 1);
      const float* q = c.act + b * actw;
      const float* kcur = q + hd;
      const float* vcur = q + 2 * hd;
      // scores: lane owns key positions t = lane, lane+32, ...
      if (mine)
        for (int t = lane; t <= p.step; t += 32) sc[t] = 0.f;
      if (p.step > 0) {
        const int rpc = rows_per_chunk_f32(hd, p.Smax);
        for (int d0 = 0; d0 < hd; d0 += rpc) {
          const int nd = min(rpc, hd - d0);
          const float* Kc = reinterpret_cast<const float*>(ring.acquire());
          if (mine) {
            for (int t = lane; t < p.step; t += 32) {
              float s = sc[t];
              for (int d = 0; d < nd; d++) s = fmaf(q[d0 + d], Kc[d * p.Smax + t], s);
              sc[t] = s;
            }
          }
          ring.release();
        }
      }
      float inv = 0.f;
      if (mine) {
        float s = 0.f;
        for (int d = lane; d < hd; d += 32) s = fmaf(q[d], kcur[d], s);
        s = warp_sum(s);
        if (lane == 0) sc[p.step] = s;
        __syncwarp();
        float mx = -INFINITY;
        for (int t = lane; t <= p.step; t += 32) mx = fmaxf(mx, sc[t] * scale);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int t = lane; t <= p.step; t += 32) {
          const float e = expf(sc[t] * scale - mx);
          sc[t] = e;
          sum += e;
        }
        inv = 1.0f / warp_sum(sum);
        __syncwarp();
      }
      float o0 = 0.f, o1 = 0.f;  // lane owns dims d = lane, lane + 32 (hd <= 64)
      if (p.step > 0) {
        const int rpc = rows_per_chunk_f32(p.step, hd);
        for (int r0 = 0; r0 < p.step; r0 += rpc) {
          const int nr = min(rpc, p.step - r0);
          const float* Vc = reinterpret_cast<const float*>(ring.acquire());
          if (mine) {
            for (int t = 0; t < nr; t++) {
              const float pt = sc[r0 + t];
              if (lane < hd) o0 = fmaf(pt, Vc[t * hd + lane], o0);
              if (lane + 32 < hd) o1 = fmaf(pt, Vc[t * hd + lane + 32], o1);
            }
          }
          ring.release();
        }
      }
      if (mine) {
        const float pl = sc[p.step];
        if (lane < hd) c.att[b * attw + lane] = fmaf(pl, vcur[lane], o0) * inv;
        if (lane + 32 < hd) c.att[b * attw + lane + 32] = fmaf(pl, vcur[lane + 32], o1) * inv;
      }
    }
  }
  csync();
  prof_mark(c, 4);
  gemm_ring<NB>(ring, c, c.att, attw, hd, D, nullptr, c.hs, D);
  prof_mark(c, 5);
  store_partial(p, c, NB, b0, partA + (int64_t)h * p.B * D);
  prof_mark(c, 6);
}

// ============================== phase B =================================
__device__ __forceinline__ void produce_cross(const DecoderParams& p, int l, int item, int NB, Ring& ring,
                                              const unsigned char* active) {
  const int H = p.H, hd = p.hd, D = p.D;
  const int h = item % H, b0 = (item / H) * NB;
  if (!tile_active(p, active, NB, b0)) return;
  const DecLayerWeights& w = p.layers[l];
  produce_block_f32(ring, w.wqc + (int64_t)h * D * hd, D, hd);
  for (int b = 0; b < NB; b++) {
    if (b0 + b >= p.B || !active[b0 + b]) continue;
    const int64_t bh = ((int64_t)l * p.B + (b0 + b)) * H + h;
    produce_block_f16(ring, p.kc + bh * hd * p.Tpad, hd, p.Tpad);   // K^T [hd][Tpad]
    produce_block_f16(ring, p.vc + bh * p.Tpad * hd, p.Tpad, hd);   // V   [Tpad][hd]
  }
  produce_block_f32(ring, w.woc + (int64_t)h * hd * D, hd, D);
}

template <int NB>
__device__ void phase_cross(const DecoderParams& p, int l, int item, Ctx& c, Ring& ring,
                            const float* hrd, float* hwr, const float* partA, float* partB) {
  const int D = p.D, hd = p.hd, H = p.H;
  const int h = item % H;
  const int b0 = (item / H) * NB;
  if (!tile_active(p, c.active, NB, b0)) return;
  const DecLayerWeights& w = p.layers[l];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int actw = c.actw, attw = c.attw;

  load_flags(p, c, NB, b0, true);
  resolve_rows(p, c, NB, b0, hrd, hwr, partA, H, nullptr, h == 0);
  prof_mark(c, 11);
  layernorm_rows(c.hs, c.xs, NB, D);
  prof_mark(c, 12);
  gemm_ring<NB>(ring, c, c.xs, D, D, hd, nullptr, c.act, actw);
  prof_mark(c, 13);

  const float scale = rsqrtf((float)hd);
  const int Tpad = p.Tpad;
  const int tpr = hd >> 2;           // threads per V row (4 halves = 8 bytes each)
  const int G = kConsumers / tpr;    // V rows per pass
  float* ps = c.ps;
  float* red_max = c.red;            // [8] warp maxima
  float* red_sum = c.red + 8;        // [8] warp sums
  float* pv = c.red + 16;            // [G][hd] PV partials (G * hd <= 1024)
  for (int b = 0; b < NB; b++) {
    if (c.flags[b]) continue;  // uniform
    const int T = c.flags[32 + b];   // encoder length, staged by load_flags
    const float* q = c.act + b * actw;
    // ---- scores over K^T chunks (rows = head dims); thread j owns t = 4j .. 4j+3 ----
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const int t4 = threadIdx.x * 4;
    {
      const int rpc = rows_per_chunk_f16(hd, Tpad);
      for (int d0 = 0; d0 < hd; d0 += rpc) {
        const int nd = min(rpc, hd - d0);
        const __half* Kc = reinterpret_cast<const __half*>(ring.acquire());
        if (t4 < Tpad) {
#pragma unroll 4
          for (int d = 0; d < nd; d++) {
            const uint2 u = *reinterpret_cast<const uint2*>(Kc + d * Tpad + t4);
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            const float qd = q[d0 + d];
            s0 = fmaf(qd, f0.x, s0);
            s1 = fmaf(qd, f0.y, s1);
            s2 = fmaf(qd, f1.x, s2);
            s3 = fmaf(qd, f1.y, s3);
          }
        }
        ring.release();
      }
    }
    float lmax = -INFINITY;
    if (t4 < Tpad) {
      s0 = (t4 + 0 < T) ? s0 * scale : -INFINITY;
      s1 = (t4 + 1 < T) ? s1 * scale : -INFINITY;
      s2 = (t4 + 2 < T) ? s2 * scale : -INFINITY;
      s3 = (t4 + 3 < T) ? s3 * scale : -INFINITY;
      lmax = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
    }
    lmax = warp_max(lmax);
    if (lane == 0) red_max[warp] = lmax;
    csync();                                   // (1) maxima visible; previous utterance fully done
    float mx = red_max[0];
#pragma unroll
    for (int i = 1; i < kWarpsC; i++) mx = fmaxf(mx, red_max[i]);
    float lsum = 0.f;
    if (t4 < Tpad) {
      s0 = (t4 + 0 < T) ? expf(s0 - mx) : 0.f;
      s1 = (t4 + 1 < T) ? expf(s1 - mx) : 0.f;
      s2 = (t4 + 2 < T) ? expf(s2 - mx) : 0.f;
      s3 = (t4 + 3 < T) ? expf(s3 - mx) : 0.f;
      *reinterpret_cast<float4*>(&ps[t4]) = make_float4(s0, s1, s2, s3);
      lsum = (s0 + s1) + (s2 + s3);
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red_sum[warp] = lsum;
    csync();                                   // (2) probabilities and sums visible
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < kWarpsC; i++) tot += red_sum[i];
    const float inv = 1.0f / tot;
    // ---- PV over V chunks (rows = time); thread (g, dq) owns 4 dims of rows g, g+G, ... ----
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int g = threadIdx.x / tpr, dq = threadIdx.x - g * tpr;
    {
      const int rpc = rows_per_chunk_f16(Tpad, hd);
      for (int r0 = 0; r0 < Tpad; r0 += rpc) {
        const int nr = min(rpc, Tpad - r0);
        const __half* Vc = reinterpret_cast<const __half*>(ring.acquire());
        if (g < G) {
          const int tend = min(nr, T - r0);
#pragma unroll 4
          for (int t = g; t < tend; t += G) {
            const uint2 u = *reinterpret_cast<const uint2*>(Vc + t * hd + dq * 4);
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            const float pt = ps[r0 + t];
            a0 = fmaf(pt, f0.x, a0);
            a1 = fmaf(pt, f0.y, a1);
            a2 = fmaf(pt, f1.x, a2);
            a3 = fmaf(pt, f1.y, a3);
          }
        }
        ring.release();
      }
    }
    if (g < G) *reinterpret_cast<float4*>(&pv[g * hd + dq * 4]) = make_float4(a0, a1, a2, a3);
    csync();                                   // (3) PV partials visible
    if (threadIdx.x < hd) {
      float o = 0.f;
      for (int gg = 0; gg < G; gg++) o += pv[gg * hd + threadIdx.x];
      c.att[b * attw + threadIdx.x] = o * inv;
    }
    // no sync here: the next utterance only overwrites red_max before its csync (1), ps after
    // it and pv after its csync (2) -- by then every thread has left this reduction.
  }
  csync();
  prof_mark(c, 14);
  gemm_ring<NB>(ring, c, c.att, attw, hd, D, nullptr, c.hs, D);
  prof_mark(c, 15);
  store_partial(p, c, NB, b0, partB + (int64_t)h * p.B * D);
  prof_mark(c, 16);
}

// ============================== phase C =================================
__device__ __forceinline__ void produce_mlp(const DecoderParams& p, int l, int item, int NB, Ring& ring,
                                            const unsigned char* active) {
  const int D = p.D, IC = p.IC;
  const int ch = item % p.n_chunk, b0 = (item / p.n_chunk) * NB;
  if (!tile_active(p, active, NB, b0)) return;
  const DecLayerWeights& w = p.layers[l];
  produce_block_f32(ring, w.w1 + (int64_t)ch * D * 2 * IC, D, 2 * IC);
  produce_block_f32(ring, w.w2 + (int64_t)ch * IC * D, IC, D);
}

template <int NB>
__device__ void phase_mlp(const DecoderParams& p, int l, int item, Ctx& c, Ring& ring,
                          const float* hrd, float* hwr, const float* partB, float* partC) {
  const int D = p.D, IC = p.IC;
  const int ch = item % p.n_chunk;
  const int b0 = (item / p.n_chunk) * NB;
  if (!tile_active(p, c.active, NB, b0)) return;
  const DecLayerWeights& w = p.layers[l];
  const int actw = c.actw, attw = c.attw;
  for (int i = threadIdx.x; i < 2 * IC; i += kConsumers) c.bias[i] = w.b1[(int64_t)ch * 2 * IC + i];
  load_flags(p, c, NB, b0);
  resolve_rows(p, c, NB, b0, hrd, hwr, partB, p.H, nullptr, ch == 0);
  prof_mark(c, 21);
  layernorm_rows(c.hs, c.xs, NB, D);
  prof_mark(c, 22);
  gemm_ring<NB>(ring, c, c.xs, D, D, 2 * IC, c.bias, c.act, actw);
  prof_mark(c, 23);
  for (int i = threadIdx.x; i < NB * IC; i += kConsumers) {
    const int b = i / IC, j = i - b * IC;
    const float up = c.act[b * actw + j];
    const float gate = c.act[b * actw + IC + j];
    c.att[b * attw + j] = gate / (1.0f + expf(-gate)) * up;  // silu(gate) * up
  }
  csync();
  gemm_ring<NB>(ring, c, c.att, attw, IC, D, nullptr, c.hs, D);
  prof_mark(c, 25);
  store_partial(p, c, NB, b0, partC + (int64_t)ch * p.B * D);
  prof_mark(c, 26);
}

// ============================== phase F =================================
template <int NB>
__device__ void phase_final_ln(const DecoderParams& p, int item, Ctx& c, const float* hrd,
                               const float* partC) {
  const int D = p.D;
  const int b0 = item * NB;
  load_flags(p, c, NB, b0);
  resolve_rows(p, c, NB, b0, hrd, nullptr, partC, p.n_chunk, p.layers[p.L - 1].b2, false);
  layernorm_rows(c.hs, c.xs, NB, D);
  for (int i = threadIdx.x; i < NB * D; i += kConsumers) {
    const int b = i / D;
    if (b0 + b < p.B) p.xfin[(int64_t)(b0 + b) * D + (i - b * D)] = c.xs[i];
  }
  csync();
}

// ============================== phase G =================================
// Tied-embedding logits on the tensor cores (mma.sync m16n8k16, bf16x3 split, fp32
// accumulate): per pass a [32 utterances x D] tile of final-LN rows times the CTA's
// vocab slab embS[item] = [D][VCP] (k-major, row pitch VCP = vchunk + 4 keeps the
// B-fragment loads bank-conflict free), streamed through the ring 16 rows (one k-step)
// at a time.  The per-utterance argmax is fused: no logits leave the SM.
__device__ __forceinline__ int slab_rows_per_chunk(int vcp) {
  int r = (kStageBytes / (vcp * 4)) & ~15;
  return r < 16 ? 16 : r;
}
__device__ __forceinline__ void produce_logits(const DecoderParams& p, int item, Ring& ring) {
  const float* slab = p.embS + (int64_t)item * p.D * p.vcp;
  const int rpc = slab_rows_per_chunk(p.vcp);
  for (int b0 = 0; b0 < p.B; b0 += kLogitsTile)
    for (int k0 = 0; k0 < p.D; k0 += rpc) {
      const int rows = min(rpc, p.D - k0);
      ring.produce(slab + (size_t)k0 * p.vcp, (uint32_t)rows * p.vcp * 4);
    }
}

__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ void phase_logits(const DecoderParams& p, int item, Ctx& c, Ring& ring) {
  const int D = p.D, V = p.V, VC = p.vchunk, VCP = p.vcp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int XP = c.xpitch;
  float* xg = c.xg;                     // [32][XP] fp32 rows of the final LN
  const int NT = VC >> 3;               // n-tiles of 8 vocab entries (<= 32)
  const int parity = p.step & 1;
  const int rpc = slab_rows_per_chunk(VCP);
  for (int b0 = 0; b0 < p.B; b0 += kLogitsTile) {
    const int nb = min(kLogitsTile, p.B - b0);
    for (int i = threadIdx.x; i < kLogitsTile * D; i += kConsumers) {
      const int b = i / D, k = i - b * D;
      xg[b * XP + k] = (b < nb) ? p.xfin[(int64_t)(b0 + b) * D + k] : 0.f;
    }
    csync();
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0.f;
    for (int k0 = 0; k0 < D; k0 += rpc) {
      const int rows = min(rpc, D - k0);
      const float* W = reinterpret_cast<const float*>(ring.acquire());
      for (int kk = 0; kk < rows; kk += 16) {   // D % 16 == 0
        uint32_t ahi[2][4], alo[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          const float* xr0 = xg + (mt * 16 + g) * XP + k0 + kk + 2 * t;
          const float* xr1 = xr0 + 8 * XP;
          const float2 v0 = *reinterpret_cast<const float2*>(xr0);
          const float2 v1 = *reinterpret_cast<const float2*>(xr1);
          const float2 v2 = *reinterpret_cast<const float2*>(xr0 + 8);
          const float2 v3 = *reinterpret_cast<const float2*>(xr1 + 8);
          split_bf16x2(v0.x, v0.y, ahi[mt][0], alo[mt][0]);
          split_bf16x2(v1.x, v1.y, ahi[mt][1], alo[mt][1]);
          split_bf16x2(v2.x, v2.y, ahi[mt][2], alo[mt][2]);
          split_bf16x2(v3.x, v3.y, ahi[mt][3], alo[mt][3]);
        }
        uint32_t bhi[4][2], blo[4][2];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int nt = warp + j * kWarpsC;
          if (nt < NT) {  // warp-uniform
            const float* wr = W + (kk + 2 * t) * VCP + nt * 8 + g;
            split_bf16x2(wr[0], wr[VCP], bhi[j][0], blo[j][0]);
            split_bf16x2(wr[8 * VCP], wr[9 * VCP], bhi[j][1], blo[j][1]);
          }
        }
        // three split products; within each pass the 8 accumulators are independent
#pragma unroll
        for (int pass = 0; pass < 3; pass++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if (warp + j * kWarpsC < NT) {
#pragma unroll
              for (int mt = 0; mt < 2; mt++) {
                if (pass == 0) mma_bf16(acc[mt][j], alo[mt], bhi[j]);
                else if (pass == 1) mma_bf16(acc[mt][j], ahi[mt], blo[j]);
                else mma_bf16(acc[mt][j], ahi[mt], bhi[j]);
              }
            }
          }
      }
      ring.release();
    }
    // accumulator (mt, j): rows mt*16+g (c0,c1) and +8 (c2,c3), vocab item*VC + nt*8 + 2t (+1)
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
#pragma unroll
      for (int hrow = 0; hrow < 2; hrow++) {
        const int row = mt * 16 + g + hrow * 8;
        float bv = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int nt = warp + j * kWarpsC;
          if (nt < NT) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const int v = item * VC + nt * 8 + 2 * t + e;
              const float x = acc[mt][j][hrow * 2 + e];
              if (v < V) {
                if (p.logits_out && row < nb) p.logits_out[(int64_t)(b0 + row) * V + v] = x;
                if (x > bv || (x == bv && v < bi)) { bv = x; bi = v; }  // NaN never wins
              }
            }
          }
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {  // the 4 lanes of a group share the row
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (t == 0) { c.argv[warp * kLogitsTile + row] = bv; c.argi[warp * kLogitsTile + row] = bi; }
      }
    }
    csync();
    if (threadIdx.x < nb) {
      const int b = threadIdx.x;
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int w2 = 0; w2 < kWarpsC; w2++) {
        const float ov = c.argv[w2 * kLogitsTile + b];
        const int oi = c.argi[w2 * kLogitsTile + b];
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      p.cand_val[((int64_t)parity * p.n_vchunk + item) * p.B + b0 + b] = bv;
      p.cand_idx[((int64_t)parity * p.n_vchunk + item) * p.B + b0 + b] = bi;
    }
    csync();
  }
}

template <int NB, int NBM>
__global__ void __launch_bounds__(kThreads2, 1)
decoder_step2_kernel(const __grid_constant__ DecoderParams p) {
  if (*p.n_active == 0) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NBmax = NB > NBM ? NB : NBM;
  const SmemLayout2 L = smem_layout2(NBmax, p.B, p.D, p.hd, p.IC, p.Tpad, p.Smax, p.smem_limit);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
  unsigned char* active = smem_raw + L.active;
  Ring ring;
  ring.full = bars;
  ring.empty = bars + 16;
  ring.data = reinterpret_cast<char*>(smem_raw + L.ring);
  ring.ns = L.ns;
  ring.idx = 0;

  // snapshot of the done flags: the work list of this launch
  for (int b = threadIdx.x; b < p.B; b += kThreads2) active[b] = p.done[b] ? 0 : 1;
  if (threadIdx.x == 0) {
    for (int i = 0; i < L.ns; i++) {
      mbar_init(&ring.full[i], 1);
      mbar_init(&ring.empty[i], kWarpsC);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const unsigned G = gridDim.x;
  const int n_bt = (p.B + NB - 1) / NB;
  const int n_btm = (p.B + NBM - 1) / NBM;

  if (threadIdx.x >= kConsumers) {
    // ======================= producer warp =======================
    if (threadIdx.x == kConsumers) {
      for (int l = 0; l < p.L; l++) {
        for (int it = blockIdx.x; it < n_bt * p.H; it += G) produce_self(p, l, it, NB, ring, active);
        for (int it = blockIdx.x; it < n_bt * p.H; it += G) produce_cross(p, l, it, NB, ring, active);
        for (int it = blockIdx.x; it < n_btm * p.n_chunk; it += G) produce_mlp(p, l, it, NBM, ring, active);
      }
      for (int it = blockIdx.x; it < p.n_vchunk; it += G) produce_logits(p, it, ring);
    }
    return;
  }

  // ========================= consumers =========================
  Ctx c;
  c.hs = reinterpret_cast<float*>(smem_raw + L.hs);
  c.xs = reinterpret_cast<float*>(smem_raw + L.xs);
  c.act = reinterpret_cast<float*>(smem_raw + L.act);
  c.att = reinterpret_cast<float*>(smem_raw + L.att);
  c.red = reinterpret_cast<float*>(smem_raw + L.red);
  c.ps = reinterpret_cast<float*>(smem_raw + L.ps);
  c.sc = reinterpret_cast<float*>(smem_raw + L.sc);
  c.flags = reinterpret_cast<int*>(smem_raw + L.flags);
  c.argv = reinterpret_cast<float*>(smem_raw + L.argv);
  c.argi = reinterpret_cast<int*>(smem_raw + L.argi);
  c.active = active;
  c.actw = L.actw;
  c.attw = L.attw;
  {
    float* rope = reinterpret_cast<float*>(smem_raw + L.rope);
    const int half_rot = p.rot_dim >> 1;  // <= 64
    for (int i = threadIdx.x; i < half_rot; i += kConsumers) {
      rope[i] = p.rope_cos[(int64_t)p.step * half_rot + i];
      rope[64 + i] = p.rope_sin[(int64_t)p.step * half_rot + i];
    }
    c.rope = rope;
    c.bias = reinterpret_cast<float*>(smem_raw + L.bias);
    c.xg = reinterpret_cast<float*>(smem_raw);
    c.xpitch = L.xpitch;
  }
  c.prof = reinterpret_cast<unsigned long long*>(p.prof);
  c.prof_n = 0;
  prof_mark(c, 0);

  unsigned nbar = p.barrier[1];
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
        phase_self<NB>(p, l, it, c, ring, hrd, hwr, partC, partA);
      prof_mark(c, 7);
      grid_barrier(p.barrier, (++nbar) * G);
      prof_mark(c, 8);
      ph++;
    }
    {
      const float* hrd = p.hbuf + (int64_t)(ph & 1) * BD;
      float* hwr = p.hbuf + (int64_t)((ph + 1) & 1) * BD;
      for (int it = blockIdx.x; it < n_bt * p.H; it += G)
        phase_cross<NB>(p, l, it, c, ring, hrd, hwr, partA, partB);
      prof_mark(c, 17);
      grid_barrier(p.barrier, (++nbar) * G);
      prof_mark(c, 18);
      ph++;
    }
    {
      const float* hrd = p.hbuf + (int64_t)(ph & 1) * BD;
      float* hwr = p.hbuf + (int64_t)((ph + 1) & 1) * BD;
      for (int it = blockIdx.x; it < n_btm * p.n_chunk; it += G)
        phase_mlp<NBM>(p, l, it, c, ring, hrd, hwr, partB, partC);
      prof_mark(c, 27);
      grid_barrier(p.barrier, (++nbar) * G);
      prof_mark(c, 28);
      ph++;
    }
  }
  {
    const float* hrd = p.hbuf + (int64_t)(ph & 1) * BD;
    for (int it = blockIdx.x; it < p.B; it += G) phase_final_ln<1>(p, it, c, hrd, partC);
    prof_mark(c, 31);
    grid_barrier(p.barrier, (++nbar) * G);
    prof_mark(c, 32);
  }
  for (int it = blockIdx.x; it < p.n_vchunk; it += G) phase_logits(p, it, c, ring);
  prof_mark(c, 33);
  grid_barrier(p.barrier, (++nbar) * G);
  prof_mark(c, 34);
  if (blockIdx.x == 0) {
    int* cnt = c.flags;
    if (threadIdx.x == 0) *cnt = 0;
    csync();
    int n = 0;
    for (int b = threadIdx.x; b < p.B; b += kConsumers) n += p.done[b] ? 0 : 1;
    if (n) atomicAdd(cnt, n);
    csync();
    if (threadIdx.x == 0) {
      *p.n_active = *cnt;
      p.barrier[1] = nbar;
    }
  }
}

template <int NB, int NBM>
void launch_variant2(const DecoderParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = decoder_step2_kernel<NB, NBM>;
  static size_t configured_smem = 0;
  if (smem > configured_smem) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_smem = smem;
  }
  void* args[] = {const_cast<DecoderParams*>(&p)};
  CUDA_CHECK(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kThreads2), args, smem, stream));
}

}  // namespace

size_t decoder_step2_smem_bytes(const DecoderParams& p) {
  int nb, nbm;
  decoder_tiles_for_batch(p.B, nb, nbm);
  const SmemLayout2 L = smem_layout2(nb > nbm ? nb : nbm, p.B, p.D, p.hd, p.IC, p.Tpad, p.Smax, p.smem_limit);
  if (L.ns < 2) throw std::runtime_error("decoder v2: not enough shared memory for the operand ring");
  return (size_t)L.total;
}

void launch_decoder_step2(const DecoderParams& p, int grid, cudaStream_t stream) {
  int nb, nbm;
  decoder_tiles_for_batch(p.B, nb, nbm);
  const size_t smem = decoder_step2_smem_bytes(p);
  if (nb == 1 && nbm == 1) launch_variant2<1, 1>(p, grid, smem, stream);
  else if (nb == 1) launch_variant2<1, 2>(p, grid, smem, stream);
  else if (nb == 2) launch_variant2<2, 4>(p, grid, smem, stream);
  else if (nb == 4) launch_variant2<4, 8>(p, grid, smem, stream);
  else if (nb == 8) launch_variant2<8, 16>(p, grid, smem, stream);
  else launch_variant2<16, 16>(p, grid, smem, stream);
}

}  // namespace msb
