// Persistent decoder-step kernel, v4: one thread-block CLUSTER per group of utterances, layers resident in the
// cluster, exchanges through distributed shared memory.  Small batches (<= 8 utterances per cluster).
//
// Why (measured on B200, profiles/r2b_*, scripts/ubench/icache.cu): a hand-off between CTAs through L2 costs three
// round trips of ~0.65 us (release, poll, first load) and the weight-stationary v3 kernel needs 5 of them per layer;
// and one CTA per SM executing a long program is instruction-fetch bound once the code it walks exceeds ~128 KB.
// So this kernel keeps a whole decoder layer inside a cluster of CS CTAs (16, or 8) that owns U <= 8 utterances:
//   * every CTA holds a copy of the cluster's residual rows in shared memory;
//   * every dense contraction is split across the ranks (head h's q|k|v and cross q on the ranks of head h; output
//     projections and fc1 by output feature; fc2 by input slice), fp32 SIMT from k-major weight slices that arrive
//     through the TMA bulk-copy ring -- ONE split-K routine, driven by a small stage table, so the layer's code is
//     a few thousand instructions and stays in the instruction cache;
//   * results are exchanged with st/ld.shared::cluster and a cluster-scope mbarrier (6 per layer, ~0.3 us each)
//     instead of L2 round trips; producers never take part in it.
// Every cluster streams all layer weights (U <= 8 => the weights are read NC times per step out of L2: 8 x 36 MB
// for moonshine-tiny at batch 32, which is why large batches stay on the weight-stationary v3 kernel); cross K/V
// is streamed with an L2 evict-first policy so the weights stay resident.
// The step ends like v3: final LayerNorm rows to global, ONE grid-wide hand-off, tied head on tcgen05 + fused argmax.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <atomic>

#include "common.h"
#include "kernels.h"

namespace msb {

namespace {

constexpr int kConsumers = 256;
constexpr int kProducers = 2;
constexpr int kThreads4 = kConsumers + 32 * kProducers;
constexpr int kWarpsC = kConsumers / 32;
constexpr int kStageBytes = 32768;
constexpr int kUmax = 8;                         // utterances per cluster
constexpr int kMaxPairs = 8;                     // float2 per lane of a LayerNorm row: D <= 512
constexpr long long kSpinLimit = 4000000000LL;   // ~2 s of SM cycles
constexpr int kPlaneBatch = 4;
constexpr int kBiasEntries = 512;

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
__device__ __forceinline__ bool poisoned(const unsigned* err) { return *reinterpret_cast<const volatile unsigned*>(err) != 0u; }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / bulk copy ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t a, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(a), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned* err) {
  const uint32_t a = smem_u32(bar);
  if (mbar_try(a, parity)) return;
  const long long t0 = clock64();
  unsigned polls = 0;
  while (!mbar_try(a, parity)) {
    if ((++polls & 15u) == 0u) {
      if (poisoned(err)) return;
      if (clock64() - t0 > kSpinLimit) { atomicExch(err, 1u); return; }
    }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// ---- cluster: rank, distributed shared memory, cluster-scope barrier built on one mbarrier per CTA ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster(uint32_t raddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_cluster(uint32_t raddr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_hw_sync() {  // every thread of the cluster (kernel start only)
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Consumer-side cluster barrier: after the CTA barrier one thread arrives (release.cluster) on the barrier of every
// rank and waits (acquire.cluster) on its own, which collects one arrival per rank.  Remote stores issued before it
// are visible to every rank after it.  Producer warps are not involved.
// TWO barrier objects used alternately: a rank that has left barrier k arrives for k+1 on the OTHER object, so
// its arrival can never be counted into a slower rank's still-open phase k (it cannot reach k+2 before everyone has
// arrived for k+1, i.e. finished arriving for k).
__device__ __forceinline__ void cluster_barrier(uint64_t* cbars, uint32_t& phase, int CS, unsigned* err) {
  csync();
  uint64_t* cbar = cbars + (phase & 1u);
  const uint32_t parity = (phase >> 1) & 1u;
  if (threadIdx.x < 32) {
    const uint32_t a = smem_u32(cbar);
    if ((int)threadIdx.x < CS) {  // lane r arrives on rank r's barrier: the remote arrives overlap
      const uint32_t ra = mapa(a, threadIdx.x);
      asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) {
      uint32_t done = 0;
      const long long t0 = clock64();
      unsigned polls = 0;
      while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) break;
        if ((++polls & 15u) == 0u) {
          if (poisoned(err)) break;
          if (clock64() - t0 > kSpinLimit) { atomicExch(err, 1u); break; }
        }
      }
    }
  }
  phase++;
  csync();
}

// ---- tcgen05 (logits phase only) ----
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
  const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah));
  const __nv_bfloat16 bl = __float2bfloat16_rn(b - __bfloat162float(bh));
  hi = (uint32_t)__bfloat16_as_ushort(ah) | ((uint32_t)__bfloat16_as_ushort(bh) << 16);
  lo = (uint32_t)__bfloat16_as_ushort(al) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
}
__device__ __forceinline__ uint32_t plane_off(int NXp, int r, int k) {
  return (uint32_t)((k >> 5) * NXp * 128 + r * 64 + ((((k >> 3) & 3) ^ ((r >> 1) & 3)) << 4) + (k & 7) * 2);
}

// ---- shared memory layout ----
struct Smem4 {
  int bars, active, flags, rope, argv, argi, sbias, rowflag, ent_key, ent_val, ent_n;  // byte offsets
  int hres, xs, attg, qkv, act, actT, part, ytmp, red, sc, ps, redx, xg, ring;
  int qw, isp, dsp, ub, ns, nxl, total;
};
__host__ __device__ inline int logits_rows4(int B, int D) {
  int nx = (B + 15) & ~15;
  if (nx > 64) nx = 64;
  while (nx > 16 && nx * D * 4 > 80 * 1024) nx -= 16;
  return nx;
}
__host__ __device__ inline Smem4 smem_layout4(int B, int D, int hd, int I, int CS, int U, int Tpad, int Smax, int smem_limit) {
  Smem4 L;
  int o = 0;
  auto take = [&](int bytes) { int r = o; o += (bytes + 15) / 16 * 16; return r; };
  L.ub = U <= 6 ? (U < 1 ? 1 : U) : 8;   // rows the dense routine computes (template instances 1..6, 8)
  L.qw = 3 * hd;
  L.isp = I / CS;
  L.dsp = (D / CS + 3) & ~3;
  L.bars = take((2 * 16 + 4) * 8);  // ring full / empty, accumulator barrier, two cluster barriers, TMEM base
  L.active = take(16);
  L.flags = take(64 * 4);
  L.rope = take(128 * 4);
  L.argv = take(kWarpsC * 64 * 4);
  L.argi = take(kWarpsC * 64 * 4);
  L.sbias = take(384 * 4);
  L.rowflag = take(384);
  L.ent_key = take(kBiasEntries * 4);
  L.ent_val = take(kBiasEntries * 4);
  L.ent_n = take(16);
  o = (o + 1023) / 1024 * 1024;
  L.xg = o;  // logits x planes alias everything from here to the ring (the layer state is dead by then)
  L.hres = take(kUmax * D * 4);
  L.xs = take(kUmax * D * 4);
  L.attg = take(kUmax * D * 4);
  L.part = take(kUmax * D * 4);
  L.qkv = take(kUmax * L.qw * 4);
  L.act = take(kUmax * (2 * L.isp) * 4);   // fc1 output (value | gate interleaved)
  L.actT = take(L.isp * 8 * 4);            // gated product, transposed [I/CS][8]
  L.ytmp = take(kUmax * L.dsp * 4);
  L.red = take(L.ub * 1024 * 4 + 64);       // split-K partials: (k-slice, utterance) x output features <= 1024 floats / utterance
  L.sc = take(kWarpsC * (Smax + 4) * 4);
  L.ps = take(2 * Tpad * 4);
  L.redx = take((32 + 1024) * 4);
  L.nxl = logits_rows4(B, D);
  if (L.xg + L.nxl * D * 4 > o) o = L.xg + L.nxl * D * 4;
  o = (o + 1023) / 1024 * 1024;
  L.ring = o;
  int ns = (smem_limit - o) / kStageBytes;
  if (ns > 16) ns = 16;
  L.ns = ns;
  L.total = o + ns * kStageBytes;
  return L;
}

// ---- ring cursor (registers; uniform over the consumer threads) ----
struct Cur {
  int st;
  uint32_t par;
};
__device__ __forceinline__ void cur_advance(Cur& c, int ns) {
  if (++c.st == ns) { c.st = 0; c.par ^= 1u; }
}
struct RingRef {
  uint64_t* full;
  uint64_t* empty;
  char* data;
  unsigned* err;
  int ns;
};
__device__ __forceinline__ const char* ring_acquire(const RingRef& r, const Cur& c) {
  mbar_wait(&r.full[c.st], c.par, r.err);
  return r.data + (size_t)c.st * kStageBytes;
}
__device__ __forceinline__ void ring_release(const RingRef& r, Cur& c) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&r.empty[c.st]);
  cur_advance(c, r.ns);
}
// producers: every producer lane walks the whole chunk sequence and issues every kProducers-th chunk
struct PCur {
  int st;
  uint32_t par;
  int turn;
};
__device__ __forceinline__ void ring_produce(const RingRef& r, PCur& c, const void* src, uint32_t bytes, uint64_t policy) {
  if (c.turn == 0) {
    mbar_wait(&r.empty[c.st], c.par ^ 1u, r.err);
    mbar_expect_tx(&r.full[c.st], bytes);
    bulk_g2s(r.data + (size_t)c.st * kStageBytes, src, bytes, &r.full[c.st], policy);
    c.turn = kProducers;
  }
  c.turn--;
  if (++c.st == r.ns) { c.st = 0; c.par ^= 1u; }
}
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {  // bytes: multiple of 16
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ int rows_per_chunk_f32(int K, int N) {
  int r = kStageBytes / (N * 4);
  return r < 1 ? 1 : (r > K ? K : r);
}
__device__ __forceinline__ int rows_per_chunk_f16(int rows, int cols) {
  int r = (kStageBytes / (cols * 2)) & ~1;
  if (r < 2) r = 2;
  return r > rows ? rows : r;
}
__device__ __forceinline__ void produce_f32(const RingRef& r, PCur& c, const float* Wt, int K, int N, uint64_t pol) {
  const int rpc = rows_per_chunk_f32(K, N);
  for (int k0 = 0; k0 < K; k0 += rpc) {
    const int rows = min(rpc, K - k0);
    ring_produce(r, c, Wt + (size_t)k0 * N, (uint32_t)rows * N * 4, pol);
  }
}
__device__ __forceinline__ void produce_f16(const RingRef& r, PCur& c, const __half* M, int rows, int cols, uint64_t pol) {
  const int rpc = rows_per_chunk_f16(rows, cols);
  for (int r0 = 0; r0 < rows; r0 += rpc) {
    const int n = min(rpc, rows - r0);
    ring_produce(r, c, M + (size_t)r0 * cols, (uint32_t)n * cols * 2, pol);
  }
}

// ---- token bookkeeping (same rules as v3) ----
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
  if (finished && writer && lane == 0) p.done[b] = p.step + 1;
  if (tok_in < 0 || tok_in >= p.V) tok_in = 0;
  return tok_in;
}

// ---- LayerNorm of the cluster's residual rows (no affine: gamma lives in the next weight block): warp per row ----
__device__ __forceinline__ void ln_rows(const float* hres, float* xs, int U, int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < U) {
    const float* h = hres + warp * D;
    float2 v[kMaxPairs];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPairs; i++) {
      const int k = 2 * lane + 64 * i;
      v[i] = (k < D) ? *reinterpret_cast<const float2*>(h + k) : make_float2(0.f, 0.f);
      s += v[i].x + v[i].y;
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPairs; i++)
      if (2 * lane + 64 * i < D) {
        const float dx = v[i].x - mean, dy = v[i].y - mean;
        q += dx * dx + dy * dy;
      }
    const float rstd = rsqrtf(warp_sum(q) / D + 1e-5f);
#pragma unroll
    for (int i = 0; i < kMaxPairs; i++) {
      const int k = 2 * lane + 64 * i;
      if (k < D) {  // transposed: xT[k][u]
        xs[(size_t)k * 8 + warp] = (v[i].x - mean) * rstd;
        xs[(size_t)(k + 1) * 8 + warp] = (v[i].y - mean) * rstd;
      }
    }
  }
  csync();
}

// ---- the one dense routine: out[u][n] = sum_k xT[k][u] * Wt[k][n] (+ bias[n]), Wt k-major fp32 through the ring ----
// threads = (k-slice s, 4 features n4); the activations are stored TRANSPOSED, 8 utterances per input row (32 B), so
// one LDS.128 brings an input of 4 utterances; 4 weight rows per trip, every load issued before the first FMA;
// the partial sums of the k-slices are reduced through shared memory.
struct Prof {
  unsigned long long* buf;
  int n;
};
__device__ __forceinline__ void pmark(Prof& pf, int tag) {
  if (pf.buf != nullptr && threadIdx.x == 0 && pf.n < 512) {
    pf.buf[(size_t)blockIdx.x * 512 + pf.n] = ((unsigned long long)clock64() << 8) | (unsigned)tag;
    pf.n++;
  }
}
template <int UB>
__device__ __forceinline__ void gemv(const RingRef& ring, Cur& cur, const float* xT, int K, int N, const float* __restrict__ bias,
                                     float* red, float* out, int ldo, Prof& pf) {
  const int N4 = N >> 2;
  int S = kConsumers / N4;
  if (S > K) S = K;
  const int t = threadIdx.x;
  const int n4 = t % N4, s = t / N4;
  const bool on = s < S;
  float acc[UB][4];
#pragma unroll
  for (int b = 0; b < UB; b++) acc[b][0] = acc[b][1] = acc[b][2] = acc[b][3] = 0.f;
  const int rpc = rows_per_chunk_f32(K, N);
  for (int k0 = 0; k0 < K; k0 += rpc) {
    const int rows = min(rpc, K - k0);
    const float4* W = reinterpret_cast<const float4*>(ring_acquire(ring, cur));
    pmark(pf, 50);
    if (on) {
#pragma unroll 1
      for (int r = s; r < rows; r += 4 * S) {
        float4 w[4], xa[4], xb[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int rr = r + e * S;
          const bool ok = rr < rows;
          w[e] = ok ? W[rr * N4 + n4] : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4* xr = reinterpret_cast<const float4*>(xT + (size_t)(k0 + (ok ? rr : r)) * 8);
          xa[e] = xr[0];
          if (UB > 4) xb[e] = xr[1];
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float xv[8] = {xa[e].x, xa[e].y, xa[e].z, xa[e].w, UB > 4 ? xb[e].x : 0.f, UB > 4 ? xb[e].y : 0.f,
                               UB > 4 ? xb[e].z : 0.f, UB > 4 ? xb[e].w : 0.f};
#pragma unroll
          for (int b = 0; b < UB; b++) {
            acc[b][0] = fmaf(xv[b], w[e].x, acc[b][0]);
            acc[b][1] = fmaf(xv[b], w[e].y, acc[b][1]);
            acc[b][2] = fmaf(xv[b], w[e].z, acc[b][2]);
            acc[b][3] = fmaf(xv[b], w[e].w, acc[b][3]);
          }
        }
      }
    }
    pmark(pf, 51);
    ring_release(ring, cur);
  }
  if (on) {
#pragma unroll
    for (int b = 0; b < UB; b++)
      *reinterpret_cast<float4*>(&red[(s * UB + b) * N + n4 * 4]) = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
  }
  csync();
  pmark(pf, 52);
  for (int i = threadIdx.x; i < UB * N4; i += kConsumers) {
    const int b = i / N4, c4 = i - b * N4;
    float4 v = bias ? __ldg(reinterpret_cast<const float4*>(bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s2 = 0; s2 < S; s2++) {
      const float4 q = *reinterpret_cast<const float4*>(&red[(s2 * UB + b) * N + c4 * 4]);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    *reinterpret_cast<float4*>(&out[b * ldo + c4 * 4]) = v;
  }
  csync();
}

// x planes of the logits pass from the fp32 final rows (global, written by every cluster before the grid hand-off)
__device__ __forceinline__ void planes_from_rows(unsigned char* planes, int NXp, const float* src, int64_t ld, int g0, int nb, int K,
                                                 const unsigned char* done_mask /*global int32 done[]*/, const int* done, int step) {
  (void)done_mask;
  const int k8n = (K + 31) / 32 * 4;
  const int nitems = NXp * k8n;
  for (int i0 = threadIdx.x; i0 < nitems; i0 += kPlaneBatch * kConsumers) {
    float4 va[kPlaneBatch], vb[kPlaneBatch];
#pragma unroll
    for (int u = 0; u < kPlaneBatch; u++) {
      const int i = i0 + u * kConsumers;
      const int r = i / k8n, k8 = i - r * k8n;
      va[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      vb[u] = va[u];
      if (i < nitems && r < nb) {
        const int d = done[g0 + r];
        if (d == 0 || d > step) {
          const float* s = src + (int64_t)(g0 + r) * ld + k8 * 8;
          if (k8 * 8 < K) va[u] = *reinterpret_cast<const float4*>(s);
          if (k8 * 8 + 4 < K) vb[u] = *reinterpret_cast<const float4*>(s + 4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kPlaneBatch; u++) {
      const int i = i0 + u * kConsumers;
      if (i >= nitems) break;
      const int r = i / k8n, k8 = i - r * k8n;
      uint4 hi, lo;
      split_bf16x2(va[u].x, va[u].y, hi.x, lo.x);
      split_bf16x2(va[u].z, va[u].w, hi.y, lo.y);
      split_bf16x2(vb[u].x, vb[u].y, hi.z, lo.z);
      split_bf16x2(vb[u].z, vb[u].w, hi.w, lo.w);
      unsigned char* at = planes + plane_off(NXp, r, k8 * 8);
      *reinterpret_cast<uint4*>(at) = hi;
      *reinterpret_cast<uint4*>(at + NXp * 64) = lo;
    }
  }
}

__global__ void __launch_bounds__(kThreads4, 1) decoder_step4_kernel(const __grid_constant__ DecoderParams p) {
  if (*p.n_active == 0) return;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int CS = p.c4_cs, U = p.c4_u;
  const int D = p.D, hd = p.hd, H = p.H, I = p.I;
  const Smem4 L = smem_layout4(p.B, D, hd, I, CS, U, p.Tpad, p.Smax, p.smem_limit);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + L.bars);
  unsigned* err = p.sync3 + 1;
  RingRef ring;
  ring.full = bars;
  ring.empty = bars + 16;
  ring.data = reinterpret_cast<char*>(smem_raw + L.ring);
  ring.err = err;
  ring.ns = L.ns;
  uint64_t* acc_bar = bars + 32;
  uint64_t* cbar = bars + 33;  // [2]
  uint32_t& tmem_base_smem = *reinterpret_cast<uint32_t*>(bars + 35);

  const int rank = (int)cluster_ctarank();
  const int cid = (int)cluster_id_x();
  const int u0 = cid * U;                                   // first utterance of this cluster
  const int nu = max(0, min(U, p.B - u0));                  // utterances this cluster holds
  const int RH = CS / H;                                    // ranks per head
  const int hh = rank / RH, sub = rank - hh * RH;
  const int ds = D / CS, dsp = L.dsp, is = L.isp;
  unsigned char* active = smem_raw + L.active;

  if (threadIdx.x < kUmax) {
    const int b = u0 + threadIdx.x;
    int a = 0;
    if ((int)threadIdx.x < nu) {
      const int d = p.done[b];
      a = (d == 0 || d > p.step) ? 1 : 0;
    }
    active[threadIdx.x] = (unsigned char)a;
  }
  const unsigned epoch = *reinterpret_cast<const volatile unsigned*>(p.sync3);
  if (threadIdx.x == 0) {
    for (int i = 0; i < L.ns; i++) {
      mbar_init(&ring.full[i], 1);
      mbar_init(&ring.empty[i], kWarpsC);
    }
    mbar_init(acc_bar, 1);
    mbar_init(cbar, (uint32_t)CS);
    mbar_init(cbar + 1, (uint32_t)CS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  cluster_hw_sync();  // every rank's barriers exist before anyone arrives on them remotely
  const uint32_t tmem_base = tmem_base_smem;
  bool any_active = false;
  for (int u = 0; u < nu; u++) any_active |= active[u] != 0;

  uint64_t pol_keep, pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));

  const int n_vjobs = p.n_vchunk;
  const int G = (int)gridDim.x;

  if (threadIdx.x >= kConsumers) {
    // ======================= producer warps (one lane each) =======================
    if ((threadIdx.x & 31) == 0) {
      PCur pc;
      const int pidx = (threadIdx.x - kConsumers) >> 5;
      pc.st = 0; pc.par = 0; pc.turn = pidx;
      if (any_active) {
        for (int l = 0; l < p.L; l++) {
          const DecLayerWeights& w = p.layers[l];
          produce_f32(ring, pc, w.wqkv + (int64_t)hh * D * 3 * hd, D, 3 * hd, pol_keep);
          if (p.step > 0 && pc.turn == 0) {  // self K/V prefix of this rank's items: ask L2 for it a stage ahead
            for (int u = sub; u < nu; u += RH) {
              if (!active[u]) continue;
              const int64_t bh = ((int64_t)l * p.B + (u0 + u)) * H + hh;
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.ks + bh * hd * p.Smax), "r"((uint32_t)(hd * p.Smax * 4)) : "memory");
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.vs + bh * p.Smax * hd), "r"((uint32_t)(p.step * hd * 4)) : "memory");
            }
          }
          produce_f32(ring, pc, w.c4_wo + (int64_t)rank * D * dsp, D, dsp, pol_keep);
          produce_f32(ring, pc, w.wqc + (int64_t)hh * D * hd, D, hd, pol_keep);
          for (int u = sub; u < nu; u += RH) {
            if (!active[u]) continue;
            const int64_t bh = ((int64_t)l * p.B + (u0 + u)) * H + hh;
            produce_f16(ring, pc, p.kc + bh * hd * p.Tpad, hd, p.Tpad, pol_stream);
            produce_f16(ring, pc, p.vc + bh * p.Tpad * hd, p.Tpad, hd, pol_stream);
          }
          produce_f32(ring, pc, w.c4_woc + (int64_t)rank * D * dsp, D, dsp, pol_keep);
          produce_f32(ring, pc, w.c4_w1 + (int64_t)rank * D * 2 * is, D, 2 * is, pol_keep);
          produce_f32(ring, pc, w.c4_w2 + (int64_t)rank * is * D, is, D, pol_keep);
        }
      }
      // logits: the CTA's vocab chunks as plane-packed slabs, two 32-wide k-blocks of one m-tile per chunk
      const int VC = p.vchunk, n_mt = (VC + 127) >> 7, nkb = D >> 5;
      for (int b0 = 0; b0 < p.B; b0 += L.nxl) {
        for (int item = (int)blockIdx.x; item < n_vjobs; item += G) {
          const unsigned char* slab = reinterpret_cast<const unsigned char*>(p.embP) + (size_t)item * VC * D * 4;
          size_t mt_off = 0;
          for (int mt = 0; mt < n_mt; mt++) {
            const int R = min(128, VC - mt * 128);
            for (int kb = 0; kb < nkb; kb += 2) {
              const int n = min(2, nkb - kb);
              ring_produce(ring, pc, slab + mt_off + (size_t)kb * R * 128, (uint32_t)(n * R * 128), pol_stream);
            }
            mt_off += (size_t)R * D * 4;
          }
        }
      }
    }
    return;
  }

  // ========================= consumers =========================
  float* hres = reinterpret_cast<float*>(smem_raw + L.hres);
  float* xs = reinterpret_cast<float*>(smem_raw + L.xs);
  float* attg = reinterpret_cast<float*>(smem_raw + L.attg);
  float* part = reinterpret_cast<float*>(smem_raw + L.part);
  float* qkv = reinterpret_cast<float*>(smem_raw + L.qkv);
  float* act = reinterpret_cast<float*>(smem_raw + L.act);
  float* actT = reinterpret_cast<float*>(smem_raw + L.actT);
  float* ytmp = reinterpret_cast<float*>(smem_raw + L.ytmp);
  float* red = reinterpret_cast<float*>(smem_raw + L.red);
  float* scb = reinterpret_cast<float*>(smem_raw + L.sc);
  float* psb = reinterpret_cast<float*>(smem_raw + L.ps);
  float* redx = reinterpret_cast<float*>(smem_raw + L.redx);
  int* flags = reinterpret_cast<int*>(smem_raw + L.flags);
  float* rope = reinterpret_cast<float*>(smem_raw + L.rope);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qw = L.qw;
  Cur cur;
  cur.st = 0; cur.par = 0;
  uint32_t cpar = 0;
  Prof pf;
  pf.buf = reinterpret_cast<unsigned long long*>(p.prof);
  pf.n = 0;
  auto mark = [&](int tag) { pmark(pf, tag); };
  mark(0);
  {
    const int half_rot = p.rot_dim >> 1;
    for (int i = threadIdx.x; i < half_rot; i += kConsumers) {
      rope[i] = p.rope_cos[(int64_t)p.step * half_rot + i];
      rope[64 + i] = p.rope_sin[(int64_t)p.step * half_rot + i];
    }
    if (threadIdx.x < kUmax) flags[32 + threadIdx.x] = ((int)threadIdx.x < nu && active[threadIdx.x]) ? p.enc_len[u0 + threadIdx.x] : 0;
  }

  if (any_active) {
    // rows beyond the cluster's utterances (and padding columns) feed the dense routine too: keep them finite
    for (int i = threadIdx.x * 4; i < (L.sc - L.hres) / 4; i += kConsumers * 4)
      *reinterpret_cast<float4*>(smem_raw + L.hres + i * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    csync();
    // ---- tokens + embedding: every rank builds the same residual rows ----
    if (warp < nu && active[warp]) {
      const int tok = step_prologue_warp(p, u0 + warp, rank == 0);
      for (int k = lane * 4; k < D; k += 128)
        *reinterpret_cast<float4*>(hres + warp * D + k) = __ldg(reinterpret_cast<const float4*>(p.embed + (int64_t)tok * D + k));
    } else if (warp < kUmax) {
      for (int k = lane * 4; k < D; k += 128) *reinterpret_cast<float4*>(hres + warp * D + k) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    csync();
    mark(1);
    const float scale = rsqrtf((float)hd);
    const uint32_t attg_s = smem_u32(attg), hres_s = smem_u32(hres), part_s = smem_u32(part);
#pragma unroll 1
    for (int l = 0; l < p.L; l++) {
      const DecLayerWeights& w = p.layers[l];
#pragma unroll 1
      for (int stage = 0; stage < 6; stage++) {
        // ---------- before the contraction ----------
        if (stage == 0 || stage == 2 || stage == 4) ln_rows(hres, xs, kUmax < U ? kUmax : U, D);
        // ---------- the contraction of this stage (ONE code instance, table-driven) ----------
        const float* x; int K, N; const float* bias = nullptr; float* out; int ldo;
        switch (stage) {   // x: transposed activations [K][8]
          case 0: x = xs; K = D; N = 3 * hd; out = qkv; ldo = qw; break;
          case 1: x = attg; K = D; N = dsp; out = ytmp; ldo = dsp; break;
          case 2: x = xs; K = D; N = hd; out = qkv; ldo = qw; break;
          case 3: x = attg; K = D; N = dsp; out = ytmp; ldo = dsp; break;
          case 4: x = xs; K = D; N = 2 * is; bias = w.c4_b1 + (int64_t)rank * 2 * is; out = act; ldo = 2 * is; break;
          default: x = actT; K = is; N = D; out = part; ldo = D; break;
        }
        switch (L.ub) {
          case 8: gemv<8>(ring, cur, x, K, N, bias, red, out, ldo, pf); break;
          case 6: gemv<6>(ring, cur, x, K, N, bias, red, out, ldo, pf); break;
          case 5: gemv<5>(ring, cur, x, K, N, bias, red, out, ldo, pf); break;
          case 4: gemv<4>(ring, cur, x, K, N, bias, red, out, ldo, pf); break;
          case 3: gemv<3>(ring, cur, x, K, N, bias, red, out, ldo, pf); break;
          case 2: gemv<2>(ring, cur, x, K, N, bias, red, out, ldo, pf); break;
          default: gemv<1>(ring, cur, x, K, N, bias, red, out, ldo, pf); break;
        }
        mark(10 + stage);
        // ---------- after it ----------
        if (stage == 0) {
          // RoPE (interleaved pairs) on q and k of every row, then K/V append + causal self-attention of this
          // rank's utterances (u = sub, sub + RH, ...): one warp per utterance, prefix straight from global
          const int half_rot = p.rot_dim >> 1;
          for (int i = threadIdx.x; i < nu * 2 * half_rot; i += kConsumers) {
            const int b = i / (2 * half_rot);
            const int r = i - b * 2 * half_rot;
            const int which = r / half_rot, pr = r - which * half_rot;
            const float cs = rope[pr], sn = rope[64 + pr];
            float* v = qkv + b * qw + which * hd + 2 * pr;
            const float x0 = v[0], x1 = v[1];
            v[0] = x0 * cs - x1 * sn;
            v[1] = x1 * cs + x0 * sn;
          }
          csync();
          mark(40);
          {
            const int u = sub + warp * RH;     // warp j takes this rank's j-th utterance
            if (u < nu && active[u]) {
              const int64_t bh = ((int64_t)l * p.B + (u0 + u)) * H + hh;
              float* Kt = p.ks + bh * hd * p.Smax;
              float* Vr = p.vs + bh * p.Smax * hd;
              const float* q = qkv + u * qw;
              const float* kcur = q + hd;
              const float* vcur = q + 2 * hd;
              for (int d = lane; d < hd; d += 32) {
                Kt[(int64_t)d * p.Smax + p.step] = kcur[d];
                Vr[(int64_t)p.step * hd + d] = vcur[d];
              }
              float* sc = scb + warp * (p.Smax + 4);
              float mx = -INFINITY;
              for (int t0 = 0; t0 < p.step; t0 += 64) {
                const int ta = t0 + lane, tb = t0 + 32 + lane;
                const bool va = ta < p.step, vb = tb < p.step;
                float sa = 0.f, sb = 0.f;
#pragma unroll 1
                for (int d0 = 0; d0 < hd; d0 += 20) {   // 40 independent loads in flight per lane and trip
                  float ka[20], kb[20];
#pragma unroll
                  for (int e = 0; e < 20; e++) {
                    const bool in = d0 + e < hd;
                    ka[e] = (va && in) ? __ldg(Kt + (int64_t)(d0 + e) * p.Smax + ta) : 0.f;
                    kb[e] = (vb && in) ? __ldg(Kt + (int64_t)(d0 + e) * p.Smax + tb) : 0.f;
                  }
#pragma unroll
                  for (int e = 0; e < 20; e++) {
                    const float qd = d0 + e < hd ? q[d0 + e] : 0.f;
                    sa = fmaf(qd, ka[e], sa);
                    sb = fmaf(qd, kb[e], sb);
                  }
                }
                if (va) { sa *= scale; sc[ta] = sa; mx = fmaxf(mx, sa); }
                if (vb) { sb *= scale; sc[tb] = sb; mx = fmaxf(mx, sb); }
              }
              {
                float s = 0.f;
                for (int d = lane; d < hd; d += 32) s = fmaf(q[d], kcur[d], s);
                s = warp_sum(s) * scale;
                if (lane == 0) sc[p.step] = s;
                mx = fmaxf(mx, s);
              }
              mx = warp_max(mx);
              __syncwarp();
              mark(41);
              float sum = 0.f;
              for (int t = lane; t <= p.step; t += 32) {
                const float e = expf(sc[t] - mx);
                sc[t] = e;
                sum += e;
              }
              const float inv = 1.0f / warp_sum(sum);
              __syncwarp();
              float o0 = 0.f, o1 = 0.f;
              const bool has0 = lane < hd, has1 = lane + 32 < hd;
              int t = 0;
#pragma unroll 1
              for (; t + 24 <= p.step; t += 24) {
                float v0[24], v1[24];
#pragma unroll
                for (int e = 0; e < 24; e++) {
                  v0[e] = has0 ? __ldg(Vr + (int64_t)(t + e) * hd + lane) : 0.f;
                  v1[e] = has1 ? __ldg(Vr + (int64_t)(t + e) * hd + lane + 32) : 0.f;
                }
#pragma unroll
                for (int e = 0; e < 24; e++) {
                  const float pt = sc[t + e];
                  o0 = fmaf(pt, v0[e], o0);
                  o1 = fmaf(pt, v1[e], o1);
                }
              }
              for (; t < p.step; t++) {
                const float pt = sc[t];
                if (has0) o0 = fmaf(pt, __ldg(Vr + (int64_t)t * hd + lane), o0);
                if (has1) o1 = fmaf(pt, __ldg(Vr + (int64_t)t * hd + lane + 32), o1);
              }
              const float pl = sc[p.step];
              mark(43);
              // attention output of (u, head hh) -> every rank's gathered copy
              if (has0) {
                const float v = fmaf(pl, vcur[lane], o0) * inv;
                const uint32_t a = attg_s + (uint32_t)(((hh * hd + lane) * 8 + u) * 4);
                for (int r = 0; r < CS; r++) st_cluster(mapa(a, (uint32_t)r), v);
              }
              if (has1) {
                const float v = fmaf(pl, vcur[lane + 32], o1) * inv;
                const uint32_t a = attg_s + (uint32_t)(((hh * hd + lane + 32) * 8 + u) * 4);
                for (int r = 0; r < CS; r++) st_cluster(mapa(a, (uint32_t)r), v);
              }
            }
          }
          mark(44);
          cluster_barrier(cbar, cpar, CS, err);
        } else if (stage == 1 || stage == 3) {
          // new residual slice of this rank -> every rank's copy
          for (int i = threadIdx.x; i < nu * ds; i += kConsumers) {
            const int u = i / ds, j = i - u * ds;
            const int n = rank * ds + j;
            const float v = hres[u * D + n] + ytmp[u * dsp + j];
            const uint32_t a = hres_s + (uint32_t)((u * D + n) * 4);
            for (int r = 0; r < CS; r++) st_cluster(mapa(a, (uint32_t)r), v);
          }
          cluster_barrier(cbar, cpar, CS, err);
        } else if (stage == 2) {
          // cross-attention of this rank's utterances over the fp16 cross K/V (ring), two at a time (one per half CTA)
          const int Tpad = p.Tpad;
          const int tpr = hd >> 2;
          const bool halves = Tpad <= 512;
          const int half = halves ? (int)(threadIdx.x >> 7) : 0;
          const int gtid = halves ? (int)(threadIdx.x & 127) : (int)threadIdx.x;
          const int gthreads = halves ? 128 : kConsumers;
          const int gwarps = gthreads >> 5, gwarp = gtid >> 5;
          const int Gv = gthreads / tpr;
          float* ps = psb + (halves ? half * Tpad : 0);
          float* red_max = redx + half * 8;
          float* red_sum = redx + 16 + half * 8;
          float* pv = redx + 32 + half * 512;
          int j = 0;
          for (int u = sub; u < nu; u += RH) {
            if (!active[u]) continue;  // uniform
            const bool mine = !halves || ((j & 1) == half);
            j++;
            const int T = flags[32 + u];
            const float* q = qkv + u * qw;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            const int t4 = gtid * 4;
            {
              const int rpc = rows_per_chunk_f16(hd, Tpad);
              for (int d0 = 0; d0 < hd; d0 += rpc) {
                const int nd = min(rpc, hd - d0);
                const __half* Kc = reinterpret_cast<const __half*>(ring_acquire(ring, cur));
                if (mine && t4 < Tpad) {
#pragma unroll 4
                  for (int d = 0; d < nd; d++) {
                    const uint2 uu = *reinterpret_cast<const uint2*>(Kc + d * Tpad + t4);
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&uu.x));
                    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&uu.y));
                    const float qd = q[d0 + d];
                    s0 = fmaf(qd, f0.x, s0); s1 = fmaf(qd, f0.y, s1);
                    s2 = fmaf(qd, f1.x, s2); s3 = fmaf(qd, f1.y, s3);
                  }
                }
                ring_release(ring, cur);
              }
            }
            mark(45);
            float inv = 0.f;
            if (mine) {
              float lmax = -INFINITY;
              if (t4 < Tpad) {
                s0 = (t4 + 0 < T) ? s0 * scale : -INFINITY;
                s1 = (t4 + 1 < T) ? s1 * scale : -INFINITY;
                s2 = (t4 + 2 < T) ? s2 * scale : -INFINITY;
                s3 = (t4 + 3 < T) ? s3 * scale : -INFINITY;
                lmax = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
              }
              lmax = warp_max(lmax);
              if (lane == 0) red_max[gwarp] = lmax;
              if (halves) { if (half == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory"); } else csync();
              float mx = red_max[0];
              for (int i = 1; i < gwarps; i++) mx = fmaxf(mx, red_max[i]);
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
              if (lane == 0) red_sum[gwarp] = lsum;
              if (halves) { if (half == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory"); } else csync();
              float tot = 0.f;
              for (int i = 0; i < gwarps; i++) tot += red_sum[i];
              inv = 1.0f / tot;
              if (p.xattn_out != nullptr && t4 < Tpad) {
                float* dst = p.xattn_out + (((((int64_t)(u0 + u) * p.L + l) * H + hh) * p.xattn_steps + p.step) * Tpad + t4);
                *reinterpret_cast<float4*>(dst) = make_float4(s0 * inv, s1 * inv, s2 * inv, s3 * inv);
              }
            }
            mark(46);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            const int g = gtid / tpr, dq = gtid - g * tpr;
            {
              const int rpc = rows_per_chunk_f16(Tpad, hd);
              for (int r0 = 0; r0 < Tpad; r0 += rpc) {
                const int nr = min(rpc, Tpad - r0);
                const __half* Vc = reinterpret_cast<const __half*>(ring_acquire(ring, cur));
                if (mine && g < Gv) {
                  const int tend = min(nr, T - r0);
#pragma unroll 4
                  for (int t = g; t < tend; t += Gv) {
                    const uint2 uu = *reinterpret_cast<const uint2*>(Vc + t * hd + dq * 4);
                    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&uu.x));
                    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&uu.y));
                    const float pt = ps[r0 + t];
                    a0 = fmaf(pt, f0.x, a0); a1 = fmaf(pt, f0.y, a1);
                    a2 = fmaf(pt, f1.x, a2); a3 = fmaf(pt, f1.y, a3);
                  }
                }
                ring_release(ring, cur);
              }
            }
            mark(47);
            if (mine) {
              if (g < Gv) *reinterpret_cast<float4*>(&pv[g * hd + dq * 4]) = make_float4(a0, a1, a2, a3);
              if (halves) { if (half == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory"); } else csync();
              if (gtid < hd) {
                float o = 0.f;
                for (int gg = 0; gg < Gv; gg++) o += pv[gg * hd + gtid];
                const float v = o * inv;
                const uint32_t a = attg_s + (uint32_t)(((hh * hd + gtid) * 8 + u) * 4);
                for (int r = 0; r < CS; r++) st_cluster(mapa(a, (uint32_t)r), v);
              }
            }
          }
          cluster_barrier(cbar, cpar, CS, err);
        } else if (stage == 4) {
          // silu(gate) * value: fc1 columns are interleaved (2j = value j, 2j+1 = gate j); result into the transposed
          // actT[j][u] the fc2 slice reads
          float vals[8];
          int cnt = 0;
          for (int i = threadIdx.x; i < nu * is && cnt < 8; i += kConsumers, cnt++) {
            const int u = i / is, jj = i - u * is;
            const float up = act[u * 2 * is + 2 * jj], gate = act[u * 2 * is + 2 * jj + 1];
            vals[cnt] = gate / (1.0f + expf(-gate)) * up;
          }
          cnt = 0;
          for (int i = threadIdx.x; i < nu * is && cnt < 8; i += kConsumers, cnt++) {
            const int u = i / is, jj = i - u * is;
            actT[jj * 8 + u] = vals[cnt];
          }
          csync();
        } else {
          // fc2 partials of every rank are in place: reduce this rank's output slice over the ranks, add bias and
          // residual, and hand the new residual slice to every rank
          cluster_barrier(cbar, cpar, CS, err);
          mark(48);
          for (int i = threadIdx.x; i < nu * ds; i += kConsumers) {
            const int u = i / ds, j = i - u * ds;
            const int n = rank * ds + j;
            float v = hres[u * D + n] + __ldg(w.b2 + n);
            const uint32_t a = part_s + (uint32_t)((u * D + n) * 4);
            for (int r = 0; r < CS; r++) v += ld_cluster(mapa(a, (uint32_t)r));
            const uint32_t ha = hres_s + (uint32_t)((u * D + n) * 4);
            for (int r = 0; r < CS; r++) st_cluster(mapa(ha, (uint32_t)r), v);
          }
          mark(49);
          cluster_barrier(cbar, cpar, CS, err);
        }
        mark(20 + stage);
      }
    }
    // ---- final LayerNorm rows of the cluster's utterances -> global (rank u % CS takes row u) ----
    if (warp == 0) {
      for (int u = rank; u < nu; u += CS) {
        if (!active[u]) continue;
        const float* h = hres + u * D;
        float2 v[kMaxPairs];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxPairs; i++) {
          const int k = 2 * lane + 64 * i;
          v[i] = (k < D) ? *reinterpret_cast<const float2*>(h + k) : make_float2(0.f, 0.f);
          s += v[i].x + v[i].y;
        }
        const float mean = warp_sum(s) / D;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < kMaxPairs; i++)
          if (2 * lane + 64 * i < D) {
            const float dx = v[i].x - mean, dy = v[i].y - mean;
            q += dx * dx + dy * dy;
          }
        const float rstd = rsqrtf(warp_sum(q) / D + 1e-5f);
#pragma unroll
        for (int i = 0; i < kMaxPairs; i++) {
          const int k = 2 * lane + 64 * i;
          if (k < D) *reinterpret_cast<float2*>(p.xfin + (int64_t)(u0 + u) * D + k) = make_float2((v[i].x - mean) * rstd, (v[i].y - mean) * rstd);
        }
      }
    }
  }
  // ---- the one grid-wide hand-off of the step: every CTA signals, every CTA waits for all of them ----
  unsigned* counters = p.sync3 + 32;
  csync();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counters), "r"(1u) : "memory");
    const unsigned target = (epoch + 1u) * (unsigned)G;
    const long long t0 = clock64();
    unsigned polls = 0;
    while ((int)(ld_acquire(counters) - target) < 0) {
      if ((++polls & 15u) == 0u) {
        if (poisoned(err)) break;
        if (clock64() - t0 > kSpinLimit) { atomicExch(err, 1u); break; }
      }
    }
  }
  csync();
  mark(30);

  // ============================== logits (tcgen05 + fused argmax), as in v3 ==============================
  {
    const int V = p.V, VC = p.vchunk;
    const int n_mt = (VC + 127) >> 7, nkb = D >> 5;
    const int nx = L.nxl;
    const int parity = p.step & 1;
    unsigned char* xp = smem_raw + L.xg;
    float* sbias = reinterpret_cast<float*>(smem_raw + L.sbias);
    unsigned char* rowflag = smem_raw + L.rowflag;
    int* ent_key = reinterpret_cast<int*>(smem_raw + L.ent_key);
    float* ent_val = reinterpret_cast<float*>(smem_raw + L.ent_val);
    int* ent_n = reinterpret_cast<int*>(smem_raw + L.ent_n);
    float* argv = reinterpret_cast<float*>(smem_raw + L.argv);
    int* argi = reinterpret_cast<int*>(smem_raw + L.argi);
    int acc_phase = 0;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nx >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#pragma unroll 1
    for (int b0 = 0; b0 < p.B; b0 += nx) {
      const int nb = min(nx, p.B - b0);
      planes_from_rows(xp, nx, p.xfin, D, b0, nb, D, nullptr, p.done, p.step);  // once per pass, every chunk of this CTA reuses them
#pragma unroll 1
      for (int item = (int)blockIdx.x; item < n_vjobs; item += G) {
        const bool biased = p.bias_static != nullptr || p.bias_dyn_n != nullptr;
        if (biased) {
          for (int i = threadIdx.x; i < 384; i += kConsumers) {
            const int v = item * VC + i;
            sbias[i] = (p.bias_static != nullptr && i < VC && v < V) ? __ldg(p.bias_static + v) : 0.f;
            rowflag[i] = 0;
          }
          if (threadIdx.x == 0) *ent_n = 0;
          csync();
          if (p.bias_dyn_n != nullptr) {
            const int cap = p.bias_dyn_cap;
            for (int idx = threadIdx.x; idx < nb * cap; idx += kConsumers) {
              const int b = idx / cap, k = idx - b * cap;
              if (k < p.bias_dyn_n[b0 + b]) {
                const int loc = p.bias_dyn_ids[(int64_t)(b0 + b) * cap + k] - item * VC;
                if (loc >= 0 && loc < VC) {
                  const int slot = atomicAdd(ent_n, 1);
                  if (slot >= kBiasEntries) atomicExch(err, 2u);
                  if (slot < kBiasEntries) {
                    ent_key[slot] = (loc << 9) | b;
                    ent_val[slot] = p.bias_dyn_val[(int64_t)(b0 + b) * cap + k];
                    rowflag[loc] = 1;
                  }
                }
              }
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        csync();
        const int warp_u = (int)uniform_u32((uint32_t)warp);
        int nchunks = 0;
        for (int mt = 0; mt < n_mt; mt++) nchunks += (nkb + 1) >> 1;
        if (warp_u == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t xb = uniform_u32(smem_u32(xp));
          const uint32_t tm = uniform_u32(tmem_base);
          for (int mt = 0; mt < n_mt; mt++) {
            const int R = min(128, VC - mt * 128);
            const uint32_t tmem_d = tm + (uint32_t)(mt * nx);
            for (int kb = 0; kb < nkb; kb += 2) {
              const int n = min(2, nkb - kb);
              mbar_wait(&ring.full[cur.st], cur.par, err);
              __syncwarp();
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t a_base = uniform_u32(smem_u32(ring.data + (size_t)cur.st * kStageBytes));
              const uint32_t empty_bar = uniform_u32(smem_u32(&ring.empty[cur.st]));
              if (elect_one()) {
                for (int q = 0; q < n; q++) {
                  const uint32_t a_hi = a_base + (uint32_t)(q * R * 128), a_lo = a_hi + (uint32_t)(R * 64);
                  const uint32_t b_hi = xb + (uint32_t)((kb + q) * nx * 128), b_lo = b_hi + (uint32_t)(nx * 64);
#pragma unroll
                  for (int jj = 0; jj < 2; jj++) {
                    const uint32_t ko = (uint32_t)jj * 32u;
                    const uint32_t first = (kb | q | jj) ? 1u : 0u;
                    umma_bf16(tmem_d, make_desc_sw64(a_lo + ko), make_desc_sw64(b_hi + ko), idesc, first);
                    umma_bf16(tmem_d, make_desc_sw64(a_hi + ko), make_desc_sw64(b_lo + ko), idesc, 1u);
                    umma_bf16(tmem_d, make_desc_sw64(a_hi + ko), make_desc_sw64(b_hi + ko), idesc, 1u);
                  }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty_bar) : "memory");
              }
              __syncwarp();
              cur_advance(cur, ring.ns);
            }
          }
          if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(acc_bar)) : "memory");
          __syncwarp();
        } else if (lane == 0) {
          for (int i = 0; i < nchunks; i++) {
            mbar_wait(&ring.full[cur.st], cur.par, err);
            mbar_arrive(&ring.empty[cur.st]);
            cur_advance(cur, ring.ns);
          }
        } else {
          for (int i = 0; i < nchunks; i++) cur_advance(cur, ring.ns);
        }
        __syncwarp();
        mbar_wait(acc_bar, (uint32_t)(acc_phase & 1), err);
        acc_phase++;
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // running best (key, index) of this warp per utterance column; warps 0-3 take m-tiles 0, 2, warps 4-7 m-tile 1
        for (int b = lane; b < 64; b += 32) { argv[warp * 64 + b] = 0.f; argi[warp * 64 + b] = 0x7fffffff; }
        __syncwarp();
        for (int mt_w = warp >> 2; mt_w < n_mt; mt_w += 2) {
        const int vrow = mt_w * 128 + (warp & 3) * 32 + lane;
        const int v = item * VC + vrow;
        const bool vok = (vrow < VC) && (v < V);
        for (int cb = 0; cb < nx; cb += 16) {
          uint32_t r[16];
          {
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt_w * nx + cb);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
          if (biased && vok) {
            const float sb = sbias[vrow];
#pragma unroll
            for (int e = 0; e < 16; e++) r[e] = __float_as_uint(__uint_as_float(r[e]) + sb);
            if (rowflag[vrow]) {
              const int n_ent = min(*ent_n, kBiasEntries);
              for (int s2 = 0; s2 < n_ent; s2++) {
                const int key = ent_key[s2];
                const int col = (key & 511) - cb;
                if ((key >> 9) == vrow && col >= 0 && col < 16) {
                  const float add = ent_val[s2];
#pragma unroll
                  for (int e = 0; e < 16; e++)
                    if (e == col) r[e] = __float_as_uint(__uint_as_float(r[e]) + add);
                }
              }
            }
          }
#pragma unroll
          for (int e = 0; e < 16; e++) {
            const int b = cb + e;
            const float val = __uint_as_float(r[e]);
            if (vok && p.logits_out && b < nb) p.logits_out[(int64_t)(b0 + b) * V + v] = val;
            uint32_t key = r[e];
            key = (key & 0x80000000u) ? ~key : (key | 0x80000000u);
            if (!vok || val != val) key = 0u;
            const uint32_t mx = __reduce_max_sync(0xffffffffu, key);
            const uint32_t who = __ballot_sync(0xffffffffu, key == mx);
            if (lane == 0 && mx != 0u) {
              const int idx = item * VC + mt_w * 128 + (warp & 3) * 32 + (__ffs(who) - 1);
              const uint32_t cur_key = __float_as_uint(argv[warp * 64 + b]);
              if (mx > cur_key || (mx == cur_key && idx < argi[warp * 64 + b])) {
                argv[warp * 64 + b] = __uint_as_float(mx);
                argi[warp * 64 + b] = idx;
              }
            }
          }
        }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        csync();
        if (threadIdx.x < nb) {
          const int b = threadIdx.x;
          uint32_t bk = 0u;
          int bi = 0x7fffffff;
          for (int w2 = 0; w2 < kWarpsC; w2++) {
            const uint32_t ok = __float_as_uint(argv[w2 * 64 + b]);
            const int oi = argi[w2 * 64 + b];
            if (ok > bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
          }
          const float bv = bk == 0u ? -INFINITY : __uint_as_float((bk & 0x80000000u) ? (bk & 0x7fffffffu) : ~bk);
          p.cand_val[((int64_t)parity * p.n_vchunk + item) * p.B + b0 + b] = bv;
          p.cand_idx[((int64_t)parity * p.n_vchunk + item) * p.B + b0 + b] = bi;
        }
        csync();
      }
    }
  }
  mark(31);
  // ---- end of step: CTA 0 waits for every CTA's candidates, counts what is left, opens the next epoch ----
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(counters + 32), "r"(1u) : "memory");
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      const unsigned target = (epoch + 1u) * (unsigned)G;
      const long long t0 = clock64();
      unsigned polls = 0;
      while ((int)(ld_acquire(counters + 32) - target) < 0) {
        if ((++polls & 15u) == 0u) {
          if (poisoned(err)) break;
          if (clock64() - t0 > kSpinLimit) { atomicExch(err, 1u); break; }
        }
      }
    }
    csync();
    int* cnt = flags;
    if (threadIdx.x == 0) *cnt = 0;
    csync();
    int n = 0;
    for (int b = threadIdx.x; b < p.B; b += kConsumers) n += __ldcg(p.done + b) ? 0 : 1;
    if (n) atomicAdd(cnt, n);
    csync();
    if (threadIdx.x == 0) {
      *p.n_active = *cnt;
      p.sync3[0] = epoch + 1u;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  csync();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

}  // namespace

// Cluster size (16, else 8) and how many such clusters are co-resident with this much shared memory per CTA.
// (On the B200s measured here: 7 clusters of 16 -- one GPC cannot host a 16-CTA cluster -- or 15 of 8.)
int decoder_step4_cluster_size(int device, int H, size_t smem_hint, int* clusters) {
  (void)device;
  auto kern = decoder_step4_kernel;
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_hint);
  *clusters = 0;
  for (int cs : {16, 8}) {
    if (cs % H != 0) continue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 4);
    cfg.blockDim = dim3(kThreads4);
    cfg.dynamicSmemBytes = smem_hint;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (std::getenv("MOONSHINE_B200_VERBOSE"))
      MSB_LOGF("decoder v4: cluster size %d, %zu B shared memory: %d co-resident clusters (%s)", cs, smem_hint, n, cudaGetErrorString(e));
    cudaGetLastError();
    if (e == cudaSuccess && n >= 4) {
      *clusters = n > 16 ? 16 : n;
      return cs;
    }
  }
  return 0;
}

bool decoder_step4_supported(const DecoderParams& p) {
  const int CS = p.c4_cs;
  if (CS != 8 && CS != 16) return false;
  if (p.layers[0].c4_wo == nullptr) return false;
  const int NC = p.c4_nc;
  if (NC < 1) return false;
  return p.D % 32 == 0 && p.D <= 64 * kMaxPairs && p.D % CS == 0 && p.I % CS == 0 && ((2 * p.I / CS) % 4) == 0 &&
         p.hd % 4 == 0 && p.hd <= 64 && CS % p.H == 0 && p.rot_dim <= 128 && p.Tpad <= 1024 && p.B <= kUmax * NC &&
         (3 * p.hd) % 4 == 0;
}

void decoder_step4_plan(DecoderParams& p) {
  p.c4_u = (p.B + p.c4_nc - 1) / p.c4_nc;
  if (p.c4_u == 7) p.c4_u = 8;  // dense-routine instances: 1..6 and 8 rows
}

size_t decoder_step4_smem_bytes(const DecoderParams& p) {
  const Smem4 L = smem_layout4(p.B, p.D, p.hd, p.I, p.c4_cs, p.c4_u, p.Tpad, p.Smax, p.smem_limit);
  if (L.ns < 2) throw std::runtime_error("decoder v4: not enough shared memory for the operand ring");
  return (size_t)L.total;
}

void launch_decoder_step4(const DecoderParams& p, cudaStream_t stream) {
  const size_t smem = decoder_step4_smem_bytes(p);
  auto kern = decoder_step4_kernel;
  static SmemAttrCache cache;
  if (cache.needs(smem)) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.c4_cs * p.c4_nc);
  cfg.blockDim = dim3(kThreads4);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  // Cooperative as well as clustered: the one grid-wide hand-off before the vocabulary projection needs every
  // cluster resident, and two handles (or two contexts of one handle) may launch on the same device at once.
  // (MOONSHINE_B200_V4_COOP=0: Nsight Compute cannot replay a launch that is both clustered and cooperative)
  static std::atomic<int> coop{[] { const char* e = std::getenv("MOONSHINE_B200_V4_COOP"); return (e && e[0] == '0') ? 0 : 1; }()};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.c4_cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = coop.load() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess && cfg.numAttrs == 2) {
    MSB_LOGF("decoder v4: cooperative cluster launch refused (%s); launching clustered only", cudaGetErrorString(e));
    cudaGetLastError();
    coop.store(0);
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, p);
  }
  CUDA_CHECK(e);
}

}  // namespace msb
