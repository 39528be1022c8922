// Fused encoder self-attention on tcgen05 + TMEM:  O = softmax(scale * Q K^T [+ window]) V
// for one 128-query tile of one (utterance, head) per CTA, fp32 in / fp32 out, bf16x3 split on both
// contractions.  Replaces three launches (QK^T GEMM, row softmax, PV GEMM) and the
// 4 x B*H*T*T fp32 score traffic through HBM between them: the score tile never leaves the SM.
//
//   1. Q tile [128 x hd] and the K rows of the key range [klo, khi) are read as fp32, split to bf16 hi / lo and
//      stored K-major SWIZZLE_64B; one thread issues tcgen05.mma (M=128, N<=256 chunks) into TMEM columns
//      [0, Tk): the WHOLE score row of every query lives in TMEM (Tk <= 448 columns).
//   2. softmax straight out of TMEM: thread = (query row, column half).  Pass 1 row max, pass 2
//      exp(scale * s - max) -- written as bf16 hi / lo A-operand planes in the shared memory K no longer needs --
//      and row sums.  Exact two-pass softmax, no online rescaling.
//   3. P V per round of <= 224 keys: V^T rows (d) x keys are the K-major B operand; accumulators in TMEM columns
//      [448, 512).  Epilogue divides by the row sum and writes the head's slice of the attention output.
// Classic encoder: key range = the whole utterance (T <= 448, i.e. clips up to 10.8 s; longer clips take the
// unfused path).  Streaming encoder: key range = [m0 - past, m0 + 127 + future], so only the band is computed.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.h"
#include "kernels.h"

namespace msb {

namespace {

constexpr int kAttnThreads = 256;
constexpr int kMaxTk = 448;           // score columns in TMEM
constexpr int kOCol = 448;            // O accumulator columns [448, 512)
constexpr int kRoundKeys = 224;       // keys per P*V round (7 k-blocks of 32)
constexpr int kQBytes = 2 * 2 * 128 * 64;              // 2 k-blocks x (hi | lo) x 128 rows x 64 B = 32 KB
constexpr int kKPBytes = 2 * 2 * kMaxTk * 64;          // K planes (2 k-blocks) == P planes of one round = 112 KB
constexpr int kVBytes = (kRoundKeys / 32) * 2 * 64 * 64;  // 7 k-blocks x (hi | lo) x 64 rows x 64 B = 56 KB
constexpr long long kSpinLimit = 4000000000LL;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ unsigned int g_attention_error = 0;  // watchdog flag (see attention_tc_error_async)
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
    if (clock64() - t0 > kSpinLimit) {  // poison instead of trapping the context: the host raises after its next synchronise
      atomicExch(&g_attention_error, 1u);
      break;
    }
  }
}
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;  // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ uint32_t idesc_for(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a converged warp (tcgen05.mma wants uniform-register operands: issued under elect.sync inside
// warp-uniform control flow it is a few instructions; from a divergent `if (tid == X)` the compiler wraps every
// MMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop, ~95 ns per MMA measured on B200).
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

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 8 floats -> one 16-byte chunk of bf16 hi and one of bf16 lo
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 f = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// byte offset of 16-byte chunk `cc` (0..3) of row `r` inside a [rows x 64 B] SWIZZLE_64B plane
__device__ __forceinline__ uint32_t sw64(int r, int cc) { return (uint32_t)r * 64u + (uint32_t)((cc ^ ((r >> 1) & 3)) << 4); }

// rows x hd fp32 matrix (row stride ld) -> k-blocks of [hi rows x 64 B | lo rows x 64 B]; rows >= valid and
// columns >= hd are zero.  Every 8-column group must start 16-byte aligned (hd % 4 == 0, ld % 4 == 0).
__device__ __forceinline__ void stage_rows(const float* __restrict__ src, int64_t ld, int valid, int rows, int hd,
                                           int nkb, unsigned char* dst) {
  const int chunks = nkb * 4;
  const int total = rows * chunks;
  constexpr int U = 8;  // independent loads in flight per thread
  for (int i0 = threadIdx.x; i0 < total; i0 += U * kAttnThreads) {
    float4 a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int i = i0 + u * kAttnThreads;
      const int r = i / chunks, c8 = i - r * chunks;
      a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      b[u] = a[u];
      if (i < total && r < valid) {
        const float* p = src + (int64_t)r * ld + c8 * 8;
        if (c8 * 8 < hd) a[u] = *reinterpret_cast<const float4*>(p);
        if (c8 * 8 + 4 < hd) b[u] = *reinterpret_cast<const float4*>(p + 4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int i = i0 + u * kAttnThreads;
      if (i >= total) break;
      const int r = i / chunks, c8 = i - r * chunks;
      const float v[8] = {a[u].x, a[u].y, a[u].z, a[u].w, b[u].x, b[u].y, b[u].z, b[u].w};
      uint4 hi, lo;
      split8(v, hi, lo);
      const int kb = c8 >> 2, cc = c8 & 3;
      unsigned char* base = dst + (size_t)kb * 2 * rows * 64;
      *reinterpret_cast<uint4*>(base + sw64(r, cc)) = hi;
      *reinterpret_cast<uint4*>(base + (size_t)rows * 64 + sw64(r, cc)) = lo;
    }
  }
}

__global__ void __launch_bounds__(kAttnThreads, 1) attention_tc_kernel(const __grid_constant__ AttnParams p) {
  const int z = blockIdx.y;
  const int T = p.Tz[z];
  const int m0 = blockIdx.x * 128;
  if (m0 >= T) return;  // uniform per CTA, before any allocation
  const int hd = p.hd;
  const int nkb = (hd + 31) >> 5;       // 32-wide k-blocks of the head dimension
  const int ksteps = (hd + 15) >> 4;    // k16 MMA steps that hold data
  int klo = 0, khi = T;
  if (p.win_past >= 0) {
    klo = max(0, m0 - p.win_past);
    khi = min(T, m0 + 128 + p.win_future);
  }
  const int nk = khi - klo;
  const int Tk = (nk + 15) & ~15;       // <= kMaxTk (host guarantees)

  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* Qs = smem;                       // [kb][hi | lo][128 x 64 B]
  unsigned char* KP = Qs + kQBytes;               // K: [kb][hi | lo][Tk x 64 B];  P round: [kblock][hi | lo][128 x 64 B]
  unsigned char* Vs = KP + kKPBytes;              // [kblock][hi | lo][64 x 64 B]
  __shared__ __align__(8) uint64_t mma_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float red_max[2][128];
  __shared__ float red_sum[2][128];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&mma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- stage Q and K ----
  const float* Q = p.qk + p.offQ[z] + (int64_t)m0 * p.ldqk;
  const float* K = p.qk + p.offK[z] + (int64_t)klo * p.ldqk;
  stage_rows(Q, p.ldqk, min(128, T - m0), 128, hd, nkb, Qs);
  stage_rows(K, p.ldqk, nk, Tk, hd, nkb, KP);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_smem;
  uint32_t bar_phase = 0;

  // ---- S = Q K^T into TMEM columns [0, Tk) ----  (warp 0, one elected lane, operands made provably uniform)
  const int warp_u = (int)uniform_u32(threadIdx.x >> 5);
  const uint32_t tmem_u = uniform_u32(tmem);
  const uint32_t qs_u = uniform_u32(smem_u32(Qs)), kp_u = uniform_u32(smem_u32(KP));
  if (warp_u == 0) {
    if (elect_one()) {
      for (int n0 = 0; n0 < Tk; n0 += 256) {
        const int nc = min(256, Tk - n0);
        const uint32_t idesc = idesc_for(nc);
        for (int ks = 0; ks < ksteps; ks++) {
          const int kb = ks >> 1;
          const uint32_t ko = (uint32_t)(ks & 1) * 32u;
          const uint32_t a_hi = qs_u + (uint32_t)(kb * 2 * 128 * 64) + ko, a_lo = a_hi + 128 * 64;
          const uint32_t b_hi = kp_u + (uint32_t)(kb * 2 * Tk * 64) + (uint32_t)n0 * 64u + ko;
          const uint32_t b_lo = b_hi + (uint32_t)Tk * 64u;
          umma_bf16(tmem_u + n0, make_desc_sw64(a_lo), make_desc_sw64(b_hi), idesc, ks ? 1u : 0u);
          umma_bf16(tmem_u + n0, make_desc_sw64(a_hi), make_desc_sw64(b_lo), idesc, 1u);
          umma_bf16(tmem_u + n0, make_desc_sw64(a_hi), make_desc_sw64(b_hi), idesc, 1u);
        }
      }
      umma_commit(&mma_bar);
    }
    __syncwarp();
  }
  mbar_wait(&mma_bar, bar_phase & 1);
  bar_phase++;
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- softmax: thread = (row r, column half) ----
  const int r = (warp & 3) * 32 + lane;
  const int half = warp >> 2;
  const int m = m0 + r;                                  // query position
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  int vlo = 0, vhi = nk;                                 // valid column range of this row
  if (p.win_past >= 0) {
    vlo = max(0, m - p.win_past - klo);
    vhi = min(nk, m + p.win_future + 1 - klo);
  }
  const int ngroups = Tk >> 4;
  const int g_lo = half ? (ngroups + 1) / 2 : 0, g_hi = half ? ngroups : (ngroups + 1) / 2;
  float mx = -INFINITY;
  for (int g = g_lo; g < g_hi; g += 2) {
    uint32_t v0[16], v1[16];
    const bool two = g + 1 < g_hi;  // warp-uniform
    tmem_ld16(lane_addr + (uint32_t)(g * 16), v0);
    if (two) tmem_ld16(lane_addr + (uint32_t)(g * 16 + 16), v1);
    tmem_ld_wait();
    // raw maxima (scale > 0 is applied once at the end); groups entirely inside the row's valid range skip
    // the per-column bounds checks
    const int j0 = g * 16;
    if (j0 >= vlo && j0 + 16 <= vhi) {
#pragma unroll
      for (int e = 0; e < 16; e++) mx = fmaxf(mx, __uint_as_float(v0[e]));
    } else {
#pragma unroll
      for (int e = 0; e < 16; e++)
        if (j0 + e >= vlo && j0 + e < vhi) mx = fmaxf(mx, __uint_as_float(v0[e]));
    }
    if (two) {
      if (j0 + 16 >= vlo && j0 + 32 <= vhi) {
#pragma unroll
        for (int e = 0; e < 16; e++) mx = fmaxf(mx, __uint_as_float(v1[e]));
      } else {
#pragma unroll
        for (int e = 0; e < 16; e++)
          if (j0 + 16 + e >= vlo && j0 + 16 + e < vhi) mx = fmaxf(mx, __uint_as_float(v1[e]));
      }
    }
  }
  red_max[half][r] = mx;
  __syncthreads();
  mx = fmaxf(red_max[0][r], red_max[1][r]);
  if (mx == -INFINITY) mx = 0.f;  // rows past the utterance end: every probability is 0
  // exp(scale * (s - max)) as one FMA + ex2: exp2(s * c - max * c), c = scale * log2(e)
  const float c2 = p.scale * 1.4426950408889634f;
  const float mxc = mx * c2;

  // ---- P V in rounds of <= 224 keys ----
  float sum = 0.f;
  const int NV = 64;  // B-operand rows of V^T (head dims, zero-padded)
  const uint32_t idesc_pv = idesc_for(NV);
  const float* V = p.vt + p.offV[z];
  const bool v_vec = ((klo & 3) == 0) && ((p.ldv & 3) == 0);
  for (int i = threadIdx.x * 16; i < kVBytes; i += kAttnThreads * 16)  // padding rows d >= hd stay zero
    *reinterpret_cast<uint4*>(Vs + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int k0 = 0, round = 0; k0 < Tk; k0 += kRoundKeys, round++) {
    const int rk = min(kRoundKeys, Tk - k0);        // keys this round (multiple of 16)
    const int rkb = (rk + 31) >> 5;                 // k-blocks this round
    // P planes for columns [k0, k0 + rk): this thread's half of the round's 16-column groups
    {
      const int rg = rk >> 4;
      const int h_lo = half ? (rg + 1) / 2 : 0, h_hi = half ? rg : (rg + 1) / 2;
      for (int g = h_lo; g < h_hi; g++) {
        uint32_t v[16];
        tmem_ld16(lane_addr + (uint32_t)(k0 + g * 16), v);
        tmem_ld_wait();
        float e[16];
        const int j0 = k0 + g * 16;
        if (j0 >= vlo && j0 + 16 <= vhi) {
#pragma unroll
          for (int q = 0; q < 16; q++) {
            e[q] = exp2f(fmaf(__uint_as_float(v[q]), c2, -mxc));
            sum += e[q];
          }
        } else {
#pragma unroll
          for (int q = 0; q < 16; q++) {
            const int j = j0 + q;
            e[q] = (j >= vlo && j < vhi) ? exp2f(fmaf(__uint_as_float(v[q]), c2, -mxc)) : 0.f;
            sum += e[q];
          }
        }
        const int kblock = g >> 1;                   // 32 keys per block, 16 per group
        unsigned char* base = KP + (size_t)kblock * 2 * 128 * 64;
#pragma unroll
        for (int c = 0; c < 2; c++) {
          float w[8];
#pragma unroll
          for (int q = 0; q < 8; q++) w[q] = e[c * 8 + q];
          uint4 hi, lo;
          split8(w, hi, lo);
          const int cc = (g & 1) * 2 + c;
          *reinterpret_cast<uint4*>(base + sw64(r, cc)) = hi;
          *reinterpret_cast<uint4*>(base + 128 * 64 + sw64(r, cc)) = lo;
        }
      }
      if ((rk & 31) && half == 1) {  // odd number of 16-groups: zero the unused half of the last k-block
        unsigned char* base = KP + (size_t)(rkb - 1) * 2 * 128 * 64;
        const uint4 zz = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int cc = 2; cc < 4; cc++) {
          *reinterpret_cast<uint4*>(base + sw64(r, cc)) = zz;
          *reinterpret_cast<uint4*>(base + 128 * 64 + sw64(r, cc)) = zz;
        }
      }
    }
    // V^T planes: rows d (0..63), columns = keys of the round
    {
      // only the hd real rows are staged per round; rows [hd, 64) were zeroed once before the loop
      const int chunks = rkb * 4;
      const int total = hd * chunks;
      constexpr int U = 4;
      for (int i0 = threadIdx.x; i0 < total; i0 += U * kAttnThreads) {
        float v[U][8];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int i = i0 + u * kAttnThreads;
          const int d = i / chunks, c8 = i - d * chunks;
#pragma unroll
          for (int e = 0; e < 8; e++) v[u][e] = 0.f;
          if (i < total) {
            const int j0 = k0 + c8 * 8;                // column (key - klo) of the first element
            const float* src = V + (int64_t)d * p.ldv + klo + j0;
            if (v_vec && j0 + 8 <= nk) {
              const float4 a = *reinterpret_cast<const float4*>(src);
              const float4 b = *reinterpret_cast<const float4*>(src + 4);
              v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
              v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
            } else {
#pragma unroll
              for (int e = 0; e < 8; e++)
                if (j0 + e < nk) v[u][e] = src[e];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          const int i = i0 + u * kAttnThreads;
          if (i >= total) break;
          const int d = i / chunks, c8 = i - d * chunks;
          uint4 hi, lo;
          split8(v[u], hi, lo);
          const int kblock = c8 >> 2, cc = c8 & 3;
          unsigned char* base = Vs + (size_t)kblock * 2 * NV * 64;
          *reinterpret_cast<uint4*>(base + sw64(d, cc)) = hi;
          *reinterpret_cast<uint4*>(base + NV * 64 + sw64(d, cc)) = lo;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp_u == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t vs_u = uniform_u32(smem_u32(Vs));
      const int steps = (int)uniform_u32((uint32_t)(rk >> 4));  // k16 steps with data
      const uint32_t round_u = uniform_u32((uint32_t)round);
      if (elect_one()) {
        for (int s = 0; s < steps; s++) {
          const int kblock = s >> 1;
          const uint32_t ko = (uint32_t)(s & 1) * 32u;
          const uint32_t a_hi = kp_u + (uint32_t)(kblock * 2 * 128 * 64) + ko, a_lo = a_hi + 128 * 64;
          const uint32_t b_hi = vs_u + (uint32_t)(kblock * 2 * NV * 64) + ko, b_lo = b_hi + NV * 64;
          const uint32_t acc = (round_u | (uint32_t)s) ? 1u : 0u;
          umma_bf16(tmem_u + kOCol, make_desc_sw64(a_lo), make_desc_sw64(b_hi), idesc_pv, acc);
          umma_bf16(tmem_u + kOCol, make_desc_sw64(a_hi), make_desc_sw64(b_lo), idesc_pv, 1u);
          umma_bf16(tmem_u + kOCol, make_desc_sw64(a_hi), make_desc_sw64(b_hi), idesc_pv, 1u);
        }
        umma_commit(&mma_bar);
      }
      __syncwarp();
    }
    // the next round overwrites P / V: wait until the tensor core has read them
    mbar_wait(&mma_bar, bar_phase & 1);
    bar_phase++;
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // ---- epilogue: O / sum -> out[m][h * hd + d] ----
  red_sum[half][r] = sum;
  __syncthreads();
  const float inv = 1.0f / (red_sum[0][r] + red_sum[1][r]);
  {
    float* dst = p.out + p.offO[z] + (int64_t)m * p.ldo;
#pragma unroll
    for (int g = 0; g < 2; g++) {
      const int c0 = half * 32 + g * 16;
      if (c0 < hd) {  // warp-uniform: the TMEM load is executed by the whole warp
        uint32_t v[16];
        tmem_ld16(lane_addr + (uint32_t)(kOCol + c0), v);
        tmem_ld_wait();
        if (m < T) {
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            if (c0 + q < hd) {
              *reinterpret_cast<float4*>(dst + c0 + q) =
                  make_float4(__uint_as_float(v[q]) * inv, __uint_as_float(v[q + 1]) * inv,
                              __uint_as_float(v[q + 2]) * inv, __uint_as_float(v[q + 3]) * inv);
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

}  // namespace

bool attention_tc_supported(int max_t, int hd, int win_past, int win_future) {
  if (hd > 64 || (hd & 3)) return false;
  const int span = win_past >= 0 ? std::min(max_t, 128 + win_past + win_future) : max_t;
  return ((span + 15) & ~15) <= kMaxTk;
}

void launch_attention_tc(const AttnParams& p, int groups, int max_t, cudaStream_t stream) {
  if (groups == 0 || max_t == 0) return;
  static SmemAttrCache cache;
  const size_t smem = (size_t)kQBytes + kKPBytes + kVBytes + 1024;
  if (cache.needs(smem))
    CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((max_t + 127) / 128, groups);
  attention_tc_kernel<<<grid, kAttnThreads, smem, stream>>>(p);
}

void attention_tc_error_async(unsigned int* pinned_dst, cudaStream_t stream) {
  CUDA_CHECK(cudaMemcpyFromSymbolAsync(pinned_dst, g_attention_error, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, stream));
}
void attention_tc_clear_error(cudaStream_t stream) {
  const unsigned int zero = 0;
  CUDA_CHECK(cudaMemcpyToSymbolAsync(g_attention_error, &zero, sizeof(unsigned int), 0, cudaMemcpyHostToDevice, stream));
}

}  // namespace msb
