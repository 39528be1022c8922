// Microbenchmark behind DESIGN.md's statement on the decoder's operand ring: how fast can ONE SM pull
// data through `cp.async.bulk` + mbarrier stages, as a function of stage size, stage count and the number of
// bulk copies a stage is split into; source either private to the CTA (HBM stream) or shared by all CTAs
// (L2 hits, like the layer weights that every batch tile re-reads).  Not on the product path.
#include <cstdint>

#include "common.h"
#include "kernels.h"

namespace msb {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__global__ void ring_bw_kernel(const unsigned char* src, int64_t bytes_per_cta, int stage_bytes, int stages, int nsub,
                               int shared_src, int64_t shared_span) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 32;
  unsigned char* data = smem + 1024;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t nchunks = bytes_per_cta / stage_bytes;
  const unsigned char* base = shared_src ? src : src + (int64_t)blockIdx.x * bytes_per_cta;
  // nsub < 0: |nsub| producer warps, chunk i issued by warp i % P (one whole-stage copy each)
  const int P = nsub < 0 ? -nsub : 1;
  if (nsub < 0) nsub = 1;
  const int consumer_thread = 32 * P;
  if ((threadIdx.x & 31) == 0 && (int)(threadIdx.x >> 5) < P) {  // producers
    const int sub = stage_bytes / nsub;
    for (int64_t i = threadIdx.x >> 5; i < nchunks; i += P) {
      const int s = (int)(i % stages);
      mbar_wait(&empty[s], (uint32_t)(((i / stages) & 1) ^ 1));
      mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
      const int64_t off = shared_src ? (i * stage_bytes) % shared_span : i * stage_bytes;
      for (int q = 0; q < nsub; q++)
        bulk_g2s(data + (size_t)s * stage_bytes + (size_t)q * sub, base + off + (int64_t)q * sub, (uint32_t)sub, &full[s]);
    }
  } else if ((int)threadIdx.x == consumer_thread) {  // consumer: releases a stage as soon as it has landed
    for (int64_t i = 0; i < nchunks; i++) {
      const int s = (int)(i % stages);
      mbar_wait(&full[s], (uint32_t)((i / stages) & 1));
      mbar_arrive(&empty[s]);
    }
  }
}

}  // namespace

float ring_bandwidth_test(int64_t bytes_per_cta, int stage_bytes, int stages, int nsub, int shared_src, int grid) {
  if (stages > 32 || stage_bytes * stages + 1024 > 220 * 1024 || (nsub > 0 && stage_bytes % (16 * nsub)) || nsub < -8 || nsub == 0) return -1.f;
  const int64_t shared_span = 2 << 20;
  const size_t total = shared_src ? (size_t)shared_span : (size_t)bytes_per_cta * grid;
  unsigned char* buf = nullptr;
  CUDA_CHECK(cudaMalloc(&buf, total));
  CUDA_CHECK(cudaMemset(buf, 1, total));
  const size_t smem = (size_t)stage_bytes * stages + 1024;
  CUDA_CHECK(cudaFuncSetAttribute(ring_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  ring_bw_kernel<<<grid, 320, smem>>>(buf, bytes_per_cta, stage_bytes, stages, nsub, shared_src, shared_span);
  CUDA_CHECK(cudaEventRecord(e0));
  ring_bw_kernel<<<grid, 320, smem>>>(buf, bytes_per_cta, stage_bytes, stages, nsub, shared_src, shared_span);
  CUDA_CHECK(cudaEventRecord(e1));
  CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  return ms;
}

}  // namespace msb
