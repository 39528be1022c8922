// Microbenchmark: how fast does ONE CTA per SM execute straight-line code that does not fit the instruction caches?
// (The persistent decoder kernel runs each phase's code once per layer: if cold code is fetch-bound, code size is the
// first-order cost.)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu && ./icache
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void __launch_bounds__(320, 1) body(float* out, int iters, long long* cycles) {
  float a = threadIdx.x * 1e-3f, b = a + 1.f, c = a + 2.f, d = a + 3.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < N; i++) {  // 4 independent chains: issue-bound when the code is resident
      a = fmaf(a, 1.000001f, 0.5f + i);
      b = fmaf(b, 0.999999f, 0.25f + i);
      c = fmaf(c, 1.000002f, 0.125f + i);
      d = fmaf(d, 0.999998f, 0.0625f + i);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
}

template <int N>
void run(int threads, int iters) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 320 * 4);
  cudaMalloc(&cyc, 8);
  for (int rep = 0; rep < 3; rep++) body<N><<<148, threads>>>(out, iters, cyc);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double instr = 4.0 * N * iters;
  printf("body %6d instr (%4d KB)  threads %3d  iters %4d : %.2f cycles/instr per warp  (%.1f us per pass at 1.9 GHz)\n", 4 * N,
         4 * N * 16 / 1024, threads, iters, (double)h / instr, (double)h / iters / 1900.0);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int threads : {32, 256}) {
    run<2048>(threads, 12);     // 128 KB
    run<2560>(threads, 10);     // 160 KB
    run<3072>(threads, 8);      // 192 KB
    run<3584>(threads, 8);      // 224 KB
    run<4096>(threads, 6);      // 256 KB
    run<5120>(threads, 6);      // 320 KB
    run<6144>(threads, 4);      // 384 KB
    run<8192>(threads, 4);      // 512 KB
  }
  return 0;
}
