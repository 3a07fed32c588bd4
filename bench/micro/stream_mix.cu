// stream_mix.cu — what HBM delivers for the read:write MIX of each fused half-step, without any
// stencil: nr read streams and nw write streams of n doubles each, one element of every stream per
// thread per step, grid-stride.  The plain copy (1 read : 1 write) is what MEASURED_PEAKS.json
// quotes; the B half-step of the fast path is 6 reads : 3 writes, the D/E half-step 9 : 6.
// Build and run on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a stream_mix.cu -o stream_mix
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

template <int NR, int NW> struct Ptrs {
  const double *r[NR];
  double *w[NW];
};

template <int NR, int NW, int UNROLL>
__global__ void __launch_bounds__(256) mix_kernel(const __grid_constant__ Ptrs<NR, NW> P, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * UNROLL) {
    double v[UNROLL][NR];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int k = 0; k < NR; ++k)
        v[u][k] = (i + u * stride < n) ? __ldg(P.r[k] + i + u * stride) : 0.0;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      double s = 0;
#pragma unroll
      for (int k = 0; k < NR; ++k)
        s += v[u][k];
#pragma unroll
      for (int k = 0; k < NW; ++k)
        if (i + u * stride < n) P.w[k][i + u * stride] = s + k;
    }
  }
}

template <int NR, int NW, int UNROLL> static void run(size_t n, int ctas_per_sm) {
  Ptrs<NR, NW> P;
  for (int k = 0; k < NR; ++k) {
    cudaMalloc((void **)&P.r[k], n * 8);
    cudaMemset((void *)P.r[k], 0, n * 8);
  }
  for (int k = 0; k < NW; ++k) cudaMalloc((void **)&P.w[k], n * 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int grid = 148 * ctas_per_sm;
  float best = 1e30f;
  for (int it = 0; it < 8; ++it) {
    cudaEventRecord(a);
    mix_kernel<NR, NW, UNROLL><<<grid, 256>>>(P, n);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it >= 2 && ms < best) best = ms;
  }
  printf("{\"reads\": %d, \"writes\": %d, \"unroll\": %d, \"ctas_per_sm\": %d, \"n\": %zu, \"ms\": %.4f, \"GBps\": %.1f}\n", NR, NW,
         UNROLL, ctas_per_sm, n, best, (NR + NW) * 8.0 * n / best / 1e6);
  for (int k = 0; k < NR; ++k) cudaFree((void *)P.r[k]);
  for (int k = 0; k < NW; ++k) cudaFree(P.w[k]);
}

int main(int argc, char **argv) {
  const size_t n = argc > 1 ? (size_t)atoll(argv[1]) : (size_t)120 * 1000 * 1000;
  for (int c : {4, 8}) {
    run<1, 1, 4>(n, c);
    run<2, 1, 4>(n, c);
    run<6, 3, 2>(n, c);
    run<6, 3, 4>(n, c);
    run<9, 6, 1>(n, c);
    run<9, 6, 2>(n, c);
    run<12, 6, 1>(n, c);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "%s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
