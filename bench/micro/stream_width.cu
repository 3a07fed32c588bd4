// stream_width.cu — does the width of a thread's access bound the HBM rate of a many-stream kernel?
// The same read:write mixes as stream_mix.cu (6:3 = B half-step, 9:6 = D/E half-step), one element of
// every stream per thread per step, with the element 4, 8 or 16 bytes wide (float, float2 = double,
// float4 = double2): a warp request of 128, 256 or 512 bytes.  If the single-precision fast path
// (128-byte requests, 69 % of the roofline) is limited by requests in flight rather than by bytes, the
// float rows of this table stop short of the float2 / float4 rows at the same occupancy.
// Build and run on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a stream_width.cu -o stream_width
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

template <int NR, int NW, typename E> struct Ptrs {
  const E *r[NR];
  E *w[NW];
};
__device__ __forceinline__ float sum(float v) { return v; }
__device__ __forceinline__ float sum(float2 v) { return v.x + v.y; }
__device__ __forceinline__ float sum(float4 v) { return v.x + v.y + v.z + v.w; }
__device__ __forceinline__ void splat(float &o, float s) { o = s; }
__device__ __forceinline__ void splat(float2 &o, float s) { o = make_float2(s, s); }
__device__ __forceinline__ void splat(float4 &o, float s) { o = make_float4(s, s, s, s); }

template <int NR, int NW, int UNROLL, typename E, int MINB>
__global__ void __launch_bounds__(256, MINB) mix_kernel(const __grid_constant__ Ptrs<NR, NW, E> P, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * UNROLL) {
    E v[UNROLL][NR];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int k = 0; k < NR; ++k)
        if (i + u * stride < n) v[u][k] = __ldg(P.r[k] + i + u * stride);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float s = 0;
#pragma unroll
      for (int k = 0; k < NR; ++k)
        s += sum(v[u][k]);
      E o;
      splat(o, s);
#pragma unroll
      for (int k = 0; k < NW; ++k)
        if (i + u * stride < n) P.w[k][i + u * stride] = o;
    }
  }
}

template <int NR, int NW, int UNROLL, typename E, int MINB> static void run(size_t bytes_per_stream, const char *ename) {
  const size_t n = bytes_per_stream / sizeof(E);
  Ptrs<NR, NW, E> P;
  for (int k = 0; k < NR; ++k) {
    cudaMalloc((void **)&P.r[k], n * sizeof(E));
    cudaMemset((void *)P.r[k], 0, n * sizeof(E));
  }
  for (int k = 0; k < NW; ++k) cudaMalloc((void **)&P.w[k], n * sizeof(E));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int grid = 148 * MINB;
  float best = 1e30f;
  for (int it = 0; it < 8; ++it) {
    cudaEventRecord(a);
    mix_kernel<NR, NW, UNROLL, E, MINB><<<grid, 256>>>(P, n);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it >= 2 && ms < best) best = ms;
  }
  printf("{\"elem\": \"%s\", \"bytes_per_thread_access\": %d, \"reads\": %d, \"writes\": %d, \"unroll\": %d, \"ctas_per_sm\": %d, "
         "\"loads_in_flight_per_thread\": %d, \"ms\": %.4f, \"GBps\": %.1f}\n",
         ename, (int)sizeof(E), NR, NW, UNROLL, MINB, NR * UNROLL, best, (NR + NW) * (double)sizeof(E) * n / best / 1e6);
  fflush(stdout);
  for (int k = 0; k < NR; ++k) cudaFree((void *)P.r[k]);
  for (int k = 0; k < NW; ++k) cudaFree(P.w[k]);
}

template <int NR, int NW> static void sweep(size_t bytes) {
  run<NR, NW, 1, float, 4>(bytes, "float");
  run<NR, NW, 2, float, 4>(bytes, "float");
  run<NR, NW, 1, float, 8>(bytes, "float");
  run<NR, NW, 2, float, 8>(bytes, "float");
  run<NR, NW, 1, float2, 4>(bytes, "float2");
  run<NR, NW, 2, float2, 4>(bytes, "float2");
  run<NR, NW, 1, float2, 8>(bytes, "float2");
  run<NR, NW, 1, float4, 4>(bytes, "float4");
  run<NR, NW, 1, float4, 8>(bytes, "float4");
}

int main(int argc, char **argv) {
  const size_t bytes = argc > 1 ? (size_t)atoll(argv[1]) : (size_t)480 * 1000 * 1000;
  sweep<6, 3>(bytes);
  sweep<9, 6>(bytes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "%s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
