// pml_shapes.cu — where the PML half-step kernel loses its time: one chunk shape at a time.
//
// The PML chunks of BASELINE configs[1] (512^3, PML 1.0 at resolution 10) are thin slabs: faces of
// 10 x 492 x 492 cells normal to x, y or z, edges of 10 x 10 x 492, corners of 10^3.  bench.py times
// them all in one launch; this harness builds the job descriptor of ONE such chunk (same component
// variants as fields_chunk::step_db / update_eh emit: PML in dsig = d_c + 1, f_u level for
// dsigu = d_c + 2, f_w ODE for dsigw = d_c) and times the product kernels on it with CUDA events,
// reporting algorithmic GB/s per shape, kernel form, planes per CTA and CTAs per SM.
//
// Build (from the repo root) and run on the GPU box:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Iinclude -Imeep_b200/csrc \
//        bench/micro/pml_shapes.cu -o gpurun_out/pml_shapes && gpurun_out/pml_shapes
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "meep_b200.h"
#include "fused.cuh"

using namespace mb200;

#define CK(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) {                                                                       \
      fprintf(stderr, "%s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));                  \
      exit(1);                                                                                     \
    }                                                                                              \
  } while (0)

__global__ void fill_kernel(double *p, size_t n, double v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = v + 1e-9 * (double)(i & 1023);
}

static std::vector<void *> g_allocs;
static double *dalloc(size_t n, double v) {
  double *p;
  CK(cudaMalloc((void **)&p, n * sizeof(double)));
  fill_kernel<<<1024, 256>>>(p, n, v);
  g_allocs.push_back(p);
  return p;
}
static void free_all() {
  for (void *p : g_allocs)
    cudaFree(p);
  g_allocs.clear();
}

struct Chunk {
  mb200_step3_job_t J;
  double alg_bytes;
  double cells;
};

// One chunk of n[0] x n[1] x n[2] cells with PML in the directions of pml_mask (bit d), D-E half
// (backward differences, chi1inv, fused E update) when eh_half is true, else the B-H half.
static Chunk make_chunk(const int n[3], int pml_mask, bool plain, int t1) {
  Chunk K;
  memset(&K, 0, sizeof(K));
  mb200_step3_job_t &J = K.J;
  const int64_t N[3] = {n[0] + 1, n[1] + 1, n[2] + 1};
  const size_t ntot = (size_t)N[0] * N[1] * N[2];
  for (int d = 0; d < 3; ++d)
    J.n[d] = n[d];
  J.stride[2] = 1;
  J.stride[1] = N[2];
  J.stride[0] = N[1] * N[2];
  J.reserved = t1;
  J.dt = 0.05;
  J.ix_lo = 0;
  J.ix_hi = n[0];
  J.noepi_lo = 0;
  J.noepi_n = 0;
  double *G[3];
  for (int d = 0; d < 3; ++d)
    G[d] = dalloc(ntot, 0.25);
  // PML tables per direction
  double *sig[3], *kap[3], *sinv[3];
  for (int d = 0; d < 3; ++d) {
    sig[d] = dalloc(2 * N[d] + 4, 0.01);
    kap[d] = dalloc(2 * N[d] + 4, 1.0);
    sinv[d] = dalloc(2 * N[d] + 4, 0.99);
  }
  double arrays = 3; // the three g arrays, read once
  for (int c = 0; c < 3; ++c) {
    mb200_step3_comp_t &C = J.c[c];
    const int d1 = (c + 1) % 3, d2 = (c + 2) % 3;
    for (int d = 0; d < 3; ++d) {
      C.lo[d] = 1;
      C.hi[d] = n[d];
      C.metal_lo[d] = C.metal_hi[d] = -1;
    }
    C.f = dalloc(ntot, 0.5);
    C.g1 = G[d2];
    C.g2 = G[d1];
    C.s1 = -J.stride[d1];
    C.s2 = -J.stride[d2];
    C.dtdx = 0.5;
    C.e = dalloc(ntot, 0.0);
    C.u = dalloc(ntot, 0.9);
    arrays += 4; // f read + write, u read, e write
    auto table = [&](mb200_pml_t &P, int d) {
      P.sig = sig[d];
      P.kap = kap[d];
      P.siginv = sinv[d];
      P.k0 = 1;
      P.ks[0] = P.ks[1] = P.ks[2] = 0;
      P.ks[d] = 2;
    };
    if (!plain) {
      if (pml_mask & (1 << d1)) table(C.pml, d1);
      if (pml_mask & (1 << d2)) {
        table(C.pmlu, d2);
        C.fu = dalloc(ntot, 0.5);
        arrays += 2;
      }
      if (pml_mask & (1 << c)) {
        table(C.pmlw, c);
        C.fw = dalloc(ntot, 0.5);
        arrays += 3; // fw read + write, e read
      }
    }
  }
  K.cells = (double)n[0] * n[1] * n[2];
  K.alg_bytes = arrays * 8.0 * K.cells;
  return K;
}

struct Table {
  mb200_step3_job_t *d_jobs;
  int64_t *d_prefix;
  int64_t tiles;
  int njobs;
};
static Table upload(const std::vector<mb200_step3_job_t> &jobs) {
  Table T;
  std::vector<int64_t> prefix(jobs.size() + 1, 0);
  for (size_t j = 0; j < jobs.size(); ++j)
    prefix[j + 1] = prefix[j] + step3_tiles(jobs[j]);
  T.tiles = prefix.back();
  T.njobs = (int)jobs.size();
  CK(cudaMalloc((void **)&T.d_jobs, sizeof(jobs[0]) * jobs.size()));
  CK(cudaMalloc((void **)&T.d_prefix, sizeof(int64_t) * prefix.size()));
  CK(cudaMemcpy(T.d_jobs, jobs.data(), sizeof(jobs[0]) * jobs.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(T.d_prefix, prefix.data(), sizeof(int64_t) * prefix.size(), cudaMemcpyHostToDevice));
  return T;
}

template <typename F> static float time_best(F launch) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int it = 0; it < 7; ++it) {
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (it >= 2 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return best;
}

static void set_flags(int pair, int lean) {
  CK(cudaMemcpyToSymbol(g_pml_pair, &pair, sizeof(int)));
  CK(cudaMemcpyToSymbol(g_pml_lean, &lean, sizeof(int)));
}

static void report(const char *shape, const char *form, int t1, const Chunk &K, float ms) {
  printf("{\"shape\": \"%s\", \"form\": \"%s\", \"t1\": %d, \"cells\": %.0f, \"alg_MB\": %.1f, \"ms\": %.4f, \"GBps\": %.0f}\n",
         shape, form, t1, K.cells, K.alg_bytes / 1e6, ms, K.alg_bytes / ms / 1e6);
  fflush(stdout);
}

int main(int argc, char **argv) {
  const int n_int = argc > 1 ? atoi(argv[1]) : 492, thick = argc > 2 ? atoi(argv[2]) : 10;
  struct Shape {
    const char *name;
    int n[3];
    int mask;
  };
  const Shape shapes[] = {
      {"face_x", {thick, n_int, n_int}, 1}, {"face_y", {n_int, thick, n_int}, 2}, {"face_z", {n_int, n_int, thick}, 4},
      {"edge_xy", {thick, thick, n_int}, 3}, {"edge_yz", {n_int, thick, thick}, 6}, {"edge_xz", {thick, n_int, thick}, 5},
      {"corner", {thick, thick, thick}, 7},
      {"cube_x", {n_int / 4, n_int / 2, n_int / 2}, 1}, // thick chunk with the face_x variants: shape effect removed
      {"all26", {0, 0, 0}, 0},                          // the 26 PML chunks of the cell in one launch, as bench.py runs them
  };
  for (const Shape &S : shapes) {
    for (int t1 : {8, 16, 32}) {
      Chunk K;
      std::vector<mb200_step3_job_t> jobs;
      if (strcmp(S.name, "all26")) {
        K = make_chunk(S.n, S.mask, false, t1);
        jobs.push_back(K.J);
      }
      else {
        memset(&K, 0, sizeof(K));
        for (int m = 1; m < 27; ++m) { // position (-1,0,+1)^3 except the centre
          int pos[3] = {m % 3, (m / 3) % 3, m / 9}, n[3], mask = 0;
          for (int d = 0; d < 3; ++d) {
            n[d] = pos[d] == 0 ? n_int : thick;
            if (pos[d] != 0) mask |= 1 << d;
          }
          Chunk Q = make_chunk(n, mask, false, t1);
          jobs.push_back(Q.J);
          K.cells += Q.cells;
          K.alg_bytes += Q.alg_bytes;
        }
      }
      Table T = upload(jobs);
      const unsigned g3 = (unsigned)(3 * T.tiles), g1 = (unsigned)T.tiles;
      CK(cudaDeviceSynchronize());
      struct Form {
        const char *name;
        int pair, lean, minb;
      };
      const Form forms[] = {{"c4_pair10_lean", 10, 1, 4}, {"c4_pair9_lean", 9, 1, 4}, {"c4_single_lean", 0, 1, 4},
                            {"c4_single_wide", 0, 0, 4},  {"c3_pair10_lean", 10, 1, 3}, {"c5_single_lean", 0, 1, 5}};
      for (const Form &F : forms) {
        set_flags(F.pair, F.lean);
        float ms;
        if (F.minb == 4) ms = time_best([&] { step3c_kernel<double, 4><<<g3, kThreads>>>(T.d_jobs, T.d_prefix, T.njobs); });
        else if (F.minb == 3) ms = time_best([&] { step3c_kernel<double, 3><<<g3, kThreads>>>(T.d_jobs, T.d_prefix, T.njobs); });
        else ms = time_best([&] { step3c_kernel<double, 5><<<g3, kThreads>>>(T.d_jobs, T.d_prefix, T.njobs); });
        report(S.name, F.name, t1, K, ms);
      }
      if (t1 == 16) {
        float ms = time_best([&] { step3_kernel<double><<<g1, kThreads>>>(T.d_jobs, T.d_prefix, T.njobs); });
        report(S.name, "three_components_per_thread", t1, K, ms);
      }
      cudaFree(T.d_jobs);
      cudaFree(T.d_prefix);
      free_all();
      // the same shape without PML through the fast-path kernel: what the shape alone costs
      if (t1 == 16) {
        Chunk P = make_chunk(S.n, 0, true, t1);
        if (!strcmp(S.name, "all26")) continue;
        Table TP = upload(std::vector<mb200_step3_job_t>(1, P.J));
        float ms = time_best([&] { step3_plain_kernel<double><<<(unsigned)TP.tiles, kThreads>>>(TP.d_jobs, TP.d_prefix, 1); });
        report(S.name, "plain_fast_path_same_shape", t1, P, ms);
        cudaFree(TP.d_jobs);
        cudaFree(TP.d_prefix);
        free_all();
      }
    }
  }
  return 0;
}
