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

namespace mb200 {
// ---- fast-path chunks and PML chunks in ONE launch ------------------------------------------------
// EXPERIMENT (negative result, kept here and not in the product): launched one after the other, the
// fast-path kernel runs at the HBM rate and the PML kernel, which is latency-bound on its thin slabs,
// leaves a third of it unused.  In one grid with the two kinds of CTA interleaved (of every U = P + Q consecutive CTAs, Q take PML units
// and P fast-path tiles, spread evenly) every SM holds both kinds at any time: the PML CTAs wait on
// their loads while the fast-path CTAs keep the memory system busy.  Measured (B200, 512^3 and 1024^3
// cells): 2.89 / 23.1 ms mixed against 2.85 / 22.4 ms for the two launches — the joint register
// allocation (64 registers, 1.6 KB of spill code) costs the fast path more than the mixing gains.  jobs[0 .. n_plain) are the
// fast-path jobs, jobs[n_plain .. n_plain + n_pml) the others; a PML unit is (tile, component).
template <typename T, int MINB, int BUDGET = step3c_budget<T>(MINB)>
__global__ void __launch_bounds__(kThreads, MINB)
    step3_mixed_kernel(const mb200_step3_job_t *__restrict__ jobs, const int64_t *__restrict__ prefix_plain,
                       int n_plain, const int64_t *__restrict__ prefix_pml, int n_pml, int64_t units_plain,
                       int64_t units_pml) {
  __shared__ mb200_step3_job_t J;
  const int64_t U = units_plain + units_pml, b = (int64_t)blockIdx.x;
  const int64_t q0 = b * units_pml / U, q1 = (b + 1) * units_pml / U;
  int64_t tile;
  if (q1 > q0) {
    stage_job_at(&J, jobs + n_plain, prefix_pml, n_pml, q0 / 3, &tile);
    step3c_thread<T, BUDGET, false>(J, (int)(q0 % 3), tile, threadIdx.x);
  }
  else {
    stage_job_at(&J, jobs, prefix_plain, n_plain, b - q0, &tile);
    step3_plain_thread<T>(J, tile, threadIdx.x);
  }
}

} // namespace mb200

// EXPERIMENT: the interior march alone (no masks, no table look-ups, descriptor in constant space) over every
// tile of a job whose components own the whole box — what a launch restricted to interior tiles would run at
template <typename T, int MINB, bool EPI, int NP>
__global__ void __launch_bounds__(mb200::kThreads, MINB) lean_kernel(const __grid_constant__ mb200_step3_job_t J) {
  using namespace mb200;
  const mb200_box_t box = step3_box(J);
  int ix0, ix_end, iy, iz;
  if (!box_thread_point(box, (int64_t)blockIdx.x, threadIdx.x, ix0, ix_end, iy, iz, step3_t1(J))) return;
  if (iy < 1 || iz < 1) return;
  if (ix0 < 1) ix0 = 1;
  const int64_t i = box_index(box, ix0, iy, iz);
  const bool metal_yz[3] = {false, false, false};
  step3_plain_fast<T, EPI, EPI, NP>(J, i, box.s[0], ix0, ix_end, metal_yz);
}

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

static void report(const char *shape, const char *form, int t1, const Chunk &K, float ms) {
  printf("{\"shape\": \"%s\", \"form\": \"%s\", \"t1\": %d, \"cells\": %.0f, \"alg_MB\": %.1f, \"ms\": %.4f, \"GBps\": %.0f}\n",
         shape, form, t1, K.cells, K.alg_bytes / 1e6, ms, K.alg_bytes / ms / 1e6);
  fflush(stdout);
}

int main(int argc, char **argv) {
  // pml_shapes [interior cells per edge] [PML thickness in cells] [shape filter] [form filter] [planes per CTA]
  const int n_int = argc > 1 ? atoi(argv[1]) : 492, thick = argc > 2 ? atoi(argv[2]) : 10;
  const char *shape_filter = argc > 3 && strcmp(argv[3], "all") ? argv[3] : nullptr;
  const char *form_filter = argc > 4 && strcmp(argv[4], "all") ? argv[4] : nullptr;
  const int only_t1 = argc > 5 ? atoi(argv[5]) : 0;
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
    if (shape_filter && !strstr(S.name, shape_filter)) continue;
    for (int t1 : {8, 16, 32}) {
      if (only_t1 && t1 != only_t1) continue;
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
        int minb, budget;
      };
      // budget = values in flight per thread; planes per iteration = budget / operands of the variant (1..4)
      const Form forms[] = {{"c4_b10", 4, 10}, {"c4_b20", 4, 20}, {"c3_b20", 3, 20}, {"c3_b28", 3, 28},
                            {"c2_b30", 2, 30}, {"c2_b40", 2, 40}, {"c2_b52", 2, 52}};
      for (const Form &F : forms) {
        if (form_filter && !strstr(F.name, form_filter)) continue;
        float ms = 0;
#define RUN(MB, BU)                                                                                            \
  if (F.minb == MB && F.budget == BU)                                                                        \
    ms = time_best([&] { step3c_kernel<double, MB, false, BU><<<g3, kThreads>>>(T.d_jobs, T.d_prefix, T.njobs); });
        RUN(4, 10) RUN(4, 20) RUN(3, 20) RUN(3, 28) RUN(2, 30) RUN(2, 40) RUN(2, 52)
#undef RUN
        report(S.name, F.name, t1, K, ms);
      }
      if (t1 == 16 && !form_filter) {
        float ms = time_best([&] { step3_kernel<double><<<g1, kThreads>>>(T.d_jobs, T.d_prefix, T.njobs); });
        report(S.name, "three_components_per_thread", t1, K, ms);
      }
      cudaFree(T.d_jobs);
      cudaFree(T.d_prefix);
      free_all();
      // the same shape without PML through the fast-path kernel: what the shape alone costs
      if (t1 == 16 && !form_filter) {
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
  // ---- fast path: table-driven masked march (the product's default) vs the interior march alone
  if (shape_filter && !strcmp(shape_filter, "lean")) {
    const int t1 = 16;
    const int ni[3] = {n_int, n_int, n_int};
    for (int half = 0; half < 2; ++half) {
      Chunk P = make_chunk(ni, 0, true, t1);
      if (half == 0) { // B half: H aliases B, no epilogue
        for (int c = 0; c < 3; ++c)
          P.J.c[c].e = nullptr, P.J.c[c].u = nullptr;
        P.alg_bytes = 9 * 8.0 * P.cells;
      }
      Table TP = upload(std::vector<mb200_step3_job_t>(1, P.J));
      const unsigned g = (unsigned)TP.tiles;
      const char *hn = half ? "DE" : "B";
      char name[64];
#define REP(T, label, launch)                                                                                  \
  {                                                                                                            \
    Chunk Q = P;                                                                                               \
    Q.alg_bytes = P.alg_bytes * sizeof(T) / 8;                                                                 \
    const float ms = time_best([&] { launch; });                                                               \
    snprintf(name, sizeof(name), "%s_%s_%s", hn, #T, label);                                                   \
    report("interior", name, t1, Q, ms);                                                                       \
  }
      REP(double, "masked", (step3_plain_kernel<double><<<g, kThreads>>>(TP.d_jobs, TP.d_prefix, 1)))
      REP(float, "masked", (step3_plain_kernel<float><<<g, kThreads>>>(TP.d_jobs, TP.d_prefix, 1)))
      if (half == 0) {
        REP(double, "lean_c4_np1", (lean_kernel<double, 4, false, 1><<<g, kThreads>>>(P.J)))
        REP(double, "lean_c4_np2", (lean_kernel<double, 4, false, 2><<<g, kThreads>>>(P.J)))
        REP(double, "lean_c3_np3", (lean_kernel<double, 3, false, 3><<<g, kThreads>>>(P.J)))
        REP(float, "lean_c4_np2", (lean_kernel<float, 4, false, 2><<<g, kThreads>>>(P.J)))
        REP(float, "lean_c4_np3", (lean_kernel<float, 4, false, 3><<<g, kThreads>>>(P.J)))
        REP(float, "lean_c4_np4", (lean_kernel<float, 4, false, 4><<<g, kThreads>>>(P.J)))
        REP(float, "lean_c5_np2", (lean_kernel<float, 5, false, 2><<<g, kThreads>>>(P.J)))
      }
      else {
        REP(double, "lean_c4_np1", (lean_kernel<double, 4, true, 1><<<g, kThreads>>>(P.J)))
        REP(double, "lean_c3_np2", (lean_kernel<double, 3, true, 2><<<g, kThreads>>>(P.J)))
        REP(float, "lean_c4_np1", (lean_kernel<float, 4, true, 1><<<g, kThreads>>>(P.J)))
        REP(float, "lean_c4_np2", (lean_kernel<float, 4, true, 2><<<g, kThreads>>>(P.J)))
        REP(float, "lean_c5_np1", (lean_kernel<float, 5, true, 1><<<g, kThreads>>>(P.J)))
        REP(float, "lean_c3_np3", (lean_kernel<float, 3, true, 3><<<g, kThreads>>>(P.J)))
      }
      // the product's lean + shell launches (MEEP_B200_PLAIN_LEAN=1): full box, two x-slabs, (y, z) shell columns
      {
        const Step3Shell S = step3_shell(P.J);
        int *d_cols;
        CK(cudaMalloc((void **)&d_cols, sizeof(int) * (S.cols.size() + 1)));
        CK(cudaMemcpy(d_cols, S.cols.data(), sizeof(int) * S.cols.size(), cudaMemcpyHostToDevice));
        Table TS = upload(S.slabs);
        const int ncols = (int)S.cols.size();
        const dim3 gc((unsigned)((ncols + kThreads - 1) / kThreads), (unsigned)((S.hi[0] - S.lo[0] + t1) / t1));
#define SHELL(T)                                                                                               \
  (step3_lean_kernel<T><<<g, kThreads>>>(P.J),                                                                 \
   step3_plain_kernel<T><<<(unsigned)TS.tiles, kThreads>>>(TS.d_jobs, TS.d_prefix, TS.njobs),                  \
   step3_cols_kernel<T><<<gc, kThreads>>>(TP.d_jobs, d_cols, ncols, S.lo[0], S.hi[0], t1))
        REP(double, "lean_plus_shell", SHELL(double))
        REP(float, "lean_plus_shell", SHELL(float))
#undef SHELL
        fprintf(stderr, "full box [%d..%d]x[%d..%d]x[%d..%d], %d shell columns, %lld slab tiles of %lld\n", S.lo[0], S.hi[0],
                S.lo[1], S.hi[1], S.lo[2], S.hi[2], ncols, (long long)TS.tiles, (long long)TP.tiles);
        cudaFree(d_cols);
        cudaFree(TS.d_jobs);
        cudaFree(TS.d_prefix);
      }
#undef REP
      cudaFree(TP.d_jobs);
      cudaFree(TP.d_prefix);
      free_all();
    }
    return 0;
  }
  // ---- the interior chunk (fast path) and the 26 PML chunks: one after the other vs one mixed grid
  if (!shape_filter || !strcmp(shape_filter, "mixed")) {
    const int t1 = 16;
    std::vector<mb200_step3_job_t> jobs;
    const int ni[3] = {n_int, n_int, n_int};
    Chunk P = make_chunk(ni, 0, true, t1), K;
    memset(&K, 0, sizeof(K));
    jobs.push_back(P.J);
    for (int m = 1; m < 27; ++m) {
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
    Table TP = upload(std::vector<mb200_step3_job_t>(1, jobs[0]));
    Table TQ = upload(std::vector<mb200_step3_job_t>(jobs.begin() + 1, jobs.end()));
    Table TA = upload(jobs); // (only the job array of this one is used)
    const float ms_p = time_best([&] { step3_plain_kernel<double><<<(unsigned)TP.tiles, kThreads>>>(TP.d_jobs, TP.d_prefix, 1); });
    report("interior", "plain_alone", t1, P, ms_p);
    const float ms_q4 = time_best([&] { step3c_kernel<double, 4><<<(unsigned)(3 * TQ.tiles), kThreads>>>(TQ.d_jobs, TQ.d_prefix, TQ.njobs); });
    report("all26", "pml_alone_c4", t1, K, ms_q4);
    const float ms_q3 = time_best([&] { step3c_kernel<double, 3><<<(unsigned)(3 * TQ.tiles), kThreads>>>(TQ.d_jobs, TQ.d_prefix, TQ.njobs); });
    report("all26", "pml_alone_c3", t1, K, ms_q3);
    Chunk S = P;
    S.cells += K.cells;
    S.alg_bytes += K.alg_bytes;
    const float ms_seq = time_best([&] {
      step3_plain_kernel<double><<<(unsigned)TP.tiles, kThreads>>>(TP.d_jobs, TP.d_prefix, 1);
      step3c_kernel<double, 4><<<(unsigned)(3 * TQ.tiles), kThreads>>>(TQ.d_jobs, TQ.d_prefix, TQ.njobs);
    });
    report("interior+all26", "two_launches_c4", t1, S, ms_seq);
    const unsigned gm = (unsigned)(TP.tiles + 3 * TQ.tiles);
    const float ms_m4 = time_best([&] {
      step3_mixed_kernel<double, 4><<<gm, kThreads>>>(TA.d_jobs, TP.d_prefix, 1, TQ.d_prefix, TQ.njobs, TP.tiles, 3 * TQ.tiles);
    });
    report("interior+all26", "mixed_c4", t1, S, ms_m4);
    const float ms_m3 = time_best([&] {
      step3_mixed_kernel<double, 3><<<gm, kThreads>>>(TA.d_jobs, TP.d_prefix, 1, TQ.d_prefix, TQ.njobs, TP.tiles, 3 * TQ.tiles);
    });
    report("interior+all26", "mixed_c3", t1, S, ms_m3);
    free_all();
  }
  return 0;
}
