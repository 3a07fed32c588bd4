// lean_property.cpp — property test of the lean + shell decomposition of the fast path
// (meep_b200/csrc/fused.cuh: step3_full_box / step3_shell / step3_lean_thread / step3_plain_column),
// compiled for the host (the same headers the CUDA build and the emulator use).
//
// For many random "plain" jobs — chunk sizes from 1 to 37 cells per direction, owned ranges that
// start at 0 or 1 and end at n-1 or n per component, forward or backward differences, metal planes
// on any face or none, x-slab restrictions, planes without the epilogue, with / without chi1inv,
// with / without the E/H epilogue, several planes-per-CTA settings — the arrays produced by
//   (a) the masked march over every tile of the job (the round-1 form), and
//   (b) the lean march over the full box + the two x-slab jobs + the shell columns
// must be bit-identical: every point updated exactly once, by the same arithmetic.
//
// usage: lean_property <seed> <number of jobs>       (prints "OK <jobs> <lean jobs>" or a diagnosis)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <random>
#include <vector>

#include "../../meep_b200/csrc/fused.cuh"

using namespace mb200;

template <typename T> struct Arrays {
  std::vector<T> g[3], f[3], e[3], u[3];
};

template <typename T>
static void fill(Arrays<T> &A, size_t ntot, std::mt19937 &rng) {
  std::uniform_real_distribution<double> d(-1.0, 1.0);
  for (int c = 0; c < 3; ++c) {
    A.g[c].resize(ntot);
    A.f[c].resize(ntot);
    A.e[c].resize(ntot);
    A.u[c].resize(ntot);
    for (size_t i = 0; i < ntot; ++i) {
      A.g[c][i] = (T)d(rng);
      A.f[c][i] = (T)d(rng);
      A.e[c][i] = (T)d(rng);
      A.u[c][i] = (T)(0.5 + 0.25 * d(rng));
    }
  }
}

template <typename T> static void bind(mb200_step3_job_t &J, Arrays<T> &A, bool epi, bool hasu) {
  for (int c = 0; c < 3; ++c) {
    J.c[c].f = A.f[c].data();
    J.c[c].g1 = A.g[(c + 2) % 3].data();
    J.c[c].g2 = A.g[(c + 1) % 3].data();
    J.c[c].e = epi ? A.e[c].data() : nullptr;
    J.c[c].u = (epi && hasu) ? A.u[c].data() : nullptr;
  }
}

template <typename T> static int run_one(std::mt19937 &rng, int &lean_jobs) {
  std::uniform_int_distribution<int> size(1, 37), coin(0, 1), three(0, 2), t1s(0, 3);
  mb200_step3_job_t J;
  memset(&J, 0, sizeof(J));
  for (int d = 0; d < 3; ++d)
    J.n[d] = size(rng);
  if (coin(rng)) J.n[2] = 20 + size(rng); // (patch-mode tiling needs n3 >= 24; thin boxes use the flattened mode)
  J.stride[2] = 1;
  J.stride[1] = J.n[2] + 1;
  J.stride[0] = (int64_t)(J.n[1] + 1) * (J.n[2] + 1);
  const int t1v[4] = {0, 4, 16, 5};
  J.reserved = t1v[t1s(rng)];
  J.dt = 0.05;
  const bool backward = coin(rng) != 0; // D/E half: differences towards -stride; B/H half: towards +stride
  const bool epi = coin(rng) != 0, hasu = coin(rng) != 0;
  for (int c = 0; c < 3; ++c) {
    mb200_step3_comp_t &C = J.c[c];
    const int d1 = (c + 1) % 3, d2 = (c + 2) % 3;
    for (int d = 0; d < 3; ++d) {
      // the neighbour read must stay inside the array
      const bool reads = d == d1 || d == d2;
      C.lo[d] = (reads && backward) ? 1 : coin(rng);
      C.hi[d] = (reads && !backward) ? J.n[d] - 1 : J.n[d] - coin(rng);
      C.metal_lo[d] = C.metal_hi[d] = -1;
      if (epi && three(rng) == 0) C.metal_lo[d] = C.lo[d];
      if (epi && three(rng) == 0 && C.hi[d] != C.lo[d]) C.metal_hi[d] = C.hi[d];
    }
    C.s1 = (backward ? -1 : 1) * J.stride[d1];
    C.s2 = (backward ? -1 : 1) * J.stride[d2];
    C.dtdx = 0.4 + 0.1 * c;
  }
  J.ix_lo = 0;
  J.ix_hi = J.n[0];
  if (three(rng) == 0) { // an x-slab of the chunk
    std::uniform_int_distribution<int> px(0, J.n[0]);
    int a = px(rng), b = px(rng);
    if (a > b) std::swap(a, b);
    J.ix_lo = a;
    J.ix_hi = b;
  }
  J.noepi_lo = 0;
  J.noepi_n = 0;
  if (epi && three(rng) == 0) {
    std::uniform_int_distribution<int> px(0, J.n[0]);
    J.noepi_lo = px(rng);
    J.noepi_n = 1 + coin(rng);
  }
  const size_t ntot = (size_t)(J.n[0] + 1) * (J.n[1] + 1) * (J.n[2] + 1);
  Arrays<T> A, B;
  fill(A, ntot, rng);
  B = A;

  // (a) the masked march over every tile
  bind(J, A, epi, hasu);
  if (!step3_is_plain(J)) {
    printf("internal: the random job is not a fast-path job\n");
    return 1;
  }
  const int64_t tiles = step3_tiles(J);
  for (int64_t t = 0; t < tiles; ++t)
    for (int tid = 0; tid < kThreads; ++tid)
      step3_plain_thread<T>(J, t, tid);

  // (b) lean + shell (the launches of launch_step3, thread by thread)
  bind(J, B, epi, hasu);
  const Step3Shell S = step3_shell(J);
  if (S.lean) {
    ++lean_jobs;
    for (int64_t t = 0; t < tiles; ++t)
      for (int tid = 0; tid < kThreads; ++tid)
        step3_lean_thread<T>(J, t, tid);
    for (const mb200_step3_job_t &slab : S.slabs) {
      const int64_t st = step3_tiles(slab);
      for (int64_t t = 0; t < st; ++t)
        for (int tid = 0; tid < kThreads; ++tid)
          step3_plain_thread<T>(slab, t, tid);
    }
    const int row = J.n[2] + 1, planes = step3_t1(J);
    for (int col : S.cols)
      for (int x0 = S.lo[0]; x0 <= S.hi[0]; x0 += planes)
        step3_plain_column<T>(J, x0, x0 + planes < S.hi[0] + 1 ? x0 + planes : S.hi[0] + 1, col / row, col % row);
  }
  else {
    for (int64_t t = 0; t < tiles; ++t)
      for (int tid = 0; tid < kThreads; ++tid)
        step3_plain_thread<T>(J, t, tid);
  }
  for (int c = 0; c < 3; ++c) {
    if (memcmp(A.f[c].data(), B.f[c].data(), ntot * sizeof(T)) || memcmp(A.e[c].data(), B.e[c].data(), ntot * sizeof(T))) {
      for (size_t i = 0; i < ntot; ++i)
        if (A.f[c][i] != B.f[c][i] || A.e[c][i] != B.e[c][i]) {
          const int ix = (int)(i / J.stride[0]), iy = (int)((i % J.stride[0]) / J.stride[1]), iz = (int)(i % J.stride[1]);
          printf("MISMATCH component %d at (%d,%d,%d): f %.17g vs %.17g, e %.17g vs %.17g; n = %d %d %d, slab %d..%d, "
                 "full box x %d..%d y %d..%d z %d..%d, epi %d hasu %d backward %d t1 %d\n",
                 c, ix, iy, iz, (double)A.f[c][i], (double)B.f[c][i], (double)A.e[c][i], (double)B.e[c][i], J.n[0], J.n[1],
                 J.n[2], J.ix_lo, J.ix_hi, S.lo[0], S.hi[0], S.lo[1], S.hi[1], S.lo[2], S.hi[2], (int)epi, (int)hasu,
                 (int)backward, J.reserved);
          return 1;
        }
    }
  }
  return 0;
}

int main(int argc, char **argv) {
  const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1;
  const int jobs = argc > 2 ? atoi(argv[2]) : 200;
  std::mt19937 rng(seed);
  int lean_jobs = 0;
  for (int k = 0; k < jobs; ++k) {
    const int rc = (k % 2) ? run_one<float>(rng, lean_jobs) : run_one<double>(rng, lean_jobs);
    if (rc) {
      printf("job %d of seed %u failed\n", k, seed);
      return 1;
    }
  }
  printf("OK %d %d\n", jobs, lean_jobs);
  return 0;
}
