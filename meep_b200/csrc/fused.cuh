// fused.cuh — the fused half-step kernel (MB200_K_STEP3, see include/meep_b200.h).
//
// One pass over a 3-D chunk updates all three components of D (or B) — the three step_curl
// calls of fields_chunk::step_db (reference src/step_db.cpp:47-127) — and, where the host
// found it legal, applies the diagonal update_eh (src/update_eh.cpp:190-195) to the freshly
// computed value while it is still in a register.  HBM traffic per cell-step drops from
// 33R (three curl passes + three EDHB passes per field pair) to the algorithmic 24R.
//
// Thread mapping: the index box [0..n]^3 (owned points of every component plus the not-owned
// planes, which are masked per component) is tiled exactly like the generic box kernels:
// 256 threads cover a (4 x 64) patch of (y,z) and march kT1 planes in x, so +/-y neighbours
// hit L1 and z neighbours sit in the same 128-byte lines.
#ifndef MEEP_B200_FUSED_CUH
#define MEEP_B200_FUSED_CUH

#include "kernels.cuh"

namespace mb200 {

MB200_HD mb200_box_t step3_box(const mb200_step3_job_t &J) {
  mb200_box_t b;
  for (int d = 0; d < 3; ++d) {
    b.s[d] = J.stride[d];
    b.n[d] = J.n[d] + 1;
  }
  // restrict to the slab ix_lo..ix_hi along direction 0
  const int lo = J.ix_lo > 0 ? J.ix_lo : 0;
  const int hi = J.ix_hi < J.n[0] ? J.ix_hi : J.n[0];
  b.idx0 = (int64_t)lo * J.stride[0];
  b.n[0] = hi - lo + 1;
  b.reserved = lo; // first ix of the box
  return b;
}
MB200_HD int64_t step3_tiles(const mb200_step3_job_t &J) { return box_tiles(step3_box(J)); }
// owned points of component C inside the job's slab
inline double step3_comp_points(const mb200_step3_job_t &J, const mb200_step3_comp_t &C) {
  double q = 1;
  for (int d = 0; d < 3; ++d) {
    int lo = C.lo[d], hi = C.hi[d];
    if (d == 0) {
      if (J.ix_lo > lo) lo = J.ix_lo;
      if (J.ix_hi < hi) hi = J.ix_hi;
    }
    q *= hi >= lo ? (double)(hi - lo + 1) : 0.0;
  }
  return q;
}
inline double step3_points(const mb200_step3_job_t &J) {
  double p = 0;
  for (int c = 0; c < 3; ++c)
    if (J.c[c].f) p += step3_comp_points(J, J.c[c]);
  return p / 3.0; // cells (each cell has three components)
}
inline double step3_bytes(const mb200_step3_job_t &J, double R) {
  double bytes = 0;
  // distinct g arrays are read once
  const void *g[6];
  int ng = 0;
  for (int c = 0; c < 3; ++c) {
    const mb200_step3_comp_t &C = J.c[c];
    if (!C.f) continue;
    const double q = step3_comp_points(J, C);
    int arrays = 2 + (C.pmlu.sig ? 2 : 0) + (C.cnd ? 2 + (C.pml.sig ? 2 : 0) : 0);
    if (C.e) arrays += 1 + (C.u ? 1 : 0) + (C.pmlw.sig ? 3 : 0);
    const void *gs[2] = {C.g1, C.g2};
    for (int k = 0; k < 2; ++k) {
      bool seen = gs[k] == nullptr;
      for (int m = 0; m < ng && !seen; ++m)
        seen = g[m] == gs[k];
      if (!seen) {
        g[ng++] = gs[k];
        arrays += 1;
      }
    }
    bytes += R * arrays * q;
  }
  return bytes;
}

// one point, one component
template <typename T>
MB200_HD void step3_comp_point(const mb200_step3_comp_t &C, int variant, int64_t i, int ix, int iy,
                               int iz, T dt2) {
  if (ix < C.lo[0] || ix > C.hi[0] || iy < C.lo[1] || iy > C.hi[1] || iz < C.lo[2] ||
      iz > C.hi[2])
    return;
  const int k = pml_k(C.pml, ix, iy, iz), ku = pml_k(C.pmlu, ix, iy, iz);
  const T d = curl_point_any<T>(C, variant, i, k, ku, (T)C.dtdx, dt2);
  if (C.e) {
    const bool metal = ix == C.metal_lo[0] || ix == C.metal_hi[0] || iy == C.metal_lo[1] ||
                       iy == C.metal_hi[1] || iz == C.metal_lo[2] || iz == C.metal_hi[2];
    edhb_diag<T>(C, i, pml_k(C.pmlw, ix, iy, iz), metal ? T(0) : d);
  }
}

// ---- fast path ---------------------------------------------------------------------------------
// "plain" job: all three components present, none with PML / f_u / conductivity (the interior
// chunk of a PML-padded cell: ~89 % of the cells of BASELINE config 2), fused E/H update
// without f_w.  All loads of a grid point (3 f + 12 g + 3 chi1inv) are issued before the first
// store, so each warp keeps ~4.6 KB in flight instead of ~1.3 KB.
MB200_HD bool step3_is_plain(const mb200_step3_job_t &J) {
  for (int c = 0; c < 3; ++c) {
    const mb200_step3_comp_t &C = J.c[c];
    if (!C.f || curl_variant(C) != 1) return false;
    if (C.e && C.pmlw.sig) return false;
  }
  return true;
}

template <typename T>
MB200_HD void step3_plain_thread(const mb200_step3_job_t &J, int64_t tile, int tid) {
  const mb200_box_t box = step3_box(J);
  int ix0, ix_end, iy, iz;
  if (!box_thread_point(box, tile, tid, ix0, ix_end, iy, iz)) return;
  int64_t i = box_index(box, ix0, iy, iz);
  const int64_t sx = box.s[0];
  ix0 += box.reserved;
  ix_end += box.reserved;
  bool myz[3], metal_yz[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const mb200_step3_comp_t &C = J.c[c];
    myz[c] = iy >= C.lo[1] && iy <= C.hi[1] && iz >= C.lo[2] && iz <= C.hi[2];
    metal_yz[c] = iy == C.metal_lo[1] || iy == C.metal_hi[1] || iz == C.metal_lo[2] ||
                  iz == C.metal_hi[2];
  }
  for (int ix = ix0; ix < ix_end; ++ix, i += sx) {
    T fv[3], a1[3], c1[3], c2[3], a2[3], uv[3];
    bool m[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { // ---- all loads first
      const mb200_step3_comp_t &C = J.c[c];
      m[c] = myz[c] && ix >= C.lo[0] && ix <= C.hi[0];
      if (m[c]) {
        const T *g1 = (const T *)C.g1, *g2 = (const T *)C.g2;
        fv[c] = ((const T *)C.f)[i];
        a1[c] = ldro(g1 + i + C.s1);
        c1[c] = ldro(g1 + i);
        c2[c] = ldro(g2 + i);
        a2[c] = ldro(g2 + i + C.s2);
        uv[c] = (C.e && C.u) ? ldro((const T *)C.u + i) : T(1);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { // ---- then arithmetic + stores
      const mb200_step3_comp_t &C = J.c[c];
      if (m[c]) {
        T dg = a1[c] - c1[c];
        dg = dg + c2[c] - a2[c];
        const T d = fv[c] - (T)C.dtdx * dg;
        ((T *)C.f)[i] = d;
        if (C.e) {
          const bool metal = metal_yz[c] || ix == C.metal_lo[0] || ix == C.metal_hi[0];
          const T dd = metal ? T(0) : d;
          ((T *)C.e)[i] = C.u ? dd * uv[c] : dd;
        }
      }
    }
  }
}

template <typename T>
MB200_HD void step3_thread(const mb200_step3_job_t &J, int64_t tile, int tid) {
  const mb200_box_t box = step3_box(J);
  int ix0, ix_end, iy, iz;
  if (!box_thread_point(box, tile, tid, ix0, ix_end, iy, iz)) return;
  const T dt2 = (T)J.dt * T(0.5);
  const int v0 = J.c[0].f ? curl_variant(J.c[0]) : -1;
  const int v1 = J.c[1].f ? curl_variant(J.c[1]) : -1;
  const int v2 = J.c[2].f ? curl_variant(J.c[2]) : -1;
  int64_t i = box_index(box, ix0, iy, iz);
  const int64_t sx = box.s[0];
  ix0 += box.reserved; // loop index -> array index along direction 0
  ix_end += box.reserved;
  for (int ix = ix0; ix < ix_end; ++ix, i += sx) {
    if (v0 >= 0) step3_comp_point<T>(J.c[0], v0, i, ix, iy, iz, dt2);
    if (v1 >= 0) step3_comp_point<T>(J.c[1], v1, i, ix, iy, iz, dt2);
    if (v2 >= 0) step3_comp_point<T>(J.c[2], v2, i, ix, iy, iz, dt2);
  }
}

#ifdef __CUDACC__

template <typename T>
__global__ void __launch_bounds__(kThreads)
    step3_kernel(const mb200_step3_job_t *__restrict__ jobs,
                 const int64_t *__restrict__ tile_prefix, int njobs) {
  __shared__ mb200_step3_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  step3_thread<T>(J, tile, threadIdx.x);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    step3_plain_kernel(const mb200_step3_job_t *__restrict__ jobs,
                       const int64_t *__restrict__ tile_prefix, int njobs) {
  __shared__ mb200_step3_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  step3_plain_thread<T>(J, tile, threadIdx.x);
}

template <typename T>
static void launch_step3(const mb200_step3_job_t *jobs, const int64_t *prefix, int njobs,
                         int64_t tiles, bool all_plain, cudaStream_t s) {
  if (all_plain)
    step3_plain_kernel<T><<<dim3((unsigned)tiles), dim3(kThreads), 0, s>>>(jobs, prefix, njobs);
  else
    step3_kernel<T><<<dim3((unsigned)tiles), dim3(kThreads), 0, s>>>(jobs, prefix, njobs);
}

#endif // __CUDACC__

} // namespace mb200
#endif
