// kernels.cuh — sm_100a kernels of the FDTD hot path (generic, table-driven versions).
//
// Launch model: one grid per plan run.  A plan holds a device table of jobs (each job = the
// argument list of one reference inner-loop call, see include/meep_b200.h) and a prefix sum of
// per-job tile counts; every CTA binary-searches its job, stages the descriptor in shared
// memory and processes one tile.  So a whole phase of fields::step over all 27+ chunks of a
// GPU is ONE launch instead of (chunks x components x cmp) launches.
//
// Tile shapes for 3-loop ("box") jobs.  The arrays keep the reference layout (last loop
// stride 1).  A CTA of 256 threads covers a (t2 x t3) patch of loops 2,3 and marches over up
// to MB200_T1 planes of loop 1, so the +/-s2 neighbour rows of a stencil hit L1 and the
// per-thread index/PML setup is amortised.  For thin boxes (n3 < 24, e.g. z-PML slabs with
// n3 = 10) loops 2,3 are flattened instead so that lanes stay dense.
#ifndef MEEP_B200_KERNELS_CUH
#define MEEP_B200_KERNELS_CUH

#include "point_ops.h"

namespace mb200 {

constexpr int kThreads = 256;
constexpr int kT1 = 8; // loop-1 planes marched per CTA
constexpr int kItems1D = 4;  // elements per thread for streaming 1-D jobs
constexpr int kDftPts = 64;  // monitor points per CTA in dft_kernel
constexpr int kFluxPts = 256; // points per CTA in flux_kernel
constexpr int kFmpBlocks = 4; // zero-flag blocks per CTA in fmp_kernel

// host+device: tile decomposition of a box
struct BoxTiling {
  int t3;      // threads along loop 3 (0 => flattened mode)
  int nb23;    // tiles over loops (2,3)
  int nb3;     // tiles along loop 3 (patch mode)
  int nb1;     // tiles along loop 1
};

MB200_HD BoxTiling box_tiling(const mb200_box_t &b, int t1 = kT1) {
  BoxTiling t;
  if (b.n[2] >= 48) t.t3 = 64;
  else if (b.n[2] >= 24) t.t3 = 32;
  else t.t3 = 0;
  if (t.t3) {
    const int t2 = kThreads / t.t3;
    t.nb3 = (b.n[2] + t.t3 - 1) / t.t3;
    t.nb23 = t.nb3 * ((b.n[1] + t2 - 1) / t2);
  }
  else {
    const int64_t nq = (int64_t)b.n[1] * b.n[2];
    t.nb3 = 1;
    t.nb23 = (int)((nq + kThreads - 1) / kThreads);
  }
  t.nb1 = (b.n[0] + t1 - 1) / t1;
  return t;
}
MB200_HD int64_t box_tiles(const mb200_box_t &b, int t1 = kT1) {
  if (b.n[0] <= 0 || b.n[1] <= 0 || b.n[2] <= 0) return 0;
  BoxTiling t = box_tiling(b, t1);
  return (int64_t)t.nb1 * t.nb23;
}

// map (tile, thread) -> loop indices; returns false if this thread has no (i2,i3)
MB200_HD bool box_thread_point(const mb200_box_t &b, int64_t tile, int tid, int &i1_0, int &i1_end,
                               int &i2, int &i3, int t1 = kT1) {
  const BoxTiling t = box_tiling(b, t1);
  const int b1 = (int)(tile / t.nb23);
  const int b23 = (int)(tile - (int64_t)b1 * t.nb23);
  i1_0 = b1 * t1;
  i1_end = i1_0 + t1 < b.n[0] ? i1_0 + t1 : b.n[0];
  if (t.t3) {
    const int b2 = b23 / t.nb3, b3 = b23 - b2 * t.nb3;
    const int l3 = tid & (t.t3 - 1), l2 = tid / t.t3;
    i3 = b3 * t.t3 + l3;
    i2 = b2 * (kThreads / t.t3) + l2;
    return i3 < b.n[2] && i2 < b.n[1];
  }
  else {
    const int64_t q = (int64_t)b23 * kThreads + tid;
    if (q >= (int64_t)b.n[1] * b.n[2]) return false;
    i2 = (int)(q / b.n[2]);
    i3 = (int)(q - (int64_t)i2 * b.n[2]);
    return true;
  }
}


// ---- per-thread bodies of the box kernels (shared with the test-only emulator) ------------------
template <typename T> MB200_HD void curl_thread(const mb200_curl_job_t &J, int64_t tile, int tid) {
  int i1_0, i1_end, i2, i3;
  if (!box_thread_point(J.box, tile, tid, i1_0, i1_end, i2, i3)) return;
  const int variant = curl_variant(J);
  const T dtdx = (T)J.dtdx, dt2 = (T)J.dt * T(0.5);
  int64_t i = box_index(J.box, i1_0, i2, i3);
  int k = pml_k(J.pml, i1_0, i2, i3), ku = pml_k(J.pmlu, i1_0, i2, i3);
  const int64_t s1 = J.box.s[0];
  const int dk = J.pml.ks[0], dku = J.pmlu.ks[0];
  switch (variant) {
#define MB200_CASE(v)                                                                              \
  case v:                                                                                          \
    for (int i1 = i1_0; i1 < i1_end; ++i1, i += s1, k += dk, ku += dku)                            \
      curl_point<T, ((v)&8) != 0, ((v)&4) != 0, ((v)&2) != 0, ((v)&1) != 0>(J, i, k, ku, dtdx,     \
                                                                             dt2);                 \
    break;
    MB200_CASE(0) MB200_CASE(1) MB200_CASE(2) MB200_CASE(3) MB200_CASE(4) MB200_CASE(5)
    MB200_CASE(6) MB200_CASE(7) MB200_CASE(8) MB200_CASE(9) MB200_CASE(10) MB200_CASE(11)
    MB200_CASE(12) MB200_CASE(13) MB200_CASE(14) MB200_CASE(15)
#undef MB200_CASE
  }
}

constexpr int kBatch = 4; // loop-1 planes whose loads are issued together
constexpr int kEdhbT1 = 16;   // loop-1 planes marched per CTA in edhb_kernel
constexpr int kEdhbBatch = 8; // planes in flight per thread for the plain E = chi1inv D update
constexpr int kZBlocksPerCta = 16; // zero-block Lorentz kernel: blocks walked by one CTA

// step_update_EDHB.  Diagonal, linear jobs (the common case when the update could not be fused
// into the D/B pass: f_minus_p present, 1-D/2-D grids, tiled update_eh) take a batched path that
// issues the loads of kBatch planes before the first store; everything else goes point by point.
template <typename T> MB200_HD void edhb_thread(const mb200_edhb_job_t &J, int64_t tile, int tid) {
  int i1_0, i1_end, i2, i3;
  if (!box_thread_point(J.box, tile, tid, i1_0, i1_end, i2, i3, kEdhbT1)) return;
  int64_t i = box_index(J.box, i1_0, i2, i3);
  int kw = pml_k(J.pmlw, i1_0, i2, i3);
  const int64_t s1 = J.box.s[0];
  const int dk = J.pmlw.ks[0];
  if (!J.u1 && !J.u2 && !J.chi3) {
    T *f = (T *)J.f, *fw = (T *)J.fw;
    const T *g = (const T *)J.g, *u = (const T *)J.u;
    const T *sigw = (const T *)J.pmlw.sig, *kapw = (const T *)J.pmlw.kap;
    if (!sigw) {
      // E = chi1inv D (two loads and a store per point): kEdhbBatch planes in flight per thread
      for (int i1 = i1_0; i1 < i1_end; i1 += kEdhbBatch, i += kEdhbBatch * s1) {
        T gv[kEdhbBatch], uv[kEdhbBatch];
#pragma unroll
        for (int k = 0; k < kEdhbBatch; ++k)
          if (i1 + k < i1_end) {
            const int64_t idx = i + k * s1;
            gv[k] = ldro(g + idx);
            uv[k] = u ? ldro(u + idx) : T(1);
          }
#pragma unroll
        for (int k = 0; k < kEdhbBatch; ++k)
          if (i1 + k < i1_end) f[i + k * s1] = u ? gv[k] * uv[k] : gv[k];
      }
      return;
    }
    for (int i1 = i1_0; i1 < i1_end; i1 += kBatch, i += kBatch * s1, kw += kBatch * dk) {
      T gv[kBatch], uv[kBatch], fv[kBatch], fwv[kBatch], kv[kBatch], sv[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k)
        if (i1 + k < i1_end) {
          const int64_t idx = i + k * s1;
          gv[k] = ldro(g + idx);
          uv[k] = u ? ldro(u + idx) : T(1);
          fv[k] = f[idx];
          fwv[k] = fw[idx];
          kv[k] = ldro(kapw + kw + k * dk);
          sv[k] = ldro(sigw + kw + k * dk);
        }
#pragma unroll
      for (int k = 0; k < kBatch; ++k)
        if (i1 + k < i1_end) {
          const int64_t idx = i + k * s1;
          const T val = u ? gv[k] * uv[k] : gv[k];
          fw[idx] = val; // src/step_generic.cpp:596-602
          f[idx] = fv[k] + ((kv[k] + sv[k]) * val - (kv[k] - sv[k]) * fwv[k]);
        }
    }
    return;
  }
  if (J.u1 && J.u2 && !J.chi3 && !J.pmlw.sig) {
    // full 3x3 chi1inv outside PML (anisotropic media; src/step_generic.cpp:588-615): 14 loads and one
    // store per point, two planes in flight (same expression as edhb_point)
    T *f = (T *)J.f;
    const T *g = (const T *)J.g, *g1 = (const T *)J.g1, *g2 = (const T *)J.g2;
    const T *u = (const T *)J.u, *u1 = (const T *)J.u1, *u2 = (const T *)J.u2;
    const int64_t s = J.s, sa = J.s1, sb = J.s2;
    auto value = [&](int64_t q) {
      return ldro(g + q) * ldro(u + q) + offdiag(u1, g1, q, s, sa) + offdiag(u2, g2, q, s, sb);
    };
    int i1 = i1_0;
    for (; i1 + 1 < i1_end; i1 += 2, i += 2 * s1) {
      const T v0 = value(i), v1 = value(i + s1);
      f[i] = v0;
      f[i + s1] = v1;
    }
    if (i1 < i1_end) f[i] = value(i);
    return;
  }
  for (int i1 = i1_0; i1 < i1_end; ++i1, i += s1, kw += dk)
    edhb_point<T>(J, i, kw);
}

// lorentzian update_P: the isotropic case (src/susceptibility.cpp:251-257) is batched the same way
template <typename T>
MB200_HD void lorentz_thread(const mb200_lorentz_job_t &J, int64_t tile, int tid) {
  int i1_0, i1_end, i2, i3;
  if (!box_thread_point(J.box, tile, tid, i1_0, i1_end, i2, i3)) return;
  int64_t i = box_index(J.box, i1_0, i2, i3);
  const int64_t s1 = J.box.s[0];
  if (!J.s1) {
    T *p = (T *)J.p, *pp = (T *)J.pp;
    const T *w = (const T *)J.w, *s = (const T *)J.s;
    const T gamma1inv = (T)J.gamma1inv, gamma1 = (T)J.gamma1, omega0dtsqr = (T)J.omega0dtsqr,
            omega0dtsqr_denom = (T)J.omega0dtsqr_denom;
    for (int i1 = i1_0; i1 < i1_end; i1 += kBatch, i += kBatch * s1) {
      T pv[kBatch], ppv[kBatch], sv[kBatch], wv[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k)
        if (i1 + k < i1_end) {
          const int64_t idx = i + k * s1;
          pv[k] = p[idx];
          ppv[k] = pp[idx];
          sv[k] = ldro(s + idx);
          wv[k] = ldro(w + idx);
        }
#pragma unroll
      for (int k = 0; k < kBatch; ++k)
        if (i1 + k < i1_end) {
          const int64_t idx = i + k * s1;
          p[idx] = gamma1inv * (pv[k] * (2 - omega0dtsqr_denom) - gamma1 * ppv[k] +
                                omega0dtsqr * (sv[k] * wv[k]));
          pp[idx] = pv[k];
        }
    }
    return;
  }
  for (int i1 = i1_0; i1 < i1_end; ++i1, i += s1)
    lorentz_point<T>(J, i);
}

template <typename T> MB200_HD void beta_thread(const mb200_beta_job_t &J, int64_t tile, int tid) {
  int i1_0, i1_end, i2, i3;
  if (!box_thread_point(J.box, tile, tid, i1_0, i1_end, i2, i3)) return;
  int64_t i = box_index(J.box, i1_0, i2, i3);
  int k = pml_k(J.pml, i1_0, i2, i3), ku = pml_k(J.pmlu, i1_0, i2, i3);
  const int64_t s1 = J.box.s[0];
  const int dk = J.pml.ks[0], dku = J.pmlu.ks[0];
  const T fac = J.cyl ? (T)J.betadt / (T)(J.r_is2 + 2 * i2) : (T)J.betadt;
  for (int i1 = i1_0; i1 < i1_end; ++i1, i += s1, k += dk, ku += dku)
    beta_point<T>(J, i, k, ku, fac);
}

// list form: one thread per transfer.  Run-length form: the tiles after the PHASE tiles hold
// kHaloRunsPerTile runs each, one warp per run, lanes striding through its elements.
// one run per warp: the kernel is latency-bound (ncu: 92 % of cycles without an eligible warp, achieved
// occupancy 35 % of a theoretical 75 % with four runs per warp — a few long-running CTAs at the
// tail), so the work is spread over four times as many warps, each with eight loads in flight
constexpr int kHaloRunsPerWarp = 1;
constexpr int kHaloRunsPerTile = (kThreads / 32) * kHaloRunsPerWarp;
MB200_HD int64_t halo_list_tiles(const mb200_halo_job_t &J) {
  const int64_t n = J.nrun > 0 ? J.n_phase : J.n_phase + J.n_negate + J.n_copy;
  return (n + kThreads - 1) / kThreads;
}
template <typename T> MB200_HD void halo_thread(const mb200_halo_job_t &J, int64_t tile, int tid) {
  const int64_t lt = halo_list_tiles(J);
  if (tile < lt) {
    const int64_t n = tile * kThreads + tid;
    if (n < (J.nrun > 0 ? J.n_phase : halo_count(J))) halo_transfer<T>(J, n);
    return;
  }
  const int64_t r0 = (tile - lt) * kHaloRunsPerTile + (tid / 32) * kHaloRunsPerWarp;
  for (int64_t r = r0; r < r0 + kHaloRunsPerWarp && r < J.nrun; ++r) {
    const mb200_halo_run_t run = J.runs[r];
    // eight loads in flight per lane before the first store
    for (int e = tid % 32; e < run.n; e += 256) {
      T v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (e + 32 * q < run.n)
          v[q] = ldmut((const T *)(uintptr_t)(run.src0 + (int64_t)(e + 32 * q) * run.dsrc));
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (e + 32 * q < run.n)
          stout((T *)(uintptr_t)(run.dst0 + (int64_t)(e + 32 * q) * run.ddst), run.negate ? -v[q] : v[q]);
    }
  }
}

template <typename T> MB200_HD void gyro_thread(const mb200_gyro_job_t &J, int64_t tile, int tid) {
  int i1_0, i1_end, i2, i3;
  if (!box_thread_point(J.box, tile, tid, i1_0, i1_end, i2, i3)) return;
  int64_t i = box_index(J.box, i1_0, i2, i3);
  const int64_t s1 = J.box.s[0];
  for (int i1 = i1_0; i1 < i1_end; ++i1, i += s1)
    gyro_point<T>(J, i);
}

template <typename T>
MB200_HD void noise_thread(const mb200_noise_job_t &J, int64_t tile, int tid, const double *noise) {
  int i1_0, i1_end, i2, i3;
  if (!box_thread_point(J.box, tile, tid, i1_0, i1_end, i2, i3)) return;
  T *p = (T *)J.p;
  for (int i1 = i1_0; i1 < i1_end; ++i1) {
    const int64_t i = box_index(J.box, i1, i2, i3);
    const int64_t k = J.slot + ((int64_t)i1 * J.box.n[1] + i2) * J.box.n[2] + i3;
    p[i] = (T)((double)p[i] + noise[k]);
  }
}

template <typename T> MB200_HD void average_thread(const mb200_average_job_t &J, int64_t tile, int tid) {
  T *f = (T *)J.f;
  const T *b = (const T *)J.backup;
  for (int k = 0; k < kItems1D; ++k) {
    const int64_t i = (tile * kItems1D + k) * kThreads + tid;
    if (i < J.n) f[i] = (T)(0.5 * (f[i] + b[i]));
  }
}

template <typename T> MB200_HD void bfast_thread(const mb200_bfast_job_t &J, int64_t tile, int tid) {
  int i1_0, i1_end, i2, i3;
  if (!box_thread_point(J.box, tile, tid, i1_0, i1_end, i2, i3)) return;
  int64_t i = box_index(J.box, i1_0, i2, i3);
  int k = pml_k(J.pml, i1_0, i2, i3), ku = pml_k(J.pmlu, i1_0, i2, i3);
  const int64_t s1 = J.box.s[0];
  const int dk = J.pml.ks[0], dku = J.pmlu.ks[0];
  for (int i1 = i1_0; i1 < i1_end; ++i1, i += s1, k += dk, ku += dku)
    bfast_point<T>(J, i, k, ku);
}

template <typename T> MB200_HD void cylr0_thread(const mb200_cylr0_job_t &J, int64_t tile, int tid) {
  int i1_0, i1_end, i2, i3;
  if (!box_thread_point(J.box, tile, tid, i1_0, i1_end, i2, i3)) return;
  int64_t i = box_index(J.box, i1_0, i2, i3);
  int k = pml_k(J.pml, i1_0, i2, i3), ku = pml_k(J.pmlu, i1_0, i2, i3);
  const int64_t s1 = J.box.s[0];
  const int dk = J.pml.ks[0], dku = J.pmlu.ks[0];
  for (int i1 = i1_0; i1 < i1_end; ++i1, i += s1, k += dk, ku += dku)
    cylr0_point<T>(J, i, k, ku);
}

template <typename T> MB200_HD void cylint_thread(const mb200_cylint_job_t &J, int64_t tile, int tid) {
  const int64_t iz = tile * kThreads + tid;
  if (iz < J.sr) cylint_column<T>(J, iz);
}

// ---- zero-block variant of the isotropic Lorentz update (see mb200_lorentz_job_t) ------------------
// usable when the job runs over the standard 3-D layout: s = {(n2+1)(n3+1)-like, row, 1}
MB200_HD bool lorentz_blocked_ok(const mb200_lorentz_job_t &J) {
  return J.pzero && J.szero && !J.s1 && J.ntot > 0 && J.box.s[2] == 1 && J.box.s[1] > 0 &&
         J.box.s[0] > 0 && J.box.s[0] % J.box.s[1] == 0;
}
MB200_HD bool lorentz_owned(const mb200_lorentz_job_t &J, int64_t idx) {
  const int64_t s0 = J.box.s[0], s1 = J.box.s[1];
  const int64_t a1 = idx / s0, r = idx - a1 * s0, a2 = r / s1, a3 = r - a2 * s1;
  const int64_t l1 = J.box.idx0 / s0, lr = J.box.idx0 - l1 * s0, l2 = lr / s1, l3 = lr - l2 * s1;
  return a1 >= l1 && a1 < l1 + J.box.n[0] && a2 >= l2 && a2 < l2 + J.box.n[1] && a3 >= l3 &&
         a3 < l3 + J.box.n[2];
}
// one element; returns true if the (new) p and pp are both zero
template <typename T> MB200_HD bool lorentz_blocked_point(const mb200_lorentz_job_t &J, int64_t idx) {
  T *p = (T *)J.p, *pp = (T *)J.pp;
  if (!lorentz_owned(J, idx)) return p[idx] == T(0) && pp[idx] == T(0);
  const T pcur = p[idx];
  const T pn = (T)J.gamma1inv * (pcur * (2 - (T)J.omega0dtsqr_denom) - (T)J.gamma1 * pp[idx] +
                                 (T)J.omega0dtsqr * (ldro((const T *)J.s + idx) * ldro((const T *)J.w + idx)));
  p[idx] = pn;
  pp[idx] = pcur;
  return pn == T(0) && pcur == T(0);
}

#ifdef __CUDACC__

// ---- job lookup: tile_prefix[j] <= tile < tile_prefix[j+1] -------------------------------------
__device__ __forceinline__ int find_job(const int64_t *__restrict__ tile_prefix, int njobs,
                                        int64_t tile) {
  int lo = 0, hi = njobs; // invariant: prefix[lo] <= tile < prefix[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(tile_prefix + mid) <= tile) lo = mid;
    else hi = mid;
  }
  return lo;
}

template <typename JOB>
__device__ __forceinline__ void stage_job_at(JOB *sm, const JOB *__restrict__ jobs,
                                             const int64_t *__restrict__ tile_prefix, int njobs,
                                             int64_t tile, int64_t *tile_in_job) {
  __shared__ int s_job;
  __shared__ int64_t s_tile;
  if (threadIdx.x == 0) {
    const int j = find_job(tile_prefix, njobs, tile);
    s_job = j;
    s_tile = tile - __ldg(tile_prefix + j);
  }
  __syncthreads();
  const int *src = reinterpret_cast<const int *>(jobs + s_job);
  int *dst = reinterpret_cast<int *>(sm);
  for (int k = threadIdx.x; k < (int)(sizeof(JOB) / sizeof(int)); k += blockDim.x)
    dst[k] = __ldg(src + k);
  __syncthreads();
  *tile_in_job = s_tile;
}
template <typename JOB>
__device__ __forceinline__ void stage_job(JOB *sm, const JOB *__restrict__ jobs,
                                          const int64_t *__restrict__ tile_prefix, int njobs,
                                          int64_t *tile_in_job) {
  stage_job_at(sm, jobs, tile_prefix, njobs, (int64_t)blockIdx.x, tile_in_job);
}

// ---- step_curl ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    curl_kernel(const mb200_curl_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                int njobs) {
  __shared__ mb200_curl_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  curl_thread<T>(J, tile, threadIdx.x);
}

// ---- step_update_EDHB --------------------------------------------------------------------------
// group != NULL: group[j] is the first job of the component triple job j belongs to (group[j] == j and
// group[j + 1] != j for a single job); the CTAs of a triple are dealt out tile by tile (see plan_create)
template <typename T>
__global__ void __launch_bounds__(kThreads)
    edhb_kernel(const mb200_edhb_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                int njobs, const int *__restrict__ group) {
  __shared__ mb200_edhb_job_t J;
  __shared__ int s_job;
  __shared__ int64_t s_tile;
  if (threadIdx.x == 0) {
    int j = find_job(tile_prefix, njobs, (int64_t)blockIdx.x);
    int64_t t = (int64_t)blockIdx.x - __ldg(tile_prefix + j);
    if (group) {
      const int j0 = __ldg(group + j);
      if (j0 != j || (j + 1 < njobs && __ldg(group + j + 1) == j)) { // member of a triple
        const int64_t local = (int64_t)blockIdx.x - __ldg(tile_prefix + j0);
        j = j0 + (int)(local % 3);
        t = local / 3;
      }
    }
    s_job = j;
    s_tile = t;
  }
  __syncthreads();
  const int *src = reinterpret_cast<const int *>(jobs + s_job);
  int *dst = reinterpret_cast<int *>(&J);
  for (int k = threadIdx.x; k < (int)(sizeof(J) / sizeof(int)); k += blockDim.x)
    dst[k] = __ldg(src + k);
  __syncthreads();
  edhb_thread<T>(J, s_tile, threadIdx.x);
}

// ---- lorentzian update_P -----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    lorentz_kernel(const mb200_lorentz_job_t *__restrict__ jobs,
                   const int64_t *__restrict__ tile_prefix, int njobs) {
  __shared__ mb200_lorentz_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  lorentz_thread<T>(J, tile, threadIdx.x);
}

// ---- gyrotropic update_P ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    gyro_kernel(const mb200_gyro_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                int njobs) {
  __shared__ mb200_gyro_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  gyro_thread<T>(J, tile, threadIdx.x);
}

// ---- noise term of noisy_lorentzian_susceptibility ------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    noise_kernel(const mb200_noise_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                 int njobs, const double *__restrict__ noise) {
  __shared__ mb200_noise_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  noise_thread<T>(J, tile, threadIdx.x, noise);
}

// ---- average_with_backup ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    average_kernel(const mb200_average_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                   int njobs) {
  __shared__ mb200_average_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  average_thread<T>(J, tile, threadIdx.x);
}

// ---- step_bfast ----------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    bfast_kernel(const mb200_bfast_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                 int njobs) {
  __shared__ mb200_bfast_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  bfast_thread<T>(J, tile, threadIdx.x);
}

// ---- cylindrical coordinates --------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    cylint_kernel(const mb200_cylint_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                  int njobs) {
  __shared__ mb200_cylint_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  cylint_thread<T>(J, tile, threadIdx.x);
}
template <typename T>
__global__ void __launch_bounds__(kThreads)
    cylr0_kernel(const mb200_cylr0_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                 int njobs) {
  __shared__ mb200_cylr0_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  cylr0_thread<T>(J, tile, threadIdx.x);
}

// ---- step_beta ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    beta_kernel(const mb200_beta_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                int njobs) {
  __shared__ mb200_beta_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  beta_thread<T>(J, tile, threadIdx.x);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    lorentz_blocked_kernel(const mb200_lorentz_job_t *__restrict__ jobs,
                           const int64_t *__restrict__ tile_prefix, int njobs,
                           unsigned long long *__restrict__ work) {
  __shared__ mb200_lorentz_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  // a CTA walks kZBlocksPerCta consecutive blocks, so that skipped blocks cost a flag test and
  // not a CTA launch
  const int64_t nblocks = (J.ntot + MB200_ZBLOCK - 1) / MB200_ZBLOCK;
  int done = 0; // blocks actually updated (the measured-bytes accounting: work[0])
  // Everything the march needs from the descriptor, read once (the compiler cannot keep shared-memory
  // fields in registers across the barrier and the flag store of each block).  The owned-box test of
  // lorentz_owned in 32-bit arithmetic when the array has fewer than 2^31 elements: its 64-bit
  // divisions (two per element, three more per block for the box corner) were most of the kernel's
  // instructions — ~400 per element against 6 memory accesses.
  const int64_t ntot = J.ntot, s0 = J.box.s[0], s1 = J.box.s[1], idx0 = J.box.idx0;
  const bool small = ntot < ((int64_t)1 << 31);
  const int64_t c1 = idx0 / s0, cr = idx0 - c1 * s0, c2 = cr / s1, c3 = cr - c2 * s1;
  const unsigned us0 = (unsigned)s0, us1 = (unsigned)s1, l1 = (unsigned)c1, l2 = (unsigned)c2, l3 = (unsigned)c3;
  const unsigned n1 = (unsigned)J.box.n[0], n2 = (unsigned)J.box.n[1], n3 = (unsigned)J.box.n[2];
  auto owned = [&](int64_t idx) -> bool {
    if (small) {
      const unsigned u = (unsigned)idx, a1 = u / us0, r = u - a1 * us0, a2 = r / us1, a3 = r - a2 * us1;
      return a1 - l1 < n1 && a2 - l2 < n2 && a3 - l3 < n3; // (unsigned: below the corner wraps to huge)
    }
    const int64_t a1 = idx / s0, r = idx - a1 * s0, a2 = r / s1, a3 = r - a2 * s1;
    return a1 >= c1 && a1 < c1 + n1 && a2 >= c2 && a2 < c2 + n2 && a3 >= c3 && a3 < c3 + n3;
  };
  T *const p = (T *)J.p, *const pp = (T *)J.pp;
  const T *const sg = (const T *)J.s, *const wg = (const T *)J.w;
  const uint8_t *const szero = J.szero;
  uint8_t *const pzero = J.pzero;
  const T gamma1inv = (T)J.gamma1inv, gamma1 = (T)J.gamma1, omega0dtsqr = (T)J.omega0dtsqr,
          two_minus_denom = 2 - (T)J.omega0dtsqr_denom;
  for (int64_t b = tile * kZBlocksPerCta; b < (tile + 1) * kZBlocksPerCta && b < nblocks; ++b) {
    if (szero[b] && pzero[b]) continue; // sigma = P = P_prev = 0 here: nothing changes
    ++done;
    const int64_t base = b * MB200_ZBLOCK + threadIdx.x;
    bool zero = true;
    // the four elements of a thread: all loads before the first store (the stores to P / P_prev
    // would otherwise fence the loads of the next element: one memory latency per element)
    constexpr int kPer = MB200_ZBLOCK / kThreads;
    T pc[kPer], ppv[kPer], sv[kPer], wv[kPer];
    bool own[kPer], in[kPer];
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
      const int64_t idx = base + (int64_t)r * kThreads;
      in[r] = idx < ntot;
      own[r] = in[r] && owned(idx);
      pc[r] = in[r] ? p[idx] : T(0);
      ppv[r] = in[r] ? pp[idx] : T(0);
      sv[r] = own[r] ? ldro(sg + idx) : T(0);
      wv[r] = own[r] ? ldro(wg + idx) : T(0);
    }
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
      const int64_t idx = base + (int64_t)r * kThreads;
      if (!in[r]) continue;
      if (!own[r]) { // not-owned points are not updated; they still count for the zero flag
        zero = zero && pc[r] == T(0) && ppv[r] == T(0);
        continue;
      }
      // same expression as lorentz_blocked_point / src/susceptibility.cpp:251-257
      const T pn = gamma1inv * (pc[r] * two_minus_denom - gamma1 * ppv[r] + omega0dtsqr * (sv[r] * wv[r]));
      p[idx] = pn;
      pp[idx] = pc[r];
      zero = zero && pn == T(0) && pc[r] == T(0);
    }
    const int allzero = __syncthreads_and(zero ? 1 : 0);
    if (threadIdx.x == 0) pzero[b] = allzero ? 1 : 0;
  }
  if (threadIdx.x == 0 && done) atomicAdd(work, (unsigned long long)done);
}

template <typename T>
__global__ void block_zero_flags_kernel(const T *__restrict__ arr, int64_t n, uint8_t *flags) {
  const int64_t base = (int64_t)blockIdx.x * MB200_ZBLOCK + threadIdx.x;
  bool zero = true;
  for (int r = 0; r < MB200_ZBLOCK / kThreads; ++r) {
    const int64_t idx = base + (int64_t)r * kThreads;
    if (idx < n && arr[idx] != T(0)) zero = false;
  }
  const int allzero = __syncthreads_and(zero ? 1 : 0);
  if (threadIdx.x == 0) flags[blockIdx.x] = allzero ? 1 : 0;
}

// ---- 1-D jobs ----------------------------------------------------------------------------------

__device__ int g_fmp_simple = 0;
// A CTA takes kFmpBlocks consecutive zero-flag blocks (MB200_ZBLOCK elements each).  The flags of
// its blocks are read once, by the first threads, into shared memory: read per element they put a
// dependent byte load in front of every polarisation load (two memory latencies per element instead
// of one), and with one block per CTA the job look-up cost as much as the block itself.
template <typename T>
__global__ void __launch_bounds__(kThreads, 4)
    fmp_kernel(const mb200_fmp_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
               int njobs, unsigned long long *__restrict__ work) {
  __shared__ mb200_fmp_job_t J;
  __shared__ uint8_t s_zero[kFmpBlocks][MB200_MAX_P];
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  static_assert(kThreads * kItems1D == MB200_ZBLOCK, "kItems1D elements per thread and block");
  static_assert(kFmpBlocks * MB200_MAX_P <= kThreads, "one thread per (block, pole) flag");
  const int64_t ntot = J.ntot, nblocks = (ntot + MB200_ZBLOCK - 1) / MB200_ZBLOCK, b0 = tile * kFmpBlocks;
  const int np = J.np;
  if (threadIdx.x < kFmpBlocks * MB200_MAX_P) {
    const int blk = threadIdx.x / MB200_MAX_P, k = threadIdx.x % MB200_MAX_P;
    const int64_t b = b0 + blk;
    s_zero[blk][k] = (k < np && b < nblocks && J.pzero[k] && J.pzero[k][b]) ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) { // polarisation blocks this CTA reads (the others are known to be zero): work[1]
    int nread = 0;
    for (int blk = 0; blk < kFmpBlocks && b0 + blk < nblocks; ++blk)
      for (int k = 0; k < np; ++k)
        if (!s_zero[blk][k]) ++nread;
    if (nread) atomicAdd(work + 1, (unsigned long long)nread);
  }
  T *const fmp = (T *)J.fmp;
  const T *const d = (const T *)J.d;
  if (g_fmp_simple) { // MEEP_B200_FMP_SIMPLE=1: element by element as in round 1 (A/B switch)
    for (int blk = 0; blk < kFmpBlocks; ++blk)
#pragma unroll
      for (int r = 0; r < kItems1D; ++r) {
        const int64_t i = (b0 + blk) * MB200_ZBLOCK + threadIdx.x + (int64_t)r * kThreads;
        if (i < ntot) fmp_point<T>(J, i);
      }
    return;
  }
  for (int blk = 0; blk < kFmpBlocks; ++blk) {
    const int64_t base = (b0 + blk) * MB200_ZBLOCK + threadIdx.x;
    if (base - threadIdx.x >= ntot) break;
    // two elements at a time, all their loads first: 2 (1 + n_pol) values in flight per thread at four
    // CTAs per SM (four elements at once need 98 registers: two CTAs per SM, measured 30 % slower)
#pragma unroll
    for (int h = 0; h < kItems1D; h += 2) {
      T v[2], pv[2][MB200_MAX_P];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int64_t i = base + (int64_t)(h + r) * kThreads;
        const bool in = i < ntot;
        v[r] = in ? (d ? ldro(d + i) : fmp[i]) : T(0);
#pragma unroll
        for (int k = 0; k < MB200_MAX_P; ++k)
          pv[r][k] = (in && k < np && !s_zero[blk][k]) ? ldro((const T *)J.p[k] + i) : T(0);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int64_t i = base + (int64_t)(h + r) * kThreads;
        T x = v[r];
#pragma unroll
        for (int k = 0; k < MB200_MAX_P; ++k)
          if (k < np) x -= pv[r][k];
        if (i < ntot) fmp[i] = x;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    source_kernel(const mb200_src_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                  int njobs, const double *__restrict__ scalars) {
  __shared__ mb200_src_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  const int64_t j = tile * kThreads + threadIdx.x;
  if (j < J.npts) source_point<T>(J, j, scalars);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    halo_kernel(const mb200_halo_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                int njobs) {
  __shared__ mb200_halo_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  halo_thread<T>(J, tile, threadIdx.x);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    zero_kernel(const mb200_zero_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                int njobs) {
  __shared__ mb200_zero_job_t J;
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  const int64_t n = tile * kThreads + threadIdx.x;
  if (n < J.n) *(T *)(uintptr_t)J.ptrs[n] = T(0);
}

// ---- dft_chunk::update_dft ---------------------------------------------------------------------
// A CTA takes kDftPts consecutive monitor points (in IVEC_LOOP_COUNTER order): the first
// kDftPts threads form the weighted/averaged field values into shared memory, then all threads
// sweep the (point, frequency) plane in dft-array order, so the read-modify-write of the
// point-major/frequency-minor dft array is fully coalesced.

template <typename T>
__global__ void __launch_bounds__(kThreads)
    dft_kernel(const mb200_dft_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
               int njobs, const T *__restrict__ phases) {
  __shared__ mb200_dft_job_t J;
  __shared__ T s_fr[kDftPts], s_fi[kDftPts];
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  const int64_t npts = (int64_t)J.box.n[0] * J.box.n[1] * J.box.n[2];
  const int64_t p0 = tile * kDftPts;
  const int np = (int)min((int64_t)kDftPts, npts - p0);
  if ((int)threadIdx.x < np) {
    const int64_t p = p0 + threadIdx.x;
    const int n23 = J.box.n[1] * J.box.n[2];
    const int i1 = (int)(p / n23);
    const int r = (int)(p - (int64_t)i1 * n23);
    const int i2 = r / J.box.n[2], i3 = r - i2 * J.box.n[2];
    T fr, fi;
    dft_field_value<T>(J, i1, i2, i3, fr, fi);
    s_fr[threadIdx.x] = fr;
    s_fi[threadIdx.x] = fi;
  }
  __syncthreads();
  const int nomega = J.nomega;
  const bool is_complex = J.f_im != nullptr;
  const T *ph = phases + 2 * (int64_t)J.phase_slot;
  T *dft = (T *)J.dft + 2 * p0 * nomega;
  const int total = np * nomega;
  for (int e = threadIdx.x; e < total; e += kThreads) {
    const int lp = e / nomega, w = e - lp * nomega;
    dft_accumulate<T>(dft + 2 * (int64_t)e, is_complex, __ldg(ph + 2 * w), __ldg(ph + 2 * w + 1),
                      s_fr[lp], s_fi[lp]);
  }
}

// ---- dft_flux::flux inner sum ------------------------------------------------------------------
// Deterministic two-stage tree, no atomics.  Stage 1: a CTA takes kFluxPts consecutive monitor
// points of one job; warp v walks points v, v+8, ... with its lanes on 32 consecutive frequencies
// (the dft arrays are point-major/frequency-minor, so every access is a contiguous 32 x 2R line),
// accumulating Re(E conj(H)) in double; the eight warps' sums are combined through shared memory
// in a fixed order and written to partial[tile][frequency].  Stage 2: one warp per frequency
// strides over the tiles and folds its lanes with __shfl_down_sync.  The order of every addition
// is a function of the problem size only: two calls give bit-identical spectra.

template <typename T>
__global__ void __launch_bounds__(kThreads)
    flux_partial_kernel(const mb200_flux_job_t *__restrict__ jobs, const int64_t *__restrict__ tile_prefix,
                        int njobs, double *__restrict__ partial, int nomega_stride) {
  __shared__ mb200_flux_job_t J;
  __shared__ double s_acc[kThreads / 32][32];
  int64_t tile;
  stage_job(&J, jobs, tile_prefix, njobs, &tile);
  const int64_t p0 = tile * kFluxPts;
  const int np = (int)min((int64_t)kFluxPts, J.npts - p0);
  const T *e = (const T *)J.e + 2 * p0 * J.nomega, *h = (const T *)J.h + 2 * p0 * J.nomega;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  double *out = partial + (int64_t)blockIdx.x * nomega_stride;
  for (int w0 = 0; w0 < J.nomega; w0 += 32) {
    const int w = w0 + lane;
    double acc = 0;
    if (w < J.nomega)
      for (int p = warp; p < np; p += kThreads / 32) {
        const int64_t o = 2 * ((int64_t)p * J.nomega + w);
        // Re(E conj(H)) = Er*Hr + Ei*Hi, formed in realnum then widened (src/dft.cpp:547-550)
        acc += (double)(e[o] * h[o] + e[o + 1] * h[o + 1]);
      }
    s_acc[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && w < J.nomega) {
      double sum = s_acc[0][lane];
#pragma unroll
      for (int v = 1; v < kThreads / 32; ++v)
        sum += s_acc[v][lane];
      out[w] = sum;
    }
    __syncthreads();
  }
}

// out[w] += sum over tiles of partial[tile][w]; one warp per frequency
__global__ void __launch_bounds__(kThreads)
    flux_final_kernel(const double *__restrict__ partial, int64_t ntiles, int nomega, int nomega_stride,
                      double *__restrict__ out) {
  const int w = blockIdx.x * (kThreads / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (w >= nomega) return; // (whole warps leave together)
  double acc = 0;
  for (int64_t t = lane; t < ntiles; t += 32)
    acc += partial[t * nomega_stride + w];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    acc += __shfl_down_sync(0xffffffffu, acc, d);
  if (lane == 0) out[w] += acc;
}

// ---- finiteness probe --------------------------------------------------------------------------
template <typename T>
__global__ void check_finite_kernel(const uint64_t *__restrict__ ptrs, int64_t n, int32_t *flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const T v = *(const T *)(uintptr_t)ptrs[i];
    if (!isfinite(v)) *flag = 1;
  }
}

#endif // __CUDACC__

} // namespace mb200
#endif
