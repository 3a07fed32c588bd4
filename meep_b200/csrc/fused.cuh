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

#include <vector>

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
// planes of direction 0 marched per CTA: J.reserved (0 = default)
MB200_HD int step3_t1(const mb200_step3_job_t &J) { return J.reserved > 0 ? J.reserved : kT1; }
// does plane ix get the fused E/H epilogue? (not the planes that hold source points)
MB200_HD bool step3_epi_plane(const mb200_step3_job_t &J, int ix) {
  return ix < J.noepi_lo || ix >= J.noepi_lo + J.noepi_n;
}
MB200_HD int64_t step3_tiles(const mb200_step3_job_t &J) {
  return box_tiles(step3_box(J), step3_t1(J));
}
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
    int epi_arrays = 0;
    if (C.e) epi_arrays = 1 + (C.u ? 1 : 0) + (C.pmlw.sig ? 3 : 0);
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
    if (epi_arrays) { // the epilogue skips the source planes
      double frac = 1.0;
      int lo = C.lo[0] > J.ix_lo ? C.lo[0] : J.ix_lo, hi = C.hi[0] < J.ix_hi ? C.hi[0] : J.ix_hi;
      if (hi >= lo && J.noepi_n > 0) {
        const int last = J.noepi_lo + J.noepi_n - 1;
        const int nlo = J.noepi_lo > lo ? J.noepi_lo : lo, nhi = last < hi ? last : hi;
        if (nhi >= nlo) frac = 1.0 - (double)(nhi - nlo + 1) / (double)(hi - lo + 1);
      }
      bytes += R * epi_arrays * q * frac;
    }
  }
  return bytes;
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

// "Use" a loaded value without emitting an instruction: everything the compiler must have issued
// to produce the value stays above this point, so loads cannot be sunk below the first store of
// a marching step (they were, for the chi1inv operands: two memory latencies per plane instead of one).
MB200_HD void keep_above(double v) {
#if defined(__CUDA_ARCH__)
  asm volatile("" ::"d"(v));
#else
  (void)v;
#endif
}
MB200_HD void keep_above(float v) {
#if defined(__CUDA_ARCH__)
  asm volatile("" ::"f"(v));
#else
  (void)v;
#endif
}

// The march of an interior thread of the fast path (see step3_thread_full): all three components
// updated on every plane, cyclic curl operands G0..G2 with each centre value loaded once, one
// 32-bit cursor for every array (all arrays of a chunk share an index space of < 2^32 elements) so
// that an access is one IMAD.WIDE.U32 against a base held in the constant bank.  EPI: the diagonal
// E = chi1inv D (or H = B / mu) epilogue is fused; HASU: chi1inv is not identically 1.  NP x-planes
// are loaded before the first store (the kernel responds to loads in flight per thread).  Same
// arithmetic, in the same order, as step3_plain_general.
template <typename T, bool EPI, bool HASU, int NP>
MB200_HD void step3_plain_fast(const mb200_step3_job_t &J, int64_t i, int64_t sx, int ix0, int ix_end,
                               const bool (&metal_yz)[3]) {
  const T *G0 = (const T *)J.c[1].g1, *G1 = (const T *)J.c[2].g1, *G2 = (const T *)J.c[0].g1;
  const unsigned s10 = (unsigned)J.c[0].s1, s20 = (unsigned)J.c[0].s2, s11 = (unsigned)J.c[1].s1,
                 s21 = (unsigned)J.c[1].s2, s12 = (unsigned)J.c[2].s1, s22 = (unsigned)J.c[2].s2;
  T *f0 = (T *)J.c[0].f, *f1 = (T *)J.c[1].f, *f2 = (T *)J.c[2].f;
  const T *u0 = (const T *)J.c[0].u, *u1 = (const T *)J.c[1].u, *u2 = (const T *)J.c[2].u;
  T *e0 = (T *)J.c[0].e, *e1 = (T *)J.c[1].e, *e2 = (T *)J.c[2].e;
  const T k0 = (T)J.c[0].dtdx, k1 = (T)J.c[1].dtdx, k2 = (T)J.c[2].dtdx;
  const unsigned sxu = (unsigned)sx;
  struct Plane {
    T g0, g1, g2, a10, a20, a11, a21, a12, a22, v0, v1, v2, w0, w1, w2;
  };
  auto load = [&](unsigned q, Plane &p) {
    p.g0 = ldro(G0 + q), p.g1 = ldro(G1 + q), p.g2 = ldro(G2 + q);
    p.a10 = ldro(G2 + (q + s10)), p.a20 = ldro(G1 + (q + s20)); // component 0: g1 = G2, g2 = G1
    p.a11 = ldro(G0 + (q + s11)), p.a21 = ldro(G2 + (q + s21)); // component 1: g1 = G0, g2 = G2
    p.a12 = ldro(G1 + (q + s12)), p.a22 = ldro(G0 + (q + s22)); // component 2: g1 = G1, g2 = G0
    p.v0 = f0[q], p.v1 = f1[q], p.v2 = f2[q];
    p.w0 = p.w1 = p.w2 = T(1);
    if (HASU) p.w0 = ldro(u0 + q), p.w1 = ldro(u1 + q), p.w2 = ldro(u2 + q);
  };
  auto finish = [&](unsigned q, int ix, const Plane &p) {
    T dg = p.a10 - p.g2;
    dg = dg + p.g1 - p.a20;
    const T d0 = p.v0 - k0 * dg;
    dg = p.a11 - p.g0;
    dg = dg + p.g2 - p.a21;
    const T d1 = p.v1 - k1 * dg;
    dg = p.a12 - p.g1;
    dg = dg + p.g0 - p.a22;
    const T d2 = p.v2 - k2 * dg;
    f0[q] = d0;
    f1[q] = d1;
    f2[q] = d2;
    if (EPI && step3_epi_plane(J, ix)) {
      const T dd0 = metal_yz[0] ? T(0) : d0, dd1 = metal_yz[1] ? T(0) : d1, dd2 = metal_yz[2] ? T(0) : d2;
      e0[q] = HASU ? dd0 * p.w0 : dd0;
      e1[q] = HASU ? dd1 * p.w1 : dd1;
      e2[q] = HASU ? dd2 * p.w2 : dd2;
    }
  };
  unsigned q = (unsigned)i;
  int ix = ix0;
  if (NP > 1)
    for (; ix + NP <= ix_end; ix += NP, q += NP * sxu) {
      Plane p[NP];
#pragma unroll
      for (int k = 0; k < NP; ++k)
        load(q + k * sxu, p[k]);
      // every load stays above the first store (they were sunk below it otherwise: two memory
      // latencies per iteration instead of one)
      keep_above(p[NP - 1].a22);
      keep_above(p[NP - 1].v2);
      if (HASU) keep_above(p[NP - 1].w2);
#pragma unroll
      for (int k = 0; k < NP; ++k)
        finish(q + k * sxu, ix + k, p[k]);
    }
  for (; ix < ix_end; ++ix, q += sxu) {
    Plane p;
    load(q, p);
    keep_above(p.a22);
    keep_above(p.v2);
    if (HASU) keep_above(p.w2);
    finish(q, ix, p);
  }
}

// The general march of the fast path: per-plane masks (edge CTAs, planes outside a component's
// owned range, metal planes), all loads of a grid point before its first store.
template <typename T>
MB200_HD void step3_plain_general(const mb200_step3_job_t &J, int64_t i, int64_t sx, int ix0, int ix_end,
                                  const bool (&myz)[3], const bool (&metal_yz)[3]) {
  const int nlo = J.noepi_lo, nhi = J.noepi_lo + J.noepi_n;
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
        if (C.e && (ix < nlo || ix >= nhi)) {
          const bool metal = metal_yz[c] || ix == C.metal_lo[0] || ix == C.metal_hi[0];
          const T dd = metal ? T(0) : d;
          ((T *)C.e)[i] = C.u ? dd * uv[c] : dd;
        }
      }
    }
  }
}

// A "full" thread of the fast path: every component is updated on every plane it marches, no metal
// x-plane is crossed, and the curl operands are the three arrays of the other field type in cyclic
// order (g1 of component c = G[(c+2)%3], g2 = G[(c+1)%3]: src/fields.cpp:428-456), so the centre
// value of each is loaded once for the two components that use it: 15 loads and 6 stores per point
// instead of 18 and 6, no per-plane predicates, cursors instead of index arithmetic.  Full threads
// are marched by step3_lean_kernel, all others by the masked march (step3_plain_general).
MB200_HD bool step3_job_leanable(const mb200_step3_job_t &J) {
  bool ok = J.c[0].g1 == J.c[1].g2 && J.c[1].g1 == J.c[2].g2 && J.c[2].g1 == J.c[0].g2 &&
            (int64_t)(J.n[0] + 1) * (J.n[1] + 1) * (J.n[2] + 1) < ((int64_t)1 << 32);
  for (int c = 0; c < 3; ++c)
    ok = ok && (J.c[c].e != nullptr) == (J.c[0].e != nullptr) && (J.c[c].u != nullptr) == (J.c[0].u != nullptr);
  return ok;
}
// The full box of a job: the index box on which every thread is full — every component owned in
// all three directions, no metal x-plane inside (metal y / z planes are handled by the lean march).
// Empty (lo > hi in some direction) if there is none.
MB200_HD void step3_full_box(const mb200_step3_job_t &J, int (&lo)[3], int (&hi)[3]) {
  for (int d = 0; d < 3; ++d) {
    lo[d] = 0;
    hi[d] = J.n[d];
    for (int c = 0; c < 3; ++c) {
      if (J.c[c].lo[d] > lo[d]) lo[d] = J.c[c].lo[d];
      if (J.c[c].hi[d] < hi[d]) hi[d] = J.c[c].hi[d];
    }
  }
  if (J.ix_lo > lo[0]) lo[0] = J.ix_lo;
  if (J.ix_hi < hi[0]) hi[0] = J.ix_hi;
  // metal x-planes are boundary planes of the owned box: cut them off (one at a time is enough —
  // after a cut the next plane is tested against the shrunken range)
  for (int pass = 0; pass < 6; ++pass)
    for (int c = 0; c < 3; ++c) {
      const int m[2] = {J.c[c].metal_lo[0], J.c[c].metal_hi[0]};
      for (int k = 0; k < 2; ++k)
        if (m[k] >= lo[0] && m[k] <= hi[0]) {
          if (m[k] - lo[0] <= hi[0] - m[k]) lo[0] = m[k] + 1;
          else hi[0] = m[k] - 1;
        }
    }
}

// one thread of the masked march at loop point (ix0 .. ix_end, iy, iz) of the job (array indices)
template <typename T>
MB200_HD void step3_plain_column(const mb200_step3_job_t &J, int ix0, int ix_end, int iy, int iz) {
  const int64_t i = (int64_t)ix0 * J.stride[0] + (int64_t)iy * J.stride[1] + (int64_t)iz * J.stride[2];
  bool myz[3], metal_yz[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const mb200_step3_comp_t &C = J.c[c];
    myz[c] = iy >= C.lo[1] && iy <= C.hi[1] && iz >= C.lo[2] && iz <= C.hi[2];
    metal_yz[c] = iy == C.metal_lo[1] || iy == C.metal_hi[1] || iz == C.metal_lo[2] ||
                  iz == C.metal_hi[2];
  }
  step3_plain_general<T>(J, i, J.stride[0], ix0, ix_end, myz, metal_yz);
}

template <typename T>
MB200_HD void step3_plain_thread(const mb200_step3_job_t &J, int64_t tile, int tid) {
  const mb200_box_t box = step3_box(J);
  int ix0, ix_end, iy, iz;
  if (!box_thread_point(box, tile, tid, ix0, ix_end, iy, iz, step3_t1(J))) return;
  step3_plain_column<T>(J, ix0 + box.reserved, ix_end + box.reserved, iy, iz);
}

// The points of the full box of one job (descriptor in kernel-parameter space: every field is a
// constant-bank operand), tiled like the job itself with the march clipped to the full x-range.
// Planes in flight: two in double without the epilogue (the B half-step of a chunk whose H aliases
// B: 12 operands per plane), one with it; two in single precision.
template <typename T>
MB200_HD void step3_lean_thread(const mb200_step3_job_t &J, int64_t tile, int tid) {
  const mb200_box_t box = step3_box(J);
  int ix0, ix_end, iy, iz;
  if (!box_thread_point(box, tile, tid, ix0, ix_end, iy, iz, step3_t1(J))) return;
  ix0 += box.reserved;
  ix_end += box.reserved;
  int lo[3], hi[3];
  step3_full_box(J, lo, hi);
  if (iy < lo[1] || iy > hi[1] || iz < lo[2] || iz > hi[2]) return;
  if (ix0 < lo[0]) ix0 = lo[0];
  if (ix_end > hi[0] + 1) ix_end = hi[0] + 1;
  if (ix0 >= ix_end) return;
  const int64_t i = (int64_t)ix0 * J.stride[0] + (int64_t)iy * J.stride[1] + (int64_t)iz * J.stride[2];
  bool metal_yz[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const mb200_step3_comp_t &C = J.c[c];
    metal_yz[c] = iy == C.metal_lo[1] || iy == C.metal_hi[1] || iz == C.metal_lo[2] || iz == C.metal_hi[2];
  }
  constexpr int kB = 2, kE = sizeof(T) == 4 ? 2 : 1; // (measured: bench/micro/pml_shapes.cu "lean")
  const bool epi = J.c[0].e != nullptr, hasu = epi && J.c[0].u != nullptr;
  if (hasu) step3_plain_fast<T, true, true, kE>(J, i, J.stride[0], ix0, ix_end, metal_yz);
  else if (epi) step3_plain_fast<T, true, false, kE>(J, i, J.stride[0], ix0, ix_end, metal_yz);
  else step3_plain_fast<T, false, false, kB>(J, i, J.stride[0], ix0, ix_end, metal_yz);
}

// host side, at plan creation: the shell of a plain job
struct Step3Shell {
  bool lean;                           // the job has a non-empty full box and the lean march applies
  int lo[3], hi[3];                    // the full box
  std::vector<int> cols;               // (y, z) columns outside it
  std::vector<mb200_step3_job_t> slabs; // x-planes outside it (copies of the job with ix_lo / ix_hi set)
};
inline Step3Shell step3_shell(const mb200_step3_job_t &J) {
  Step3Shell S;
  step3_full_box(J, S.lo, S.hi);
  S.lean = step3_job_leanable(J) && S.lo[0] <= S.hi[0] && S.lo[1] <= S.hi[1] && S.lo[2] <= S.hi[2];
  if (!S.lean) return S;
  const int row = J.n[2] + 1;
  for (int iy = 0; iy <= J.n[1]; ++iy)
    for (int iz = 0; iz <= J.n[2]; ++iz)
      if (iy < S.lo[1] || iy > S.hi[1] || iz < S.lo[2] || iz > S.hi[2]) S.cols.push_back(iy * row + iz);
  const int first = J.ix_lo > 0 ? J.ix_lo : 0, last = J.ix_hi < J.n[0] ? J.ix_hi : J.n[0];
  if (S.lo[0] > first) {
    mb200_step3_job_t A = J;
    A.ix_lo = first;
    A.ix_hi = S.lo[0] - 1;
    S.slabs.push_back(A);
  }
  if (S.hi[0] < last) {
    mb200_step3_job_t B = J;
    B.ix_lo = S.hi[0] + 1;
    B.ix_hi = last;
    S.slabs.push_back(B);
  }
  return S;
}

// ---- general path --------------------------------------------------------------------------------
// Any mix of the 16 step_curl variants per component (PML in f, f_u level, conductivity with or
// without f_cond) plus the diagonal update_eh with or without the f_w ODE.  The variants are
// one formula with terms removed (reference src/step_generic.cpp:78-84); here the terms are
// switched by per-component flags that are uniform over the CTA, and — as in the fast path —
// every load of a grid point (all three components) is issued before its first store.
//
//   x      = fu if the f_u level exists, else f                       (the value the curl drives)
//   no PML : x' = CND ? ((1 - dt/2 cnd) x - curl) cndinv : x - curl                (lines 85-153)
//   PML    : x' = ((kap - sig) x - curl) siginv                                     (lines 179-193)
//            with conductivity: fcnd' = ((1 - dt/2 cnd) fcnd - curl) cndinv,
//                               x'    = ((kap - sig) x + (fcnd' - fcnd)) siginv     (lines 158-178)
//   f_u    : fu = x',  f' = siginvu ((kapu - sigu) f + x' - fu_old)                 (lines 112-153)
//   else   : f' = x'
template <typename T> struct Step3Vals {
  T f, a1, c1, c2, a2;       // field and the four curl operands
  T fu, fcnd, cnd, cndinv;   // aux levels
  T kms, sinv, kmsu, sinvu;  // (kap - sig), siginv for dsig and dsigu
  T u, fw, e, kapw, sigw;    // fused update_eh operands
};

template <typename T>
MB200_HD void step3_load(const mb200_step3_comp_t &C, int64_t i, int ix, int iy, int iz,
                         Step3Vals<T> &v) {
  const T *g1 = (const T *)C.g1, *g2 = (const T *)C.g2;
  v.f = ldmut((const T *)C.f + i);
  v.a1 = ldro(g1 + i + C.s1);
  v.c1 = ldro(g1 + i);
  v.c2 = ldro(g2 + i);
  v.a2 = ldro(g2 + i + C.s2);
  if (C.pmlu.sig) {
    const int ku = pml_k(C.pmlu, ix, iy, iz);
    v.fu = ldmut((const T *)C.fu + i);
    v.kmsu = ldro((const T *)C.pmlu.kap + ku) - ldro((const T *)C.pmlu.sig + ku);
    v.sinvu = ldro((const T *)C.pmlu.siginv + ku);
  }
  if (C.cnd) {
    v.cnd = ldro((const T *)C.cnd + i);
    v.cndinv = ldro((const T *)C.cndinv + i);
    if (C.pml.sig) v.fcnd = ldmut((const T *)C.fcnd + i);
  }
  if (C.pml.sig) {
    const int k = pml_k(C.pml, ix, iy, iz);
    v.kms = ldro((const T *)C.pml.kap + k) - ldro((const T *)C.pml.sig + k);
    v.sinv = ldro((const T *)C.pml.siginv + k);
  }
  if (C.e) {
    v.u = C.u ? ldro((const T *)C.u + i) : T(1);
    if (C.pmlw.sig) {
      const int kw = pml_k(C.pmlw, ix, iy, iz);
      v.fw = ldmut((const T *)C.fw + i);
      v.e = ldmut((const T *)C.e + i);
      v.kapw = ldro((const T *)C.pmlw.kap + kw);
      v.sigw = ldro((const T *)C.pmlw.sig + kw);
    }
  }
}

template <typename T>
MB200_HD void step3_compute_store(const mb200_step3_comp_t &C, int64_t i, bool metal, T dt2,
                                  const Step3Vals<T> &v, bool epi_plane = true) {
  T dg = v.a1 - v.c1;
  dg = dg + v.c2 - v.a2;
  const T curl = (T)C.dtdx * dg;
  const bool FU = C.pmlu.sig != nullptr, PML = C.pml.sig != nullptr, CND = C.cnd != nullptr;
  const T x = FU ? v.fu : v.f;
  T xn;
  if (!PML) {
    if (CND) xn = ((1 - dt2 * v.cnd) * x - curl) * v.cndinv;
    else xn = x - curl;
  }
  else if (CND) {
    const T fcn = ((1 - dt2 * v.cnd) * v.fcnd - curl) * v.cndinv;
    stout((T *)C.fcnd + i, fcn);
    xn = (v.kms * x + (fcn - v.fcnd)) * v.sinv;
  }
  else
    xn = (v.kms * x - curl) * v.sinv;
  T fn;
  if (FU) {
    stout((T *)C.fu + i, xn);
    fn = v.sinvu * (v.kmsu * v.f + xn - v.fu);
  }
  else
    fn = xn;
  stout((T *)C.f + i, fn);
  if (C.e && epi_plane) { // fused diagonal update_eh (src/step_generic.cpp:682-699, 774-782)
    const T d = metal ? T(0) : fn;
    const T val = C.u ? d * v.u : d;
    if (C.pmlw.sig) {
      stout((T *)C.fw + i, val);
      stout((T *)C.e + i, v.e + ((v.kapw + v.sigw) * val - (v.kapw - v.sigw) * v.fw));
    }
    else
      stout((T *)C.e + i, val);
  }
}

template <typename T>
MB200_HD void step3_thread(const mb200_step3_job_t &J, int64_t tile, int tid) {
  const mb200_box_t box = step3_box(J);
  int ix0, ix_end, iy, iz;
  if (!box_thread_point(box, tile, tid, ix0, ix_end, iy, iz, step3_t1(J))) return;
  const T dt2 = (T)J.dt * T(0.5);
  int64_t i = box_index(box, ix0, iy, iz);
  const int64_t sx = box.s[0];
  ix0 += box.reserved; // loop index -> array index along direction 0
  ix_end += box.reserved;
  bool myz[3], metal_yz[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const mb200_step3_comp_t &C = J.c[c];
    myz[c] = C.f && iy >= C.lo[1] && iy <= C.hi[1] && iz >= C.lo[2] && iz <= C.hi[2];
    metal_yz[c] = iy == C.metal_lo[1] || iy == C.metal_hi[1] || iz == C.metal_lo[2] ||
                  iz == C.metal_hi[2];
  }
  for (int ix = ix0; ix < ix_end; ++ix, i += sx) {
    Step3Vals<T> v[3];
    bool m[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      m[c] = myz[c] && ix >= J.c[c].lo[0] && ix <= J.c[c].hi[0];
      if (m[c]) step3_load<T>(J.c[c], i, ix, iy, iz, v[c]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (m[c])
        step3_compute_store<T>(J.c[c], i,
                               metal_yz[c] || ix == J.c[c].metal_lo[0] || ix == J.c[c].metal_hi[0],
                               dt2, v[c], step3_epi_plane(J, ix));
  }
}

// One component per thread (general kernel, split form).  The three components of a chunk are
// independent of each other inside one D/B(+E/H) pass — each reads only arrays of the other field
// type and writes its own f / f_u / f_cond / e / f_w — so the PML chunks can also be walked with
// one component per thread and three times as many CTAs: a third of the operands per thread
// (descriptor fields stay in registers across the march instead of being re-read from shared
// memory; twice the resident warps).
// The march of one thread, specialised at compile time on which auxiliary levels exist (the
// general kernel spends most of its issue slots on predicates and dead operands otherwise):
// PML = PML-in-f (dsig), FU = f_u level (dsigu), CND = conductivity, EPI = 0 no E/H epilogue,
// 1 diagonal update_eh, 2 diagonal update_eh through the f_w ODE.
template <typename T, bool PML, bool FU, bool CND, int EPI>
MB200_HD void step3c_march(const mb200_step3_job_t &J, const mb200_step3_comp_t &C, int64_t i,
                           int64_t sx, int ix0, int ix_end, int iy, int iz) {
  constexpr bool FW = EPI == 2;
  const bool HASU = EPI != 0 && C.u != nullptr;
  const T dtdx = (T)C.dtdx, dt2 = (T)J.dt * T(0.5);
  // everything the march needs, taken out of the (shared-memory) descriptor once: array cursors
  // at the first point, table cursors
  T *pf = (T *)C.f + i;
  const T *g1 = (const T *)C.g1 + i, *g2 = (const T *)C.g2 + i;
  const int64_t s1 = C.s1, s2 = C.s2;
  T *pfu = FU ? (T *)C.fu + i : nullptr;
  T *pfcnd = (CND && PML) ? (T *)C.fcnd + i : nullptr;
  const T *pcnd = CND ? (const T *)C.cnd + i : nullptr;
  const T *pcndinv = CND ? (const T *)C.cndinv + i : nullptr;
  T *pe = EPI ? (T *)C.e + i : nullptr;
  T *pfw = FW ? (T *)C.fw + i : nullptr;
  const T *pu = HASU ? (const T *)C.u + i : nullptr;
  // 1-D PML tables: cursor at ix0 and step per x-plane (0 unless the PML direction is x; then the
  // table values are the same for the whole march and are read once)
  const int dk = PML ? C.pml.ks[0] : 0, dku = FU ? C.pmlu.ks[0] : 0, dkw = FW ? C.pmlw.ks[0] : 0;
  const T *tsig = nullptr, *tkap = nullptr, *tsinv = nullptr;
  const T *tsigu = nullptr, *tkapu = nullptr, *tsinvu = nullptr, *tsigw = nullptr, *tkapw = nullptr;
  T kms = 0, sinv = 0, kmsu = 0, sinvu = 0, kapw = 0, sigw = 0;
  if (PML) {
    const int k = pml_k(C.pml, ix0, iy, iz);
    tsig = (const T *)C.pml.sig + k;
    tkap = (const T *)C.pml.kap + k;
    tsinv = (const T *)C.pml.siginv + k;
    kms = ldro(tkap) - ldro(tsig);
    sinv = ldro(tsinv);
  }
  if (FU) {
    const int ku = pml_k(C.pmlu, ix0, iy, iz);
    tsigu = (const T *)C.pmlu.sig + ku;
    tkapu = (const T *)C.pmlu.kap + ku;
    tsinvu = (const T *)C.pmlu.siginv + ku;
    kmsu = ldro(tkapu) - ldro(tsigu);
    sinvu = ldro(tsinvu);
  }
  if (FW) {
    const int kw = pml_k(C.pmlw, ix0, iy, iz);
    tsigw = (const T *)C.pmlw.sig + kw;
    tkapw = (const T *)C.pmlw.kap + kw;
    kapw = ldro(tkapw);
    sigw = ldro(tsigw);
  }
  bool metal_yz = false;
  int mlo = -1, mhi = -1;
  if (EPI) {
    metal_yz = iy == C.metal_lo[1] || iy == C.metal_hi[1] || iz == C.metal_lo[2] ||
               iz == C.metal_hi[2];
    mlo = C.metal_lo[0];
    mhi = C.metal_hi[0];
  }
  const int nlo = J.noepi_lo, nhi = J.noepi_lo + J.noepi_n; // planes without the epilogue (registers, not LDS per plane)

  for (int ix = ix0; ix < ix_end; ++ix) {
    // ---- all loads of this point (as step3_load) ----
    const T f = ldmut(pf);
    const T a1 = ldro(g1 + s1), c1 = ldro(g1), c2 = ldro(g2), a2 = ldro(g2 + s2);
    T fu = 0, fcnd = 0, cnd = 0, cndinv = 0, u = 1, fw = 0, e = 0;
    if (FU) fu = ldmut(pfu);
    if (CND) {
      cnd = ldro(pcnd);
      cndinv = ldro(pcndinv);
      if (PML) fcnd = ldmut(pfcnd);
    }
    if (HASU) u = ldro(pu);
    if (FW) {
      fw = ldmut(pfw);
      e = ldmut(pe);
    }
    if (PML && dk) {
      kms = ldro(tkap) - ldro(tsig);
      sinv = ldro(tsinv);
    }
    if (FU && dku) {
      kmsu = ldro(tkapu) - ldro(tsigu);
      sinvu = ldro(tsinvu);
    }
    if (FW && dkw) {
      kapw = ldro(tkapw);
      sigw = ldro(tsigw);
    }
    // ---- arithmetic and stores (as step3_compute_store) ----
    T dg = a1 - c1;
    dg = dg + c2 - a2;
    const T curl = dtdx * dg;
    const T x = FU ? fu : f;
    T xn;
    if (!PML) {
      if (CND) xn = ((1 - dt2 * cnd) * x - curl) * cndinv;
      else xn = x - curl;
    }
    else if (CND) {
      const T fcn = ((1 - dt2 * cnd) * fcnd - curl) * cndinv;
      stout(pfcnd, fcn);
      xn = (kms * x + (fcn - fcnd)) * sinv;
    }
    else
      xn = (kms * x - curl) * sinv;
    T fn;
    if (FU) {
      stout(pfu, xn);
      fn = sinvu * (kmsu * f + xn - fu);
    }
    else
      fn = xn;
    stout(pf, fn);
    if (EPI && (ix < nlo || ix >= nhi)) { // fused diagonal update_eh (src/step_generic.cpp:682-699, 774-782)
      const bool metal = metal_yz || ix == mlo || ix == mhi;
      const T d = metal ? T(0) : fn;
      const T val = HASU ? d * u : d;
      if (FW) {
        stout(pfw, val);
        stout(pe, e + ((kapw + sigw) * val - (kapw - sigw) * fw));
      }
      else
        stout(pe, val);
    }
    // ---- advance the cursors by one x-plane ----
    pf += sx;
    g1 += sx;
    g2 += sx;
    if (FU) pfu += sx;
    if (CND) {
      pcnd += sx;
      pcndinv += sx;
      if (PML) pfcnd += sx;
    }
    if (EPI) pe += sx;
    if (FW) pfw += sx;
    if (HASU) pu += sx;
    if (PML) {
      tsig += dk;
      tkap += dk;
      tsinv += dk;
    }
    if (FU) {
      tsigu += dku;
      tkapu += dku;
      tsinvu += dku;
    }
    if (FW) {
      tsigw += dkw;
      tkapw += dkw;
    }
  }
}

#ifdef __CUDACC__
__device__ int g_pml_pair = 4; // MEEP_B200_PML_PAIR=0|1: one x-plane per iteration for every variant (A/B switch)
// The same march with NP x-planes loaded before the first store.  The PML kernel is latency-bound
// (ncu, 512^3: issue slots 25 % busy, DRAM 52 %, DRAM bytes = 1.03 x algorithmic): with one component
// per thread a plane keeps only 6-13 loads in flight per thread against 15-18 in the fast path, so the
// register budget of the launch configuration is spent on planes: NP = budget / operands per point.
// All arrays of a chunk share one index space, so a single 32-bit cursor addresses every operand and
// the descriptor fields stay in shared memory (a 64-bit cursor per array, as in step3c_march, costs
// 2 registers per array: measured 4 % slower at one plane, bench/micro/pml_shapes.cu).
template <typename T> struct Step3cVals {
  T f, a1, c1, c2, a2, fu, fcnd, cnd, cndinv, u, fw, e, kms, sinv, kmsu, sinvu, kapw, sigw;
};
template <typename T, bool PML, bool FU, bool CND, int EPI, int NP>
__device__ __forceinline__ void step3c_multi_march(const mb200_step3_job_t &J, const mb200_step3_comp_t &C,
                                                   int64_t i, int64_t sx, int ix0, int ix_end, int iy, int iz) {
  constexpr bool FW = EPI == 2;
  const bool HASU = EPI != 0 && C.u != nullptr;
  const T dtdx = (T)C.dtdx, dt2 = (T)J.dt * T(0.5);
  const unsigned sxu = (unsigned)sx, s1 = (unsigned)C.s1, s2 = (unsigned)C.s2;
  const int dk = PML ? C.pml.ks[0] : 0, dku = FU ? C.pmlu.ks[0] : 0, dkw = FW ? C.pmlw.ks[0] : 0;
  int k = PML ? pml_k(C.pml, ix0, iy, iz) : 0, ku = FU ? pml_k(C.pmlu, ix0, iy, iz) : 0,
      kw = FW ? pml_k(C.pmlw, ix0, iy, iz) : 0;
  bool metal_yz = false;
  int mlo = -1, mhi = -1;
  if (EPI) {
    metal_yz = iy == C.metal_lo[1] || iy == C.metal_hi[1] || iz == C.metal_lo[2] || iz == C.metal_hi[2];
    mlo = C.metal_lo[0];
    mhi = C.metal_hi[0];
  }
  const int nlo = J.noepi_lo, nhi = J.noepi_lo + J.noepi_n;
  typedef Step3cVals<T> V;
  auto load = [&](unsigned q, int k, int ku, int kw, V &v) {
    v.f = ldmut((const T *)C.f + q);
    v.a1 = ldro((const T *)C.g1 + (q + s1));
    v.c1 = ldro((const T *)C.g1 + q);
    v.c2 = ldro((const T *)C.g2 + q);
    v.a2 = ldro((const T *)C.g2 + (q + s2));
    if (FU) {
      v.fu = ldmut((const T *)C.fu + q);
      v.kmsu = ldro((const T *)C.pmlu.kap + ku) - ldro((const T *)C.pmlu.sig + ku);
      v.sinvu = ldro((const T *)C.pmlu.siginv + ku);
    }
    if (CND) {
      v.cnd = ldro((const T *)C.cnd + q);
      v.cndinv = ldro((const T *)C.cndinv + q);
      if (PML) v.fcnd = ldmut((const T *)C.fcnd + q);
    }
    if (PML) {
      v.kms = ldro((const T *)C.pml.kap + k) - ldro((const T *)C.pml.sig + k);
      v.sinv = ldro((const T *)C.pml.siginv + k);
    }
    v.u = HASU ? ldro((const T *)C.u + q) : T(1);
    if (FW) {
      v.fw = ldmut((const T *)C.fw + q);
      v.e = ldmut((const T *)C.e + q);
      v.kapw = ldro((const T *)C.pmlw.kap + kw);
      v.sigw = ldro((const T *)C.pmlw.sig + kw);
    }
  };
  auto finish = [&](unsigned q, int ix, const V &v) { // arithmetic and stores: as step3c_march
    T dg = v.a1 - v.c1;
    dg = dg + v.c2 - v.a2;
    const T curl = dtdx * dg;
    const T x = FU ? v.fu : v.f;
    T xn;
    if (!PML) {
      if (CND) xn = ((1 - dt2 * v.cnd) * x - curl) * v.cndinv;
      else xn = x - curl;
    }
    else if (CND) {
      const T fcn = ((1 - dt2 * v.cnd) * v.fcnd - curl) * v.cndinv;
      stout((T *)C.fcnd + q, fcn);
      xn = (v.kms * x + (fcn - v.fcnd)) * v.sinv;
    }
    else
      xn = (v.kms * x - curl) * v.sinv;
    T fn;
    if (FU) {
      stout((T *)C.fu + q, xn);
      fn = v.sinvu * (v.kmsu * v.f + xn - v.fu);
    }
    else
      fn = xn;
    stout((T *)C.f + q, fn);
    if (EPI && (ix < nlo || ix >= nhi)) {
      const bool metal = metal_yz || ix == mlo || ix == mhi;
      const T d = metal ? T(0) : fn;
      const T val = HASU ? d * v.u : d;
      if (FW) {
        stout((T *)C.fw + q, val);
        stout((T *)C.e + q, v.e + ((v.kapw + v.sigw) * val - (v.kapw - v.sigw) * v.fw));
      }
      else
        stout((T *)C.e + q, val);
    }
  };
  unsigned q = (unsigned)i;
  int ix = ix0;
  if (NP > 1 && g_pml_pair > 1)
    for (; ix + NP <= ix_end; ix += NP, q += NP * sxu, k += NP * dk, ku += NP * dku, kw += NP * dkw) {
      V v[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p)
        load(q + p * sxu, k + p * dk, ku + p * dku, kw + p * dkw, v[p]);
#pragma unroll
      for (int p = 0; p < NP; ++p)
        finish(q + p * sxu, ix + p, v[p]);
    }
  for (; ix < ix_end; ++ix, q += sxu, k += dk, ku += dku, kw += dkw) {
    V a;
    load(q, k, ku, kw, a);
    finish(q, ix, a);
  }
}
#endif

// BUDGET: values a thread may hold in flight (what the registers of the launch configuration leave
// after addressing: see step3c_kernel); WIDE: a chunk with 2^32 or more elements per array (64-bit cursors)
template <typename T, bool PML, bool FU, bool CND, int EPI, int BUDGET, bool WIDE>
MB200_HD void step3c_dispatch(const mb200_step3_job_t &J, const mb200_step3_comp_t &C, int64_t i, int64_t sx,
                              int ix0, int ix_end, int iy, int iz) {
#ifdef __CUDA_ARCH__
  if (!WIDE) {
    constexpr int kOperands = 5 + (FU ? 3 : 0) + (CND ? (PML ? 3 : 2) : 0) + (PML ? 2 : 0) + (EPI ? 1 : 0) +
                              (EPI == 2 ? 4 : 0);
    constexpr int kPlanes = BUDGET / kOperands < 1 ? 1 : (BUDGET / kOperands > 4 ? 4 : BUDGET / kOperands);
    step3c_multi_march<T, PML, FU, CND, EPI, kPlanes>(J, C, i, sx, ix0, ix_end, iy, iz);
    return;
  }
#endif
  step3c_march<T, PML, FU, CND, EPI>(J, C, i, sx, ix0, ix_end, iy, iz);
}

template <typename T, int BUDGET = 0, bool WIDE = true>
MB200_HD void step3c_thread(const mb200_step3_job_t &J, int c, int64_t tile, int tid) {
  const mb200_box_t box = step3_box(J);
  int ix0, ix_end, iy, iz;
  if (!box_thread_point(box, tile, tid, ix0, ix_end, iy, iz, step3_t1(J))) return;
  const mb200_step3_comp_t &C = J.c[c];
  if (!C.f || iy < C.lo[1] || iy > C.hi[1] || iz < C.lo[2] || iz > C.hi[2]) return;
  int64_t i = box_index(box, ix0, iy, iz);
  const int64_t sx = box.s[0];
  ix0 += box.reserved; // loop index -> array index along direction 0
  ix_end += box.reserved;
  if (ix0 < C.lo[0]) {
    i += (int64_t)(C.lo[0] - ix0) * sx;
    ix0 = C.lo[0];
  }
  if (ix_end > C.hi[0] + 1) ix_end = C.hi[0] + 1;
  if (ix0 >= ix_end) return;
  const int epi = C.e ? (C.pmlw.sig ? 2 : 1) : 0;
  const int variant = (C.pml.sig ? 12 : 0) + (C.pmlu.sig ? 6 : 0) + (C.cnd ? 3 : 0) + epi;
  switch (variant) { // CTA-uniform
#define MB200_S3C(v)                                                                               \
  case v:                                                                                          \
    step3c_dispatch<T, ((v) / 12) != 0, (((v) / 6) % 2) != 0, (((v) / 3) % 2) != 0, (v) % 3, BUDGET, WIDE>( \
        J, C, i, sx, ix0, ix_end, iy, iz);                                                         \
    break;
    MB200_S3C(0) MB200_S3C(1) MB200_S3C(2) MB200_S3C(3) MB200_S3C(4) MB200_S3C(5)
    MB200_S3C(6) MB200_S3C(7) MB200_S3C(8) MB200_S3C(9) MB200_S3C(10) MB200_S3C(11)
    MB200_S3C(12) MB200_S3C(13) MB200_S3C(14) MB200_S3C(15) MB200_S3C(16) MB200_S3C(17)
    MB200_S3C(18) MB200_S3C(19) MB200_S3C(20) MB200_S3C(21) MB200_S3C(22) MB200_S3C(23)
#undef MB200_S3C
  }
}

// does any array of the job have 2^32 or more elements (the 32-bit cursors of the multi-plane march)?
inline bool step3_wide(const mb200_step3_job_t &J) {
  return (int64_t)(J.n[0] + 1) * (J.n[1] + 1) * (J.n[2] + 1) >= ((int64_t)1 << 32);
}

#ifdef __CUDACC__

// values in flight per thread that MINB CTAs of 256 threads per SM leave room for: 65536 / (256 MINB)
// registers, ~24 of them for addressing, bounds and constants; a double takes two
template <typename T> constexpr int step3c_budget(int minb) {
  return (65536 / (kThreads * minb) - (minb == 3 ? 29 : 24)) / (int)(sizeof(T) / 4);
}
template <typename T, int MINB, bool WIDE = false, int BUDGET = step3c_budget<T>(MINB)>
__global__ void __launch_bounds__(kThreads, MINB)
    step3c_kernel(const mb200_step3_job_t *__restrict__ jobs,
                  const int64_t *__restrict__ tile_prefix, int njobs) {
  __shared__ mb200_step3_job_t J;
  int64_t tile;
  // consecutive CTAs take the three components of the same tile: the curl operands they share
  // are fetched from DRAM once and found in L2 by the other two
  stage_job_at(&J, jobs, tile_prefix, njobs, (int64_t)(blockIdx.x / 3), &tile);
  step3c_thread<T, BUDGET, WIDE>(J, (int)(blockIdx.x % 3), tile, threadIdx.x);
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
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

// ---- the fast path as lean + shell launches per job -------------------------------------------------
// step3_lean_kernel marches the full box of ONE job (step3_full_box) with the lean march.  What is
// left is a shell one or two points thick: the x-planes below and above the full box go through the
// ordinary masked kernel as two slab jobs (ix_lo / ix_hi restricted copies of the job), and the
// (y, z) columns outside the full box — listed by the host — are marched over the full x-range by
// step3_cols_kernel, one thread per column.  The three launches write disjoint points and read only
// arrays that none of them writes.  History: one kernel holding both marches was slower than the masked
// march alone (joint register allocation); re-walking every TILE that holds a non-full thread with the
// masked march (16 % of the tiles at 512^3: the iz = 0 column sits in every first z-tile) cost more than
// the lean march gained (profiles/r2ab_*); with the shell as slabs + columns the 1024^3 step went
// 37.70 -> 36.8 ms with every job lean (profiles/r2ad_*), and the D-E half in double is left to the
// masked march (see launch_step3).
template <typename T>
__global__ void __launch_bounds__(kThreads, 4)
    step3_lean_kernel(const __grid_constant__ mb200_step3_job_t J) {
  step3_lean_thread<T>(J, (int64_t)blockIdx.x, threadIdx.x);
}
// cols[k] = iy * (n[2] + 1) + iz of the k-th shell column; blockIdx.y = chunk of planes of the full x-range
template <typename T>
__global__ void __launch_bounds__(kThreads)
    step3_cols_kernel(const mb200_step3_job_t *__restrict__ job, const int *__restrict__ cols, int ncols,
                      int x_lo, int x_hi, int planes) {
  __shared__ mb200_step3_job_t J;
  const int *src = reinterpret_cast<const int *>(job);
  int *dst = reinterpret_cast<int *>(&J);
  for (int k = threadIdx.x; k < (int)(sizeof(J) / sizeof(int)); k += blockDim.x)
    dst[k] = __ldg(src + k);
  __syncthreads();
  const int k = (int)blockIdx.x * kThreads + (int)threadIdx.x;
  if (k >= ncols) return;
  const int col = __ldg(cols + k), row = J.n[2] + 1;
  const int ix0 = x_lo + (int)blockIdx.y * planes;
  const int ix_end = ix0 + planes < x_hi + 1 ? ix0 + planes : x_hi + 1;
  if (ix0 < ix_end) step3_plain_column<T>(J, ix0, ix_end, col / row, col % row);
}
constexpr int kMaxJobLaunches = 8; // more plain jobs than this: one table-driven launch

// ---- job table in kernel-parameter (constant) space ---------------------------------------------
// The job descriptor is CTA-uniform.  Staging it in shared memory costs an LDS (and a short-
// scoreboard stall) for every pointer/flag use inside the marching loop; passed by value as a
// __grid_constant__ parameter it lives in the constant bank, where uniform operands are read by
// the uniform datapath without occupying the LSU.  CUDA >= 12.1 allows 32764 bytes of
// parameters, i.e. up to kParamJobs descriptors per launch (more jobs => more launches).
constexpr int kParamJobs = (32764 - 16) / (int)(sizeof(mb200_step3_job_t) + sizeof(int64_t)) - 1;
struct Step3Params {
  int njobs;
  int pad;
  int64_t prefix[kParamJobs + 1];
  mb200_step3_job_t jobs[kParamJobs];
};
static_assert(sizeof(Step3Params) <= 32764, "kernel parameter space exceeded");

template <typename T, bool PLAIN>
__global__ void __launch_bounds__(kThreads, PLAIN ? 1 : 2)
    step3_param_kernel(const __grid_constant__ Step3Params P) {
  int j = 0;
  while (j + 1 < P.njobs && P.prefix[j + 1] <= (int64_t)blockIdx.x)
    ++j;
  const int64_t tile = (int64_t)blockIdx.x - P.prefix[j];
  if (PLAIN) step3_plain_thread<T>(P.jobs[j], tile, threadIdx.x);
  else step3_thread<T>(P.jobs[j], tile, threadIdx.x);
}

// host copies of the job table / prefix are needed to pass them by value
template <typename T>
static void launch_step3_params(const mb200_step3_job_t *h_jobs, const int64_t *h_prefix, int njobs,
                                bool all_plain, cudaStream_t s) {
  for (int j0 = 0; j0 < njobs; j0 += kParamJobs) {
    Step3Params P;
    P.njobs = njobs - j0 < kParamJobs ? njobs - j0 : kParamJobs;
    P.pad = 0;
    for (int k = 0; k <= P.njobs; ++k)
      P.prefix[k] = h_prefix[j0 + k] - h_prefix[j0];
    for (int k = 0; k < P.njobs; ++k)
      P.jobs[k] = h_jobs[j0 + k];
    const int64_t tiles = P.prefix[P.njobs];
    if (tiles <= 0) continue;
    if (all_plain)
      step3_param_kernel<T, true><<<dim3((unsigned)tiles), dim3(kThreads), 0, s>>>(P);
    else
      step3_param_kernel<T, false><<<dim3((unsigned)tiles), dim3(kThreads), 0, s>>>(P);
  }
}

// what plan_create prepares for the lean + shell launches of an all-plain plan
struct Step3LeanPlan {
  struct Job {
    bool lean;
    int x_lo, x_hi, ncols;
    const int *d_cols;
    int nslabs;
    const mb200_step3_job_t *d_slabs; // [nslabs]
    const int64_t *d_slab_prefix;     // [nslabs + 1]
    const int64_t *d_own_prefix;      // {0, tiles of the job}: the masked march over the whole job
    int64_t slab_tiles;
  };
  std::vector<Job> jobs;
};

// returns the number of kernels launched
template <typename T>
static int launch_step3(const mb200_step3_job_t *jobs, const int64_t *prefix, int njobs,
                        int64_t tiles, bool all_plain, int split, cudaStream_t s,
                        const mb200_step3_job_t *h_jobs, const int64_t *h_prefix, const Step3LeanPlan *lean) {
  if (all_plain && lean && (int)lean->jobs.size() == njobs) {
    int launched = 0;
    for (int j = 0; j < njobs; ++j) {
      const Step3LeanPlan::Job &L = lean->jobs[j];
      const int64_t t = h_prefix[j + 1] - h_prefix[j];
      if (t <= 0) continue;
      // Where the lean march pays (B200, bench/micro/pml_shapes.cu "lean", 492^3 interior; lean + shell
      // against the masked march): double without the epilogue -9.8 %, double with it +0.7 %, single
      // -17 % and -6 %.  So: every job in single precision, the jobs without the epilogue in double.
      const bool pays = sizeof(T) == 4 || h_jobs[j].c[0].e == nullptr;
      if (!L.lean || !pays) { // (or no full box / not the cyclic operand layout): the masked march for the whole job
        step3_plain_kernel<T><<<dim3((unsigned)t), dim3(kThreads), 0, s>>>(jobs + j, L.d_own_prefix, 1);
        ++launched;
        continue;
      }
      step3_lean_kernel<T><<<dim3((unsigned)t), dim3(kThreads), 0, s>>>(h_jobs[j]);
      ++launched;
      if (L.slab_tiles > 0) {
        step3_plain_kernel<T><<<dim3((unsigned)L.slab_tiles), dim3(kThreads), 0, s>>>(L.d_slabs, L.d_slab_prefix, L.nslabs);
        ++launched;
      }
      if (L.ncols > 0) {
        ++launched;
        const int planes = step3_t1(h_jobs[j]);
        const dim3 grid((unsigned)((L.ncols + kThreads - 1) / kThreads), (unsigned)((L.x_hi - L.x_lo + planes) / planes));
        step3_cols_kernel<T><<<grid, dim3(kThreads), 0, s>>>(jobs + j, L.d_cols, L.ncols, L.x_lo, L.x_hi, planes);
      }
    }
    return launched;
  }
  else if (all_plain)
    step3_plain_kernel<T><<<dim3((unsigned)tiles), dim3(kThreads), 0, s>>>(jobs, prefix, njobs);
  else if (split) { // MEEP_B200_SPLIT_PML=2..6: CTAs per SM (128 / 85 / 64 / 51 / 42 registers per thread)
    bool wide = false;
    for (int j = 0; j < njobs; ++j)
      wide = wide || step3_wide(h_jobs[j]);
    const dim3 grid((unsigned)(3 * tiles)), block(kThreads);
    if (wide) step3c_kernel<T, 4, true><<<grid, block, 0, s>>>(jobs, prefix, njobs);
    else if (split == 2) step3c_kernel<T, 2><<<grid, block, 0, s>>>(jobs, prefix, njobs);
    else if (split == 3) step3c_kernel<T, 3><<<grid, block, 0, s>>>(jobs, prefix, njobs);
    else if (split == 5) step3c_kernel<T, 5><<<grid, block, 0, s>>>(jobs, prefix, njobs);
    else if (split == 6) step3c_kernel<T, 6><<<grid, block, 0, s>>>(jobs, prefix, njobs);
    else step3c_kernel<T, 4><<<grid, block, 0, s>>>(jobs, prefix, njobs);
  }
  else
    step3_kernel<T><<<dim3((unsigned)tiles), dim3(kThreads), 0, s>>>(jobs, prefix, njobs);
  return 1;
}

#endif // __CUDACC__

} // namespace mb200
#endif
