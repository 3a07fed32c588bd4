// point_ops.h — per-grid-point arithmetic of the FDTD hot path, written once and compiled
// (a) as __device__ code into the sm_100a kernels (kernels.cuh) and
// (b) as host code into the test-only emulator (tests/emu), which exists so that the host-side
//     engine logic can be exercised in a container without a GPU.  The shipped library contains
//     no host execution path.
//
// Each function restates one inner-loop body of the reference; citations give file:line in the
// reference repository.  Expression order follows the reference so that differences are limited
// to FMA contraction (rel-L2 ~1e-16/step, far inside the 1e-12 parity gate).
#ifndef MEEP_B200_POINT_OPS_H
#define MEEP_B200_POINT_OPS_H

#include <stdint.h>
#include <math.h>
#include "../../include/meep_b200.h"

#ifdef __CUDACC__
#define MB200_HD __host__ __device__ __forceinline__
#else
#define MB200_HD inline
#endif

namespace mb200 {

// load from an array that no kernel of the current launch writes (the "other" field type,
// materials, PML tables): ld.global.nc lets the compiler hoist these loads above the stores of
// the arrays being updated, which is what keeps enough bytes in flight per warp.
template <typename T> MB200_HD T ldro(const T *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// Loads / stores of the arrays a kernel updates.  Job descriptors hand us generic pointers; every
// array lives in global memory (mb200_malloc), and saying so (ld.global / st.global instead of
// generic LD / ST) lets the compiler keep CTA-uniform descriptor fields, which sit in shared
// memory, in registers across the stores of a marching loop instead of re-reading them.
template <typename T> MB200_HD T ldmut(const T *p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}
template <typename T> MB200_HD void stout(T *p, T v) {
#if defined(__CUDA_ARCH__)
  __stcg(p, v);
#else
  *p = v;
#endif
}

// index of loop point (i1,i2,i3) in a box
MB200_HD int64_t box_index(const mb200_box_t &b, int i1, int i2, int i3) {
  return b.idx0 + (int64_t)i1 * b.s[0] + (int64_t)i2 * b.s[1] + (int64_t)i3 * b.s[2];
}

// KDEF (src/meep_internals.hpp:222-223)
MB200_HD int pml_k(const mb200_pml_t &p, int i1, int i2, int i3) {
  return ((p.k0 + p.ks[0] * i1) + p.ks[1] * i2) + p.ks[2] * i3;
}

// ------------------------------------------------------------------------------------------------
// step_curl (src/step_generic.cpp:65-249): the 16 specialised loops are this one body with
// terms removed.  PML: dsig != NO_DIRECTION; FU: dsigu != NO_DIRECTION; CND: cnd != NULL;
// G2: second curl term present.  Callers have already applied the g1==NULL swap (lines 72-76).
// curl_apply: the read-modify-write of f / fu / fcnd given curl = dtdx*(dg1 - dg2).
// JOB is mb200_curl_job_t or mb200_step3_comp_t (same member names).  Returns the new f[i].
template <typename T, bool PML, bool FU, bool CND, typename JOB>
MB200_HD T curl_apply(const JOB &J, int64_t i, int k, int ku, T curl, T dt2) {
  T *f = (T *)J.f;
  T fnew;
  if (!PML) {
    if (!FU) {
      if (CND) { // lines 90-99
        const T *cnd = (const T *)J.cnd, *cndinv = (const T *)J.cndinv;
        fnew = ((1 - dt2 * ldro(cnd + i)) * f[i] - curl) * ldro(cndinv + i);
      }
      else // lines 102-109
        fnew = f[i] - curl;
    }
    else { // lines 112-153
      T *fu = (T *)J.fu;
      const T *sigu = (const T *)J.pmlu.sig, *kapu = (const T *)J.pmlu.kap,
              *siginvu = (const T *)J.pmlu.siginv;
      const T fprev = fu[i];
      T fun;
      if (CND) {
        const T *cnd = (const T *)J.cnd, *cndinv = (const T *)J.cndinv;
        fun = ((1 - dt2 * ldro(cnd + i)) * fprev - curl) * ldro(cndinv + i);
      }
      else
        fun = fprev - curl;
      fu[i] = fun;
      fnew = ldro(siginvu + ku) * ((ldro(kapu + ku) - ldro(sigu + ku)) * f[i] + fun - fprev);
    }
  }
  else {
    const T *sig = (const T *)J.pml.sig, *kap = (const T *)J.pml.kap,
            *siginv = (const T *)J.pml.siginv;
    if (!FU) { // lines 157-194
      if (CND) {
        const T *cnd = (const T *)J.cnd, *cndinv = (const T *)J.cndinv;
        T *fcnd = (T *)J.fcnd;
        const T fcnd_prev = fcnd[i];
        const T fcn = ((1 - dt2 * ldro(cnd + i)) * fcnd_prev - curl) * ldro(cndinv + i);
        fcnd[i] = fcn;
        fnew = ((ldro(kap + k) - ldro(sig + k)) * f[i] + (fcn - fcnd_prev)) * ldro(siginv + k);
      }
      else
        fnew = ((ldro(kap + k) - ldro(sig + k)) * f[i] - curl) * ldro(siginv + k);
    }
    else { // lines 195-247 (most general case 201-211)
      T *fu = (T *)J.fu;
      const T *sigu = (const T *)J.pmlu.sig, *kapu = (const T *)J.pmlu.kap,
              *siginvu = (const T *)J.pmlu.siginv;
      const T fprev = fu[i];
      T fun;
      if (CND) {
        const T *cnd = (const T *)J.cnd, *cndinv = (const T *)J.cndinv;
        T *fcnd = (T *)J.fcnd;
        const T fcnd_prev = fcnd[i];
        const T fcn = ((1 - dt2 * ldro(cnd + i)) * fcnd_prev - curl) * ldro(cndinv + i);
        fcnd[i] = fcn;
        fun = ((ldro(kap + k) - ldro(sig + k)) * fprev + (fcn - fcnd_prev)) * ldro(siginv + k);
      }
      else
        fun = ((ldro(kap + k) - ldro(sig + k)) * fprev - curl) * ldro(siginv + k);
      fu[i] = fun;
      fnew = ldro(siginvu + ku) * ((ldro(kapu + ku) - ldro(sigu + ku)) * f[i] + fun - fprev);
    }
  }
  f[i] = fnew;
  return fnew;
}

// the g-difference of step_curl: g1[i+s1] - g1[i] + g2[i] - g2[i+s2]
template <typename T, bool G2, typename JOB> MB200_HD T curl_dg(const JOB &J, int64_t i) {
  const T *g1 = (const T *)J.g1;
  T dg = ldro(g1 + i + J.s1) - ldro(g1 + i);
  if (G2) {
    const T *g2 = (const T *)J.g2;
    dg = dg + ldro(g2 + i) - ldro(g2 + i + J.s2);
  }
  return dg;
}

template <typename T, bool PML, bool FU, bool CND, bool G2, typename JOB>
MB200_HD T curl_point(const JOB &J, int64_t i, int k, int ku, T dtdx, T dt2) {
  return curl_apply<T, PML, FU, CND>(J, i, k, ku, dtdx * curl_dg<T, G2>(J, i), dt2);
}

// runtime-flag front end (flags are uniform over a job)
template <typename JOB> MB200_HD int curl_variant(const JOB &J) {
  return (J.pml.sig ? 8 : 0) | (J.pmlu.sig ? 4 : 0) | (J.cnd ? 2 : 0) | (J.g2 ? 1 : 0);
}

template <typename T, typename JOB>
MB200_HD T curl_point_any(const JOB &J, int variant, int64_t i, int k, int ku, T dtdx, T dt2) {
  switch (variant) {
#define MB200_CASE(v)                                                                              \
  case v:                                                                                          \
    return curl_point<T, ((v)&8) != 0, ((v)&4) != 0, ((v)&2) != 0, ((v)&1) != 0>(J, i, k, ku,      \
                                                                                 dtdx, dt2);
    MB200_CASE(0) MB200_CASE(1) MB200_CASE(2) MB200_CASE(3) MB200_CASE(4) MB200_CASE(5)
    MB200_CASE(6) MB200_CASE(7) MB200_CASE(8) MB200_CASE(9) MB200_CASE(10) MB200_CASE(11)
    MB200_CASE(12) MB200_CASE(13) MB200_CASE(14) MB200_CASE(15)
#undef MB200_CASE
  }
  return T(0);
}

// ------------------------------------------------------------------------------------------------
// step_beta (src/step_generic.cpp:255-333): the eight specialised loops as one body
// (fac = betadt, or the_m / r for the cylindrical i*m/r terms of src/step_db.cpp:178-280)
template <typename T>
MB200_HD void beta_point(const mb200_beta_job_t &J, int64_t i, int k, int ku, T fac) {
  T *f = (T *)J.f;
  const T *g = (const T *)J.g;
  T df = fac * ldro(g + i);
  if (J.cndinv) df = df * ldro((const T *)J.cndinv + i);
  if (J.pml.siginv) {
    if (J.cndinv) ((T *)J.fcnd)[i] += df;
    df = df * ldro((const T *)J.pml.siginv + k);
  }
  if (J.pmlu.siginv) {
    ((T *)J.fu)[i] += df;
    f[i] += ldro((const T *)J.pmlu.siginv + ku) * df;
  }
  else
    f[i] += df;
}

// ------------------------------------------------------------------------------------------------
// gyrotropic_susceptibility::update_P (src/susceptibility.cpp:445-584), one loop point
// OFFDIAGW (line 443)
template <typename T> MB200_HD T offdiagw(const T *g, int64_t i, int64_t sx, int64_t s) {
  return T(0.25) * (ldro(g + i) + ldro(g + i - sx) + ldro(g + i + s) + ldro(g + (i + s) - sx));
}
template <typename T> MB200_HD void gyro_point(const mb200_gyro_job_t &J, int64_t i) {
  T *p0 = (T *)J.p[0], *p1 = (T *)J.p[1], *p2 = (T *)J.p[2];
  T *pp0 = (T *)J.pp[0], *pp1 = (T *)J.pp[1], *pp2 = (T *)J.pp[2];
  const T *w0 = (const T *)J.w[0], *w1 = (const T *)J.w[1], *w2 = (const T *)J.w[2];
  const T si = ldro((const T *)J.s + i);
  const T P0 = p0[i], P1 = p1[i], P2 = p2[i], Q0 = pp0[i], Q1 = pp1[i], Q2 = pp2[i];
  const T g01 = (T)J.gt[0][1], g02 = (T)J.gt[0][2], g10 = (T)J.gt[1][0], g12 = (T)J.gt[1][2],
          g20 = (T)J.gt[2][0], g21 = (T)J.gt[2][1];
  T r0, r1, r2;
  if (J.model == 0) {
    const T diag = (T)J.c[0], gamma1 = (T)J.c[1], omega0dtsqr = (T)J.c[2], pt = (T)J.c[3];
    r0 = diag * P0 - gamma1 * Q0 + omega0dtsqr * si * ldro(w0 + i) - pt * g01 * Q1 - pt * g02 * Q2;
    r1 = diag * P1 - gamma1 * Q1 + (w1 ? omega0dtsqr * si * offdiagw(w1, i, J.is1, J.is) : T(0)) -
         pt * g10 * Q0 - pt * g12 * Q2;
    r2 = diag * P2 - gamma1 * Q2 + (w2 ? omega0dtsqr * si * offdiagw(w2, i, J.is2, J.is) : T(0)) -
         pt * g21 * Q1 - pt * g20 * Q0;
  }
  else {
    const T omega2pidt = (T)J.c[0], g2pidt = (T)J.c[1], alpha = (T)J.c[2], dt2pi = (T)J.c[3];
    const T q0 = -omega2pidt * P0 + T(0.5) * alpha * Q0 + dt2pi * si * ldro(w0 + i);
    const T q1 = -omega2pidt * P1 + T(0.5) * alpha * Q1 +
                 dt2pi * si * (w1 ? offdiagw(w1, i, J.is1, J.is) : T(0));
    const T q2 = -omega2pidt * P2 + T(0.5) * alpha * Q2 +
                 dt2pi * si * (w2 ? offdiagw(w2, i, J.is2, J.is) : T(0));
    r0 = T(0.5) * Q0 - g2pidt * P0 + g01 * q1 + g02 * q2;
    r1 = T(0.5) * Q1 - g2pidt * P1 + g12 * q2 + g10 * q0;
    r2 = T(0.5) * Q2 - g2pidt * P2 + g20 * q0 + g21 * q1;
  }
  pp0[i] = P0;
  pp1[i] = P1;
  pp2[i] = P2;
  p0[i] = (T)J.inv[0][0] * r0 + (T)J.inv[0][1] * r1 + (T)J.inv[0][2] * r2;
  p1[i] = (T)J.inv[1][0] * r0 + (T)J.inv[1][1] * r1 + (T)J.inv[1][2] * r2;
  p2[i] = (T)J.inv[2][0] * r0 + (T)J.inv[2][1] * r1 + (T)J.inv[2][2] * r2;
}

// ------------------------------------------------------------------------------------------------
// step_bfast (src/step_generic.cpp:335-530): the sixteen specialised loops as one body
template <typename T> MB200_HD void bfast_point(const mb200_bfast_job_t &J, int64_t i, int k, int ku) {
  T *f = (T *)J.f, *F = (T *)J.F;
  const T *g1 = (const T *)J.g1, *g2 = (const T *)J.g2;
  const T k1 = (T)J.k1, k2 = (T)J.k2;
  const T F_prev = F[i];
  T Fn;
  if (g2)
    Fn = (k1 * (ldro(g1 + i + J.s1) + ldro(g1 + i)) - k2 * (ldro(g2 + i + J.s2) + ldro(g2 + i))) - F_prev;
  else if (!J.pml.siginv && !J.pmlu.siginv && !J.cnd)
    Fn = k1 * (ldro(g1 + i + J.s1) + ldro(g1 + i)); // (line 372: this variant does not subtract F)
  else
    Fn = k1 * (ldro(g1 + i + J.s1) + ldro(g1 + i)) - F_prev;
  F[i] = Fn;
  T df = Fn - F_prev;
  if (J.cnd) df = df * ldro((const T *)J.cndinv + i);
  if (J.pml.siginv) {
    if (J.cnd) ((T *)J.fcnd)[i] += df;
    df = df * ldro((const T *)J.pml.siginv + k);
  }
  if (J.pmlu.siginv) {
    ((T *)J.fu)[i] += df;
    f[i] += ldro((const T *)J.pmlu.siginv + ku) * df;
  }
  else
    f[i] += df;
}

// ------------------------------------------------------------------------------------------------
// cylindrical r = 0 row (src/step_db.cpp:300-321 and 350-371): one body for both loops
template <typename T> MB200_HD void cylr0_point(const mb200_cylr0_job_t &J, int64_t i, int k, int ku) {
  T *the_f = (T *)J.f, *fu = (T *)J.fu, *fcnd = (T *)J.fcnd;
  const T *fp = (const T *)J.fp, *fm = (const T *)J.fm;
  const T fprev = the_f[i];
  T dfcnd;
  if (J.mode == 0)
    dfcnd = (T)((double)ldro(fp + i) * J.c);
  else
    dfcnd = (T)(J.c * (double)(ldro(fp + i) - ldro(fp + i - J.sd) - (T)J.mult * ldro(fm + i)));
  if (fcnd) {
    const T dt2 = (T)(J.dt * 0.5);
    const T fcnd_prev = fcnd[i];
    fcnd[i] = ((1 - dt2 * ldro((const T *)J.cnd + i)) * fcnd[i] + dfcnd) * ldro((const T *)J.cndinv + i);
    dfcnd = fcnd[i] - fcnd_prev;
  }
  const T *kap = (const T *)J.pml.kap, *sig = (const T *)J.pml.sig, *siginv = (const T *)J.pml.siginv;
  the_f[i] = ((kap ? ldro(kap + k) - ldro(sig + k) : T(1)) * the_f[i] + dfcnd) *
             (siginv ? ldro(siginv + k) : T(1));
  if (fu) {
    const T *kapu = (const T *)J.pmlu.kap, *sigu = (const T *)J.pmlu.sig,
            *siginvu = (const T *)J.pmlu.siginv;
    fu[i] = ldro(siginvu + ku) *
            ((kapu ? ldro(kapu + ku) - ldro(sigu + ku) : T(1)) * fu[i] + the_f[i] - fprev);
  }
}

// cylindrical helper array (src/step_db.cpp:104-116): one z column, serial in r
template <typename T> MB200_HD void cylint_column(const mb200_cylint_job_t &J, int64_t iz) {
  T *out = (T *)J.out;
  const T *fp = (const T *)J.fp;
  const T ir0 = (T)J.ir0;
  const int64_t sr = J.sr;
  T acc = 0;
  out[iz] = 0;
  T prev = ldro(fp + iz);
  for (int64_t ir = 1; ir <= J.nr; ++ir) {
    const T rinv = (T)(1.0 / ((double)((T)ir + ir0) - 0.5));
    const int64_t idx = ir * sr + iz;
    const T cur = ldro(fp + idx);
    acc = acc + rinv * (cur * ((T)ir + ir0) - prev * ((T)(ir - 1) + ir0));
    out[idx] = acc;
    prev = cur;
  }
}

// ------------------------------------------------------------------------------------------------
// step_update_EDHB (src/step_generic.cpp:566-785).  Callers have applied the swap of line 573.
// calc_nonlinear_u: lines 542-547.
template <typename T> MB200_HD T calc_nonlinear_u(T Dsqr, T Di, T chi1inv, T chi2, T chi3) {
  T c2 = Di * chi2 * (chi1inv * chi1inv);
  T c3 = Dsqr * chi3 * (chi1inv * chi1inv * chi1inv);
  return (1 + c2 + 2 * c3) / (1 + 2 * c2 + 3 * c3);
}

// OFFDIAG (lines 580-581)
template <typename T>
MB200_HD T offdiag(const T *u, const T *g, int64_t i, int64_t s, int64_t sx) {
  return T(0.25) * ((ldro(g + i) + ldro(g + i - sx)) * ldro(u + i) +
                    (ldro(g + i + s) + ldro(g + (i + s) - sx)) * ldro(u + i + s));
}

// store val = (u g) into f, through the fw ODE in PML (lines 596-602)
template <typename T>
MB200_HD void edhb_store(T *f, T *fw, const mb200_pml_t &pmlw, int64_t i, int kw, T val) {
  if (pmlw.sig) {
    const T *sigw = (const T *)pmlw.sig, *kapw = (const T *)pmlw.kap;
    const T fwprev = fw[i], kapwkw = ldro(kapw + kw), sigwkw = ldro(sigw + kw);
    fw[i] = val;
    f[i] += (kapwkw + sigwkw) * val - (kapwkw - sigwkw) * fwprev;
  }
  else
    f[i] = val;
}

template <typename T> MB200_HD void edhb_point(const mb200_edhb_job_t &J, int64_t i, int kw) {
  T *f = (T *)J.f;
  const T *g = (const T *)J.g, *g1 = (const T *)J.g1, *g2 = (const T *)J.g2;
  const T *u = (const T *)J.u, *u1 = (const T *)J.u1, *u2 = (const T *)J.u2;
  const T *chi2 = (const T *)J.chi2, *chi3 = (const T *)J.chi3;
  const int64_t s = J.s, s1 = J.s1, s2 = J.s2;
  const T gs = ldro(g + i);
  T val;
  if (u1 && u2) { // 3x3 (lines 588-615, 703-722)
    const T us = ldro(u + i);
    val = gs * us + offdiag(u1, g1, i, s, s1) + offdiag(u2, g2, i, s, s2);
    if (chi3) {
      T g1s = ldro(g1 + i) + ldro(g1 + i + s) + ldro(g1 + i - s1) + ldro(g1 + i + (s - s1));
      T g2s = ldro(g2 + i) + ldro(g2 + i + s) + ldro(g2 + i - s2) + ldro(g2 + i + (s - s2));
      val = val * calc_nonlinear_u(gs * gs + T(0.0625) * (g1s * g1s + g2s * g2s), gs, us, ldro(chi2 + i),
                                   ldro(chi3 + i));
    }
  }
  else if (u1) { // 2x2 (lines 616-639, 723-740)
    const T us = ldro(u + i);
    val = gs * us + offdiag(u1, g1, i, s, s1);
    if (chi3) {
      T g1s = ldro(g1 + i) + ldro(g1 + i + s) + ldro(g1 + i - s1) + ldro(g1 + i + (s - s1));
      val = val * calc_nonlinear_u(gs * gs + T(0.0625) * (g1s * g1s), gs, us, ldro(chi2 + i), ldro(chi3 + i));
    }
  }
  else if (chi3) { // diagonal, nonlinear (lines 644-681, 745-773)
    const T us = ldro(u + i);
    T dsqr;
    if (g1 && g2) {
      T g1s = ldro(g1 + i) + ldro(g1 + i + s) + ldro(g1 + i - s1) + ldro(g1 + i + (s - s1));
      T g2s = ldro(g2 + i) + ldro(g2 + i + s) + ldro(g2 + i - s2) + ldro(g2 + i + (s - s2));
      dsqr = gs * gs + T(0.0625) * (g1s * g1s + g2s * g2s);
    }
    else if (g1) {
      T g1s = ldro(g1 + i) + ldro(g1 + i + s) + ldro(g1 + i - s1) + ldro(g1 + i + (s - s1));
      dsqr = gs * gs + T(0.0625) * (g1s * g1s);
    }
    else
      dsqr = gs * gs;
    val = (gs * us) * calc_nonlinear_u(dsqr, gs, us, ldro(chi2 + i), ldro(chi3 + i));
  }
  else if (u) // lines 682-691, 774-779
    val = gs * ldro(u + i);
  else // lines 692-699, 781-782
    val = gs;

  edhb_store<T>(f, (T *)J.fw, J.pmlw, i, kw, val);
}

// diagonal, linear update_eh fused behind a curl update (mb200_step3_comp_t): e = u * d
template <typename T> MB200_HD void edhb_diag(const mb200_step3_comp_t &C, int64_t i, int kw, T d) {
  const T val = C.u ? d * ldro((const T *)C.u + i) : d;
  edhb_store<T>((T *)C.e, (T *)C.fw, C.pmlw, i, kw, val);
}

// ------------------------------------------------------------------------------------------------
// lorentzian_susceptibility::update_P (src/susceptibility.cpp:188-262)
template <typename T> MB200_HD void lorentz_point(const mb200_lorentz_job_t &J, int64_t i) {
  T *p = (T *)J.p, *pp = (T *)J.pp;
  const T *w = (const T *)J.w, *s = (const T *)J.s;
  const T *w1 = (const T *)J.w1, *s1 = (const T *)J.s1, *w2 = (const T *)J.w2,
          *s2 = (const T *)J.s2;
  const T gamma1inv = (T)J.gamma1inv, gamma1 = (T)J.gamma1, omega0dtsqr = (T)J.omega0dtsqr,
          omega0dtsqr_denom = (T)J.omega0dtsqr_denom;
  if (s1 && s2) { // 3x3 (lines 227-240)
    if (ldro(s + i) != 0) {
      T pcur = p[i];
      p[i] = gamma1inv * (pcur * (2 - omega0dtsqr_denom) - gamma1 * pp[i] +
                          omega0dtsqr * (ldro(s + i) * ldro(w + i) + offdiag(s1, w1, i, J.is, J.is1) +
                                         offdiag(s2, w2, i, J.is, J.is2)));
      pp[i] = pcur;
    }
  }
  else if (s1) { // 2x2 (lines 241-250)
    if (ldro(s + i) != 0) {
      T pcur = p[i];
      p[i] = gamma1inv * (pcur * (2 - omega0dtsqr_denom) - gamma1 * pp[i] +
                          omega0dtsqr * (ldro(s + i) * ldro(w + i) + offdiag(s1, w1, i, J.is, J.is1)));
      pp[i] = pcur;
    }
  }
  else { // isotropic (lines 251-257)
    T pcur = p[i];
    p[i] = gamma1inv *
           (pcur * (2 - omega0dtsqr_denom) - gamma1 * pp[i] + omega0dtsqr * (ldro(s + i) * ldro(w + i)));
    pp[i] = pcur;
  }
}

// ------------------------------------------------------------------------------------------------
// f_minus_p = D - sum P (src/update_eh.cpp:114-123, src/susceptibility.cpp:264-281)
template <typename T> MB200_HD void fmp_point(const mb200_fmp_job_t &J, int64_t i) {
  T *fmp = (T *)J.fmp;
  T v = J.d ? ldro((const T *)J.d + i) : fmp[i];
  T pv[MB200_MAX_P];
  const int64_t zb = i / MB200_ZBLOCK;
#pragma unroll
  for (int k = 0; k < MB200_MAX_P; ++k) // all loads first (they are independent)
    pv[k] = (k < J.np && !(J.pzero[k] && J.pzero[k][zb])) ? ldro((const T *)J.p[k] + i) : T(0);
#pragma unroll
  for (int k = 0; k < MB200_MAX_P; ++k)
    if (k < J.np) v -= pv[k];
  fmp[i] = v;
}

// ------------------------------------------------------------------------------------------------
// step_source (src/step.cpp:295-318) / dipole subtraction (src/update_eh.cpp:128-138)
template <typename T>
MB200_HD void source_point(const mb200_src_job_t &J, int64_t j, const double *scalars) {
  const int64_t i = J.index[j];
  const double ar = J.amp[2 * j], ai = J.amp[2 * j + 1];
  const double cr = scalars[2 * J.scalar_slot], ci = scalars[2 * J.scalar_slot + 1];
  double Ar = ar * cr - ai * ci, Ai = ar * ci + ai * cr; // amp[j] * current() (or dipole())
  if (J.mode == 0) {
    Ar *= J.dt;
    Ai *= J.dt;
    if (J.cndinv) {
      const double ci_ = (double)((const T *)J.cndinv)[i];
      Ar *= ci_;
      Ai *= ci_;
    }
  }
  T *fr = (T *)J.f_re, *fi = (T *)J.f_im;
  fr[i] = (T)((double)fr[i] - Ar);
  if (fi) fi[i] = (T)((double)fi[i] - Ai);
}

// ------------------------------------------------------------------------------------------------
// chunk-boundary transfer n of a halo job (src/step.cpp:178-221)
template <typename T> MB200_HD void halo_transfer(const mb200_halo_job_t &J, int64_t n) {
  if (n < J.n_phase) { // CONNECT_PHASE: complex multiply, two realnums per transfer
    const T *sr = (const T *)(uintptr_t)J.src[2 * n], *si = (const T *)(uintptr_t)J.src[2 * n + 1];
    T *dr = (T *)(uintptr_t)J.dst[2 * n], *di = (T *)(uintptr_t)J.dst[2 * n + 1];
    const T pr = ((const T *)J.phase)[2 * n], pi = ((const T *)J.phase)[2 * n + 1];
    const T vr = *sr, vi = *si;
    const T outr = pr * vr - pi * vi, outi = pr * vi + pi * vr;
    *dr = outr;
    *di = outi;
    if (J.dst_flag) {
      if (outr != T(0)) *(uint8_t *)(uintptr_t)J.dst_flag[2 * n] = 0;
      if (outi != T(0)) *(uint8_t *)(uintptr_t)J.dst_flag[2 * n + 1] = 0;
    }
  }
  else {
    const int64_t m = n + J.n_phase; // list position (phase entries occupy 2*n_phase slots)
    const T v = *(const T *)(uintptr_t)J.src[m];
    *(T *)(uintptr_t)J.dst[m] = (n < J.n_phase + J.n_negate) ? -v : v;
    if (J.dst_flag && v != T(0)) *(uint8_t *)(uintptr_t)J.dst_flag[m] = 0;
  }
}
MB200_HD int64_t halo_count(const mb200_halo_job_t &J) {
  return J.n_phase + J.n_negate + J.n_copy;
}
// element e of run r (run-length form of the NEGATE || COPY entries)
template <typename T> MB200_HD void halo_run_transfer(const mb200_halo_run_t &r, int e) {
  const T v = *(const T *)(uintptr_t)(r.src0 + (int64_t)e * r.dsrc);
  *(T *)(uintptr_t)(r.dst0 + (int64_t)e * r.ddst) = r.negate ? -v : v;
}

// ------------------------------------------------------------------------------------------------
// dft_chunk::update_dft (src/dft.cpp:266-308)
// IVEC_LOOP_WEIGHT1x (src/meep/vec.hpp:372-378)
MB200_HD double loop_weight1(double s0, double s1, double e0, double e1, int i, int n) {
  return (i > 1 && i < n - 2)
             ? 1.0
             : (i == 0 ? s0 : (i == 1 ? s1 : i == n - 1 ? e0 : (i == n - 2 ? e1 : 1.0)));
}

// weighted, Yee->centre averaged field value(s) at loop point (i1,i2,i3): lines 277-294
template <typename T>
MB200_HD void dft_field_value(const mb200_dft_job_t &J, int i1, int i2, int i3, T &fr, T &fi) {
  const int64_t idx = box_index(J.box, i1, i2, i3);
  double w;
  if (J.use_weights) {
    // IVEC_LOOP_WEIGHT (src/meep/vec.hpp:381-383) with dV = dV0 + dV1*loop_i2
    w = loop_weight1(J.wgt_s0[2], J.wgt_s1[2], J.wgt_e0[2], J.wgt_e1[2], i3, J.box.n[2]) *
        (loop_weight1(J.wgt_s0[1], J.wgt_s1[1], J.wgt_e0[1], J.wgt_e1[1], i2, J.box.n[1]) *
         ((J.dV0 + J.dV1 * i2) *
          loop_weight1(J.wgt_s0[0], J.wgt_s1[0], J.wgt_e0[0], J.wgt_e1[0], i1, J.box.n[0])));
    if (J.sqrt_weights) w = sqrt(w);
  }
  else
    w = 1.0;
  const T *re = (const T *)J.f_re, *im = (const T *)J.f_im;
  const int64_t a1 = J.avg1, a2 = J.avg2;
  if (a2) {
    fr = (T)((w * 0.25) * (re[idx] + re[idx + a1] + re[idx + a2] + re[idx + (a1 + a2)]));
    fi = im ? (T)((w * 0.25) * (im[idx] + im[idx + a1] + im[idx + a2] + im[idx + (a1 + a2)]))
            : T(0);
  }
  else if (a1) {
    fr = (T)((w * 0.5) * (re[idx] + re[idx + a1]));
    fi = im ? (T)((w * 0.5) * (im[idx] + im[idx + a1])) : T(0);
  }
  else {
    fr = (T)(w * re[idx]);
    fi = im ? (T)(w * im[idx]) : T(0);
  }
}

// accumulate one frequency: lines 296-306
template <typename T>
MB200_HD void dft_accumulate(T *dft2, bool is_complex, T pr, T pi, T fr, T fi) {
  if (is_complex) {
    dft2[0] += pr * fr - pi * fi;
    dft2[1] += pr * fi + pi * fr;
  }
  else {
    dft2[0] += fr * pr;
    dft2[1] += fr * pi;
  }
}

} // namespace mb200
#endif
