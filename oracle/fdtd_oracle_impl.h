/* fdtd_oracle_impl.h — body of the oracle, included twice by fdtd_oracle.c with
 * REAL = double / float and SUF = _f64 / _f32.   TEST INFRASTRUCTURE ONLY (see fdtd_oracle.c). */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* reference src/meep_internals.hpp:217-226 (KSTRIDE_DEF/KDEF) */
static int FN(kidx)(const mb200_pml_t *p, int i1, int i2, int i3) {
  return ((p->k0 + p->ks[0] * i1) + p->ks[1] * i2) + p->ks[2] * i3;
}

/* reference src/step_generic.cpp:65-249 (step_curl), most general case 201-211 with absent
 * terms dropped exactly as the specialised loops drop them. */
void FN(oracle_step_curl)(const mb200_curl_job_t *J) {
  REAL *f = (REAL *)J->f, *fu = (REAL *)J->fu, *fcnd = (REAL *)J->fcnd;
  const REAL *g1 = (const REAL *)J->g1, *g2 = (const REAL *)J->g2;
  const REAL *cnd = (const REAL *)J->cnd, *cndinv = (const REAL *)J->cndinv;
  const REAL *sig = (const REAL *)J->pml.sig, *kap = (const REAL *)J->pml.kap,
             *siginv = (const REAL *)J->pml.siginv;
  const REAL *sigu = (const REAL *)J->pmlu.sig, *kapu = (const REAL *)J->pmlu.kap,
             *siginvu = (const REAL *)J->pmlu.siginv;
  int64_t s1 = J->s1, s2 = J->s2;
  REAL dtdx = (REAL)J->dtdx;
  const REAL dt2 = (REAL)J->dt * (REAL)0.5;
  if (!g1) { /* lines 72-76 */
    const REAL *t = g1; g1 = g2; g2 = t;
    int64_t ts = s1; s1 = s2; s2 = ts;
    dtdx = -dtdx;
  }
  for (int i1 = 0; i1 < J->box.n[0]; ++i1)
    for (int i2 = 0; i2 < J->box.n[1]; ++i2)
      for (int i3 = 0; i3 < J->box.n[2]; ++i3) {
        const int64_t i = J->box.idx0 + i1 * J->box.s[0] + i2 * J->box.s[1] + i3 * J->box.s[2];
        REAL curl;
        if (g2) curl = dtdx * (g1[i + s1] - g1[i] + g2[i] - g2[i + s2]);
        else curl = dtdx * (g1[i + s1] - g1[i]);
        if (!sig) {
          if (!sigu) {
            if (cnd) f[i] = ((1 - dt2 * cnd[i]) * f[i] - curl) * cndinv[i];
            else f[i] -= curl;
          }
          else {
            const int ku = FN(kidx)(&J->pmlu, i1, i2, i3);
            const REAL fprev = fu[i];
            if (cnd) fu[i] = ((1 - dt2 * cnd[i]) * fprev - curl) * cndinv[i];
            else fu[i] -= curl;
            f[i] = siginvu[ku] * ((kapu[ku] - sigu[ku]) * f[i] + fu[i] - fprev);
          }
        }
        else {
          const int k = FN(kidx)(&J->pml, i1, i2, i3);
          if (!sigu) {
            if (cnd) {
              const REAL fcnd_prev = fcnd[i];
              fcnd[i] = ((1 - dt2 * cnd[i]) * fcnd[i] - curl) * cndinv[i];
              f[i] = ((kap[k] - sig[k]) * f[i] + (fcnd[i] - fcnd_prev)) * siginv[k];
            }
            else f[i] = ((kap[k] - sig[k]) * f[i] - curl) * siginv[k];
          }
          else {
            const int ku = FN(kidx)(&J->pmlu, i1, i2, i3);
            const REAL fprev = fu[i];
            if (cnd) {
              const REAL fcnd_prev = fcnd[i];
              fcnd[i] = ((1 - dt2 * cnd[i]) * fcnd[i] - curl) * cndinv[i];
              fu[i] = ((kap[k] - sig[k]) * fu[i] + (fcnd[i] - fcnd_prev)) * siginv[k];
            }
            else fu[i] = ((kap[k] - sig[k]) * fu[i] - curl) * siginv[k];
            f[i] = siginvu[ku] * ((kapu[ku] - sigu[ku]) * f[i] + fu[i] - fprev);
          }
        }
      }
}

/* reference src/step_generic.cpp:255-333 (step_beta) */
void FN(oracle_step_beta)(const mb200_beta_job_t *J) {
  REAL *f = (REAL *)J->f, *fu = (REAL *)J->fu, *fcnd = (REAL *)J->fcnd;
  const REAL *g = (const REAL *)J->g, *cndinv = (const REAL *)J->cndinv;
  const REAL *siginv = (const REAL *)J->pml.siginv, *siginvu = (const REAL *)J->pmlu.siginv;
  const REAL the_m = (REAL)J->betadt;
  if (!g) return;
  for (int i1 = 0; i1 < J->box.n[0]; ++i1)
    for (int i2 = 0; i2 < J->box.n[1]; ++i2)
      for (int i3 = 0; i3 < J->box.n[2]; ++i3) {
        const int64_t i = J->box.idx0 + i1 * J->box.s[0] + i2 * J->box.s[1] + i3 * J->box.s[2];
        /* the cylindrical i*m/r loops of reference src/step_db.cpp:203-280 are these same eight
         * loops with rinv = the_m / (loop_is2 + 2 * loop_i2) in place of betadt */
        const REAL betadt = J->cyl ? the_m / (J->r_is2 + 2 * i2) : the_m;
        if (siginv) {
          const int k = FN(kidx)(&J->pml, i1, i2, i3);
          if (siginvu) {
            const int ku = FN(kidx)(&J->pmlu, i1, i2, i3);
            REAL df;
            if (cndinv) {
              REAL dfcnd = betadt * g[i] * cndinv[i];
              fcnd[i] += dfcnd;
              fu[i] += (df = dfcnd * siginv[k]);
            }
            else fu[i] += (df = betadt * g[i] * siginv[k]);
            f[i] += siginvu[ku] * df;
          }
          else if (cndinv) {
            REAL dfcnd = betadt * g[i] * cndinv[i];
            fcnd[i] += dfcnd;
            f[i] += dfcnd * siginv[k];
          }
          else f[i] += betadt * g[i] * siginv[k];
        }
        else if (siginvu) {
          const int ku = FN(kidx)(&J->pmlu, i1, i2, i3);
          REAL df;
          if (cndinv) fu[i] += (df = betadt * g[i] * cndinv[i]);
          else fu[i] += (df = betadt * g[i]);
          f[i] += siginvu[ku] * df;
        }
        else if (cndinv) f[i] += betadt * g[i] * cndinv[i];
        else f[i] += betadt * g[i];
      }
}

/* reference src/step_generic.cpp:335-530 (step_bfast); g1 != NULL on entry (the caller applies the
 * swap of lines 342-346) */
void FN(oracle_step_bfast)(const mb200_bfast_job_t *J) {
  REAL *f = (REAL *)J->f, *fu = (REAL *)J->fu, *fcnd = (REAL *)J->fcnd, *F = (REAL *)J->F;
  const REAL *g1 = (const REAL *)J->g1, *g2 = (const REAL *)J->g2;
  const REAL *cnd = (const REAL *)J->cnd, *cndinv = (const REAL *)J->cndinv;
  const REAL *siginv = (const REAL *)J->pml.siginv, *siginvu = (const REAL *)J->pmlu.siginv;
  const REAL k1 = (REAL)J->k1, k2 = (REAL)J->k2;
  const int64_t s1 = J->s1, s2 = J->s2;
  for (int i1 = 0; i1 < J->box.n[0]; ++i1)
    for (int i2 = 0; i2 < J->box.n[1]; ++i2)
      for (int i3 = 0; i3 < J->box.n[2]; ++i3) {
        const int64_t i = J->box.idx0 + i1 * J->box.s[0] + i2 * J->box.s[1] + i3 * J->box.s[2];
        const int k = FN(kidx)(&J->pml, i1, i2, i3), ku = FN(kidx)(&J->pmlu, i1, i2, i3);
        REAL F_prev = F[i];
        if (g2) F[i] = (k1 * (g1[i + s1] + g1[i]) - k2 * (g2[i + s2] + g2[i])) - F[i];
        else if (!siginv && !siginvu && !cnd) F[i] = k1 * (g1[i + s1] + g1[i]); /* line 372 */
        else F[i] = k1 * (g1[i + s1] + g1[i]) - F[i];
        if (!siginv) {
          if (!siginvu) {
            if (cnd) f[i] += (F[i] - F_prev) * cndinv[i];
            else f[i] += (F[i] - F_prev);
          }
          else {
            REAL df;
            if (cnd) fu[i] += (df = (F[i] - F_prev) * cndinv[i]);
            else fu[i] += (df = (F[i] - F_prev));
            f[i] += siginvu[ku] * df;
          }
        }
        else if (!siginvu) {
          if (cnd) {
            REAL dfcnd = (F[i] - F_prev) * cndinv[i];
            fcnd[i] += dfcnd;
            f[i] += dfcnd * siginv[k];
          }
          else f[i] += (F[i] - F_prev) * siginv[k];
        }
        else {
          REAL df;
          if (cnd) {
            REAL dfcnd = (F[i] - F_prev) * cndinv[i];
            fcnd[i] += dfcnd;
            fu[i] += (df = dfcnd * siginv[k]);
          }
          else fu[i] += (df = (F[i] - F_prev) * siginv[k]);
          f[i] += siginvu[ku] * df;
        }
      }
}

/* reference src/boundaries.cpp:310-313 (fields_chunk::zero_metal); also the ZERO_Z rows of
 * src/step_db.cpp:283,322-327,372-376,406-461 */
void FN(oracle_zero_metal)(const mb200_zero_job_t *J) {
  for (int64_t i = 0; i < J->n; ++i)
    *(REAL *)(uintptr_t)J->ptrs[i] = 0;
}

/* reference src/susceptibility.cpp:445-584 (gyrotropic_susceptibility::update_P), the loop body
 * for one driving-field component in its rotated frame (see mb200_gyro_job_t) */
#define OFFDIAGW(g, sx, s) (0.25 * (g[i] + g[i - sx] + g[i + s] + g[i + s - sx]))
void FN(oracle_gyrotropic_update_P)(const mb200_gyro_job_t *J) {
  REAL *p0 = (REAL *)J->p[0], *p1 = (REAL *)J->p[1], *p2 = (REAL *)J->p[2];
  REAL *pp0 = (REAL *)J->pp[0], *pp1 = (REAL *)J->pp[1], *pp2 = (REAL *)J->pp[2];
  const REAL *w0 = (const REAL *)J->w[0], *w1 = (const REAL *)J->w[1], *w2 = (const REAL *)J->w[2];
  const REAL *s = (const REAL *)J->s;
  const int64_t is = J->is, is1 = J->is1, is2 = J->is2;
  REAL gt[3][3], inv[3][3];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      gt[a][b] = (REAL)J->gt[a][b];
      inv[a][b] = (REAL)J->inv[a][b];
    }
  for (int i1 = 0; i1 < J->box.n[0]; ++i1)
    for (int i2 = 0; i2 < J->box.n[1]; ++i2)
      for (int i3 = 0; i3 < J->box.n[2]; ++i3) {
        const int64_t i = J->box.idx0 + i1 * J->box.s[0] + i2 * J->box.s[1] + i3 * J->box.s[2];
        REAL r0, r1, r2;
        if (J->model == 0) {
          const REAL diag = (REAL)J->c[0], gamma1 = (REAL)J->c[1], omega0dtsqr = (REAL)J->c[2],
                     pt = (REAL)J->c[3];
          r0 = diag * p0[i] - gamma1 * pp0[i] + omega0dtsqr * s[i] * w0[i] - pt * gt[0][1] * pp1[i] -
               pt * gt[0][2] * pp2[i];
          r1 = diag * p1[i] - gamma1 * pp1[i] + (w1 ? omega0dtsqr * s[i] * OFFDIAGW(w1, is1, is) : 0) -
               pt * gt[1][0] * pp0[i] - pt * gt[1][2] * pp2[i];
          r2 = diag * p2[i] - gamma1 * pp2[i] + (w2 ? omega0dtsqr * s[i] * OFFDIAGW(w2, is2, is) : 0) -
               pt * gt[2][1] * pp1[i] - pt * gt[2][0] * pp0[i];
        }
        else {
          const REAL omega2pidt = (REAL)J->c[0], g2pidt = (REAL)J->c[1], alpha = (REAL)J->c[2],
                     dt2pi = (REAL)J->c[3];
          REAL q0 = -omega2pidt * p0[i] + 0.5 * alpha * pp0[i] + dt2pi * s[i] * w0[i];
          REAL q1 = -omega2pidt * p1[i] + 0.5 * alpha * pp1[i] +
                    dt2pi * s[i] * (w1 ? OFFDIAGW(w1, is1, is) : 0);
          REAL q2 = -omega2pidt * p2[i] + 0.5 * alpha * pp2[i] +
                    dt2pi * s[i] * (w2 ? OFFDIAGW(w2, is2, is) : 0);
          r0 = 0.5 * pp0[i] - g2pidt * p0[i] + gt[0][1] * q1 + gt[0][2] * q2;
          r1 = 0.5 * pp1[i] - g2pidt * p1[i] + gt[1][2] * q2 + gt[1][0] * q0;
          r2 = 0.5 * pp2[i] - g2pidt * p2[i] + gt[2][0] * q0 + gt[2][1] * q1;
        }
        pp0[i] = p0[i];
        pp1[i] = p1[i];
        pp2[i] = p2[i];
        p0[i] = inv[0][0] * r0 + inv[0][1] * r1 + inv[0][2] * r2;
        p1[i] = inv[1][0] * r0 + inv[1][1] * r1 + inv[1][2] * r2;
        p2[i] = inv[2][0] * r0 + inv[2][1] * r1 + inv[2][2] * r2;
      }
}
#undef OFFDIAGW

/* reference src/susceptibility.cpp:331-334: p[i] += gaussian_random(0, amp sqrt(s[i])), with the
 * random numbers handed in (drawn by the caller in loop order) */
void FN(oracle_add_noise)(const mb200_noise_job_t *J, const double *noise) {
  REAL *p = (REAL *)J->p;
  int64_t k = J->slot;
  for (int i1 = 0; i1 < J->box.n[0]; ++i1)
    for (int i2 = 0; i2 < J->box.n[1]; ++i2)
      for (int i3 = 0; i3 < J->box.n[2]; ++i3) {
        const int64_t i = J->box.idx0 + i1 * J->box.s[0] + i2 * J->box.s[1] + i3 * J->box.s[2];
        p[i] += noise[k++];
      }
}

/* reference src/energy_and_flux.cpp:139-147 (fields_chunk::average_with_backup) */
void FN(oracle_average_with_backup)(const mb200_average_job_t *J) {
  REAL *fc = (REAL *)J->f;
  const REAL *backup = (const REAL *)J->backup;
  for (int64_t i = 0; i < J->n; i++)
    fc[i] = 0.5 * (fc[i] + backup[i]);
}

/* reference src/step_db.cpp:104-116: running sum over r of 1/r d(r f_p)/dr */
void FN(oracle_cyl_rderiv_int)(const mb200_cylint_job_t *J) {
  REAL *out = (REAL *)J->out;
  const REAL *f_p = (const REAL *)J->fp;
  const REAL ir0 = (REAL)J->ir0;
  const int64_t sr = J->sr;
  for (int64_t iz = 0; iz < sr; ++iz)
    out[iz] = 0;
  for (int64_t ir = 1; ir <= J->nr; ++ir) {
    REAL rinv = 1.0 / ((ir + ir0) - 0.5);
    for (int64_t iz = 0; iz < sr; ++iz) {
      int64_t idx = ir * sr + iz;
      out[idx] = out[idx - sr] + rinv * (f_p[idx] * (ir + ir0) - f_p[idx - sr] * ((ir - 1) + ir0));
    }
  }
}

/* reference src/step_db.cpp:285-321 (m == 0, Dz at r = 0) and 329-371 (|m| == 1, Dp / Br) */
void FN(oracle_cyl_origin)(const mb200_cylr0_job_t *J) {
  REAL *the_f = (REAL *)J->f, *fu = (REAL *)J->fu, *fcnd = (REAL *)J->fcnd;
  const REAL *f_p = (const REAL *)J->fp, *f_m = (const REAL *)J->fm;
  const REAL *cnd = (const REAL *)J->cnd, *cndinv = (const REAL *)J->cndinv;
  const REAL *sig = (const REAL *)J->pml.sig, *kap = (const REAL *)J->pml.kap,
             *siginv = (const REAL *)J->pml.siginv;
  const REAL *sigu = (const REAL *)J->pmlu.sig, *kapu = (const REAL *)J->pmlu.kap,
             *siginvu = (const REAL *)J->pmlu.siginv;
  const REAL dt2 = J->dt * 0.5, f_m_mult = (REAL)J->mult;
  const int64_t sd = J->sd;
  for (int i1 = 0; i1 < J->box.n[0]; ++i1)
    for (int i2 = 0; i2 < J->box.n[1]; ++i2)
      for (int i3 = 0; i3 < J->box.n[2]; ++i3) {
        const int64_t i = J->box.idx0 + i1 * J->box.s[0] + i2 * J->box.s[1] + i3 * J->box.s[2];
        REAL fprev = the_f[i];
        REAL dfcnd = J->mode == 0 ? f_p[i] * J->c : J->c * (f_p[i] - f_p[i - sd] - f_m_mult * f_m[i]);
        if (fcnd) {
          REAL fcnd_prev = fcnd[i];
          fcnd[i] = ((1 - dt2 * cnd[i]) * fcnd[i] + dfcnd) * cndinv[i];
          dfcnd = fcnd[i] - fcnd_prev;
        }
        const int k = FN(kidx)(&J->pml, i1, i2, i3), ku = FN(kidx)(&J->pmlu, i1, i2, i3);
        the_f[i] = ((kap ? kap[k] - sig[k] : 1) * the_f[i] + dfcnd) * (siginv ? siginv[k] : 1);
        if (fu) fu[i] = siginvu[ku] * ((kapu ? kapu[ku] - sigu[ku] : 1) * fu[i] + the_f[i] - fprev);
      }
}

/* reference src/step_generic.cpp:542-547 */
static REAL FN(nonlinear_u)(REAL Dsqr, REAL Di, REAL chi1inv, REAL chi2, REAL chi3) {
  REAL c2 = Di * chi2 * (chi1inv * chi1inv);
  REAL c3 = Dsqr * chi3 * (chi1inv * chi1inv * chi1inv);
  return (1 + c2 + 2 * c3) / (1 + 2 * c2 + 3 * c3);
}

/* reference src/step_generic.cpp:566-785 (step_update_EDHB) */
void FN(oracle_step_update_EDHB)(const mb200_edhb_job_t *J) {
  REAL *f = (REAL *)J->f, *fw = (REAL *)J->fw;
  const REAL *g = (const REAL *)J->g, *g1 = (const REAL *)J->g1, *g2 = (const REAL *)J->g2;
  const REAL *u = (const REAL *)J->u, *u1 = (const REAL *)J->u1, *u2 = (const REAL *)J->u2;
  const REAL *chi2 = (const REAL *)J->chi2, *chi3 = (const REAL *)J->chi3;
  const REAL *sigw = (const REAL *)J->pmlw.sig, *kapw = (const REAL *)J->pmlw.kap;
  int64_t s = J->s, s1 = J->s1, s2 = J->s2;
  if (!f) return;
  if ((!g1 && g2) || (g1 && g2 && !u1 && u2)) { /* lines 573-577 */
    const REAL *t = g1; g1 = g2; g2 = t;
    t = u1; u1 = u2; u2 = t;
    int64_t ts = s1; s1 = s2; s2 = ts;
  }
#define OFFD(u_, g_, sx) \
  ((REAL)0.25 * ((g_[i] + g_[i - sx]) * u_[i] + (g_[i + s] + g_[(i + s) - sx]) * u_[i + s]))
  for (int i1 = 0; i1 < J->box.n[0]; ++i1)
    for (int i2 = 0; i2 < J->box.n[1]; ++i2)
      for (int i3 = 0; i3 < J->box.n[2]; ++i3) {
        const int64_t i = J->box.idx0 + i1 * J->box.s[0] + i2 * J->box.s[1] + i3 * J->box.s[2];
        const REAL gs = g[i];
        REAL val;
        if (u1 && u2) {
          const REAL us = u[i];
          val = gs * us + OFFD(u1, g1, s1) + OFFD(u2, g2, s2);
          if (chi3) {
            REAL g1s = g1[i] + g1[i + s] + g1[i - s1] + g1[i + (s - s1)];
            REAL g2s = g2[i] + g2[i + s] + g2[i - s2] + g2[i + (s - s2)];
            val = val * FN(nonlinear_u)(gs * gs + (REAL)0.0625 * (g1s * g1s + g2s * g2s), gs, us,
                                        chi2[i], chi3[i]);
          }
        }
        else if (u1) {
          const REAL us = u[i];
          val = gs * us + OFFD(u1, g1, s1);
          if (chi3) {
            REAL g1s = g1[i] + g1[i + s] + g1[i - s1] + g1[i + (s - s1)];
            val = val * FN(nonlinear_u)(gs * gs + (REAL)0.0625 * (g1s * g1s), gs, us, chi2[i],
                                        chi3[i]);
          }
        }
        else if (chi3) {
          const REAL us = u[i];
          REAL dsq = gs * gs;
          if (g1 && g2) {
            REAL g1s = g1[i] + g1[i + s] + g1[i - s1] + g1[i + (s - s1)];
            REAL g2s = g2[i] + g2[i + s] + g2[i - s2] + g2[i + (s - s2)];
            dsq = gs * gs + (REAL)0.0625 * (g1s * g1s + g2s * g2s);
          }
          else if (g1) {
            REAL g1s = g1[i] + g1[i + s] + g1[i - s1] + g1[i + (s - s1)];
            dsq = gs * gs + (REAL)0.0625 * (g1s * g1s);
          }
          val = (gs * us) * FN(nonlinear_u)(dsq, gs, us, chi2[i], chi3[i]);
        }
        else if (u) val = gs * u[i];
        else val = gs;
        if (sigw) { /* lines 596-602 */
          const int kw = FN(kidx)(&J->pmlw, i1, i2, i3);
          const REAL fwprev = fw[i], kapwkw = kapw[kw], sigwkw = sigw[kw];
          fw[i] = val;
          f[i] += (kapwkw + sigwkw) * fw[i] - (kapwkw - sigwkw) * fwprev;
        }
        else f[i] = val;
      }
#undef OFFD
}

/* reference src/susceptibility.cpp:188-262 (lorentzian_susceptibility::update_P), one (c,cmp) */
void FN(oracle_lorentzian_update_P)(const mb200_lorentz_job_t *J) {
  REAL *p = (REAL *)J->p, *pp = (REAL *)J->pp;
  const REAL *w = (const REAL *)J->w, *s = (const REAL *)J->s;
  const REAL *w1 = (const REAL *)J->w1, *s1 = (const REAL *)J->s1;
  const REAL *w2 = (const REAL *)J->w2, *s2 = (const REAL *)J->s2;
  const REAL gamma1inv = (REAL)J->gamma1inv, gamma1 = (REAL)J->gamma1,
             omega0dtsqr = (REAL)J->omega0dtsqr, omega0dtsqr_denom = (REAL)J->omega0dtsqr_denom;
  const int64_t is = J->is, is1 = J->is1, is2 = J->is2;
#define OFFD(u_, g_, sx) \
  ((REAL)0.25 * ((g_[i] + g_[i - sx]) * u_[i] + (g_[i + is] + g_[(i + is) - sx]) * u_[i + is]))
  for (int i1 = 0; i1 < J->box.n[0]; ++i1)
    for (int i2 = 0; i2 < J->box.n[1]; ++i2)
      for (int i3 = 0; i3 < J->box.n[2]; ++i3) {
        const int64_t i = J->box.idx0 + i1 * J->box.s[0] + i2 * J->box.s[1] + i3 * J->box.s[2];
        if (s1 && s2) {
          if (s[i] != 0) {
            REAL pcur = p[i];
            p[i] = gamma1inv * (pcur * (2 - omega0dtsqr_denom) - gamma1 * pp[i] +
                                omega0dtsqr * (s[i] * w[i] + OFFD(s1, w1, is1) + OFFD(s2, w2, is2)));
            pp[i] = pcur;
          }
        }
        else if (s1) {
          if (s[i] != 0) {
            REAL pcur = p[i];
            p[i] = gamma1inv * (pcur * (2 - omega0dtsqr_denom) - gamma1 * pp[i] +
                                omega0dtsqr * (s[i] * w[i] + OFFD(s1, w1, is1)));
            pp[i] = pcur;
          }
        }
        else {
          REAL pcur = p[i];
          p[i] = gamma1inv *
                 (pcur * (2 - omega0dtsqr_denom) - gamma1 * pp[i] + omega0dtsqr * (s[i] * w[i]));
          pp[i] = pcur;
        }
      }
#undef OFFD
}

/* reference src/update_eh.cpp:114-123 + src/susceptibility.cpp:264-281 */
void FN(oracle_subtract_P)(const mb200_fmp_job_t *J) {
  REAL *fmp = (REAL *)J->fmp;
  if (J->d) memcpy(fmp, J->d, sizeof(REAL) * (size_t)J->ntot);
  for (int k = 0; k < J->np; ++k) {
    const REAL *p = (const REAL *)J->p[k];
    for (int64_t i = 0; i < J->ntot; ++i)
      fmp[i] -= p[i];
  }
}

/* reference src/step.cpp:295-318 (mode 0) and src/update_eh.cpp:128-138 (mode 1) */
void FN(oracle_step_source)(const mb200_src_job_t *J, const double *scalars) {
  REAL *fr = (REAL *)J->f_re, *fi = (REAL *)J->f_im;
  const REAL *cndinv = (const REAL *)J->cndinv;
  const double complex sc = scalars[2 * J->scalar_slot] + I * scalars[2 * J->scalar_slot + 1];
  for (int64_t j = 0; j < J->npts; ++j) {
    const int64_t i = J->index[j];
    const double complex amp = J->amp[2 * j] + I * J->amp[2 * j + 1];
    double complex A = amp * sc;
    if (J->mode == 0) {
      A = A * J->dt;
      if (cndinv) A = A * (double)cndinv[i];
    }
    fr[i] -= creal(A);
    if (fi) fi[i] -= cimag(A);
  }
}

/* reference src/step.cpp:172-223 (gather into a block, then scatter with phase/negate/copy) */
void FN(oracle_step_boundaries)(const mb200_halo_job_t *J) {
  const int64_t nlist = 2 * J->n_phase + (J->nrun > 0 ? 0 : J->n_negate + J->n_copy);
  REAL *block = (REAL *)malloc(sizeof(REAL) * (size_t)(nlist ? nlist : 1));
  for (int64_t n = 0; n < nlist; ++n)
    block[n] = *(const REAL *)(uintptr_t)J->src[n];
  const REAL *ph = (const REAL *)J->phase;
  int64_t o = 0;
  for (int64_t n = 0; n < J->n_phase; ++n) {
    const REAL pr = ph[2 * n], pi = ph[2 * n + 1], vr = block[2 * n], vi = block[2 * n + 1];
    *(REAL *)(uintptr_t)J->dst[2 * n] = pr * vr - pi * vi;
    *(REAL *)(uintptr_t)J->dst[2 * n + 1] = pr * vi + pi * vr;
  }
  o = 2 * J->n_phase;
  if (J->nrun > 0) { /* the same NEGATE || COPY transfers, listed as constant-stride runs */
    free(block);
    int64_t total = 0;
    for (int64_t r = 0; r < J->nrun; ++r)
      total += J->runs[r].n;
    block = (REAL *)malloc(sizeof(REAL) * (size_t)(total ? total : 1));
    int64_t q = 0;
    for (int64_t r = 0; r < J->nrun; ++r) /* gather everything first, as the comm block does */
      for (int e = 0; e < J->runs[r].n; ++e)
        block[q++] = *(const REAL *)(uintptr_t)(J->runs[r].src0 + (int64_t)e * J->runs[r].dsrc);
    q = 0;
    for (int64_t r = 0; r < J->nrun; ++r)
      for (int e = 0; e < J->runs[r].n; ++e, ++q)
        *(REAL *)(uintptr_t)(J->runs[r].dst0 + (int64_t)e * J->runs[r].ddst) =
            J->runs[r].negate ? -block[q] : block[q];
    free(block);
    return;
  }
  for (int64_t n = 0; n < J->n_negate; ++n)
    *(REAL *)(uintptr_t)J->dst[o + n] = -block[o + n];
  o += J->n_negate;
  for (int64_t n = 0; n < J->n_copy; ++n)
    *(REAL *)(uintptr_t)J->dst[o + n] = block[o + n];
  free(block);
}

/* reference src/meep/vec.hpp:372-378 */
static double FN(w1x)(double s0, double s1, double e0, double e1, int i, int n) {
  return (i > 1 && i < n - 2)
             ? 1.0
             : (i == 0 ? s0 : (i == 1 ? s1 : i == n - 1 ? e0 : (i == n - 2 ? e1 : 1.0)));
}

/* reference src/dft.cpp:266-308 (dft_chunk::update_dft); phases = dft_phase[] */
void FN(oracle_update_dft)(const mb200_dft_job_t *J, const REAL *phases) {
  const REAL *re = (const REAL *)J->f_re, *im = (const REAL *)J->f_im;
  REAL *dft = (REAL *)J->dft;
  const REAL *ph = phases + 2 * (int64_t)J->phase_slot;
  const int numcmp = im ? 2 : 1;
  int64_t idx_dft = 0;
  for (int i1 = 0; i1 < J->box.n[0]; ++i1)
    for (int i2 = 0; i2 < J->box.n[1]; ++i2)
      for (int i3 = 0; i3 < J->box.n[2]; ++i3, ++idx_dft) {
        const int64_t idx = J->box.idx0 + i1 * J->box.s[0] + i2 * J->box.s[1] + i3 * J->box.s[2];
        double w;
        if (J->use_weights) {
          w = FN(w1x)(J->wgt_s0[2], J->wgt_s1[2], J->wgt_e0[2], J->wgt_e1[2], i3, J->box.n[2]) *
              (FN(w1x)(J->wgt_s0[1], J->wgt_s1[1], J->wgt_e0[1], J->wgt_e1[1], i2, J->box.n[1]) *
               ((J->dV0 + J->dV1 * i2) *
                FN(w1x)(J->wgt_s0[0], J->wgt_s1[0], J->wgt_e0[0], J->wgt_e1[0], i1, J->box.n[0])));
          if (J->sqrt_weights) w = sqrt(w);
        }
        else w = 1.0;
        REAL f[2] = {0, 0};
        for (int cmp = 0; cmp < numcmp; ++cmp) {
          const REAL *a = cmp ? im : re;
          if (J->avg2)
            f[cmp] = (w * 0.25) * (a[idx] + a[idx + J->avg1] + a[idx + J->avg2] +
                                   a[idx + (J->avg1 + J->avg2)]);
          else if (J->avg1) f[cmp] = (w * 0.5) * (a[idx] + a[idx + J->avg1]);
          else f[cmp] = w * a[idx];
        }
        for (int k = 0; k < J->nomega; ++k) {
          REAL *d = dft + 2 * (J->nomega * idx_dft + k);
          if (numcmp == 2) {
            d[0] += ph[2 * k] * f[0] - ph[2 * k + 1] * f[1];
            d[1] += ph[2 * k] * f[1] + ph[2 * k + 1] * f[0];
          }
          else {
            d[0] += f[0] * ph[2 * k];
            d[1] += f[0] * ph[2 * k + 1];
          }
        }
      }
}

/* reference src/dft.cpp:542-556 inner sum */
void FN(oracle_dft_flux)(const mb200_flux_job_t *J) {
  const REAL *e = (const REAL *)J->e, *h = (const REAL *)J->h;
  for (int64_t k = 0; k < J->npts; ++k)
    for (int i = 0; i < J->nomega; ++i) {
      const int64_t o = 2 * (k * J->nomega + i);
      J->out[i] += (double)(e[o] * h[o] + e[o + 1] * h[o + 1]);
    }
}

#undef FN
#undef CAT
#undef CAT_
