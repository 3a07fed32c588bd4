// update_pols.cpp — B200 replacement for the reference translation unit src/update_pols.cpp,
// plus the two hot member functions of src/susceptibility.cpp
// (lorentzian_susceptibility::update_P / subtract_P).
//
// fields_chunk::update_pols keeps the reference's lazy allocation of the polarisation state
// (new_internal_data / init_internal_data run on the host, reference code) and then calls
// lorentzian_susceptibility::update_P — OUR definition, which emits one mb200_lorentz_job_t
// per (component, cmp) instead of looping (reference src/susceptibility.cpp:188-262).
#include <assert.h>
#include <string.h>
#include <typeinfo>

#include "engine.hpp"
#include "hostmem.hpp"
#include "loop_desc.hpp"
#include "meep_internals.hpp"

using namespace std;
using namespace meep_b200;

namespace meep {

void fields::update_pols(field_type ft) {
  Engine &E = Engine::get(this);
  Scope scope(E, this);
  run_phase(E, this, PH_POLS, ft, true, [&]() {
    for (int i = 0; i < num_chunks; i++)
      if (chunks[i]->is_mine())
        if (chunks[i]->update_pols(ft)) {
          chunk_connections_valid = false;
          assert(changed_materials);
        }
  });
}

bool fields_chunk::update_pols(field_type ft) {
  Engine *E = Engine::current();
  if (!E || !E->recording()) meep::abort("meep_b200: fields_chunk::update_pols outside a phase");
  bool allocated_fields = false;

  realnum *w[NUM_FIELD_COMPONENTS][2];
  FOR_COMPONENTS(c) DOCMP2 { w[c][cmp] = f_w[c][cmp] ? f_w[c][cmp] : f[c][cmp]; }

  for (polarization_state *p = pol[ft]; p; p = p->next) {
    const bool gyro = typeid(*p->s) == typeid(gyrotropic_susceptibility);
    const bool noisy = typeid(*p->s) == typeid(noisy_lorentzian_susceptibility);
    if (!gyro && !noisy && typeid(*p->s) != typeid(lorentzian_susceptibility))
      meep::abort("meep_b200: only lorentzian_susceptibility (Lorentz/Drude, also noisy) and "
                  "gyrotropic_susceptibility polarisations are supported on the device path");

    // Lazily allocate internal polarization data (host block laid out by the reference;
    // the device twin starts at zero exactly like init_internal_data's memset):
    if (!p->data) {
      p->data = p->s->new_internal_data(f, gv);
      if (p->data) {
        p->s->init_internal_data(f, dt, gv, p->data);
        const std::pair<realnum *, size_t> blk = polarisation_block(p->s, p->data);
        if (blk.second) {
          E->ensure_from(blk.first, blk.second, NULL);
          // init_internal_data just memset the block: hand the zero pages back (they read as zero
          // again on demand) — six Au poles are three times the field arrays' footprint
          if (E->lazy_host) release_interior(blk.first, blk.second);
        }
        allocated_fields = true;
      }
    }

    // Finally, timestep the polarizations (emits jobs):
    if (gyro)
      static_cast<const gyrotropic_susceptibility *>(p->s)->gyrotropic_susceptibility::update_P(
          w, f_w_prev, dt, gv, p->data);
    else if (noisy)
      static_cast<const noisy_lorentzian_susceptibility *>(p->s)->noisy_lorentzian_susceptibility::update_P(
          w, f_w_prev, dt, gv, p->data);
    else
      static_cast<const lorentzian_susceptibility *>(p->s)->lorentzian_susceptibility::update_P(
          w, f_w_prev, dt, gv, p->data);
  }

  return allocated_fields;
}

// Job emission for one susceptibility on one chunk.  W holds HOST pointers (as in the
// reference); they are translated to device addresses through the Engine's mirror table.
void lorentzian_susceptibility::update_P(realnum *W[NUM_FIELD_COMPONENTS][2],
                                         realnum *W_prev[NUM_FIELD_COMPONENTS][2], realnum dt,
                                         const grid_volume &gv, void *P_internal_data) const {
  Engine *E = Engine::current();
  if (!E || !E->recording())
    meep::abort("meep_b200: lorentzian_susceptibility::update_P called outside a device phase "
                "(this build has no CPU time-stepping path)");
  if (!P_internal_data) return;
  Recorder &R = E->rec();
  lorentzian_data_layout *d = (lorentzian_data_layout *)P_internal_data;
  // constants in realnum arithmetic, exactly as src/susceptibility.cpp:192-195
  const realnum omega2pi = 2 * pi * omega_0, g2pi = gamma * 2 * pi;
  const realnum omega0dtsqr = omega2pi * omega2pi * dt * dt;
  const realnum gamma1inv = 1 / (1 + g2pi * dt / 2), gamma1 = (1 - g2pi * dt / 2);
  const realnum omega0dtsqr_denom = no_omega_0_denominator ? 0 : omega0dtsqr;
  (void)W_prev; // unused;

  FOR_COMPONENTS(c) DOCMP2 {
    if (d->P[c][cmp]) {
      const realnum *w = W[c][cmp], *s = sigma[c][component_direction(c)];
      if (w && s) {
        realnum *p = d->P[c][cmp], *pp = d->P_prev[c][cmp];

        // directions/strides for offdiagonal terms, similar to update_eh
        const direction dd = component_direction(c);
        const ptrdiff_t is = gv.stride(dd) * (is_magnetic(c) ? -1 : +1);
        direction d1 = cycle_direction(gv.dim, dd, 1);
        component c1 = direction_component(c, d1);
        ptrdiff_t is1 = gv.stride(d1) * (is_magnetic(c) ? -1 : +1);
        const realnum *w1 = W[c1][cmp];
        const realnum *s1 = w1 ? sigma[c][d1] : NULL;
        direction d2 = cycle_direction(gv.dim, dd, 2);
        component c2 = direction_component(c, d2);
        ptrdiff_t is2 = gv.stride(d2) * (is_magnetic(c) ? -1 : +1);
        const realnum *w2 = W[c2][cmp];
        const realnum *s2 = w2 ? sigma[c][d2] : NULL;

        if (s2 && !s1) { // make s1 the non-NULL one if possible
          std::swap(d1, d2);
          std::swap(c1, c2);
          std::swap(is1, is2);
          std::swap(w1, w2);
          std::swap(s1, s2);
        }
        mb200_lorentz_job_t J;
        memset(&J, 0, sizeof(J));
        J.box = make_box(gv, gv.little_owned_corner(c), gv.big_corner()); // PLOOP_OVER_VOL_OWNED
        J.p = E->dev(p);
        J.pp = E->dev(pp);
        J.w = E->dev(w);
        J.s = E->dev(s);
        if (s1) {
          J.w1 = E->dev(w1);
          J.s1 = E->dev(s1);
        }
        if (s1 && s2) {
          J.w2 = E->dev(w2);
          J.s2 = E->dev(s2);
        }
        J.is = is;
        J.is1 = is1;
        J.is2 = is2;
        J.gamma1inv = gamma1inv;
        J.gamma1 = gamma1;
        J.omega0dtsqr = omega0dtsqr;
        J.omega0dtsqr_denom = omega0dtsqr_denom;
        J.ntot = (int64_t)gv.ntot();
        if (!s1 && gv.dim == D3 && !E->suppress_zero_skip) { // isotropic, standard 3-D layout: zero-block skipping
          // (flags start as "unknown"; the kernel establishes them with its first update)
          J.pzero = E->pzero_flags(p, gv.ntot(), false);
          J.szero = J.pzero ? E->szero_flags(s, gv.ntot()) : NULL;
          if (!J.szero) J.pzero = NULL;
        }
        if (!J.pzero) E->pzero_drop(p); // this array is updated by the plain kernel: no flags
        if (J.box.n[0] > 0 && J.box.n[1] > 0 && J.box.n[2] > 0) R.lorentz.push_back(J);
      }
    }
  }
}

// The reference calls this from fields_chunk::update_eh; our update_eh folds the subtraction
// into its f_minus_p job (update_eh.cpp), so reaching this function means some CPU code path
// is trying to time-step on stale host arrays.
void lorentzian_susceptibility::subtract_P(field_type, realnum *[NUM_FIELD_COMPONENTS][2],
                                           void *) const {
  meep::abort("meep_b200: lorentzian_susceptibility::subtract_P: this build has no CPU "
              "time-stepping path");
}

// noisy_lorentzian_susceptibility::update_P (reference src/susceptibility.cpp:317-339): the
// Lorentzian update, then p[i] += gaussian_random(0, amp sqrt(sigma[i])).  The numbers are drawn
// by the reference's generator on the host at every run (Engine::run), here only the jobs and the
// generators are recorded.
void noisy_lorentzian_susceptibility::update_P(realnum *W[NUM_FIELD_COMPONENTS][2],
                                               realnum *W_prev[NUM_FIELD_COMPONENTS][2], realnum dt,
                                               const grid_volume &gv, void *P_internal_data) const {
  Engine *E = Engine::current();
  if (!E || !E->recording())
    meep::abort("meep_b200: noisy_lorentzian_susceptibility::update_P outside a phase: this build has no "
                "CPU time-stepping path");
  // (the noise makes P non-zero wherever sigma is: no zero-block bookkeeping for these arrays)
  E->suppress_zero_skip = true;
  lorentzian_susceptibility::update_P(W, W_prev, dt, gv, P_internal_data);
  E->suppress_zero_skip = false;
  if (!P_internal_data) return;
  Recorder &R = E->rec();
  lorentzian_data_layout *d = (lorentzian_data_layout *)P_internal_data;

  const realnum g2pi = gamma * 2 * pi;
  const realnum w2pi = omega_0 * 2 * pi;
  const realnum amp = w2pi * noise_amp * sqrt(g2pi) * dt * dt / (1 + g2pi * dt / 2);

  FOR_COMPONENTS(c) DOCMP2 {
    if (d->P[c][cmp]) {
      const realnum *s = sigma[c][component_direction(c)];
      if (s) {
        mb200_noise_job_t J;
        memset(&J, 0, sizeof(J));
        J.box = make_box(gv, gv.little_owned_corner(c), gv.big_corner()); // LOOP_OVER_VOL_OWNED
        if (J.box.n[0] <= 0 || J.box.n[1] <= 0 || J.box.n[2] <= 0) continue;
        J.p = E->dev(d->P[c][cmp]);
        J.slot = 0;
        for (const NoiseGen &g : R.noise_gens)
          J.slot += (int64_t)g.box.n[0] * g.box.n[1] * g.box.n[2];
        NoiseGen g;
        g.amp = amp;
        g.sigma = s;
        g.box = J.box;
        R.noise.push_back(J);
        R.noise_gens.push_back(g);
      }
    }
  }
}

// gyrotropic_susceptibility::update_P (reference src/susceptibility.cpp:445-584): same constants,
// computed in realnum as there; one mb200_gyro_job_t per (component, cmp), arrays and tensors
// handed over in the rotated frame (d0, d1, d2) of that component.
void gyrotropic_susceptibility::update_P(realnum *W[NUM_FIELD_COMPONENTS][2],
                                         realnum *W_prev[NUM_FIELD_COMPONENTS][2], realnum dt,
                                         const grid_volume &gv, void *P_internal_data) const {
  Engine *E = Engine::current();
  if (!E || !E->recording())
    meep::abort("meep_b200: gyrotropic_susceptibility::update_P outside a phase: this build has no CPU "
                "time-stepping path");
  if (!P_internal_data) return;
  Recorder &R = E->rec();
  gyrotropy_data_layout *d = (gyrotropy_data_layout *)P_internal_data;
  const realnum omega2pidt = 2 * pi * omega_0 * dt;
  const realnum g2pidt = 2 * pi * gamma * dt;
  (void)W_prev; // unused;

  realnum c4[4], gd, gx, gy, gz;
  int model_id;
  switch (model) {
    case GYROTROPIC_LORENTZIAN:
    case GYROTROPIC_DRUDE: {
      const realnum omega0dtsqr = omega2pidt * omega2pidt;
      const realnum gamma1 = (1 - g2pidt / 2);
      const realnum diag = 2 - (model == GYROTROPIC_DRUDE ? 0 : omega0dtsqr);
      const realnum pt = pi * dt;
      gd = (1 + g2pidt / 2);
      gx = pt * gyro_tensor[Y][Z];
      gy = pt * gyro_tensor[Z][X];
      gz = pt * gyro_tensor[X][Y];
      c4[0] = diag;
      c4[1] = gamma1;
      c4[2] = omega0dtsqr;
      c4[3] = pt;
      model_id = 0;
    } break;
    case GYROTROPIC_SATURATED: {
      const realnum dt2pi = 2 * pi * dt;
      gd = 0.5;
      gx = -0.5 * alpha * gyro_tensor[Y][Z];
      gy = -0.5 * alpha * gyro_tensor[Z][X];
      gz = -0.5 * alpha * gyro_tensor[X][Y];
      c4[0] = omega2pidt;
      c4[1] = g2pidt;
      c4[2] = alpha;
      c4[3] = dt2pi;
      model_id = 1;
    } break;
    default: meep::abort("meep_b200: unknown gyrotropy model"); return;
  }
  // Precalculate 3x3 matrix inverse, exploiting skew symmetry (lines 466-474 = 519-527)
  const realnum invdet = 1.0 / gd / (gd * gd + gx * gx + gy * gy + gz * gz);
  const realnum inv[3][3] = {{invdet * (gd * gd + gx * gx), invdet * (gx * gy + gd * gz),
                              invdet * (gx * gz - gd * gy)},
                             {invdet * (gy * gx - gd * gz), invdet * (gd * gd + gy * gy),
                              invdet * (gy * gz + gd * gx)},
                             {invdet * (gz * gx + gd * gy), invdet * (gz * gy - gd * gx),
                              invdet * (gd * gd + gz * gz)}};

  FOR_COMPONENTS(c) DOCMP2 {
    if (d->P[c][cmp][0]) {
      const direction d0 = component_direction(c);
      const realnum *w0 = W[c][cmp], *s = sigma[c][d0];
      if (!w0 || !s || (d0 != X && d0 != Y && d0 != Z))
        meep::abort("gyrotropic media require 3D Cartesian fields\n");
      const direction d1 = cycle_direction(gv.dim, d0, 1);
      const direction d2 = cycle_direction(gv.dim, d0, 2);
      const realnum *w1 = W[direction_component(c, d1)][cmp];
      const realnum *w2 = W[direction_component(c, d2)][cmp];
      const direction ds[3] = {d0, d1, d2};
      if (!d->P_prev[c][cmp][d1] || !d->P_prev[c][cmp][d2])
        meep::abort("gyrotropic media require 3D Cartesian fields\n");
      if (sigma[c][d1] || sigma[c][d2])
        meep::abort("gyrotropic media do not support anisotropic sigma\n");
      mb200_gyro_job_t J;
      memset(&J, 0, sizeof(J));
      J.box = make_box(gv, gv.little_owned_corner(c), gv.big_corner()); // LOOP_OVER_VOL_OWNED
      for (int a = 0; a < 3; ++a) {
        J.p[a] = E->dev(d->P[c][cmp][ds[a]]);
        J.pp[a] = E->dev(d->P_prev[c][cmp][ds[a]]);
        for (int b = 0; b < 3; ++b) {
          J.gt[a][b] = gyro_tensor[ds[a]][ds[b]];
          J.inv[a][b] = inv[ds[a]][ds[b]];
        }
      }
      J.w[0] = E->dev(w0);
      J.w[1] = E->dev(w1);
      J.w[2] = E->dev(w2);
      J.s = E->dev(s);
      J.is = gv.stride(d0) * (is_magnetic(c) ? -1 : +1);
      J.is1 = gv.stride(d1) * (is_magnetic(c) ? -1 : +1);
      J.is2 = gv.stride(d2) * (is_magnetic(c) ? -1 : +1);
      for (int k = 0; k < 4; ++k)
        J.c[k] = c4[k];
      J.model = model_id;
      if (J.box.n[0] > 0 && J.box.n[1] > 0 && J.box.n[2] > 0) R.gyro.push_back(J);
    }
  }
}

// (folded into the f_minus_p job of update_eh.cpp, like the Lorentzian one)
void gyrotropic_susceptibility::subtract_P(field_type, realnum *[NUM_FIELD_COMPONENTS][2],
                                           void *) const {
  meep::abort("meep_b200: gyrotropic_susceptibility::subtract_P: this build has no CPU "
              "time-stepping path");
}

} // namespace meep
