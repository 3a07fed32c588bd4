// update_eh.cpp — B200 replacement for the reference translation unit src/update_eh.cpp.
//
// Same decisions as the reference (which components need f_minus_p, lazy split of H from B,
// lazy f_w, off-diagonal chi1inv pointers: src/update_eh.cpp:66-215); the three stages
//   f_minus_p = D - sum P        (memcpy + subtract_P, lines 114-123)   -> mb200_fmp_job_t
//   f_minus_p -= dipole sources  (lines 128-138)                         -> mb200_src_job_t mode 1
//   E = chi1inv * f_minus_p      (STEP_UPDATE_EDHB, lines 190-195)       -> mb200_edhb_job_t
// are emitted as jobs and executed as at most three launches for all chunks together.
#include <assert.h>
#include <algorithm>
#include <string.h>
#include <typeinfo>

#include "engine.hpp"
#include "loop_desc.hpp"
#include "meep_internals.hpp"

using namespace std;
using namespace meep_b200;

namespace meep {

void fields::update_eh(field_type ft, bool skip_w_components) {
  if (ft != E_stuff && ft != H_stuff) meep::abort("update_eh only works with E/H");
  Engine &E = Engine::get(this);
  Scope scope(E, this);

  // split the chunks' volume into subdomains for tiled execution of update_eh loop
  // (src/update_eh.cpp:31-50; the tiles only change the order of independent point updates)
  for (int i = 0; i < num_chunks; i++)
    if (chunks[i]->is_mine() && (changed_materials || chunks[i]->gvs_eh[ft].empty())) {
      bool is_aniso = false;
      FOR_FT_COMPONENTS(ft, cc) {
        const direction d_c = component_direction(cc);
        const direction d_1 = cycle_direction(chunks[i]->gv.dim, d_c, 1);
        const direction d_2 = cycle_direction(chunks[i]->gv.dim, d_c, 2);
        if (chunks[i]->s->chi1inv[cc][d_1] && chunks[i]->s->chi1inv[cc][d_2]) {
          is_aniso = true;
          break;
        }
      }
      const size_t ntiles_before = chunks[i]->gvs_eh[ft].size();
      if (!chunks[i]->gvs_eh[ft].empty()) chunks[i]->gvs_eh[ft].clear();
      if (loop_tile_base_eh > 0 && is_aniso) {
        split_into_tiles(chunks[i]->gv, &chunks[i]->gvs_eh[ft], loop_tile_base_eh);
        check_tiles(chunks[i]->gv, chunks[i]->gvs_eh[ft]);
      }
      else { chunks[i]->gvs_eh[ft].push_back(chunks[i]->gv); }
      if (chunks[i]->gvs_eh[ft].size() != ntiles_before) E.phase(PH_EH, ft).valid = false;
    }

  bool cw = false;
  for (int i = 0; i < num_chunks; i++)
    if (chunks[i]->is_mine() && chunks[i]->doing_solve_cw) cw = true;

  run_phase(E, this, PH_EH, ft, E.in_step && !skip_w_components && !cw, [&]() {
    for (int i = 0; i < num_chunks; i++)
      if (chunks[i]->is_mine())
        if (chunks[i]->update_eh(ft, skip_w_components)) {
          chunk_connections_valid = false; // E/H allocated - reconnect chunks
          assert(changed_materials);
        }
  });
}

bool fields_chunk::needs_W_prev(component c) const {
  for (susceptibility *chiP = s->chiP[type(c)]; chiP; chiP = chiP->next)
    if (chiP->needs_W_prev()) return true;
  return false;
}

bool fields_chunk::update_eh(field_type ft, bool skip_w_components) {
  Engine *E = Engine::current();
  if (!E || !E->recording()) meep::abort("meep_b200: fields_chunk::update_eh outside a phase");
  Recorder &R = E->rec();
  field_type ft2 = ft == E_stuff ? D_stuff : B_stuff; // for sources etc.
  bool allocated_eh = false;
  const size_t nbytes = gv.ntot() * sizeof(realnum);

  bool have_int_sources = false;
  if (!doing_solve_cw) {
    for (const src_vol &sv : sources[ft2]) {
      if (sv.t()->is_integrated) {
        have_int_sources = true;
        break;
      }
    }
  }

  FOR_FT_COMPONENTS(ft, ec) {
    component dc = field_type_component(ft2, ec);
    DOCMP {
      bool need_fmp = false;
      if (f[ec][cmp]) {
        need_fmp = have_int_sources;
        for (polarization_state *p = pol[ft]; p && !need_fmp; p = p->next)
          need_fmp = need_fmp || p->s->needs_P(ec, cmp, f);
      }
      if (need_fmp) {
        if (!f_minus_p[dc][cmp]) {
          f_minus_p[dc][cmp] = new realnum[gv.ntot()];
          memset(f_minus_p[dc][cmp], 0, nbytes);
          E->ensure(f_minus_p[dc][cmp], nbytes, true, 1);
        }
      }
      else if (f_minus_p[dc][cmp]) { // remove unneeded f_minus_p
        E->forget(f_minus_p[dc][cmp]);
        delete[] f_minus_p[dc][cmp];
        f_minus_p[dc][cmp] = 0;
      }
    }
  }
  bool have_f_minus_p = false;
  FOR_FT_COMPONENTS(ft2, dc) {
    if (f_minus_p[dc][0]) {
      have_f_minus_p = true;
      break;
    }
  }

  const size_t ntot = s->gv.ntot();

  if (have_f_minus_p && doing_solve_cw)
    meep::abort("dispersive materials are not yet implemented for solve_cw");

  //////////////////////////////////////////////////////////////////////////
  // First, initialize f_minus_p to D - P, if necessary (one fused pass per array)

  FOR_FT_COMPONENTS(ft, ec) if (f[ec][0]) {
    component dc = field_type_component(ft2, ec);
    DOCMP if (f_minus_p[dc][cmp]) {
      mb200_fmp_job_t J;
      memset(&J, 0, sizeof(J));
      J.fmp = E->dev(f_minus_p[dc][cmp]);
      J.d = E->dev(f[dc][cmp]);
      J.ntot = (int64_t)ntot;
      // subtract_P of every polarisation that has data (src/susceptibility.cpp:264-281)
      for (polarization_state *p = pol[ft]; p; p = p->next)
        if (p->data) {
          const realnum *P = nullptr;
          if (typeid(*p->s) == typeid(lorentzian_susceptibility) ||
              typeid(*p->s) == typeid(noisy_lorentzian_susceptibility))
            P = ((const lorentzian_data_layout *)p->data)->P[ec][cmp];
          else if (typeid(*p->s) == typeid(gyrotropic_susceptibility)) // (src/susceptibility.cpp:586-602)
            P = ((const gyrotropy_data_layout *)p->data)->P[ec][cmp][component_direction(ec)];
          else
            meep::abort("meep_b200: only lorentzian_susceptibility and gyrotropic_susceptibility "
                        "polarisations are supported");
          if (P) {
            if (J.np == MB200_MAX_P) { // flush and continue in place
              R.fmp.push_back(J);
              J.d = nullptr;
              J.np = 0;
              memset(J.pzero, 0, sizeof(J.pzero));
            }
            J.pzero[J.np] = E->pzero_lookup(P);
            J.p[J.np++] = E->dev(P);
          }
        }
      R.fmp.push_back(J);
    }
  }

  //////////////////////////////////////////////////////////////////////////
  // Next, subtract time-integrated sources (i.e. polarizations, not currents)

  if (have_f_minus_p && !doing_solve_cw) {
    for (const src_vol &sv : sources[ft2]) {
      if (sv.t()->is_integrated && f[sv.c][0] && ft == type(sv.c)) {
        component c = field_type_component(ft2, sv.c);
        const size_t np = sv.num_points();
        if (!np) continue;
        std::vector<int64_t> idx(np);
        std::vector<double> amp(2 * np);
        for (size_t j = 0; j < np; ++j) {
          idx[j] = (int64_t)sv.index_at(j);
          amp[2 * j] = sv.amplitude_at(j).real();
          amp[2 * j + 1] = sv.amplitude_at(j).imag();
        }
        mb200_src_job_t J;
        memset(&J, 0, sizeof(J));
        J.f_re = E->dev(f_minus_p[c][0]);
        J.f_im = is_real ? nullptr : E->dev(f_minus_p[c][1]);
        J.index = (const int64_t *)E->aux_upload(idx.data(), np * sizeof(int64_t));
        J.amp = (const double *)E->aux_upload(amp.data(), 2 * np * sizeof(double));
        J.npts = (int64_t)np;
        J.dt = dt;
        J.scalar_slot = (int32_t)R.dip_times.size();
        J.mode = 1;
        R.dip.push_back(J);
        R.dip_times.push_back(sv.t());
      }
    }
  }

  //////////////////////////////////////////////////////////////////////////
  // Finally, compute E = chi1inv * D

  realnum *dmp[NUM_FIELD_COMPONENTS][2];
  FOR_FT_COMPONENTS(ft2, dc) DOCMP2 {
    dmp[dc][cmp] = f_minus_p[dc][cmp] ? f_minus_p[dc][cmp] : f[dc][cmp];
  }

  for (size_t i = 0; i < gvs_eh[ft].size(); ++i) {
    DOCMP FOR_FT_COMPONENTS(ft, ec) {
      if (f[ec][cmp]) {
        if (type(ec) != ft) meep::abort("bug in FOR_FT_COMPONENTS");
        component dc = field_type_component(ft2, ec);
        const direction d_ec = component_direction(ec);
        const ptrdiff_t s_ec = gv.stride(d_ec) * (ft == H_stuff ? -1 : +1);
        const direction d_1 = cycle_direction(gv.dim, d_ec, 1);
        const component dc_1 = direction_component(dc, d_1);
        const ptrdiff_t s_1 = gv.stride(d_1) * (ft == H_stuff ? -1 : +1);
        const direction d_2 = cycle_direction(gv.dim, d_ec, 2);
        const component dc_2 = direction_component(dc, d_2);
        const ptrdiff_t s_2 = gv.stride(d_2) * (ft == H_stuff ? -1 : +1);

        direction dsigw0 = d_ec;
        direction dsigw = s->sigsize[dsigw0] > 1 ? dsigw0 : NO_DIRECTION;

        // lazily allocate any E/H fields that are needed (H==B initially)
        if (i == 0 && f[ec][cmp] == f[dc][cmp] &&
            (s->chi1inv[ec][d_ec] || have_f_minus_p || dsigw != NO_DIRECTION)) {
          f[ec][cmp] = new realnum[gv.ntot()];
          memcpy(f[ec][cmp], f[dc][cmp], nbytes);
          E->ensure_from(f[ec][cmp], nbytes, f[dc][cmp]);
          allocated_eh = true;
        }

        // lazily allocate W auxiliary field
        if (i == 0 && !f_w[ec][cmp] && dsigw != NO_DIRECTION) {
          f_w[ec][cmp] = new realnum[gv.ntot()];
          memcpy(f_w[ec][cmp], f[ec][cmp], nbytes);
          E->ensure_from(f_w[ec][cmp], nbytes, f[ec][cmp]);
          if (needs_W_notowned(ec)) allocated_eh = true; // communication needed
        }

        // for solve_cw, when W exists we get W and E from special variables
        if (f_w[ec][cmp] && skip_w_components) continue;

        if (i == 0 && needs_W_prev(ec))
          meep::abort("meep_b200: susceptibilities that need W_prev are not supported on the "
                      "device path");

        if (f[ec][cmp] != f[dc][cmp]) {
          const ivec is = gvs_eh[ft][i].little_owned_corner0(ec), ie = gvs_eh[ft][i].big_corner();
          mb200_edhb_job_t J;
          memset(&J, 0, sizeof(J));
          J.box = make_box(gv, is, ie);
          J.f = E->dev(f[ec][cmp]);
          J.g = E->dev(dmp[dc][cmp]);
          J.g1 = E->dev(dmp[dc_1][cmp]);
          J.g2 = E->dev(dmp[dc_2][cmp]);
          J.u = E->dev(s->chi1inv[ec][d_ec]);
          J.u1 = dmp[dc_1][cmp] ? E->dev(s->chi1inv[ec][d_1]) : NULL;
          J.u2 = dmp[dc_2][cmp] ? E->dev(s->chi1inv[ec][d_2]) : NULL;
          J.s = s_ec;
          J.s1 = s_1;
          J.s2 = s_2;
          J.chi2 = E->dev(s->chi2[ec]);
          J.chi3 = E->dev(s->chi3[ec]);
          J.fw = E->dev(f_w[ec][cmp]);
          if (dsigw != NO_DIRECTION)
            J.pmlw = make_pml(gv, is, dsigw, E->dev(s->sig[dsigw]), E->dev(s->kap[dsigw]), NULL);
          // swap g1 and g2 (src/step_generic.cpp:573-577)
          if ((!J.g1 && J.g2) || (J.g1 && J.g2 && !J.u1 && J.u2)) {
            std::swap(J.g1, J.g2);
            std::swap(J.u1, J.u2);
            std::swap(J.s1, J.s2);
          }
          if (!J.u1 && J.u2) meep::abort("bug - didn't swap off-diagonal terms!?");
          if (E->in_step) {
            // this chunk's E/H update was folded into the D/B pass (step_db.cpp), except for
            // the planes that hold source points
            auto fz = E->fused_eh[ft].find(this);
            if (fz != E->fused_eh[ft].end()) {
              const int slab_lo = fz->second.first, slab_hi = fz->second.second;
              if (slab_lo > slab_hi) continue;
              // 3-D: loop 1 is X; first loop plane has array index idx0 / stride_x
              const int ix0 = (int)(J.box.idx0 / J.box.s[0]);
              const int lo = std::max(slab_lo, ix0), hi = std::min(slab_hi, ix0 + J.box.n[0] - 1);
              if (lo > hi) continue;
              J.box.idx0 += (int64_t)(lo - ix0) * J.box.s[0];
              J.pmlw.k0 += J.pmlw.ks[0] * (lo - ix0);
              J.box.n[0] = hi - lo + 1;
            }
          }
          if (J.box.n[0] > 0 && J.box.n[1] > 0 && J.box.n[2] > 0) R.edhb.push_back(J);

          if (gv.dim == Dcyl) { // the r = 0 row (src/update_eh.cpp:197-209)
            const ivec is0 = gvs_eh[ft][i].little_owned_corner(ec);
            if (is0.r() == 0) {
              ivec ie0 = gvs_eh[ft][i].big_corner();
              ie0.set_direction(meep::R, 0);
              /* NULL off-diagonal terms: they must be zero at r=0 for an axisymmetric structure */
              mb200_edhb_job_t J0;
              memset(&J0, 0, sizeof(J0));
              J0.box = make_box(gv, is0, ie0);
              J0.f = E->dev(f[ec][cmp]);
              J0.g = E->dev(dmp[dc][cmp]);
              J0.u = E->dev(s->chi1inv[ec][d_ec]);
              J0.s = s_ec;
              J0.s1 = s_1;
              J0.s2 = s_2;
              J0.chi2 = E->dev(s->chi2[ec]);
              J0.chi3 = E->dev(s->chi3[ec]);
              J0.fw = E->dev(f_w[ec][cmp]);
              if (dsigw != NO_DIRECTION)
                J0.pmlw = make_pml(gv, is0, dsigw, E->dev(s->sig[dsigw]), E->dev(s->kap[dsigw]), NULL);
              if (J0.box.n[0] > 0 && J0.box.n[1] > 0 && J0.box.n[2] > 0) R.edhb.push_back(J0);
            }
          }
        }
      }
    }
  }

  return allocated_eh;
}

} // namespace meep
