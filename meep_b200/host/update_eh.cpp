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
#include "hostmem.hpp"
#include "loop_desc.hpp"
#include "meep_internals.hpp"

using namespace std;
using namespace meep_b200;

namespace meep {

namespace {
// sign of the neighbour offsets used by the off-diagonal averages: E looks at +stride, H at -stride
inline ptrdiff_t signed_stride(const grid_volume &gv, direction d, field_type ft) {
  return ft == H_stuff ? -gv.stride(d) : gv.stride(d);
}
} // namespace

// ---- fields::update_eh: tile bookkeeping (src/update_eh.cpp:31-50), then one recorded phase ------
void fields::update_eh(field_type ft, bool skip_w_components) {
  if (ft != E_stuff && ft != H_stuff) meep::abort("update_eh only works with E/H");
  Engine &E = Engine::get(this);
  Scope scope(E, this);

  // The reference may cut an anisotropic chunk into tiles for cache reuse; the tiles only change
  // the order of independent point updates, but they are part of the chunk state other code can
  // look at, so they are maintained exactly as there.
  bool solving_cw = false;
  for (int ic = 0; ic < num_chunks; ic++) {
    fields_chunk *fc = chunks[ic];
    if (!fc->is_mine()) continue;
    solving_cw = solving_cw || fc->doing_solve_cw;
    std::vector<grid_volume> &tiles = fc->gvs_eh[ft];
    if (!changed_materials && !tiles.empty()) continue;
    bool offdiag_both = false; // a component with both off-diagonal chi1inv arrays
    FOR_FT_COMPONENTS(ft, cc) {
      const direction dc = component_direction(cc);
      offdiag_both = offdiag_both || (fc->s->chi1inv[cc][cycle_direction(fc->gv.dim, dc, 1)] &&
                                      fc->s->chi1inv[cc][cycle_direction(fc->gv.dim, dc, 2)]);
    }
    const size_t had = tiles.size();
    tiles.clear();
    if (offdiag_both && loop_tile_base_eh > 0) {
      split_into_tiles(fc->gv, &tiles, loop_tile_base_eh);
      check_tiles(fc->gv, tiles);
    }
    else
      tiles.push_back(fc->gv);
    if (tiles.size() != had) E.phase(PH_EH, ft).valid = false;
  }

  const bool cacheable = E.in_step && !skip_w_components && !solving_cw;
  run_phase(E, this, PH_EH, ft, cacheable, [&]() {
    for (int ic = 0; ic < num_chunks; ic++) {
      if (!chunks[ic]->is_mine()) continue;
      if (chunks[ic]->update_eh(ft, skip_w_components)) {
        chunk_connections_valid = false; // new E/H arrays: the chunks must be reconnected
        assert(changed_materials);
      }
    }
  });
}

bool fields_chunk::needs_W_prev(component c) const {
  bool any = false;
  for (susceptibility *sus = s->chiP[type(c)]; sus && !any; sus = sus->next)
    any = sus->needs_W_prev();
  return any;
}

// ---- fields_chunk::update_eh: job emission for one chunk (src/update_eh.cpp:66-215) ---------------
bool fields_chunk::update_eh(field_type ft, bool skip_w_components) {
  Engine *E = Engine::current();
  if (!E || !E->recording()) meep::abort("meep_b200: fields_chunk::update_eh outside a phase");
  Recorder &R = E->rec();
  const field_type ft_db = ft == E_stuff ? D_stuff : B_stuff; // the D/B family feeding this update
  const size_t npts = gv.ntot(), nbytes = npts * sizeof(realnum);
  bool allocated_eh = false;

  // ---- stage 0: which D/B components need a scratch copy "D minus polarisations" -------------
  // (needed when a polarisation contributes to the component, or when integrated sources exist)
  bool integrated_sources = false;
  if (!doing_solve_cw)
    for (const src_vol &sv : sources[ft_db])
      integrated_sources = integrated_sources || sv.t()->is_integrated;

  bool any_fmp = false;
  FOR_FT_COMPONENTS(ft, ec) {
    const component dc = field_type_component(ft_db, ec);
    DOCMP {
      bool wanted = false;
      if (f[ec][cmp]) {
        wanted = integrated_sources;
        for (polarization_state *p = pol[ft]; p && !wanted; p = p->next)
          wanted = p->s->needs_P(ec, cmp, f);
      }
      realnum *&scratch = f_minus_p[dc][cmp];
      if (wanted && !scratch) {
        scratch = new_zeroed_lazily(npts);
        E->ensure(scratch, nbytes, true, 1);
      }
      else if (!wanted && scratch) { // no longer needed: release host and device copies
        E->forget(scratch);
        delete[] scratch;
        scratch = 0;
      }
    }
  }
  FOR_FT_COMPONENTS(ft_db, dc) { any_fmp = any_fmp || f_minus_p[dc][0] != 0; }
  if (any_fmp && doing_solve_cw)
    meep::abort("dispersive materials are not yet implemented for solve_cw");

  // ---- stage 1: scratch = D - sum of polarisations, one fused pass per array ------------------
  // (reference: memcpy followed by subtract_P of every susceptibility, src/update_eh.cpp:114-123)
  const size_t ntot_s = s->gv.ntot();
  FOR_FT_COMPONENTS(ft, ec) {
    if (!f[ec][0]) continue;
    const component dc = field_type_component(ft_db, ec);
    DOCMP {
      if (!f_minus_p[dc][cmp]) continue;
      mb200_fmp_job_t J;
      memset(&J, 0, sizeof(J));
      J.fmp = E->dev(f_minus_p[dc][cmp]);
      J.d = E->dev(f[dc][cmp]);
      J.ntot = (int64_t)ntot_s;
      for (polarization_state *p = pol[ft]; p; p = p->next) {
        if (!p->data) continue;
        const realnum *P = nullptr;
        if (typeid(*p->s) == typeid(lorentzian_susceptibility) ||
            typeid(*p->s) == typeid(noisy_lorentzian_susceptibility))
          P = ((const lorentzian_data_layout *)p->data)->P[ec][cmp];
        else if (typeid(*p->s) == typeid(gyrotropic_susceptibility)) // (src/susceptibility.cpp:586-602)
          P = ((const gyrotropy_data_layout *)p->data)->P[ec][cmp][component_direction(ec)];
        else
          meep::abort("meep_b200: only lorentzian_susceptibility and gyrotropic_susceptibility "
                      "polarisations are supported");
        if (!P) continue;
        if (J.np == MB200_MAX_P) { // job full: flush it and keep subtracting in place
          R.fmp.push_back(J);
          J.d = nullptr;
          J.np = 0;
          memset(J.pzero, 0, sizeof(J.pzero));
        }
        J.pzero[J.np] = E->pzero_lookup(P);
        J.p[J.np++] = E->dev(P);
      }
      R.fmp.push_back(J);
    }
  }

  // ---- stage 2: subtract the dipole moments of integrated sources (src/update_eh.cpp:128-138) ---
  if (any_fmp && !doing_solve_cw)
    for (const src_vol &sv : sources[ft_db]) {
      if (!sv.t()->is_integrated || !f[sv.c][0] || ft != type(sv.c)) continue;
      const size_t np = sv.num_points();
      if (!np) continue;
      const component c = field_type_component(ft_db, sv.c);
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

  // ---- stage 3: E = chi1inv (D - P), one job per (tile, part, component) ------------------------
  // operand of the update for a D/B component: the scratch copy where there is one
  auto operand = [&](component dc, int cmp) -> realnum * {
    return f_minus_p[dc][cmp] ? f_minus_p[dc][cmp] : f[dc][cmp];
  };
  const std::vector<grid_volume> &tiles = gvs_eh[ft];
  for (size_t it = 0; it < tiles.size(); ++it) {
    const bool first_tile = it == 0;
    DOCMP FOR_FT_COMPONENTS(ft, ec) {
      if (!f[ec][cmp]) continue;
      if (type(ec) != ft) meep::abort("bug in FOR_FT_COMPONENTS");
      const component dc = field_type_component(ft_db, ec);
      // the component's own direction and the two cyclic ones (off-diagonal partners)
      direction dir[3];
      component dcomp[3];
      ptrdiff_t stride[3];
      for (int k = 0; k < 3; ++k) {
        dir[k] = k == 0 ? component_direction(ec) : cycle_direction(gv.dim, component_direction(ec), k);
        dcomp[k] = k == 0 ? dc : direction_component(dc, dir[k]);
        stride[k] = signed_stride(gv, dir[k], ft);
      }
      const direction dsigw = s->sigsize[dir[0]] > 1 ? dir[0] : NO_DIRECTION; // PML along the component

      if (first_tile) {
        // E/H start out aliased to D/B; split them when the update is not the identity
        if (f[ec][cmp] == f[dc][cmp] && (s->chi1inv[ec][dir[0]] || any_fmp || dsigw != NO_DIRECTION)) {
          // (device twin copied from the device copy of D/B; the host array is filled by the next download)
          f[ec][cmp] = new_zeroed_lazily(npts);
          E->ensure_from(f[ec][cmp], nbytes, f[dc][cmp]);
          allocated_eh = true;
        }
        // the W auxiliary field of the PML ODE
        if (dsigw != NO_DIRECTION && !f_w[ec][cmp]) {
          f_w[ec][cmp] = new_zeroed_lazily(npts);
          E->ensure_from(f_w[ec][cmp], nbytes, f[ec][cmp]);
          if (needs_W_notowned(ec)) allocated_eh = true; // its halo must be communicated
        }
      }
      if (skip_w_components && f_w[ec][cmp]) continue; // solve_cw supplies W and E itself
      if (first_tile && needs_W_prev(ec))
        meep::abort("meep_b200: susceptibilities that need W_prev are not supported on the "
                    "device path");
      if (f[ec][cmp] == f[dc][cmp]) continue; // still aliased: nothing to compute

      // a job over [lo, hi] with or without the off-diagonal operands
      auto make_job = [&](const ivec &lo, const ivec &hi, bool offdiag) {
        mb200_edhb_job_t J;
        memset(&J, 0, sizeof(J));
        J.box = make_box(gv, lo, hi);
        J.f = E->dev(f[ec][cmp]);
        J.g = E->dev(operand(dcomp[0], cmp));
        J.u = E->dev(s->chi1inv[ec][dir[0]]);
        J.s = stride[0];
        J.s1 = stride[1];
        J.s2 = stride[2];
        if (offdiag) {
          realnum *o1 = operand(dcomp[1], cmp), *o2 = operand(dcomp[2], cmp);
          J.g1 = E->dev(o1);
          J.g2 = E->dev(o2);
          J.u1 = o1 ? E->dev(s->chi1inv[ec][dir[1]]) : NULL;
          J.u2 = o2 ? E->dev(s->chi1inv[ec][dir[2]]) : NULL;
        }
        J.chi2 = E->dev(s->chi2[ec]);
        J.chi3 = E->dev(s->chi3[ec]);
        J.fw = E->dev(f_w[ec][cmp]);
        if (dsigw != NO_DIRECTION)
          J.pmlw = make_pml(gv, lo, dsigw, E->dev(s->sig[dsigw]), E->dev(s->kap[dsigw]), NULL);
        return J;
      };

      mb200_edhb_job_t J = make_job(tiles[it].little_owned_corner0(ec), tiles[it].big_corner(), true);
      // the kernel wants the first off-diagonal slot filled first (src/step_generic.cpp:573-577)
      if ((!J.g1 && J.g2) || (J.g1 && J.g2 && !J.u1 && J.u2)) {
        std::swap(J.g1, J.g2);
        std::swap(J.u1, J.u2);
        std::swap(J.s1, J.s2);
      }
      if (!J.u1 && J.u2) meep::abort("bug - didn't swap off-diagonal terms!?");
      bool emit = true;
      if (E->in_step) {
        // this chunk's E/H update was folded into the D/B pass (step_db.cpp), except for the
        // planes that hold source points: restrict the job to those planes
        auto fz = E->fused_eh[ft].find(this);
        if (fz != E->fused_eh[ft].end()) {
          const int slab_lo = fz->second.first, slab_hi = fz->second.second;
          const int ix0 = (int)(J.box.idx0 / J.box.s[0]); // 3-D: loop 1 is X
          const int lo = std::max(slab_lo, ix0), hi = std::min(slab_hi, ix0 + J.box.n[0] - 1);
          if (slab_lo > slab_hi || lo > hi)
            emit = false;
          else {
            J.box.idx0 += (int64_t)(lo - ix0) * J.box.s[0];
            J.pmlw.k0 += J.pmlw.ks[0] * (lo - ix0);
            J.box.n[0] = hi - lo + 1;
          }
        }
      }
      if (!emit) continue;
      if (J.box.n[0] > 0 && J.box.n[1] > 0 && J.box.n[2] > 0) R.edhb.push_back(J);

      if (gv.dim == Dcyl) { // the r = 0 row is updated separately (src/update_eh.cpp:197-209):
        // the off-diagonal terms must vanish there for an axisymmetric structure
        const ivec lo0 = tiles[it].little_owned_corner(ec);
        if (lo0.r() == 0) {
          ivec hi0 = tiles[it].big_corner();
          hi0.set_direction(meep::R, 0);
          const mb200_edhb_job_t J0 = make_job(lo0, hi0, false);
          if (J0.box.n[0] > 0 && J0.box.n[1] > 0 && J0.box.n[2] > 0) R.edhb.push_back(J0);
        }
      }
    }
  }

  return allocated_eh;
}

} // namespace meep
