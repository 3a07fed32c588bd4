// step_db.cpp — B200 replacement for the reference translation unit src/step_db.cpp.
//
// fields_chunk::step_db keeps the reference's per-component set-up (which arrays feed which
// curl term, PML directions, lazy allocation of f_u / f_cond: src/step_db.cpp:44-127) but
// instead of calling STEP_CURL it emits one mb200_curl_job_t per call; the Engine batches the
// jobs of all chunks into one launch (and fuses the three components of a 3-D chunk).
#include <assert.h>
#include <string.h>

#include "engine.hpp"
#include "hostmem.hpp"
#include "loop_desc.hpp"
#include "meep_internals.hpp"
#include <stdio.h>
#include <vector>

using namespace std;
using namespace meep_b200;

namespace meep {

void fields::step_db(field_type ft) {
  if (ft != B_stuff && ft != D_stuff) meep::abort("step_db only works with B/D");
  Engine &E = Engine::get(this);
  Scope scope(E, this);
  // (called by reference/user code outside fields::step(): never cached, never fused with E/H)
  E.connections_valid = chunk_connections_valid;
  run_phase(E, this, PH_DB, ft, E.in_step, [&]() {
    // the E/H fusion decisions are remade with this recording: the matching update_eh phase
    // must be re-recorded against them
    // (only for the cached in-step plans; a stand-alone call from reference code is recorded
    // unfused, run once and discarded, and must leave the in-step plans alone)
    if (E.in_step) {
      const field_type fte = ft == D_stuff ? E_stuff : H_stuff;
      E.fused_eh[fte].clear();
      E.free_phase(E.phase(PH_EH, fte));
    }
    for (int i = 0; i < num_chunks; i++)
      if (chunks[i]->is_mine())
        if (chunks[i]->step_db(ft)) {
          chunk_connections_valid = false;
          assert(changed_materials);
        }
  });
}

// Decide whether fields_chunk::update_eh(E or H) of this chunk may be folded into the D (or B)
// pass, i.e. whether E = chi1inv * D is a purely local, linear, diagonal operation here
// (reference src/update_eh.cpp:66-195: no f_minus_p, no off-diagonal chi1inv, no chi2/chi3,
// every array it needs already allocated), and describe the fused update per component.
static void plan_eh_fusion_impl(fields_chunk *fc, field_type ft, int cmp, Recorder::Group &grp,
                                Engine *E, bool doing_solve_cw) {
  const field_type fte = ft == D_stuff ? E_stuff : H_stuff;
  structure_chunk *s = fc->s;
  const grid_volume &gv = fc->gv;
#define WHY(msg) do { if (E->verbose) fprintf(stderr, "meep_b200: chunk %d ft %d: no E/H fusion: %s\n", fc->chunk_idx, (int)ft, msg); return; } while (0)
  if (doing_solve_cw) return;
  if (!E->connections_valid) WHY("chunk connections (and metal lists) not rebuilt yet");
  if (cmp > 0 && !E->fused_eh[fte].count(fc)) return; // follow the decision made for cmp 0
  if (fc->gvs_eh[fte].size() > 1) return; // tiled update_eh keeps its own loop order
  for (const src_vol &sv : fc->get_sources(ft))
    if (sv.t()->is_integrated) WHY("integrated source"); // needs f_minus_p
  bool any = false;
  Recorder::Epilogue epi[3];
  int k = 0;
  FOR_FT_COMPONENTS(ft, dc) {
    if (!fc->f[dc][cmp]) continue; // (not a component of this grid, e.g. Dr/Dp in 3-D)
    if (k >= 3) WHY("more than three components");
    const component ec = field_type_component(fte, dc);
    const direction d_ec = component_direction(ec);
    const direction d_1 = cycle_direction(gv.dim, d_ec, 1), d_2 = cycle_direction(gv.dim, d_ec, 2);
    if (!fc->f[ec][cmp]) WHY("missing E/H component");
    for (polarization_state *p = fc->pol[fte]; p; p = p->next)
      if (p->s->needs_P(ec, cmp, fc->f)) WHY("polarisation");
    if (fc->f_minus_p[dc][cmp]) WHY("f_minus_p");
    if (s->chi1inv[ec][d_1] || s->chi1inv[ec][d_2] || s->chi2[ec] || s->chi3[ec]) WHY("offdiag/chi3");
    const direction dsigw = s->sigsize[d_ec] > 1 ? d_ec : NO_DIRECTION;
    if (fc->f[ec][cmp] == fc->f[dc][cmp]) {
      // H aliases B: update_eh is a no-op here, unless it is about to split them (first step)
      if (s->chi1inv[ec][d_ec] || dsigw != NO_DIRECTION) WHY("H about to split from B");
      ++k;
      continue;
    }
    if (dsigw != NO_DIRECTION && !fc->f_w[ec][cmp]) WHY("f_w not allocated yet"); // lazily allocated by update_eh
    Recorder::Epilogue &P = epi[k];
    P.e = E->dev(fc->f[ec][cmp]);
    P.u = E->dev(s->chi1inv[ec][d_ec]);
    P.fw = E->dev(fc->f_w[ec][cmp]);
    memset(&P.pmlw, 0, sizeof(P.pmlw));
    if (dsigw != NO_DIRECTION) {
      // array-index based lookup: k = 2*i_d + (yee offset of ec along d): evaluate at index 0
      const ivec is0 = gv.little_corner() + gv.iyee_shift(ec);
      P.pmlw = make_pml(gv, is0, dsigw, E->dev(s->sig[dsigw]), E->dev(s->kap[dsigw]), NULL);
    }
    // Points that zero_metal(ft) clears right after this update (src/step.cpp:243-245): they
    // must form whole boundary planes of the owned box, which the kernel can mask.
    {
      const realnum *base = fc->f[dc][cmp];
      const ivec is = gv.little_owned_corner0(dc), ie = gv.big_corner();
      const mb200_box_t ob = make_box(gv, is, ie);
      int lo[3], hi[3];
      int64_t rem = ob.idx0;
      const int64_t st[3] = {(int64_t)gv.stride(X), (int64_t)gv.stride(Y), (int64_t)gv.stride(Z)};
      for (int d = 0; d < 3; ++d) {
        lo[d] = (int)(rem / st[d]);
        rem -= (int64_t)lo[d] * st[d];
        hi[d] = lo[d] + ob.n[d] - 1;
      }
      int64_t cnt_lo[3] = {0, 0, 0}, cnt_hi[3] = {0, 0, 0}, total = 0;
      std::vector<int64_t> zidx;
      for (size_t z = 0; z < fc->num_zeroes[ft]; ++z) {
        const realnum *p = fc->zeroes[ft][z];
        if (p < base || p >= base + gv.ntot()) continue;
        zidx.push_back((int64_t)(p - base));
      }
      for (int64_t idx : zidx) {
        int64_t r = idx;
        int c3[3];
        for (int d = 0; d < 3; ++d) {
          c3[d] = (int)(r / st[d]);
          r -= (int64_t)c3[d] * st[d];
        }
        ++total;
        for (int d = 0; d < 3; ++d) {
          if (c3[d] == lo[d]) ++cnt_lo[d];
          if (c3[d] == hi[d]) ++cnt_hi[d];
        }
      }
      if (total) {
        int64_t plane[3];
        for (int d = 0; d < 3; ++d)
          plane[d] = (int64_t)ob.n[(d + 1) % 3] * ob.n[(d + 2) % 3];
        for (int d = 0; d < 3; ++d) {
          if (cnt_lo[d] == plane[d]) P.metal_lo[d] = lo[d];
          if (cnt_hi[d] == plane[d] && hi[d] != lo[d]) P.metal_hi[d] = hi[d];
        }
        // every zeroed point must lie on one of the flagged planes
        int64_t covered = 0;
        for (int64_t idx : zidx) {
          int64_t r = idx;
          bool on = false;
          for (int d = 0; d < 3; ++d) {
            const int c = (int)(r / st[d]);
            r -= (int64_t)c * st[d];
            if (c == P.metal_lo[d] || c == P.metal_hi[d]) on = true;
          }
          if (on) ++covered;
        }
        if (covered != total) WHY("metal points do not form whole boundary planes");
      }
    }
    any = true;
    ++k;
  }
  if (k != 3) WHY("fewer than three components");
  if (!any) WHY("nothing to fold"); // all aliased: plain fused curl
  // planes that hold source points: step_source must run between the D and the E update there
  int lo = 1, hi = 0;
  const ptrdiff_t sx = gv.stride(X);
  for (const src_vol &sv : fc->get_sources(ft))
    for (size_t j = 0; j < sv.num_points(); ++j) {
      const int ix = (int)(sv.index_at(j) / sx);
      if (lo > hi) lo = hi = ix;
      if (ix < lo) lo = ix;
      if (ix > hi) hi = ix;
    }
  if (lo <= hi && (hi - lo + 1) * 2 > gv.nx() + 1) WHY("sources everywhere"); // not worth it
  grp.fuse_eh = true;
  grp.slab_lo = lo;
  grp.slab_hi = hi;
  for (int c = 0; c < 3; ++c)
    grp.epi[c] = epi[c];
  if (cmp == 0)
    E->fused_eh[fte][fc] = std::make_pair(lo, hi);
  else if (!E->fused_eh[fte].count(fc))
    meep::abort("meep_b200: internal error: real and imaginary parts disagree on E/H fusion");
}

bool fields_chunk::step_db(field_type ft) {
  Engine *E = Engine::current();
  if (!E || !E->recording()) meep::abort("meep_b200: fields_chunk::step_db outside a phase");
  Recorder &R = E->rec();
  bool allocated_u = false;
  const size_t nbytes = gv.ntot() * sizeof(realnum);

  // cylindrical: the helper array of src/step_db.cpp:93-116 is device scratch, one per real /
  // imaginary part (the jobs of both parts run in the same launch)
  void *rderiv_int[2] = {nullptr, nullptr};
  const bool use_bfast = bfast_scaled_k[0] || bfast_scaled_k[1] || bfast_scaled_k[2];

  for (const auto &sub_gv : gvs_tiled) {
    DOCMP {
      Recorder::Group grp;
      grp.first = (int)R.curl.size();
      grp.fc = this;
      grp.cmp = cmp;
      FOR_FT_COMPONENTS(ft, cc) {
        if (f[cc][cmp]) {
          const component c_p = plus_component[cc], c_m = minus_component[cc];
          const direction d_deriv_p = plus_deriv_direction[cc];
          const direction d_deriv_m = minus_deriv_direction[cc];
          const direction d_c = component_direction(cc);
          const bool have_p = have_plus_deriv[cc];
          const bool have_m = have_minus_deriv[cc];
          const direction dsig0 = cycle_direction(gv.dim, d_c, 1);
          const direction dsig = s->sigsize[dsig0] > 1 ? dsig0 : NO_DIRECTION;
          const direction dsigu0 = cycle_direction(gv.dim, d_c, 2);
          const direction dsigu = s->sigsize[dsigu0] > 1 ? dsigu0 : NO_DIRECTION;
          ptrdiff_t stride_p = have_p ? gv.stride(d_deriv_p) : 0;
          ptrdiff_t stride_m = have_m ? gv.stride(d_deriv_m) : 0;
          realnum *f_p = have_p ? f[c_p][cmp] : NULL;
          realnum *f_m = have_m ? f[c_m][cmp] : NULL;
          realnum *the_f = f[cc][cmp];

          // lazy allocation, mirrored on the device (src/step_db.cpp:67-75): the host arrays
          // must exist because the reference's connection tables and accessors use them
          if (dsig != NO_DIRECTION && s->conductivity[cc][d_c] && !f_cond[cc][cmp]) {
            f_cond[cc][cmp] = new_zeroed_lazily(gv.ntot());
            E->ensure_from(f_cond[cc][cmp], nbytes, NULL);
          }
          if (dsigu != NO_DIRECTION && !f_u[cc][cmp]) {
            // (the device twin is initialised from the device copy of the_f; the host array is
            //  only an address until the next download fills it)
            f_u[cc][cmp] = new_zeroed_lazily(gv.ntot());
            E->ensure_from(f_u[cc][cmp], nbytes, the_f);
            allocated_u = true;
          }
          if (use_bfast && !f_bfast[cc][cmp]) {
            f_bfast[cc][cmp] = new_zeroed_lazily(gv.ntot());
            E->ensure_from(f_bfast[cc][cmp], nbytes, NULL);
          }

          if (ft == D_stuff) { // strides are opposite sign for H curl
            stride_p = -stride_p;
            stride_m = -stride_m;
          }

          void *g1_dev = E->dev(f_p), *g2_dev = E->dev(f_m);
          if (gv.dim == Dcyl) switch (d_c) { // (src/step_db.cpp:86-122)
              case meep::R:
                g1_dev = nullptr; // im/r Fz term will be handled separately
                break;
              case P: break; // curl works normally for phi component
              case Z: {
                g2_dev = nullptr; // im/r Fr term will be handled separately
                /* the z component needs 1/r d(r Fp)/dr: as in the reference, step_curl is given
                   the running sum over r of that quantity instead of Fp (see its comment at
                   src/step_db.cpp:93-103) */
                if (!rderiv_int[cmp]) {
                  rderiv_int[cmp] = E->aux_alloc(nbytes);
                  mb200_cylint_job_t CJ;
                  memset(&CJ, 0, sizeof(CJ));
                  CJ.out = rderiv_int[cmp];
                  CJ.fp = E->dev(f_p);
                  CJ.nr = gv.nr();
                  CJ.sr = gv.nz() + 1;
                  const realnum ir0 = gv.origin_r() * gv.a + 0.5 * gv.iyee_shift(c_p).in_direction(meep::R);
                  CJ.ir0 = ir0;
                  R.cylint.push_back(CJ);
                }
                g1_dev = rderiv_int[cmp];
                break;
              }
              default: meep::abort("bug - non-cylindrical field component in Dcyl");
            }

          const ivec is = sub_gv.little_owned_corner0(cc), ie = sub_gv.big_corner();
          mb200_curl_job_t J;
          memset(&J, 0, sizeof(J));
          J.box = make_box(gv, is, ie);
          J.f = E->dev(the_f);
          J.g1 = g1_dev;
          J.g2 = g2_dev;
          J.s1 = stride_p;
          J.s2 = stride_m;
          J.dtdx = (realnum)Courant;
          J.dt = (realnum)dt;
          if (!J.g1) { // swap g1 and g2 (src/step_generic.cpp:72-76)
            std::swap(J.g1, J.g2);
            std::swap(J.s1, J.s2);
            J.dtdx = -J.dtdx;
          }
          if (dsig != NO_DIRECTION)
            J.pml = make_pml(gv, is, dsig, E->dev(s->sig[dsig]), E->dev(s->kap[dsig]),
                             E->dev(s->siginv[dsig]));
          if (dsigu != NO_DIRECTION)
            J.pmlu = make_pml(gv, is, dsigu, E->dev(s->sig[dsigu]), E->dev(s->kap[dsigu]),
                              E->dev(s->siginv[dsigu]));
          J.fu = E->dev(f_u[cc][cmp]);
          J.cnd = E->dev(s->conductivity[cc][d_c]);
          J.cndinv = E->dev(s->condinv[cc][d_c]);
          J.fcnd = E->dev(f_cond[cc][cmp]);
          if (!J.g1) continue; // no curl term at all (step_curl returns at once: step_generic.cpp:72-77)
          if (J.box.n[0] <= 0 || J.box.n[1] <= 0 || J.box.n[2] <= 0) continue;
          R.curl.push_back(J);

          if (use_bfast) { // STEP_BFAST right after STEP_CURL of the same component (lines 129-143)
            realnum k1 = have_m ? bfast_scaled_k[component_index(c_m)] : 0; // puts k1 in direction of g2
            realnum k2 = have_p ? bfast_scaled_k[component_index(c_p)] : 0; // puts k2 in direction of g1
            if (ft == D_stuff) {
              k1 = -k1;
              k2 = -k2;
            }
            mb200_bfast_job_t BJ;
            memset(&BJ, 0, sizeof(BJ));
            BJ.box = J.box;
            BJ.f = J.f;
            BJ.g1 = g1_dev;
            BJ.g2 = g2_dev;
            BJ.s1 = stride_p;
            BJ.s2 = stride_m;
            BJ.k1 = k1;
            BJ.k2 = k2;
            if (!BJ.g1) { // swap g1 and g2, and k1/k2 with them (src/step_generic.cpp:342-346)
              std::swap(BJ.g1, BJ.g2);
              std::swap(BJ.s1, BJ.s2);
              std::swap(BJ.k1, BJ.k2);
            }
            BJ.pml = J.pml;
            BJ.pmlu = J.pmlu;
            BJ.fu = J.fu;
            BJ.cnd = J.cnd;
            BJ.cndinv = J.cndinv;
            BJ.fcnd = J.fcnd;
            BJ.F = E->dev(f_bfast[cc][cmp]);
            R.bfast.push_back(BJ);
          }
        }
      }
      grp.count = (int)R.curl.size() - grp.first;
      if (gvs_tiled.size() == 1 && grp.count > 0) {
        if (grp.count == 3 && gv.dim == D3 && E->fuse && E->in_step && !use_bfast)
          plan_eh_fusion_impl(this, ft, cmp, grp, E, doing_solve_cw);
        R.curl_groups.push_back(grp);
      }
    }
  }
  /* In 2d with beta != 0, add beta terms (see the reference's comment at
     src/step_db.cpp:148-160); one mb200_beta_job_t per STEP_BETA call (lines 161-175). */
  if (gv.dim == D2 && beta != 0) DOCMP for (direction d_c = X; d_c <= Y; d_c = direction(d_c + 1)) {
      component cc = direction_component(first_field_component(ft), d_c);
      component c_g = direction_component(ft == D_stuff ? Hx : Ex, d_c == X ? Y : X);
      realnum *the_f = f[cc][cmp];
      const realnum *g = f[c_g][1 - cmp] ? f[c_g][1 - cmp] : f[c_g][cmp];
      if (!the_f || !g) continue; // step_beta returns immediately without g (src/step_generic.cpp:259)
      const direction dsig0 = cycle_direction(gv.dim, d_c, 1);
      const direction dsig = s->sigsize[dsig0] > 1 ? dsig0 : NO_DIRECTION;
      const direction dsigu0 = cycle_direction(gv.dim, d_c, 2);
      const direction dsigu = s->sigsize[dsigu0] > 1 ? dsigu0 : NO_DIRECTION;
      const realnum betadt = 2 * pi * beta * dt * (d_c == X ? +1 : -1) *
                             (f[c_g][1 - cmp] ? (ft == D_stuff ? -1 : +1) * (2 * cmp - 1) : 1);
      const ivec is = gv.little_owned_corner0(cc), ie = gv.big_corner();
      mb200_beta_job_t J;
      memset(&J, 0, sizeof(J));
      J.box = make_box(gv, is, ie);
      J.f = E->dev(the_f);
      J.g = E->dev(g);
      J.betadt = betadt;
      if (dsig != NO_DIRECTION) J.pml = make_pml(gv, is, dsig, NULL, NULL, E->dev(s->siginv[dsig]));
      if (dsigu != NO_DIRECTION) J.pmlu = make_pml(gv, is, dsigu, NULL, NULL, E->dev(s->siginv[dsigu]));
      J.fu = E->dev(f_u[cc][cmp]);
      J.cndinv = E->dev(s->condinv[cc][d_c]);
      J.fcnd = E->dev(f_cond[cc][cmp]);
      if (J.box.n[0] > 0 && J.box.n[1] > 0 && J.box.n[2] > 0) R.beta.push_back(J);
    }

  // in cylindrical coordinates, we now have to add the i*m/r terms (src/step_db.cpp:177-280):
  // the eight loops there are the eight loops of step_beta with a factor the_m / r
  if (gv.dim == Dcyl && m != 0) DOCMP FOR_FT_COMPONENTS(ft, cc) {
      const direction d_c = component_direction(cc);
      if (f[cc][cmp] && (d_c == meep::R || d_c == Z)) {
        const component c_g = d_c == meep::R ? plus_component[cc] : minus_component[cc];
        const realnum *g = f[c_g][1 - cmp];
        if (!g) meep::abort("meep_b200: cylindrical fields with m != 0 must be complex");
        const direction dsig = cycle_direction(gv.dim, d_c, 1);
        const direction dsigu = cycle_direction(gv.dim, d_c, 2);
        const ivec is = gv.little_owned_corner0(cc), ie = gv.big_corner();
        const realnum the_m =
            2 * m * (1 - 2 * cmp) * (1 - 2 * (ft == B_stuff)) * (1 - 2 * (d_c == meep::R)) * Courant;
        mb200_beta_job_t J;
        memset(&J, 0, sizeof(J));
        J.box = make_box(gv, is, ie);
        J.f = E->dev(f[cc][cmp]);
        J.g = E->dev(g);
        J.betadt = the_m;
        J.cyl = 1;
        J.r_is2 = is.yucky_val(1);
        if (s->sigsize[dsig] > 1) J.pml = make_pml(gv, is, dsig, NULL, NULL, E->dev(s->siginv[dsig]));
        if (s->sigsize[dsigu] > 1)
          J.pmlu = make_pml(gv, is, dsigu, NULL, NULL, E->dev(s->siginv[dsigu]));
        J.fu = E->dev(f_u[cc][cmp]);
        J.cndinv = E->dev(s->condinv[cc][d_c]);
        J.fcnd = E->dev(f_cond[cc][cmp]);
        if (J.pmlu.siginv && !J.fu) meep::abort("meep_b200: cylindrical m/r term: missing f_u");
        if (J.pml.siginv && J.cndinv && !J.fcnd) meep::abort("meep_b200: cylindrical m/r term: missing f_cond");
        if (J.box.n[0] > 0 && J.box.n[1] > 0 && J.box.n[2] > 0) R.beta.push_back(J);
      }
    }

  // deal with the r=0 boundary conditions for m=0 and m=1 (src/step_db.cpp:282-462)
  if (gv.dim == Dcyl && gv.origin_r() == 0.0) DOCMP {
      const int nz = gv.nz();
      auto zero_z = [&](realnum *array, int row) { // ZERO_Z(array + row * (nz + 1))
        if (!array) return;
        std::vector<uint64_t> zp(nz + 1);
        const uint64_t base = E->dev_addr(array + (size_t)row * (nz + 1));
        for (int k = 0; k <= nz; ++k)
          zp[k] = base + (uint64_t)k * sizeof(realnum);
        mb200_zero_job_t zj;
        zj.ptrs = (const uint64_t *)E->aux_upload(zp.data(), zp.size() * 8);
        zj.n = (int64_t)zp.size();
        R.cylzero.push_back(zj);
      };
      auto zero3 = [&](component c, int row) {
        zero_z(f[c][cmp], row);
        zero_z(f_cond[c][cmp], row);
        zero_z(f_u[c][cmp], row);
      };
      // the common tail of the two r = 0 loops (lines 300-321 and 350-371)
      auto origin_job = [&](component cc, direction d_c, int mode, const realnum *fp, const realnum *fm,
                            int sd, double c, double mult) {
        const direction dsig = cycle_direction(gv.dim, d_c, 1);
        const direction dsigu = cycle_direction(gv.dim, d_c, 2);
        const bool have_sig = s->sigsize[dsig] > 1, have_sigu = s->sigsize[dsigu] > 1;
        realnum *fu = have_sigu && f_u[cc][cmp] ? f[cc][cmp] : 0;
        realnum *the_f = fu ? f_u[cc][cmp] : f[cc][cmp];
        ivec is = gv.little_owned_corner(cc);
        ivec ie = gv.big_owned_corner(cc);
        ie.set_direction(meep::R, 0);
        mb200_cylr0_job_t J;
        memset(&J, 0, sizeof(J));
        J.box = make_box(gv, is, ie);
        J.f = E->dev(the_f);
        J.fu = E->dev(fu);
        J.fp = E->dev(fp);
        J.fm = fm ? E->dev(fm) : nullptr;
        J.sd = sd;
        J.c = c;
        J.mult = mult;
        J.dt = dt;
        J.mode = mode;
        J.fcnd = E->dev(f_cond[cc][cmp]);
        J.cnd = E->dev(s->conductivity[cc][d_c]);
        J.cndinv = E->dev(s->condinv[cc][d_c]);
        if (J.fcnd && (!J.cnd || !J.cndinv)) meep::abort("meep_b200: r=0 update: f_cond without conductivity");
        if (have_sig)
          J.pml = make_pml(gv, is, dsig, E->dev(s->sig[dsig]), E->dev(s->kap[dsig]), E->dev(s->siginv[dsig]));
        if (have_sigu)
          J.pmlu = make_pml(gv, is, dsigu, E->dev(s->sig[dsigu]), E->dev(s->kap[dsigu]),
                            E->dev(s->siginv[dsigu]));
        if (J.box.n[0] > 0 && J.box.n[1] > 0 && J.box.n[2] > 0) R.cylr0.push_back(J);
      };
      if (m == 0 && ft == D_stuff && f[Dz][cmp]) {
        // d(Dz)/dt = (1/r) * d(r*Hp)/dr
        origin_job(Dz, Z, 0, f[Hp][cmp], NULL, 0, Courant * 4, 0);
        zero3(Dp, 0);
      }
      else if (m == 0 && ft == B_stuff && f[Br][cmp]) { zero3(Br, 0); }
      else if (fabs(m) == 1) {
        // D_stuff: d(Dp)/dt = d(Hr)/dz - d(Hz)/dr
        // B_stuff: d(Br)/dt = d(Ep)/dz - i*m*Ez/r
        component cc = ft == D_stuff ? Dp : Br;
        if (!f[cc][cmp]) continue;
        const realnum *f_p = f[ft == D_stuff ? Hr : Ep][cmp];
        if (ft != D_stuff && !f[Ez][1 - cmp])
          meep::abort("meep_b200: cylindrical fields with m != 0 must be complex");
        const realnum *f_m = ft == D_stuff ? f[Hz][cmp] : (f[Ez][1 - cmp] + (nz + 1));
        const int sd = ft == D_stuff ? +1 : -1;
        const realnum f_m_mult = ft == D_stuff ? 2 : (1 - 2 * cmp) * m;
        origin_job(cc, component_direction(cc), 1, f_p, f_m, sd, sd * Courant, f_m_mult);
        if (ft == D_stuff) zero3(Dz, 0);
      }
      else if (m != 0) { // m != {0,+1,-1}
        // (see the reference's comments at src/step_db.cpp:389-405 and 433-439)
        int nrows = 1;
        if (zero_fields_near_cylorigin) {
          const double rmax = fabs(m) - int(gv.origin_r() * gv.a + 0.5);
          nrows = 0;
          for (int r = 0; r <= gv.nr() && r < rmax; r++)
            nrows++;
        }
        for (int r = 0; r < nrows; r++) {
          if (ft == D_stuff) {
            zero3(Dr, r);
            zero3(Dp, r);
            zero3(Dz, r);
          }
          else {
            zero3(Br, r);
            zero3(Bp, r);
            zero3(Bz, r);
          }
        }
      }
    }

  return allocated_u;
}

} // namespace meep
