// step_db.cpp — B200 replacement for the reference translation unit src/step_db.cpp.
//
// fields_chunk::step_db keeps the reference's per-component set-up (which arrays feed which
// curl term, PML directions, lazy allocation of f_u / f_cond: src/step_db.cpp:44-127) but
// instead of calling STEP_CURL it emits one mb200_curl_job_t per call; the Engine batches the
// jobs of all chunks into one launch (and fuses the three components of a 3-D chunk).
#include <assert.h>
#include <string.h>

#include "engine.hpp"
#include "loop_desc.hpp"
#include "meep_internals.hpp"

using namespace std;
using namespace meep_b200;

namespace meep {

void fields::step_db(field_type ft) {
  if (ft != B_stuff && ft != D_stuff) meep::abort("step_db only works with B/D");
  Engine &E = Engine::get(this);
  Scope scope(E, this);
  run_phase(E, this, PH_DB, ft, true, [&]() {
    for (int i = 0; i < num_chunks; i++)
      if (chunks[i]->is_mine())
        if (chunks[i]->step_db(ft)) {
          chunk_connections_valid = false;
          assert(changed_materials);
        }
  });
}

bool fields_chunk::step_db(field_type ft) {
  Engine *E = Engine::current();
  if (!E || !E->recording()) meep::abort("meep_b200: fields_chunk::step_db outside a phase");
  Recorder &R = E->rec();
  bool allocated_u = false;
  const size_t nbytes = gv.ntot() * sizeof(realnum);

  if (gv.dim == Dcyl)
    meep::abort("meep_b200: cylindrical coordinates are not supported on the device path yet");
  if (gv.dim == D2 && beta != 0)
    meep::abort("meep_b200: 2d beta != 0 is not supported on the device path yet");
  if (bfast_scaled_k[0] || bfast_scaled_k[1] || bfast_scaled_k[2])
    meep::abort("meep_b200: BFAST is not supported on the device path yet");

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
            f_cond[cc][cmp] = new realnum[gv.ntot()];
            memset(f_cond[cc][cmp], 0, nbytes);
            E->ensure_from(f_cond[cc][cmp], nbytes, NULL);
          }
          if (dsigu != NO_DIRECTION && !f_u[cc][cmp]) {
            f_u[cc][cmp] = new realnum[gv.ntot()];
            memcpy(f_u[cc][cmp], the_f, nbytes); // (host copy is refreshed on the next download)
            E->ensure_from(f_u[cc][cmp], nbytes, the_f);
            allocated_u = true;
          }

          if (ft == D_stuff) { // strides are opposite sign for H curl
            stride_p = -stride_p;
            stride_m = -stride_m;
          }

          const ivec is = sub_gv.little_owned_corner0(cc), ie = sub_gv.big_corner();
          mb200_curl_job_t J;
          memset(&J, 0, sizeof(J));
          J.box = make_box(gv, is, ie);
          J.f = E->dev(the_f);
          J.g1 = E->dev(f_p);
          J.g2 = E->dev(f_m);
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
          if (!J.g1) continue; // no curl term at all (cannot happen for allocated components)
          if (J.box.n[0] <= 0 || J.box.n[1] <= 0 || J.box.n[2] <= 0) continue;
          R.curl.push_back(J);
        }
      }
      grp.count = (int)R.curl.size() - grp.first;
      if (gvs_tiled.size() == 1 && grp.count > 0) R.curl_groups.push_back(grp);
    }
  }
  return allocated_u;
}

} // namespace meep
