// dft_hot.cpp — B200 replacements for the hot half of the reference's src/dft.cpp:
// fields::update_dfts (250-255), fields_chunk::update_dfts (257-264), dft_chunk::update_dft
// (266-308) and dft_flux::flux (542-556).  The set-up half of dft.cpp (add_dft*, dft_chunk
// constructor, save/load, ...) stays the reference's; the interposed definitions below
// override these symbols only.
#include <string.h>

#include "engine.hpp"
#include "loop_desc.hpp"
#include "meep_internals.hpp"

using namespace std;
using namespace meep_b200;

namespace meep {

void fields::update_dfts() {
  am_now_working_on(FourierTransforming);
  bool any = false;
  for (int i = 0; i < num_chunks && !any; i++)
    if (chunks[i]->is_mine() && chunks[i]->dft_chunks) any = true;
  if (any) {
    Engine &E = Engine::get(this);
    Scope scope(E, this);
    run_phase(E, this, PH_DFT, 0, true, [&]() {
      for (int i = 0; i < num_chunks; i++)
        if (chunks[i]->is_mine()) chunks[i]->update_dfts(time(), time() - 0.5 * dt, t);
    });
  }
  finished_working();
}

// Records one job per dft_chunk (all decimation factors; Engine::run applies the
// `current_step % decimation == 0` gate of src/dft.cpp:260 at launch time).
void fields_chunk::update_dfts(double timeE, double timeH, int current_step) {
  (void)timeE;
  (void)timeH;
  (void)current_step;
  if (doing_solve_cw) return;
  Engine *E = Engine::current();
  if (!E || !E->recording()) meep::abort("meep_b200: fields_chunk::update_dfts outside a phase");
  Recorder &R = E->rec();
  for (dft_chunk *cur = dft_chunks; cur; cur = cur->next_in_chunk) {
    if (!f[cur->c][0]) continue;
    const int dec = cur->get_decimation_factor();
    mb200_dft_job_t J;
    memset(&J, 0, sizeof(J));
    J.box = make_box(gv, cur->is, cur->ie);
    J.f_re = E->dev(f[cur->c][0]);
    J.f_im = f[cur->c][1] ? E->dev(f[cur->c][1]) : NULL;
    J.avg1 = cur->avg1;
    J.avg2 = cur->avg2;
    for (int k = 0; k < 3; ++k) {
      const direction dk = gv.yucky_direction(k);
      J.wgt_s0[k] = cur->s0.in_direction(dk);
      J.wgt_s1[k] = cur->s1.in_direction(dk);
      J.wgt_e0[k] = cur->e0.in_direction(dk);
      J.wgt_e1[k] = cur->e1.in_direction(dk);
    }
    J.dV0 = cur->dV0;
    J.dV1 = cur->dV1;
    J.use_weights = cur->include_dV_and_interp_weights ? 1 : 0;
    J.sqrt_weights = cur->sqrt_dV_and_interp_weights ? 1 : 0;
    J.dft = E->dev(cur->dft);
    J.nomega = (int32_t)cur->omega.size();
    int slot = 0;
    for (dft_chunk *prev : R.dft_chunks[dec])
      slot += (int)prev->omega.size();
    J.phase_slot = slot;
    R.dft[dec].push_back(J);
    R.dft_chunks[dec].push_back(cur);
  }
}

void dft_chunk::update_dft(double) {
  meep::abort("meep_b200: dft_chunk::update_dft: this build has no CPU time-stepping path "
              "(DFT accumulation runs inside fields::step on the device)");
}

// Sum over the E/H dft_chunk lists walked in lock-step, on the device (double accumulation).
double *dft_flux::flux() {
  const size_t Nfreq = freq.size();
  double *F = new double[Nfreq];
  for (size_t i = 0; i < Nfreq; ++i)
    F[i] = 0;
  Engine *E = nullptr;
  for (dft_chunk *curE = this->E; curE && !E; curE = curE->next_in_dft)
    E = Engine::owner_of(curE->dft);
  if (E && E->state == Engine::DEVICE_NEWER) {
    std::vector<mb200_flux_job_t> jobs;
    void *d_out = nullptr;
    check(mb200_malloc(E->ctx, Nfreq * sizeof(double), &d_out), "malloc(flux)");
    check(mb200_memset(E->ctx, d_out, 0, Nfreq * sizeof(double)), "memset(flux)");
    for (dft_chunk *curE = this->E, *curH = H; curE && curH;
         curE = curE->next_in_dft, curH = curH->next_in_dft) {
      mb200_flux_job_t J;
      memset(&J, 0, sizeof(J));
      J.e = E->dev(curE->dft);
      J.h = E->dev(curH->dft);
      J.npts = (int64_t)curE->N;
      J.nomega = (int32_t)Nfreq;
      J.out = (double *)d_out;
      if (J.npts > 0) jobs.push_back(J);
    }
    if (!jobs.empty())
      check(mb200_dft_flux(E->ctx, E->dtype, jobs.data(), (int)jobs.size()), "mb200_dft_flux");
    check(mb200_d2h(E->ctx, F, d_out, Nfreq * sizeof(double)), "d2h(flux)");
    E->stats.d2h_bytes += Nfreq * sizeof(double);
    mb200_free(E->ctx, d_out);
  }
  else {
    // the DFT arrays are current on the host (no step taken since they were last synchronised,
    // or they were loaded/scaled by host code): plain host sum, as src/dft.cpp:547-550
    for (dft_chunk *curE = this->E, *curH = H; curE && curH;
         curE = curE->next_in_dft, curH = curH->next_in_dft)
      for (size_t k = 0; k < curE->N; ++k)
        for (size_t i = 0; i < Nfreq; ++i)
          F[i] += real(curE->dft[k * Nfreq + i] * conj(curH->dft[k * Nfreq + i]));
  }
  double *Fsum = new double[Nfreq];
  sum_to_all(F, Fsum, int(Nfreq));
  delete[] F;
  return Fsum;
}

} // namespace meep
