// sync_magnetic.cpp — fields::synchronize_magnetic_fields / restore_magnetic_fields on the device
// (reference src/energy_and_flux.cpp:97-178).
//
// Energy, flux-in-box and output routines bring H/B to the time of E/D by stepping B and H half a
// step forward and averaging with the saved values, then put the saved values back.  The
// reference does the save / average / restore with host memcpy loops; here the copies are
// device-to-device, the half step is the ordinary recorded phases, the average is one launch
// (mb200_average_job_t), and nothing crosses PCIe: readers that follow (loop_in_chunks,
// get_field, ...) download on demand through the hooks in hooks.cpp.
#include "engine.hpp"
#include "meep_internals.hpp"

using namespace meep_b200;

namespace meep {

namespace {
// the arrays fields_chunk::backup_component saves for component c (lines 97-118)
template <typename F> void for_each_backed_up_array(fields_chunk *fc, component c, F fn) {
  for (int cmp = 0; cmp < 2; ++cmp) {
    if (c < NUM_FIELD_COMPONENTS && fc->f[c][cmp] &&
        // in mu=1 regions where H==B, don't bother to backup H
        !(is_magnetic(c) && fc->f[c][cmp] == fc->f[direction_component(Bx, component_direction(c))][cmp])) {
      fn(fc->f[c][cmp]);
      fn(fc->f_u[c][cmp]);
      fn(fc->f_w[c][cmp]);
      fn(fc->f_cond[c][cmp]);
      fn(fc->f_bfast[c][cmp]);
    }
  }
}
} // namespace

void fields::synchronize_magnetic_fields() {
  if (synchronized_magnetic_fields++) return; // already synched
  Engine &E = Engine::get(this);
  E.keep_on_device++;
  {
    Scope scope(E, this);
    for (int i = 0; i < num_chunks; i++)
      if (chunks[i]->is_mine()) {
        FOR_B_COMPONENTS(c) {
          for_each_backed_up_array(chunks[i], c, [&](realnum *a) { E.backup_array(a); });
        }
        FOR_MAGNETIC_COMPONENTS(c) {
          for_each_backed_up_array(chunks[i], c, [&](realnum *a) { E.backup_array(a); });
        }
      }
    am_now_working_on(Stepping);
    calc_sources(time()); // for B sources
    step_db(B_stuff);
    step_source(B_stuff);
    step_boundaries(B_stuff);
    calc_sources(time() + 0.5 * dt); // for integrated H sources
    update_eh(H_stuff);
    step_boundaries(H_stuff);
    finished_working();
    // average_with_backup (lines 139-147): only the field arrays themselves, f[c]
    std::vector<mb200_average_job_t> jobs;
    for (int i = 0; i < num_chunks; i++)
      if (chunks[i]->is_mine()) {
        auto add = [&](component c) {
          for (int cmp = 0; cmp < 2; ++cmp) {
            realnum *fc = chunks[i]->f[c][cmp];
            if (!fc || !E.has_backup(fc)) continue;
            // (B and H alias where mu = 1: average the shared array once)
            bool dup = false;
            for (const mb200_average_job_t &j : jobs)
              dup = dup || j.f == E.dev(fc);
            if (dup) continue;
            mb200_average_job_t j;
            j.f = E.dev(fc);
            j.backup = E.backups[fc];
            j.n = (int64_t)chunks[i]->gv.ntot();
            jobs.push_back(j);
          }
        };
        FOR_B_COMPONENTS(c) { add(c); }
        FOR_MAGNETIC_COMPONENTS(c) { add(c); }
      }
    if (!jobs.empty())
      check(mb200_average_with_backup(E.ctx, E.dtype, jobs.data(), (int)jobs.size()),
            "mb200_average_with_backup");
  }
  E.keep_on_device--;
}

void fields::restore_magnetic_fields() {
  if (!synchronized_magnetic_fields      // already restored
      || --synchronized_magnetic_fields) // not ready to restore yet
    return;
  Engine &E = Engine::get(this);
  E.keep_on_device++;
  {
    Scope scope(E, this);
    for (int i = 0; i < num_chunks; i++)
      if (chunks[i]->is_mine()) {
        FOR_B_COMPONENTS(c) {
          for_each_backed_up_array(chunks[i], c, [&](realnum *a) { E.restore_array(a); });
        }
        FOR_MAGNETIC_COMPONENTS(c) {
          for_each_backed_up_array(chunks[i], c, [&](realnum *a) { E.restore_array(a); });
        }
      }
  }
  E.keep_on_device--;
}

} // namespace meep
