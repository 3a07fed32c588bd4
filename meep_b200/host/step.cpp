// step.cpp — B200 replacement for the reference translation unit src/step.cpp.
//
// Defines exactly the symbols that TU defines (fields::step, step_boundaries,
// process_incoming_chunk_data, step_source, calc_sources, phase_material and the fields_chunk
// counterparts), compiled against the reference's UNMODIFIED meep.hpp.  The schedule of
// fields::step (reference src/step.cpp:35-139) is kept verbatim; every phase is executed on
// the device through the C ABI in include/meep_b200.h.  There is no CPU time-stepping path.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <vector>

#include "engine.hpp"
#include "loop_desc.hpp"
#include "meep_internals.hpp"

using namespace std;
using namespace meep_b200;

namespace meep {

// One half of a time step: the D/B family `db` is advanced from the curl of the other family,
// then the E/H family `eh` is derived from it (with its W and polarisation exchanges in between).
// fields::step() is two of these, B/H first; the order inside is the reference's
// (src/step.cpp:62-121) and is what every parity test checks.
namespace {
struct HalfStep {
  field_type db, eh, w, p;
  time_sink t_update_db, t_bnd_db, t_update_eh, t_bnd_w, t_bnd_p, t_bnd_eh;
};
const HalfStep kHalves[2] = {
    {B_stuff, H_stuff, WH_stuff, PH_stuff, FieldUpdateB, BoundarySteppingB, FieldUpdateH,
     BoundarySteppingWH, BoundarySteppingPH, BoundarySteppingH},
    {D_stuff, E_stuff, WE_stuff, PE_stuff, FieldUpdateD, BoundarySteppingD, FieldUpdateE,
     BoundarySteppingWE, BoundarySteppingPE, BoundarySteppingE}};
} // namespace

void fields::step() {
  Engine &E = Engine::get(this);

  // Stepping happens on unsynchronised fields: undo any number of pending
  // synchronize_magnetic_fields() calls now and redo one at the end (device-side, sync_magnetic.cpp)
  const int sync_depth = synchronized_magnetic_fields;
  if (sync_depth) {
    synchronized_magnetic_fields = 1;
    restore_magnetic_fields();
  }

  am_now_working_on(Stepping);

  // progress line (same text and cadence as the reference's)
  if (!t) {
    last_step_output_wall_time = wall_time();
    last_step_output_t = t;
  }
  if (verbosity > 0 && wall_time() > last_step_output_wall_time + MEEP_MIN_OUTPUT_TIME) {
    master_printf("on time step %d (time=%g), %g s/step\n", t, time(),
                  (wall_time() - last_step_output_wall_time) / (t - last_step_output_t));
    if (sync_depth) master_printf("  (doing expensive timestepping of synched fields)\n");
    last_step_output_wall_time = wall_time();
    last_step_output_t = t;
  }

  {
    // host-side material state that the device mirrors: conductivity inverses are refreshed by
    // reference code and re-uploaded when they changed
    E.materials_dirty = E.materials_dirty || changed_materials;
    E.cw_mode = false;
    for (int i = 0; i < num_chunks; i++) {
      const bool mine = chunks[i]->is_mine();
      if (mine && chunks[i]->s->condinv_stale) E.materials_dirty = true;
      if (mine && chunks[i]->doing_solve_cw) E.cw_mode = true;
      chunks[i]->s->update_condinv();
    }
    E.in_step = true;
    Scope scope(E, this);

    phase_material();

    // The reference refreshes the conductivity inverses AFTER phase_material (src/step.cpp:58-62):
    // mix_with() changes the conductivities (and may allocate them) and marks condinv stale.  The
    // loop above served the mirror set-up of the ordinary case; this one catches a phase-in step.
    {
      bool refreshed = false;
      for (int i = 0; i < num_chunks; i++) {
        if (chunks[i]->is_mine() && chunks[i]->s->condinv_stale) refreshed = true;
        chunks[i]->s->update_condinv();
      }
      if (refreshed) {
        E.scan(this); // condinv arrays may be new
        E.upload_materials();
        E.invalidate_plans();
      }
    }

    // Phase timers.  The reference brackets each phase with a host wall clock (timing_scope,
    // src/step.cpp:64-121); launches are asynchronous here, so with verbosity > 1 (or
    // MEEP_B200_TIMERS=1) each phase boundary is a CUDA event on the engine's stream instead and
    // the device times are credited to the same time_sinks when the step ends (see below).
    const bool dev_timers = E.device_timers || verbosity > 1;
    auto mark = [&](time_sink s) {
      if (dev_timers) check(mb200_mark(E.ctx, (int)s), "mb200_mark");
    };
    double t_trace = wall_time();
    auto trace = [&](const char *what) { // MEEP_B200_VERBOSE=1: where the host time of a (first) step goes
      if (!E.verbose) return;
      const double now = wall_time();
      if (now - t_trace > 0.05) fprintf(stderr, "meep_b200: step %d: %s took %.3f s on the host\n", t, what, now - t_trace);
      t_trace = now;
    };
    trace("phase_material / condinv");
    time_sink_to_duration_map discard; // host launch time of a phase is not phase time
    auto phase_clock = [&](time_sink s) {
      if (dev_timers) return timing_scope(&discard, s);
      return with_timing_scope(s);
    };
    for (int h = 0; h < 2; ++h) {
      const HalfStep &H = kHalves[h];
      const double t_half = time() + 0.5 * dt * h;
      calc_sources(t_half); // currents driving this family
      mark(H.t_update_db);
      {
        auto timer = phase_clock(H.t_update_db);
        step_db(H.db);
      }
      trace("step_db");
      mark(Stepping);
      step_source(H.db);
      mark(H.t_bnd_db);
      {
        auto timer = phase_clock(H.t_bnd_db);
        step_boundaries(H.db);
      }
      trace("step_boundaries(D/B)");
      calc_sources(t_half + 0.5 * dt); // integrated sources enter the E/H update
      mark(H.t_update_eh);
      {
        auto timer = phase_clock(H.t_update_eh);
        update_eh(H.eh);
      }
      trace("update_eh");
      mark(H.t_bnd_w);
      {
        auto timer = phase_clock(H.t_bnd_w);
        step_boundaries(H.w);
      }
      mark(Stepping);
      update_pols(H.eh);
      mark(H.t_bnd_p);
      {
        auto timer = phase_clock(H.t_bnd_p);
        step_boundaries(H.p);
      }
      mark(H.t_bnd_eh);
      {
        auto timer = phase_clock(H.t_bnd_eh);
        step_boundaries(H.eh);
      }
      trace("step_boundaries(W, P, E/H) + update_pols");
      mark(Stepping);
      if (fluxes) {
        // legacy flux planes (flux_vol, src/meep.hpp:2322-2351) integrate host arrays through
        // loop_in_chunks: the interposed loop_in_chunks downloads just the planes they read
        E.force_reader_sync = true;
        E.forced_download_done = false;
        if (h == 0) fluxes->update_half();
        else fluxes->update();
        E.force_reader_sync = false;
      }
    }

    t += 1;
    mark(FourierTransforming);
    update_dfts();
    mark(Stepping);
    if (dev_timers) {
      // One synchronisation per step.  The per-phase sinks (FieldUpdateB ... BoundarySteppingE)
      // receive their device time; the time the host spent waiting for the device is what the
      // reference's exclusive sinks measure, so it is moved from Stepping (where the wait above
      // is clocked) to Boundaries / FourierTransforming for the phases that belong there.
      int tags[64], n = 0;
      double ms[64];
      const double w0 = wall_time();
      check(mb200_marks_collect(E.ctx, tags, ms, 64, &n), "mb200_marks_collect");
      (void)w0;
      for (int k = 0; k < n; ++k) {
        const time_sink sink = (time_sink)tags[k];
        const double sec = ms[k] * 1e-3;
        if (sink == Stepping) continue; // already clocked by the enclosing am_now_working_on(Stepping)
        if (sink == FourierTransforming) {
          times_spent[FourierTransforming] += sec;
          times_spent[Stepping] -= sec;
          continue;
        }
        times_spent[sink] += sec;
        if (sink >= BoundarySteppingB && sink <= BoundarySteppingE) {
          times_spent[Boundaries] += sec;
          times_spent[Stepping] -= sec;
        }
      }
    }
    finished_working();

    changed_materials = false; // any material changes were handled in connect_chunks()

    // NaN/Inf check (reference src/step.cpp:137-138 reads D_EnergyDensity at the cell centre on
    // the host every step): a device-side probe of the same grid points, read back every
    // MEEP_B200_NAN_CHECK_EVERY steps so that stepping stays asynchronous.
    E.stats.steps++;
    E.check_probe(this, false);
  }
  E.in_step = false;

  if (E.eager) E.sync_host();

  if (sync_depth) {
    synchronize_magnetic_fields();
    synchronized_magnetic_fields = sync_depth;
  }
}

void fields::phase_material() {
  bool changed = false;
  if (is_phasing()) {
    Engine &E = Engine::get(this);
    Scope scope(E, this);
    for (int i = 0; i < num_chunks; i++)
      if (chunks[i]->is_mine()) {
        chunks[i]->phase_material(phasein_time);
        changed = changed || chunks[i]->new_s;
      }
    phasein_time--;
    am_now_working_on(MpiAllTime);
    bool changed_mpi = or_to_all(changed);
    finished_working();
    if (changed_mpi) {
      // mix_with() rewrote the material arrays on the host (possibly re-allocating them)
      E.scan(this);
      E.upload_materials();
      E.invalidate_plans();
      calc_sources(time() + 0.5 * dt); // for integrated H sources
      update_eh(H_stuff);              // ensure H = 1/mu * B
      step_boundaries(H_stuff);
      calc_sources(time() + dt); // for integrated E sources
      update_eh(E_stuff);        // ensure E = 1/eps * D
      step_boundaries(E_stuff);
    }
  }
}

void fields_chunk::phase_material(int phasein_time) {
  if (new_s && phasein_time > 0) {
    changing_structure();
    s->mix_with(new_s, 1.0 / phasein_time);
  }
}

// Host-side scatter of one received comm block (reference src/step.cpp:172-223).  On the device
// path chunk pairs living on the same GPU are exchanged by the halo plan (step_boundaries
// below) and never pass through comm_blocks, so this entry point only remains for API
// compatibility; there is no multi-process transport in this build.
void fields::process_incoming_chunk_data(field_type, const chunk_pair &) {
  meep::abort("meep_b200: process_incoming_chunk_data: comm blocks are unpacked on the device by "
              "fields::step_boundaries; there is no host-side scatter in this build");
}

void fields::step_boundaries(field_type ft) {
  Engine &E = Engine::get(this);
  Scope scope(E, this);

  // connect_chunks() re-connects if !chunk_connections_valid — and the flag it tests is first
  // and-ed over all processes when materials changed (sync_chunk_connections), so another rank's
  // lazy allocation can force a re-connection here that the local flag did not announce: detect
  // the re-connection itself (connect_the_chunks in connect.cpp bumps connect_epoch).
  // Comparing against the epoch of the last invalidation also covers a re-connection made
  // outside this function (user code calling connect_chunks()).
  const bool refresh = E.refresh_local; // Engine::refresh_deferred_halos: same-process pairs only
  connect_chunks();
  const bool was_valid = E.connect_epoch == E.plans_epoch;
  if (!was_valid) {
    E.invalidate_plans();
    E.plans_epoch = E.connect_epoch;
  }

  // Peer-memory links follow the chunk connections (connect.cpp bumps connect_epoch); all
  // processes reach the first in-step exchange after a (collective) re-connection together.
  if (E.in_step && E.p2p && count_processors() > 1 && E.links_epoch != E.connect_epoch) {
    std::map<std::pair<int, int>, std::pair<size_t, size_t> > counts;
    FOR_FIELD_TYPES(ft2) {
      for (int j = 0; j < num_chunks; j++)
        for (int i = 0; i < num_chunks; i++) {
          const bool j_mine = chunks[j]->is_mine(), i_mine = chunks[i]->is_mine();
          if (j_mine == i_mine) continue;
          const size_t tot = comm_size_tot(ft2, chunk_pair{j, i});
          if (!tot) continue;
          std::pair<size_t, size_t> &c =
              counts[std::make_pair((int)ft2, j_mine ? chunks[i]->n_proc() : chunks[j]->n_proc())];
          (j_mine ? c.first : c.second) += tot;
        }
    }
    E.rebuild_links(counts);
  }
  const bool use_links = E.in_step && E.p2p && count_processors() > 1;

  // Exchange merging.  The not-owned D (B) values are read by nothing before the E (H) exchange
  // unless update_eh needs neighbouring D (B) points (off-diagonal chi1inv, chi2/chi3: the g1/g2
  // reads of src/step_generic.cpp:580-581,592-593).  When no chunk of this process does, the D
  // (B) connections are carried out together with the E (H) ones: half as many launches and —
  // across GPUs — half as many (latency-bound) transfers per time step, same final arrays.
  const field_type partner = ft == D_stuff ? E_stuff : (ft == B_stuff ? H_stuff : ft);
  if (!refresh && (!was_valid || changed_materials || !E.defer_known)) {
    // (re)decide, together with all other processes — these three conditions are the same on
    // every rank, as the reference itself requires for sync_chunk_connections()
    for (field_type fdb : {D_stuff, B_stuff}) {
      const field_type feh = fdb == D_stuff ? E_stuff : H_stuff;
      bool ok = E.merge_exchanges && !fluxes;
      for (int i = 0; i < num_chunks && ok; i++) {
        if (!chunks[i]->is_mine()) continue;
        const structure_chunk *sc = chunks[i]->s;
        FOR_FT_COMPONENTS(feh, ec) {
          const direction d_ec = component_direction(ec);
          const direction d_1 = cycle_direction(chunks[i]->gv.dim, d_ec, 1);
          const direction d_2 = cycle_direction(chunks[i]->gv.dim, d_ec, 2);
          if (sc->chi1inv[ec][d_1] || sc->chi1inv[ec][d_2] || sc->chi2[ec] || sc->chi3[ec]) ok = false;
        }
      }
      if (count_processors() > 1) ok = and_to_all(ok);
      if (ok != E.defer_ok[fdb]) E.invalidate_plans();
      E.defer_ok[fdb] = ok;
    }
    E.defer_known = true;
  }
  const bool defer = E.in_step && partner != ft && E.defer_ok[ft];
  const bool take_partner = E.in_step && (ft == E_stuff || ft == H_stuff) &&
                            E.deferred_exchange[ft == E_stuff ? D_stuff : B_stuff];

  // Deferred local copies.  The D (B) connections that ride with the E (H) exchange only keep the
  // not-owned D (B) values identical to the reference's arrays — nothing on the device reads them
  // (that is what defer_ok established) unless a DFT monitor accumulates a D or B component.  For
  // pairs of chunks on the SAME device they are therefore not carried out every step at all: the
  // copies are made once, from the then-current owner values, before anybody can look — before a
  // download and before any stand-alone entry point (Engine::refresh_deferred_halos).  At 512^3
  // that is 41 % of all halo transfers.  Pairs that cross devices keep riding with E (H): a
  // refresh must not need the other process.
  const field_type ftdb = ft == E_stuff ? D_stuff : B_stuff;
  bool defer_local = false;
  if (take_partner && E.defer_local) {
    defer_local = true;
    for (int i = 0; i < num_chunks && defer_local; i++)
      if (chunks[i]->is_mine())
        for (dft_chunk *d = chunks[i]->dft_chunks; d; d = d->next_in_chunk)
          if (is_D(d->c) || is_B(d->c)) defer_local = false;
  }

  // The same for the polarisation halos (PE/PH: one value per Lorentzian pole and boundary point,
  // src/susceptibility.cpp:283-295): not-owned P values only feed the not-owned points of
  // f_minus_p, which a diagonal update_eh never reads — for the six-pole Au sphere they are six
  // times the E halo, in list form because the zero-block flags travel with them.
  const bool pol_ft = ft == PE_stuff || ft == PH_stuff;
  const bool defer_pol = pol_ft && E.in_step && !refresh && E.defer_local && E.defer_known &&
                         E.defer_ok[ft == PE_stuff ? D_stuff : B_stuff];

  am_now_working_on(Boundaries);
  run_phase(E, this, PH_BND, ft, E.in_step, [&]() {
    Recorder &R = E.rec();
    if (E.in_step && partner != ft) {
      // the decision is remade with this recording: re-record the partner phase against it
      E.deferred_exchange[ft] = defer;
      E.free_phase(E.phase(PH_BND, partner));
    }
    if (take_partner) E.local_deferred[ftdb] = defer_local;
    for (int i = 0; i < num_chunks; i++) {
      if (!chunks[i]->is_mine() || refresh) continue; // (a refresh only copies)
      // Do the metals first!  (fields_chunk::zero_metal, src/boundaries.cpp:310-313)
      const size_t nz = chunks[i]->num_zeroes[ft];
      if (nz) {
        std::vector<uint64_t> zp(nz);
        for (size_t k = 0; k < nz; ++k)
          zp[k] = E.dev_addr(chunks[i]->zeroes[ft][k]);
        mb200_zero_job_t zj;
        zj.ptrs = (const uint64_t *)E.aux_upload(zp.data(), nz * 8);
        zj.n = (int64_t)nz;
        R.zero.push_back(zj);
      }
    }
    // Every chunk pair (j -> i) in a fixed global order (all processes enumerate the pairs
    // identically, so the grouped sends and receives below match up).  Same-process pairs go
    // straight from connections_out[j] to connections_in[i] (positions match: both vectors are
    // built by the same traversal, src/boundaries.cpp:476-594).  For a pair that crosses
    // processes the comm block of src/step.cpp:256-267 (PHASE || NEGATE || COPY) is packed into
    // / unpacked from a contiguous device buffer and moved device-to-device.
    const size_t Rsz = sizeof(realnum);
    auto dev_list = [&](const std::vector<realnum *> &v, std::vector<uint64_t> &out) {
      for (realnum *p : v)
        out.push_back(E.dev_addr(p));
    };
    std::map<int, size_t> link_send_off, link_recv_off;
    std::vector<double> cost_halo, cost_pack, cost_unpack;
    std::vector<field_type> fts;
    if (take_partner) fts.push_back(ft == E_stuff ? D_stuff : B_stuff);
    if (!(E.in_step && partner != ft && defer)) fts.push_back(ft);
    for (field_type ftl : fts)
    for (int j = 0; j < num_chunks; j++)
      for (int i = 0; i < num_chunks; i++) {
        const field_type ft_outer = ft;
        const field_type ft = ftl; // (the body below is written for "the field type being connected")
        const chunk_pair pair{j, i};
        const size_t tot = comm_size_tot(ft, pair);
        if (!tot) continue;
        const bool j_mine = chunks[j]->is_mine(), i_mine = chunks[i]->is_mine();
        if (!j_mine && !i_mine) continue;
        if (refresh && !(j_mine && i_mine)) continue;
        if (defer_local && ft != ft_outer && j_mine && i_mine) continue; // copied by the next refresh instead
        if (defer_pol && j_mine && i_mine) continue;
        uint64_t block = 0; // device comm block for a cross-process pair
        if (j_mine != i_mine && use_links) {
          // the block is a slice of the arena in the RECEIVER's memory: packed straight into the
          // neighbour's HBM / unpacked from mine
          const int peer = j_mine ? chunks[i]->n_proc() : chunks[j]->n_proc();
          const int lk = E.find_link((int)ft, peer);
          if (lk < 0) meep::abort("meep_b200: no peer-memory link for a cross-process chunk pair");
          size_t &off = (j_mine ? link_send_off : link_recv_off)[lk];
          const P2PLink &L = E.links[lk];
          if (off + tot > (j_mine ? L.send_count : L.recv_count))
            meep::abort("meep_b200: peer-memory arena overflow (stale links)");
          block = (uint64_t)(uintptr_t)((char *)(j_mine ? L.theirs : L.mine) + kArenaHeader) + off * Rsz;
          off += tot;
          if (std::find(R.links.begin(), R.links.end(), lk) == R.links.end()) R.links.push_back(lk);
        }
        else if (j_mine != i_mine) {
          block = (uint64_t)(uintptr_t)E.aux_alloc(tot * Rsz);
          mb200_xfer_t x;
          x.peer = j_mine ? chunks[i]->n_proc() : chunks[j]->n_proc();
          x.reserved = 0;
          x.buf = (void *)(uintptr_t)block;
          x.count = (int64_t)tot;
          (j_mine ? R.sends : R.recvs).push_back(x);
        }
        size_t off = 0; // position inside the comm block
        for (connect_phase ip : all_connect_phases) {
          const comms_key key = {ft, ip, pair};
          const size_t n = get_comm_size(key);
          if (!n) continue;
          std::vector<uint64_t> src, dst;
          if (j_mine) {
            auto it = chunks[j]->connections_out.find(key);
            if (it == chunks[j]->connections_out.end() || it->second.size() != n)
              meep::abort("meep_b200: inconsistent outgoing chunk connection table");
            dev_list(it->second, src);
          }
          if (i_mine) {
            auto it = chunks[i]->connections_in.find(key);
            if (it == chunks[i]->connections_in.end() || it->second.size() != n)
              meep::abort("meep_b200: inconsistent incoming chunk connection table");
            dev_list(it->second, dst);
          }
          mb200_halo_job_t hj;
          memset(&hj, 0, sizeof(hj));
          if (j_mine && !i_mine) { // pack: plain copy into the outgoing block
            for (size_t k = 0; k < n; ++k)
              dst.push_back(block + (off + k) * Rsz);
            hj.n_copy = (int64_t)n;
          }
          else {
            if (!j_mine) // unpack: the block is the source
              for (size_t k = 0; k < n; ++k)
                src.push_back(block + (off + k) * Rsz);
            if (ip == CONNECT_PHASE) {
              const std::vector<std::complex<realnum> > &ph = chunks[i]->connection_phases.at(key);
              if (ph.size() * 2 != n) meep::abort("meep_b200: inconsistent connection phase table");
              hj.phase = E.aux_upload(ph.data(), ph.size() * sizeof(std::complex<realnum>));
              hj.n_phase = (int64_t)ph.size();
            }
            else if (ip == CONNECT_NEGATE)
              hj.n_negate = (int64_t)n;
            else
              hj.n_copy = (int64_t)n;
          }
          bool need_flags = false;
          if (i_mine && (ft == PE_stuff || ft == PH_stuff) && E.zero_skip)
            for (realnum *p : chunks[i]->connections_in.at(key))
              if (E.pzero_flag_addr(p)) need_flags = true;
          // NEGATE / COPY transfers between regular chunk faces: runs of constant stride
          // (40 bytes per run instead of 16 bytes of addresses per value)
          std::vector<mb200_halo_run_t> host_runs; // (kept for the cost estimate below)
          if (hj.n_phase == 0 && !need_flags && E.halo_runs && n >= 64) {
            std::vector<mb200_halo_run_t> &runs = host_runs;
            size_t k = 0;
            while (k < n) {
              mb200_halo_run_t r;
              r.src0 = src[k];
              r.dst0 = dst[k];
              r.dsrc = r.ddst = 0;
              r.negate = hj.n_negate > 0 ? 1 : 0; // (pack jobs copy; the sign is applied on arrival)
              size_t len = 1;
              if (k + 1 < n) {
                r.dsrc = (int64_t)(src[k + 1] - src[k]);
                r.ddst = (int64_t)(dst[k + 1] - dst[k]);
                len = 2;
                while (k + len < n && len < (size_t)1 << 30 &&
                       (int64_t)(src[k + len] - src[k + len - 1]) == r.dsrc &&
                       (int64_t)(dst[k + len] - dst[k + len - 1]) == r.ddst)
                  ++len;
              }
              r.n = (int32_t)len;
              runs.push_back(r);
              k += len;
            }
            if (E.verbose) {
              int32_t longest = 0;
              for (const mb200_halo_run_t &r : runs)
                longest = std::max(longest, r.n);
              fprintf(stderr, "meep_b200: halo pair (%d -> %d) ft %d: %zu values in %zu runs, longest %d\n", j, i, (int)ft, n,
                      runs.size(), (int)longest);
            }
            if (runs.size() * 8 <= n) { // worth it: at least 8 values per run on average
              hj.runs = (const mb200_halo_run_t *)E.aux_upload(runs.data(), runs.size() * sizeof(runs[0]));
              hj.nrun = (int64_t)runs.size();
              src.clear();
              dst.clear();
            }
          }
          hj.src = (const uint64_t *)E.aux_upload(src.data(), src.size() * 8);
          hj.dst = (const uint64_t *)E.aux_upload(dst.data(), dst.size() * 8);
          if (need_flags) {
            // polarisation values arriving from a neighbour: keep the zero-block flags exact
            const std::vector<realnum *> &in = chunks[i]->connections_in.at(key);
            std::vector<uint64_t> fl(in.size());
            bool any = false;
            for (size_t k = 0; k < in.size(); ++k) {
              fl[k] = E.pzero_flag_addr(in[k]);
              any = any || fl[k];
            }
            if (any) {
              // (elements without flags point at a scratch byte)
              uint64_t scratch = (uint64_t)(uintptr_t)E.aux_alloc(8);
              for (size_t k = 0; k < fl.size(); ++k)
                if (!fl[k]) fl[k] = scratch;
              hj.dst_flag = (const uint64_t *)E.aux_upload(fl.data(), fl.size() * 8);
            }
          }
          // estimated cost: a value that is not next to its predecessor in memory moves a whole
          // 32-byte sector on each side (z-normal faces: one element per row)
          double cost = 0;
          if (hj.nrun > 0)
            for (const mb200_halo_run_t &r : host_runs)
              cost += (double)r.n * ((std::abs(r.dsrc) > (int64_t)Rsz || std::abs(r.ddst) > (int64_t)Rsz) ? 4.0 : 1.0);
          else
            cost = 4.0 * (double)n;
          std::vector<mb200_halo_job_t> &tab = j_mine == i_mine ? R.halo : (j_mine ? R.pack : R.unpack);
          (j_mine == i_mine ? cost_halo : (j_mine ? cost_pack : cost_unpack)).push_back(cost);
          tab.push_back(hj);
          off += n;
        }
      }
    // Longest jobs first.  The transfers of a phase are independent (every destination is a
    // not-owned point with exactly one owner, no source is written in the phase), so the order of
    // the jobs in a launch is free; CTAs are dispatched in table order, and the slow jobs — the
    // strided z-normal faces — used to sit wherever the chunk pairs put them: at 512^3 the SMs were
    // idle for 28 % of the kernel (ncu: SM active cycles / elapsed) waiting for a few late CTAs.
    auto by_cost = [](std::vector<mb200_halo_job_t> &tab, const std::vector<double> &cost) {
      if (tab.size() != cost.size() || tab.size() < 2) return;
      std::vector<size_t> order(tab.size());
      for (size_t k = 0; k < order.size(); ++k)
        order[k] = k;
      std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return cost[a] > cost[b]; });
      std::vector<mb200_halo_job_t> sorted;
      for (size_t k : order)
        sorted.push_back(tab[k]);
      tab.swap(sorted);
    };
    if (E.halo_sort) {
      by_cost(R.halo, cost_halo);
      by_cost(R.pack, cost_pack);
      by_cost(R.unpack, cost_unpack);
    }
  });
  if (E.in_step && take_partner && E.local_deferred[ftdb]) E.halos_stale[ftdb] = true;
  if (defer_pol) E.halos_stale[ft] = true;
  finished_working();
}

void fields::step_source(field_type ft, bool including_integrated) {
  if (ft != D_stuff && ft != B_stuff) meep::abort("only step_source(D/B) is okay");
  Engine &E = Engine::get(this);
  Scope scope(E, this);
  bool cw = false;
  for (int i = 0; i < num_chunks; i++)
    if (chunks[i]->is_mine() && chunks[i]->doing_solve_cw) cw = true;
  run_phase(E, this, PH_SRC, ft, !including_integrated && !cw, [&]() {
    for (int i = 0; i < num_chunks; i++)
      if (chunks[i]->is_mine()) chunks[i]->step_source(ft, including_integrated);
  });
}

// Emits one source job per src_vol (reference src/step.cpp:295-318).
void fields_chunk::step_source(field_type ft, bool including_integrated) {
  Engine *E = Engine::current();
  if (!E || !E->recording()) meep::abort("meep_b200: fields_chunk::step_source outside a phase");
  if (doing_solve_cw && !including_integrated) return;
  Recorder &R = E->rec();
  for (const src_vol &sv : sources[ft]) {
    component c = direction_component(first_field_component(ft), component_direction(sv.c));
    const realnum *cndinv = s->condinv[c][component_direction(sv.c)];
    if ((including_integrated || !sv.t()->is_integrated) && f[c][0] &&
        ((ft == D_stuff && is_electric(sv.c)) || (ft == B_stuff && is_magnetic(sv.c)))) {
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
      J.f_re = E->dev(f[c][0]);
      J.f_im = is_real ? nullptr : E->dev(f[c][1]);
      J.cndinv = E->dev(cndinv);
      J.index = (const int64_t *)E->aux_upload(idx.data(), np * sizeof(int64_t));
      J.amp = (const double *)E->aux_upload(amp.data(), 2 * np * sizeof(double));
      J.npts = (int64_t)np;
      J.dt = dt;
      J.scalar_slot = (int32_t)R.src_times.size();
      J.mode = 0;
      R.src.push_back(J);
      R.src_times.push_back(sv.t());
    }
  }
}

void fields::calc_sources(double tim) {
  for (src_time *s = sources; s; s = s->next)
    s->update(tim, dt);
  for (int i = 0; i < num_chunks; i++)
    if (chunks[i]->is_mine()) chunks[i]->calc_sources(tim);
}

void fields_chunk::calc_sources(double time) {
  (void)time; // unused;
}

} // namespace meep
