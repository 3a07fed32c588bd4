// connect.cpp — fast replacements for fields::connect_the_chunks and fields::find_metals
// (reference src/boundaries.cpp:315-345 and 368-638).
//
// These build the tables that fields::step_boundaries consumes (connections_in/out,
// connection_phases, comm_sizes, zeroes).  The reference versions are
// O(not-owned points x chunks) with several hash-map lookups per point and a full-volume scan for
// metal points: 14 s at 256^3 / 27 chunks, ~1 min at 512^3, and they grow with the SQUARE of the
// chunk count in a multi-GPU run.  Here the same tables (same contents, same order) are produced
// with: the not-owned slabs walked in runs (rows) whose table entries are arithmetic progressions —
// the per-point work of the reference (locate_component_point, on_metal_boundary, owner search,
// polarisation matching) is done once per run, i.e. the walk is analytic in the box intersections
// of slabs, owned volumes and the cell; table vectors resolved once per (chunk pair, field type,
// phase class); a single pass (sizes are read off the finished vectors); metal points found by
// walking only the boundary planes of the owned volume.  Host code only — nothing here touches field values.
#include <algorithm>
#include <atomic>
#include <deque>
#include <stdlib.h>
#include <map>
#include <thread>
#include <vector>

#include "engine.hpp"
#include "meep_internals.hpp"

using namespace std;

namespace meep {

namespace {

bool phase_isclose(std::complex<double> thephase, double realphase) {
  return fabs(thephase.imag()) < 1e-13 && fabs(thephase.real() - realphase) < 1e-13;
}
connect_phase connect_phase_from_phase(std::complex<double> thephase) {
  return phase_isclose(thephase, 1.0)    ? CONNECT_COPY
         : phase_isclose(thephase, -1.0) ? CONNECT_NEGATE
                                         : CONNECT_PHASE;
}

} // namespace

// fields::find_metals (src/boundaries.cpp:315-345) lists, in LOOP_OVER_VOL_OWNED order, the owned
// points for which on_metal_boundary() holds.  That predicate only compares one coordinate at a
// time with at most three values per direction (src/boundaries.cpp:191-205), so the metal points
// of a chunk are whole coordinate planes: flag the loop indices that sit on such a plane once per
// direction, then walk rows — a row on a flagged plane is taken whole, any other row contributes
// its (at most three) flagged points.  Same list, same order, no per-point work.
void fields::find_metals() {
  const double t_start = wall_time();
  // metal coordinates per direction
  std::vector<int> metal_coord[5];
  LOOP_OVER_DIRECTIONS(gv.dim, d) {
    if (user_volume.has_boundary(High, d) && boundaries[High][d] == Metallic)
      metal_coord[d].push_back(user_volume.big_corner().in_direction(d));
    if (boundaries[Low][d] == Magnetic) metal_coord[d].push_back(user_volume.little_corner().in_direction(d) + 1);
    if (boundaries[Low][d] == Metallic) metal_coord[d].push_back(user_volume.little_corner().in_direction(d));
  }
  for (int i = 0; i < num_chunks; i++)
    if (chunks[i]->is_mine()) {
      const grid_volume vi = chunks[i]->gv;
      FOR_FIELD_TYPES(ft) {
        delete[] chunks[i]->zeroes[ft];
        std::vector<realnum *> found;
        DOCMP FOR_COMPONENTS(c) {
          if (type(c) != ft || !chunks[i]->f[c][cmp]) continue;
          realnum *base = chunks[i]->f[c][cmp];
          // geometry of LOOP_OVER_VOL_OWNED(vi, c, n)
          const ivec is = vi.little_owned_corner(c), ie = vi.big_corner();
          const ptrdiff_t is_[3] = {is.yucky_val(0), is.yucky_val(1), is.yucky_val(2)};
          const ptrdiff_t nn[3] = {(ie.yucky_val(0) - is_[0]) / 2 + 1, (ie.yucky_val(1) - is_[1]) / 2 + 1,
                                   (ie.yucky_val(2) - is_[2]) / 2 + 1};
          if (nn[0] <= 0 || nn[1] <= 0 || nn[2] <= 0) continue;
          const direction dd[3] = {vi.yucky_direction(0), vi.yucky_direction(1), vi.yucky_direction(2)};
          const ptrdiff_t ss[3] = {vi.stride(dd[0]), vi.stride(dd[1]), vi.stride(dd[2])};
          const ivec rel = is - vi.little_corner();
          const ptrdiff_t idx0 =
              rel.yucky_val(0) / 2 * ss[0] + rel.yucky_val(1) / 2 * ss[1] + rel.yucky_val(2) / 2 * ss[2];
          std::vector<char> flag[3];
          std::vector<ptrdiff_t> flagged3;
          for (int k = 0; k < 3; ++k) {
            flag[k].assign((size_t)nn[k], 0);
            if (!has_direction(gv.dim, dd[k])) continue; // (a loop of one iteration over a direction the grid lacks)
            for (int mc : metal_coord[dd[k]]) {
              const ptrdiff_t off = mc - is_[k];
              if (off >= 0 && off % 2 == 0 && off / 2 < nn[k]) flag[k][(size_t)(off / 2)] = 1;
            }
          }
          for (ptrdiff_t i3 = 0; i3 < nn[2]; ++i3)
            if (flag[2][(size_t)i3]) flagged3.push_back(i3);
          for (ptrdiff_t i1 = 0; i1 < nn[0]; ++i1)
            for (ptrdiff_t i2 = 0; i2 < nn[1]; ++i2) {
              realnum *row = base + idx0 + i1 * ss[0] + i2 * ss[1];
              if (flag[0][(size_t)i1] || flag[1][(size_t)i2])
                for (ptrdiff_t i3 = 0; i3 < nn[2]; ++i3)
                  found.push_back(row + i3 * ss[2]);
              else
                for (ptrdiff_t i3 : flagged3)
                  found.push_back(row + i3 * ss[2]);
            }
        }
        typedef realnum *realnum_ptr;
        chunks[i]->num_zeroes[ft] = found.size();
        chunks[i]->zeroes[ft] = new realnum_ptr[found.size()];
        std::copy(found.begin(), found.end(), chunks[i]->zeroes[ft]);
      }
    }
  if (getenv("MEEP_B200_VERBOSE") && atoi(getenv("MEEP_B200_VERBOSE")))
    master_printf("meep_b200: find_metals: %.3f s\n", wall_time() - t_start);
}

void fields::connect_the_chunks() {
  const double t_start = wall_time();
  if (meep_b200::Engine *e = meep_b200::Engine::find(this)) e->connect_epoch++;
  // (see the reference's comment at src/boundaries.cpp:369-377)
  std::vector<int> B_redundant(num_chunks * 2 * 5);
  for (int i = 0; i < num_chunks; ++i)
    FOR_H_AND_B(hc, bc) {
      B_redundant[5 * (num_chunks + i) + bc - Bx] = chunks[i]->f[hc][0] == chunks[i]->f[bc][0];
    }
  am_now_working_on(MpiAllTime);
  and_to_all(B_redundant.data() + 5 * num_chunks, B_redundant.data(), 5 * num_chunks);
  finished_working();

  bool needs_W_notowned[NUM_FIELD_COMPONENTS];
  FOR_COMPONENTS(c) { needs_W_notowned[c] = false; }
  FOR_E_AND_H(c) {
    for (int i = 0; i < num_chunks; i++)
      needs_W_notowned[c] = needs_W_notowned[c] || chunks[i]->needs_W_notowned(c);
  }
  am_now_working_on(MpiAllTime);
  FOR_E_AND_H(c) { needs_W_notowned[c] = or_to_all(needs_W_notowned[c]); }
  finished_working();

  comm_sizes.clear();

  // per (field type, phase class, other chunk j) caches for the chunk i being processed
  struct Slot {
    std::vector<realnum *> *in = nullptr, *out = nullptr;
    std::vector<std::complex<realnum> > *phases = nullptr;
    size_t count = 0; // realnums of this key when neither vector is ours to fill (never both NULL)
  };

  // A chunk of another process matters only if one of its not-owned points can be owned by one
  // of OUR chunks.  Without periodic boundaries or symmetries a not-owned point is at most one
  // pixel outside its chunk, so a bounding-box test against our chunks prunes the rest (in an
  // N-process run this keeps the work per process O(own surface) instead of O(total surface)).
  bool can_prune = S.multiplicity() == 1;
  LOOP_OVER_DIRECTIONS(gv.dim, d) {
    if (boundaries[High][d] == Periodic || boundaries[Low][d] == Periodic) can_prune = false;
  }
  auto touches_mine = [&](const grid_volume &vi) {
    const ivec lo = vi.little_corner() - one_ivec(vi.dim) * 2, hi = vi.big_corner() + one_ivec(vi.dim) * 2;
    for (int j = 0; j < num_chunks; j++) {
      if (!chunks[j]->is_mine()) continue;
      const ivec jl = chunks[j]->gv.little_corner(), jh = chunks[j]->gv.big_corner();
      bool overlap = true;
      LOOP_OVER_DIRECTIONS(vi.dim, d) {
        if (hi.in_direction(d) < jl.in_direction(d) || lo.in_direction(d) > jh.in_direction(d))
          overlap = false;
      }
      if (overlap) return true;
    }
    return false;
  };

  // One worker call per chunk i.  A worker only touches chunk i's own incoming tables; the
  // outgoing tables (which belong to the OTHER chunk j of each pair) are staged in `res` and
  // merged serially afterwards, so chunks can be processed by several host threads.
  struct Staged {
    size_t k; // slot index = ((field type, phase class), j)
    std::vector<realnum *> out;
  };
  struct ChunkResult {
    bool done = false;
    std::deque<Staged> staged;       // in first-touch order
    std::vector<size_t> touched_list; // slot indices in first-touch order
    std::vector<size_t> sizes;        // comm size per touched slot
  };
  std::vector<ChunkResult> results(num_chunks);
  const size_t nslots = (size_t)NUM_FIELD_TYPES * NUM_CONNECT_PHASE_TYPES * num_chunks;

  auto process_chunk = [&](int i) {
    const grid_volume &vi = chunks[i]->gv;
    const bool i_is_mine = chunks[i]->is_mine();
    if (!i_is_mine && can_prune && !touches_mine(vi)) return;
    ChunkResult &res = results[i];
    res.done = true;
    std::vector<Slot> slots(nslots);
    std::vector<char> touched(nslots, 0);
    std::vector<size_t> &touched_list = res.touched_list;
    auto slot = [&](field_type f, connect_phase ip, int j) -> Slot & {
      const size_t k = ((size_t)f * NUM_CONNECT_PHASE_TYPES + (size_t)ip) * num_chunks + j;
      if (!touched[k]) {
        touched[k] = 1;
        touched_list.push_back(k);
        const comms_key key = {f, ip, {j, i}};
        Slot &s = slots[k];
        if (i_is_mine) {
          s.in = &chunks[i]->connections_in[key];
          if (ip == CONNECT_PHASE) s.phases = &chunks[i]->connection_phases[key];
        }
        if (chunks[j]->is_mine()) {
          res.staged.push_back(Staged{k, {}});
          s.out = &res.staged.back().out;
        }
      }
      return slots[k];
    };
    int last_j = i;

    // The not-owned slabs of chunk i are walked in the reference's order (LOOP_OVER_VOL_NOTOWNED,
    // src/meep/vec.hpp:180-186: slab by slab, loops 1,2,3 nested) but in RUNS along the innermost
    // loop direction of extent > 1: the first point of a run goes through the reference's
    // per-point logic (locate_component_point, on_metal_boundary, owner search); how far the same
    // answer holds — same periodic image, same owner chunk, no metal plane — follows from the
    // chunk and cell bounds along that direction, and the last point of the run is re-checked.
    // All table entries of a run are arithmetic progressions.  With symmetries or cylindrical
    // coordinates runs have length 1 (the per-point walk).
    const bool runs_ok = S.multiplicity() == 1 && gv.dim != Dcyl;
    const int ncmp = 2 - is_real;
    struct PolPair { // a matching pair of polarisations of chunks i and j (src/boundaries.cpp:556-591)
      polarization_state *pi, *pj, *po;
      size_t ni, cni;
    };
    std::vector<PolPair> pol_pairs;

    auto head_of_run = [&](component corig, const ivec &here, component &c, ivec &there,
                           std::complex<double> &thephase, int &j) -> bool {
      c = corig;
      there = here;
      if (!locate_component_point(&c, &there, &thephase) || on_metal_boundary(there)) return false;
      // the chunk that owns `there` (ownership is exclusive): try the previous owner first
      j = -1;
      if (chunks[last_j]->gv.owns(there)) j = last_j;
      else
        for (int jj = 0; jj < num_chunks; jj++)
          if (chunks[jj]->gv.owns(there)) {
            j = jj;
            break;
          }
      if (j < 0) return false;
      last_j = j;
      return true;
    };

    FOR_COMPONENTS(corig) {
      if (!have_component(corig)) continue;
      ivec sis(vi.dim, 0), sie(vi.dim, 0);
      for (int ib = 0; vi.get_boundary_icorners(corig, ib, &sis, &sie); ib++) {
        // geometry of LOOP_OVER_IVECS(vi, sis, sie, n)
        const ptrdiff_t is_[3] = {sis.yucky_val(0), sis.yucky_val(1), sis.yucky_val(2)};
        const ptrdiff_t nn[3] = {(sie.yucky_val(0) - is_[0]) / 2 + 1, (sie.yucky_val(1) - is_[1]) / 2 + 1,
                                 (sie.yucky_val(2) - is_[2]) / 2 + 1};
        if (nn[0] <= 0 || nn[1] <= 0 || nn[2] <= 0) continue;
        const direction dd[3] = {vi.yucky_direction(0), vi.yucky_direction(1), vi.yucky_direction(2)};
        const ptrdiff_t ss[3] = {vi.stride(dd[0]), vi.stride(dd[1]), vi.stride(dd[2])};
        const ivec rel = sis - vi.little_corner();
        const ptrdiff_t idx0 =
            rel.yucky_val(0) / 2 * ss[0] + rel.yucky_val(1) / 2 * ss[1] + rel.yucky_val(2) / 2 * ss[2];
        const int rdim = nn[2] > 1 ? 2 : (nn[1] > 1 ? 1 : 0); // loops after rdim have one iteration
        const direction rd = dd[rdim];
        const ptrdiff_t lim0 = rdim > 0 ? nn[0] : 1, lim1 = rdim > 1 ? nn[1] : 1;
        for (ptrdiff_t o0 = 0; o0 < lim0; o0++)
          for (ptrdiff_t o1 = 0; o1 < lim1; o1++)
            for (ptrdiff_t k = 0; k < nn[rdim];) {
              ptrdiff_t ii[3] = {o0, o1, 0};
              ii[rdim] = k;
              ivec here(vi.dim);
              here.set_direction(dd[0], is_[0] + 2 * ii[0]);
              here.set_direction(dd[1], is_[1] + 2 * ii[1]);
              here.set_direction(dd[2], is_[2] + 2 * ii[2]);
              const ptrdiff_t n = idx0 + ii[0] * ss[0] + ii[1] * ss[1] + ii[2] * ss[2];
              component c;
              ivec there(vi.dim);
              std::complex<double> thephase;
              int j;
              if (!head_of_run(corig, here, c, there, thephase, j)) {
                ++k;
                continue;
              }
              // ---- length of the run
              ptrdiff_t len = 1;
              if (runs_ok && nn[rdim] - k > 1 && ss[rdim] != 0) {
                const int tr = there.in_direction(rd);
                int hi = chunks[j]->gv.big_corner().in_direction(rd); // owned: coordinate <= io + 2n
                hi = std::min(hi, user_volume.big_corner().in_direction(rd));
                len = std::min<ptrdiff_t>(nn[rdim] - k, (hi - tr) / 2 + 1);
                if (user_volume.has_boundary(High, rd) && boundaries[High][rd] == Metallic) {
                  const int m = user_volume.big_corner().in_direction(rd);
                  if (m > tr && (m - tr) % 2 == 0) len = std::min<ptrdiff_t>(len, (m - tr) / 2);
                }
                if (len < 1) len = 1;
                if (len > 1) { // re-check the far end with the per-point logic
                  ivec here2 = here, there2(vi.dim);
                  here2.set_direction(rd, here.in_direction(rd) + 2 * (int)(len - 1));
                  component c2;
                  std::complex<double> ph2;
                  int j2;
                  const int keep_last = last_j;
                  ivec expect = there;
                  expect.set_direction(rd, tr + 2 * (int)(len - 1));
                  if (!head_of_run(corig, here2, c2, there2, ph2, j2) || j2 != j || c2 != c || ph2 != thephase ||
                      there2 != expect)
                    len = 1;
                  last_j = keep_last;
                }
              }
              const bool j_is_mine = chunks[j]->is_mine();
              if ((!i_is_mine && !j_is_mine) ||
                  (is_B(corig) && is_B(c) && B_redundant[5 * i + corig - Bx] && B_redundant[5 * j + c - Bx])) {
                k += len;
                continue;
              }

              const connect_phase ip = connect_phase_from_phase(thephase);
              const ptrdiff_t m = chunks[j]->gv.index(c, there);
              const ptrdiff_t dn = ss[rdim], dm = chunks[j]->gv.stride(rd);
              const std::complex<realnum> ph(thephase.real(), thephase.imag());

              {
                Slot &s = slot(type(c), ip, j);
                if (i_is_mine) {
                  if (ip == CONNECT_PHASE) s.phases->insert(s.phases->end(), (size_t)len, ph);
                  realnum *base[2] = {chunks[i]->f[corig][0] + n, ncmp > 1 ? chunks[i]->f[corig][1] + n : nullptr};
                  for (ptrdiff_t q = 0; q < len; ++q)
                    for (int cmp = 0; cmp < ncmp; cmp++)
                      s.in->push_back(base[cmp] + q * dn);
                }
                if (j_is_mine) {
                  realnum *base[2] = {chunks[j]->f[c][0] + m, ncmp > 1 ? chunks[j]->f[c][1] + m : nullptr};
                  for (ptrdiff_t q = 0; q < len; ++q)
                    for (int cmp = 0; cmp < ncmp; cmp++)
                      s.out->push_back(base[cmp] + q * dm);
                }
              }

              if (needs_W_notowned[corig]) {
                Slot &s = slot(is_electric(corig) ? WE_stuff : WH_stuff, ip, j);
                if (i_is_mine) {
                  if (ip == CONNECT_PHASE) s.phases->insert(s.phases->end(), (size_t)len, ph);
                  realnum *base[2];
                  for (int cmp = 0; cmp < ncmp; cmp++)
                    base[cmp] = (chunks[i]->f_w[corig][cmp] ? chunks[i]->f_w[corig][cmp] : chunks[i]->f[corig][cmp]) + n;
                  for (ptrdiff_t q = 0; q < len; ++q)
                    for (int cmp = 0; cmp < ncmp; cmp++)
                      s.in->push_back(base[cmp] + q * dn);
                }
                if (j_is_mine) {
                  realnum *base[2];
                  for (int cmp = 0; cmp < ncmp; cmp++)
                    base[cmp] = (chunks[j]->f_w[c][cmp] ? chunks[j]->f_w[c][cmp] : chunks[j]->f[c][cmp]) + m;
                  for (ptrdiff_t q = 0; q < len; ++q)
                    for (int cmp = 0; cmp < ncmp; cmp++)
                      s.out->push_back(base[cmp] + q * dm);
                }
              }

              if (is_electric(corig) || is_magnetic(corig)) {
                const field_type f = is_electric(corig) ? PE_stuff : PH_stuff;
                // which polarisation pairs exchange internal values here: the same for every point
                // of the run (depends on the components and the data blocks only)
                pol_pairs.clear();
                for (polarization_state *pi = chunks[i]->pol[type(corig)]; pi; pi = pi->next)
                  for (polarization_state *pj = chunks[j]->pol[type(c)]; pj; pj = pj->next)
                    if (*pi->s == *pj->s) {
                      polarization_state *po = NULL;
                      if (pi->data && i_is_mine)
                        po = pi;
                      else if (pj->data && j_is_mine)
                        po = pj;
                      if (po) {
                        const size_t ni = po->s->num_internal_notowned_needed(corig, po->data);
                        const size_t cni = po->s->num_cinternal_notowned_needed(corig, po->data);
                        if (ni || cni) pol_pairs.push_back(PolPair{pi, pj, po, ni, cni});
                      }
                    }
                if (!pol_pairs.empty())
                  for (ptrdiff_t q = 0; q < len; ++q) { // point-major, as the reference pushes them
                    const ptrdiff_t nq = n + q * dn, mq = m + q * dm;
                    for (const PolPair &pp : pol_pairs) {
                      if (pp.ni) {
                        Slot &s = slot(f, CONNECT_COPY, j);
                        for (size_t kk = 0; kk < pp.ni; ++kk) {
                          if (i_is_mine) s.in->push_back(pp.po->s->internal_notowned_ptr(kk, corig, nq, pp.pi->data));
                          if (j_is_mine) s.out->push_back(pp.po->s->internal_notowned_ptr(kk, c, mq, pp.pj->data));
                        }
                      }
                      if (pp.cni) {
                        Slot &s = slot(f, ip, j);
                        for (size_t kk = 0; kk < pp.cni; ++kk) {
                          if (i_is_mine) {
                            if (ip == CONNECT_PHASE) s.phases->push_back(ph);
                            for (int cmp = 0; cmp < ncmp; cmp++)
                              s.in->push_back(pp.po->s->cinternal_notowned_ptr(kk, corig, cmp, nq, pp.pi->data));
                          }
                          if (j_is_mine)
                            for (int cmp = 0; cmp < ncmp; cmp++)
                              s.out->push_back(pp.po->s->cinternal_notowned_ptr(kk, c, cmp, mq, pp.pj->data));
                        }
                      }
                    }
                  }
              }
              k += len;
            }
      }
    }

    // sizes of the comm blocks of every pair (j -> i) that was touched (src/boundaries.cpp:406-451)
    for (size_t k : touched_list) {
      const Slot &s = slots[k];
      res.sizes.push_back(s.in ? s.in->size() : (s.out ? s.out->size() : 0));
    }
  };

  // workers: chunks are independent (see above); MEEP_B200_HOST_THREADS overrides the count
  {
    int nthreads = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("MEEP_B200_HOST_THREADS")) nthreads = atoi(e);
    const int local_ranks = getenv("LOCAL_WORLD_SIZE") ? atoi(getenv("LOCAL_WORLD_SIZE")) : count_processors();
    if (!getenv("MEEP_B200_HOST_THREADS") && local_ranks > 1) nthreads /= local_ranks;
    nthreads = std::max(1, std::min(nthreads, std::min(num_chunks, 16)));
    std::atomic<int> next(0);
    auto worker = [&]() {
      for (int i = next.fetch_add(1); i < num_chunks; i = next.fetch_add(1))
        process_chunk(i);
    };
    if (nthreads == 1)
      worker();
    else {
      std::vector<std::thread> pool;
      for (int t = 0; t < nthreads; ++t)
        pool.emplace_back(worker);
      for (std::thread &t : pool)
        t.join();
    }
  }

  // merge, in chunk order
  for (int i = 0; i < num_chunks; i++) {
    ChunkResult &res = results[i];
    if (!res.done) continue;
    std::vector<std::vector<realnum *> *> staged_of(nslots, nullptr);
    for (Staged &st : res.staged)
      staged_of[st.k] = &st.out;
    for (size_t q = 0; q < res.touched_list.size(); ++q) {
      const size_t k = res.touched_list[q], sz = res.sizes[q];
      const int j = (int)(k % num_chunks);
      const size_t fi = k / num_chunks;
      const comms_key key = {field_type(fi / NUM_CONNECT_PHASE_TYPES),
                             connect_phase(fi % NUM_CONNECT_PHASE_TYPES), {j, i}};
      if (sz) {
        comm_sizes[key] = sz;
        if (staged_of[k]) chunks[j]->connections_out[key] = std::move(*staged_of[k]);
      }
      else { // nothing was connected under this key after all: leave no empty table behind
        chunks[i]->connections_in.erase(key);
        chunks[i]->connection_phases.erase(key);
      }
    }
  }

  // (the host comm_blocks of the reference are not allocated: comm blocks live in HBM,
  //  fields::step_boundaries in step.cpp)

  // comms_sequence_for_field (the reference's ordered list of MPI sends/receives, consumed only
  // by its own fields::step_boundaries) has no reader in this build: our step_boundaries walks
  // comm_sizes in a fixed global pair order and moves the blocks device-to-device.
  FOR_FIELD_TYPES(f) { comms_sequence_for_field[f].clear(); }
  if (getenv("MEEP_B200_VERBOSE") && atoi(getenv("MEEP_B200_VERBOSE")))
    master_printf("meep_b200: connect_the_chunks: %d chunks, %.3f s\n", num_chunks,
                  wall_time() - t_start);
}

} // namespace meep
