// engine.cpp — see engine.hpp.
#include "engine.hpp"
#include <execinfo.h>
#include "meep_internals.hpp"
#include "comm.hpp"
#include "hostmem.hpp"
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <typeinfo>
#include <unordered_map>

using namespace meep;

namespace meep_b200 {

Engine *Engine::current_ = nullptr;
uint64_t Engine::source_generation = 0;

static std::unordered_map<const fields *, Engine *> &table() {
  static std::unordered_map<const fields *, Engine *> t;
  return t;
}

void check(int rc, const char *what) {
  if (rc) meep::abort("meep_b200: %s failed: %s", what, mb200_last_error());
}

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

Engine &Engine::get(fields *f) {
  auto &t = table();
  auto it = t.find(f);
  if (it != t.end()) return *it->second;
  Engine *e = new Engine(f);
  t[f] = e;
  return *e;
}

Engine *Engine::find(const fields *f) {
  auto &t = table();
  auto it = t.find(f);
  return it == t.end() ? nullptr : it->second;
}

void Engine::drop(const fields *f) {
  auto &t = table();
  auto it = t.find(f);
  if (it == t.end()) return;
  delete it->second;
  t.erase(it);
}

bool Engine::mirrors(const void *host) const {
  const uintptr_t a = (uintptr_t)host;
  auto it = arrs_.upper_bound(a);
  if (it == arrs_.begin()) return false;
  --it;
  return a >= it->first && a < it->first + it->second.bytes;
}

Engine *Engine::owner_of(const void *p) {
  if (!p) return nullptr;
  for (auto &kv : table())
    if (kv.second->mirrors(p)) return kv.second;
  return nullptr;
}

void Engine::for_each(const std::function<void(Engine &)> &fn) {
  for (auto &kv : table())
    fn(*kv.second);
}

Engine::Engine(fields *f) : self(f) {
  if (mb200_abi_version() != MB200_ABI_VERSION)
    meep::abort("meep_b200: libmeepb200 ABI version mismatch");
  if (mb200_device_count() < 1)
    meep::abort("meep_b200: no CUDA device visible — this build of libmeep has no CPU "
                "time-stepping path (fields::step requires a B200)");
  int device = env_int("MEEP_B200_DEVICE", env_int("LOCAL_RANK", 0));
  if (device >= mb200_device_count()) device = device % mb200_device_count();
  const double t_init = meep::wall_time();
  check(mb200_init(device, &ctx), "mb200_init");
  if (env_int("MEEP_B200_VERBOSE", 0))
    fprintf(stderr, "meep_b200: device context on GPU %d ready in %.3f s\n", device, meep::wall_time() - t_init);
  emulated = dlsym(RTLD_DEFAULT, "mb200_is_emulator") != nullptr;
  fuse = env_int("MEEP_B200_FUSE", 1) != 0;
  eager = env_int("MEEP_B200_EAGER", 0) != 0;
  lazy_host = env_int("MEEP_B200_LAZY_HOST", 1) != 0;
  release_host = env_int("MEEP_B200_RELEASE_HOST", 1);
  verbose = env_int("MEEP_B200_VERBOSE", 0) != 0;
  device_timers = env_int("MEEP_B200_TIMERS", 0) != 0;
  merge_exchanges = env_int("MEEP_B200_MERGE_EXCHANGES", 1) != 0;
  defer_local = env_int("MEEP_B200_DEFER_LOCAL", 1) != 0;
  if (getenv("MEEP_B200_REGION_SYNC")) region_fraction_limit = atof(getenv("MEEP_B200_REGION_SYNC"));
  zero_skip = env_int("MEEP_B200_ZERO_SKIP", 1) != 0;
  p2p = env_int("MEEP_B200_P2P", 1) != 0;
  halo_runs = env_int("MEEP_B200_HALO_RUNS", 1) != 0;
  halo_sort = env_int("MEEP_B200_HALO_SORT", 1) != 0;
  plain_t1 = env_int("MEEP_B200_PLAIN_T1", 16);
  if (plain_t1 < 0 || plain_t1 > 64) plain_t1 = 16;
  pml_t1 = env_int("MEEP_B200_PML_T1", 16);
  if (pml_t1 < 1 || pml_t1 > 64) pml_t1 = 16;
  emulated = dlsym(RTLD_DEFAULT, "mb200_is_emulator") != nullptr;
  nan_check_every = env_int("MEEP_B200_NAN_CHECK_EVERY", 16);
  if (nan_check_every < 1) nan_check_every = 1;
}

Engine::~Engine() {
  if (current_ == this) current_ = nullptr;
  if (verbose)
    fprintf(stderr, "meep_b200: engine summary: %lld steps, %lld uploads, %lld downloads, %lld region downloads, "
            "%.3f MB h2d, %.3f MB d2h, %lld plan builds\n", (long long)stats.steps, (long long)stats.uploads,
            (long long)stats.downloads, (long long)stats.region_downloads, stats.h2d_bytes / 1e6, stats.d2h_bytes / 1e6,
            (long long)stats.plan_builds);
  recording_ = false;
  invalidate_plans();
  drop_links();
  drop_backups();
  for (auto &kv : arrs_)
    mb200_free(ctx, kv.second.dev);
  arrs_.clear();
  if (probe_flag_) mb200_free(ctx, probe_flag_);
  for (auto &kv : pzero_)
    mb200_free(ctx, kv.second.dev);
  for (auto &kv : szero_)
    mb200_free(ctx, kv.second.dev);
  if (comm) mb200_comm_destroy(comm);
  mb200_destroy(ctx);
}

// ---- mirror --------------------------------------------------------------------------------------

void *Engine::dev(const void *host) const {
  if (!host) return nullptr;
  const uintptr_t a = (uintptr_t)host;
  // table translation asks for millions of consecutive elements of the same array
  if (a >= hit_lo_ && a < hit_hi_) return hit_dev_ + (a - hit_lo_);
  auto it = arrs_.upper_bound(a);
  if (it != arrs_.begin()) {
    --it;
    if (a >= it->first && a < it->first + it->second.bytes) {
      hit_lo_ = it->first;
      hit_hi_ = it->first + it->second.bytes;
      hit_dev_ = (char *)it->second.dev;
      return hit_dev_ + (a - hit_lo_);
    }
  }
  meep::abort("meep_b200: host pointer %p is not part of any mirrored array", host);
  return nullptr;
}

void *Engine::ensure(const void *host, size_t bytes, bool is_field, int init) {
  if (!host || bytes == 0) return nullptr;
  const uintptr_t a = (uintptr_t)host;
  auto it = arrs_.find(a);
  if (it != arrs_.end()) {
    if (it->second.bytes == bytes) {
      it->second.seen = true;
      return it->second.dev;
    }
    mb200_free(ctx, it->second.dev); // same address, different size: the host re-allocated
    arrs_.erase(it);
    hit_lo_ = hit_hi_ = 0;
    if (backups.count(host)) {
      mb200_free(ctx, backups[host]);
      backups.erase(host);
    }
    invalidate_plans();
  }
  Arr arr;
  arr.bytes = bytes;
  arr.is_field = is_field;
  arr.seen = true;
  check(mb200_malloc(ctx, bytes, &arr.dev), "mb200_malloc");
  if (init == 0) {
    upload_array(host, arr.dev, bytes);
    arr.fresh = true; // just uploaded: the bulk uploads of this enter() skip it
  }
  arrs_[a] = arr;
  invalidate_plans();
  return arr.dev;
}

// Host array -> device twin.  An array whose interior pages were never touched still holds the
// zeros it was allocated with (hostmem.hpp): its twin is a device memset plus the two partial
// pages at its ends, and the host pages stay unborn.
void Engine::upload_array(const void *host, void *dev, size_t bytes) {
  char *lo, *hi;
  if (lazy_host && interior_untouched(host, bytes) && page_interior(host, bytes, &lo, &hi)) {
    const size_t head = (size_t)(lo - (const char *)host), tail = (size_t)((const char *)host + bytes - hi);
    check(mb200_memset(ctx, dev, 0, bytes), "mb200_memset");
    if (head) check(mb200_h2d(ctx, dev, host, head), "mb200_h2d");
    if (tail) check(mb200_h2d(ctx, (char *)dev + (bytes - tail), hi, tail), "mb200_h2d");
    stats.h2d_bytes += head + tail;
    return;
  }
  check(mb200_h2d(ctx, dev, host, bytes), "mb200_h2d");
  stats.h2d_bytes += bytes;
}

void Engine::ensure_from(const void *host, size_t bytes, const void *src_host) {
  void *d = ensure(host, bytes, true, 1);
  if (src_host) check(mb200_d2d(ctx, d, dev(src_host), bytes), "mb200_d2d");
  else check(mb200_memset(ctx, d, 0, bytes), "mb200_memset");
}

void Engine::forget(const void *host) {
  auto it = arrs_.find((uintptr_t)host);
  if (it == arrs_.end()) return;
  invalidate_plans();
  mb200_free(ctx, it->second.dev);
  arrs_.erase(it);
  hit_lo_ = hit_hi_ = 0;
  if (backups.count(host)) {
    mb200_free(ctx, backups[host]);
    backups.erase(host);
  }
}

void *Engine::aux_upload(const void *host, size_t bytes) {
  void *d = nullptr;
  if (!recording_) meep::abort("meep_b200: aux_upload outside a phase recording");
  check(mb200_malloc(ctx, bytes ? bytes : 8, &d), "mb200_malloc(aux)");
  if (bytes) check(mb200_h2d(ctx, d, host, bytes), "mb200_h2d(aux)");
  rec_aux_.push_back(d);
  return d;
}

void *Engine::aux_alloc(size_t bytes) {
  void *d = nullptr;
  if (!recording_) meep::abort("meep_b200: aux_alloc outside a phase recording");
  check(mb200_malloc(ctx, bytes ? bytes : 8, &d), "mb200_malloc(aux)");
  rec_aux_.push_back(d);
  return d;
}

// ---- device-side backups (synchronize_magnetic_fields) ----------------------------------------------

void Engine::backup_array(const void *host) {
  if (!host) return;
  auto it = arrs_.find((uintptr_t)host);
  if (it == arrs_.end()) meep::abort("meep_b200: backup of an array that has no device mirror");
  void *&b = backups[host];
  if (!b) check(mb200_malloc(ctx, it->second.bytes, &b), "mb200_malloc(backup)");
  check(mb200_d2d(ctx, b, it->second.dev, it->second.bytes), "mb200_d2d(backup)");
}

void Engine::restore_array(const void *host) {
  auto b = backups.find(host);
  if (!host || b == backups.end()) return;
  auto it = arrs_.find((uintptr_t)host);
  if (it == arrs_.end()) return;
  check(mb200_d2d(ctx, it->second.dev, b->second, it->second.bytes), "mb200_d2d(restore)");
}

void Engine::drop_backups() {
  for (auto &kv : backups)
    mb200_free(ctx, kv.second);
  backups.clear();
}

// ---- peer-memory links ----------------------------------------------------------------------------

int Engine::find_link(int ft, int rank) const {
  for (size_t k = 0; k < links.size(); ++k)
    if (links[k].ft == ft && links[k].rank == rank) return (int)k;
  return -1;
}

// Collective over all processes whenever any link exists anywhere (the callers guarantee this:
// links exist on both ends or on neither).
void Engine::drop_links() {
  if (links.empty()) return;
  // neighbours may still be storing into my arenas, and I into theirs: drain the device, meet the
  // other processes, unmap, meet again, free.
  mb200_sync(ctx);
  comm_barrier();
  for (P2PLink &p : links)
    if (p.theirs) mb200_ipc_close(ctx, p.theirs);
  comm_barrier();
  for (P2PLink &p : links)
    if (p.mine) mb200_free(ctx, p.mine);
  links.clear();
}

void Engine::rebuild_links(const std::map<std::pair<int, int>, std::pair<size_t, size_t> > &counts) {
  // boundary phases recorded against the old arenas are stale
  for (int ft = 0; ft < NUM_FIELD_TYPES; ++ft)
    free_phase(phases_[PH_BND][ft]);
  // "does anybody have links" must be answered identically everywhere: links are symmetric, and
  // every process runs this function at the same point of the schedule
  int any_old = links.empty() ? 0 : 1;
  comm_allreduce(&any_old, sizeof(int), 1, [](void *acc, const void *in, size_t n) {
    for (size_t i = 0; i < n; ++i)
      ((int *)acc)[i] = ((int *)acc)[i] > ((const int *)in)[i] ? ((int *)acc)[i] : ((const int *)in)[i];
  }, true);
  if (any_old) {
    mb200_sync(ctx);
    comm_barrier();
    for (P2PLink &p : links)
      if (p.theirs) mb200_ipc_close(ctx, p.theirs);
    comm_barrier();
    for (P2PLink &p : links)
      if (p.mine) mb200_free(ctx, p.mine);
    links.clear();
  }
  links_epoch = connect_epoch;
  const size_t Rsz = sizeof(meep::realnum);
  int ok = 1;
  std::vector<char> handles(counts.size() * 64), peer_handles(counts.size() * 64);
  size_t k = 0;
  for (const auto &kv : counts) {
    P2PLink p;
    p.ft = kv.first.first;
    p.rank = kv.first.second;
    p.send_count = kv.second.first;
    p.recv_count = kv.second.second;
    const size_t bytes = kArenaHeader + p.recv_count * Rsz;
    if (mb200_malloc(ctx, bytes, &p.mine) != 0 || mb200_memset(ctx, p.mine, 0, bytes) != 0 ||
        mb200_ipc_export(ctx, p.mine, &handles[64 * k]) != 0)
      ok = 0;
    links.push_back(p);
    ++k;
  }
  if (mb200_sync(ctx) != 0) ok = 0; // the zeroed headers must be in place before anybody signals
  // swap the handles (std::map order = (ft, peer): both ends list their common links by ft order)
  std::vector<HostMsg> sends, recvs;
  for (size_t i = 0; i < links.size(); ++i) {
    sends.push_back(HostMsg{links[i].rank, &handles[64 * i], 64});
    recvs.push_back(HostMsg{links[i].rank, &peer_handles[64 * i], 64});
  }
  comm_sendrecv_all(sends, recvs);
  for (size_t i = 0; i < links.size() && ok; ++i)
    if (mb200_ipc_import(ctx, &peer_handles[64 * i], &links[i].theirs) != 0) ok = 0;
  comm_allreduce(&ok, sizeof(int), 1, [](void *acc, const void *in, size_t n) {
    for (size_t i = 0; i < n; ++i)
      ((int *)acc)[i] = ((int *)acc)[i] < ((const int *)in)[i] ? ((int *)acc)[i] : ((const int *)in)[i];
  }, true);
  if (!ok) {
    // no peer mapping between some pair of GPUs: everybody falls back to the NCCL exchange
    if (verbose || comm_rank() == 0)
      fprintf(stderr, "meep_b200: peer-memory exchange unavailable (%s); using NCCL send/recv\n",
              mb200_last_error());
    for (P2PLink &p : links)
      if (p.theirs) mb200_ipc_close(ctx, p.theirs);
    comm_barrier();
    for (P2PLink &p : links)
      if (p.mine) mb200_free(ctx, p.mine);
    links.clear();
    p2p = false;
    return;
  }
  comm_barrier();
  if (verbose) {
    size_t tot = 0;
    for (const P2PLink &p : links)
      tot += p.recv_count * Rsz;
    fprintf(stderr, "meep_b200[%d]: %zu peer-memory links, %.1f MB of arenas\n", comm_rank(), links.size(),
            tot / 1e6);
  }
}

// One communicator per Engine; the 128-byte id is made on rank 0 and broadcast over the socket
// runtime (comm.hpp).  Collective: every rank reaches this from its first step_boundaries.
void Engine::ensure_comm() {
  if (comm || comm_size() == 1) return;
  char id[128];
  memset(id, 0, sizeof(id));
  if (comm_rank() == 0) check(mb200_comm_unique_id(id), "mb200_comm_unique_id");
  comm_broadcast(0, id, sizeof(id));
  check(mb200_comm_create(ctx, comm_rank(), comm_size(), id, &comm), "mb200_comm_create");
}

void Engine::free_phase(Phase &ph) {
  ph.links.clear();
  for (Launch &l : ph.launches)
    if (l.plan) mb200_plan_destroy(ctx, l.plan);
  ph.launches.clear();
  for (void *p : ph.aux)
    mb200_free(ctx, p);
  ph.aux.clear();
  ph.valid = false;
  ph.one_shot = false;
}

void Engine::invalidate_plans() {
  if (recording_) {
    pending_invalidate_ = true;
    return;
  }
  for (int id = 0; id < PH_COUNT; ++id)
    for (int ft = 0; ft < NUM_FIELD_TYPES; ++ft)
      free_phase(phases_[id][ft]);
  for (int ft = 0; ft < NUM_FIELD_TYPES; ++ft) {
    fused_eh[ft].clear();
    deferred_exchange[ft] = false;
  }
  if (probe_ptrs_) {
    mb200_free(ctx, probe_ptrs_);
    probe_ptrs_ = nullptr;
    probe_n_ = 0;
  }
}

static inline void hash_mix(uint64_t &h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
}

// Everything that determines the job tables: which arrays exist (by address), chunk shapes,
// sources and DFT monitors.
uint64_t Engine::fingerprint(fields *f) const {
  uint64_t h = 1469598103934665603ULL;
  hash_mix(h, (uint64_t)f->num_chunks);
  hash_mix(h, (uint64_t)f->is_real);
  hash_mix(h, source_generation);
  for (int i = 0; i < f->num_chunks; ++i) {
    fields_chunk *fc = f->chunks[i];
    if (!fc->is_mine()) continue;
    hash_mix(h, (uint64_t)fc->gv.ntot());
    hash_mix(h, (uint64_t)(uintptr_t)fc->s);
    FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) {
      hash_mix(h, (uint64_t)(uintptr_t)fc->f[c][cmp]);
      hash_mix(h, (uint64_t)(uintptr_t)fc->f_u[c][cmp]);
      hash_mix(h, (uint64_t)(uintptr_t)fc->f_w[c][cmp]);
      hash_mix(h, (uint64_t)(uintptr_t)fc->f_cond[c][cmp]);
      hash_mix(h, (uint64_t)(uintptr_t)fc->f_bfast[c][cmp]);
      hash_mix(h, (uint64_t)(uintptr_t)fc->f_minus_p[c][cmp]);
      hash_mix(h, (uint64_t)(uintptr_t)fc->f_w_prev[c][cmp]);
    }
    const structure_chunk *s = fc->s;
    FOR_COMPONENTS(c) {
      hash_mix(h, (uint64_t)(uintptr_t)s->chi2[c]);
      hash_mix(h, (uint64_t)(uintptr_t)s->chi3[c]);
      FOR_DIRECTIONS(d) {
        hash_mix(h, (uint64_t)(uintptr_t)s->chi1inv[c][d]);
        hash_mix(h, (uint64_t)(uintptr_t)s->conductivity[c][d]);
        hash_mix(h, (uint64_t)(uintptr_t)s->condinv[c][d]);
      }
    }
    for (int d = 0; d < 6; ++d) {
      hash_mix(h, (uint64_t)(uintptr_t)s->sig[d]);
      hash_mix(h, (uint64_t)s->sigsize[d]);
    }
    FOR_FIELD_TYPES(ft) {
      for (polarization_state *p = fc->pol[ft]; p; p = p->next) {
        hash_mix(h, (uint64_t)(uintptr_t)p->data);
        hash_mix(h, (uint64_t)(uintptr_t)p->s);
      }
      for (const src_vol &sv : fc->get_sources(ft)) {
        hash_mix(h, (uint64_t)sv.num_points());
        hash_mix(h, (uint64_t)(uintptr_t)sv.t());
        hash_mix(h, (uint64_t)sv.c);
        if (sv.num_points()) hash_mix(h, (uint64_t)(uintptr_t)&sv.amplitude_at(0));
      }
    }
    for (dft_chunk *d = fc->dft_chunks; d; d = d->next_in_chunk) {
      hash_mix(h, (uint64_t)(uintptr_t)d);
      hash_mix(h, (uint64_t)(uintptr_t)d->dft);
      hash_mix(h, (uint64_t)d->get_decimation_factor());
    }
  }
  return h;
}

void Engine::scan(fields *f) {
  for (auto &kv : arrs_)
    kv.second.seen = false;
  const size_t R = sizeof(realnum);
  for (int i = 0; i < f->num_chunks; ++i) {
    fields_chunk *fc = f->chunks[i];
    if (!fc->is_mine()) continue;
    const size_t nb = fc->gv.ntot() * R;
    FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) {
      ensure(fc->f[c][cmp], nb, true);
      ensure(fc->f_u[c][cmp], nb, true);
      ensure(fc->f_w[c][cmp], nb, true);
      ensure(fc->f_cond[c][cmp], nb, true);
      ensure(fc->f_bfast[c][cmp], nb, true);
      // f_minus_p is scratch that is fully rewritten before every use: no upload needed
      ensure(fc->f_minus_p[c][cmp], nb, true, 1);
      if (fc->f_w_prev[c][cmp])
        meep::abort("meep_b200: susceptibilities that need W_prev (multilevel atoms) are not "
                    "supported on the device path");
    }
    const structure_chunk *s = fc->s;
    FOR_COMPONENTS(c) {
      ensure(s->chi2[c], nb, false);
      ensure(s->chi3[c], nb, false);
      FOR_DIRECTIONS(d) {
        ensure(s->chi1inv[c][d], nb, false);
        ensure(s->conductivity[c][d], nb, false);
        ensure(s->condinv[c][d], nb, false);
      }
    }
    for (int d = 0; d < 5; ++d)
      if (s->sigsize[d] > 1) {
        ensure(s->sig[d], s->sigsize[d] * R, false);
        ensure(s->kap[d], s->sigsize[d] * R, false);
        ensure(s->siginv[d], s->sigsize[d] * R, false);
      }
    FOR_FIELD_TYPES(ft) {
      if (ft != E_stuff && ft != H_stuff) continue;
      for (susceptibility *sus = s->chiP[ft]; sus; sus = sus->next)
        FOR_COMPONENTS(c) FOR_DIRECTIONS(d) ensure(sus->sigma[c][d], nb, false);
      for (polarization_state *p = fc->pol[ft]; p; p = p->next)
        if (p->data) {
          const std::pair<realnum *, size_t> blk = polarisation_block(p->s, p->data);
          if (blk.second) ensure(blk.first, blk.second, true);
        }
    }
    for (dft_chunk *d = fc->dft_chunks; d; d = d->next_in_chunk)
      ensure(d->dft, d->N * d->omega.size() * 2 * R, true);
  }
  // drop mirrors of arrays the host no longer has
  for (auto it = arrs_.begin(); it != arrs_.end();) {
    if (!it->second.seen) {
      mb200_free(ctx, it->second.dev);
      it = arrs_.erase(it);
      hit_lo_ = hit_hi_ = 0;
      invalidate_plans();
    }
    else
      ++it;
  }
}

uint8_t *Engine::pzero_flags(const void *host_P, size_t ntot, bool known_zero) {
  if (!zero_skip || !host_P) return nullptr;
  auto it = pzero_.find((uintptr_t)host_P);
  if (it != pzero_.end() && it->second.ntot == ntot) return it->second.dev;
  if (it != pzero_.end()) {
    mb200_free(ctx, it->second.dev);
    pzero_.erase(it);
  }
  Flags fl;
  fl.ntot = ntot;
  fl.nblocks = (ntot + MB200_ZBLOCK - 1) / MB200_ZBLOCK;
  void *d = nullptr;
  check(mb200_malloc(ctx, fl.nblocks, &d), "malloc(pzero)");
  check(mb200_memset(ctx, d, known_zero ? 1 : 0, fl.nblocks), "memset(pzero)");
  fl.dev = (uint8_t *)d;
  pzero_[(uintptr_t)host_P] = fl;
  return fl.dev;
}

void Engine::pzero_drop(const void *host_P) {
  auto it = pzero_.find((uintptr_t)host_P);
  if (it == pzero_.end()) return;
  mb200_free(ctx, it->second.dev);
  pzero_.erase(it);
  invalidate_plans(); // f_minus_p jobs may hold the flag pointer
}

uint8_t *Engine::pzero_lookup(const void *host_elem) const {
  const uintptr_t a = (uintptr_t)host_elem;
  auto it = pzero_.upper_bound(a);
  if (it == pzero_.begin()) return nullptr;
  --it;
  if (a >= it->first && a < it->first + it->second.ntot * sizeof(realnum)) return it->second.dev;
  return nullptr;
}

uint64_t Engine::pzero_flag_addr(const void *host_elem) const {
  const uintptr_t a = (uintptr_t)host_elem;
  auto it = pzero_.upper_bound(a);
  if (it == pzero_.begin()) return 0;
  --it;
  if (a < it->first || a >= it->first + it->second.ntot * sizeof(realnum)) return 0;
  const size_t idx = (a - it->first) / sizeof(realnum);
  return (uint64_t)(uintptr_t)(it->second.dev + idx / MB200_ZBLOCK);
}

const uint8_t *Engine::szero_flags(const void *host_sigma, size_t ntot) {
  if (!zero_skip || !host_sigma) return nullptr;
  Flags &fl = szero_[(uintptr_t)host_sigma];
  if (fl.dev && fl.ntot != ntot) {
    mb200_free(ctx, fl.dev);
    fl = Flags();
  }
  if (!fl.dev) {
    fl.ntot = ntot;
    fl.nblocks = (ntot + MB200_ZBLOCK - 1) / MB200_ZBLOCK;
    void *d = nullptr;
    check(mb200_malloc(ctx, fl.nblocks, &d), "malloc(szero)");
    fl.dev = (uint8_t *)d;
    fl.fresh = false;
  }
  if (!fl.fresh) {
    check(mb200_block_zero_flags(ctx, dtype, dev(host_sigma), (int64_t)ntot, fl.dev), "block_zero_flags");
    fl.fresh = true;
  }
  return fl.dev;
}

void Engine::upload_fields() {
  // host values replace the device ones: nothing is known about zero blocks any more
  for (auto &kv : pzero_)
    check(mb200_memset(ctx, kv.second.dev, 0, kv.second.nblocks), "memset(pzero)");
  for (auto &kv : arrs_)
    if (kv.second.is_field) {
      if (!kv.second.fresh) upload_array((const void *)kv.first, kv.second.dev, kv.second.bytes);
      kv.second.fresh = false;
    }
  stats.uploads++;
  host_resident = true;
}

void Engine::download_fields() {
  // all copies queued, then ONE synchronisation (a small run has tens of arrays: a stream
  // synchronisation per array made a download cost more than the step it follows)
  for (auto &kv : arrs_)
    if (kv.second.is_field) {
      check(mb200_d2h_async(ctx, (void *)kv.first, kv.second.dev, kv.second.bytes), "d2h");
      stats.d2h_bytes += kv.second.bytes;
    }
  check(mb200_sync(ctx), "device synchronisation");
  stats.downloads++;
  host_resident = true;
}

// The device copy is the authoritative one: the host pages of the field arrays only hold stale
// values.  MEEP_B200_RELEASE_HOST: 0 = keep them, 1 (default) = drop them once after an upload
// (a run that never downloads then never holds field pages on the host: 1024^3 on one GPU needs
// ~30 GB of host RAM instead of ~125 GB), 2 = also after every download.
void Engine::release_host_fields() {
  if (!host_resident || release_host == 0 || state != DEVICE_NEWER) return;
  if (release_host == 1 && stats.downloads > released_at_download) {
    host_resident = false; // a reader pulled the arrays: assume it will again, keep the pages
    released_at_download = stats.downloads;
    return;
  }
  for (auto &kv : arrs_)
    if (kv.second.is_field) release_interior((void *)kv.first, kv.second.bytes);
  host_resident = false;
  released_at_download = stats.downloads;
}

void Engine::upload_materials() {
  for (auto &kv : szero_)
    kv.second.fresh = false; // recomputed on next use
  for (auto &kv : arrs_)
    if (!kv.second.is_field) {
      if (!kv.second.fresh) {
        check(mb200_h2d(ctx, kv.second.dev, (const void *)kv.first, kv.second.bytes), "h2d");
        stats.h2d_bytes += kv.second.bytes;
      }
      kv.second.fresh = false;
    }
  materials_dirty = false;
}

// The same-device D/B halo copies that fields::step_boundaries postponed: made now, from the
// current owner values, through the ordinary exchange code restricted to same-process pairs (no
// other process is needed, so any single rank may do this at any time).
void Engine::refresh_deferred_halos() {
  bool any = false;
  for (field_type ft : {B_stuff, D_stuff, PH_stuff, PE_stuff})
    any = any || halos_stale[ft];
  if (refresh_local || !any) return;
  const bool was_in_step = in_step;
  refresh_local = true;
  in_step = false;
  keep_on_device++;
  for (field_type ft : {B_stuff, D_stuff, PH_stuff, PE_stuff})
    if (halos_stale[ft]) {
      halos_stale[ft] = false;
      self->step_boundaries(ft);
    }
  keep_on_device--;
  in_step = was_in_step;
  refresh_local = false;
}

void Engine::sync_host() {
  if (state == DEVICE_NEWER) {
    refresh_deferred_halos();
    download_fields();
    state = COHERENT;
    // whoever reads the arrays on the host now must not be handed NaNs (or stale halos after a
    // peer time-out) silently: the per-step probe is only read back every nan_check_every steps
    check(mb200_sync(ctx), "device synchronisation");
    if (probe_flag_ && probe_n_) {
      int32_t flag = 0;
      check(mb200_d2h(ctx, &flag, probe_flag_, 4), "d2h(flag)");
      stats.d2h_bytes += 4;
      if (flag) meep::abort("simulation fields are NaN or Inf");
    }
  }
}

void Engine::sync_host_region(const volume &where) {
  // (force_reader_sync: a reader called from INSIDE fields::step, where the device copy is newer
  // although `state` only says so when the step ends — the legacy flux_vol planes)
  const bool forced = force_reader_sync;
  if (state != DEVICE_NEWER && !forced) return;
  fields *f = self;
  bool ok = region_fraction_limit > 0 && f->S.multiplicity() == 1 && f->gv.dim == D3 && where.dim == D3;
  LOOP_OVER_DIRECTIONS(f->gv.dim, d) {
    if (f->boundaries[High][d] == Periodic || f->boundaries[Low][d] == Periodic) ok = false;
  }
  if (!ok) {
    if (forced) {
      // (several loop_in_chunks calls of one flux update: nothing runs on the device in between)
      if (!forced_download_done) download_fields();
      forced_download_done = true;
    }
    else sync_host();
    return;
  }
  struct Box {
    realnum *host;
    int64_t rows, row_elems, lo[3], cnt[3];
  };
  std::vector<Box> boxes;
  double part = 0, total = 0;
  const ivec lo_i = f->gv.round_vec(where.get_min_corner()), hi_i = f->gv.round_vec(where.get_max_corner());
  for (int i = 0; i < f->num_chunks; ++i) {
    fields_chunk *fc = f->chunks[i];
    if (!fc->is_mine()) continue;
    const grid_volume &gv = fc->gv;
    const direction ds[3] = {X, Y, Z};
    int64_t lo[3], cnt[3];
    bool empty = false;
    for (int k = 0; k < 3; ++k) {
      // array index range along direction k touched by a reader of `where`: 2 pixels of margin for
      // interpolation weights and centred-grid averages
      const int io = gv.little_corner().in_direction(ds[k]);
      int a = (lo_i.in_direction(ds[k]) - io) / 2 - 3, b = (hi_i.in_direction(ds[k]) - io) / 2 + 3;
      if (a < 0) a = 0;
      if (b > gv.num_direction(ds[k])) b = gv.num_direction(ds[k]);
      if (b < a) empty = true;
      lo[k] = a;
      cnt[k] = b - a + 1;
    }
    const double ntot = (double)gv.ntot();
    FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) {
      realnum *p = fc->f[c][cmp];
      if (!p) continue;
      if (is_magnetic(c) && p == fc->f[direction_component(Bx, component_direction(c))][cmp]) continue; // H aliases B
      total += ntot;
      if (empty) continue;
      part += (double)cnt[0] * cnt[1] * cnt[2];
      boxes.push_back(Box{p, gv.ny() + 1, gv.nz() + 1, {lo[0], lo[1], lo[2]}, {cnt[0], cnt[1], cnt[2]}});
    }
  }
  if (total == 0 || part > region_fraction_limit * total) {
    if (forced) {
      // (several loop_in_chunks calls of one flux update: nothing runs on the device in between)
      if (!forced_download_done) download_fields();
      forced_download_done = true;
    }
    else sync_host();
    return;
  }
  refresh_deferred_halos();
  for (const Box &b : boxes) {
    check(mb200_d2h_box(ctx, b.host, dev(b.host), sizeof(realnum), b.rows, b.row_elems, b.lo[0], b.lo[1], b.lo[2],
                        b.cnt[0], b.cnt[1], b.cnt[2]),
          "mb200_d2h_box");
    stats.d2h_bytes += (double)b.cnt[0] * b.cnt[1] * b.cnt[2] * sizeof(realnum);
  }
  check(mb200_sync(ctx), "device synchronisation");
  host_resident = true;
  stats.region_downloads++;
  if (verbose && stats.region_downloads <= 8)
    fprintf(stderr, "meep_b200: sub-volume download: %zu boxes, %.3f MB (%.2f %% of the field arrays)\n",
            boxes.size(), part * sizeof(realnum) / 1e6, 100.0 * part / total);
  // (state stays DEVICE_NEWER: only the box is current on the host)
}

void Engine::enter(fields *f) {
  if (depth++ > 0) return;
  current_ = this;
  // a stand-alone entry point (reference/user code running a piece of the schedule by itself) may
  // read or propagate not-owned D/B values: bring them up to date first
  if (!in_step && state == DEVICE_NEWER) refresh_deferred_halos();
  const uint64_t fp = fingerprint(f);
  const double t0 = verbose ? meep::wall_time() : 0;
  if (fp != last_fingerprint || arrs_.empty()) {
    // arrays that are new to us are uploaded from the host by ensure(); existing mirrors keep
    // their (possibly newer) device contents
    scan(f);
    last_fingerprint = fp;
    invalidate_plans();
    materials_dirty = true;
    if (verbose) {
      mb200_sync(ctx);
      fprintf(stderr, "meep_b200: mirror scan (%zu arrays, %.2f GB on device): %.3f s\n", arrs_.size(),
              mb200_bytes_allocated(ctx) / 1e9, meep::wall_time() - t0);
    }
  }
  const double t1 = verbose ? meep::wall_time() : 0;
  const bool up_mat = materials_dirty, up_fld = state == HOST_NEWER;
  if (materials_dirty) upload_materials();
  if (state == HOST_NEWER) {
    upload_fields();
    state = COHERENT;
  }
  for (auto &kv : arrs_)
    kv.second.fresh = false;
  if (verbose && (up_mat || up_fld)) {
    mb200_sync(ctx);
    fprintf(stderr, "meep_b200: upload (materials %d, fields %d): %.3f s\n", (int)up_mat, (int)up_fld,
            meep::wall_time() - t1);
  }
}

void Engine::leave(fields *f, bool modified) {
  if (--depth > 0) return;
  const double t_leave = verbose ? meep::wall_time() : 0;
  if (modified) state = DEVICE_NEWER;
  if (in_step && !cw_mode && !eager) release_host_fields();
  // lazily allocated arrays changed the pointer set: remember the new fingerprint so the next
  // step does not rescan
  last_fingerprint = fingerprint(f);
  current_ = nullptr;
  // solve_cw (src/cw_fields.cpp:25-73) gathers / scatters the host arrays directly around every
  // fields::step(): in that mode the host copy is the master between calls
  // (cw_mode is set by fields::step)
  if ((!in_step && !keep_on_device) || cw_mode) {
    // The caller is reference/user code running a piece of the schedule by itself (e.g.
    // synchronize_magnetic_fields, initialize_field): it works on the host arrays right
    // before and after this call, so hand them back and assume it will modify them.
    sync_host();
    state = HOST_NEWER;
    if (getenv("MEEP_B200_TRACE_STANDALONE")) {
      static int shown = 0;
      if (shown++ < 24) {
        void *bt[12];
        const int n = backtrace(bt, 12);
        fprintf(stderr, "meep_b200: stand-alone entry point (host copy handed back):\n");
        backtrace_symbols_fd(bt, n, 2);
      }
    }
  }
  if (verbose && meep::wall_time() - t_leave > 0.05)
    fprintf(stderr, "meep_b200: leaving an entry point took %.3f s on the host\n", meep::wall_time() - t_leave);
}

// ---- finiteness probe (stands in for the host read of src/step.cpp:137-138) ----------------------

void Engine::setup_probe(fields *f) {
  if (probe_ptrs_) return;
  std::vector<uint64_t> ptrs;
  const ivec centre = f->gv.round_vec(f->gv.center());
  for (int i = 0; i < f->num_chunks; ++i) {
    fields_chunk *fc = f->chunks[i];
    if (!fc->is_mine()) continue;
    FOR_COMPONENTS(c) {
      if (!fc->f[c][0] || !(is_D(c) || is_B(c) || is_electric(c) || is_magnetic(c))) continue;
      // nearest Yee point of component c at or above the cell centre, if this chunk holds it
      ivec p = centre;
      LOOP_OVER_DIRECTIONS(f->gv.dim, d) {
        int v = p.in_direction(d), par = fc->gv.iyee_shift(c).in_direction(d) & 1;
        if ((v & 1) != par) v += 1;
        p.set_direction(d, v);
      }
      if (!fc->gv.owns(p)) continue;
      const ptrdiff_t idx = fc->gv.index(c, p);
      for (int cmp = 0; cmp < 2; ++cmp)
        if (fc->f[c][cmp]) ptrs.push_back(dev_addr(fc->f[c][cmp] + idx));
    }
  }
  probe_n_ = (int64_t)ptrs.size();
  if (!probe_n_) return;
  check(mb200_malloc(ctx, ptrs.size() * 8, &probe_ptrs_), "malloc(probe)");
  check(mb200_h2d(ctx, probe_ptrs_, ptrs.data(), ptrs.size() * 8), "h2d(probe)");
  if (!probe_flag_) {
    void *p = nullptr;
    check(mb200_malloc(ctx, 8, &p), "malloc(flag)");
    probe_flag_ = (int32_t *)p;
    check(mb200_memset(ctx, probe_flag_, 0, 8), "memset(flag)");
  }
}

void Engine::check_probe(fields *f, bool force) {
  setup_probe(f);
  if (!probe_n_) return;
  check(mb200_check_finite(ctx, dtype, (const uint64_t *)probe_ptrs_, probe_n_, probe_flag_),
        "check_finite");
  if (force || (stats.steps % nan_check_every) == 0) {
    int32_t flag = 0;
    check(mb200_d2h(ctx, &flag, probe_flag_, 4), "d2h(flag)");
    stats.d2h_bytes += 4;
    if (flag) meep::abort("simulation fields are NaN or Inf");
  }
}

// ---- phases --------------------------------------------------------------------------------------

static mb200_plan *make_plan(Engine &E, int kind, const void *jobs, size_t n) {
  if (!n) return nullptr;
  mb200_plan *p = nullptr;
  check(mb200_plan_create(E.ctx, kind, E.dtype, jobs, (int)n, &p), "mb200_plan_create");
  E.stats.plan_builds++;
  return p;
}

static void push(Phase &ph, int kind, mb200_plan *p) {
  if (!p) return;
  Launch l;
  l.kind = kind;
  l.plan = p;
  ph.launches.push_back(l);
}

// try to turn the three step_curl jobs of one chunk/cmp into fused jobs (one per x-slab)
static bool fuse_group(const Recorder &R, const Recorder::Group &g,
                       std::vector<mb200_step3_job_t> &out_jobs) {
  const fields_chunk *fc = g.fc;
  if (fc->gv.dim != D3 || g.count != 3) return false;
  mb200_step3_job_t out;
  memset(&out, 0, sizeof(out));
  const direction dirs[3] = {X, Y, Z};
  for (int d = 0; d < 3; ++d) {
    out.n[d] = fc->gv.num_direction(dirs[d]);
    out.stride[d] = fc->gv.stride(dirs[d]);
  }
  out.dt = R.curl[g.first].dt;
  out.ix_lo = 0;
  out.ix_hi = out.n[0];
  out.noepi_lo = 0;
  out.noepi_n = 0;
  for (int c = 0; c < 3; ++c) {
    const mb200_curl_job_t &J = R.curl[g.first + c];
    if (!J.g1 || !J.g2) return false;
    mb200_step3_comp_t &C = out.c[c];
    // box -> inclusive index ranges (3-D: loops 1,2,3 are X,Y,Z)
    int64_t rem = J.box.idx0;
    for (int d = 0; d < 3; ++d) {
      if (J.box.s[d] != out.stride[d]) return false;
      C.lo[d] = (int)(rem / out.stride[d]);
      rem -= (int64_t)C.lo[d] * out.stride[d];
      C.hi[d] = C.lo[d] + J.box.n[d] - 1;
    }
    C.f = J.f;
    C.g1 = J.g1;
    C.g2 = J.g2;
    C.s1 = J.s1;
    C.s2 = J.s2;
    C.dtdx = J.dtdx;
    C.pml = J.pml;
    C.pmlu = J.pmlu;
    // re-base PML lookups from loop indices to array indices
    for (int d = 0; d < 3; ++d) {
      C.pml.k0 -= C.pml.ks[d] * C.lo[d];
      C.pmlu.k0 -= C.pmlu.ks[d] * C.lo[d];
    }
    C.fu = J.fu;
    C.cnd = J.cnd;
    C.cndinv = J.cndinv;
    C.fcnd = J.fcnd;
    C.e = nullptr;
    for (int d = 0; d < 3; ++d)
      C.metal_lo[d] = C.metal_hi[d] = -1;
  }
  if (!g.fuse_eh) {
    out_jobs.push_back(out);
    return true;
  }
  mb200_step3_job_t fused = out;
  for (int c = 0; c < 3; ++c) {
    fused.c[c].e = g.epi[c].e;
    fused.c[c].u = g.epi[c].u;
    fused.c[c].fw = g.epi[c].fw;
    fused.c[c].pmlw = g.epi[c].pmlw;
    for (int d = 0; d < 3; ++d) {
      fused.c[c].metal_lo[d] = g.epi[c].metal_lo[d];
      fused.c[c].metal_hi[d] = g.epi[c].metal_hi[d];
    }
  }
  // planes that hold source points are updated without the epilogue (step_source runs between the
  // D and the E update there: src/step.cpp:98-109; update_eh covers them afterwards): same job
  fused.noepi_lo = g.slab_lo;
  fused.noepi_n = g.slab_hi - g.slab_lo + 1;
  out_jobs.push_back(fused);
  return true;
}

void Engine::end_record(Phase &ph, PhaseId id, fields *f) {
  (void)f;
  Recorder &R = rec_;
  free_phase(ph);
  switch (id) {
    case PH_DB: {
      std::vector<mb200_step3_job_t> s3;
      std::vector<mb200_curl_job_t> rest;
      std::vector<char> used(R.curl.size(), 0);
      if (fuse)
        for (const Recorder::Group &g : R.curl_groups) {
          if (fuse_group(R, g, s3)) {
            for (int k = 0; k < g.count; ++k)
              used[g.first + k] = 1;
          }
          else if (g.fuse_eh)
            meep::abort("meep_b200: internal error: E/H fusion promised for a chunk that cannot "
                        "use the fused kernel");
        }
      for (size_t k = 0; k < R.curl.size(); ++k)
        if (!used[k]) rest.push_back(R.curl[k]);
      // two launches: the fast-path kernel for plain (non-PML) chunks, the general one for the rest
      std::vector<mb200_step3_job_t> s3_plain, s3_gen;
      for (const mb200_step3_job_t &j : s3) {
        bool plain = true;
        for (int c = 0; c < 3; ++c) {
          const mb200_step3_comp_t &C = j.c[c];
          if (!C.f || C.pml.sig || C.pmlu.sig || C.cnd || !C.g2 || (C.e && C.pmlw.sig)) plain = false;
        }
        if (!plain) {
          mb200_step3_job_t g = j;
          g.reserved = pml_t1; // thin PML slabs: shorter marches, more CTAs in flight
          s3_gen.push_back(g);
        }
        else {
          mb200_step3_job_t g = j;
          if (plain_t1 > 0) g.reserved = plain_t1;
          s3_plain.push_back(g);
        }
      }
      push(ph, MB200_K_CYLINT, make_plan(*this, MB200_K_CYLINT, R.cylint.data(), R.cylint.size()));
      push(ph, MB200_K_STEP3, make_plan(*this, MB200_K_STEP3, s3_plain.data(), s3_plain.size()));
      push(ph, MB200_K_STEP3, make_plan(*this, MB200_K_STEP3, s3_gen.data(), s3_gen.size()));
      push(ph, MB200_K_CURL, make_plan(*this, MB200_K_CURL, rest.data(), rest.size()));
      push(ph, MB200_K_BFAST, make_plan(*this, MB200_K_BFAST, R.bfast.data(), R.bfast.size()));
      push(ph, MB200_K_BETA, make_plan(*this, MB200_K_BETA, R.beta.data(), R.beta.size()));
      push(ph, MB200_K_CYLR0, make_plan(*this, MB200_K_CYLR0, R.cylr0.data(), R.cylr0.size()));
      push(ph, MB200_K_ZERO, make_plan(*this, MB200_K_ZERO, R.cylzero.data(), R.cylzero.size()));
      break;
    }
    case PH_SRC: {
      mb200_plan *p = make_plan(*this, MB200_K_SOURCE, R.src.data(), R.src.size());
      if (p) {
        Launch l;
        l.kind = MB200_K_SOURCE;
        l.plan = p;
        l.src_times = R.src_times;
        l.src_dipole = false;
        ph.launches.push_back(l);
      }
      break;
    }
    case PH_BND:
      push(ph, MB200_K_ZERO, make_plan(*this, MB200_K_ZERO, R.zero.data(), R.zero.size()));
      if (!R.links.empty()) { // peer-memory mode: the pack jobs store straight into the neighbours' HBM
        // What the neighbours wait for goes first: pack (a few MB over NVLink), signal, and only
        // then the same-device copies — the neighbours' blocks arrive while those run, so the wait
        // before the unpack has the NVLink latency and most of the skew between GPUs behind it.
        ph.links = R.links;
        Launch pre, sig, wait, post;
        pre.kind = KIND_P2P_PRE;
        sig.kind = KIND_P2P_SIGNAL;
        wait.kind = KIND_P2P_WAIT;
        post.kind = KIND_P2P_POST;
        ph.launches.push_back(pre);
        push(ph, MB200_K_HALO, make_plan(*this, MB200_K_HALO, R.pack.data(), R.pack.size()));
        ph.launches.push_back(sig);
        push(ph, MB200_K_HALO, make_plan(*this, MB200_K_HALO, R.halo.data(), R.halo.size()));
        ph.launches.push_back(wait);
        push(ph, MB200_K_HALO, make_plan(*this, MB200_K_HALO, R.unpack.data(), R.unpack.size()));
        ph.launches.push_back(post);
        break;
      }
      push(ph, MB200_K_HALO, make_plan(*this, MB200_K_HALO, R.pack.data(), R.pack.size()));
      push(ph, MB200_K_HALO, make_plan(*this, MB200_K_HALO, R.halo.data(), R.halo.size()));
      if (!R.sends.empty() || !R.recvs.empty()) {
        Launch l;
        l.kind = KIND_EXCHANGE;
        l.sends = R.sends;
        l.recvs = R.recvs;
        ph.launches.push_back(l);
      }
      push(ph, MB200_K_HALO, make_plan(*this, MB200_K_HALO, R.unpack.data(), R.unpack.size()));
      break;
    case PH_EH: {
      push(ph, MB200_K_FMP, make_plan(*this, MB200_K_FMP, R.fmp.data(), R.fmp.size()));
      mb200_plan *p = make_plan(*this, MB200_K_SOURCE, R.dip.data(), R.dip.size());
      if (p) {
        Launch l;
        l.kind = MB200_K_SOURCE;
        l.plan = p;
        l.src_times = R.dip_times;
        l.src_dipole = true;
        ph.launches.push_back(l);
      }
      push(ph, MB200_K_EDHB, make_plan(*this, MB200_K_EDHB, R.edhb.data(), R.edhb.size()));
      break;
    }
    case PH_POLS: {
      // zero-block jobs and plain jobs use different kernels (one plan each)
      std::vector<mb200_lorentz_job_t> blocked, plainj;
      for (const mb200_lorentz_job_t &j : R.lorentz)
        (j.pzero ? blocked : plainj).push_back(j);
      push(ph, MB200_K_LORENTZ, make_plan(*this, MB200_K_LORENTZ, blocked.data(), blocked.size()));
      push(ph, MB200_K_LORENTZ, make_plan(*this, MB200_K_LORENTZ, plainj.data(), plainj.size()));
      push(ph, MB200_K_GYRO, make_plan(*this, MB200_K_GYRO, R.gyro.data(), R.gyro.size()));
      if (mb200_plan *p = make_plan(*this, MB200_K_NOISE, R.noise.data(), R.noise.size())) {
        Launch l;
        l.kind = MB200_K_NOISE;
        l.plan = p;
        l.noise_gens = R.noise_gens;
        ph.launches.push_back(l);
      }
      break;
    }
    case PH_DFT:
      for (auto &kv : R.dft) {
        mb200_plan *p = make_plan(*this, MB200_K_DFT, kv.second.data(), kv.second.size());
        if (!p) continue;
        Launch l;
        l.kind = MB200_K_DFT;
        l.plan = p;
        l.dft_chunks = R.dft_chunks[kv.first];
        l.decimation = kv.first;
        ph.launches.push_back(l);
      }
      break;
    default: break;
  }
  if (verbose) {
    static const char *names[] = {"step_db", "step_source", "step_boundaries", "update_eh",
                                  "update_pols", "update_dfts"};
    fprintf(stderr, "meep_b200: recorded %s:", names[id]);
    for (const Launch &l : ph.launches)
      if (l.kind >= KIND_P2P_PRE)
        fprintf(stderr, " [p2p %s]", l.kind == KIND_P2P_PRE ? "wait consumed" : l.kind == KIND_P2P_SIGNAL ? "signal packed"
                                     : l.kind == KIND_P2P_WAIT ? "wait packed" : "signal consumed");
      else if (l.kind == KIND_EXCHANGE)
        fprintf(stderr, " [exchange: %zu sends, %zu recvs]", l.sends.size(), l.recvs.size());
      else
        fprintf(stderr, " [kind %d: %.0f points, %.3f MB]", l.kind, mb200_plan_points(l.plan),
                mb200_plan_bytes(l.plan) / 1e6);
    fprintf(stderr, "\n");
  }
  ph.aux.swap(rec_aux_);
  rec_aux_.clear();
  ph.valid = true;
  rec_ = Recorder();
  recording_ = false;
  if (pending_invalidate_) { // arrays appeared while recording: every OTHER phase is stale,
    pending_invalidate_ = false; // and this one must be re-recorded after it has run once
    ph.one_shot = true;
    for (int i = 0; i < PH_COUNT; ++i)
      for (int ft = 0; ft < NUM_FIELD_TYPES; ++ft)
        if (&phases_[i][ft] != &ph) free_phase(phases_[i][ft]);
    if (probe_ptrs_) {
      mb200_free(ctx, probe_ptrs_);
      probe_ptrs_ = nullptr;
      probe_n_ = 0;
    }
  }
}

void Engine::run(Phase &ph, fields *f) {
  for (Launch &l : ph.launches) {
    if (l.kind == KIND_P2P_PRE || l.kind == KIND_P2P_SIGNAL || l.kind == KIND_P2P_WAIT || l.kind == KIND_P2P_POST) {
      // all sequence words of this phase in one launch (a 2x2x2 leaf has up to 7 neighbours)
      std::vector<uint64_t *> flags;
      std::vector<uint64_t> values;
      for (int k : ph.links) {
        P2PLink &p = links[k];
        if (l.kind == KIND_P2P_PRE) {
          // what I stored into the neighbour's arena last time must have been consumed
          p.seq += 1;
          if (p.send_count) {
            flags.push_back((uint64_t *)((char *)p.mine + 8));
            values.push_back(p.seq - 1);
          }
        }
        else if (l.kind == KIND_P2P_SIGNAL) {
          if (p.send_count) {
            flags.push_back((uint64_t *)p.theirs);
            values.push_back(p.seq);
          }
        }
        else if (l.kind == KIND_P2P_WAIT) {
          if (p.recv_count) {
            flags.push_back((uint64_t *)p.mine);
            values.push_back(p.seq);
          }
        }
        else if (p.recv_count) {
          flags.push_back((uint64_t *)((char *)p.theirs + 8));
          values.push_back(p.seq);
        }
      }
      const bool waiting = l.kind == KIND_P2P_PRE || l.kind == KIND_P2P_WAIT;
      for (size_t k0 = 0; k0 < flags.size(); k0 += MB200_MAX_FLAGS) {
        const int n = (int)std::min<size_t>(MB200_MAX_FLAGS, flags.size() - k0);
        if (waiting)
          check(mb200_flag_wait_many(ctx, (const uint64_t *const *)(flags.data() + k0), values.data() + k0, n),
                "flag_wait");
        else
          check(mb200_flag_signal_many(ctx, flags.data() + k0, values.data() + k0, n), "flag_signal");
      }
      continue;
    }
    if (l.kind == KIND_EXCHANGE) {
      if (emulated) {
        // emulator: "device" buffers are host memory; move them through the socket runtime
        const size_t R = sizeof(realnum);
        std::vector<HostMsg> s, r;
        for (const mb200_xfer_t &x : l.sends)
          s.push_back(HostMsg{x.peer, x.buf, (size_t)x.count * R});
        for (const mb200_xfer_t &x : l.recvs)
          r.push_back(HostMsg{x.peer, x.buf, (size_t)x.count * R});
        comm_sendrecv_all(s, r);
      }
      else {
        ensure_comm();
        check(mb200_comm_exchange(ctx, comm, dtype, l.sends.data(), (int)l.sends.size(),
                                  l.recvs.data(), (int)l.recvs.size()),
              "mb200_comm_exchange");
      }
      continue;
    }
    if (l.kind == MB200_K_NOISE) {
      // the reference's generator, in the reference's order (chunk, susceptibility, component,
      // loop point): src/susceptibility.cpp:326-337
      std::vector<double> noise;
      for (const NoiseGen &g : l.noise_gens)
        for (int i1 = 0; i1 < g.box.n[0]; ++i1)
          for (int i2 = 0; i2 < g.box.n[1]; ++i2)
            for (int i3 = 0; i3 < g.box.n[2]; ++i3) {
              const int64_t i = g.box.idx0 + i1 * g.box.s[0] + i2 * g.box.s[1] + i3 * g.box.s[2];
              noise.push_back(meep::gaussian_random(0, (realnum)g.amp * sqrt(g.sigma[i])));
            }
      check(mb200_plan_run(ctx, l.plan, noise.data(), noise.size() * sizeof(double)), "run(noise)");
      stats.h2d_bytes += noise.size() * sizeof(double);
      continue;
    }
    if (l.kind == MB200_K_SOURCE) {
      std::vector<double> scal(2 * l.src_times.size());
      for (size_t k = 0; k < l.src_times.size(); ++k) {
        const std::complex<double> v =
            l.src_dipole ? l.src_times[k]->dipole() : l.src_times[k]->current();
        scal[2 * k] = v.real();
        scal[2 * k + 1] = v.imag();
      }
      check(mb200_plan_run(ctx, l.plan, scal.data(), scal.size() * sizeof(double)), "run(source)");
      stats.h2d_bytes += scal.size() * sizeof(double);
    }
    else if (l.kind == MB200_K_DFT) {
      if (f->t % l.decimation != 0) continue;
      // phase tables: dft_phase[i] = polar(1, omega_i * t) * scale, computed in double on the
      // host and narrowed to complex<realnum> (reference src/dft.cpp:270-271)
      std::vector<std::complex<realnum> > ph_tab;
      const double timeE = f->time(), timeH = f->time() - 0.5 * f->dt;
      for (dft_chunk *d : l.dft_chunks) {
        const double tm = is_H_or_B(d->c) ? timeH : timeE;
        for (size_t i = 0; i < d->omega.size(); ++i) {
          d->dft_phase[i] = std::polar(1.0, d->omega[i] * tm) * d->scale;
          ph_tab.push_back(d->dft_phase[i]);
        }
      }
      check(mb200_plan_run(ctx, l.plan, ph_tab.data(),
                           ph_tab.size() * sizeof(std::complex<realnum>)),
            "run(dft)");
      stats.h2d_bytes += ph_tab.size() * sizeof(std::complex<realnum>);
    }
    else
      check(mb200_plan_run(ctx, l.plan, nullptr, 0), "mb200_plan_run");
  }
}

} // namespace meep_b200
