// engine.hpp — device mirror + plan cache behind the unchanged meep::fields API.
//
// One Engine per meep::fields object (side table keyed by the fields pointer; the reference's
// meep.hpp is compiled unmodified, so no member can be added).  The Engine
//   * mirrors every per-chunk array the hot path touches (fields_chunk::f, f_u, f_w, f_cond,
//     f_minus_p, polarisation P/P_prev, dft_chunk::dft; structure_chunk::chi1inv, conductivity,
//     condinv, chi2, chi3, sig/kap/siginv, susceptibility sigma) in HBM, in the reference's own
//     array layout, so every index/stride/pointer offset the host code computes is valid on
//     the device as is;
//   * records, per phase of fields::step(), the jobs our replacement member functions emit,
//     turns them into plans (include/meep_b200.h) and replays the plans every step until the
//     chunk layout changes;
//   * tracks which side (host/device) holds the current field values.
#ifndef MEEP_B200_ENGINE_HPP
#define MEEP_B200_ENGINE_HPP

#include <complex>
#include <cstdint>
#include <functional>
#include <map>
#include <typeinfo>
#include <utility>
#include <stddef.h>
#include <string>
#include <vector>

#include "meep.hpp"
#include "../../include/meep_b200.h"

namespace meep_b200 {

using meep::realnum;

enum PhaseId {
  PH_DB = 0,     // step_db(ft)
  PH_SRC,        // step_source(ft)
  PH_BND,        // step_boundaries(ft)
  PH_EH,         // update_eh(ft)
  PH_POLS,       // update_pols(ft)
  PH_DFT,        // update_dfts
  PH_COUNT
};

// one noisy polarisation array: noise[k] = gaussian_random(0, amp * sqrt(sigma[idx(k)])) over the
// loop box, drawn on the host before every launch (reference src/susceptibility.cpp:331-334)
struct NoiseGen {
  double amp;
  const meep::realnum *sigma; // host array
  mb200_box_t box;
};

// a recorded launch: plan + how to build its per-run side data
struct Launch {
  int kind = -1;
  mb200_plan *plan = nullptr;
  // NOISE: the generators whose numbers are concatenated at run time
  std::vector<NoiseGen> noise_gens;
  // SOURCE: src_time objects whose current()/dipole() fill the scalar slots at run time
  std::vector<const meep::src_time *> src_times;
  bool src_dipole = false;
  // DFT: the dft_chunks whose phase tables are concatenated at run time
  std::vector<meep::dft_chunk *> dft_chunks;
  int decimation = 1;
  // EXCHANGE (kind == KIND_EXCHANGE): comm blocks sent to / received from other processes
  std::vector<mb200_xfer_t> sends, recvs;
};
enum { KIND_EXCHANGE = 100, KIND_P2P_PRE = 101, KIND_P2P_SIGNAL = 102, KIND_P2P_POST = 103, KIND_P2P_WAIT = 104 };

// Peer-memory exchange: one link per (field type, neighbour process).  `mine` lives in MY HBM and
// is written by the neighbour's pack kernel over NVLink; `theirs` is the neighbour's arena for me
// (CUDA IPC mapping).  Arena layout: 64-byte header ([0] "packed" sequence word, [8] "consumed"
// sequence word, both written by the OTHER side), then the comm blocks of every chunk pair
// between the two processes in the global pair order.
struct P2PLink {
  int ft = 0, rank = -1;
  void *mine = nullptr, *theirs = nullptr;
  size_t recv_count = 0, send_count = 0; // realnums
  uint64_t seq = 0;                      // exchanges done over this link
};
constexpr size_t kArenaHeader = 64;

struct Phase {
  bool valid = false;
  bool one_shot = false; // recorded while the array set was changing: run once, then re-record
  std::vector<Launch> launches;
  std::vector<void *> aux; // device side tables (index lists, ...) owned by this phase
  std::vector<int> links; // peer-memory exchange: indices into Engine::links (empty: none / NCCL)
};

// job recorder used while a phase is being (re)built
struct Recorder {
  std::vector<mb200_curl_job_t> curl;
  std::vector<mb200_beta_job_t> beta; // 2-D exp(i beta z) terms, run after the curl jobs
  // cylindrical coordinates: helper arrays (before the curl jobs), r = 0 rows and zeroed rows (after)
  // noise terms of noisy_lorentzian_susceptibility: device jobs + what the host needs to draw the
  // numbers each step (amp, host sigma array, loop box), in the reference's order
  std::vector<mb200_noise_job_t> noise;
  std::vector<NoiseGen> noise_gens;
  std::vector<mb200_gyro_job_t> gyro;   // gyrotropic polarisations (update_pols phase)
  std::vector<mb200_bfast_job_t> bfast; // BFAST corrections, after the curl jobs
  std::vector<mb200_cylint_job_t> cylint;
  std::vector<mb200_cylr0_job_t> cylr0;
  std::vector<mb200_zero_job_t> cylzero;
  std::vector<mb200_edhb_job_t> edhb;
  std::vector<mb200_lorentz_job_t> lorentz;
  std::vector<mb200_fmp_job_t> fmp;
  std::vector<mb200_src_job_t> src;       // mode 0 (currents)
  std::vector<const meep::src_time *> src_times;
  std::vector<mb200_src_job_t> dip;       // mode 1 (integrated dipoles)
  std::vector<const meep::src_time *> dip_times;
  std::vector<mb200_halo_job_t> halo;   // same-process pairs
  std::vector<mb200_halo_job_t> pack;   // packing of outgoing comm blocks (into the neighbour's HBM / a send buffer)
  std::vector<mb200_halo_job_t> unpack; // scatter of received comm blocks
  std::vector<mb200_xfer_t> sends, recvs;
  std::vector<int> links;
  std::vector<mb200_zero_job_t> zero;
  std::map<int, std::vector<mb200_dft_job_t> > dft; // by decimation factor
  std::map<int, std::vector<meep::dft_chunk *> > dft_chunks;
  // grouping info for the fused kernel: curl jobs [first, first+count) belong to one chunk/cmp
  struct Epilogue { // fused update_eh of one component (valid when e != NULL)
    void *e = nullptr;
    const void *u = nullptr;
    void *fw = nullptr;
    mb200_pml_t pmlw; // re-based to array indices
    int metal_lo[3] = {-1, -1, -1}, metal_hi[3] = {-1, -1, -1};
  };
  struct Group {
    int first, count;
    const meep::fields_chunk *fc;
    int cmp;
    bool fuse_eh = false;       // the E/H update of this chunk is folded into the D/B pass
    int slab_lo = 1, slab_hi = 0; // planes (array index along direction 0) that hold source points
    Epilogue epi[3];
  };
  std::vector<Group> curl_groups;
};

struct Stats {
  int64_t steps = 0;
  int64_t uploads = 0, downloads = 0;
  double h2d_bytes = 0, d2h_bytes = 0;
  int64_t plan_builds = 0;
  int64_t region_downloads = 0;
};

class Engine {
public:
  static Engine &get(meep::fields *f);          // create on first use
  static Engine *find(const meep::fields *f);   // NULL if none
  static void drop(const meep::fields *f);      // fields destroyed
  static Engine *current() { return current_; }
  static uint64_t source_generation; // bumped by the interposed fields_chunk::add_source
  // the Engine whose mirror table contains host address p (NULL if none)
  static Engine *owner_of(const void *p);
  bool mirrors(const void *p) const;
  static void for_each(const std::function<void(Engine &)> &fn);

  explicit Engine(meep::fields *f);
  ~Engine();

  // ---- coherence ------------------------------------------------------------------------------
  enum State { HOST_NEWER, COHERENT, DEVICE_NEWER };
  State state = HOST_NEWER;
  int depth = 0; // nesting of our own entry points (0 => called from reference/user code)
  bool in_step = false; // the outermost entry point is fields::step()
  bool merge_exchanges = true; // MEEP_B200_MERGE_EXCHANGES=0: one exchange per field type
  bool defer_known = false;
  bool defer_ok[meep::NUM_FIELD_TYPES] = {}; // global decision: D/B connections may be merged
  bool deferred_exchange[meep::NUM_FIELD_TYPES] = {}; // D/B connections ride with the E/H ones
  bool connections_valid = false; // fields::chunk_connections_valid at the last step_db
  // same-device D/B halo copies postponed until somebody can look (see fields::step_boundaries)
  bool defer_local = true;        // MEEP_B200_DEFER_LOCAL=0: carry them out every step
  bool local_deferred[meep::NUM_FIELD_TYPES] = {}; // the cached E/H exchange plan leaves them out
  bool halos_stale[meep::NUM_FIELD_TYPES] = {};    // steps were taken since the last refresh
  bool refresh_local = false;     // inside refresh_deferred_halos
  meep::fields *self = nullptr;
  void refresh_deferred_halos();

  // Brackets every interposed entry point.  On the outermost entry it validates the mirror
  // (array set, materials) and uploads field arrays if the host copy is (or may be) newer.
  void enter(meep::fields *f);
  void leave(meep::fields *f, bool modified_fields);
  // make the host arrays current (called by the interposed readers in hooks.cpp)
  void sync_host();
  // make current only what a reader of the sub-volume `where` can touch (the field arrays of the
  // chunks it overlaps, restricted to the index box of `where` + 2 pixels); falls back to
  // sync_host() when images of the volume could be read (symmetries, periodic boundaries) or the
  // box is most of the cell.  The device copy stays the authoritative one.
  void sync_host_region(const meep::volume &where);
  bool halo_sort = true; // MEEP_B200_HALO_SORT=0: exchange jobs in chunk-pair order
  bool force_reader_sync = false;
  bool forced_download_done = false; // the arrays were all downloaded since force_reader_sync was raised
  double region_fraction_limit = 0.3; // MEEP_B200_REGION_SYNC (0 disables sub-volume downloads)
  void mark_host_dirty() { if (state != DEVICE_NEWER) state = HOST_NEWER; }

  // ---- mirror ---------------------------------------------------------------------------------
  struct Arr {
    void *dev = nullptr;
    size_t bytes = 0;
    bool is_field = false; // true: device-authoritative between steps; false: material (host)
    bool seen = false;
    bool fresh = false; // uploaded by ensure() during the current enter()
  };
  // device address of a host element pointer (NULL -> NULL); aborts if the array is unknown
  void *dev(const void *host) const;
  uint64_t dev_addr(const void *host) const { return (uint64_t)(uintptr_t)dev(host); }
  // register (or find) a mirror; init: 0 = upload host contents, 1 = leave uninitialised
  void *ensure(const void *host, size_t bytes, bool is_field, int init = 0);
  // lazily allocated array created inside a step: host array `host` was just new[]'d;
  // its device twin is initialised from device array `src_host` (or zeros if NULL).
  void ensure_from(const void *host, size_t bytes, const void *src_host);
  void forget(const void *host);
  void scan(meep::fields *f);          // (re)register everything reachable from f
  // zero-block flags (include/meep_b200.h: mb200_lorentz_job_t): one byte per MB200_ZBLOCK
  // elements of a polarisation array pair (P, P_prev) / of a susceptibility sigma array
  uint8_t *pzero_flags(const void *host_P, size_t ntot, bool known_zero); // create on first use
  void pzero_drop(const void *host_P);
  uint8_t *pzero_lookup(const void *host_elem) const; // flag array covering a P element (or NULL)
  uint64_t pzero_flag_addr(const void *host_elem) const;
  const uint8_t *szero_flags(const void *host_sigma, size_t ntot);
  bool zero_skip = true; // MEEP_B200_ZERO_SKIP=0 disables
  void upload_array(const void *host, void *dev, size_t bytes);
  void release_host_fields();
  bool lazy_host = true;    // MEEP_B200_LAZY_HOST=0: always copy host arrays, never probe their pages
  int release_host = 1;     // MEEP_B200_RELEASE_HOST (see release_host_fields)
  bool host_resident = false; // field pages may be resident on the host (set by up/downloads)
  int64_t released_at_download = 0;
  void upload_fields();
  void download_fields();
  void upload_materials();

  // ---- phases ---------------------------------------------------------------------------------
  Phase &phase(PhaseId id, int ft) { return phases_[id][ft]; }
  // drop every cached plan.  While a phase is being recorded (a lazily allocated array just
  // appeared) the drop is deferred to end_record and spares the phase being recorded.
  void invalidate_plans();
  void free_phase(Phase &ph);
  void begin_record() { rec_ = Recorder(); recording_ = true; }
  Recorder &rec() { return rec_; }
  bool recording() const { return recording_; }
  void end_record(Phase &ph, PhaseId id, meep::fields *f);
  void run(Phase &ph, meep::fields *f);

  // ---- misc -----------------------------------------------------------------------------------
  mb200_ctx *ctx = nullptr;
  mb200_comm *comm = nullptr; // inter-process exchange (created on first use when WORLD_SIZE > 1)
  bool emulated = false;      // the C ABI is served by the test-only emulator
  int plain_t1 = 16;          // MEEP_B200_PLAIN_T1: same for the fast-path kernel (0: kernel default, 8)
  int pml_t1 = 16;            // MEEP_B200_PML_T1: x-planes marched per CTA in PML chunks
  bool halo_runs = true;      // MEEP_B200_HALO_RUNS=0: plain address lists for every halo job
  // device-side copies made by fields::synchronize_magnetic_fields (host array -> backup buffer)
  std::map<const void *, void *> backups;
  void backup_array(const void *host);
  void restore_array(const void *host);
  bool has_backup(const void *host) const { return backups.count(host) != 0; }
  void drop_backups();
  bool suppress_zero_skip = false; // set while recording the Lorentzian part of a noisy susceptibility
  int keep_on_device = 0;     // > 0: a stand-alone phase call leaves the arrays in HBM (no hand-back)
  bool cw_mode = false;       // inside solve_cw: the host arrays are the master between steps
  bool p2p = true;            // MEEP_B200_P2P=0: move comm blocks with NCCL instead of peer stores
  // peer-memory links (see P2PLink).  (Re)built collectively by the first in-step
  // step_boundaries after the chunks were (re)connected; counts[(ft, peer)] = (send, recv).
  std::vector<P2PLink> links;
  int connect_epoch = 0, links_epoch = -1;
  int plans_epoch = -1; // connect_epoch the cached plans were recorded against
  void rebuild_links(const std::map<std::pair<int, int>, std::pair<size_t, size_t> > &counts);
  void drop_links(); // collective when links exist
  int find_link(int ft, int rank) const;
  void ensure_comm();
  // plain device buffer owned by the phase being recorded
  void *aux_alloc(size_t bytes);
  int dtype = sizeof(realnum) == 8 ? MB200_F64 : MB200_F32;
  Stats stats;
  bool fuse = true;        // MEEP_B200_FUSE=0 disables the fused step3 path
  bool verbose = false;    // MEEP_B200_VERBOSE=1: print the recorded plans
  bool device_timers = false; // MEEP_B200_TIMERS=1 (or verbosity > 1): CUDA-event phase timers in fields::step
  bool eager = false;      // MEEP_B200_EAGER=1: download after every step (debug/safety)
  int nan_check_every = 16;
  // finiteness probe
  void setup_probe(meep::fields *f);
  void check_probe(meep::fields *f, bool force);

  // chunks whose update_eh(ft) was folded into step_db of the matching D/B type in the current
  // plan generation: value = [slab_lo, slab_hi] planes that were NOT fused (source planes;
  // lo > hi: the whole chunk is fused).  Written by step_db, read by update_eh, in-step only.
  std::map<const meep::fields_chunk *, std::pair<int, int> > fused_eh[meep::NUM_FIELD_TYPES];

  uint64_t fingerprint(meep::fields *f) const;
  uint64_t last_fingerprint = 0;
  bool materials_dirty = true;

private:
  std::map<uintptr_t, Arr> arrs_; // keyed by host base address
  mutable uintptr_t hit_lo_ = 0, hit_hi_ = 0; // last array dev() resolved
  mutable char *hit_dev_ = nullptr;
  struct Flags {
    uint8_t *dev = nullptr;
    size_t ntot = 0, nblocks = 0;
    bool fresh = false; // szero: computed since the last material upload
  };
  std::map<uintptr_t, Flags> pzero_, szero_; // keyed by the host address of the array
  std::vector<void *> rec_aux_; // side tables uploaded while recording the current phase
  bool pending_invalidate_ = false;
  Phase phases_[PH_COUNT][meep::NUM_FIELD_TYPES];
  Recorder rec_;
  bool recording_ = false;
  void *probe_ptrs_ = nullptr;
  int64_t probe_n_ = 0;
  int32_t *probe_flag_ = nullptr;
  static Engine *current_;
  friend struct Scope;

public:
  // device side table owned by the phase being recorded
  void *aux_upload(const void *host, size_t bytes);
};

// RAII bracket for interposed entry points
struct Scope {
  Engine &E;
  meep::fields *f;
  bool modified;
  Scope(Engine &e, meep::fields *ff, bool mod = true) : E(e), f(ff), modified(mod) { E.enter(f); }
  ~Scope() { E.leave(f, modified); }
};

void check(int rc, const char *what);

// run one phase: replay the cached plans, or (re)record them by calling `record`
template <typename F>
inline void run_phase(Engine &E, meep::fields *f, PhaseId id, int ft, bool cacheable, F record) {
  if (cacheable) {
    Phase &ph = E.phase(id, ft);
    if (!ph.valid) {
      const double t0 = E.verbose ? meep::wall_time() : 0;
      E.begin_record();
      record();
      E.end_record(ph, id, f);
      if (E.verbose)
        fprintf(stderr, "meep_b200: phase %d/%d recorded in %.3f s (host)\n", (int)id, ft,
                meep::wall_time() - t0);
    }
    E.run(ph, f);
    if (ph.one_shot) E.free_phase(ph);
  }
  else { // one-off (solve_cw variants etc.): record, run, discard
    Phase tmp;
    E.begin_record();
    record();
    E.end_record(tmp, id, f);
    E.run(tmp, f);
    E.free_phase(tmp);
  }
}


// restated layout of the file-local struct in the reference's src/susceptibility.cpp:98-104
// (the block new_internal_data/init_internal_data allocate; ABI between reference and us)
struct lorentzian_data_layout {
  size_t sz_data;
  size_t ntot;
  realnum *P[meep::NUM_FIELD_COMPONENTS][2];
  realnum *P_prev[meep::NUM_FIELD_COMPONENTS][2];
  realnum data[1];
};

// likewise src/susceptibility.cpp:374-380 (gyrotropic_susceptibility)
struct gyrotropy_data_layout {
  size_t sz_data;
  size_t ntot;
  realnum *P[meep::NUM_FIELD_COMPONENTS][2][3];
  realnum *P_prev[meep::NUM_FIELD_COMPONENTS][2][3];
  realnum data[1];
};

// the polarisation block of a susceptibility the device path supports: {start, bytes} of its data
// area (0 bytes: nothing allocated), aborting for the kinds that are not supported
inline std::pair<realnum *, size_t> polarisation_block(const meep::susceptibility *s, void *data) {
  if (!data) return std::make_pair((realnum *)nullptr, (size_t)0);
  if (typeid(*s) == typeid(meep::lorentzian_susceptibility) ||
      typeid(*s) == typeid(meep::noisy_lorentzian_susceptibility)) {
    lorentzian_data_layout *d = (lorentzian_data_layout *)data;
    const size_t hdr = offsetof(lorentzian_data_layout, data);
    return std::make_pair(d->data, d->sz_data > hdr ? d->sz_data - hdr : 0);
  }
  if (typeid(*s) == typeid(meep::gyrotropic_susceptibility)) {
    gyrotropy_data_layout *d = (gyrotropy_data_layout *)data;
    const size_t hdr = offsetof(gyrotropy_data_layout, data);
    return std::make_pair(d->data, d->sz_data > hdr ? d->sz_data - hdr : 0);
  }
  meep::abort("meep_b200: only lorentzian_susceptibility (Lorentz/Drude, also noisy) and "
              "gyrotropic_susceptibility polarisations are supported on the device path");
  return std::make_pair((realnum *)nullptr, (size_t)0);
}

} // namespace meep_b200
#endif
