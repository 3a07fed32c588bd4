// hooks.cpp — keeps the reference's HOST-side readers and writers of field / DFT arrays correct
// while the arrays live in HBM.
//
// libmeep_b200 is linked in front of (or LD_PRELOADed over) the installed libmeep.  The
// definitions below interpose a handful of reference entry points: each one first makes the
// host arrays current (or reads the few values it needs straight from the device), then
// forwards to the reference's own definition found with dlsym(RTLD_NEXT, <its own symbol>).
// Nothing here computes fields.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

#include "engine.hpp"
#include "hostmem.hpp"
#include "meep_internals.hpp"

using namespace std;
using namespace meep_b200;

namespace {

// The reference's definition of the function we were called from: look up the caller's own
// (mangled) symbol name with dladdr and ask the dynamic linker for the next definition.
__attribute__((noinline)) void *next_definition_of_caller() {
  Dl_info info;
  void *ra = __builtin_extract_return_addr(__builtin_return_address(0));
  if (!dladdr(ra, &info) || !info.dli_sname) {
    fprintf(stderr, "meep_b200: cannot resolve the interposed symbol of a hook\n");
    abort();
  }
  void *p = dlsym(RTLD_NEXT, info.dli_sname);
  if (!p) {
    fprintf(stderr,
            "meep_b200: %s is not defined by any library loaded after libmeep_b200 — "
            "libmeep_b200 must be linked/preloaded in FRONT of the reference libmeep\n",
            info.dli_sname);
    abort();
  }
  return p;
}

// every Engine whose fields may be read by the caller
void sync_all_hosts() {
  Engine::for_each([](Engine &E) { E.sync_host(); });
}

} // namespace

namespace meep {

// ---- lifetime -------------------------------------------------------------------------------------
extern "C" {
void mb200_hook_fields_D1(fields *self) __asm__("_ZN4meep6fieldsD1Ev");
void mb200_hook_fields_D2(fields *self) __asm__("_ZN4meep6fieldsD2Ev");
}
void mb200_hook_fields_D1(fields *self) {
  static void (*next)(fields *) = (void (*)(fields *))next_definition_of_caller();
  Engine::drop(self);
  next(self);
}
void mb200_hook_fields_D2(fields *self) {
  static void (*next)(fields *) = (void (*)(fields *))next_definition_of_caller();
  Engine::drop(self);
  next(self);
}

// A DFT monitor removed mid-run (dft_flux::remove() etc. -> delete of its dft_chunks, reference
// src/dft.cpp:131-144) frees the host `dft` array: its device mirror must go with it, or a later
// download would write into freed memory and a new monitor whose array lands on the same address
// would inherit the old accumulations.
extern "C" {
void mb200_hook_dft_chunk_D1(dft_chunk *self) __asm__("_ZN4meep9dft_chunkD1Ev");
void mb200_hook_dft_chunk_D2(dft_chunk *self) __asm__("_ZN4meep9dft_chunkD2Ev");
}
static void forget_dft_mirror(dft_chunk *self) {
  if (Engine *E = Engine::owner_of(self->dft)) E->forget(self->dft);
}
void mb200_hook_dft_chunk_D1(dft_chunk *self) {
  static void (*next)(dft_chunk *) = (void (*)(dft_chunk *))next_definition_of_caller();
  forget_dft_mirror(self);
  next(self);
}
void mb200_hook_dft_chunk_D2(dft_chunk *self) {
  static void (*next)(dft_chunk *) = (void (*)(dft_chunk *))next_definition_of_caller();
  forget_dft_mirror(self);
  next(self);
}

// ---- point probes: every fields::get_field variant funnels into this one (reference
//      src/monitor.cpp:128-133).  Reads the one or two values from HBM when the device copy is
//      the current one, so monitoring a point every step does not download whole arrays.
complex<double> fields_chunk::get_field(component c, const ivec &iloc) const {
  if (!is_mine() || !f[c][0]) return 0.0;
  const ptrdiff_t idx = gv.index(c, iloc);
  Engine *E = Engine::owner_of(f[c][0]);
  if (E && E->state == Engine::DEVICE_NEWER) {
    realnum re = 0, im = 0;
    check(mb200_d2h(E->ctx, &re, E->dev(f[c][0] + idx), sizeof(realnum)), "d2h(get_field)");
    if (f[c][1])
      check(mb200_d2h(E->ctx, &im, E->dev(f[c][1] + idx), sizeof(realnum)), "d2h(get_field)");
    E->stats.d2h_bytes += (f[c][1] ? 2 : 1) * sizeof(realnum);
    return complex<double>(re, im);
  }
  return f[c][1] ? complex<double>(f[c][0][idx], f[c][1][idx]) : complex<double>(f[c][0][idx]);
}

// ---- bulk readers: integrate, get_array_slice, output_hdf5, max_abs, energy/flux in box,
//      add_dft / add_source set-up all go through loop_in_chunks (src/loop_in_chunks.cpp:339).
void fields::loop_in_chunks(field_chunkloop chunkloop, void *chunkloop_data, const volume &where,
                            component cgrid, bool use_symmetry, bool snap_unit_dims) {
  typedef void (*fn)(fields *, field_chunkloop, void *, const volume &, component, bool, bool);
  static fn next = (fn)next_definition_of_caller();
  // only what a reader of `where` can touch is brought to the host (a flux plane, a slice, a box to
  // integrate over): Engine::sync_host_region
  if (Engine *E = Engine::find(this)) E->sync_host_region(where);
  next(this, chunkloop, chunkloop_data, where, cgrid, use_symmetry, snap_unit_dims);
}

// ---- field arrays are born without host pages -------------------------------------------------------
// fields_chunk::alloc_f (src/fields.cpp:480-504) restated: same arrays, same aliasing of H to B,
// but the zero contents are zero-fill-on-demand pages (hostmem.hpp) instead of a store loop, so a
// simulation that is set up and then stepped on the device never materialises them on the host.
bool fields_chunk::alloc_f(component c) {
  bool changed = false;
  if (is_mine()) DOCMP {
      if (!f[c][cmp]) {
        changed = true;
        if (is_magnetic(c)) {
          const component bc = direction_component(Bx, component_direction(c));
          if (!f[bc][cmp]) f[bc][cmp] = new_zeroed_lazily(gv.ntot());
          f[c][cmp] = f[bc][cmp];
        }
        else
          f[c][cmp] = new_zeroed_lazily(gv.ntot());
      }
    }
  return changed;
}

// ---- sources added mid-run -------------------------------------------------------------------------
// fields_chunk::add_source (src/sources.cpp) may MERGE a new source into an existing src_vol
// (add_amplitudes_from: same component, src_time and indices) — the amplitudes change in place and
// nothing else does.  A global generation counter, part of every Engine's fingerprint, makes the
// cached source plans (which hold uploaded amplitude copies) be re-recorded.
void fields_chunk::add_source(field_type ft, src_vol &&src) {
  typedef void (*fn)(fields_chunk *, field_type, src_vol &&);
  static fn next = (fn)next_definition_of_caller();
  Engine::source_generation++;
  next(this, ft, std::move(src));
}

// ---- host-side writers of field arrays -----------------------------------------------------------
#define MB200_WRITER_HOOK(NAME)                                                                    \
  void fields::NAME() {                                                                            \
    typedef void (*fn)(fields *);                                                                  \
    static fn next = (fn)next_definition_of_caller();                                              \
    Engine *E = Engine::find(this);                                                                \
    if (E) E->sync_host();                                                                         \
    next(this);                                                                                    \
    if (E) E->state = Engine::HOST_NEWER;                                                          \
  }
MB200_WRITER_HOOK(zero_fields)
MB200_WRITER_HOOK(use_real_fields)
MB200_WRITER_HOOK(remove_susceptibilities)
#undef MB200_WRITER_HOOK

void fields::initialize_field(component c, complex<double> func(const vec &)) {
  typedef void (*fn)(fields *, component, complex<double> (*)(const vec &));
  static fn next = (fn)next_definition_of_caller();
  Engine *E = Engine::find(this);
  if (E) E->sync_host();
  if (E) E->state = Engine::HOST_NEWER;
  next(this, c, func);
  if (E) E->state = Engine::HOST_NEWER;
}

void fields::load(const char *filename, bool single_parallel_file) {
  typedef void (*fn)(fields *, const char *, bool);
  static fn next = (fn)next_definition_of_caller();
  Engine *E = Engine::find(this);
  if (E) E->sync_host();
  next(this, filename, single_parallel_file);
  if (E) E->state = Engine::HOST_NEWER;
}

void fields::dump(const char *filename, bool single_parallel_file) {
  typedef void (*fn)(fields *, const char *, bool);
  static fn next = (fn)next_definition_of_caller();
  if (Engine *E = Engine::find(this)) E->sync_host();
  next(this, filename, single_parallel_file);
}

// ---- host-side readers / writers of DFT arrays ---------------------------------------------------
void dft_chunk::scale_dft(complex<double> scale_) {
  typedef void (*fn)(dft_chunk *, complex<double>);
  static fn next = (fn)next_definition_of_caller();
  Engine *E = Engine::owner_of(dft);
  if (E) E->sync_host();
  next(this, scale_);
  if (E) E->state = Engine::HOST_NEWER;
}

void dft_chunk::operator-=(const dft_chunk &chunk) {
  typedef void (*fn)(dft_chunk *, const dft_chunk &);
  static fn next = (fn)next_definition_of_caller();
  Engine *E = Engine::owner_of(dft), *E2 = Engine::owner_of(chunk.dft);
  if (E) E->sync_host();
  if (E2) E2->sync_host();
  next(this, chunk);
  if (E) E->state = Engine::HOST_NEWER;
}

double dft_chunk::norm2(grid_volume fgv) const {
  typedef double (*fn)(const dft_chunk *, grid_volume);
  static fn next = (fn)next_definition_of_caller();
  if (Engine *E = Engine::owner_of(dft)) E->sync_host();
  return next(this, fgv);
}

std::vector<complex<double> > dft_flux::complexflux() {
  typedef std::vector<complex<double> > (*fn)(dft_flux *);
  static fn next = (fn)next_definition_of_caller();
  sync_all_hosts();
  return next(this);
}

complex<double> dft_chunk::process_dft_component(int rank, direction *ds, ivec min_corner,
                                                 ivec max_corner, int num_freq, h5file *file,
                                                 realnum *buffer, int reim,
                                                 complex<realnum> *field_array, void *mode1_data,
                                                 void *mode2_data, int ic_conjugate,
                                                 bool retain_interp_weights, fields *parent) {
  typedef complex<double> (*fn)(dft_chunk *, int, direction *, ivec, ivec, int, h5file *,
                                realnum *, int, complex<realnum> *, void *, void *, int, bool,
                                fields *);
  static fn next = (fn)next_definition_of_caller();
  sync_all_hosts();
  return next(this, rank, ds, min_corner, max_corner, num_freq, file, buffer, reim, field_array,
              mode1_data, mode2_data, ic_conjugate, retain_interp_weights, parent);
}

void save_dft_hdf5(dft_chunk *dft_chunks, const char *name, h5file *file, const char *dprefix,
                   bool single_parallel_file) {
  typedef void (*fn)(dft_chunk *, const char *, h5file *, const char *, bool);
  static fn next = (fn)next_definition_of_caller();
  sync_all_hosts();
  next(dft_chunks, name, file, dprefix, single_parallel_file);
}

void load_dft_hdf5(dft_chunk *dft_chunks, const char *name, h5file *file, const char *dprefix,
                   bool single_parallel_file) {
  typedef void (*fn)(dft_chunk *, const char *, h5file *, const char *, bool);
  static fn next = (fn)next_definition_of_caller();
  sync_all_hosts();
  next(dft_chunks, name, file, dprefix, single_parallel_file);
  Engine::for_each([](Engine &E) { E.state = Engine::HOST_NEWER; });
}

void dft_near2far::farfield_lowlevel(complex<double> *EH, const vec &x, double freq_) {
  typedef void (*fn)(dft_near2far *, complex<double> *, const vec &, double);
  static fn next = (fn)next_definition_of_caller();
  sync_all_hosts();
  next(this, EH, x, freq_);
}

double *dft_force::force() {
  typedef double *(*fn)(dft_force *);
  static fn next = (fn)next_definition_of_caller();
  sync_all_hosts();
  return next(this);
}

double *dft_energy::electric() {
  typedef double *(*fn)(dft_energy *);
  static fn next = (fn)next_definition_of_caller();
  sync_all_hosts();
  return next(this);
}

double *dft_energy::magnetic() {
  typedef double *(*fn)(dft_energy *);
  static fn next = (fn)next_definition_of_caller();
  sync_all_hosts();
  return next(this);
}

void dft_ldos::update(fields &f) {
  typedef void (*fn)(dft_ldos *, fields &);
  static fn next = (fn)next_definition_of_caller();
  if (Engine *E = Engine::find(&f)) E->sync_host();
  next(this, f);
}

} // namespace meep

// ---- explicit controls for code that touches fields_chunk arrays directly -------------------------
extern "C" {
// make the host arrays of `f` current (e.g. before reading fields_chunk::f yourself)
void meep_b200_sync_host(meep::fields *f) {
  if (Engine *E = Engine::find(f)) E->sync_host();
}
// tell the engine that host code modified the arrays of `f` (re-uploaded before the next step)
void meep_b200_mark_host_dirty(meep::fields *f) {
  if (Engine *E = Engine::find(f)) {
    E->sync_host();
    E->state = Engine::HOST_NEWER;
  }
}
// counters: [0] steps, [1] H2D bytes, [2] D2H bytes, [3] kernel launches, [4] plan builds,
// [5] full uploads, [6] full downloads, [7] device bytes allocated
void meep_b200_get_stats(meep::fields *f, double out[8]) {
  for (int i = 0; i < 8; ++i)
    out[i] = 0;
  if (Engine *E = Engine::find(f)) {
    out[0] = (double)E->stats.steps;
    out[1] = E->stats.h2d_bytes;
    out[2] = E->stats.d2h_bytes;
    out[3] = (double)mb200_launch_count(E->ctx);
    out[4] = (double)E->stats.plan_builds;
    out[5] = (double)E->stats.uploads;
    out[6] = (double)E->stats.downloads;
    out[7] = (double)mb200_bytes_allocated(E->ctx);
  }
}
// the C-ABI context driving `f` (for CUDA-event timing / profiling from a bench driver)
mb200_ctx *meep_b200_ctx(meep::fields *f) {
  Engine *E = Engine::find(f);
  return E ? E->ctx : nullptr;
}
}
