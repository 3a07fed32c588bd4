// bench_driver.cpp — builds BASELINE.json's benchmark workloads with the reference's public C++
// API (meep::structure / meep::fields, exactly as a user program would) and times
// fields::step().  The same source is built three ways (meep_b200/build.py):
//
//   meep_b200/lib/libmeep_b200_bench_<p>.so  linked with libmeep_b200 in FRONT of the reference
//                                            libmeep: fields::step() runs on the B200.  Loaded
//                                            in-process by bench.py / __graft_entry__.py (ctypes).
//   meep_b200/lib/bench_ref_<p>              executable linked against the unmodified reference
//                                            ONLY: the CPU baseline / `--impl reference` arm.
//
// Workloads (SURVEY §8d "synthetic inputs"; resolution 10, Courant 0.5, real fields):
//   c2 : n^3 cells, eps = 12 cube of half the cell size, pml(1.0) on all faces, Gaussian Ez
//        point dipole (BASELINE config 2 at n = 512; config 5 at n = 1024)
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <complex>
#include <algorithm>
#include <string>

#include <omp.h>

#include "meep.hpp"
using namespace meep;

namespace {

double g_L = 1, g_cx = 0.5, g_cy = 0.5, g_cz = 0.5;

// thread-safe (so set_chi1inv runs under OpenMP: anisotropic_averaging.cpp:252) eps = 12 cube
class cube_material : public material_function {
public:
  virtual double chi1p1(field_type ft, const vec &r) {
    if (ft != E_stuff) return 1.0;
    const double dx = fabs(r.x() - g_cx), dy = fabs(r.y() - g_cy), dz = fabs(r.z() - g_cz);
    const double m = std::max(dx, std::max(dy, dz));
    return m < 0.25 * g_L ? 12.0 : 1.0;
  }
  virtual double eps(const vec &r) { return chi1p1(E_stuff, r); }
  virtual bool is_thread_safe() const { return true; }
};

// C3: Drude + 5 Lorentz "Au" sphere of radius 0.2 L (constants: reference python/materials.py:340-364)
class sphere_sigma : public material_function {
public:
  double scale;
  explicit sphere_sigma(double s) : scale(s) {}
  double inside(const vec &r) const {
    const double dx = r.x() - g_cx, dy = r.y() - g_cy, dz = r.z() - g_cz;
    return dx * dx + dy * dy + dz * dz < 0.04 * g_L * g_L ? 1.0 : 0.0;
  }
  virtual double chi1p1(field_type, const vec &r) { return scale * inside(r); }
  virtual void sigma_row(component c, double sigrow[3], const vec &r) {
    sigrow[0] = sigrow[1] = sigrow[2] = 0.0;
    sigrow[component_index(c)] = scale * inside(r);
  }
  virtual bool is_thread_safe() const { return true; }
};

class vacuum_material : public material_function {
public:
  virtual double chi1p1(field_type, const vec &) { return 1.0; }
  virtual double eps(const vec &) { return 1.0; }
  virtual bool is_thread_safe() const { return true; }
};

// C4: Si ring in the xy-plane, subpixel smoothing -> off-diagonal chi1inv
class ring_material : public material_function {
public:
  virtual double chi1p1(field_type ft, const vec &r) {
    if (ft != E_stuff) return 1.0;
    const double x = r.x() - g_cx, y = r.y() - g_cy, rr = sqrt(x * x + y * y);
    const bool inz = fabs(r.z() - g_cz) < 0.11 * 2 * g_cz;
    return (rr > 0.3 * 2 * g_cx && rr < 0.4 * 2 * g_cx && inz) ? 12.0 : 1.0;
  }
  virtual double eps(const vec &r) { return chi1p1(E_stuff, r); }
  virtual bool is_thread_safe() const { return true; }
};

struct Bench {
  dft_flux *flux = nullptr;
  structure *s = nullptr;
  fields *f = nullptr;
  grid_volume gv;
  double cells = 0;
  vec probe_pt;
  // the eight probe points of mb200_bench_probes sit around this point, reading this component: next to
  // the source, in a component the source excites, so that the values are non-zero after a few tens of
  // steps (a cross-rank / cross-build check that reads zeros checks nothing)
  vec probe_anchor;
  component probe_comp = Ez;
};

} // namespace

extern "C" {

void *mb200_bench_create3d(const char *workload, int nx, int ny, int nz, int num_chunks);

// returns NULL on failure (message on stderr)
void *mb200_bench_create(const char *workload, int n, int num_chunks) {
  return mb200_bench_create3d(workload, n, n, n, num_chunks);
}

// nx x ny x nz cells; the eps = 12 cube is centred and half as wide as the SHORTEST edge.  With
// more than one process (torchrun: RANK / WORLD_SIZE) the reference's own split_by_cost cuts the
// cell into count_processors() leaves (num_chunks = 0) and every process owns one of them.
void *mb200_bench_create3d(const char *workload, int nx, int ny, int nz, int num_chunks) {
  static initialize *mpi = nullptr;
  if (!mpi) {
    static int argc = 1;
    static char arg0[] = "bench";
    static char *argvv[] = {arg0, nullptr};
    static char **argv = argvv;
    mpi = new initialize(argc, argv);
  }
  verbosity = 0;
  try {
    Bench *b = new Bench();
    const double a = 10.0;
    if (std::string(workload) == "c2") {
      const int nmin = std::min(nx, std::min(ny, nz));
      g_L = nmin / a;
      g_cx = 0.5 * nx / a;
      g_cy = 0.5 * ny / a;
      g_cz = 0.5 * nz / a;
      b->gv = vol3d(nx / a, ny / a, nz / a, a);
      cube_material mat;
      const bool trace = getenv("MEEP_B200_VERBOSE") && atoi(getenv("MEEP_B200_VERBOSE"));
      double t0 = wall_time();
      b->s = new structure(b->gv, mat, pml(1.0), identity(), num_chunks, 0.5, false);
      if (trace) fprintf(stderr, "bench: structure %.3f s\n", wall_time() - t0);
      t0 = wall_time();
      b->f = new fields(b->s);
      b->f->use_real_fields();
      gaussian_src_time src(0.15, 0.1);
      src.is_integrated = false; // a current source, the Python front end's default
      b->f->add_point_source(Ez, src, b->gv.center() + vec(0.05, 0.05, 0.05));
      if (trace) fprintf(stderr, "bench: fields + source %.3f s\n", wall_time() - t0);
      b->probe_pt = b->gv.center() + vec(0.35, 0.25, 0.15);
      b->probe_anchor = b->gv.center();
      b->cells = (double)nx * ny * nz;
    }
    else if (std::string(workload) == "c3") {
      // BASELINE config 3 (3D Drude-Lorentz Au nanoparticle + 100-frequency DFT flux box)
      const int nmin = std::min(nx, std::min(ny, nz));
      g_L = nmin / a;
      g_cx = 0.5 * nx / a;
      g_cy = 0.5 * ny / a;
      g_cz = 0.5 * nz / a;
      b->gv = vol3d(nx / a, ny / a, nz / a, a);
      vacuum_material vac;
      b->s = new structure(b->gv, vac, pml(1.0), identity(), num_chunks, 0.5, false);
      // unit length 0.1 um (python/materials.py um_scale = 0.1), i.e. 10 nm pixels: with 1 um units the
      // Drude term has omega_p dt = 2.0 at resolution 10 and the run diverges in the reference itself
      const double eV = 0.1 / 1.23984193;
      const double frq[6] = {1e-3, 0.415 * eV, 0.830 * eV, 2.969 * eV, 4.304 * eV, 13.32 * eV};
      const double gam[6] = {0.053 * eV, 0.241 * eV, 0.345 * eV, 0.870 * eV, 2.494 * eV, 2.214 * eV};
      const double wp = 9.03 * eV;
      const double fstr[6] = {0.760, 0.024, 0.010, 0.071, 0.601, 4.384};
      for (int k = 0; k < 6; ++k) {
        sphere_sigma sg(fstr[k] * wp * wp / (frq[k] * frq[k]));
        b->s->add_susceptibility(sg, E_stuff, lorentzian_susceptibility(frq[k], gam[k], k == 0));
      }
      b->f = new fields(b->s);
      b->f->use_real_fields();
      gaussian_src_time src(0.4, 0.3);
      src.is_integrated = false;
      b->f->add_point_source(Ez, src, vec(0.15 * nx / a, g_cy, g_cz));
      volume box(vec(0.25 * nx / a, 0.25 * ny / a, 0.25 * nz / a),
                 vec(0.75 * nx / a, 0.75 * ny / a, 0.75 * nz / a));
      b->flux = new dft_flux(b->f->add_dft_flux_box(box, 0.24, 0.56, 100));
      b->probe_pt = b->gv.center() + vec(0.35, 0.25, 0.15);
      b->probe_anchor = vec(0.15 * nx / a, g_cy, g_cz); // the source
      b->cells = (double)nx * ny * nz;
    }
    else if (std::string(workload) == "c4") {
      // BASELINE config 4 (anisotropic subpixel-smoothed Si ring, off-diagonal chi1inv)
      g_L = std::min(nx, ny) / a;
      g_cx = 0.5 * nx / a;
      g_cy = 0.5 * ny / a;
      g_cz = 0.5 * nz / a;
      b->gv = vol3d(nx / a, ny / a, nz / a, a);
      ring_material mat;
      b->s = new structure(b->gv, mat, pml(1.0), identity(), num_chunks, 0.5, true, 1e-2, 2000);
      b->f = new fields(b->s);
      b->f->use_real_fields();
      gaussian_src_time src(0.3, 0.2);
      src.is_integrated = false;
      b->f->add_point_source(Hz, src, vec(g_cx + 0.35 * nx / a, g_cy, g_cz));
      b->probe_pt = b->gv.center() + vec(0.35, 0.25, 0.05);
      b->probe_anchor = vec(g_cx + 0.35 * nx / a, g_cy, g_cz); // the (Hz) source: Ey next to it
      b->probe_comp = Ey;
      b->cells = (double)nx * ny * nz;
    }
    else {
      fprintf(stderr, "mb200_bench_create: unknown workload %s\n", workload);
      delete b;
      return nullptr;
    }
    return b;
  } catch (std::exception &e) {
    fprintf(stderr, "mb200_bench_create: %s\n", e.what());
    return nullptr;
  }
}

int mb200_bench_step(void *h, int nsteps) {
  Bench *b = (Bench *)h;
  try {
    const bool trace = getenv("MEEP_B200_VERBOSE") && atoi(getenv("MEEP_B200_VERBOSE"));
    for (int i = 0; i < nsteps; ++i) {
      const double t0 = trace ? wall_time() : 0;
      b->f->step();
      if (trace && b->f->t <= 4) fprintf(stderr, "bench: step %d took %.3f s (host)\n", b->f->t, wall_time() - t0);
    }
    return 0;
  } catch (std::exception &e) {
    fprintf(stderr, "mb200_bench_step: %s\n", e.what());
    return 1;
  }
}

// what a user's monitoring loop does after each step: read one field value on the host
double mb200_bench_probe(void *h) {
  Bench *b = (Bench *)h;
  return real(b->f->get_field(Ez, b->probe_pt));
}

// Ez at n fixed points of the cell (the same physical points whatever the sharding); every rank
// gets every value (fields::get_field reduces over the processes)
void mb200_bench_probes(void *h, double *out, int n) {
  Bench *b = (Bench *)h;
  // within ~10 pixels of the cell centre, on both sides of it: the source sits next to the centre (the
  // fields are non-zero there after a few tens of steps), and the centre is where the leaves of a
  // 2x2x2 partition meet, so the points belong to different ranks
  const vec c = b->probe_anchor;
  for (int k = 0; k < n; ++k) {
    const double s = (k % 2 ? -1.0 : 1.0) * k * 0.1;
    const vec p = c + vec(0.35 + s, 0.25 - s, 0.15 + s);
    out[k] = real(b->f->get_field(b->probe_comp, p));
  }
}
const char *mb200_bench_probe_component(void *h) { return component_name(((Bench *)h)->probe_comp); }

double mb200_bench_cells(void *h) { return ((Bench *)h)->cells; }
int mb200_bench_num_chunks(void *h) { return ((Bench *)h)->f->num_chunks; }
int mb200_bench_time_step(void *h) { return ((Bench *)h)->f->t; }
void *mb200_bench_fields(void *h) { return ((Bench *)h)->f; }

// total bytes of field-like host arrays (what a cold start uploads / a final sync downloads)
double mb200_bench_field_bytes(void *h) {
  Bench *b = (Bench *)h;
  double bytes = 0;
  for (int i = 0; i < b->f->num_chunks; ++i) {
    fields_chunk *fc = b->f->chunks[i];
    if (!fc->is_mine()) continue;
    FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) {
      if (fc->f[c][cmp] && !(is_magnetic(c) &&
                             fc->f[c][cmp] == fc->f[direction_component(Bx, component_direction(c))][cmp]))
        bytes += fc->gv.ntot() * sizeof(realnum);
      if (fc->f_u[c][cmp]) bytes += fc->gv.ntot() * sizeof(realnum);
      if (fc->f_w[c][cmp]) bytes += fc->gv.ntot() * sizeof(realnum);
      if (fc->f_cond[c][cmp]) bytes += fc->gv.ntot() * sizeof(realnum);
    }
  }
  return bytes;
}

// algorithmic HBM bytes per time step (SURVEY §8d accounting, evaluated on the actual chunk
// layout): every array element that a half-step must read is counted once, every element it
// must write once; neighbour re-reads are free (on chip); D->E fused.
double mb200_bench_algorithmic_bytes_per_step(void *h) {
  Bench *b = (Bench *)h;
  const double R = sizeof(realnum);
  double bytes = 0;
  for (int i = 0; i < b->f->num_chunks; ++i) {
    fields_chunk *fc = b->f->chunks[i];
    if (!fc->is_mine()) continue;
    const double owned = (double)fc->gv.nx() * fc->gv.ny() * fc->gv.nz();
    int arrays = 0;
    FOR_COMPONENTS(c) {
      if (!fc->f[c][0]) continue;
      const component bc = direction_component(Bx, component_direction(c));
      if (is_B(c)) arrays += 2;                                   // B r/w
      if (is_D(c)) arrays += 2;                                   // D r/w
      if (is_electric(c)) arrays += 1 /* read by the B curl */ + 1 /* written by update_eh */;
      if (is_magnetic(c)) {
        arrays += 1;                                              // read by the D curl
        if (fc->f[c][0] != fc->f[bc][0]) arrays += 1;             // separate H: written
      }
      if (fc->f_u[c][0]) arrays += 2;
      if (fc->f_cond[c][0]) arrays += 2;
      if (fc->f_w[c][0]) arrays += 2 + 1;                         // fw r/w + E/H read (+=)
      if (is_electric(c) || is_magnetic(c)) {
        FOR_DIRECTIONS(d) if (fc->s->chi1inv[c][d]) arrays += 1;  // diagonal + off-diagonal chi1inv
        // off-diagonal terms forbid the D->E fusion: D is re-read after its halo exchange
        const direction dd = component_direction(c);
        FOR_DIRECTIONS(d) if (d != dd && fc->s->chi1inv[c][d]) { arrays += 1; break; }
      }
    }
    // Lorentz/Drude polarisations: P, P_prev r/w + sigma r per polarised component (5R, fused
    // with the E update: SURVEY 8d)
    FOR_FIELD_TYPES(ft) for (polarization_state *p = fc->pol[ft]; p; p = p->next) FOR_COMPONENTS(c) {
      if (p->s->sigma[c][component_direction(c)] && fc->f[c][0]) arrays += 5;
    }
    bytes += R * arrays * owned;
  }
  return bytes;
}

// sum of the flux spectrum (C3): forces the DFT arrays through dft_flux::flux()
double mb200_bench_flux_sum(void *h) {
  Bench *b = (Bench *)h;
  if (!b->flux) return 0;
  double *F = b->flux->flux(), s = 0;
  for (size_t i = 0; i < b->flux->freq.size(); ++i)
    s += F[i];
  delete[] F;
  return s;
}

// DFT state bytes (C3) and monitor points
double mb200_bench_dft_bytes(void *h) {
  Bench *b = (Bench *)h;
  double bytes = 0;
  for (int i = 0; i < b->f->num_chunks; ++i)
    for (dft_chunk *d = b->f->chunks[i]->dft_chunks; d; d = d->next_in_chunk)
      bytes += 2.0 * sizeof(realnum) * d->N * d->omega.size();
  return bytes;
}

void mb200_bench_destroy(void *h) {
  Bench *b = (Bench *)h;
  if (!b) return;
  delete b->flux;
  delete b->f;
  delete b->s;
  delete b;
}

} // extern "C"

#ifdef MB200_BENCH_MAIN
// CPU arm: time fields::step() of the unmodified reference on the host cores.
// usage: bench_ref_<p> <workload> <n> <warmup> <steps> [num_chunks]
int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <workload> <n> <warmup> <steps> [num_chunks]\n", argv[0]);
    return 2;
  }
  const int n = atoi(argv[2]), warm = atoi(argv[3]), steps = atoi(argv[4]);
  omp_set_num_threads(omp_get_max_threads());
  const int nchunks = argc > 5 ? atoi(argv[5]) : 0;
  double t0 = wall_time();
  void *h = mb200_bench_create(argv[1], n, nchunks);
  if (!h) return 1;
  double t_setup = wall_time() - t0;
  t0 = wall_time();
  if (mb200_bench_step(h, warm)) return 1; // includes the one-time connect_the_chunks
  double t_warm = wall_time() - t0;
  t0 = wall_time();
  if (mb200_bench_step(h, steps)) return 1;
  double t = wall_time() - t0;
  printf("{\"workload\": \"%s\", \"n\": %d, \"cells\": %.0f, \"steps\": %d, \"warmup\": %d, "
         "\"seconds\": %.6f, \"cells_per_s\": %.6e, \"setup_s\": %.3f, \"warmup_s\": %.3f, "
         "\"probe\": %.17g, \"num_chunks\": %d, \"realnum_bytes\": %d}\n",
         argv[1], n, mb200_bench_cells(h), steps, warm, t, mb200_bench_cells(h) * steps / t, t_setup,
         t_warm, mb200_bench_probe(h), mb200_bench_num_chunks(h), (int)sizeof(realnum));
  mb200_bench_destroy(h);
  return 0;
}
#endif
