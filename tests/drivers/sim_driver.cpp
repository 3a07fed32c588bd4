// sim_driver.cpp — parity driver written against the reference's public C++ API (meep.hpp).
//
// The SAME source is linked twice:
//   sim_ref_<prec>   : against the unmodified reference build only (oracle/_ref)  -> oracle arm
//   sim_b200_<prec>  : with libmeep_b200 in front of it                            -> device arm
//   (sim_emu_<prec>  : drop-in + the test-only emulator, for the GPU-less CI here)
// It builds one of the named simulations below with meep::structure / meep::fields exactly as
// the reference's own tests do (tests/known_results.cpp, tests/three_d.cpp, tests/pml.cpp,
// tests/flux.cpp, tests/bend-flux-ll.cpp), runs N steps, and dumps every field-like array
// (f, f_u, f_w, f_cond, polarisation P/P_prev, dft arrays) plus flux spectra and a few
// get_field probes to a binary file that tests/test_parity_*.py compares with rel-L2.
//
// usage: sim_driver <case> <nsteps> <out.bin> [num_chunks]
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <complex>
#include <string>
#include <vector>

#include "meep.hpp"
using namespace meep;
using std::complex;

// ---- dump format: repeated records { u32 name_len; name; u32 elem_size; u64 count; data } ------
static FILE *g_out = NULL;
static void dump(const std::string &name, const void *data, unsigned elem_size, size_t count) {
  unsigned nl = (unsigned)name.size();
  unsigned long long c = count;
  fwrite(&nl, 4, 1, g_out);
  fwrite(name.data(), 1, nl, g_out);
  fwrite(&elem_size, 4, 1, g_out);
  fwrite(&c, 8, 1, g_out);
  fwrite(data, elem_size, count, g_out);
}

// restated layout of the reference's file-local lorentzian_data (src/susceptibility.cpp:98-104)
struct lorentzian_data_layout {
  size_t sz_data;
  size_t ntot;
  realnum *P[NUM_FIELD_COMPONENTS][2];
  realnum *P_prev[NUM_FIELD_COMPONENTS][2];
  realnum data[1];
};

static void sync_host(fields &f) {
  typedef void (*fn)(fields *);
  fn s = (fn)dlsym(RTLD_DEFAULT, "meep_b200_sync_host");
  if (s) s(&f);
}

static void dump_fields(fields &f) {
  sync_host(f);
  char nm[128];
  for (int i = 0; i < f.num_chunks; ++i) {
    fields_chunk *fc = f.chunks[i];
    const size_t n = fc->gv.ntot();
    FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) {
      const char *cn = component_name(c);
      // B aliases H where the reference left them shared: dump once, under the B name
      if (is_magnetic(c) && fc->f[c][cmp] &&
          fc->f[c][cmp] == fc->f[direction_component(Bx, component_direction(c))][cmp])
        continue;
#define DUMP(arr, tag)                                                                             \
  if (fc->arr[c][cmp]) {                                                                           \
    snprintf(nm, sizeof nm, "chunk%d." tag ".%s.%d", i, cn, cmp);                                  \
    dump(nm, fc->arr[c][cmp], sizeof(realnum), n);                                                 \
  }
      DUMP(f, "f")
      DUMP(f_u, "f_u")
      DUMP(f_w, "f_w")
      DUMP(f_cond, "f_cond")
      DUMP(f_bfast, "f_bfast")
#undef DUMP
    }
    FOR_FIELD_TYPES(ft) {
      int ip = 0;
      for (polarization_state *p = fc->pol[ft]; p; p = p->next, ++ip)
        if (p->data && dynamic_cast<const gyrotropic_susceptibility *>(p->s)) {
          struct gyro_layout { // src/susceptibility.cpp:374-380
            size_t sz_data, ntot;
            realnum *P[NUM_FIELD_COMPONENTS][2][3];
            realnum *P_prev[NUM_FIELD_COMPONENTS][2][3];
          } *d = (gyro_layout *)p->data;
          FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) for (int k = 0; k < 3; ++k) if (d->P[c][cmp][k]) {
            snprintf(nm, sizeof nm, "chunk%d.P%d.%s.%d.%d", i, ip, component_name(c), cmp, k);
            dump(nm, d->P[c][cmp][k], sizeof(realnum), n);
            snprintf(nm, sizeof nm, "chunk%d.Pprev%d.%s.%d.%d", i, ip, component_name(c), cmp, k);
            dump(nm, d->P_prev[c][cmp][k], sizeof(realnum), n);
          }
        }
        else if (p->data) {
          lorentzian_data_layout *d = (lorentzian_data_layout *)p->data;
          FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) if (d->P[c][cmp]) {
            snprintf(nm, sizeof nm, "chunk%d.P%d.%s.%d", i, ip, component_name(c), cmp);
            dump(nm, d->P[c][cmp], sizeof(realnum), n);
            snprintf(nm, sizeof nm, "chunk%d.Pprev%d.%s.%d", i, ip, component_name(c), cmp);
            dump(nm, d->P_prev[c][cmp], sizeof(realnum), n);
          }
        }
    }
    int id = 0;
    for (dft_chunk *d = fc->dft_chunks; d; d = d->next_in_chunk, ++id) {
      snprintf(nm, sizeof nm, "chunk%d.dft%d.%s", i, id, component_name(d->c));
      dump(nm, d->dft, sizeof(realnum), 2 * d->N * d->omega.size());
    }
  }
}

static void dump_flux(const char *name, dft_flux &fl) {
  double *F = fl.flux();
  dump(name, F, sizeof(double), fl.freq.size());
  delete[] F;
}

// ---- materials ---------------------------------------------------------------------------------
static double g_L = 1;
static double one(const vec &) { return 1.0; }
static double eps_box(const vec &r) { // eps = 12 inside the central half-size cube, 1 outside
  double m = 0;
  LOOP_OVER_DIRECTIONS(r.dim, d) {
    double x = fabs(r.in_direction(d) - 0.5 * g_L);
    if (x > m) m = x;
  }
  return m < 0.25 * g_L ? 12.0 : 1.0;
}
static double sphere(const vec &r) { // 1 inside a sphere of radius 0.2 L at the centre
  double r2 = 0;
  LOOP_OVER_DIRECTIONS(r.dim, d) {
    double x = r.in_direction(d) - 0.5 * g_L;
    r2 += x * x;
  }
  return r2 < 0.04 * g_L * g_L ? 1.0 : 0.0;
}
static double cond_slab(const vec &r) { return r.in_direction(r.dim == D1 ? Z : X) > 0.5 * g_L ? 0.7 : 0.0; }
static double chi3_box(const vec &r) { return eps_box(r) > 1 ? 0.3 : 0.0; }
static double eps_smooth(const vec &r) { // smooth profile -> anisotropic averaging gives off-diag
  double s = 0;
  LOOP_OVER_DIRECTIONS(r.dim, d) { s += (d + 1) * r.in_direction(d); }
  return 6.5 + 5.5 * sin(1.7 * s);
}
static double eps_ring(const vec &r) { // Si ring in the xy-plane (BASELINE config 4, scaled)
  double x = r.x() - 0.5 * g_L, y = r.y() - 0.5 * g_L;
  double rr = sqrt(x * x + y * y);
  bool inz = r.dim == D3 ? fabs(r.z() - 0.5 * g_L) < 0.11 * g_L : true;
  return (rr > 0.3 * g_L && rr < 0.4 * g_L && inz) ? 12.0 : 1.0;
}

// explicit off-diagonal material as in the reference's tests/pml.cpp:13-44
class offdiag_material : public material_function {
public:
  offdiag_material(double od) : offdiag(od) {}
  virtual bool has_mu() { return true; }
  virtual void eff_chi1inv_row(component c, double chi1inv_row[3], const volume &v, double tol,
                               int maxeval) {
    (void)v; (void)tol; (void)maxeval;
    double detinv = 1.0 / (1 + 2 * offdiag);
    if (component_direction(c) == X) {
      chi1inv_row[0] = (1 + offdiag) * detinv; chi1inv_row[1] = -offdiag * detinv; chi1inv_row[2] = 0.0;
    }
    else if (component_direction(c) == Y) {
      chi1inv_row[0] = -offdiag * detinv; chi1inv_row[1] = (1 + offdiag) * detinv; chi1inv_row[2] = 0.0;
    }
    else { chi1inv_row[0] = 0.0; chi1inv_row[1] = 0.0; chi1inv_row[2] = 1.0; }
  }
  double offdiag;
};

static void probes(fields &f, const grid_volume &gv) {
  // exercises the interposed point-probe path (fields::get_field -> fields_chunk::get_field)
  std::vector<double> v;
  const component cs[] = {Ex, Ey, Ez, Hx, Hy, Hz, Dx, Dy, Dz};
  vec p = gv.center();
  for (component c : cs)
    if (gv.has_field(c) && f.have_component(c)) {
      complex<double> z = f.get_field(c, p);
      v.push_back(z.real());
      v.push_back(z.imag());
    }
  dump("probe.center", v.data(), sizeof(double), v.size());
}

int main(int argc, char **argv) {
  initialize mpi(argc, argv);
  verbosity = 0;
  if (argc < 4) {
    fprintf(stderr, "usage: %s <case> <nsteps> <out.bin> [num_chunks]\n", argv[0]);
    return 2;
  }
  const std::string cs = argv[1];
  const int nsteps = atoi(argv[2]);
  const int num_chunks = argc > 4 ? atoi(argv[4]) : 0;
  std::string outname = argv[3];
  if (count_processors() > 1) outname += ".rank" + std::to_string(my_rank());
  g_out = fopen(outname.c_str(), "wb");
  if (!g_out) { perror(argv[3]); return 2; }
  const double a = 10.0;

  if (cs == "c2_3d_pml" || cs == "c2_3d_pml_integrated" || cs == "c2_3d_pml_complex") {
    // BASELINE config 2 (scaled twin): 3-D eps=12 cube + PML on all faces, Gaussian Ez dipole
    g_L = getenv("MB200_TEST_L") ? atof(getenv("MB200_TEST_L")) : 3.2;
    grid_volume gv = vol3d(g_L, g_L, g_L, a);
    structure s(gv, eps_box, pml(1.0), identity(), num_chunks);
    fields f(&s);
    if (cs != "c2_3d_pml_complex") f.use_real_fields();
    gaussian_src_time src(0.15, 0.1);
    src.is_integrated = (cs == "c2_3d_pml_integrated");
    f.add_point_source(Ez, src, gv.center() + vec(0.05, 0.05, 0.05));
    const double w0 = wall_time();
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
    if (getenv("MB200_DUMP_TIMES")) { // phase timers as fields::time_spent_on reports them (not a parity quantity)
      const time_sink sinks[] = {Stepping, Boundaries, FourierTransforming, FieldUpdateB, FieldUpdateH, FieldUpdateD,
                                 FieldUpdateE, BoundarySteppingB, BoundarySteppingH, BoundarySteppingD, BoundarySteppingE};
      std::vector<double> tv;
      for (time_sink sk : sinks) tv.push_back(f.time_spent_on(sk)[0]);
      tv.push_back(wall_time() - w0);
      dump("times", tv.data(), sizeof(double), tv.size());
    }
  }
  else if (cs == "3d_metal") { // no PML: metallic walls (zero_metal path)
    g_L = 1.6;
    grid_volume gv = vol3d(g_L, 1.2, 1.0, a);
    structure s(gv, eps_box, no_pml(), identity(), num_chunks);
    fields f(&s);
    f.add_point_source(Ez, 0.2, 3.0, 0.0, 2.0, gv.center(), complex<double>(0, -2 * pi * 0.2));
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "3d_bloch") { // fully periodic with Bloch phases (CONNECT_PHASE path)
    g_L = 1.0;
    grid_volume gv = vol3d(1.0, 1.0, 1.0, a);
    structure s(gv, eps_box, no_pml(), identity(), num_chunks);
    fields f(&s);
    f.add_point_source(Ez, 0.2, 3.0, 0.0, 2.0, gv.center(), complex<double>(0, -2 * pi * 0.2));
    f.use_bloch(vec(0.3, 0.5, 0.8));
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "3d_bfast" || cs == "2d_bfast") {
    // BFAST (src/step_db.cpp:129-143, step_bfast): oblique-incidence Bloch-periodic cell, PML along
    // y (f_u / PML-in-f variants), a conducting slab (cnd / f_cond variants), complex fields
    g_L = 1.0;
    grid_volume gv = cs == "3d_bfast" ? vol3d(1.0, 3.0, 1.0, a) : voltwo(1.0, 3.0, a);
    structure s(gv, eps_box, pml(1.0, Y), identity(), num_chunks, 0.4);
    s.set_conductivity(Dz, cond_slab);
    s.set_conductivity(Dx, cond_slab);
    std::vector<double> bk = {0.2, 0.0, cs == "3d_bfast" ? 0.1 : 0.0};
    fields f(&s, 0.0, 0.0, true, 0, 0, bk);
    f.add_point_source(Ez, 0.4, 3.0, 0.0, 2.0, gv.center(), complex<double>(0, -2 * pi * 0.2));
    f.add_point_source(Hz, 0.4, 3.0, 0.0, 2.0, gv.center(), 1.0);
    f.use_bloch(X, 0.0);
    if (cs == "3d_bfast") f.use_bloch(Z, 0.0);
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "3d_phase_in" || cs == "3d_phase_in_cond" || cs == "3d_bloch_change") {
    // fields::phase_in_material (material arrays rewritten on the host every step while phasing,
    // src/step.cpp:141-166) / use_bloch changed mid-run (boundary phases and connections rebuilt)
    g_L = 1.6;
    grid_volume gv = vol3d(1.6, 1.6, 1.2, a);
    const bool phase_in = cs != "3d_bloch_change";
    structure s(gv, one, phase_in ? pml(0.3) : no_pml(), identity(), num_chunks);
    structure s2(gv, eps_box, phase_in ? pml(0.3) : no_pml(), identity(), num_chunks);
    if (cs == "3d_phase_in_cond") { // the phased-in structure INTRODUCES conductivity (condinv appears mid-run)
      s2.set_conductivity(Dx, cond_slab);
      s2.set_conductivity(Dy, cond_slab);
      s2.set_conductivity(Dz, cond_slab);
      s2.set_conductivity(By, cond_slab);
    }
    fields f(&s);
    gaussian_src_time src(0.5, 0.4);
    f.add_point_source(Ez, src, vec(0.8, 0.8, 0.6));
    if (cs == "3d_bloch_change") f.use_bloch(vec(0.1, 0.2, 0.0));
    for (int i = 0; i < nsteps / 3; ++i) f.step();
    if (phase_in) f.phase_in_material(&s2, 0.5 * (nsteps / 3) * f.dt);
    else f.use_bloch(vec(0.3, -0.1, 0.25));
    for (int i = 0; i < 2 * (nsteps / 3); ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "3d_midrun_changes") {
    // things user code does between steps: a source and a DFT monitor added mid-run, sources
    // removed, fields::reset(), then stepping again (plans must follow; host writers must be seen)
    g_L = 1.6;
    grid_volume gv = vol3d(1.6, 1.6, 1.6, a);
    structure s(gv, eps_box, pml(0.3), identity(), num_chunks);
    s.add_susceptibility(sphere, E_stuff, lorentzian_susceptibility(0.8, 0.05));
    fields f(&s);
    f.use_real_fields();
    gaussian_src_time src(0.5, 0.4);
    f.add_point_source(Ez, src, vec(0.8, 0.8, 0.8));
    const int q = nsteps / 4;
    for (int i = 0; i < q; ++i) f.step();
    continuous_src_time cw(0.45, 0.0, f.time());
    f.add_point_source(Hy, cw, vec(0.6, 0.9, 0.7));
    for (int i = 0; i < q; ++i) f.step();
    volume box(vec(0.5, 0.5, 0.5), vec(1.1, 1.1, 1.1));
    dft_flux fl = f.add_dft_flux_box(box, 0.3, 0.7, 5);
    for (int i = 0; i < q; ++i) f.step();
    dump_flux("flux.before_reset", fl);
    // a monitor removed mid-run and a new one of the same shape added (its arrays are likely to
    // land on the addresses just freed): the new one must start from zero, and the host readers
    // below must not be handed the freed arrays
    fl.remove();
    dft_flux fl2 = f.add_dft_flux_box(box, 0.3, 0.7, 5);
    f.remove_sources();
    for (int i = 0; i < 3; ++i) f.step();
    std::vector<double> e0 = {f.field_energy()};
    f.reset();
    f.add_point_source(Ex, src, vec(0.9, 0.7, 0.8));
    for (int i = 0; i < q; ++i) f.step();
    e0.push_back(f.field_energy());
    dump("energy", e0.data(), sizeof(double), e0.size());
    dump_flux("flux.second_monitor", fl2);
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "3d_tiled") {
    // loop tiling of step_db / update_eh (fields ctor loop_tile_base_db / _eh; split_into_tiles,
    // src/step_db.cpp:47, src/update_eh.cpp:31-50): anisotropic medium so that update_eh tiles too
    g_L = 2.0;
    grid_volume gv = vol3d(2.0, 1.8, 1.6, a);
    structure s(gv, eps_smooth, pml(0.4), identity(), num_chunks, 0.5, true, 1e-2, 2000);
    fields f(&s, 0.0, 0.0, true, 6, 5);
    f.use_real_fields();
    gaussian_src_time src(0.4, 0.3);
    f.add_point_source(Ez, src, vec(0.9, 0.8, 0.7));
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "3d_sync_magnetic") {
    // fields::synchronize_magnetic_fields / restore_magnetic_fields (src/energy_and_flux.cpp:149-178)
    // around an energy evaluation, stepping on afterwards, and a dump in the synchronised state
    g_L = 2.0;
    grid_volume gv = vol3d(2.0, 2.0, 2.0, a);
    structure s(gv, eps_box, pml(0.5), identity(), num_chunks);
    fields f(&s);
    f.use_real_fields();
    f.add_point_source(Ez, 0.3, 3.0, 0.0, 2.0, gv.center() + vec(0.05, 0.05, 0.05));
    f.add_point_source(Hx, 0.3, 3.0, 0.0, 2.0, gv.center());
    std::vector<double> en;
    for (int i = 0; i < nsteps; ++i) {
      f.step();
      if (i % 7 == 3) {
        en.push_back(f.field_energy());
        en.push_back(f.magnetic_energy_in_box(gv.surroundings()));
      }
    }
    dump("energy", en.data(), sizeof(double), en.size());
    // stepping while synchronised (restore -> step -> re-synchronise inside fields::step, twice
    // nested), then a dump in the synchronised state
    f.synchronize_magnetic_fields();
    f.synchronize_magnetic_fields();
    for (int i = 0; i < 5; ++i) f.step();
    dump_fields(f);
    f.restore_magnetic_fields();
    f.restore_magnetic_fields();
  }
  else if (cs == "3d_flux_planes") {
    // host readers of SUB-VOLUMES while the arrays live on the device: legacy flux planes (flux_vol,
    // updated twice per step inside fields::step), flux / energy in a box, a field integral over a
    // thin slab — each brings only the index box it reads to the host (Engine::sync_host_region)
    g_L = 3.0;
    grid_volume gv = vol3d(3.0, 2.6, 2.2, a);
    structure s(gv, eps_box, pml(0.5), identity(), num_chunks);
    fields f(&s);
    f.use_real_fields();
    gaussian_src_time src(0.4, 0.3);
    f.add_point_source(Ez, src, vec(1.1, 1.3, 1.0));
    flux_vol *fx = f.add_flux_plane(vec(1.9, 0.7, 0.6), vec(1.9, 1.9, 1.6));
    flux_vol *fy = f.add_flux_plane(vec(0.8, 1.9, 0.6), vec(2.0, 1.9, 1.6));
    flux_vol *fz = f.add_flux_plane(vec(0.8, 0.7, 1.6), vec(2.0, 1.9, 1.6));
    std::vector<double> rec;
    const volume box(vec(0.9, 0.9, 0.8), vec(1.5, 1.6, 1.3));
    const volume slab(vec(0.7, 0.7, 1.2), vec(2.1, 1.9, 1.3));
    for (int i = 0; i < nsteps; ++i) {
      f.step();
      rec.push_back(fx->flux());
      rec.push_back(fy->flux());
      rec.push_back(fz->flux());
      if (i % 5 == 4) {
        rec.push_back(f.electric_energy_in_box(box));
        rec.push_back(f.flux_in_box(X, volume(vec(1.6, 0.8, 0.7), vec(1.6, 1.8, 1.5))));
        rec.push_back(f.field_energy_in_box(slab));
      }
    }
    dump("monitors", rec.data(), sizeof(double), rec.size());
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "3d_xperiodic_ypml") {
    g_L = 1.0;
    grid_volume gv = vol3d(1.0, 3.0, 1.0, a);
    structure s(gv, one, pml(1.0, Y), identity(), num_chunks);
    fields f(&s);
    f.add_point_source(Ez, 0.2, 3.0, 0.0, 2.0, gv.center(), complex<double>(0, -2 * pi * 0.2));
    f.use_bloch(X, 0.1);
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "2d_mirror_sym" || cs == "3d_rotate_sym") {
    // symmetry-reduced cells (reference tests/symmetry.cpp): connections with -1 / complex phases
    // between a chunk and its own mirror / rotated image
    if (cs == "2d_mirror_sym") {
      g_L = 2.0;
      grid_volume gv = voltwo(2.0, 1.6, a);
      const symmetry S = mirror(X, gv) + mirror(Y, gv);
      structure s(gv, one, pml(0.4), S, num_chunks);
      fields f(&s);
      f.add_point_source(Ez, 0.7, 2.5, 0.0, 4.0, gv.center());
      f.add_point_source(Hz, 0.6, 2.0, 0.0, 4.0, vec(0.7, 0.55));
      for (int i = 0; i < nsteps; ++i) f.step();
      probes(f, gv);
      dump_fields(f);
    }
    else {
      g_L = 1.2;
      grid_volume gv = vol3d(1.2, 1.2, 1.0, a);
      const symmetry S = rotate4(Z, gv);
      structure s(gv, one, no_pml(), S, num_chunks);
      fields f(&s);
      f.add_point_source(Ez, 0.7, 2.5, 0.0, 4.0, gv.center());
      f.add_point_source(Hz, 0.6, 2.0, 0.0, 4.0, gv.center());
      f.use_bloch(vec(0.0, 0.0, 0.2));
      for (int i = 0; i < nsteps; ++i) f.step();
      probes(f, gv);
      dump_fields(f);
    }
  }
  else if (cs == "2d_beta" || cs == "2d_beta_real") {
    // 2-D cell with an exp(i beta z) dependence (fields ctor argument beta; Python's kz_2d):
    // step_beta couples the TE and TM families (src/step_db.cpp:148-175)
    g_L = 2.4;
    grid_volume gv = voltwo(2.4, 2.0, a);
    structure s(gv, eps_box, pml(0.5), identity(), num_chunks);
    s.set_conductivity(Dx, cond_slab);
    fields f(&s, 0.0, 0.37);
    if (cs == "2d_beta_real") f.use_real_fields();
    gaussian_src_time src(0.4, 0.3);
    f.add_point_source(Ez, src, gv.center());
    f.add_point_source(Hz, src, vec(1.0, 0.8));
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs.rfind("cyl_", 0) == 0) {
    // cylindrical coordinates (src/step_db.cpp:86-122,177-462; src/update_eh.cpp:197-209):
    // cyl_m0 / cyl_m1 / cyl_mneg1 / cyl_m2 (fields zeroed near r=0) / cyl_m3_nozero (r=0 boundary
    // condition only, smaller Courant factor) / cyl_m1_cond (conductivity + PML => f_cond) /
    // cyl_m1_flux (DFT flux through an r = const surface and a z = const disc)
    g_L = 2.0;
    const double rsize = 2.0, zsize = 2.6;
    grid_volume gv = volcyl(rsize, zsize, a);
    struct rod : public material_function { // eps = 9 rod of radius 0.7 between z = 0.8 and 1.8
      virtual double chi1p1(field_type, const vec &r) {
        return (r.r() < 0.7 && r.z() > 0.8 && r.z() < 1.8) ? 9.0 : 1.0;
      }
      virtual double conductivity(component, const vec &r) { return cond && r.z() > 1.3 ? 0.6 : 0.0; }
      virtual bool has_conductivity(component c) { return cond && is_D(c); }
      virtual bool has_mu() { return false; }
      bool cond = false;
    } mat;
    double m = 0;
    bool zero_near = true;
    double courant = 0.5;
    if (cs == "cyl_m0") m = 0;
    else if (cs == "cyl_m1" || cs == "cyl_m1_cond" || cs == "cyl_m1_flux") m = 1;
    else if (cs == "cyl_mneg1") m = -1;
    else if (cs == "cyl_m2") m = 2;
    else if (cs == "cyl_m3_nozero") { m = 3; zero_near = false; courant = 0.25; }
    else { fprintf(stderr, "unknown case %s\n", cs.c_str()); return 2; }
    mat.cond = (cs == "cyl_m1_cond");
    structure s(gv, mat, pml(0.5), identity(), num_chunks, courant);
    fields f(&s, m, 0.0, zero_near);
    gaussian_src_time src(0.35, 0.3);
    f.add_point_source(Ep, src, veccyl(0.45, 1.2));
    f.add_point_source(Ez, src, veccyl(0.85, 0.9));
    f.add_point_source(Hr, src, veccyl(0.3, 1.5));
    std::vector<dft_flux> fluxes;
    if (cs == "cyl_m1_flux") {
      volume side(veccyl(1.2, 0.7), veccyl(1.2, 1.9));
      volume top(veccyl(0.0, 1.9), veccyl(1.2, 1.9));
      fluxes.push_back(f.add_dft_flux_plane(side, 0.2, 0.5, 7));
      fluxes.push_back(f.add_dft_flux_plane(top, 0.2, 0.5, 7));
    }
    for (int i = 0; i < nsteps; ++i) f.step();
    for (size_t k = 0; k < fluxes.size(); ++k)
      dump_flux(k == 0 ? "flux.side" : "flux.top", fluxes[k]);
    dump_fields(f);
  }
  else if (cs == "2d_bend_flux") {
    // BASELINE config 1 restated (tests/bend-flux-ll.cpp:47-61,137-187): 2-D Ez
    // waveguide bend, eps = 12, PML, two DFT flux planes; scaled to 8 x 16 for test time
    const double sx = 8, sy = 16, w = 1, pad = 2, dpml = 1.0;
    const double wvg_ycen = -0.5 * (sy - w - 2 * pad), wvg_xcen = 0.5 * (sx - w - 2 * pad);
    grid_volume gv = voltwo(sx, sy, a);
    gv.center_origin();
    struct wvg : public material_function {
      double sx, sy, w, ycen, xcen;
      virtual double chi1p1(field_type, const vec &r) {
        bool horiz = fabs(r.y() - ycen) <= 0.5 * w && r.x() <= xcen + 0.5 * w;
        bool vert = fabs(r.x() - xcen) <= 0.5 * w && r.y() >= ycen - 0.5 * w;
        return (horiz || vert) ? 12.0 : 1.0;
      }
      virtual double eps(const vec &r) { return chi1p1(E_stuff, r); }
    } m;
    m.sx = sx; m.sy = sy; m.w = w; m.ycen = wvg_ycen; m.xcen = wvg_xcen;
    structure s(gv, m, pml(dpml), identity(), num_chunks);
    fields f(&s);
    f.use_real_fields();
    const double fcen = 0.15, df = 0.1;
    gaussian_src_time src(fcen, df);
    src.is_integrated = false;
    volume srcv(vec(-0.5 * sx + dpml + 0.5, wvg_ycen - 0.5 * w), vec(-0.5 * sx + dpml + 0.5, wvg_ycen + 0.5 * w));
    f.add_volume_source(Ez, src, srcv);
    const int nfreq = 25;
    volume trans_v(vec(wvg_xcen - w, 0.5 * sy - dpml - 0.5), vec(wvg_xcen + w, 0.5 * sy - dpml - 0.5));
    volume refl_v(vec(-0.5 * sx + dpml + 1.5, wvg_ycen - w), vec(-0.5 * sx + dpml + 1.5, wvg_ycen + w));
    dft_flux trans = f.add_dft_flux_plane(trans_v, fcen - 0.5 * df, fcen + 0.5 * df, nfreq);
    dft_flux refl = f.add_dft_flux_plane(refl_v, fcen - 0.5 * df, fcen + 0.5 * df, nfreq);
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_flux("flux.trans", trans);
    dump_flux("flux.refl", refl);
    dump_fields(f);
  }
  else if (cs == "2d_te_pml") { // Hz polarisation, complex fields
    g_L = 3.0;
    grid_volume gv = voltwo(3.0, 2.4, a);
    structure s(gv, eps_box, pml(0.8), identity(), num_chunks);
    fields f(&s);
    f.add_point_source(Hz, 0.25, 3.0, 0.0, 2.0, gv.center(), 1.0);
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "1d_polariton") { // tests/known_results.cpp "1D polariton": Lorentzian in 1-D
    g_L = 10.0;
    grid_volume gv = volone(10.0, a);
    structure s(gv, one, no_pml(), identity(), num_chunks);
    s.add_susceptibility(one, E_stuff, lorentzian_susceptibility(0.5, 0.1));
    fields f(&s);
    f.use_real_fields();
    f.add_point_source(Ex, 0.2, 3.0, 0.0, 2.0, gv.center());
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "c3_au_sphere") {
    // BASELINE config 3 (scaled twin): Drude + 5 Lorentz "Au" sphere, all six poles with the
    // constants of the reference's python/materials.py:340-364, PML, flux box with 100 frequencies.
    // Unit length 0.1 um (materials.py's um_scale = 0.1), i.e. 10 nm pixels at resolution 10: with
    // 1 um units the Drude term has omega_p dt = 2.0 and the run diverges in the reference itself.
    // Here omega_p dt = 0.2 and max|field| stays O(1) (checked by tests/parity_util.compare).
    g_L = getenv("MB200_TEST_L") ? atof(getenv("MB200_TEST_L")) : 2.4;
    grid_volume gv = vol3d(g_L, g_L, g_L, a);
    structure s(gv, one, pml(0.6), identity(), num_chunks);
    const double um = getenv("MB200_C3_UM") ? atof(getenv("MB200_C3_UM")) : 0.1;
    const double eV = um / 1.23984193;
    const double frq[6] = {1e-3 /* Drude: value only scales sigma */, 0.415 * eV, 0.830 * eV, 2.969 * eV, 4.304 * eV, 13.32 * eV};
    const double gam[6] = {0.053 * eV, 0.241 * eV, 0.345 * eV, 0.870 * eV, 2.494 * eV, 2.214 * eV};
    const double wp = 9.03 * eV;
    const double fstr[6] = {0.760, 0.024, 0.010, 0.071, 0.601, 4.384};
    struct sig : public material_function {
      double scale;
      virtual double chi1p1(field_type, const vec &r) { return scale * sphere(r); }
      virtual void sigma_row(component c, double sigrow[3], const vec &r) {
        sigrow[0] = sigrow[1] = sigrow[2] = 0.0;
        sigrow[component_index(c)] = scale * sphere(r);
      }
    };
    for (int k = 0; k < 6; ++k) {
      sig sg;
      sg.scale = fstr[k] * wp * wp / (frq[k] * frq[k]);
      s.add_susceptibility(sg, E_stuff, lorentzian_susceptibility(frq[k], gam[k], k == 0));
    }
    fields f(&s);
    f.use_real_fields();
    const double fc = getenv("MB200_C3_FC") ? atof(getenv("MB200_C3_FC")) : 0.4;
    gaussian_src_time src(fc, 0.75 * fc); // 250 nm centre wavelength (a longer one makes the 240 nm cell quasi-static:
                                          // B is then pure cancellation noise of curl E, an ill-conditioned parity case)
    src.is_integrated = false;
    f.add_point_source(Ez, src, vec(0.15 * g_L + 0.6, 0.5 * g_L, 0.5 * g_L));
    volume box(vec(0.25 * g_L, 0.25 * g_L, 0.25 * g_L), vec(0.75 * g_L, 0.75 * g_L, 0.75 * g_L));
    dft_flux fl = f.add_dft_flux_box(box, 0.6 * fc, 1.4 * fc, 100);
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_flux("flux.box", fl);
    dump_fields(f);
  }
  else if (cs == "lorentz_3d") { // Drude + 2 Lorentz poles with moderate constants (float-safe)
    g_L = 2.4;
    grid_volume gv = vol3d(g_L, 2.0, 1.6, a);
    structure s(gv, eps_box, pml(0.5), identity(), num_chunks);
    s.add_susceptibility(sphere, E_stuff, lorentzian_susceptibility(1.0, 0.05, true));
    s.add_susceptibility(sphere, E_stuff, lorentzian_susceptibility(0.6, 0.02));
    s.add_susceptibility(eps_box, E_stuff, lorentzian_susceptibility(1.3, 0.1));
    fields f(&s);
    gaussian_src_time src(0.6, 0.5);
    f.add_point_source(Ez, src, vec(0.8, 1.0, 0.8));
    volume box(vec(0.7, 0.6, 0.5), vec(1.7, 1.4, 1.1));
    dft_flux fl = f.add_dft_flux_box(box, 0.4, 0.8, 11);
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_flux("flux.box", fl);
    dump_fields(f);
  }
  else if (cs == "gyro_lorentz_3d" || cs == "gyro_drude_3d" || cs == "gyro_saturated_3d") {
    // gyrotropic media (src/susceptibility.cpp:349-602): nine polarisation arrays per cell, coupled
    // through the bias vector; needs the not-owned W values of neighbouring chunks
    g_L = 2.0;
    grid_volume gv = vol3d(2.0, 1.6, 2.0, a);
    structure s(gv, eps_box, pml(0.4), identity(), num_chunks);
    const vec bias(0.3, -0.2, 0.9);
    if (cs == "gyro_saturated_3d")
      s.add_susceptibility(sphere, E_stuff, gyrotropic_susceptibility(bias, 0.7, 0.02, 0.05, GYROTROPIC_SATURATED));
    else
      s.add_susceptibility(sphere, E_stuff,
                           gyrotropic_susceptibility(bias, 0.8, 0.04, 0.0,
                                                     cs == "gyro_drude_3d" ? GYROTROPIC_DRUDE : GYROTROPIC_LORENTZIAN));
    fields f(&s);
    gaussian_src_time src(0.6, 0.5);
    f.add_point_source(Ez, src, vec(0.7, 0.8, 0.9));
    f.add_point_source(Ex, src, vec(1.2, 0.9, 1.1));
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "noisy_lorentz_3d") {
    // noisy_lorentzian_susceptibility (src/susceptibility.cpp:317-339): the polarisation is driven
    // by the reference's own Gaussian random numbers, drawn on the host in the reference's loop
    // order, so a seeded run is reproducible point by point
    g_L = 1.6;
    set_random_seed(20251017);
    grid_volume gv = vol3d(1.6, 1.6, 1.2, a);
    structure s(gv, eps_box, pml(0.3), identity(), num_chunks);
    s.add_susceptibility(sphere, E_stuff, noisy_lorentzian_susceptibility(0.5, 0.9, 0.06));
    s.add_susceptibility(eps_box, E_stuff, lorentzian_susceptibility(1.3, 0.1));
    fields f(&s);
    f.use_real_fields();
    gaussian_src_time src(0.6, 0.5);
    f.add_point_source(Ez, src, vec(0.6, 0.8, 0.6));
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "lorentz_aniso_sigma") {
    // Lorentzian with an anisotropic (off-diagonal) sigma tensor: the OFFDIAG terms of update_P
    // (src/susceptibility.cpp:214-247) and the exchange of not-owned W values between chunks
    // (WE_stuff connections, src/boundaries.cpp:396-404)
    g_L = 2.0;
    grid_volume gv = vol3d(2.0, 2.0, 1.6, a);
    struct aniso_sigma : public material_function {
      virtual void sigma_row(component c, double sigrow[3], const vec &r) {
        const bool in = sphere(r) > 0;
        const int k = component_index(c);
        const double m[3][3] = {{1.0, 0.3, 0.1}, {0.3, 0.8, 0.2}, {0.1, 0.2, 1.2}};
        for (int j = 0; j < 3; ++j)
          sigrow[j] = in ? m[k][j] : 0.0;
      }
    } sig;
    structure s(gv, eps_box, pml(0.4), identity(), num_chunks);
    s.add_susceptibility(sig, E_stuff, lorentzian_susceptibility(0.9, 0.05));
    fields f(&s);
    gaussian_src_time src(0.6, 0.5);
    f.add_point_source(Ez, src, vec(0.7, 1.0, 0.8));
    f.add_point_source(Ex, src, vec(1.2, 0.9, 0.7));
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "c4_aniso_ring" || cs == "aniso_smooth") {
    // BASELINE config 4 (scaled twin): subpixel-smoothed Si ring -> off-diagonal chi1inv
    g_L = 2.4;
    grid_volume gv = vol3d(g_L, g_L, 1.2, a);
    structure s(gv, cs == "aniso_smooth" ? eps_smooth : eps_ring, pml(0.5), identity(), num_chunks,
                0.5, true, 1e-2, 2000);
    fields f(&s);
    f.use_real_fields();
    gaussian_src_time src(0.3, 0.2);
    src.is_integrated = false;
    f.add_point_source(Hz, src, vec(0.5 * g_L + 0.35 * g_L, 0.5 * g_L, 0.6));
    f.add_point_source(Ez, src, vec(0.5 * g_L, 0.5 * g_L + 0.35 * g_L, 0.6));
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "offdiag_2d") { // tests/pml.cpp offdiag_material with PML
    g_L = 3.0;
    grid_volume gv = voltwo(3.0, 3.0, a);
    offdiag_material mat(0.3);
    structure s(gv, mat, pml(0.8), identity(), num_chunks);
    fields f(&s);
    f.add_point_source(Ez, 0.25, 3.0, 0.0, 2.0, gv.center(), 1.0);
    f.add_point_source(Hz, 0.25, 3.0, 0.0, 2.0, gv.center(), 1.0);
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "cond_chi3_3d") { // conductivity (+PML => f_cond) and Kerr nonlinearity
    g_L = 2.0;
    grid_volume gv = vol3d(g_L, 1.6, 1.4, a);
    structure s(gv, eps_box, pml(0.5), identity(), num_chunks);
    s.set_conductivity(Dx, cond_slab);
    s.set_conductivity(Dy, cond_slab);
    s.set_conductivity(Dz, cond_slab);
    s.set_conductivity(By, cond_slab);
    s.set_chi3(chi3_box);
    fields f(&s);
    f.use_real_fields();
    gaussian_src_time src(0.3, 0.2);
    f.add_point_source(Ez, src, gv.center(), 5.0);
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_fields(f);
  }
  else if (cs == "dft_fields_3d") { // add_dft_fields volume monitor + decimation
    g_L = 2.0;
    grid_volume gv = vol3d(g_L, g_L, g_L, a);
    structure s(gv, eps_box, pml(0.5), identity(), num_chunks);
    fields f(&s);
    gaussian_src_time src(0.3, 0.2);
    src.is_integrated = false;
    f.add_point_source(Ey, src, gv.center());
    component comps[3] = {Ex, Ey, Hz};
    volume where(vec(0.6, 0.6, 0.6), vec(1.4, 1.4, 1.4));
    dft_fields df = f.add_dft_fields(comps, 3, where, 0.2, 0.4, 7);
    volume plane(vec(0.6, 0.6, 1.3), vec(1.4, 1.4, 1.3));
    dft_flux fl = f.add_dft_flux_plane(plane, 0.2, 0.4, 9);
    for (int i = 0; i < nsteps; ++i) f.step();
    probes(f, gv);
    dump_flux("flux.plane", fl);
    dump_fields(f);
  }
  else {
    fprintf(stderr, "unknown case %s\n", cs.c_str());
    return 2;
  }
  fclose(g_out);
  return 0;
}
