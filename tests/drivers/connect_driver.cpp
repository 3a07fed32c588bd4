// connect_driver.cpp — dumps the chunk-connection tables (fields_chunk::connections_in/out,
// connection_phases, fields::comm_sizes, fields_chunk::zeroes) in a process-independent form so
// that tests can compare, entry by entry and in order, the tables built by the reference's
// connect_the_chunks / find_metals (src/boundaries.cpp:315-638; arm "ref") with the ones built by
// the drop-in's analytic replacement (meep_b200/host/connect.cpp; arm "emu"/"b200").
//
// Every pointer is written as (array id, element offset): array id = chunk * 1000 + kind * 100 +
// component * 2 + cmp for kind 0 = f, 1 = f_w; polarisation storage = chunk * 1000 + 900 + index
// of the polarisation in the chunk's list (offset relative to the start of its data block).
//
// usage: connect_driver <layout> <out.bin>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <complex>
#include <string>
#include <unordered_map>
#include <vector>

#include "meep.hpp"
using namespace meep;
using std::complex;

// fields::comm_sizes is private: reach it through an explicit-instantiation accessor (the
// reference's header stays unmodified); connections are current after a time step
typedef std::unordered_map<comms_key, size_t, comms_key_hash_fn> comm_sizes_t;
typedef comm_sizes_t fields::*comm_sizes_ptr;
comm_sizes_ptr get_comm_sizes_ptr();
template <comm_sizes_ptr P> struct comm_sizes_access {
  friend comm_sizes_ptr get_comm_sizes_ptr() { return P; }
};
template struct comm_sizes_access<&fields::comm_sizes>;

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

static double g_L = 1;
static double one(const vec &) { return 1.0; }
static double eps_box(const vec &r) {
  double m = 0;
  LOOP_OVER_DIRECTIONS(r.dim, d) {
    double x = fabs(r.in_direction(d) - 0.5 * g_L);
    if (x > m) m = x;
  }
  return m < 0.25 * g_L ? 12.0 : 1.0;
}
static double sphere(const vec &r) {
  double r2 = 0;
  LOOP_OVER_DIRECTIONS(r.dim, d) {
    double x = r.in_direction(d) - 0.5 * g_L;
    r2 += x * x;
  }
  return r2 < 0.09 * g_L * g_L ? 1.0 : 0.0;
}

struct pol_layout { // common head of the reference's lorentzian_data / gyrotropy_data blocks
  size_t sz_data, ntot;
};

// (array id, offset) of a pointer into one of chunk `i`'s arrays; ids are doubles so that one
// record holds both
static void resolve(fields &f, int i, const realnum *p, double &id, double &off) {
  fields_chunk *fc = f.chunks[i];
  const size_t n = fc->gv.ntot();
  FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) {
    if (fc->f[c][cmp] && p >= fc->f[c][cmp] && p < fc->f[c][cmp] + n) {
      // H may alias B: name it by the first component that matches (deterministic in both arms)
      id = i * 1000.0 + c * 2 + cmp;
      off = (double)(p - fc->f[c][cmp]);
      return;
    }
  }
  FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) {
    if (fc->f_w[c][cmp] && p >= fc->f_w[c][cmp] && p < fc->f_w[c][cmp] + n) {
      id = i * 1000.0 + 100 + c * 2 + cmp;
      off = (double)(p - fc->f_w[c][cmp]);
      return;
    }
  }
  FOR_FIELD_TYPES(ft) {
    int ip = 0;
    for (polarization_state *ps = fc->pol[ft]; ps; ps = ps->next, ++ip)
      if (ps->data) {
        const pol_layout *d = (const pol_layout *)ps->data;
        const char *b = (const char *)ps->data;
        if ((const char *)p >= b && (const char *)p < b + d->sz_data) {
          id = i * 1000.0 + 900 + ft * 10 + ip;
          off = (double)((const char *)p - b);
          return;
        }
      }
  }
  id = -1;
  off = -1;
}

static void dump_tables(fields &f) {
  char nm[160];
  // comm_sizes, sorted by key
  std::vector<std::vector<double> > cs;
  for (const auto &kv : f.*get_comm_sizes_ptr())
    cs.push_back({(double)kv.first.ft, (double)kv.first.phase, (double)kv.first.pair.first,
                  (double)kv.first.pair.second, (double)kv.second});
  std::sort(cs.begin(), cs.end());
  std::vector<double> flat;
  for (auto &r : cs) flat.insert(flat.end(), r.begin(), r.end());
  dump("comm_sizes", flat.data(), sizeof(double), flat.size());
  for (int i = 0; i < f.num_chunks; ++i) {
    fields_chunk *fc = f.chunks[i];
    if (!fc->is_mine()) continue;
    for (int dir = 0; dir < 2; ++dir) {
      const auto &tab = dir == 0 ? fc->connections_in : fc->connections_out;
      for (const auto &kv : tab) {
        const comms_key &k = kv.first;
        std::vector<double> rec;
        const int other = dir == 0 ? i : i; // both tables point into chunk i's own arrays
        for (realnum *p : kv.second) {
          double id, off;
          resolve(f, other, p, id, off);
          rec.push_back(id);
          rec.push_back(off);
        }
        snprintf(nm, sizeof nm, "chunk%d.%s.ft%d.ph%d.pair%d_%d", i, dir == 0 ? "in" : "out", (int)k.ft,
                 (int)k.phase, k.pair.first, k.pair.second);
        dump(nm, rec.data(), sizeof(double), rec.size());
      }
    }
    for (const auto &kv : fc->connection_phases) {
      const comms_key &k = kv.first;
      std::vector<double> rec;
      for (const complex<realnum> &z : kv.second) {
        rec.push_back(z.real());
        rec.push_back(z.imag());
      }
      snprintf(nm, sizeof nm, "chunk%d.phases.ft%d.ph%d.pair%d_%d", i, (int)k.ft, (int)k.phase, k.pair.first,
               k.pair.second);
      dump(nm, rec.data(), sizeof(double), rec.size());
    }
    FOR_FIELD_TYPES(ft) {
      std::vector<double> rec;
      for (size_t z = 0; z < fc->num_zeroes[ft]; ++z) {
        double id, off;
        resolve(f, i, fc->zeroes[ft][z], id, off);
        rec.push_back(id);
        rec.push_back(off);
      }
      snprintf(nm, sizeof nm, "chunk%d.zeroes.ft%d", i, (int)ft);
      dump(nm, rec.data(), sizeof(double), rec.size());
    }
  }
}

int main(int argc, char **argv) {
  initialize mpi(argc, argv);
  verbosity = 0;
  if (argc < 3) {
    fprintf(stderr, "usage: %s <layout> <out.bin>\n", argv[0]);
    return 2;
  }
  const std::string cs = argv[1];
  std::string outname = argv[2];
  if (count_processors() > 1) outname += ".rank" + std::to_string(my_rank());
  g_out = fopen(outname.c_str(), "wb");
  if (!g_out) { perror(argv[2]); return 2; }
  const double a = 10.0;
  const int nsteps = 2; // lazily allocated arrays (split H, f_w, polarisation blocks) exist afterwards

  if (cs == "pml27" || cs == "pml64" || cs == "pml72" || cs == "metal5" || cs == "pml27_complex") {
    // SURVEY 8e layouts: 1 leaf -> 27 chunks; 8 leaves 2x2x2 -> 64; 8 leaves 4x2x1 of a flat cell -> 72
    g_L = cs == "pml72" ? 6.4 : 3.2;
    grid_volume gv = cs == "pml72" ? vol3d(6.4, 6.4, 2.4, a) : vol3d(g_L, g_L, g_L, a);
    const int nchunks = (cs == "pml64" || cs == "pml72") ? 8 : (cs == "metal5" ? 5 : 1);
    structure s(gv, eps_box, cs == "metal5" ? no_pml() : pml(1.0), identity(), nchunks);
    fields f(&s);
    if (cs != "pml27_complex") f.use_real_fields();
    gaussian_src_time src(0.15, 0.1);
    f.add_point_source(Ez, src, gv.center() + vec(0.05, 0.05, 0.05));
    f.add_point_source(Hy, src, gv.center());
    for (int i = 0; i < nsteps; ++i) f.step();
    dump_tables(f);
  }
  else if (cs == "bloch4" || cs == "periodic_k0" || cs == "xperiodic_ypml") {
    g_L = 1.0;
    grid_volume gv = cs == "xperiodic_ypml" ? vol3d(1.0, 3.0, 1.0, a) : vol3d(1.2, 1.0, 0.8, a);
    structure s(gv, eps_box, cs == "xperiodic_ypml" ? pml(1.0, Y) : no_pml(), identity(), cs == "bloch4" ? 4 : 3);
    fields f(&s);
    f.add_point_source(Ez, 0.2, 3.0, 0.0, 2.0, gv.center(), complex<double>(0, -2 * pi * 0.2));
    if (cs == "bloch4") f.use_bloch(vec(0.3, 0.5, 0.8));
    else if (cs == "periodic_k0") f.use_bloch(vec(0.0, 0.0, 0.0)); // all phases 1: COPY connections through the wrap
    else f.use_bloch(X, 0.1);
    for (int i = 0; i < nsteps; ++i) f.step();
    dump_tables(f);
  }
  else if (cs == "mirror2d" || cs == "rotate3d") {
    if (cs == "mirror2d") {
      grid_volume gv = voltwo(2.0, 1.6, a);
      const symmetry S = mirror(X, gv) + mirror(Y, gv);
      structure s(gv, one, pml(0.4), S, 3);
      fields f(&s);
      f.add_point_source(Ez, 0.7, 2.5, 0.0, 4.0, gv.center());
      f.add_point_source(Hz, 0.6, 2.0, 0.0, 4.0, vec(0.7, 0.55));
      for (int i = 0; i < nsteps; ++i) f.step();
      dump_tables(f);
    }
    else {
      grid_volume gv = vol3d(1.2, 1.2, 1.0, a);
      const symmetry S = rotate4(Z, gv);
      structure s(gv, one, no_pml(), S, 2);
      fields f(&s);
      f.add_point_source(Ez, 0.7, 2.5, 0.0, 4.0, gv.center());
      f.use_bloch(vec(0.0, 0.0, 0.2));
      for (int i = 0; i < nsteps; ++i) f.step();
      dump_tables(f);
    }
  }
  else if (cs == "cyl3") {
    grid_volume gv = volcyl(2.0, 2.6, a);
    structure s(gv, one, pml(0.5), identity(), 3);
    fields f(&s, 1.0);
    gaussian_src_time src(0.35, 0.3);
    f.add_point_source(Ep, src, veccyl(0.45, 1.2));
    for (int i = 0; i < nsteps; ++i) f.step();
    dump_tables(f);
  }
  else if (cs == "gyro3" || cs == "aniso_sigma4") {
    // polarisation storage with not-owned internal values (gyrotropic) / not-owned W (anisotropic sigma)
    g_L = 2.0;
    grid_volume gv = vol3d(2.0, 1.6, 2.0, a);
    structure s(gv, eps_box, pml(0.4), identity(), cs == "gyro3" ? 3 : 4);
    struct aniso_sigma : public material_function {
      virtual void sigma_row(component c, double sigrow[3], const vec &r) {
        const bool in = sphere(r) > 0;
        const int k = component_index(c);
        const double m[3][3] = {{1.0, 0.3, 0.1}, {0.3, 0.8, 0.2}, {0.1, 0.2, 1.2}};
        for (int j = 0; j < 3; ++j) sigrow[j] = in ? m[k][j] : 0.0;
      }
    } sig;
    if (cs == "gyro3")
      s.add_susceptibility(sphere, E_stuff, gyrotropic_susceptibility(vec(0.3, -0.2, 0.9), 0.8, 0.04, 0.0, GYROTROPIC_LORENTZIAN));
    else
      s.add_susceptibility(sig, E_stuff, lorentzian_susceptibility(0.9, 0.05));
    fields f(&s);
    gaussian_src_time src(0.6, 0.5);
    f.add_point_source(Ez, src, vec(0.7, 0.8, 0.9));
    for (int i = 0; i < nsteps; ++i) f.step();
    dump_tables(f);
  }
  else if (cs == "bend2d") {
    grid_volume gv = voltwo(8, 16, a);
    gv.center_origin();
    structure s(gv, one, pml(1.0), identity(), 4);
    fields f(&s);
    f.use_real_fields();
    gaussian_src_time src(0.15, 0.1);
    f.add_point_source(Ez, src, vec(-2.0, -3.5));
    for (int i = 0; i < nsteps; ++i) f.step();
    dump_tables(f);
  }
  else if (cs == "1d3") {
    grid_volume gv = volone(10.0, a);
    structure s(gv, one, pml(1.0), identity(), 3);
    fields f(&s);
    f.add_point_source(Ex, 0.2, 3.0, 0.0, 2.0, gv.center());
    for (int i = 0; i < nsteps; ++i) f.step();
    dump_tables(f);
  }
  else {
    fprintf(stderr, "unknown layout %s\n", cs.c_str());
    return 2;
  }
  fclose(g_out);
  return 0;
}
