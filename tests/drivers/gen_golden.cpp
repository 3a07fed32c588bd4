// gen_golden.cpp — produces the golden vectors under tests/golden/ by calling the REFERENCE's own
// inner-loop functions (meep::step_curl, meep::step_update_EDHB,
// meep::lorentzian_susceptibility::update_P, meep::dft_chunk::update_dft) from the unmodified
// reference build in oracle/_ref on small seeded inputs.  Linked against the reference ONLY.
//
// Each case file holds: the initial arrays, the loop descriptors as derived by
// meep_b200/host/loop_desc.hpp (stored as float64 vectors), scalars, and the arrays after the call.
// tests/test_oracle.py replays the case through oracle/fdtd_oracle.c, tests/test_kernels_gpu.py
// through the CUDA C ABI.
//
// usage: gen_golden_ref_<prec> <outdir>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <complex>
#include <string>
#include <vector>

#include "meep.hpp"
#include "meep_internals.hpp"
#include "loop_desc.hpp"

using namespace meep;
using namespace meep_b200;
using std::complex;

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
static void dump_d(const std::string &name, const std::vector<double> &v) {
  dump(name, v.data(), 8, v.size());
}
static void dump_r(const std::string &name, const realnum *p, size_t n) {
  if (p) dump(name, p, sizeof(realnum), n);
}
static void dump_box(const std::string &name, const mb200_box_t &b) {
  dump_d(name, {(double)b.idx0, (double)b.s[0], (double)b.s[1], (double)b.s[2], (double)b.n[0],
                (double)b.n[1], (double)b.n[2]});
}
static void dump_pml(const std::string &name, const mb200_pml_t &p, bool present) {
  dump_d(name, {present ? 1.0 : 0.0, (double)p.k0, (double)p.ks[0], (double)p.ks[1],
                (double)p.ks[2]});
}

// deterministic pseudo-random numbers (LCG), independent of libc
static unsigned long long g_seed = 12345;
static double urand() {
  g_seed = g_seed * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)((g_seed >> 11) & ((1ULL << 53) - 1)) / (double)(1ULL << 53);
}
static realnum *rand_array(size_t n, double lo = -1, double hi = 1) {
  realnum *a = new realnum[n];
  for (size_t i = 0; i < n; ++i)
    a[i] = (realnum)(lo + (hi - lo) * urand());
  return a;
}

static void open_case(const std::string &dir, const std::string &name) {
  std::string path = dir + "/" + name + (sizeof(realnum) == 8 ? "_f64.bin" : "_f32.bin");
  g_out = fopen(path.c_str(), "wb");
  if (!g_out) { perror(path.c_str()); exit(2); }
}

// ---- step_curl ---------------------------------------------------------------------------------
static void gen_curl(const std::string &dir, const grid_volume &gv, const char *tag) {
  const size_t n = gv.ntot();
  // D-type update of the component along the first field direction: reads the low neighbours
  const component cc = gv.dim == D3 ? Dx : Dz; // 2D: TM (Dz from Hx,Hy)
  direction d1, d2;
  if (gv.dim == D3) { d1 = Y; d2 = Z; } else { d1 = X; d2 = Y; }
  const ivec is = gv.little_owned_corner0(cc), ie = gv.big_corner();
  const direction dsig0 = gv.dim == D3 ? Y : X, dsigu0 = gv.dim == D3 ? Z : Y;
  for (int variant = 0; variant < 16; ++variant) {
    const bool PML = variant & 8, FU = variant & 4, CND = variant & 2, G2 = variant & 1;
    char nm[64];
    snprintf(nm, sizeof nm, "curl_%s_v%02d", tag, variant);
    open_case(dir, nm);
    realnum *f = rand_array(n), *g1 = rand_array(n), *g2 = G2 ? rand_array(n) : NULL;
    realnum *fu = FU ? rand_array(n) : NULL, *fcnd = (CND && PML) ? rand_array(n) : NULL;
    realnum *cnd = CND ? rand_array(n, 0, 2) : NULL, *cndinv = CND ? rand_array(n, 0.5, 1) : NULL;
    const int ns = 2 * gv.num_direction(dsig0) + 2, nsu = 2 * gv.num_direction(dsigu0) + 2;
    realnum *sig = rand_array(ns, 0, 0.5), *kap = rand_array(ns, 1, 2), *siginv = rand_array(ns, 0.3, 1);
    realnum *sigu = rand_array(nsu, 0, 0.5), *kapu = rand_array(nsu, 1, 2), *siginvu = rand_array(nsu, 0.3, 1);
    const ptrdiff_t s1 = -gv.stride(d1), s2 = -gv.stride(d2);
    const realnum dtdx = 0.5, dt = 0.05;
    const direction dsig = PML ? dsig0 : NO_DIRECTION, dsigu = FU ? dsigu0 : NO_DIRECTION;
    dump_r("in.f", f, n); dump_r("in.g1", g1, n); dump_r("in.g2", g2, n);
    dump_r("in.fu", fu, n); dump_r("in.fcnd", fcnd, n); dump_r("in.cnd", cnd, n);
    dump_r("in.cndinv", cndinv, n);
    dump_r("in.sig", sig, ns); dump_r("in.kap", kap, ns); dump_r("in.siginv", siginv, ns);
    dump_r("in.sigu", sigu, nsu); dump_r("in.kapu", kapu, nsu); dump_r("in.siginvu", siginvu, nsu);
    dump_box("box", make_box(gv, is, ie));
    dump_pml("pml", make_pml(gv, is, dsig, sig, kap, siginv), PML);
    dump_pml("pmlu", make_pml(gv, is, dsigu, sigu, kapu, siginvu), FU);
    dump_d("scalars", {(double)s1, (double)s2, (double)dtdx, (double)dt});
    // THE REFERENCE CALL (dispatch macro picks the stride-1 build exactly as step_db does)
    STEP_CURL(f, cc, g1, g2, s1, s2, gv, is, ie, dtdx, dsig, sig, kap, siginv, fu, dsigu, sigu, kapu,
              siginvu, dt, cnd, cndinv, fcnd);
    dump_r("out.f", f, n); dump_r("out.fu", fu, n); dump_r("out.fcnd", fcnd, n);
    fclose(g_out);
  }
}

// ---- step_bfast ----------------------------------------------------------------------------------
static void gen_bfast(const std::string &dir, const grid_volume &gv, const char *tag) {
  const size_t n = gv.ntot();
  const component cc = gv.dim == D3 ? Dx : Dz;
  direction d1, d2;
  if (gv.dim == D3) { d1 = Y; d2 = Z; } else { d1 = X; d2 = Y; }
  const ivec is = gv.little_owned_corner0(cc), ie = gv.big_corner();
  const direction dsig0 = gv.dim == D3 ? Y : X, dsigu0 = gv.dim == D3 ? Z : Y;
  for (int variant = 0; variant < 16; ++variant) {
    const bool PML = variant & 8, FU = variant & 4, CND = variant & 2, G2 = variant & 1;
    char nm[64];
    snprintf(nm, sizeof nm, "bfast_%s_v%02d", tag, variant);
    open_case(dir, nm);
    realnum *f = rand_array(n), *g1 = rand_array(n), *g2 = G2 ? rand_array(n) : NULL, *F = rand_array(n);
    realnum *fu = FU ? rand_array(n) : NULL, *fcnd = (CND && PML) ? rand_array(n) : NULL;
    realnum *cnd = CND ? rand_array(n, 0, 2) : NULL, *cndinv = CND ? rand_array(n, 0.5, 1) : NULL;
    const int ns = 2 * gv.num_direction(dsig0) + 2, nsu = 2 * gv.num_direction(dsigu0) + 2;
    realnum *sig = rand_array(ns, 0, 0.5), *kap = rand_array(ns, 1, 2), *siginv = rand_array(ns, 0.3, 1);
    realnum *sigu = rand_array(nsu, 0, 0.5), *kapu = rand_array(nsu, 1, 2), *siginvu = rand_array(nsu, 0.3, 1);
    const ptrdiff_t s1 = -gv.stride(d1), s2 = -gv.stride(d2);
    const realnum dtdx = 0.5, dt = 0.05, k1 = 0.173, k2 = -0.291;
    const direction dsig = PML ? dsig0 : NO_DIRECTION, dsigu = FU ? dsigu0 : NO_DIRECTION;
    dump_r("in.f", f, n); dump_r("in.g1", g1, n); dump_r("in.g2", g2, n); dump_r("in.F", F, n);
    dump_r("in.fu", fu, n); dump_r("in.fcnd", fcnd, n); dump_r("in.cnd", cnd, n);
    dump_r("in.cndinv", cndinv, n);
    dump_r("in.siginv", siginv, ns); dump_r("in.siginvu", siginvu, nsu);
    dump_box("box", make_box(gv, is, ie));
    dump_pml("pml", make_pml(gv, is, dsig, sig, kap, siginv), PML);
    dump_pml("pmlu", make_pml(gv, is, dsigu, sigu, kapu, siginvu), FU);
    dump_d("scalars", {(double)s1, (double)s2, (double)k1, (double)k2});
    // THE REFERENCE CALL
    STEP_BFAST(f, cc, g1, g2, s1, s2, gv, is, ie, dtdx, dsig, sig, kap, siginv, fu, dsigu, sigu, kapu,
               siginvu, dt, cnd, cndinv, fcnd, F, k1, k2);
    dump_r("out.f", f, n); dump_r("out.fu", fu, n); dump_r("out.fcnd", fcnd, n); dump_r("out.F", F, n);
    fclose(g_out);
  }
}

// ---- step_beta -----------------------------------------------------------------------------------
static void gen_beta(const std::string &dir, const grid_volume &gv, const char *tag) {
  const size_t n = gv.ntot();
  const component cc = Dx;
  const ivec is = gv.little_owned_corner0(cc), ie = gv.big_corner();
  const direction dsig0 = Y, dsigu0 = X; // any two in-plane directions exercise both k lookups
  for (int variant = 0; variant < 8; ++variant) {
    const bool PML = variant & 4, FU = variant & 2, CND = variant & 1;
    char nm[64];
    snprintf(nm, sizeof nm, "beta_%s_v%d", tag, variant);
    open_case(dir, nm);
    realnum *f = rand_array(n), *g = rand_array(n);
    realnum *fu = FU ? rand_array(n) : NULL, *fcnd = (CND && PML) ? rand_array(n) : NULL;
    realnum *cndinv = CND ? rand_array(n, 0.5, 1) : NULL;
    const int ns = 2 * gv.num_direction(dsig0) + 2, nsu = 2 * gv.num_direction(dsigu0) + 2;
    realnum *siginv = rand_array(ns, 0.3, 1), *siginvu = rand_array(nsu, 0.3, 1);
    const realnum betadt = 0.0371;
    const direction dsig = PML ? dsig0 : NO_DIRECTION, dsigu = FU ? dsigu0 : NO_DIRECTION;
    dump_r("in.f", f, n); dump_r("in.g", g, n); dump_r("in.fu", fu, n); dump_r("in.fcnd", fcnd, n);
    dump_r("in.cndinv", cndinv, n); dump_r("in.siginv", siginv, ns); dump_r("in.siginvu", siginvu, nsu);
    dump_box("box", make_box(gv, is, ie));
    dump_pml("pml", make_pml(gv, is, dsig, NULL, NULL, siginv), PML);
    dump_pml("pmlu", make_pml(gv, is, dsigu, NULL, NULL, siginvu), FU);
    dump_d("scalars", {(double)betadt});
    STEP_BETA(f, cc, g, gv, is, ie, betadt, dsig, siginv, fu, dsigu, siginvu, cndinv, fcnd);
    dump_r("out.f", f, n); dump_r("out.fu", fu, n); dump_r("out.fcnd", fcnd, n);
    fclose(g_out);
  }
}

// ---- step_update_EDHB --------------------------------------------------------------------------
static void gen_edhb(const std::string &dir, const grid_volume &gv, const char *tag) {
  const size_t n = gv.ntot();
  const component ec = gv.dim == D3 ? Ey : Ex;
  const direction d_ec = component_direction(ec);
  const direction d_1 = cycle_direction(gv.dim, d_ec, 1), d_2 = cycle_direction(gv.dim, d_ec, 2);
  const ivec is = gv.little_owned_corner0(ec), ie = gv.big_corner();
  // (has_u, noff, chi3, pml)
  const int cfgs[][4] = {{1, 0, 0, 0}, {0, 0, 0, 0}, {1, 1, 0, 0}, {1, 2, 0, 0}, {1, 0, 1, 0}, {1, 1, 1, 0},
                         {1, 2, 1, 0}, {1, 0, 0, 1}, {0, 0, 0, 1}, {1, 1, 0, 1}, {1, 2, 0, 1}, {1, 2, 1, 1},
                         {1, 0, 1, 1}};
  for (size_t k = 0; k < sizeof(cfgs) / sizeof(cfgs[0]); ++k) {
    int has_u = cfgs[k][0], noff = cfgs[k][1], nl = cfgs[k][2], pmlw = cfgs[k][3];
    if (gv.dim == D2 && noff == 2) noff = 1; // only one in-plane partner in 2-D
    char nm[64];
    snprintf(nm, sizeof nm, "edhb_%s_c%02d", tag, (int)k);
    open_case(dir, nm);
    realnum *f = rand_array(n), *g = rand_array(n);
    // when chi3 is present the reference also reads g1/g2 neighbours if they exist
    realnum *g1 = (noff >= 1 || nl) ? rand_array(n) : NULL;
    realnum *g2 = (noff >= 2 || (nl && gv.dim == D3)) ? rand_array(n) : NULL;
    realnum *u = has_u ? rand_array(n, 0.1, 1) : NULL;
    realnum *u1 = noff >= 1 ? rand_array(n, -0.1, 0.1) : NULL, *u2 = noff >= 2 ? rand_array(n, -0.1, 0.1) : NULL;
    realnum *chi2 = nl ? rand_array(n, 0, 0.1) : NULL, *chi3 = nl ? rand_array(n, 0, 0.1) : NULL;
    realnum *fw = pmlw ? rand_array(n) : NULL;
    const int ns = 2 * gv.num_direction(d_ec) + 2;
    realnum *sigw = rand_array(ns, 0, 0.5), *kapw = rand_array(ns, 1, 2);
    const ptrdiff_t s = gv.stride(d_ec), s1 = gv.stride(d_1), s2 = gv.stride(d_2);
    const direction dsigw = pmlw ? d_ec : NO_DIRECTION;
    dump_r("in.f", f, n); dump_r("in.g", g, n); dump_r("in.g1", g1, n); dump_r("in.g2", g2, n);
    dump_r("in.u", u, n); dump_r("in.u1", u1, n); dump_r("in.u2", u2, n);
    dump_r("in.chi2", chi2, n); dump_r("in.chi3", chi3, n); dump_r("in.fw", fw, n);
    dump_r("in.sigw", sigw, ns); dump_r("in.kapw", kapw, ns);
    dump_box("box", make_box(gv, is, ie));
    dump_pml("pmlw", make_pml(gv, is, dsigw, sigw, kapw, NULL), pmlw);
    dump_d("scalars", {(double)s, (double)s1, (double)s2});
    STEP_UPDATE_EDHB(f, ec, gv, is, ie, g, g1, g2, u, u1, u2, s, s1, s2, chi2, chi3, fw, dsigw, sigw, kapw);
    dump_r("out.f", f, n); dump_r("out.fw", fw, n);
    fclose(g_out);
  }
}

// ---- lorentzian_susceptibility::update_P ---------------------------------------------------------
struct lorentzian_data_layout {
  size_t sz_data;
  size_t ntot;
  realnum *P[NUM_FIELD_COMPONENTS][2];
  realnum *P_prev[NUM_FIELD_COMPONENTS][2];
  realnum data[1];
};

static void gen_lorentz(const std::string &dir, const grid_volume &gv, const char *tag) {
  const size_t n = gv.ntot();
  for (int noff = 0; noff <= 2; ++noff)
    for (int drude = 0; drude <= 1; ++drude) {
      char nm[64];
      snprintf(nm, sizeof nm, "lorentz_%s_o%d_d%d", tag, noff, drude);
      open_case(dir, nm);
      lorentzian_susceptibility sus(0.7, 0.13, drude != 0);
      const component c = Ey;
      const direction d = component_direction(c);
      const direction d1 = cycle_direction(gv.dim, d, 1), d2 = cycle_direction(gv.dim, d, 2);
      realnum *W[NUM_FIELD_COMPONENTS][2];
      FOR_COMPONENTS(cc) { W[cc][0] = W[cc][1] = NULL; }
      FOR_ELECTRIC_COMPONENTS(cc) { W[cc][0] = rand_array(n); }
      sus.ntot = n;
      sus.sigma[c][d] = rand_array(n, 0, 1);
      for (size_t i = 0; i < n; i += 7) sus.sigma[c][d][i] = 0; // exercise the s[i] != 0 guard
      sus.trivial_sigma[c][d] = false;
      if (noff >= 1) { sus.sigma[c][d1] = rand_array(n, -0.2, 0.2); sus.trivial_sigma[c][d1] = false; }
      if (noff >= 2) { sus.sigma[c][d2] = rand_array(n, -0.2, 0.2); sus.trivial_sigma[c][d2] = false; }
      lorentzian_data_layout *data = (lorentzian_data_layout *)sus.new_internal_data(W, gv);
      sus.init_internal_data(W, 0.05, gv, data);
      if (!data->P[c][0]) { fprintf(stderr, "gen_lorentz: no P allocated\n"); exit(3); }
      for (size_t i = 0; i < n; ++i) {
        data->P[c][0][i] = (realnum)(2 * urand() - 1);
        data->P_prev[c][0][i] = (realnum)(2 * urand() - 1);
      }
      const realnum dt = 0.05;
      // constants exactly as src/susceptibility.cpp:192-195, in realnum arithmetic
      const realnum omega_0 = 0.7, gamma = 0.13;
      const realnum omega2pi = 2 * pi * omega_0, g2pi = gamma * 2 * pi;
      const realnum omega0dtsqr = omega2pi * omega2pi * dt * dt;
      const realnum gamma1inv = 1 / (1 + g2pi * dt / 2), gamma1 = (1 - g2pi * dt / 2);
      const realnum omega0dtsqr_denom = drude ? 0 : omega0dtsqr;
      dump_r("in.p", data->P[c][0], n); dump_r("in.pp", data->P_prev[c][0], n);
      dump_r("in.w", W[c][0], n); dump_r("in.s", sus.sigma[c][d], n);
      dump_r("in.w1", noff >= 1 ? W[direction_component(c, d1)][0] : NULL, n);
      dump_r("in.s1", sus.sigma[c][d1], n);
      dump_r("in.w2", noff >= 2 ? W[direction_component(c, d2)][0] : NULL, n);
      dump_r("in.s2", sus.sigma[c][d2], n);
      dump_box("box", make_box(gv, gv.little_owned_corner(c), gv.big_corner()));
      dump_d("scalars", {(double)gv.stride(d), (double)gv.stride(d1), (double)gv.stride(d2),
                         (double)gamma1inv, (double)gamma1, (double)omega0dtsqr, (double)omega0dtsqr_denom});
      sus.update_P(W, NULL, dt, gv, data);
      dump_r("out.p", data->P[c][0], n); dump_r("out.pp", data->P_prev[c][0], n);
      fclose(g_out);
      sus.delete_internal_data(data);
    }
}

// ---- gyrotropic_susceptibility::update_P -----------------------------------------------------------
struct gyro_layout { // src/susceptibility.cpp:374-380
  size_t sz_data, ntot;
  realnum *P[NUM_FIELD_COMPONENTS][2][3];
  realnum *P_prev[NUM_FIELD_COMPONENTS][2][3];
  realnum data[1];
};

static void gen_gyro(const std::string &dir, const grid_volume &gv) {
  const size_t n = gv.ntot();
  const gyrotropy_model models[3] = {GYROTROPIC_LORENTZIAN, GYROTROPIC_DRUDE, GYROTROPIC_SATURATED};
  const char *mname[3] = {"lorentz", "drude", "saturated"};
  for (int im = 0; im < 3; ++im)
    for (int nw = 1; nw <= 3; nw += 2) { // only the driving component present / all three
      char nm[64];
      snprintf(nm, sizeof nm, "gyro_%s_w%d", mname[im], nw);
      open_case(dir, nm);
      const vec bias(0.3, -0.2, 0.9);
      const realnum omega_0 = 0.7, gamma = 0.13, alpha = 0.05, dt = 0.05;
      gyrotropic_susceptibility sus(bias, omega_0, gamma, alpha, models[im]);
      const component c = Ey;
      const direction d0 = component_direction(c);
      const direction d1 = cycle_direction(gv.dim, d0, 1), d2 = cycle_direction(gv.dim, d0, 2);
      const direction ds[3] = {d0, d1, d2};
      realnum *W[NUM_FIELD_COMPONENTS][2];
      FOR_COMPONENTS(cc) { W[cc][0] = W[cc][1] = NULL; }
      W[c][0] = rand_array(n);
      if (nw == 3) {
        W[direction_component(c, d1)][0] = rand_array(n);
        W[direction_component(c, d2)][0] = rand_array(n);
      }
      sus.ntot = n;
      sus.sigma[c][d0] = rand_array(n, 0, 1);
      sus.trivial_sigma[c][d0] = false;
      gyro_layout *data = (gyro_layout *)sus.new_internal_data(W, gv);
      sus.init_internal_data(W, dt, gv, data);
      if (!data->P[c][0][0]) { fprintf(stderr, "gen_gyro: no P allocated\n"); exit(3); }
      for (int k = 0; k < 3; ++k)
        for (size_t i = 0; i < n; ++i) {
          data->P[c][0][k][i] = (realnum)(2 * urand() - 1);
          data->P_prev[c][0][k][i] = (realnum)(2 * urand() - 1);
        }
      // the reference's constants (src/susceptibility.cpp:449-474, 512-527), in realnum arithmetic
      const vec b = models[im] == GYROTROPIC_SATURATED ? bias / abs(bias) : bias;
      realnum gt[3][3];
      memset(gt, 0, sizeof gt);
      gt[X][Y] = b.z(); gt[Y][X] = -b.z(); gt[Y][Z] = b.x(); gt[Z][Y] = -b.x(); gt[Z][X] = b.y(); gt[X][Z] = -b.y();
      const realnum omega2pidt = 2 * pi * omega_0 * dt, g2pidt = 2 * pi * gamma * dt;
      realnum c4[4], gd, gx, gy, gz;
      if (models[im] != GYROTROPIC_SATURATED) {
        const realnum omega0dtsqr = omega2pidt * omega2pidt;
        const realnum gamma1 = (1 - g2pidt / 2);
        const realnum diag = 2 - (models[im] == GYROTROPIC_DRUDE ? 0 : omega0dtsqr);
        const realnum pt = pi * dt;
        gd = (1 + g2pidt / 2); gx = pt * gt[Y][Z]; gy = pt * gt[Z][X]; gz = pt * gt[X][Y];
        c4[0] = diag; c4[1] = gamma1; c4[2] = omega0dtsqr; c4[3] = pt;
      }
      else {
        const realnum dt2pi = 2 * pi * dt;
        gd = 0.5; gx = -0.5 * alpha * gt[Y][Z]; gy = -0.5 * alpha * gt[Z][X]; gz = -0.5 * alpha * gt[X][Y];
        c4[0] = omega2pidt; c4[1] = g2pidt; c4[2] = alpha; c4[3] = dt2pi;
      }
      const realnum invdet = 1.0 / gd / (gd * gd + gx * gx + gy * gy + gz * gz);
      const realnum inv[3][3] = {{invdet * (gd * gd + gx * gx), invdet * (gx * gy + gd * gz), invdet * (gx * gz - gd * gy)},
                                 {invdet * (gy * gx - gd * gz), invdet * (gd * gd + gy * gy), invdet * (gy * gz + gd * gx)},
                                 {invdet * (gz * gx + gd * gy), invdet * (gz * gy - gd * gx), invdet * (gd * gd + gz * gz)}};
      std::vector<double> gtr, invr;
      for (int a = 0; a < 3; ++a)
        for (int bb = 0; bb < 3; ++bb) {
          gtr.push_back(gt[ds[a]][ds[bb]]);
          invr.push_back(inv[ds[a]][ds[bb]]);
        }
      for (int k = 0; k < 3; ++k) {
        snprintf(nm, sizeof nm, "in.p%d", k); dump_r(nm, data->P[c][0][ds[k]], n);
        snprintf(nm, sizeof nm, "in.pp%d", k); dump_r(nm, data->P_prev[c][0][ds[k]], n);
      }
      dump_r("in.w0", W[c][0], n);
      dump_r("in.w1", W[direction_component(c, d1)][0], n);
      dump_r("in.w2", W[direction_component(c, d2)][0], n);
      dump_r("in.s", sus.sigma[c][d0], n);
      dump_box("box", make_box(gv, gv.little_owned_corner(c), gv.big_corner()));
      dump_d("scalars", {(double)gv.stride(d0), (double)gv.stride(d1), (double)gv.stride(d2), (double)c4[0],
                         (double)c4[1], (double)c4[2], (double)c4[3], (double)(im == 2 ? 1 : 0)});
      dump_d("gt", gtr);
      dump_d("inv", invr);
      sus.update_P(W, NULL, dt, gv, data); // THE REFERENCE CALL
      for (int k = 0; k < 3; ++k) {
        snprintf(nm, sizeof nm, "out.p%d", k); dump_r(nm, data->P[c][0][ds[k]], n);
        snprintf(nm, sizeof nm, "out.pp%d", k); dump_r(nm, data->P_prev[c][0][ds[k]], n);
      }
      fclose(g_out);
      sus.delete_internal_data(data);
    }
}

// ---- dft_chunk::update_dft -----------------------------------------------------------------------
static double one(const vec &) { return 1.0; }

static void gen_dft(const std::string &dir, bool three_d, bool complex_fields, const char *tag) {
  char nm[64];
  snprintf(nm, sizeof nm, "dft_%s", tag);
  open_case(dir, nm);
  grid_volume gv = three_d ? vol3d(0.9, 0.8, 0.7, 10) : voltwo(1.2, 0.9, 10);
  structure s(gv, one);
  fields f(&s);
  if (!complex_fields) f.use_real_fields();
  f.add_point_source(three_d ? Ex : Ez, 0.3, 1.0, 0.0, 2.0, gv.center());
  f.add_point_source(three_d ? Hy : Hx, 0.3, 1.0, 0.0, 2.0, gv.center());
  volume where = three_d ? volume(vec(0.2, 0.15, 0.35), vec(0.7, 0.65, 0.35))
                         : volume(vec(0.25, 0.45), vec(0.95, 0.45));
  dft_flux fl = f.add_dft_flux_plane(where, 0.2, 0.4, 5);
  // fill every allocated field array of every chunk with seeded random numbers
  for (int i = 0; i < f.num_chunks; ++i)
    FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) if (f.chunks[i]->f[c][cmp])
      for (size_t k = 0; k < f.chunks[i]->gv.ntot(); ++k)
        f.chunks[i]->f[c][cmp][k] = (realnum)(2 * urand() - 1);
  int nchunks = 0;
  const double time = 1.234;
  for (int i = 0; i < f.num_chunks; ++i) {
    fields_chunk *fc = f.chunks[i];
    for (dft_chunk *d = fc->dft_chunks; d; d = d->next_in_chunk, ++nchunks) {
      char p[64];
      snprintf(p, sizeof p, "k%d.", nchunks);
      const std::string P = p;
      const size_t nd = 2 * d->N * d->omega.size();
      for (size_t k = 0; k < nd / 2; ++k)
        d->dft[k] = complex<realnum>((realnum)(urand() - 0.5), (realnum)(urand() - 0.5));
      dump_r(P + "in.f_re", fc->f[d->c][0], fc->gv.ntot());
      dump_r(P + "in.f_im", fc->f[d->c][1], fc->gv.ntot());
      dump_r(P + "in.dft", (realnum *)d->dft, nd);
      dump_box(P + "box", make_box(fc->gv, d->is, d->ie));
      std::vector<double> w;
      for (int k = 0; k < 3; ++k) {
        const direction dk = fc->gv.yucky_direction(k);
        w.push_back(d->s0.in_direction(dk)); w.push_back(d->s1.in_direction(dk));
        w.push_back(d->e0.in_direction(dk)); w.push_back(d->e1.in_direction(dk));
      }
      dump_d(P + "weights", w);
      dump_d(P + "scalars", {(double)d->avg1, (double)d->avg2, d->dV0, d->dV1,
                             d->include_dV_and_interp_weights ? 1.0 : 0.0,
                             d->sqrt_dV_and_interp_weights ? 1.0 : 0.0, (double)d->omega.size()});
      d->update_dft(time); // THE REFERENCE CALL (fills dft_phase and accumulates)
      dump_r(P + "phase", (realnum *)d->dft_phase, 2 * d->omega.size());
      dump_r(P + "out.dft", (realnum *)d->dft, nd);
    }
  }
  dump_d("nchunks", {(double)nchunks});
  double *F = fl.flux();
  dump_d("flux", std::vector<double>(F, F + fl.freq.size()));
  delete[] F;
  fclose(g_out);
}

// ---- cylindrical step_db ---------------------------------------------------------------------------
// The cylindrical terms (src/step_db.cpp:86-122, 177-462) are inline in fields_chunk::step_db, so
// the golden vector is one whole fields::step_db(ft) of a one-chunk Dcyl cell with random arrays.
// fields::step_db is private: reach it through an explicit-instantiation accessor (no header is
// modified).  The record also holds the job parameters, derived with the reference's own tables
// (plus_component, gv.stride, little_owned_corner0, ...) so that the replay only wires arrays.
namespace {
typedef void (fields::*step_db_fn)(field_type);
template <step_db_fn F> struct step_db_access {
  friend step_db_fn get_step_db() { return F; }
};
step_db_fn get_step_db();
template struct step_db_access<&fields::step_db>;
} // namespace

static double one_eps(const vec &) { return 1.0; }

static void gen_cyl(const std::string &dir, double m, field_type ft, const char *tag) {
  grid_volume gv = volcyl(0.6, 0.5, 10); // 6 x 5 cells, origin at r = 0
  structure s(gv, one_eps, no_pml(), identity(), 1);
  fields f(&s, m);
  f.require_component(Ep);
  f.require_component(Hp);
  fields_chunk *fc = f.chunks[0];
  const size_t n = fc->gv.ntot();
  open_case(dir, std::string("cyl_") + tag);
  // random state (H aliases B without mu: fill every distinct array once)
  std::vector<realnum *> seen;
  FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) {
    realnum *a = fc->f[c][cmp];
    if (!a) continue;
    bool dup = false;
    for (realnum *q : seen) dup = dup || q == a;
    if (dup) continue;
    seen.push_back(a);
    for (size_t i = 0; i < n; ++i)
      a[i] = (realnum)(2 * urand() - 1);
  }
  char nm[64];
  FOR_COMPONENTS(c) for (int cmp = 0; cmp < 2; ++cmp) if (fc->f[c][cmp]) {
    snprintf(nm, sizeof nm, "in.f.%d.%d", (int)c, cmp);
    dump_r(nm, fc->f[c][cmp], n);
  }
  const grid_volume &g = fc->gv;
  const double Courant = fc->Courant, dt = fc->dt;
  const int nz = g.nz();
  dump_d("dims", {(double)g.nr(), (double)nz, m, Courant, dt, (double)(ft == D_stuff)});
  int k = 0, km = 0;
  FOR_FT_COMPONENTS(ft, cc) for (int cmp = 0; cmp < 2; ++cmp) if (fc->f[cc][cmp]) {
    const direction d_c = component_direction(cc);
    // the step plan (private tables of fields_chunk, src/fields.cpp:428-456) for (r, phi, z):
    // d(F_a)/dt takes +d/d(b) of the component along c and -d/d(c) of the component along b,
    // (a, b, c) a cyclic permutation of (R, P, Z)
    const direction cyc[3] = {R, P, Z};
    const int ia = d_c == R ? 0 : (d_c == P ? 1 : 2);
    const direction d_b = cyc[(ia + 1) % 3], d_cc = cyc[(ia + 2) % 3];
    const component base = ft == D_stuff ? Hr : Er;
    const component c_p = direction_component(base, d_cc), c_m = direction_component(base, d_b);
    const bool have_p = fc->f[c_p][0] != NULL, have_m = fc->f[c_m][0] != NULL;
    const direction dd_p = d_b, dd_m = d_cc;
    ptrdiff_t stride_p = have_p ? g.stride(dd_p) : 0;
    ptrdiff_t stride_m = have_m ? g.stride(dd_m) : 0;
    if (ft == D_stuff) { stride_p = -stride_p; stride_m = -stride_m; }
    int gp = have_p ? (int)c_p : -1, gm = have_m ? (int)c_m : -1, rderiv = 0;
    double ir0 = 0;
    if (d_c == R) gp = -1;
    if (d_c == Z) {
      gm = -1;
      rderiv = 1;
      ir0 = (realnum)(g.origin_r() * g.a + 0.5 * g.iyee_shift(c_p).in_direction(R));
    }
    const ivec is = g.little_owned_corner0(cc), ie = g.big_corner();
    const mb200_box_t b = make_box(g, is, ie);
    snprintf(nm, sizeof nm, "job.curl.%d", k++);
    dump_d(nm, {(double)cc, (double)cmp, (double)gp, (double)gm, (double)stride_p, (double)stride_m,
                (double)rderiv, ir0, (double)b.idx0, (double)b.s[0], (double)b.s[1], (double)b.s[2],
                (double)b.n[0], (double)b.n[1], (double)b.n[2]});
    if (m != 0 && (d_c == R || d_c == Z)) {
      const component c_g = d_c == R ? c_p : c_m;
      const realnum the_m =
          2 * m * (1 - 2 * cmp) * (1 - 2 * (ft == B_stuff)) * (1 - 2 * (d_c == R)) * Courant;
      snprintf(nm, sizeof nm, "job.mr.%d", km++);
      dump_d(nm, {(double)cc, (double)cmp, (double)c_g, (double)the_m, (double)is.yucky_val(1),
                  (double)b.idx0, (double)b.s[0], (double)b.s[1], (double)b.s[2], (double)b.n[0],
                  (double)b.n[1], (double)b.n[2]});
    }
  }
  // the r = 0 row boxes (src/step_db.cpp:297-299, 347-349)
  FOR_FT_COMPONENTS(ft, cc) {
    ivec is = g.little_owned_corner(cc), ie = g.big_owned_corner(cc);
    ie.set_direction(R, 0);
    snprintf(nm, sizeof nm, "box.r0.%d", (int)cc);
    dump_box(nm, make_box(g, is, ie));
  }
  (f.*get_step_db())(ft); // THE REFERENCE CALL
  FOR_FT_COMPONENTS(ft, cc) for (int cmp = 0; cmp < 2; ++cmp) if (fc->f[cc][cmp]) {
    snprintf(nm, sizeof nm, "out.f.%d.%d", (int)cc, cmp);
    dump_r(nm, fc->f[cc][cmp], n);
  }
  fclose(g_out);
}

int main(int argc, char **argv) {
  initialize mpi(argc, argv);
  verbosity = 0;
  if (argc < 2) { fprintf(stderr, "usage: %s <outdir>\n", argv[0]); return 2; }
  const std::string dir = argv[1];
  grid_volume g3 = vol3d(0.5, 0.4, 0.6, 10); // 5 x 4 x 6 cells
  grid_volume g2 = voltwo(0.7, 0.5, 10);     // 7 x 5 cells
  gen_curl(dir, g3, "3d");
  gen_curl(dir, g2, "2d");
  gen_beta(dir, g2, "2d");
  gen_edhb(dir, g3, "3d");
  gen_edhb(dir, g2, "2d");
  gen_lorentz(dir, g3, "3d");
  gen_dft(dir, true, false, "3d_real");
  gen_dft(dir, false, true, "2d_complex");
  gen_cyl(dir, 0, D_stuff, "m0_D");
  gen_cyl(dir, 0, B_stuff, "m0_B");
  gen_cyl(dir, 1, D_stuff, "m1_D");
  gen_cyl(dir, -1, B_stuff, "mneg1_B");
  gen_cyl(dir, 2, D_stuff, "m2_D");
  gen_cyl(dir, 3, B_stuff, "m3_B");
  gen_bfast(dir, g3, "3d");
  gen_bfast(dir, g2, "2d");
  gen_gyro(dir, g3);
  return 0;
}
