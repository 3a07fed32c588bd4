// materials_fill.cpp — structure_chunk::set_chi1inv (reference src/anisotropic_averaging.cpp:211-328)
// for the case that dominates set-up of large cells: no subpixel averaging (maxeval = 0) and a
// material_function that keeps the default eff_chi1inv_row.  SURVEY 8f rank 1: "thread-safe/bulk
// material fill".
//
// In that case the reference's loop calls eff_chi1inv_row twice per grid point (each call builds a
// `volume`, makes two virtual calls and fills a 3-vector) only to obtain
//     chi1inv[c][d_c][i] = 1 / chi1p1(ft, centre of the pixel)        and 0 for the other two,
// allocates three arrays per component and deletes the two that turn out to be identically zero.
// Here the diagonal array is filled directly — one chi1p1 call per point, at the very point the
// reference evaluates (gv.dV(here, 1).center(), so a step-function material is classified
// identically) — over all host threads when the material is thread-safe, serially otherwise.
// Every other case (subpixel averaging, an overridden eff_chi1inv_row such as tests/pml.cpp's
// off-diagonal material) is forwarded to the reference's own definition.
#include <dlfcn.h>
#include <stdlib.h>
#include <atomic>
#include <thread>
#include <vector>

#include "meep.hpp"

using namespace std;

namespace meep {

namespace {
typedef void (*eff_row_fn)(material_function *, component, double *, const volume &, double, int);
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wpmf-conversions"
// has `medium` kept material_function's own eff_chi1inv_row?  (GCC bound-member-function extension)
bool uses_default_eff_chi1inv_row(material_function &medium) {
  const eff_row_fn base = (eff_row_fn)(&material_function::eff_chi1inv_row);
  const eff_row_fn dyn = (eff_row_fn)(medium.*(&material_function::eff_chi1inv_row));
  return base == dyn;
}
#pragma GCC diagnostic pop
} // namespace

void structure_chunk::set_chi1inv(component c, material_function &medium, bool use_anisotropic_averaging,
                                  double tol, int maxeval) {
  if (use_anisotropic_averaging || !uses_default_eff_chi1inv_row(medium) ||
      (getenv("MEEP_B200_BULK_FILL") && atoi(getenv("MEEP_B200_BULK_FILL")) == 0)) {
    typedef void (*fn)(structure_chunk *, component, material_function &, bool, double, int);
    static fn next = (fn)dlsym(RTLD_NEXT, "_ZN4meep15structure_chunk11set_chi1invENS_9componentERNS_17material_functionEbdi");
    if (!next) meep::abort("meep_b200: the reference's structure_chunk::set_chi1inv was not found behind the drop-in");
    next(this, c, medium, use_anisotropic_averaging, tol, maxeval);
    return;
  }
  if (!is_mine() || !gv.has_field(c)) return;
  const field_type ft = type(c);
  if (ft != E_stuff && ft != H_stuff) meep::abort("only E or H can have chi");
  medium.set_volume(gv.pad().surroundings());
  const double smoothing_diameter = 1.0;

  const direction dc = component_direction(c);
  direction ds[3] = {X, Y, Z};
  if (gv.dim == Dcyl) {
    ds[0] = R;
    ds[1] = P;
  }
  // the off-diagonal rows of a maxeval = 0 evaluation are identically zero: the reference allocates,
  // fills and deletes them (src/anisotropic_averaging.cpp:313-321); only their absence remains
  for (int k = 0; k < 3; ++k)
    if (ds[k] != dc) {
      delete[] chi1inv[c][ds[k]];
      chi1inv[c][ds[k]] = 0;
      trivial_chi1inv[c][ds[k]] = true;
    }
  if (!chi1inv[c][dc]) chi1inv[c][dc] = new realnum[gv.ntot()];
  realnum *out = chi1inv[c][dc];

  // geometry of LOOP_OVER_IVECS(gv, little_corner + iyee_shift(c), big_corner + iyee_shift(c))
  const ivec is = gv.little_corner() + gv.iyee_shift(c), ie = gv.big_corner() + gv.iyee_shift(c);
  const ptrdiff_t is_[3] = {is.yucky_val(0), is.yucky_val(1), is.yucky_val(2)};
  const ptrdiff_t nn[3] = {(ie.yucky_val(0) - is_[0]) / 2 + 1, (ie.yucky_val(1) - is_[1]) / 2 + 1,
                           (ie.yucky_val(2) - is_[2]) / 2 + 1};
  const direction dd[3] = {gv.yucky_direction(0), gv.yucky_direction(1), gv.yucky_direction(2)};
  const ptrdiff_t ss[3] = {gv.stride(dd[0]), gv.stride(dd[1]), gv.stride(dd[2])};
  const ivec rel = is - gv.little_corner();
  const ptrdiff_t idx0 = rel.yucky_val(0) / 2 * ss[0] + rel.yucky_val(1) / 2 * ss[1] + rel.yucky_val(2) / 2 * ss[2];

  std::atomic<ptrdiff_t> next_row(0);
  std::atomic<bool> all_trivial(true);
  const ptrdiff_t nrows = nn[0] * nn[1];
  auto worker = [&]() {
    bool trivial = true;
    const ptrdiff_t batch = 16;
    for (ptrdiff_t r0 = next_row.fetch_add(batch); r0 < nrows; r0 = next_row.fetch_add(batch))
      for (ptrdiff_t r = r0; r < r0 + batch && r < nrows; ++r) {
        const ptrdiff_t i1 = r / nn[1], i2 = r % nn[1];
        ivec here(gv.dim);
        here.set_direction(dd[0], is_[0] + 2 * i1);
        here.set_direction(dd[1], is_[1] + 2 * i2);
        for (ptrdiff_t i3 = 0; i3 < nn[2]; ++i3) {
          here.set_direction(dd[2], is_[2] + 2 * i3);
          const realnum v = 1 / medium.chi1p1(ft, gv.dV(here, smoothing_diameter).center());
          out[idx0 + i1 * ss[0] + i2 * ss[1] + i3 * ss[2]] = v;
          trivial = trivial && (v == realnum(1.0));
        }
      }
    if (!trivial) all_trivial = false;
  };
  int nthreads = 1;
  if (medium.is_thread_safe()) {
    nthreads = (int)std::thread::hardware_concurrency();
    if (const char *e = getenv("MEEP_B200_HOST_THREADS")) nthreads = atoi(e);
    else if (const char *e = getenv("OMP_NUM_THREADS")) nthreads = atoi(e);
    if (nthreads < 1) nthreads = 1;
    if ((ptrdiff_t)nthreads > nrows) nthreads = (int)(nrows > 0 ? nrows : 1);
  }
  if (nthreads == 1)
    worker();
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back(worker);
    for (std::thread &t : pool)
      t.join();
  }
  trivial_chi1inv[c][dc] = all_trivial;
  if (all_trivial) { // the whole tensor is trivial: no array at all (the kernels then copy D to E)
    delete[] chi1inv[c][dc];
    chi1inv[c][dc] = 0;
  }
  medium.unset_volume();
}

} // namespace meep
