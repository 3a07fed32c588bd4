// loop_desc.hpp — reduce the reference's (grid_volume, ivec is, ivec ie) loop arguments to the
// plain integers of include/meep_b200.h, exactly as the loop macros do.
#ifndef MEEP_B200_LOOP_DESC_HPP
#define MEEP_B200_LOOP_DESC_HPP

#include "meep.hpp"
#include "../../include/meep_b200.h"

namespace meep_b200 {

// LOOP_OVER_IVECS prologue (reference src/meep/vec.hpp:151-166)
inline mb200_box_t make_box(const meep::grid_volume &gv, const meep::ivec &is,
                            const meep::ivec &ie) {
  mb200_box_t b;
  const meep::ivec rel = is - gv.little_corner();
  b.idx0 = 0;
  for (int k = 0; k < 3; ++k) {
    const meep::direction d = gv.yucky_direction(k);
    b.n[k] = (ie.yucky_val(k) - is.yucky_val(k)) / 2 + 1;
    b.s[k] = gv.stride(d);
    b.idx0 += (int64_t)(rel.yucky_val(k) / 2) * b.s[k];
  }
  b.reserved = 0;
  return b;
}

// KSTRIDE_DEF (reference src/meep_internals.hpp:217-221); dsig == NO_DIRECTION -> absent
inline mb200_pml_t make_pml(const meep::grid_volume &gv, const meep::ivec &is, meep::direction dsig,
                            const void *sig, const void *kap, const void *siginv) {
  mb200_pml_t p;
  p.sig = p.kap = p.siginv = nullptr;
  p.k0 = 0;
  p.ks[0] = p.ks[1] = p.ks[2] = 0;
  if (dsig == meep::NO_DIRECTION) return p;
  p.sig = sig;
  p.kap = kap;
  p.siginv = siginv;
  p.k0 = is.in_direction(dsig) - gv.little_corner().in_direction(dsig);
  for (int k = 0; k < 3; ++k)
    p.ks[k] = gv.yucky_direction(k) == dsig ? 2 : 0;
  return p;
}

} // namespace meep_b200
#endif
