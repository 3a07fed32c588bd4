// guards.cpp — the reference's CPU inner loops (src/step_generic.cpp and the generated
// step_generic_stride1.cpp) are interposed by definitions that abort: if any code path of an
// application linked against libmeep_b200 reaches a CPU stencil loop, it fails loudly instead
// of silently time-stepping stale host arrays.  (The device kernels are reached through
// include/meep_b200.h, not through these signatures.)
#include "meep.hpp"
#include "meep_internals.hpp"

namespace meep {

#define MB200_NO_CPU(name)                                                                         \
  meep::abort("meep_b200: " name " (CPU inner loop) was called — this build of libmeep has no "    \
              "CPU time-stepping path")

void step_curl(realnum *, component, const realnum *, const realnum *, ptrdiff_t, ptrdiff_t,
               const grid_volume &, const ivec, const ivec, realnum, direction, const realnum *,
               const realnum *, const realnum *, realnum *, direction, const realnum *,
               const realnum *, const realnum *, realnum, const realnum *, const realnum *,
               realnum *) {
  MB200_NO_CPU("step_curl");
}
void step_curl_stride1(realnum *, component, const realnum *, const realnum *, ptrdiff_t,
                       ptrdiff_t, const grid_volume &, const ivec, const ivec, realnum, direction,
                       const realnum *, const realnum *, const realnum *, realnum *, direction,
                       const realnum *, const realnum *, const realnum *, realnum,
                       const realnum *, const realnum *, realnum *) {
  MB200_NO_CPU("step_curl_stride1");
}
void step_update_EDHB(realnum *, component, const grid_volume &, const ivec, const ivec,
                      const realnum *, const realnum *, const realnum *, const realnum *,
                      const realnum *, const realnum *, ptrdiff_t, ptrdiff_t, ptrdiff_t,
                      const realnum *, const realnum *, realnum *, direction, const realnum *,
                      const realnum *) {
  MB200_NO_CPU("step_update_EDHB");
}
void step_update_EDHB_stride1(realnum *, component, const grid_volume &, const ivec, const ivec,
                              const realnum *, const realnum *, const realnum *, const realnum *,
                              const realnum *, const realnum *, ptrdiff_t, ptrdiff_t, ptrdiff_t,
                              const realnum *, const realnum *, realnum *, direction,
                              const realnum *, const realnum *) {
  MB200_NO_CPU("step_update_EDHB_stride1");
}
void step_beta(realnum *, component, const realnum *, const grid_volume &, const ivec, const ivec,
               realnum, direction, const realnum *, realnum *, direction, const realnum *,
               const realnum *, realnum *) {
  MB200_NO_CPU("step_beta");
}
void step_beta_stride1(realnum *, component, const realnum *, const grid_volume &, const ivec,
                       const ivec, realnum, direction, const realnum *, realnum *, direction,
                       const realnum *, const realnum *, realnum *) {
  MB200_NO_CPU("step_beta_stride1");
}
void step_bfast(realnum *, component, const realnum *, const realnum *, ptrdiff_t, ptrdiff_t,
                const grid_volume &, const ivec, const ivec, realnum, direction, const realnum *,
                const realnum *, const realnum *, realnum *, direction, const realnum *,
                const realnum *, const realnum *, realnum, const realnum *, const realnum *,
                realnum *, realnum *, realnum, realnum) {
  MB200_NO_CPU("step_bfast");
}
void step_bfast_stride1(realnum *, component, const realnum *, const realnum *, ptrdiff_t,
                        ptrdiff_t, const grid_volume &, const ivec, const ivec, realnum, direction,
                        const realnum *, const realnum *, const realnum *, realnum *, direction,
                        const realnum *, const realnum *, const realnum *, realnum,
                        const realnum *, const realnum *, realnum *, realnum *, realnum, realnum) {
  MB200_NO_CPU("step_bfast_stride1");
}

} // namespace meep
