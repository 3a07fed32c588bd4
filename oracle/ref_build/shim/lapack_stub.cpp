/* LAPACK is absent in this image; only multilevel-atom.cpp (out of scope) calls
 * these.  Abort if ever reached.  TEST/BASELINE INFRASTRUCTURE ONLY. */
#include <cstdio>
#include <cstdlib>
extern "C" {
#define STUB(name)                                                                                 \
  void name() {                                                                                    \
    std::fprintf(stderr, "oracle/_ref: LAPACK routine " #name " is not available\n");              \
    std::abort();                                                                                  \
  }
STUB(dgetrf_)
STUB(dgetri_)
STUB(sgetrf_)
STUB(sgetri_)
}
