// hostmem.hpp — keeping the HOST copies of device-resident arrays out of RAM.
//
// The reference keeps every field array in host memory (fields_chunk::alloc_f,
// src/fields.cpp:480-504) and the unchanged meep.hpp API exposes those pointers, so the arrays
// must stay allocated.  But while the device copy is the authoritative one nothing reads them,
// and a 1024^3 run would pin ~100 GB of host RAM for nothing.  These helpers keep the address
// range and drop (or never create) the pages behind it:
//   * new_zeroed_lazily: `new realnum[n]` whose zero contents cost no page (the interior pages
//     are handed back with madvise(MADV_DONTNEED): private anonymous memory reads as zero-fill
//     on demand afterwards);
//   * interior_untouched: no page of the array's interior is resident or swapped, i.e. the array
//     still holds the zeros it was born with — its device twin is a cudaMemset, not a copy;
//   * release_interior: drop the resident pages of an array whose contents are stale.
#ifndef MEEP_B200_HOSTMEM_HPP
#define MEEP_B200_HOSTMEM_HPP

#include <stddef.h>

#include "meep.hpp"

namespace meep_b200 {

constexpr size_t kLazyMinBytes = 1 << 20; // smaller arrays are simply memset / copied

// [*lo, *hi) = the whole pages inside [p, p + bytes); false if there is none
bool page_interior(const void *p, size_t bytes, char **lo, char **hi);
meep::realnum *new_zeroed_lazily(size_t n);
void release_interior(void *p, size_t bytes);
bool interior_untouched(const void *p, size_t bytes);

} // namespace meep_b200
#endif
