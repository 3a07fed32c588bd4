// hostmem.cpp — see hostmem.hpp.
#include "hostmem.hpp"

#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>
#include <vector>

namespace meep_b200 {

static size_t page_size() {
  static const size_t ps = (size_t)sysconf(_SC_PAGESIZE);
  return ps;
}

bool page_interior(const void *p, size_t bytes, char **lo, char **hi) {
  const size_t ps = page_size();
  const uintptr_t a = ((uintptr_t)p + ps - 1) / ps * ps, b = ((uintptr_t)p + bytes) / ps * ps;
  if (b <= a) return false;
  *lo = (char *)a;
  *hi = (char *)b;
  return true;
}

void release_interior(void *p, size_t bytes) {
  char *lo, *hi;
  if (bytes < kLazyMinBytes || !page_interior(p, bytes, &lo, &hi)) return;
  madvise(lo, (size_t)(hi - lo), MADV_DONTNEED); // failure only means the pages stay
}

meep::realnum *new_zeroed_lazily(size_t n) {
  meep::realnum *a = new meep::realnum[n];
  const size_t bytes = n * sizeof(meep::realnum);
  char *lo, *hi;
  if (bytes < kLazyMinBytes || !page_interior(a, bytes, &lo, &hi) ||
      madvise(lo, (size_t)(hi - lo), MADV_DONTNEED) != 0) {
    memset(a, 0, bytes);
    return a;
  }
  memset(a, 0, (size_t)(lo - (char *)a));
  memset(hi, 0, (size_t)((char *)a + bytes - hi));
  return a;
}

// /proc/self/pagemap: one 64-bit entry per virtual page; bit 63 = present, bit 62 = swapped
bool interior_untouched(const void *p, size_t bytes) {
  char *lo, *hi;
  if (bytes < kLazyMinBytes || !page_interior(p, bytes, &lo, &hi)) return false;
  static int fd = open("/proc/self/pagemap", O_RDONLY | O_CLOEXEC);
  if (fd < 0) return false;
  const size_t ps = page_size();
  size_t first = (uintptr_t)lo / ps, count = (size_t)(hi - lo) / ps;
  std::vector<uint64_t> buf(65536);
  while (count) {
    const size_t n = count < buf.size() ? count : buf.size();
    const ssize_t got = pread(fd, buf.data(), n * 8, (off_t)(first * 8));
    if (got != (ssize_t)(n * 8)) return false;
    for (size_t k = 0; k < n; ++k)
      if (buf[k] >> 62) return false;
    first += n;
    count -= n;
  }
  return true;
}

} // namespace meep_b200
