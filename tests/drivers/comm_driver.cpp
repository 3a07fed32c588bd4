// comm_driver.cpp — exercises the process-runtime entry points of the reference's src/mympi.cpp
// (meep::sum_to_all, broadcast, and_to_all, partial_sum_to_all, ...) as served by the MPI-free
// runtime of the drop-in (meep_b200/host/mympi_b200.cpp).  Launched as WORLD_SIZE cooperating
// processes by tests/test_host_emu.py; every rank checks the results it must see and exits
// non-zero on a mismatch.  With the plain reference build (one process) the same program checks
// the single-process identities.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <complex>
#include <vector>

#include "meep.hpp"

using namespace meep;

static int g_fail = 0;
#define CHECK(cond)                                                                                \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      fprintf(stderr, "rank %d: check failed at line %d: %s\n", my_rank(), __LINE__, #cond);      \
      g_fail = 1;                                                                                  \
    }                                                                                              \
  } while (0)

int main(int argc, char **argv) {
  initialize mpi(argc, argv);
  const int r = my_rank(), n = count_processors();
  CHECK(r >= 0 && r < n);
  CHECK(am_master() == (r == 0));
  // scalar reductions
  CHECK(sum_to_all(r + 1) == n * (n + 1) / 2);
  CHECK(sum_to_all((size_t)(r + 1)) == (size_t)(n * (n + 1) / 2));
  CHECK(fabs(sum_to_all(0.5 * (r + 1)) - 0.25 * n * (n + 1)) < 1e-12);
  CHECK(max_to_all(r) == n - 1);
  CHECK(min_to_all(r + 3) == 3);
  CHECK(fabs(max_to_all(1.5 * r) - 1.5 * (n - 1)) < 1e-12);
  CHECK(or_to_all(r == n - 1) == true);
  CHECK(or_to_all(false) == false);
  CHECK(and_to_all(true) == true);
  CHECK(and_to_all(r != n - 1) == (n == 1 ? true : false) || n == 1);
  const std::complex<double> z = sum_to_all(std::complex<double>(r, -r));
  CHECK(fabs(z.real() - 0.5 * n * (n - 1)) < 1e-12 && fabs(z.imag() + 0.5 * n * (n - 1)) < 1e-12);
  // exclusive/inclusive prefix sums (src/mympi.cpp: partial_sum_to_all = MPI_Scan, inclusive)
  CHECK(partial_sum_to_all(r + 1) == (r + 1) * (r + 2) / 2);
  CHECK(partial_sum_to_all((size_t)2) == (size_t)(2 * (r + 1)));
  // vector reductions
  std::vector<double> in(5), out(5, -1.0);
  for (int k = 0; k < 5; ++k) in[k] = r + 0.1 * k;
  sum_to_all(in.data(), out.data(), 5);
  for (int k = 0; k < 5; ++k) CHECK(fabs(out[k] - (0.5 * n * (n - 1) + 0.1 * k * n)) < 1e-12);
  std::vector<int> bi(4), bo(4, -1);
  for (int k = 0; k < 4; ++k) bi[k] = (r % (k + 1)) == 0;
  and_to_all(bi.data(), bo.data(), 4);
  for (int k = 0; k < 4; ++k) {
    int want = 1;
    for (int q = 0; q < n; ++q) want = want && ((q % (k + 1)) == 0);
    CHECK(bo[k] == want);
  }
  std::vector<double> m(3, 0.0), mo(3, -1.0);
  m[0] = r;
  sum_to_master(m.data(), mo.data(), 3);
  if (r == 0) CHECK(fabs(mo[0] - 0.5 * n * (n - 1)) < 1e-12);
  // broadcasts from every rank
  for (int root = 0; root < n; ++root) {
    double d[3] = {r == root ? 3.25 + root : -1.0, r == root ? 1.0 : -1.0, 0};
    broadcast(root, d, 3);
    CHECK(d[0] == 3.25 + root && d[1] == 1.0);
    CHECK(broadcast(root, r == root ? 41 + root : -7) == 41 + root);
    CHECK(broadcast(root, r == root) == true);
    char buf[8] = "xxxxxxx";
    if (r == root) snprintf(buf, sizeof buf, "abc%d", root % 10);
    broadcast(root, buf, 8);
    CHECK(buf[0] == 'a' && buf[3] == '0' + root % 10);
  }
  all_wait();
  if (r == 0 && !g_fail) printf("comm_driver: %d ranks ok\n", n);
  return g_fail;
}
