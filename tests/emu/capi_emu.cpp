// capi_emu.cpp — TEST-ONLY emulator of include/meep_b200.h.
//
// This container has no GPU.  To exercise the host-side engine (job construction from
// meep::fields_chunk, connection-table translation, lazy allocation, mirror protocol) in the
// `-m "not gpu"` tests, this file implements the same C ABI with "device memory" = malloc and
// every kernel = a serial loop over (tile, thread) that calls the SAME per-thread bodies
// (meep_b200/csrc/kernels.cuh, fused.cuh, point_ops.h) the CUDA kernels call.  It is built into
// tests/_build/libmeepb200_emu.so by tests/conftest.py, is never loaded by the product
// libraries, by bench.py or by __graft_entry__.py, and is not a fallback: the shipped
// libmeepb200.so has no host execution path and fails without a CUDA device.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <map>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>
#include <vector>

#include "../../meep_b200/csrc/plan_metrics.h"

using namespace mb200;

static thread_local char g_err[512] = "";
static int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

struct mb200_ctx {
  size_t bytes_allocated;
  int64_t launches;
  std::chrono::steady_clock::time_point t0;
  int64_t prof_launches[MB200_NUM_KINDS];
  double prof_bytes[MB200_NUM_KINDS];
  std::vector<std::pair<int, std::chrono::steady_clock::time_point> > marks;
};

struct mb200_plan {
  int kind, dtype, njobs;
  std::vector<char> jobs;
  std::vector<int64_t> prefix;
  double bytes, points;
};

template <typename T> static void run_plan(mb200_plan *p, const void *run) {
  for (int j = 0; j < p->njobs; ++j) {
    const int64_t ntiles = p->prefix[j + 1] - p->prefix[j];
    switch (p->kind) {
      case MB200_K_CURL: {
        const mb200_curl_job_t &J = ((const mb200_curl_job_t *)p->jobs.data())[j];
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            curl_thread<T>(J, t, tid);
        break;
      }
      case MB200_K_EDHB: {
        const mb200_edhb_job_t &J = ((const mb200_edhb_job_t *)p->jobs.data())[j];
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            edhb_thread<T>(J, t, tid);
        break;
      }
      case MB200_K_LORENTZ: {
        const mb200_lorentz_job_t &J = ((const mb200_lorentz_job_t *)p->jobs.data())[j];
        if (lorentz_blocked_ok(J)) {
          const int64_t nblocks = (J.ntot + MB200_ZBLOCK - 1) / MB200_ZBLOCK;
          for (int64_t t = 0; t < nblocks; ++t) {
            if (J.szero[t] && J.pzero[t]) continue;
            bool zero = true;
            for (int64_t idx = t * MB200_ZBLOCK; idx < (t + 1) * MB200_ZBLOCK && idx < J.ntot; ++idx)
              zero = lorentz_blocked_point<T>(J, idx) && zero;
            J.pzero[t] = zero ? 1 : 0;
          }
          break;
        }
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            lorentz_thread<T>(J, t, tid);
        break;
      }
      case MB200_K_BETA: {
        const mb200_beta_job_t &J = ((const mb200_beta_job_t *)p->jobs.data())[j];
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            beta_thread<T>(J, t, tid);
        break;
      }
      case MB200_K_NOISE: {
        const mb200_noise_job_t &J = ((const mb200_noise_job_t *)p->jobs.data())[j];
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            noise_thread<T>(J, t, tid, (const double *)run);
        break;
      }
      case MB200_K_GYRO: {
        const mb200_gyro_job_t &J = ((const mb200_gyro_job_t *)p->jobs.data())[j];
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            gyro_thread<T>(J, t, tid);
        break;
      }
      case MB200_K_AVERAGE: {
        const mb200_average_job_t &J = ((const mb200_average_job_t *)p->jobs.data())[j];
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            average_thread<T>(J, t, tid);
        break;
      }
      case MB200_K_BFAST: {
        const mb200_bfast_job_t &J = ((const mb200_bfast_job_t *)p->jobs.data())[j];
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            bfast_thread<T>(J, t, tid);
        break;
      }
      case MB200_K_CYLINT: {
        const mb200_cylint_job_t &J = ((const mb200_cylint_job_t *)p->jobs.data())[j];
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            cylint_thread<T>(J, t, tid);
        break;
      }
      case MB200_K_CYLR0: {
        const mb200_cylr0_job_t &J = ((const mb200_cylr0_job_t *)p->jobs.data())[j];
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            cylr0_thread<T>(J, t, tid);
        break;
      }
      case MB200_K_STEP3: {
        const mb200_step3_job_t &J = ((const mb200_step3_job_t *)p->jobs.data())[j];
        const bool plain = step3_is_plain(J);
        static const bool split = !(getenv("MEEP_B200_SPLIT_PML") && atoi(getenv("MEEP_B200_SPLIT_PML")) == 0);
        static const bool lean = !getenv("MEEP_B200_PLAIN_LEAN") || atoi(getenv("MEEP_B200_PLAIN_LEAN")) != 0; // (the device default)
        if (plain && lean) {
          // the lean + shell launches of the device path (fused.cuh: launch_step3), thread by thread
          const Step3Shell S = step3_shell(J);
          if (S.lean) {
            for (int64_t t = 0; t < ntiles; ++t)
              for (int tid = 0; tid < kThreads; ++tid)
                step3_lean_thread<T>(J, t, tid);
            for (const mb200_step3_job_t &slab : S.slabs) {
              const int64_t st = step3_tiles(slab);
              for (int64_t t = 0; t < st; ++t)
                for (int tid = 0; tid < kThreads; ++tid)
                  step3_plain_thread<T>(slab, t, tid);
            }
            const int row = J.n[2] + 1, planes = step3_t1(J);
            for (int col : S.cols)
              for (int x0 = S.lo[0]; x0 <= S.hi[0]; x0 += planes)
                step3_plain_column<T>(J, x0, x0 + planes < S.hi[0] + 1 ? x0 + planes : S.hi[0] + 1, col / row, col % row);
            break;
          }
        }
        for (int64_t t = 0; t < ntiles; ++t) {
          if (!plain && split) {
            for (int c = 0; c < 3; ++c)
              for (int tid = 0; tid < kThreads; ++tid)
                step3c_thread<T>(J, c, t, tid);
            continue;
          }
          for (int tid = 0; tid < kThreads; ++tid)
            if (plain) step3_plain_thread<T>(J, t, tid);
            else step3_thread<T>(J, t, tid);
        }
        break;
      }
      case MB200_K_FMP: {
        const mb200_fmp_job_t &J = ((const mb200_fmp_job_t *)p->jobs.data())[j];
        for (int64_t i = 0; i < J.ntot; ++i)
          fmp_point<T>(J, i);
        break;
      }
      case MB200_K_SOURCE: {
        const mb200_src_job_t &J = ((const mb200_src_job_t *)p->jobs.data())[j];
        for (int64_t i = 0; i < J.npts; ++i)
          source_point<T>(J, i, (const double *)run);
        break;
      }
      case MB200_K_HALO: {
        const mb200_halo_job_t &J = ((const mb200_halo_job_t *)p->jobs.data())[j];
        for (int64_t t = 0; t < ntiles; ++t)
          for (int tid = 0; tid < kThreads; ++tid)
            halo_thread<T>(J, t, tid);
        break;
      }
      case MB200_K_ZERO: {
        const mb200_zero_job_t &J = ((const mb200_zero_job_t *)p->jobs.data())[j];
        for (int64_t n = 0; n < J.n; ++n)
          *(T *)(uintptr_t)J.ptrs[n] = T(0);
        break;
      }
      case MB200_K_DFT: {
        const mb200_dft_job_t &J = ((const mb200_dft_job_t *)p->jobs.data())[j];
        const T *ph = (const T *)run + 2 * (int64_t)J.phase_slot;
        int64_t pt = 0;
        for (int i1 = 0; i1 < J.box.n[0]; ++i1)
          for (int i2 = 0; i2 < J.box.n[1]; ++i2)
            for (int i3 = 0; i3 < J.box.n[2]; ++i3, ++pt) {
              T fr, fi;
              dft_field_value<T>(J, i1, i2, i3, fr, fi);
              T *d = (T *)J.dft + 2 * pt * J.nomega;
              for (int w = 0; w < J.nomega; ++w)
                dft_accumulate<T>(d + 2 * w, J.f_im != nullptr, ph[2 * w], ph[2 * w + 1], fr, fi);
            }
        break;
      }
      case MB200_K_FLUX: {
        const mb200_flux_job_t &J = ((const mb200_flux_job_t *)p->jobs.data())[j];
        const T *e = (const T *)J.e, *h = (const T *)J.h;
        for (int64_t k = 0; k < J.npts; ++k)
          for (int w = 0; w < J.nomega; ++w) {
            const int64_t o = 2 * (k * J.nomega + w);
            J.out[w] += (double)(e[o] * h[o] + e[o + 1] * h[o + 1]);
          }
        break;
      }
    }
  }
}

extern "C" {

int mb200_abi_version(void) { return MB200_ABI_VERSION; }
const char *mb200_last_error(void) { return g_err; }
int mb200_device_count(void) { return 1; } // one emulated device
int mb200_is_emulator(void) { return 1; }  // symbol that exists ONLY in the emulator

int mb200_init(int device, mb200_ctx **out) {
  if (device != 0) return fail("emu: only device 0 exists");
  mb200_ctx *c = new mb200_ctx();
  c->bytes_allocated = 0;
  c->launches = 0;
  memset(c->prof_launches, 0, sizeof(c->prof_launches));
  memset(c->prof_bytes, 0, sizeof(c->prof_bytes));
  *out = c;
  return 0;
}
void mb200_destroy(mb200_ctx *c) { delete c; }
int mb200_sync(mb200_ctx *) { return 0; }

// "Device" memory is page-granular host memory so that an allocation can later be turned into a
// shared mapping in place (the stand-in for CUDA IPC, see mb200_ipc_export below).
struct EmuAlloc {
  size_t size;
  int fd; // memfd once exported, else -1
};
static std::map<void *, EmuAlloc> g_allocs;
static std::map<void *, size_t> g_imports;

int mb200_malloc(mb200_ctx *c, size_t bytes, void **out) {
  const size_t sz = ((bytes ? bytes : 8) + 4095) & ~(size_t)4095;
  if (posix_memalign(out, 4096, sz) != 0 || !*out) return fail("emu: out of memory");
  memset(*out, 0xA5, sz); // poison: device memory is uninitialised
  g_allocs[*out] = EmuAlloc{sz, -1};
  c->bytes_allocated += bytes;
  return 0;
}
int mb200_free(mb200_ctx *, void *p) {
  auto it = g_allocs.find(p);
  if (it != g_allocs.end()) {
    if (it->second.fd >= 0) {
      // give the pages back to private anonymous memory before the allocator reuses them
      mmap(p, it->second.size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED, -1, 0);
      close(it->second.fd);
    }
    g_allocs.erase(it);
  }
  free(p);
  return 0;
}
int mb200_memset(mb200_ctx *, void *p, int value, size_t bytes) {
  memset(p, value, bytes);
  return 0;
}
int mb200_h2d(mb200_ctx *, void *dst, const void *src, size_t bytes) {
  memcpy(dst, src, bytes);
  return 0;
}
int mb200_d2h_box(mb200_ctx *, void *host, const void *dev, size_t elem_size, int64_t rows, int64_t row_elems,
                  int64_t lo0, int64_t lo1, int64_t lo2, int64_t cnt0, int64_t cnt1, int64_t cnt2) {
  for (int64_t a = lo0; a < lo0 + cnt0; ++a)
    for (int64_t b = lo1; b < lo1 + cnt1; ++b) {
      const size_t off = (size_t)((a * rows + b) * row_elems + lo2) * elem_size;
      memcpy((char *)host + off, (const char *)dev + off, (size_t)cnt2 * elem_size);
    }
  return 0;
}
int mb200_d2h(mb200_ctx *, void *dst, const void *src, size_t bytes) {
  memcpy(dst, src, bytes);
  return 0;
}
int mb200_d2h_async(mb200_ctx *, void *dst, const void *src, size_t bytes) {
  memcpy(dst, src, bytes);
  return 0;
}
int mb200_d2d(mb200_ctx *, void *dst, const void *src, size_t bytes) {
  memmove(dst, src, bytes);
  return 0;
}
int mb200_host_alloc(size_t bytes, void **out) {
  *out = malloc(bytes ? bytes : 8);
  return *out ? 0 : fail("emu: out of memory");
}
int mb200_host_free(void *p) {
  free(p);
  return 0;
}
size_t mb200_bytes_allocated(mb200_ctx *c) { return c->bytes_allocated; }

int mb200_plan_create(mb200_ctx *, int kind, int dtype, const void *jobs, int njobs,
                      mb200_plan **out) {
  *out = nullptr;
  const size_t js = job_size_of(kind);
  if (!js) return fail("mb200_plan_create: unknown kind %d", kind);
  if (dtype != MB200_F64 && dtype != MB200_F32)
    return fail("mb200_plan_create: unknown dtype %d", dtype);
  if (njobs < 0 || (njobs > 0 && !jobs)) return fail("mb200_plan_create: bad job table");
  mb200_plan *p = new mb200_plan();
  p->kind = kind;
  p->dtype = dtype;
  p->njobs = njobs;
  p->jobs.assign((const char *)jobs, (const char *)jobs + js * njobs);
  p->prefix.assign(njobs + 1, 0);
  p->bytes = p->points = 0;
  for (int j = 0; j < njobs; ++j) {
    int64_t t;
    double b, pts;
    job_metrics(kind, dtype, jobs, j, &t, &b, &pts);
    p->prefix[j + 1] = p->prefix[j] + t;
    p->bytes += b;
    p->points += pts;
  }
  *out = p;
  return 0;
}
void mb200_plan_destroy(mb200_ctx *, mb200_plan *p) { delete p; }
double mb200_plan_bytes(const mb200_plan *p) { return p ? p->bytes : 0; }
double mb200_plan_points(const mb200_plan *p) { return p ? p->points : 0; }

int mb200_plan_run(mb200_ctx *c, mb200_plan *p, const void *run_data, size_t run_bytes) {
  if (!p) return fail("mb200_plan_run: plan == NULL");
  if (p->prefix.back() == 0) return 0;
  const bool needs_run = p->kind == MB200_K_SOURCE || p->kind == MB200_K_DFT;
  if (needs_run && (!run_data || !run_bytes))
    return fail("mb200_plan_run: kind %d needs run_data", p->kind);
  if (p->dtype == MB200_F64) run_plan<double>(p, run_data);
  else run_plan<float>(p, run_data);
  c->launches += 1;
  c->prof_launches[p->kind] += 1;
  c->prof_bytes[p->kind] += p->bytes;
  return 0;
}

static int one_shot(mb200_ctx *c, int kind, int dtype, const void *jobs, int njobs,
                    const void *run, size_t run_bytes) {
  mb200_plan *p = nullptr;
  if (mb200_plan_create(c, kind, dtype, jobs, njobs, &p)) return 1;
  int rc = mb200_plan_run(c, p, run, run_bytes);
  mb200_plan_destroy(c, p);
  return rc;
}
int mb200_step_curl(mb200_ctx *c, int dtype, const mb200_curl_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_CURL, dtype, jobs, njobs, nullptr, 0);
}
int mb200_step_update_EDHB(mb200_ctx *c, int dtype, const mb200_edhb_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_EDHB, dtype, jobs, njobs, nullptr, 0);
}
int mb200_lorentzian_update_P(mb200_ctx *c, int dtype, const mb200_lorentz_job_t *jobs,
                              int njobs) {
  return one_shot(c, MB200_K_LORENTZ, dtype, jobs, njobs, nullptr, 0);
}
int mb200_subtract_P(mb200_ctx *c, int dtype, const mb200_fmp_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_FMP, dtype, jobs, njobs, nullptr, 0);
}
int mb200_step_source(mb200_ctx *c, int dtype, const mb200_src_job_t *jobs, int njobs,
                      const double *scalars, int nslots) {
  return one_shot(c, MB200_K_SOURCE, dtype, jobs, njobs, scalars, 16 * (size_t)nslots);
}
int mb200_step_boundaries(mb200_ctx *c, int dtype, const mb200_halo_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_HALO, dtype, jobs, njobs, nullptr, 0);
}
int mb200_zero_metal(mb200_ctx *c, int dtype, const mb200_zero_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_ZERO, dtype, jobs, njobs, nullptr, 0);
}
int mb200_update_dft(mb200_ctx *c, int dtype, const mb200_dft_job_t *jobs, int njobs,
                     const void *phases, int nphases) {
  const size_t R = dtype == MB200_F64 ? 8 : 4;
  return one_shot(c, MB200_K_DFT, dtype, jobs, njobs, phases, 2 * R * (size_t)nphases);
}
int mb200_dft_flux(mb200_ctx *c, int dtype, const mb200_flux_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_FLUX, dtype, jobs, njobs, nullptr, 0);
}
int mb200_step3(mb200_ctx *c, int dtype, const mb200_step3_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_STEP3, dtype, jobs, njobs, nullptr, 0);
}
int mb200_step_beta(mb200_ctx *c, int dtype, const mb200_beta_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_BETA, dtype, jobs, njobs, nullptr, 0);
}
int mb200_add_noise(mb200_ctx *c, int dtype, const mb200_noise_job_t *jobs, int njobs, const double *noise,
                    int64_t nnoise) {
  return one_shot(c, MB200_K_NOISE, dtype, jobs, njobs, noise, sizeof(double) * (size_t)nnoise);
}
int mb200_gyrotropic_update_P(mb200_ctx *c, int dtype, const mb200_gyro_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_GYRO, dtype, jobs, njobs, nullptr, 0);
}
int mb200_average_with_backup(mb200_ctx *c, int dtype, const mb200_average_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_AVERAGE, dtype, jobs, njobs, nullptr, 0);
}
int mb200_step_bfast(mb200_ctx *c, int dtype, const mb200_bfast_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_BFAST, dtype, jobs, njobs, nullptr, 0);
}
int mb200_cyl_rderiv_int(mb200_ctx *c, int dtype, const mb200_cylint_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_CYLINT, dtype, jobs, njobs, nullptr, 0);
}
int mb200_cyl_origin(mb200_ctx *c, int dtype, const mb200_cylr0_job_t *jobs, int njobs) {
  return one_shot(c, MB200_K_CYLR0, dtype, jobs, njobs, nullptr, 0);
}

// the emulator has no device interconnect: the host engine moves "device" comm blocks (plain
// host memory here) through its socket runtime instead (meep_b200/host/step.cpp)
int mb200_comm_unique_id(void *id128) {
  memset(id128, 0, 128);
  return 0;
}
struct mb200_comm {
  int rank, nranks;
};
int mb200_comm_create(mb200_ctx *, int rank, int nranks, const void *, mb200_comm **out) {
  *out = new mb200_comm{rank, nranks};
  return 0;
}
void mb200_comm_destroy(mb200_comm *m) { delete m; }
int mb200_comm_exchange(mb200_ctx *, mb200_comm *, int, const mb200_xfer_t *, int, const mb200_xfer_t *,
                        int) {
  return fail("emu: mb200_comm_exchange is not available (the host engine uses its socket runtime)");
}

int mb200_block_zero_flags(mb200_ctx *c, int dtype, const void *arr, int64_t n, uint8_t *flags) {
  for (int64_t b = 0; b * MB200_ZBLOCK < n; ++b) {
    bool zero = true;
    for (int64_t i = b * MB200_ZBLOCK; i < (b + 1) * MB200_ZBLOCK && i < n; ++i)
      if ((dtype == MB200_F64 ? ((const double *)arr)[i] : (double)((const float *)arr)[i]) != 0) zero = false;
    flags[b] = zero ? 1 : 0;
  }
  c->launches += 1;
  return 0;
}

// Stand-in for CUDA IPC between emulated devices (= processes): the allocation is re-mapped in
// place as a shared memfd mapping; the handle names it as /proc/<pid>/fd/<fd>.
struct EmuHandle {
  int32_t pid, fd;
  uint64_t size;
};
int mb200_ipc_export(mb200_ctx *, void *devptr, void *handle64) {
  auto it = g_allocs.find(devptr);
  if (it == g_allocs.end()) return fail("emu: ipc_export of an unknown allocation");
  EmuAlloc &a = it->second;
  if (a.fd < 0) {
    int fd = memfd_create("mb200_emu_arena", 0);
    if (fd < 0 || ftruncate(fd, (off_t)a.size) != 0) return fail("emu: memfd_create failed");
    if (pwrite(fd, devptr, a.size, 0) != (ssize_t)a.size) return fail("emu: pwrite failed");
    if (mmap(devptr, a.size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0) != devptr)
      return fail("emu: could not re-map the allocation as shared memory");
    a.fd = fd;
  }
  EmuHandle h = {(int32_t)getpid(), (int32_t)a.fd, (uint64_t)a.size};
  memset(handle64, 0, 64);
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
int mb200_ipc_import(mb200_ctx *, const void *handle64, void **out) {
  EmuHandle h;
  memcpy(&h, handle64, sizeof(h));
  char path[64];
  snprintf(path, sizeof path, "/proc/%d/fd/%d", (int)h.pid, (int)h.fd);
  int fd = open(path, O_RDWR);
  if (fd < 0) return fail("emu: cannot open the peer's arena");
  void *p = mmap(nullptr, h.size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) return fail("emu: cannot map the peer's arena");
  g_imports[p] = h.size;
  *out = p;
  return 0;
}
int mb200_ipc_close(mb200_ctx *, void *imported) {
  auto it = g_imports.find(imported);
  if (it == g_imports.end()) return fail("emu: ipc_close of an unknown mapping");
  munmap(imported, it->second);
  g_imports.erase(it);
  return 0;
}
int mb200_flag_signal(mb200_ctx *, uint64_t *flag, uint64_t value);
int mb200_flag_wait(mb200_ctx *, const uint64_t *flag, uint64_t value);
int mb200_flag_signal_many(mb200_ctx *c, uint64_t *const *flags, const uint64_t *values, int n) {
  for (int k = 0; k < n; ++k)
    if (mb200_flag_signal(c, flags[k], values[k])) return 1;
  return 0;
}
int mb200_flag_wait_many(mb200_ctx *c, const uint64_t *const *flags, const uint64_t *values, int n) {
  for (int k = 0; k < n; ++k)
    if (mb200_flag_wait(c, flags[k], values[k])) return 1;
  return 0;
}
int mb200_flag_signal(mb200_ctx *, uint64_t *flag, uint64_t value) {
  __atomic_store_n(flag, value, __ATOMIC_RELEASE);
  return 0;
}
int mb200_flag_wait(mb200_ctx *, const uint64_t *flag, uint64_t value) {
  // the emulated stream is the calling thread: wait here (bounded, like the device-side spin)
  const char *e = getenv("MEEP_B200_PEER_TIMEOUT_S");
  const double limit = e && atof(e) > 0 ? atof(e) : 60.0;
  const auto t0 = std::chrono::steady_clock::now();
  while (__atomic_load_n(flag, __ATOMIC_ACQUIRE) < value) {
    sched_yield();
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit)
      return fail("emu: a wait for a neighbouring device timed out (peer exchange)");
  }
  return 0;
}

int mb200_check_finite(mb200_ctx *c, int dtype, const uint64_t *ptrs, int64_t n, int32_t *flag) {
  for (int64_t i = 0; i < n; ++i) {
    const double v = dtype == MB200_F64 ? *(const double *)(uintptr_t)ptrs[i]
                                        : (double)*(const float *)(uintptr_t)ptrs[i];
    if (!std::isfinite(v)) *flag = 1;
  }
  c->launches += 1;
  return 0;
}

int mb200_timer_start(mb200_ctx *c) {
  c->t0 = std::chrono::steady_clock::now();
  return 0;
}
int mb200_timer_stop(mb200_ctx *c, double *ms) {
  *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c->t0).count();
  return 0;
}
int mb200_profile_enable(mb200_ctx *, int) { return 0; }
int mb200_profile_reset(mb200_ctx *c) {
  memset(c->prof_launches, 0, sizeof(c->prof_launches));
  memset(c->prof_bytes, 0, sizeof(c->prof_bytes));
  return 0;
}
int mb200_profile_get(mb200_ctx *c, int kind, int64_t *launches, double *ms, double *bytes) {
  if (kind < 0 || kind >= MB200_NUM_KINDS) return fail("mb200_profile_get: bad kind");
  if (launches) *launches = c->prof_launches[kind];
  if (ms) *ms = 0;
  if (bytes) *bytes = c->prof_bytes[kind];
  return 0;
}
int mb200_mark(mb200_ctx *c, int tag) {
  c->marks.push_back(std::make_pair(tag, std::chrono::steady_clock::now()));
  return 0;
}
int mb200_marks_collect(mb200_ctx *c, int *tags, double *ms, int cap, int *n) {
  *n = 0;
  for (size_t k = 0; k + 1 < c->marks.size() && *n < cap; ++k) {
    tags[*n] = c->marks[k].first;
    ms[*n] = std::chrono::duration<double, std::milli>(c->marks[k + 1].second - c->marks[k].second).count();
    *n += 1;
  }
  c->marks.clear();
  return 0;
}
int64_t mb200_launch_count(mb200_ctx *c) { return c->launches; }

} // extern "C"
