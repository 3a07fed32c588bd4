// capi.cu — the C ABI of include/meep_b200.h on CUDA (sm_100a).
// Thin by design: context + device memory + "plans" (job table + tile prefix in HBM, one grid
// per run).  All numerical work is in kernels.cuh / point_ops.h.  No host execution path.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "kernels.cuh"
#include "fused.cuh"
#include "plan_metrics.h"

using namespace mb200;

static thread_local char g_err[512] = "";

static int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);     \
  } while (0)

struct ProfRec {
  int kind;
  cudaEvent_t a, b;
  double bytes;
  // kernels that skip known-zero blocks: bytes = base + per_block * (blocks counted by the launch)
  int slot = -1;
  double per_block = 0;
};
constexpr int kWorkSlots = 4096; // pinned host ring the per-launch work counters are copied to

struct mb200_ctx {
  int device;
  cudaStream_t stream;
  size_t bytes_allocated;
  int64_t launches;
  cudaEvent_t t0, t1;
  bool profiling;
  std::vector<ProfRec> recs;
  int64_t prof_launches[MB200_NUM_KINDS];
  double prof_ms[MB200_NUM_KINDS], prof_bytes[MB200_NUM_KINDS];
  void *run_buf;
  size_t run_cap;
  int *d_err; // latched device-side error word (flag-wait time-out)
  std::vector<std::pair<int, cudaEvent_t> > marks; // phase marks (mb200_mark)
  std::vector<cudaEvent_t> event_pool;
  unsigned long long *d_work;  // [0] Lorentz blocks updated, [1] polarisation blocks read by f_minus_p
  unsigned long long *h_work;  // pinned: 2 words per slot
  int next_slot;
};

struct mb200_plan {
  int kind, dtype, njobs;
  void *d_jobs;
  int64_t *d_prefix;
  int64_t tiles;
  double bytes, points;
  size_t job_size;
  bool all_plain; // STEP3: every job qualifies for the fast-path kernel
  Step3LeanPlan *lean;            // STEP3, all_plain, MEEP_B200_PLAIN_LEAN: the lean + shell launches of every job
  std::vector<void *> lean_allocs; // device tables of `lean`
  int *d_group;   // EDHB: first job of the component triple a job belongs to (or the job itself), see plan_create
  std::vector<char> h_jobs;       // STEP3: host copy (jobs are passed by value in param space)
  std::vector<int64_t> h_prefix;
};

// MEEP_B200_PARAMJOBS=1 passes fused-kernel job descriptors in kernel-parameter (constant) space
// instead of staging them in shared memory (experiment; see profiles/ for the comparison)
static const bool g_param_jobs = getenv("MEEP_B200_PARAMJOBS") && atoi(getenv("MEEP_B200_PARAMJOBS")) != 0;

// MEEP_B200_SPLIT_PML=0 runs the PML chunks with the three-components-per-thread general kernel
// instead of the one-component-per-thread form; 3..6 = CTAs per SM of the latter (default 4: 64
// registers).  Measured at 512^3 (profiles/r1q_*): 0 -> 1.52 ms, 3 -> 1.25, 4 -> 1.04, 5 -> 1.18,
// 6 -> 1.36 ms per step (16 planes per CTA).
static const int g_split_general = getenv("MEEP_B200_SPLIT_PML") ? atoi(getenv("MEEP_B200_SPLIT_PML")) : 4;

// Fast-path kernel form.  Default: per job, the lean march (fused.cuh: step3_lean_kernel) over the full
// box + the masked march over the shell (two x-slab jobs, a list of (y, z) columns) where that pays —
// single precision, and the half-step without the E/H epilogue in double — and one table-driven launch
// of the masked march otherwise.  MEEP_B200_PLAIN_LEAN=0: the masked march for everything (the round-1
// form).  Measured (B200; bench/micro/pml_shapes.cu "lean", profiles/r2ab_*, r2ad_*): lean + shell against
// masked over the 492^3 interior: -9.8 % (double, B half), +0.7 % (double, D-E half), -17 % / -6 % (single);
// 1024^3 step 37.70 -> 36.8 ms with every job lean.  Forms that lost, per-launch ncu times B half / D-E
// half at 512^3 double (profiles/README.md, r2d-r2g): masked march 1597 / 2350 us; ONE kernel holding both
// marches 1639 / 2790 us; the boundary TILES re-walked by the masked march 3.85 -> 4.15 ms per step; more
// CTAs per SM (5: 1894 us, 6: 2120 us for the B half); operands staged through shared memory with 8-byte
// cp.async 2518 / 2896 us.
// MEEP_B200_EDHB_INTERLEAVE=0: the three component jobs of an off-diagonal E update one after the other
static const bool g_edhb_interleave = !getenv("MEEP_B200_EDHB_INTERLEAVE") || atoi(getenv("MEEP_B200_EDHB_INTERLEAVE")) != 0;
static const bool g_plain_lean = !getenv("MEEP_B200_PLAIN_LEAN") || atoi(getenv("MEEP_B200_PLAIN_LEAN")) != 0;

// *kernels = number of kernels the plan run launched (what mb200_launch_count reports)
template <typename T>
static cudaError_t launch_plan(mb200_ctx *c, mb200_plan *p, const void *d_run, int *kernels) {
  const dim3 grid((unsigned)p->tiles), block(kThreads);
  cudaStream_t s = c->stream;
  *kernels = 1;
  switch (p->kind) {
    case MB200_K_CURL:
      curl_kernel<T><<<grid, block, 0, s>>>((const mb200_curl_job_t *)p->d_jobs, p->d_prefix,
                                            p->njobs);
      break;
    case MB200_K_EDHB:
      edhb_kernel<T><<<grid, block, 0, s>>>((const mb200_edhb_job_t *)p->d_jobs, p->d_prefix,
                                            p->njobs, p->d_group);
      break;
    case MB200_K_LORENTZ:
      if (p->all_plain) { // (for this kind: every job uses the zero-block variant)
        lorentz_blocked_kernel<T><<<grid, block, 0, s>>>((const mb200_lorentz_job_t *)p->d_jobs,
                                                         p->d_prefix, p->njobs, c->d_work);
        break;
      }
      lorentz_kernel<T><<<grid, block, 0, s>>>((const mb200_lorentz_job_t *)p->d_jobs, p->d_prefix,
                                               p->njobs);
      break;
    case MB200_K_FMP:
      fmp_kernel<T><<<grid, block, 0, s>>>((const mb200_fmp_job_t *)p->d_jobs, p->d_prefix,
                                           p->njobs, c->d_work);
      break;
    case MB200_K_SOURCE:
      source_kernel<T><<<grid, block, 0, s>>>((const mb200_src_job_t *)p->d_jobs, p->d_prefix,
                                              p->njobs, (const double *)d_run);
      break;
    case MB200_K_HALO:
      halo_kernel<T><<<grid, block, 0, s>>>((const mb200_halo_job_t *)p->d_jobs, p->d_prefix,
                                            p->njobs);
      break;
    case MB200_K_ZERO:
      zero_kernel<T><<<grid, block, 0, s>>>((const mb200_zero_job_t *)p->d_jobs, p->d_prefix,
                                            p->njobs);
      break;
    case MB200_K_DFT:
      dft_kernel<T><<<grid, block, 0, s>>>((const mb200_dft_job_t *)p->d_jobs, p->d_prefix,
                                           p->njobs, (const T *)d_run);
      break;
    case MB200_K_FLUX: {
      // (plan_create checked that the jobs share nomega and out; plan_run sized c->run_buf)
      const mb200_flux_job_t &J0 = *(const mb200_flux_job_t *)p->h_jobs.data();
      flux_partial_kernel<T><<<grid, block, 0, s>>>((const mb200_flux_job_t *)p->d_jobs, p->d_prefix,
                                                    p->njobs, (double *)d_run, J0.nomega);
      flux_final_kernel<<<dim3((unsigned)ceil_div(J0.nomega, kThreads / 32)), block, 0, s>>>(
          (const double *)d_run, p->tiles, J0.nomega, J0.nomega, J0.out);
      *kernels = 2;
      break;
    }
    case MB200_K_BETA:
      beta_kernel<T><<<grid, block, 0, s>>>((const mb200_beta_job_t *)p->d_jobs, p->d_prefix,
                                            p->njobs);
      break;
    case MB200_K_NOISE:
      noise_kernel<T><<<grid, block, 0, s>>>((const mb200_noise_job_t *)p->d_jobs, p->d_prefix, p->njobs,
                                             (const double *)d_run);
      break;
    case MB200_K_GYRO:
      gyro_kernel<T><<<grid, block, 0, s>>>((const mb200_gyro_job_t *)p->d_jobs, p->d_prefix, p->njobs);
      break;
    case MB200_K_AVERAGE:
      average_kernel<T><<<grid, block, 0, s>>>((const mb200_average_job_t *)p->d_jobs, p->d_prefix,
                                               p->njobs);
      break;
    case MB200_K_BFAST:
      bfast_kernel<T><<<grid, block, 0, s>>>((const mb200_bfast_job_t *)p->d_jobs, p->d_prefix,
                                             p->njobs);
      break;
    case MB200_K_CYLINT:
      cylint_kernel<T><<<grid, block, 0, s>>>((const mb200_cylint_job_t *)p->d_jobs, p->d_prefix,
                                              p->njobs);
      break;
    case MB200_K_CYLR0:
      cylr0_kernel<T><<<grid, block, 0, s>>>((const mb200_cylr0_job_t *)p->d_jobs, p->d_prefix,
                                             p->njobs);
      break;
    case MB200_K_STEP3:
      if (g_param_jobs) {
        launch_step3_params<T>((const mb200_step3_job_t *)p->h_jobs.data(), p->h_prefix.data(),
                               p->njobs, p->all_plain, s);
        break;
      }
      *kernels = launch_step3<T>((const mb200_step3_job_t *)p->d_jobs, p->d_prefix, p->njobs, p->tiles,
                      p->all_plain, g_split_general, s, (const mb200_step3_job_t *)p->h_jobs.data(),
                      p->h_prefix.data(), p->lean);
      break;
  }
  return cudaGetLastError();
}

extern "C" {

int mb200_abi_version(void) { return MB200_ABI_VERSION; }
const char *mb200_last_error(void) { return g_err; }

int mb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int mb200_init(int device, mb200_ctx **out) {
  if (!out) return fail("mb200_init: out == NULL");
  *out = nullptr;
  int n = 0;
  CUDA_TRY(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n)
    return fail("mb200_init: device %d not available (%d CUDA devices visible)", device, n);
  CUDA_TRY(cudaSetDevice(device));
  mb200_ctx *c = new mb200_ctx();
  c->device = device;
  c->bytes_allocated = 0;
  c->launches = 0;
  c->profiling = false;
  c->run_buf = nullptr;
  c->run_cap = 0;
  memset(c->prof_launches, 0, sizeof(c->prof_launches));
  memset(c->prof_ms, 0, sizeof(c->prof_ms));
  memset(c->prof_bytes, 0, sizeof(c->prof_bytes));
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  if (const char *e = getenv("MEEP_B200_PML_PAIR")) {
    const int v = atoi(e);
    CUDA_TRY(cudaMemcpyToSymbol(mb200::g_pml_pair, &v, sizeof(int)));
  }
  if (const char *e = getenv("MEEP_B200_FMP_SIMPLE")) {
    const int v = atoi(e);
    CUDA_TRY(cudaMemcpyToSymbol(mb200::g_fmp_simple, &v, sizeof(int)));
  }
  CUDA_TRY(cudaEventCreate(&c->t0));
  CUDA_TRY(cudaEventCreate(&c->t1));
  CUDA_TRY(cudaMalloc((void **)&c->d_err, sizeof(int)));
  CUDA_TRY(cudaMemset(c->d_err, 0, sizeof(int)));
  CUDA_TRY(cudaMalloc((void **)&c->d_work, 2 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemset(c->d_work, 0, 2 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMallocHost((void **)&c->h_work, 2 * sizeof(unsigned long long) * kWorkSlots));
  c->next_slot = 0;
  *out = c;
  return 0;
}

void mb200_destroy(mb200_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto &r : c->recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  if (c->run_buf) cudaFree(c->run_buf);
  if (c->d_err) cudaFree(c->d_err);
  for (auto &m : c->marks) cudaEventDestroy(m.second);
  for (cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
  if (c->d_work) cudaFree(c->d_work);
  if (c->h_work) cudaFreeHost(c->h_work);
  cudaEventDestroy(c->t0);
  cudaEventDestroy(c->t1);
  cudaStreamDestroy(c->stream);
  delete c;
}

int mb200_sync(mb200_ctx *c) {
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  int err = 0;
  CUDA_TRY(cudaMemcpy(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) return fail("a device-side wait for a neighbouring GPU timed out (peer exchange)");
  return 0;
}

int mb200_malloc(mb200_ctx *c, size_t bytes, void **out) {
  CUDA_TRY(cudaSetDevice(c->device));
  *out = nullptr;
  if (bytes == 0) bytes = 8;
  CUDA_TRY(cudaMalloc(out, bytes));
  c->bytes_allocated += bytes;
  return 0;
}

int mb200_free(mb200_ctx *c, void *p) {
  if (!p) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaFree(p));
  return 0;
}

int mb200_memset(mb200_ctx *c, void *p, int value, size_t bytes) {
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemsetAsync(p, value, bytes, c->stream));
  return 0;
}

int mb200_h2d(mb200_ctx *c, void *dst, const void *src, size_t bytes) {
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return 0;
}

int mb200_d2h(mb200_ctx *c, void *dst, const void *src, size_t bytes) {
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return 0;
}

int mb200_d2h_async(mb200_ctx *c, void *dst, const void *src, size_t bytes) {
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}

int mb200_d2h_box(mb200_ctx *c, void *host, const void *dev, size_t elem_size, int64_t rows, int64_t row_elems,
                  int64_t lo0, int64_t lo1, int64_t lo2, int64_t cnt0, int64_t cnt1, int64_t cnt2) {
  if (cnt0 <= 0 || cnt1 <= 0 || cnt2 <= 0) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof(p));
  // (cudaMemcpy3D: x = fastest index in BYTES, y = rows, z = planes)
  p.srcPtr = make_cudaPitchedPtr(const_cast<void *>(dev), (size_t)row_elems * elem_size, (size_t)row_elems * elem_size,
                                 (size_t)rows);
  p.dstPtr = make_cudaPitchedPtr(host, (size_t)row_elems * elem_size, (size_t)row_elems * elem_size, (size_t)rows);
  p.srcPos = make_cudaPos((size_t)lo2 * elem_size, (size_t)lo1, (size_t)lo0);
  p.dstPos = p.srcPos;
  p.extent = make_cudaExtent((size_t)cnt2 * elem_size, (size_t)cnt1, (size_t)cnt0);
  p.kind = cudaMemcpyDeviceToHost;
  CUDA_TRY(cudaMemcpy3DAsync(&p, c->stream));
  return 0;
}

int mb200_d2d(mb200_ctx *c, void *dst, const void *src, size_t bytes) {
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

int mb200_host_alloc(size_t bytes, void **out) {
  CUDA_TRY(cudaMallocHost(out, bytes ? bytes : 8));
  return 0;
}

int mb200_host_free(void *p) {
  if (p) CUDA_TRY(cudaFreeHost(p));
  return 0;
}

size_t mb200_bytes_allocated(mb200_ctx *c) { return c->bytes_allocated; }

// ---- plans -------------------------------------------------------------------------------------

int mb200_plan_create(mb200_ctx *c, int kind, int dtype, const void *jobs, int njobs,
                      mb200_plan **out) {
  if (!out) return fail("mb200_plan_create: out == NULL");
  *out = nullptr;
  const size_t js = job_size_of(kind);
  if (!js) return fail("mb200_plan_create: unknown kind %d", kind);
  if (dtype != MB200_F64 && dtype != MB200_F32)
    return fail("mb200_plan_create: unknown dtype %d", dtype);
  if (njobs < 0 || (njobs > 0 && !jobs)) return fail("mb200_plan_create: bad job table");
  CUDA_TRY(cudaSetDevice(c->device));
  mb200_plan *p = new mb200_plan();
  p->kind = kind;
  p->dtype = dtype;
  p->njobs = njobs;
  p->job_size = js;
  p->d_jobs = nullptr;
  p->d_prefix = nullptr;
  p->d_group = nullptr;
  p->lean = nullptr;
  p->bytes = p->points = 0;
  p->all_plain = kind == MB200_K_STEP3 && njobs > 0;
  if (kind == MB200_K_STEP3)
    for (int j = 0; j < njobs; ++j)
      if (!step3_is_plain(((const mb200_step3_job_t *)jobs)[j])) p->all_plain = false;
  if (kind == MB200_K_LORENTZ) {
    int nblocked = 0;
    for (int j = 0; j < njobs; ++j)
      if (lorentz_blocked_ok(((const mb200_lorentz_job_t *)jobs)[j])) ++nblocked;
    if (nblocked != 0 && nblocked != njobs) {
      delete p;
      return fail("mb200_plan_create: a Lorentz plan must not mix zero-block jobs with plain ones");
    }
    p->all_plain = njobs > 0 && nblocked == njobs;
  }
  std::vector<int64_t> prefix(njobs + 1, 0);
  for (int j = 0; j < njobs; ++j) {
    int64_t t;
    double b, pts;
    job_metrics(kind, dtype, jobs, j, &t, &b, &pts);
    prefix[j + 1] = prefix[j] + t;
    p->bytes += b;
    p->points += pts;
  }
  // E = chi1inv D with off-diagonal chi1inv: the job of one component also reads the D arrays of the
  // other two (4-point averages).  Emitted per component, each job sweeps the chunk long after the
  // previous one and every D array comes from DRAM three times.  Three consecutive jobs over the
  // same chunk (same strides, boxes within one point of each other) are therefore INTERLEAVED tile
  // by tile — CTA b of the triple takes tile b / 3 of job b % 3, as step3c_kernel does for its three
  // components — so that the shared operands are found in L2.  Each job of a triple is given the
  // tile count of the largest (surplus CTAs find an empty march).
  std::vector<int> group(njobs);
  bool any_group = false;
  if (kind == MB200_K_EDHB && g_edhb_interleave) {
    const mb200_edhb_job_t *E = (const mb200_edhb_job_t *)jobs;
    for (int j = 0; j < njobs; ++j)
      group[j] = j;
    for (int j = 0; j + 2 < njobs;) {
      bool ok = true;
      for (int k = 0; k < 3 && ok; ++k) {
        ok = (E[j + k].u1 || E[j + k].u2) && !E[j + k].pmlw.sig;
        for (int d = 0; d < 3 && ok; ++d)
          ok = E[j + k].box.s[d] == E[j].box.s[d] && abs(E[j + k].box.n[d] - E[j].box.n[d]) <= 1;
      }
      if (!ok) {
        ++j;
        continue;
      }
      for (int k = 0; k < 3; ++k) // (prefix is rebuilt below)
        group[j + k] = j;
      any_group = true;
      j += 3;
    }
    if (any_group) {
      std::vector<int64_t> tiles_of(njobs);
      for (int j = 0; j < njobs; ++j)
        tiles_of[j] = prefix[j + 1] - prefix[j];
      for (int j = 0; j < njobs; ++j)
        if (group[j] == j && j + 2 < njobs && group[j + 1] == j && group[j + 2] == j) {
          const int64_t t = std::max(tiles_of[j], std::max(tiles_of[j + 1], tiles_of[j + 2]));
          tiles_of[j] = tiles_of[j + 1] = tiles_of[j + 2] = t;
        }
      for (int j = 0; j < njobs; ++j)
        prefix[j + 1] = prefix[j] + tiles_of[j];
    }
  }
  p->tiles = prefix[njobs];
  if (kind == MB200_K_FLUX)
    for (int j = 1; j < njobs; ++j) {
      const mb200_flux_job_t *F = (const mb200_flux_job_t *)jobs;
      if (F[j].nomega != F[0].nomega || F[j].out != F[0].out) {
        delete p;
        return fail("mb200_plan_create: the jobs of a flux plan must share nomega and out");
      }
    }
  if (kind == MB200_K_STEP3 || kind == MB200_K_LORENTZ || kind == MB200_K_FLUX) {
    p->h_jobs.assign((const char *)jobs, (const char *)jobs + js * njobs);
    p->h_prefix = prefix;
  }
  if (p->tiles > 0x7fffffffLL) {
    delete p;
    return fail("mb200_plan_create: too many tiles (%lld)", (long long)p->tiles);
  }
  if (njobs > 0) {
    CUDA_TRY(cudaMalloc(&p->d_jobs, js * njobs));
    CUDA_TRY(cudaMalloc((void **)&p->d_prefix, sizeof(int64_t) * (njobs + 1)));
    CUDA_TRY(cudaMemcpyAsync(p->d_jobs, jobs, js * njobs, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(p->d_prefix, prefix.data(), sizeof(int64_t) * (njobs + 1),
                             cudaMemcpyHostToDevice, c->stream));
    if (kind == MB200_K_STEP3 && p->all_plain && g_plain_lean && njobs <= kMaxJobLaunches) {
      p->lean = new Step3LeanPlan();
      auto upload = [&](const void *host, size_t bytes) -> void * {
        void *d = nullptr;
        if (cudaMalloc(&d, bytes ? bytes : 8) != cudaSuccess) return nullptr;
        p->lean_allocs.push_back(d);
        if (bytes && cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        return d;
      };
      for (int j = 0; j < njobs; ++j) {
        const Step3Shell S = step3_shell(((const mb200_step3_job_t *)jobs)[j]);
        Step3LeanPlan::Job L;
        memset(&L, 0, sizeof(L));
        L.lean = S.lean;
        if (S.lean) {
          L.x_lo = S.lo[0];
          L.x_hi = S.hi[0];
          L.ncols = (int)S.cols.size();
          L.d_cols = (const int *)upload(S.cols.data(), S.cols.size() * sizeof(int));
          L.nslabs = (int)S.slabs.size();
          std::vector<int64_t> sp(S.slabs.size() + 1, 0);
          for (size_t k = 0; k < S.slabs.size(); ++k)
            sp[k + 1] = sp[k] + step3_tiles(S.slabs[k]);
          L.slab_tiles = sp.back();
          L.d_slabs = (const mb200_step3_job_t *)upload(S.slabs.data(), S.slabs.size() * sizeof(mb200_step3_job_t));
          L.d_slab_prefix = (const int64_t *)upload(sp.data(), sp.size() * sizeof(int64_t));
          if (!L.d_cols || !L.d_slabs || !L.d_slab_prefix) {
            mb200_plan_destroy(c, p);
            return fail("mb200_plan_create: out of device memory (lean plan)");
          }
        }
        { // the masked march for the whole job (when the lean march does not pay): its own two-entry prefix
          const int64_t own[2] = {0, prefix[j + 1] - prefix[j]};
          L.d_own_prefix = (const int64_t *)upload(own, sizeof(own));
          if (!L.d_own_prefix) {
            mb200_plan_destroy(c, p);
            return fail("mb200_plan_create: out of device memory (lean plan)");
          }
        }
        p->lean->jobs.push_back(L);
      }
    }
    if (any_group) {
      CUDA_TRY(cudaMalloc((void **)&p->d_group, sizeof(int) * njobs));
      CUDA_TRY(cudaMemcpyAsync(p->d_group, group.data(), sizeof(int) * njobs, cudaMemcpyHostToDevice, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream)); // prefix is a stack-owned vector
  }
  *out = p;
  return 0;
}

void mb200_plan_destroy(mb200_ctx *c, mb200_plan *p) {
  if (!p) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (p->d_jobs) cudaFree(p->d_jobs);
  if (p->d_prefix) cudaFree(p->d_prefix);
  if (p->d_group) cudaFree(p->d_group);
  for (void *d : p->lean_allocs)
    cudaFree(d);
  delete p->lean;
  delete p;
}

double mb200_plan_bytes(const mb200_plan *p) { return p ? p->bytes : 0; }
double mb200_plan_points(const mb200_plan *p) { return p ? p->points : 0; }

int mb200_plan_run(mb200_ctx *c, mb200_plan *p, const void *run_data, size_t run_bytes) {
  if (!p) return fail("mb200_plan_run: plan == NULL");
  if (p->tiles == 0) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  const bool needs_run = p->kind == MB200_K_SOURCE || p->kind == MB200_K_DFT || p->kind == MB200_K_NOISE;
  if (p->kind == MB200_K_FLUX) { // scratch for the per-tile partial sums (first stage of the tree)
    const size_t need = sizeof(double) * (size_t)p->tiles *
                        (size_t)((const mb200_flux_job_t *)p->h_jobs.data())->nomega;
    if (need > c->run_cap) {
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      if (c->run_buf) CUDA_TRY(cudaFree(c->run_buf));
      c->run_cap = need * 2 + 4096;
      CUDA_TRY(cudaMalloc(&c->run_buf, c->run_cap));
    }
  }
  if (needs_run) {
    if (!run_data || !run_bytes) return fail("mb200_plan_run: kind %d needs run_data", p->kind);
    if (run_bytes > c->run_cap) {
      // NB: previous launches may still be reading the old buffer
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      if (c->run_buf) CUDA_TRY(cudaFree(c->run_buf));
      c->run_cap = run_bytes * 2 + 4096;
      CUDA_TRY(cudaMalloc(&c->run_buf, c->run_cap));
    }
    CUDA_TRY(cudaMemcpyAsync(c->run_buf, run_data, run_bytes, cudaMemcpyHostToDevice, c->stream));
  }
  ProfRec rec;
  if (c->profiling) {
    rec.kind = (p->kind == MB200_K_STEP3 && !p->all_plain) ? MB200_K_STEP3_GENERAL : p->kind;
    rec.bytes = p->bytes;
    const double R = p->dtype == MB200_F64 ? 8.0 : 4.0;
    const bool blocked_lorentz = p->kind == MB200_K_LORENTZ && p->all_plain;
    if ((blocked_lorentz || p->kind == MB200_K_FMP) && c->next_slot < kWorkSlots) {
      // bytes actually moved: only the blocks the kernel did not skip (plan_metrics.h counts the
      // dense volume, which is what a kernel without zero-block skipping would move)
      rec.slot = c->next_slot++;
      if (blocked_lorentz) { // P, P_prev r/w + sigma, W r per updated element; 2 flag bytes per block
        rec.per_block = 6.0 * R * MB200_ZBLOCK;
        rec.bytes = 0;
        for (int j = 0; j < p->njobs; ++j)
          rec.bytes += 2.0 * (double)ceil_div(((const mb200_lorentz_job_t *)p->h_jobs.data())[j].ntot, MB200_ZBLOCK);
      }
      else { // D r + f_minus_p w for every element; P r only for the blocks not known to be zero
        rec.per_block = R * MB200_ZBLOCK;
        rec.bytes = 2.0 * R * p->points;
      }
      CUDA_TRY(cudaMemsetAsync(c->d_work, 0, 2 * sizeof(unsigned long long), c->stream));
    }
    CUDA_TRY(cudaEventCreate(&rec.a));
    CUDA_TRY(cudaEventCreate(&rec.b));
    CUDA_TRY(cudaEventRecord(rec.a, c->stream));
  }
  int kernels = 1;
  cudaError_t e = p->dtype == MB200_F64 ? launch_plan<double>(c, p, c->run_buf, &kernels)
                                        : launch_plan<float>(c, p, c->run_buf, &kernels);
  if (e != cudaSuccess)
    return fail("kernel launch (kind %d, %lld tiles) failed: %s", p->kind, (long long)p->tiles,
                cudaGetErrorString(e));
  c->launches += kernels;
  if (c->profiling) {
    CUDA_TRY(cudaEventRecord(rec.b, c->stream));
    if (rec.slot >= 0)
      CUDA_TRY(cudaMemcpyAsync(c->h_work + 2 * rec.slot, c->d_work, 2 * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, c->stream));
    c->recs.push_back(rec);
  }
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

int mb200_block_zero_flags(mb200_ctx *c, int dtype, const void *arr, int64_t n, uint8_t *flags) {
  if (n <= 0) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  const unsigned grid = (unsigned)ceil_div(n, MB200_ZBLOCK);
  if (dtype == MB200_F64)
    block_zero_flags_kernel<double><<<grid, kThreads, 0, c->stream>>>((const double *)arr, n, flags);
  else
    block_zero_flags_kernel<float><<<grid, kThreads, 0, c->stream>>>((const float *)arr, n, flags);
  CUDA_TRY(cudaGetLastError());
  c->launches += 1;
  return 0;
}

int mb200_check_finite(mb200_ctx *c, int dtype, const uint64_t *ptrs, int64_t n, int32_t *flag) {
  if (n <= 0) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  const unsigned grid = (unsigned)ceil_div(n, 256);
  if (dtype == MB200_F64) check_finite_kernel<double><<<grid, 256, 0, c->stream>>>(ptrs, n, flag);
  else check_finite_kernel<float><<<grid, 256, 0, c->stream>>>(ptrs, n, flag);
  CUDA_TRY(cudaGetLastError());
  c->launches += 1;
  return 0;
}

// ---- inter-process exchange (NCCL, resolved lazily so that single-GPU use needs no NCCL) -------
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8 };
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.h) return 0;
  // reuse an NCCL that the process already loaded (e.g. torch's), else the system one
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail("cannot load libnccl.so.2: %s", dlerror());
#define MB200_SYM(field, name)                                                                     \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                                      \
  if (!g_nccl.field) return fail("libnccl lacks %s", name)
  MB200_SYM(GetUniqueId, "ncclGetUniqueId");
  MB200_SYM(CommInitRank, "ncclCommInitRank");
  MB200_SYM(CommDestroy, "ncclCommDestroy");
  MB200_SYM(GroupStart, "ncclGroupStart");
  MB200_SYM(GroupEnd, "ncclGroupEnd");
  MB200_SYM(Send, "ncclSend");
  MB200_SYM(Recv, "ncclRecv");
  MB200_SYM(GetErrorString, "ncclGetErrorString");
#undef MB200_SYM
  g_nccl.h = h;
  return 0;
}
} // namespace

struct mb200_comm {
  ncclComm_t comm;
  int rank, nranks;
};

#define NCCL_TRY(expr)                                                                             \
  do {                                                                                             \
    ncclResult_t r_ = (expr);                                                                      \
    if (r_ != 0) return fail("%s failed: %s", #expr, g_nccl.GetErrorString(r_));                   \
  } while (0)

int mb200_comm_unique_id(void *id128) {
  if (load_nccl()) return 1;
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return 0;
}

int mb200_comm_create(mb200_ctx *c, int rank, int nranks, const void *id128, mb200_comm **out) {
  *out = nullptr;
  if (load_nccl()) return 1;
  CUDA_TRY(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  mb200_comm *m = new mb200_comm();
  m->rank = rank;
  m->nranks = nranks;
  NCCL_TRY(g_nccl.CommInitRank(&m->comm, nranks, id, rank));
  *out = m;
  return 0;
}

void mb200_comm_destroy(mb200_comm *m) {
  if (!m) return;
  if (g_nccl.h) g_nccl.CommDestroy(m->comm);
  delete m;
}

int mb200_comm_exchange(mb200_ctx *c, mb200_comm *m, int dtype, const mb200_xfer_t *sends,
                        int nsend, const mb200_xfer_t *recvs, int nrecv) {
  if (!m) return fail("mb200_comm_exchange: no communicator");
  if (nsend == 0 && nrecv == 0) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  const int dt = dtype == MB200_F64 ? ncclFloat64 : ncclFloat32;
  ProfRec rec;
  if (c->profiling) {
    rec.kind = MB200_K_EXCHANGE;
    rec.bytes = 0;
    const double R = dtype == MB200_F64 ? 8.0 : 4.0;
    for (int k = 0; k < nsend; ++k) rec.bytes += R * (double)sends[k].count;
    for (int k = 0; k < nrecv; ++k) rec.bytes += R * (double)recvs[k].count;
    CUDA_TRY(cudaEventCreate(&rec.a));
    CUDA_TRY(cudaEventCreate(&rec.b));
    CUDA_TRY(cudaEventRecord(rec.a, c->stream));
  }
  NCCL_TRY(g_nccl.GroupStart());
  for (int k = 0; k < nrecv; ++k)
    NCCL_TRY(g_nccl.Recv(recvs[k].buf, (size_t)recvs[k].count, dt, recvs[k].peer, m->comm, c->stream));
  for (int k = 0; k < nsend; ++k)
    NCCL_TRY(g_nccl.Send(sends[k].buf, (size_t)sends[k].count, dt, sends[k].peer, m->comm, c->stream));
  NCCL_TRY(g_nccl.GroupEnd());
  c->launches += 1;
  if (c->profiling) {
    CUDA_TRY(cudaEventRecord(rec.b, c->stream));
    c->recs.push_back(rec);
  }
  return 0;
}

// ---- peer-memory exchange ----------------------------------------------------------------------
__global__ void flag_signal_kernel(volatile uint64_t *flag, uint64_t value) {
  __threadfence_system(); // everything this stream stored before (incl. into peer HBM) is visible
  *flag = value;
}
__global__ void flag_wait_kernel(const volatile uint64_t *flag, uint64_t value, int *err,
                                 long long max_iters) {
  // bounded spin: a lost neighbour must not hang the GPU
  for (long long it = 0; it < max_iters; ++it) {
    if (*flag >= value) {
      __threadfence_system();
      return;
    }
    __nanosleep(100);
  }
  *err = 1;
}

struct FlagSet {
  uint64_t *flag[MB200_MAX_FLAGS];
  uint64_t value[MB200_MAX_FLAGS];
  int n;
};
__global__ void flag_signal_many_kernel(const __grid_constant__ FlagSet S) {
  __threadfence_system();
  if ((int)threadIdx.x < S.n) *(volatile uint64_t *)S.flag[threadIdx.x] = S.value[threadIdx.x];
}
__global__ void flag_wait_many_kernel(const __grid_constant__ FlagSet S, int *err, long long max_iters) {
  if ((int)threadIdx.x >= S.n) return;
  const volatile uint64_t *f = S.flag[threadIdx.x];
  for (long long it = 0; it < max_iters; ++it) {
    if (*f >= S.value[threadIdx.x]) {
      __threadfence_system();
      return;
    }
    __nanosleep(100);
  }
  *err = 1;
}

int mb200_ipc_export(mb200_ctx *c, void *devptr, void *handle64) {
  CUDA_TRY(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, devptr));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
int mb200_ipc_import(mb200_ctx *c, const void *handle64, void **out) {
  CUDA_TRY(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int mb200_ipc_close(mb200_ctx *c, void *imported) {
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaIpcCloseMemHandle(imported));
  return 0;
}
int mb200_flag_signal(mb200_ctx *c, uint64_t *flag, uint64_t value) {
  CUDA_TRY(cudaSetDevice(c->device));
  flag_signal_kernel<<<1, 1, 0, c->stream>>>(flag, value);
  CUDA_TRY(cudaGetLastError());
  c->launches += 1;
  return 0;
}
int mb200_flag_wait(mb200_ctx *c, const uint64_t *flag, uint64_t value) {
  CUDA_TRY(cudaSetDevice(c->device));
  // MEEP_B200_PEER_TIMEOUT_S (default 60): how long a GPU waits for its neighbour's data
  static long long max_iters = 0;
  if (!max_iters) {
    const char *e = getenv("MEEP_B200_PEER_TIMEOUT_S");
    double secs = e ? atof(e) : 60.0;
    if (!(secs > 0)) secs = 60.0;
    max_iters = (long long)(secs * 5e6); // ~200 ns per probe (100 ns sleep + a system-scope load)
  }
  ProfRec rec;
  if (c->profiling) {
    rec.kind = MB200_K_EXCHANGE;
    rec.bytes = 0;
    CUDA_TRY(cudaEventCreate(&rec.a));
    CUDA_TRY(cudaEventCreate(&rec.b));
    CUDA_TRY(cudaEventRecord(rec.a, c->stream));
  }
  flag_wait_kernel<<<1, 1, 0, c->stream>>>(flag, value, c->d_err, max_iters);
  CUDA_TRY(cudaGetLastError());
  c->launches += 1;
  if (c->profiling) {
    CUDA_TRY(cudaEventRecord(rec.b, c->stream));
    c->recs.push_back(rec);
  }
  return 0;
}

static long long peer_timeout_iters() {
  static long long max_iters = 0;
  if (!max_iters) {
    const char *e = getenv("MEEP_B200_PEER_TIMEOUT_S");
    double secs = e ? atof(e) : 60.0;
    if (!(secs > 0)) secs = 60.0;
    max_iters = (long long)(secs * 5e6);
  }
  return max_iters;
}
int mb200_flag_signal_many(mb200_ctx *c, uint64_t *const *flags, const uint64_t *values, int n) {
  if (n <= 0) return 0;
  if (n > MB200_MAX_FLAGS) return fail("mb200_flag_signal_many: more than %d flags", MB200_MAX_FLAGS);
  CUDA_TRY(cudaSetDevice(c->device));
  FlagSet S;
  S.n = n;
  for (int k = 0; k < n; ++k) {
    S.flag[k] = flags[k];
    S.value[k] = values[k];
  }
  flag_signal_many_kernel<<<1, 32, 0, c->stream>>>(S);
  CUDA_TRY(cudaGetLastError());
  c->launches += 1;
  return 0;
}
int mb200_flag_wait_many(mb200_ctx *c, const uint64_t *const *flags, const uint64_t *values, int n) {
  if (n <= 0) return 0;
  if (n > MB200_MAX_FLAGS) return fail("mb200_flag_wait_many: more than %d flags", MB200_MAX_FLAGS);
  CUDA_TRY(cudaSetDevice(c->device));
  FlagSet S;
  S.n = n;
  for (int k = 0; k < n; ++k) {
    S.flag[k] = (uint64_t *)flags[k];
    S.value[k] = values[k];
  }
  ProfRec rec;
  if (c->profiling) {
    rec.kind = MB200_K_EXCHANGE;
    rec.bytes = 0;
    CUDA_TRY(cudaEventCreate(&rec.a));
    CUDA_TRY(cudaEventCreate(&rec.b));
    CUDA_TRY(cudaEventRecord(rec.a, c->stream));
  }
  flag_wait_many_kernel<<<1, 32, 0, c->stream>>>(S, c->d_err, peer_timeout_iters());
  CUDA_TRY(cudaGetLastError());
  c->launches += 1;
  if (c->profiling) {
    CUDA_TRY(cudaEventRecord(rec.b, c->stream));
    c->recs.push_back(rec);
  }
  return 0;
}

// ---- measurement -------------------------------------------------------------------------------
int mb200_timer_start(mb200_ctx *c) {
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaEventRecord(c->t0, c->stream));
  return 0;
}
int mb200_timer_stop(mb200_ctx *c, double *ms) {
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaEventRecord(c->t1, c->stream));
  CUDA_TRY(cudaEventSynchronize(c->t1));
  float f = 0;
  CUDA_TRY(cudaEventElapsedTime(&f, c->t0, c->t1));
  *ms = f;
  return 0;
}

static int prof_collect(mb200_ctx *c) {
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (auto &r : c->recs) {
    float f = 0;
    CUDA_TRY(cudaEventElapsedTime(&f, r.a, r.b));
    c->prof_launches[r.kind] += 1;
    c->prof_ms[r.kind] += f;
    if (r.slot >= 0) // (the stream was synchronised above: the counters have landed)
      r.bytes += r.per_block * (double)c->h_work[2 * r.slot + (r.kind == MB200_K_FMP ? 1 : 0)];
    c->prof_bytes[r.kind] += r.bytes;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  c->recs.clear();
  c->next_slot = 0;
  return 0;
}

int mb200_profile_enable(mb200_ctx *c, int on) {
  if (!on && c->profiling) {
    if (prof_collect(c)) return 1;
  }
  c->profiling = on != 0;
  return 0;
}
int mb200_profile_reset(mb200_ctx *c) {
  if (prof_collect(c)) return 1;
  memset(c->prof_launches, 0, sizeof(c->prof_launches));
  memset(c->prof_ms, 0, sizeof(c->prof_ms));
  memset(c->prof_bytes, 0, sizeof(c->prof_bytes));
  return 0;
}
int mb200_profile_get(mb200_ctx *c, int kind, int64_t *launches, double *ms, double *bytes) {
  if (kind < 0 || kind >= MB200_NUM_KINDS) return fail("mb200_profile_get: bad kind");
  if (prof_collect(c)) return 1;
  if (launches) *launches = c->prof_launches[kind];
  if (ms) *ms = c->prof_ms[kind];
  if (bytes) *bytes = c->prof_bytes[kind];
  return 0;
}
int mb200_mark(mb200_ctx *c, int tag) {
  CUDA_TRY(cudaSetDevice(c->device));
  cudaEvent_t e;
  if (!c->event_pool.empty()) {
    e = c->event_pool.back();
    c->event_pool.pop_back();
  }
  else
    CUDA_TRY(cudaEventCreate(&e));
  CUDA_TRY(cudaEventRecord(e, c->stream));
  c->marks.push_back(std::make_pair(tag, e));
  return 0;
}
int mb200_marks_collect(mb200_ctx *c, int *tags, double *ms, int cap, int *n) {
  CUDA_TRY(cudaSetDevice(c->device));
  *n = 0;
  if (!c->marks.empty()) CUDA_TRY(cudaEventSynchronize(c->marks.back().second));
  for (size_t k = 0; k + 1 < c->marks.size() && *n < cap; ++k) {
    float f = 0;
    CUDA_TRY(cudaEventElapsedTime(&f, c->marks[k].second, c->marks[k + 1].second));
    tags[*n] = c->marks[k].first;
    ms[*n] = f;
    *n += 1;
  }
  for (auto &m : c->marks)
    c->event_pool.push_back(m.second);
  c->marks.clear();
  return 0;
}
int64_t mb200_launch_count(mb200_ctx *c) { return c->launches; }

} // extern "C"
