/* meep_b200.h — C ABI of the B200 (sm_100a) FDTD time-stepping engine.
 *
 * This is the drop-in boundary for ONE path of NanoComp/meep: meep::fields::step() and the
 * inner loops it drives.  Every entry point takes plain pointers and sizes only (no C++ types,
 * no torch types).  Each "job" struct is the argument list of one reference inner-loop function
 * with the grid_volume / ivec arguments already reduced to the integers the reference's loop
 * macros derive from them; the citation next to each struct names the reference interface it
 * replaces (paths relative to the reference repository root).
 *
 * Conventions
 *  - All array pointers inside jobs are DEVICE pointers obtained from mb200_malloc().
 *  - Arrays use the reference's own per-chunk layout (src/vec.cpp:482-494: (nx+1)(ny+1)(nz+1),
 *    last direction fastest); index arithmetic is therefore identical on host and device.
 *  - dtype selects realnum: MB200_F64 (reference default) or MB200_F32 (--enable-single,
 *    src/meep.hpp:42-46).  Scalars travel as double and are narrowed to realnum in the kernel,
 *    exactly where the reference narrows them (a realnum function parameter).
 *  - A "plan" is a batch of jobs of one kind uploaded once (descriptor table + tile map in HBM)
 *    and launched as ONE grid per run; plans are rebuilt only when the chunk layout changes.
 *  - Everything is asynchronous on the context's stream; mb200_sync()/mb200_d2h() synchronise.
 *  - Return value 0 = success; otherwise nonzero and mb200_last_error() describes the failure.
 *    There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MEEP_B200_H
#define MEEP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB200_ABI_VERSION 2

enum { MB200_F64 = 0, MB200_F32 = 1 };

typedef struct mb200_ctx mb200_ctx;
typedef struct mb200_plan mb200_plan;

/* ---- loop box: what LOOP_OVER_IVECS(gv, is, ie, idx) expands to (src/meep/vec.hpp:151-169):
 *      idx = idx0 + i1*s[0] + i2*s[1] + i3*s[2],  0 <= ik < n[k].  Unused loops have n = 1. */
typedef struct {
  int64_t idx0;
  int64_t s[3];
  int32_t n[3];
  int32_t reserved;
} mb200_box_t;

/* ---- PML coefficient lookup: KSTRIDE_DEF / KDEF (src/meep_internals.hpp:217-226):
 *      k = k0 + ks[0]*i1 + ks[1]*i2 + ks[2]*i3 indexes the 1-D arrays sig/kap/siginv
 *      (src/structure.cpp:665-688).  sig == NULL means "NO_DIRECTION" (term absent). */
typedef struct {
  const void *sig, *kap, *siginv;
  int32_t k0;
  int32_t ks[3];
} mb200_pml_t;

/* ---- step_curl (src/meep_internals.hpp:90-95, src/step_generic.cpp:65-249), called from
 *      fields_chunk::step_db (src/step_db.cpp:124-127).  Same semantics incl. the g1==NULL swap. */
typedef struct {
  mb200_box_t box;
  void *f;
  const void *g1, *g2;
  int64_t s1, s2;
  double dtdx, dt;
  mb200_pml_t pml;  /* dsig : sig, kap, siginv */
  mb200_pml_t pmlu; /* dsigu: sigu, kapu, siginvu */
  void *fu;
  const void *cnd, *cndinv;
  void *fcnd;
} mb200_curl_job_t;

/* ---- step_update_EDHB (src/meep_internals.hpp:97-101, src/step_generic.cpp:566-785), called
 *      from fields_chunk::update_eh (src/update_eh.cpp:190-195).  pmlw.sig/kap = sigw/kapw
 *      (pmlw.siginv is unused).  Same semantics incl. the g1/g2 swap rule (line 573). */
typedef struct {
  mb200_box_t box;
  void *f;
  const void *g, *g1, *g2;
  const void *u, *u1, *u2;
  int64_t s, s1, s2;
  const void *chi2, *chi3;
  void *fw;
  mb200_pml_t pmlw;
} mb200_edhb_job_t;

/* ---- step_beta (src/meep_internals.hpp:103-105, src/step_generic.cpp:255-333), called from
 *      fields_chunk::step_db for 2-D cells with an exp(i beta z) dependence (src/step_db.cpp:161-175):
 *      f += betadt * g (with the conductivity / PML / f_u variants).  pml.siginv = siginv,
 *      pmlu.siginv = siginvu (sig/kap unused; siginv == NULL means NO_DIRECTION). */
typedef struct {
  mb200_box_t box;
  void *f;
  const void *g;
  double betadt;
  mb200_pml_t pml, pmlu;
  void *fu;
  const void *cndinv;
  void *fcnd;
  /* cylindrical i*m/r terms (src/step_db.cpp:178-280) are the same eight loops with a factor that
   * depends on the radial loop index: cyl != 0 -> factor = betadt / (r_is2 + 2 * i2), betadt
   * holding the reference's `the_m` and r_is2 the reference's loop_is2. */
  int32_t cyl, r_is2;
} mb200_beta_job_t;

/* ---- step_bfast (src/meep_internals.hpp:107-112, src/step_generic.cpp:335-530): the BFAST
 *      correction for Bloch-periodic oblique incidence, called from fields_chunk::step_db right
 *      after step_curl of the same component (src/step_db.cpp:129-143):
 *        F' = k1 (g1[i+s1] + g1[i]) - k2 (g2[i+s2] + g2[i]) - F ;  f += (F' - F) [cndinv][siginv]...
 *      g1 != NULL (the caller applies the swap of lines 342-346, k1/k2 included).  The reference's
 *      variant without PML, f_u, conductivity AND g2 stores F' = k1 (g1[i+s1] + g1[i]) without
 *      subtracting the old F (line 372); reproduced. */
typedef struct {
  mb200_box_t box;
  void *f;
  const void *g1, *g2;
  int64_t s1, s2;
  double k1, k2;
  mb200_pml_t pml, pmlu; /* only siginv is used */
  void *fu;
  const void *cnd, *cndinv; /* cnd != NULL selects the conductivity variants (its values are unused) */
  void *fcnd;
  void *F; /* f_bfast */
} mb200_bfast_job_t;

/* ---- gyrotropic_susceptibility::update_P (src/meep.hpp:304-338, src/susceptibility.cpp:445-584):
 *      three polarisation components at the Yee position of one driving-field component.
 *      Arrays and tensors are given in the rotated frame (d0, d1, d2) = (direction of the field
 *      component, the next two cyclic directions): p[k] / pp[k] = P, P_prev along d_k,
 *      w[0] = the field itself, w[1], w[2] = the other two components (may be NULL),
 *      gt[a][b] = gyro_tensor[d_a][d_b], inv[a][b] = inv[d_a][d_b] (the reference's precomputed
 *      3x3 inverse).  model 0: GYROTROPIC_LORENTZIAN / GYROTROPIC_DRUDE (lines 454-510; c[0] = diag,
 *      c[1] = gamma1, c[2] = omega0dtsqr, c[3] = pt), model 1: GYROTROPIC_SATURATED (lines 512-578;
 *      c[0] = omega2pidt, c[1] = g2pidt, c[2] = alpha, c[3] = dt2pi). */
typedef struct {
  mb200_box_t box;
  void *p[3], *pp[3];
  const void *w[3];
  const void *s;
  int64_t is, is1, is2;
  double c[4];
  double gt[3][3], inv[3][3];
  int32_t model, reserved;
} mb200_gyro_job_t;

/* ---- noisy_lorentzian_susceptibility::update_P, the noise term (src/susceptibility.cpp:317-339):
 *      p[i] += gaussian_random(0, amp sqrt(sigma[i])) over the owned points.  The random numbers
 *      come from the reference's own generator on the host, drawn in the reference's loop order
 *      (that is what makes runs reproducible against the CPU build); they are the plan's run data:
 *      noise[slot + (i1 n2 + i2) n3 + i3] (double) is added to the point of loop indices (i1,i2,i3). */
typedef struct {
  mb200_box_t box;
  void *p;
  int64_t slot;
} mb200_noise_job_t;

/* ---- fields_chunk::average_with_backup (src/energy_and_flux.cpp:139-147), the last stage of
 *      fields::synchronize_magnetic_fields: f[i] = 0.5 * (f[i] + backup[i]) over a whole array */
typedef struct {
  void *f;
  const void *backup;
  int64_t n;
} mb200_average_job_t;

/* ---- cylindrical helper array (src/step_db.cpp:93-116): out = running sum over r of
 *      1/r d(r f_p)/dr, so that the unmodified step_curl produces the Z-component update.
 *      One thread per z column, serial in r (the reference's summation order). */
typedef struct {
  void *out;
  const void *fp;
  int64_t nr, sr; /* gv.nr(), gv.nz() + 1 */
  double ir0;
} mb200_cylint_job_t;

/* ---- r = 0 row of a cylindrical chunk with origin_r == 0 (src/step_db.cpp:285-377):
 *      mode 0 (m == 0, Dz):         dfcnd = fp[i] * c                       (c = 4 Courant)
 *      mode 1 (|m| == 1, Dp or Br): dfcnd = c * (fp[i] - fp[i-sd] - mult * fm[i])   (c = sd Courant)
 *      followed by the conductivity / PML / u updates of lines 308-320 = 358-370.
 *      f = the reference's `the_f` (f_u[cc] when fu != NULL), fu = then f[cc]. */
typedef struct {
  mb200_box_t box;
  void *f, *fu;
  const void *fp, *fm;
  int64_t sd;
  double c, mult, dt;
  int32_t mode, reserved;
  const void *cnd, *cndinv;
  void *fcnd;
  mb200_pml_t pml, pmlu;
} mb200_cylr0_job_t;

/* ---- lorentzian_susceptibility::update_P (src/susceptibility.cpp:188-262), one job per
 *      (component, cmp).  Constants are computed by the caller in realnum arithmetic exactly as
 *      lines 192-195 do.  s1/w1 (and s2/w2) NULL => isotropic / 2x2 cases. */
typedef struct {
  mb200_box_t box;
  void *p, *pp;
  const void *w, *s;
  const void *w1, *s1, *w2, *s2;
  int64_t is, is1, is2;
  double gamma1inv, gamma1, omega0dtsqr, omega0dtsqr_denom;
  /* optional zero-block skipping (isotropic jobs on the standard 3-D layout only).  The arrays
   * are viewed as blocks of MB200_ZBLOCK consecutive elements; szero[b] != 0 says sigma is
   * identically zero in block b (set by mb200_block_zero_flags when sigma is uploaded),
   * pzero[b] != 0 says P and P_prev are.  A block with both flags set is skipped — the update
   * would read zeros and write zeros (src/susceptibility.cpp:252-257 with s = p = pp = 0) —
   * otherwise it is updated and pzero[b] is recomputed from the values written.  A dispersive
   * object that fills a few percent of the cell (BASELINE config 3) then costs a few percent of
   * the polarisation traffic the reference pays on every chunk (src/susceptibility.cpp:66-75). */
  uint8_t *pzero;
  const uint8_t *szero;
  int64_t ntot;
} mb200_lorentz_job_t;
#define MB200_ZBLOCK 1024

/* ---- f_minus_p initialisation: memcpy D -> f_minus_p (src/update_eh.cpp:114-120) followed by
 *      lorentzian_susceptibility::subtract_P for each polarisation (src/susceptibility.cpp:264-281):
 *      fmp[i] = (d ? d[i] : fmp[i]) - sum_k p[k][i],  0 <= i < ntot. */
#define MB200_MAX_P 8
typedef struct {
  void *fmp;
  const void *d;
  const void *p[MB200_MAX_P];
  int32_t np;
  int32_t reserved;
  int64_t ntot;
  const uint8_t *pzero[MB200_MAX_P]; /* optional: pzero[k][b] != 0 => p[k] is zero in block b */
} mb200_fmp_job_t;

/* ---- fields_chunk::step_source (src/step.cpp:295-318) [mode 0] and the integrated-source dipole
 *      subtraction in fields_chunk::update_eh (src/update_eh.cpp:128-138) [mode 1].
 *      amp = src_vol::amp (complex<double>, interleaved), index = src_vol::index.
 *      Per run the caller supplies scalars[scalar_slot] = src_time::current() (mode 0) or
 *      src_time::dipole() (mode 1) as interleaved complex<double>.
 *      mode 0: A = amp[j]*scalar*dt*(cndinv ? cndinv[i] : 1);  mode 1: A = amp[j]*scalar;
 *      f_re[i] -= Re A;  if (f_im) f_im[i] -= Im A.   (computed in double, narrowed at the store) */
typedef struct {
  void *f_re, *f_im;
  const void *cndinv;
  const int64_t *index;
  const double *amp;
  int64_t npts;
  double dt;
  int32_t scalar_slot;
  int32_t mode;
} mb200_src_job_t;

/* ---- chunk-boundary exchange: fields::step_boundaries gather + process_incoming_chunk_data
 *      scatter (src/step.cpp:172-223, 251-278) for chunk pairs resident on the same device.
 *      src/dst are device arrays of device ADDRESSES (the translated connections_out /
 *      connections_in vectors, src/meep.hpp:1476-1479), ordered PHASE || NEGATE || COPY:
 *      first 2*n_phase entries are (re,im) pairs multiplied by phase[k] (complex<realnum>),
 *      next n_negate entries are negated, the last n_copy are copied. */
typedef struct {
  uint64_t src0, dst0; /* device addresses of the first element */
  int64_t dsrc, ddst;  /* byte strides */
  int32_t n;           /* elements */
  int32_t negate;      /* 1: NEGATE class, 0: COPY class */
} mb200_halo_run_t;

typedef struct {
  const uint64_t *src;
  const uint64_t *dst;
  const void *phase;
  int64_t n_phase, n_negate, n_copy;
  /* optional, same length as dst: address of the zero-block flag byte (see
   * mb200_lorentz_job_t.pzero) that covers dst[k]; cleared when a non-zero value is stored, so
   * that polarisation values arriving from a neighbouring chunk keep the flags exact */
  const uint64_t *dst_flag;
  /* optional run-length form of the NEGATE || COPY entries (chunk faces are regular planes, so
   * the address lists are runs of constant stride): when nrun > 0 the n_negate + n_copy
   * transfers are the elements of runs[0..nrun) and src/dst hold only the 2*n_phase PHASE
   * entries (may be NULL when n_phase == 0); dst_flag must be NULL.  16 bytes of addresses per
   * transfer become 40 bytes per run. */
  const mb200_halo_run_t *runs;
  int64_t nrun;
} mb200_halo_job_t;

/* ---- fields_chunk::zero_metal (src/boundaries.cpp:310-313): *ptrs[k] = 0. */
typedef struct {
  const uint64_t *ptrs;
  int64_t n;
} mb200_zero_job_t;

/* ---- dft_chunk::update_dft (src/dft.cpp:266-308).  wgt_* are the per-loop-direction edge
 *      weights of IVEC_LOOP_WEIGHT (src/meep/vec.hpp:372-383): s0,s1,e0,e1 .in_direction(loop_dk).
 *      Per run the caller supplies the phase table dft_phase[] (complex<realnum>, computed in
 *      double on the host as polar(1, omega*t)*scale then narrowed: src/dft.cpp:270-271);
 *      this job reads phases[phase_slot .. phase_slot+nomega).
 *      dft[nomega*p + i] += phase[i] * f,  p = (i1*n2 + i2)*n3 + i3. */
typedef struct {
  mb200_box_t box;
  const void *f_re, *f_im;
  int64_t avg1, avg2;
  double wgt_s0[3], wgt_s1[3], wgt_e0[3], wgt_e1[3];
  double dV0, dV1;
  int32_t use_weights, sqrt_weights;
  void *dft;
  int32_t nomega;
  int32_t phase_slot;
} mb200_dft_job_t;

/* ---- dft_flux::flux inner sum (src/dft.cpp:542-556): out[i] += sum_k Re(E[k*nomega+i] *
 *      conj(H[k*nomega+i])), accumulated in double on the device (out: double[nomega]). */
typedef struct {
  const void *e, *h;
  int64_t npts;
  int32_t nomega;
  int32_t reserved;
  double *out;
} mb200_flux_job_t;

/* ---- fused half-step over one 3-D chunk: the three step_curl calls that fields_chunk::step_db
 *      makes for one field type and cmp (src/step_db.cpp:47-127), and — where it is legal
 *      (diagonal chi1inv, no chi2/chi3, no f_minus_p) — the step_update_EDHB call that
 *      fields_chunk::update_eh makes for the same component (src/update_eh.cpp:190-195), in ONE
 *      pass over the chunk, so every array is read once and written once per half-step
 *      (9R + 15R = 24R bytes per cell-step).  Each component keeps exactly the
 *      step_curl / step_update_EDHB argument meaning; lo/hi are the inclusive array-index
 *      ranges of little_owned_corner0(c)..big_corner (src/meep/vec.hpp:1102-1104) and the pml
 *      lookups are re-based to array index (0,0,0): k = k0 + ks[0]*ix + ks[1]*iy + ks[2]*iz.
 *      A component with f == NULL is skipped; e == NULL means "no fused E/H update".
 *      Only array indices ix_lo <= ix <= ix_hi along direction 0 are processed, so that a chunk
 *      can be cut into slabs (e.g. the planes that hold source points, where step_source must
 *      run between the D update and the E update, are launched without the fused E update). */
typedef struct {
  int32_t lo[3], hi[3];
  void *f;
  const void *g1, *g2;
  int64_t s1, s2;
  double dtdx;
  mb200_pml_t pml, pmlu;
  void *fu;
  const void *cnd, *cndinv;
  void *fcnd;
  void *e;
  const void *u;
  void *fw;
  mb200_pml_t pmlw;
  /* planes (array index along each direction, -1 = none) on which fields_chunk::zero_metal
   * (src/boundaries.cpp:310-313) zeroes f right after this update: the fused E/H update uses
   * f = 0 there, exactly what update_eh sees in the reference's order of operations */
  int32_t metal_lo[3], metal_hi[3];
} mb200_step3_comp_t;

typedef struct {
  int32_t n[3]; /* chunk size in cells; the index box is [0..n[d]] per direction */
  int32_t reserved; /* planes of direction 0 marched per CTA (0 = default 8) */
  int64_t stride[3];
  double dt;
  int32_t ix_lo, ix_hi;
  /* the noepi_n planes from noepi_lo on (array index along direction 0; n = 0: none) are updated WITHOUT the
   * fused E/H epilogue: they hold source points, so fields::step_source must run between the D and
   * the E update there (src/step.cpp:98-109) and update_eh covers them with its own pass.  One job
   * (one launch) per chunk instead of three. */
  int32_t noepi_lo, noepi_n;
  mb200_step3_comp_t c[3];
} mb200_step3_job_t;

enum {
  MB200_K_CURL = 0,
  MB200_K_EDHB = 1,
  MB200_K_LORENTZ = 2,
  MB200_K_FMP = 3,
  MB200_K_SOURCE = 4,
  MB200_K_HALO = 5,
  MB200_K_ZERO = 6,
  MB200_K_DFT = 7,
  MB200_K_FLUX = 8,
  MB200_K_STEP3 = 9,
  MB200_K_BETA = 10,
  MB200_K_EXCHANGE = 11, /* not a plan kind: profiling slot of mb200_comm_exchange */
  MB200_K_CYLINT = 12,
  MB200_K_CYLR0 = 13,
  MB200_K_STEP3_GENERAL = 14, /* not a plan kind: profiling slot of MB200_K_STEP3 plans that run the
                                 general (PML) fused kernel; slot 9 then holds the fast-path plans */
  MB200_K_BFAST = 15,
  MB200_K_AVERAGE = 16,
  MB200_K_GYRO = 17,
  MB200_K_NOISE = 18,
  MB200_NUM_KINDS = 19
};

/* ---- context ------------------------------------------------------------------------------- */
int mb200_abi_version(void);
const char *mb200_last_error(void);
/* number of visible CUDA devices (0 if none / driver missing) */
int mb200_device_count(void);
/* create a context on CUDA device `device` (own non-blocking stream). */
int mb200_init(int device, mb200_ctx **out);
void mb200_destroy(mb200_ctx *ctx);
int mb200_sync(mb200_ctx *ctx);

/* ---- device memory ------------------------------------------------------------------------- */
int mb200_malloc(mb200_ctx *ctx, size_t bytes, void **out);
int mb200_free(mb200_ctx *ctx, void *p);
int mb200_memset(mb200_ctx *ctx, void *p, int value, size_t bytes);
int mb200_h2d(mb200_ctx *ctx, void *dst, const void *src, size_t bytes); /* async if src pinned */
int mb200_d2h(mb200_ctx *ctx, void *dst, const void *src, size_t bytes); /* synchronises */
/* the same copy queued on the context's stream WITHOUT synchronising: a download of many arrays is
 * one mb200_sync after the last of them instead of one stream synchronisation per array */
int mb200_d2h_async(mb200_ctx *ctx, void *dst, const void *src, size_t bytes);
/* Sub-box of a 3-D array in the reference layout (planes of `rows` rows of `row_elems` elements;
 * last index fastest): copies elements [lo0, lo0+cnt0) x [lo1, lo1+cnt1) x [lo2, lo2+cnt2) of the
 * device array to the same positions of the host array (one 3-D copy; does NOT synchronise — call
 * mb200_sync when all boxes are queued).  What the host readers of a sub-volume need
 * (fields::loop_in_chunks consumers: flux planes, slices, integrals over a box) instead of every array. */
int mb200_d2h_box(mb200_ctx *ctx, void *host, const void *dev, size_t elem_size, int64_t rows, int64_t row_elems,
                  int64_t lo0, int64_t lo1, int64_t lo2, int64_t cnt0, int64_t cnt1, int64_t cnt2);
int mb200_d2d(mb200_ctx *ctx, void *dst, const void *src, size_t bytes);
int mb200_host_alloc(size_t bytes, void **out); /* pinned host memory */
int mb200_host_free(void *p);
/* bytes currently allocated through mb200_malloc on this context */
size_t mb200_bytes_allocated(mb200_ctx *ctx);

/* ---- plans --------------------------------------------------------------------------------- */
/* jobs: host array of njobs structs of the kind's job type.  The table is copied; the caller
 * may free it.  Device pointers inside must stay valid for the life of the plan. */
int mb200_plan_create(mb200_ctx *ctx, int kind, int dtype, const void *jobs, int njobs,
                      mb200_plan **out);
/* run_data: per-run side input, copied to the device before launch:
 *   MB200_K_SOURCE: interleaved complex<double> scalars[], run_bytes = 16*nslots
 *   MB200_K_DFT   : complex<realnum> phase table, run_bytes = 2*sizeof(realnum)*nphases
 *   MB200_K_NOISE : double noise[], run_bytes = 8*(number of loop points of all jobs)
 *   others        : NULL / 0 */
int mb200_plan_run(mb200_ctx *ctx, mb200_plan *plan, const void *run_data, size_t run_bytes);
void mb200_plan_destroy(mb200_ctx *ctx, mb200_plan *plan);
/* algorithmic HBM bytes one run of this plan must move (each distinct array element read once
 * and written once; neighbour re-reads served on chip) — used for roofline accounting. */
double mb200_plan_bytes(const mb200_plan *plan);
/* number of loop points (cells) one run updates */
double mb200_plan_points(const mb200_plan *plan);

/* ---- one-shot, reference-shaped calls (plan_create + plan_run + plan_destroy) --------------- */
int mb200_step_curl(mb200_ctx *ctx, int dtype, const mb200_curl_job_t *jobs, int njobs);
int mb200_step_update_EDHB(mb200_ctx *ctx, int dtype, const mb200_edhb_job_t *jobs, int njobs);
int mb200_lorentzian_update_P(mb200_ctx *ctx, int dtype, const mb200_lorentz_job_t *jobs,
                              int njobs);
int mb200_subtract_P(mb200_ctx *ctx, int dtype, const mb200_fmp_job_t *jobs, int njobs);
int mb200_add_noise(mb200_ctx *ctx, int dtype, const mb200_noise_job_t *jobs, int njobs,
                    const double *noise, int64_t nnoise);
int mb200_gyrotropic_update_P(mb200_ctx *ctx, int dtype, const mb200_gyro_job_t *jobs, int njobs);
int mb200_step_source(mb200_ctx *ctx, int dtype, const mb200_src_job_t *jobs, int njobs,
                      const double *scalars, int nslots);
int mb200_step_boundaries(mb200_ctx *ctx, int dtype, const mb200_halo_job_t *jobs, int njobs);
int mb200_zero_metal(mb200_ctx *ctx, int dtype, const mb200_zero_job_t *jobs, int njobs);
int mb200_update_dft(mb200_ctx *ctx, int dtype, const mb200_dft_job_t *jobs, int njobs,
                     const void *phases, int nphases);
int mb200_dft_flux(mb200_ctx *ctx, int dtype, const mb200_flux_job_t *jobs, int njobs);
int mb200_step3(mb200_ctx *ctx, int dtype, const mb200_step3_job_t *jobs, int njobs);
int mb200_step_beta(mb200_ctx *ctx, int dtype, const mb200_beta_job_t *jobs, int njobs);
int mb200_step_bfast(mb200_ctx *ctx, int dtype, const mb200_bfast_job_t *jobs, int njobs);
int mb200_average_with_backup(mb200_ctx *ctx, int dtype, const mb200_average_job_t *jobs, int njobs);
/* cylindrical coordinates: src/step_db.cpp:93-116 and 285-377 */
int mb200_cyl_rderiv_int(mb200_ctx *ctx, int dtype, const mb200_cylint_job_t *jobs, int njobs);
int mb200_cyl_origin(mb200_ctx *ctx, int dtype, const mb200_cylr0_job_t *jobs, int njobs);

/* ---- inter-process chunk exchange (one process per GPU).  Replaces the transport half of
 *      fields::step_boundaries — comms_manager::send_real_async / receive_real_async over
 *      MPI_Isend/MPI_Irecv (src/step.cpp:233-271, src/mympi.cpp:102-128) — with device-to-device
 *      transfers over NVLink: the comm blocks (src/meep.hpp:2301-2315, one contiguous block of
 *      comm_size_tot realnums per field type and chunk pair) are packed and unpacked on the device
 *      by halo jobs and moved by one grouped ncclSend/ncclRecv per phase on the context's stream.
 *      id: 128 opaque bytes made by mb200_comm_unique_id on one rank and given to all ranks. */
typedef struct mb200_comm mb200_comm;
typedef struct {
  int32_t peer;   /* rank of the other process */
  int32_t reserved;
  void *buf;      /* contiguous device buffer */
  int64_t count;  /* realnums */
} mb200_xfer_t;
int mb200_comm_unique_id(void *id128);
int mb200_comm_create(mb200_ctx *ctx, int rank, int nranks, const void *id128, mb200_comm **out);
void mb200_comm_destroy(mb200_comm *comm);
int mb200_comm_exchange(mb200_ctx *ctx, mb200_comm *comm, int dtype, const mb200_xfer_t *sends,
                        int nsend, const mb200_xfer_t *recvs, int nrecv);

/* ---- peer-memory exchange (processes on one NVLink/NVSwitch node).  Instead of handing the comm
 *      blocks to a library, the receiving process exports its block arena with CUDA IPC, the
 *      sending process maps it, and the SAME halo kernel that gathers the outgoing values stores
 *      them straight into the neighbour's HBM over NVLink (its dst addresses are peer
 *      addresses).  Ordering between the two GPUs uses sequence words in the arenas:
 *      mb200_flag_signal publishes a value (after a system-scope fence, in stream order),
 *      mb200_flag_wait makes the stream wait until the word reaches a value (bounded spin; on
 *      time-out an error is latched and reported by the next mb200_sync). */
int mb200_ipc_export(mb200_ctx *ctx, void *devptr, void *handle64);
int mb200_ipc_import(mb200_ctx *ctx, const void *handle64, void **out);
int mb200_ipc_close(mb200_ctx *ctx, void *imported);
int mb200_flag_signal(mb200_ctx *ctx, uint64_t *flag, uint64_t value);
int mb200_flag_wait(mb200_ctx *ctx, const uint64_t *flag, uint64_t value);
/* the same for up to MB200_MAX_FLAGS words in ONE launch (a GPU has up to 7 neighbours in a 2x2x2
 * partition; one launch per word would put ~30 one-thread kernels into every time step) */
#define MB200_MAX_FLAGS 32
int mb200_flag_signal_many(mb200_ctx *ctx, uint64_t *const *flags, const uint64_t *values, int n);
int mb200_flag_wait_many(mb200_ctx *ctx, const uint64_t *const *flags, const uint64_t *values, int n);

/* ---- flags[b] = 1 if arr[b*MB200_ZBLOCK .. ) is identically zero, else 0 (n elements) */
int mb200_block_zero_flags(mb200_ctx *ctx, int dtype, const void *arr, int64_t n, uint8_t *flags);

/* ---- finiteness probe (replaces the per-step host read in fields::step, src/step.cpp:137-138):
 *      sets *flag (device int32) to 1 if any of the n listed array elements is NaN/Inf. */
int mb200_check_finite(mb200_ctx *ctx, int dtype, const uint64_t *ptrs, int64_t n, int32_t *flag);

/* ---- measurement --------------------------------------------------------------------------- */
/* CUDA-event stopwatch on the context's stream. */
int mb200_timer_start(mb200_ctx *ctx);
int mb200_timer_stop(mb200_ctx *ctx, double *ms); /* synchronises */
/* per-kind profiling: when on, every plan_run is bracketed by CUDA events. */
int mb200_profile_enable(mb200_ctx *ctx, int on);
int mb200_profile_reset(mb200_ctx *ctx);
/* sums since the last reset (synchronises): launches, device ms, algorithmic bytes */
int mb200_profile_get(mb200_ctx *ctx, int kind, int64_t *launches, double *ms, double *bytes);
/* Phase marks: mb200_mark records a CUDA event on the context's stream, labelled `tag`;
 * mb200_marks_collect waits for the last mark and returns, for every pair of consecutive marks,
 * the EARLIER mark's tag and the device time between the two (ms), then forgets all marks.
 * Replaces the host wall clocks that the reference's timing_scope puts around each phase of
 * fields::step (src/time.cpp:92-110, src/step.cpp:64-121): launches are asynchronous here, so a
 * host clock would measure launch latency.  Returns the number of intervals written (<= cap). */
int mb200_mark(mb200_ctx *ctx, int tag);
int mb200_marks_collect(mb200_ctx *ctx, int *tags, double *ms, int cap, int *n);
/* total kernels launched by this context since creation */
int64_t mb200_launch_count(mb200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MEEP_B200_H */
