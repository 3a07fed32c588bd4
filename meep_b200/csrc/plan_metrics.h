// plan_metrics.h — per-job tile counts and algorithmic-byte accounting shared by the CUDA
// implementation of the C ABI (capi.cu) and the test-only emulator (tests/emu).
#ifndef MEEP_B200_PLAN_METRICS_H
#define MEEP_B200_PLAN_METRICS_H
#include "kernels.cuh"
#include "fused.cuh"

namespace mb200 {

inline size_t job_size_of(int kind) {
  switch (kind) {
    case MB200_K_CURL: return sizeof(mb200_curl_job_t);
    case MB200_K_EDHB: return sizeof(mb200_edhb_job_t);
    case MB200_K_LORENTZ: return sizeof(mb200_lorentz_job_t);
    case MB200_K_FMP: return sizeof(mb200_fmp_job_t);
    case MB200_K_SOURCE: return sizeof(mb200_src_job_t);
    case MB200_K_HALO: return sizeof(mb200_halo_job_t);
    case MB200_K_ZERO: return sizeof(mb200_zero_job_t);
    case MB200_K_DFT: return sizeof(mb200_dft_job_t);
    case MB200_K_FLUX: return sizeof(mb200_flux_job_t);
    case MB200_K_STEP3: return sizeof(mb200_step3_job_t);
    case MB200_K_BETA: return sizeof(mb200_beta_job_t);
    case MB200_K_BFAST: return sizeof(mb200_bfast_job_t);
    case MB200_K_AVERAGE: return sizeof(mb200_average_job_t);
    case MB200_K_GYRO: return sizeof(mb200_gyro_job_t);
    case MB200_K_NOISE: return sizeof(mb200_noise_job_t);
    case MB200_K_CYLINT: return sizeof(mb200_cylint_job_t);
    case MB200_K_CYLR0: return sizeof(mb200_cylr0_job_t);
    default: return 0;
  }
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline double box_points(const mb200_box_t &b) {
  return (double)b.n[0] * (double)b.n[1] * (double)b.n[2];
}

// tiles, algorithmic bytes and points of job j
inline void job_metrics(int kind, int dtype, const void *jobs, int j, int64_t *tiles, double *bytes,
                        double *points) {
  const double R = dtype == MB200_F64 ? 8.0 : 4.0;
  switch (kind) {
    case MB200_K_CURL: {
      const mb200_curl_job_t &J = ((const mb200_curl_job_t *)jobs)[j];
      *tiles = box_tiles(J.box);
      *points = box_points(J.box);
      int arrays = 2 + 1 + (J.g1 && J.g2 ? 1 : 0) + (J.pmlu.sig ? 2 : 0) +
                   (J.cnd ? 2 + (J.pml.sig ? 2 : 0) : 0);
      *bytes = R * arrays * *points;
      break;
    }
    case MB200_K_EDHB: {
      const mb200_edhb_job_t &J = ((const mb200_edhb_job_t *)jobs)[j];
      *tiles = box_tiles(J.box, kEdhbT1);
      *points = box_points(J.box);
      int arrays = 1 + 1 + (J.u ? 1 : 0) + (J.u1 ? 2 : 0) + (J.u2 ? 2 : 0) + (J.chi3 ? 2 : 0) +
                   (J.pmlw.sig ? 3 : 0);
      *bytes = R * arrays * *points;
      break;
    }
    case MB200_K_LORENTZ: {
      const mb200_lorentz_job_t &J = ((const mb200_lorentz_job_t *)jobs)[j];
      *tiles = lorentz_blocked_ok(J) ? ceil_div(ceil_div(J.ntot, MB200_ZBLOCK), kZBlocksPerCta)
                                     : box_tiles(J.box);
      *points = box_points(J.box);
      int arrays = 4 + 2 + (J.s1 ? 2 : 0) + (J.s2 ? 2 : 0);
      *bytes = R * arrays * *points;
      break;
    }
    case MB200_K_FMP: {
      const mb200_fmp_job_t &J = ((const mb200_fmp_job_t *)jobs)[j];
      *tiles = ceil_div(ceil_div(J.ntot, (int64_t)MB200_ZBLOCK), (int64_t)kFmpBlocks);
      *points = (double)J.ntot;
      *bytes = R * (2 + J.np) * *points;
      break;
    }
    case MB200_K_SOURCE: {
      const mb200_src_job_t &J = ((const mb200_src_job_t *)jobs)[j];
      *tiles = ceil_div(J.npts, kThreads);
      *points = (double)J.npts;
      *bytes = (24 + 2 * R * (J.f_im ? 2 : 1)) * *points;
      break;
    }
    case MB200_K_HALO: {
      const mb200_halo_job_t &J = ((const mb200_halo_job_t *)jobs)[j];
      *tiles = halo_list_tiles(J) + ceil_div(J.nrun, kHaloRunsPerTile);
      *points = (double)(2 * J.n_phase + J.n_negate + J.n_copy);
      if (J.nrun > 0)
        *bytes = (16 + 2 * R) * 2.0 * J.n_phase + 2 * R * (double)(J.n_negate + J.n_copy) +
                 (double)sizeof(mb200_halo_run_t) * J.nrun;
      else
        *bytes = (16 + 2 * R) * *points;
      break;
    }
    case MB200_K_ZERO: {
      const mb200_zero_job_t &J = ((const mb200_zero_job_t *)jobs)[j];
      *tiles = ceil_div(J.n, kThreads);
      *points = (double)J.n;
      *bytes = (8 + R) * *points;
      break;
    }
    case MB200_K_DFT: {
      const mb200_dft_job_t &J = ((const mb200_dft_job_t *)jobs)[j];
      const int64_t npts = (int64_t)J.box.n[0] * J.box.n[1] * J.box.n[2];
      *tiles = ceil_div(npts, kDftPts);
      *points = (double)npts;
      *bytes = (4 * R * J.nomega + R * (J.f_im ? 2 : 1)) * *points;
      break;
    }
    case MB200_K_FLUX: {
      const mb200_flux_job_t &J = ((const mb200_flux_job_t *)jobs)[j];
      *tiles = ceil_div(J.npts, kFluxPts);
      *points = (double)J.npts;
      *bytes = 4 * R * J.nomega * *points;
      break;
    }
    case MB200_K_BETA: {
      const mb200_beta_job_t &J = ((const mb200_beta_job_t *)jobs)[j];
      *tiles = box_tiles(J.box);
      *points = box_points(J.box);
      int arrays = 2 + 1 + (J.cndinv ? 1 : 0) + (J.pmlu.siginv ? 2 : 0) +
                   (J.cndinv && J.pml.siginv ? 2 : 0);
      *bytes = R * arrays * *points;
      break;
    }
    case MB200_K_NOISE: {
      const mb200_noise_job_t &J = ((const mb200_noise_job_t *)jobs)[j];
      *tiles = box_tiles(J.box);
      *points = box_points(J.box);
      *bytes = (2 * R + 8) * *points;
      break;
    }
    case MB200_K_GYRO: {
      const mb200_gyro_job_t &J = ((const mb200_gyro_job_t *)jobs)[j];
      *tiles = box_tiles(J.box);
      *points = box_points(J.box);
      *bytes = R * (12 + 1 + 1 + (J.w[1] ? 1 : 0) + (J.w[2] ? 1 : 0)) * *points;
      break;
    }
    case MB200_K_AVERAGE: {
      const mb200_average_job_t &J = ((const mb200_average_job_t *)jobs)[j];
      *tiles = ceil_div(J.n, kThreads * kItems1D);
      *points = (double)J.n;
      *bytes = 3 * R * *points;
      break;
    }
    case MB200_K_BFAST: {
      const mb200_bfast_job_t &J = ((const mb200_bfast_job_t *)jobs)[j];
      *tiles = box_tiles(J.box);
      *points = box_points(J.box);
      int arrays = 2 + 2 + 1 + (J.g2 ? 1 : 0) + (J.cnd ? 1 : 0) + (J.pmlu.siginv ? 2 : 0) +
                   (J.cnd && J.pml.siginv ? 2 : 0);
      *bytes = R * arrays * *points;
      break;
    }
    case MB200_K_CYLINT: {
      const mb200_cylint_job_t &J = ((const mb200_cylint_job_t *)jobs)[j];
      *tiles = ceil_div(J.sr, kThreads);
      *points = (double)(J.nr + 1) * (double)J.sr;
      *bytes = 2 * R * *points;
      break;
    }
    case MB200_K_CYLR0: {
      const mb200_cylr0_job_t &J = ((const mb200_cylr0_job_t *)jobs)[j];
      *tiles = box_tiles(J.box);
      *points = box_points(J.box);
      *bytes = R * (3 + (J.fm ? 1 : 0) + (J.fcnd ? 4 : 0) + (J.fu ? 2 : 0)) * *points;
      break;
    }
    case MB200_K_STEP3: {
      const mb200_step3_job_t &J = ((const mb200_step3_job_t *)jobs)[j];
      *tiles = step3_tiles(J);
      *points = step3_points(J);
      *bytes = step3_bytes(J, R);
      break;
    }
    default: *tiles = 0; *bytes = 0; *points = 0;
  }
}


} // namespace mb200
#endif
