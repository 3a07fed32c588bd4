/* fdtd_oracle.c — CPU restatement of the inner loops of NanoComp/meep's fields::step() hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load this; the product (meep_b200/lib/*.so) never does and has no CPU execution path.
 *
 * Each function restates one reference loop nest in plain C (no OpenMP, no macros from the
 * reference) and cites the file:line it follows.  The argument structs are the ones of
 * include/meep_b200.h, but every pointer is a HOST pointer here, so the same job description
 * drives the oracle and the CUDA kernels on identical inputs.
 *
 * Pinning: tests/test_oracle.py checks the arithmetic kernels against golden vectors produced by
 * calling the reference's own functions from the unmodified reference build in oracle/_ref
 * (generator: tests/drivers/gen_golden.cpp; fixtures: tests/golden/): step_curl (32 variants),
 * step_beta, step_bfast, step_update_EDHB, lorentzian_susceptibility::update_P,
 * gyrotropic_susceptibility::update_P, dft_chunk::update_dft + dft_flux::flux, and one whole
 * cylindrical fields::step_db (which pins the r-derivative scan, the i*m/r terms, the r = 0 rows
 * and the zeroed rows).  The remaining functions are gathers / scatters / one-line updates
 * (step_source, step_boundaries, subtract_P, zero_metal, average_with_backup, add_noise) whose
 * reference form is a single statement; they are exercised against the full reference build
 * itself (oracle/_ref, which passes the reference's tests/known_results.cpp 13/13) — the second,
 * end-to-end oracle used by tests/test_host_emu.py and tests/test_parity_gpu.py on every array
 * of 40+ simulations.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/meep_b200.h"

#define REAL double
#define SUF _f64
#include "fdtd_oracle_impl.h"
#undef REAL
#undef SUF

#define REAL float
#define SUF _f32
#include "fdtd_oracle_impl.h"
#undef REAL
#undef SUF
