/* Hand-written replacement for the autoconf-generated config.h of the reference
 * (the reference is built here without autotools; see oracle/ref_build/Makefile).
 * TEST/BASELINE INFRASTRUCTURE ONLY. Only feature macros the stepping core needs;
 * MPI, HDF5, libctl, MPB, Harminv, GSL, FFTW are all absent in this image. */
#ifndef MEEP_B200_SHIM_CONFIG_H
#define MEEP_B200_SHIM_CONFIG_H
#define HAVE_IMMINTRIN_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_INTTYPES_H 1
#define HAVE_STDINT_H 1
#define HAVE_UNISTD_H 1
#define PACKAGE_VERSION "1.35.0-beta"
#define PACKAGE_NAME "meep"
#define restrict __restrict
#define F77_FUNC(name, NAME) name##_
#endif
