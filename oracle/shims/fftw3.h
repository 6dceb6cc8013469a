/* oracle/shims/fftw3.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for the FFTW3 API surface the reference touches
 * (/root/reference/src/collisions.c:11-14,59-68,82-87,270): FFTW3 is an un-vendored dependency
 * (CMakeModules/FindFFTW.cmake) and is absent from this image.  The plan executes an exact,
 * unnormalised, in-place separable dense DFT (O(N^4), N<=32), implemented in shim.c. */
#ifndef ORC_SHIM_FFTW3_H
#define ORC_SHIM_FFTW3_H
#include <stddef.h>
typedef double fftw_complex[2];
typedef struct orc_shim_plan *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in, fftw_complex *out, int sign,
                           unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void *fftw_malloc(size_t n);
void fftw_free(void *p);
#endif
