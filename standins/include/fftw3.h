/*
 * Stand-in for <fftw3.h> (+ the fftw_mpi_* slab API in fftw3-mpi.h).
 * TEST INFRASTRUCTURE ONLY: lets the unmodified reference compile without FFTW
 * (SURVEY.md section 0).  Backed by standins/shim_fft.c, a self-contained
 * CPU FFT.  Semantics reproduced: unnormalised transforms, in-place padded
 * r2c/c2r layout [x][y][2*(n2/2+1)], c2r = c2c over x,y then c2r over z with the
 * imaginary parts of the kz = 0 and kz = n2/2 inputs ignored (FFTW behaviour).
 */
#ifndef MGP_SHIM_FFTW3_H
#define MGP_SHIM_FFTW3_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef float  fftwf_complex[2];

struct mgp_shim_plan;
typedef struct mgp_shim_plan *fftw_plan;
typedef struct mgp_shim_plan *fftwf_plan;

#define FFTW_MEASURE  (0U)
#define FFTW_ESTIMATE (1U << 6)

void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);

/* complex 3-D plan, in place (SimplePofk/main.cpp:295-300 is the only user), and the calls around it */
#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
void *fftw_malloc(size_t n);
void fftw_free(void *p);
int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);

#ifdef __cplusplus
}
#endif

#endif
