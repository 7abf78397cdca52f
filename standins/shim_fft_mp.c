/*
 * FFTW-MPI stand-in for the multi-process MPI stand-in (shim_mpi_mp.c): x-slabs as fftw_mpi_local_size_3d hands
 * them out (block = ceil(n0 / ranks), 2LPT.c:50), in-place r2c / c2r with non-transposed output.
 * TEST INFRASTRUCTURE ONLY.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fftw3-mpi.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

int mgp_mp_rank(void);
int mgp_mp_size(void);
void mgp_mp_barrier(void);
void *mgp_mp_scratch(size_t bytes);

#define R double
#define PREFIX mgpd_
#include "shim_fft_impl.h"
#include "shim_fft_mp_impl.h"
#undef R
#undef PREFIX

#define R float
#define PREFIX mgpf_
#include "shim_fft_impl.h"
#include "shim_fft_mp_impl.h"
#undef R
#undef PREFIX

struct mgp_shim_plan { int kind; int is_float; int n0, n1, n2, local_n0, local_start; void *data; };

static void slab_of(ptrdiff_t n0, ptrdiff_t *local_n0, ptrdiff_t *local_0_start) {
  const ptrdiff_t P = mgp_mp_size(), me = mgp_mp_rank();
  const ptrdiff_t block = (n0 + P - 1) / P;
  ptrdiff_t s = me * block, n = block;
  if (s >= n0) { s = n0; n = 0; }
  else if (s + n > n0) n = n0 - s;
  *local_n0 = n; *local_0_start = s;
}

static struct mgp_shim_plan *mkplan(int kind, int is_float, ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2,
                                    void *in, void *out) {
  if (in != out) { fprintf(stderr, "[fftw shim] only in-place transforms are supported\n"); exit(1); }
  if (n1 % 2) { fprintf(stderr, "[fftw shim] n1 must be even\n"); exit(1); }
  struct mgp_shim_plan *p = (struct mgp_shim_plan *) malloc(sizeof(*p));
  ptrdiff_t ln, ls;
  slab_of(n0, &ln, &ls);
  p->kind = kind; p->is_float = is_float; p->n0 = (int) n0; p->n1 = (int) n1; p->n2 = (int) n2;
  p->local_n0 = (int) ln; p->local_start = (int) ls; p->data = in;
  return p;
}

void fftw_mpi_init(void) {}
void fftw_mpi_cleanup(void) {}
void fftwf_mpi_init(void) {}
void fftwf_mpi_cleanup(void) {}

ptrdiff_t fftw_mpi_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, MPI_Comm comm,
                                 ptrdiff_t *local_n0, ptrdiff_t *local_0_start) {
  (void) comm;
  slab_of(n0, local_n0, local_0_start);
  return *local_n0 * n1 * n2;
}
ptrdiff_t fftwf_mpi_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, MPI_Comm comm,
                                  ptrdiff_t *local_n0, ptrdiff_t *local_0_start) {
  return fftw_mpi_local_size_3d(n0, n1, n2, comm, local_n0, local_0_start);
}
fftw_plan fftw_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double *in, fftw_complex *out, MPI_Comm c, unsigned f) {
  (void) c; (void) f; return mkplan(0, 0, n0, n1, n2, in, out);
}
fftw_plan fftw_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex *in, double *out, MPI_Comm c, unsigned f) {
  (void) c; (void) f; return mkplan(1, 0, n0, n1, n2, in, out);
}
fftwf_plan fftwf_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, float *in, fftwf_complex *out, MPI_Comm c, unsigned f) {
  (void) c; (void) f; return mkplan(0, 1, n0, n1, n2, in, out);
}
fftwf_plan fftwf_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftwf_complex *in, float *out, MPI_Comm c, unsigned f) {
  (void) c; (void) f; return mkplan(1, 1, n0, n1, n2, in, out);
}

void fftw_execute(const fftw_plan p) {
  if (p->is_float) {
    if (p->kind == 0) mgpf_r2c_3d_mp(p->n0, p->n1, p->n2, p->local_n0, p->local_start, (float *) p->data);
    else mgpf_c2r_3d_mp(p->n0, p->n1, p->n2, p->local_n0, p->local_start, (float *) p->data);
  } else {
    if (p->kind == 0) mgpd_r2c_3d_mp(p->n0, p->n1, p->n2, p->local_n0, p->local_start, (double *) p->data);
    else mgpd_c2r_3d_mp(p->n0, p->n1, p->n2, p->local_n0, p->local_start, (double *) p->data);
  }
}
void fftwf_execute(const fftwf_plan p) { fftw_execute(p); }
void fftw_destroy_plan(fftw_plan p) { free(p); }
void fftwf_destroy_plan(fftwf_plan p) { free(p); }
