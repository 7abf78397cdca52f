/*
 * CPU FFT behind the FFTW / FFTW-MPI stand-in headers.  TEST INFRASTRUCTURE ONLY.
 * Serial slab layout: local_n0 = n0, local_0_start = 0 (one task owns the whole box),
 * which is what fftw_mpi_local_size_3d returns with a single MPI rank.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fftw3-mpi.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define R double
#define PREFIX mgpd_
#include "shim_fft_impl.h"
#undef R
#undef PREFIX

#define R float
#define PREFIX mgpf_
#include "shim_fft_impl.h"
#undef R
#undef PREFIX

struct mgp_shim_plan { int kind; int is_float; int n0, n1, n2; void *data; };

static struct mgp_shim_plan *mkplan(int kind, int is_float, ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2,
                                    void *in, void *out) {
  if (in != out) { fprintf(stderr, "[fftw shim] only in-place transforms are supported\n"); exit(1); }
  struct mgp_shim_plan *p = (struct mgp_shim_plan *) malloc(sizeof(*p));
  p->kind = kind; p->is_float = is_float; p->n0 = (int) n0; p->n1 = (int) n1; p->n2 = (int) n2; p->data = in;
  return p;
}

void fftw_mpi_init(void) {}
void fftw_mpi_cleanup(void) {}
void fftwf_mpi_init(void) {}
void fftwf_mpi_cleanup(void) {}

ptrdiff_t fftw_mpi_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, MPI_Comm comm,
                                 ptrdiff_t *local_n0, ptrdiff_t *local_0_start) {
  (void) comm; *local_n0 = n0; *local_0_start = 0; return n0 * n1 * n2;
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

/* complex 3-D transform in place, row-major [n0][n1][n2]: z rows, then y and x columns */
static void c2c_3d(int n0, int n1, int n2, double *data, int sign) {
  mgpd_plan1d *pz = mgpd_plan1d_new(n2), *py = mgpd_plan1d_new(n1), *px = mgpd_plan1d_new(n0);
  int nmax = n0 > n1 ? n0 : n1; if (n2 > nmax) nmax = n2;
  mgpd_cpx *buf = (mgpd_cpx *) malloc(sizeof(mgpd_cpx) * (size_t) nmax * 8);
  mgpd_cpx *cd = (mgpd_cpx *) data;
  for (size_t r = 0; r < (size_t) n0 * n1; r++) mgpd_fft1d(pz, cd + r * n2, sign);
  for (int x = 0; x < n0; x++) mgpd_fft_cols(py, cd + (size_t) x * n1 * n2, (size_t) n2, n2, sign, buf);
  mgpd_fft_cols(px, cd, (size_t) n1 * n2, n1 * n2, sign, buf);
  free(buf);
  mgpd_plan1d_free(pz); mgpd_plan1d_free(py); mgpd_plan1d_free(px);
}

fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex *in, fftw_complex *out, int sign, unsigned flags) {
  (void) flags;
  return mkplan(sign < 0 ? 2 : 3, 0, n0, n1, n2, in, out);
}
void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int nthreads) { (void) nthreads; }

void fftw_execute(const fftw_plan p) {
  if (p->kind >= 2) { c2c_3d(p->n0, p->n1, p->n2, (double *) p->data, p->kind == 2 ? -1 : +1); return; }
  if (p->is_float) {
    if (p->kind == 0) mgpf_r2c_3d(p->n0, p->n1, p->n2, (float *) p->data);
    else mgpf_c2r_3d(p->n0, p->n1, p->n2, (float *) p->data);
  } else {
    if (p->kind == 0) mgpd_r2c_3d(p->n0, p->n1, p->n2, (double *) p->data);
    else mgpd_c2r_3d(p->n0, p->n1, p->n2, (double *) p->data);
  }
}
void fftwf_execute(const fftwf_plan p) { fftw_execute(p); }
void fftw_destroy_plan(fftw_plan p) { free(p); }
void fftwf_destroy_plan(fftwf_plan p) { free(p); }

/* entry points for tests (ctypes): in-place padded transforms */
void mgp_shim_r2c_3d_f64(int n0, int n1, int n2, double *data) { mgpd_r2c_3d(n0, n1, n2, data); }
void mgp_shim_c2r_3d_f64(int n0, int n1, int n2, double *data) { mgpd_c2r_3d(n0, n1, n2, data); }
void mgp_shim_r2c_3d_f32(int n0, int n1, int n2, float *data) { mgpf_r2c_3d(n0, n1, n2, data); }
void mgp_shim_c2r_3d_f32(int n0, int n1, int n2, float *data) { mgpf_c2r_3d(n0, n1, n2, data); }
void mgp_shim_c2c_3d_f64(int n0, int n1, int n2, double *data, int sign) { c2c_3d(n0, n1, n2, data, sign); }
