/* Serial stand-in for <fftw3-mpi.h>; see fftw3.h in this directory. */
#ifndef MGP_SHIM_FFTW3_MPI_H
#define MGP_SHIM_FFTW3_MPI_H

#include <mpi.h>
#include <fftw3.h>

void fftw_mpi_init(void);
void fftw_mpi_cleanup(void);
ptrdiff_t fftw_mpi_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, MPI_Comm comm,
                                 ptrdiff_t *local_n0, ptrdiff_t *local_0_start);
fftw_plan fftw_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double *in,
                                   fftw_complex *out, MPI_Comm comm, unsigned flags);
fftw_plan fftw_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex *in,
                                   double *out, MPI_Comm comm, unsigned flags);

void fftwf_mpi_init(void);
void fftwf_mpi_cleanup(void);
ptrdiff_t fftwf_mpi_local_size_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, MPI_Comm comm,
                                  ptrdiff_t *local_n0, ptrdiff_t *local_0_start);
fftwf_plan fftwf_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, float *in,
                                     fftwf_complex *out, MPI_Comm comm, unsigned flags);
fftwf_plan fftwf_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftwf_complex *in,
                                     float *out, MPI_Comm comm, unsigned flags);

#endif
