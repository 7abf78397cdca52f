#ifndef MGP_SHIM_GSL_EIGEN_H
#define MGP_SHIM_GSL_EIGEN_H
/* stand-in for the part of <gsl/gsl_eigen.h> (with gsl_matrix / gsl_vector) that mm_fof.c:429-432, 548-556 uses:
 * real symmetric eigen-decomposition of the 3 x 3 inertia tensor of a halo.  TEST INFRASTRUCTURE ONLY. */
#include <stddef.h>
typedef struct { size_t size1, size2; double *data; } gsl_matrix;
typedef struct { size_t size; double *data; } gsl_vector;
typedef struct { size_t size; } gsl_eigen_symmv_workspace;
gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);
void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x);
double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j);
gsl_vector *gsl_vector_alloc(size_t n);
void gsl_vector_free(gsl_vector *v);
double gsl_vector_get(const gsl_vector *v, size_t i);
gsl_eigen_symmv_workspace *gsl_eigen_symmv_alloc(size_t n);
void gsl_eigen_symmv_free(gsl_eigen_symmv_workspace *w);
/* eigenvalues (unordered) in eval, orthonormal eigenvectors in the COLUMNS of evec; A is destroyed */
int gsl_eigen_symmv(gsl_matrix *A, gsl_vector *eval, gsl_matrix *evec, gsl_eigen_symmv_workspace *w);
#endif
