#ifndef MGP_SHIM_GSL_INTEGRATION_H
#define MGP_SHIM_GSL_INTEGRATION_H
#include <stddef.h>
#include <gsl/gsl_math.h>
typedef struct { size_t limit; } gsl_integration_workspace;
enum { GSL_INTEG_GAUSS15 = 1, GSL_INTEG_GAUSS21 = 2, GSL_INTEG_GAUSS31 = 3,
       GSL_INTEG_GAUSS41 = 4, GSL_INTEG_GAUSS51 = 5, GSL_INTEG_GAUSS61 = 6 };
gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w, double *result, double *abserr);
#endif
