#ifndef MGP_SHIM_GSL_ROOTS_H
#define MGP_SHIM_GSL_ROOTS_H
#include <gsl/gsl_math.h>
typedef struct { const char *name; } gsl_root_fsolver_type;
typedef struct {
  const gsl_root_fsolver_type *type;
  gsl_function *function;
  double root, x_lower, x_upper;
  double f_lower, f_upper;
} gsl_root_fsolver;
extern const gsl_root_fsolver_type *gsl_root_fsolver_brent;
gsl_root_fsolver *gsl_root_fsolver_alloc(const gsl_root_fsolver_type *T);
int gsl_root_fsolver_set(gsl_root_fsolver *s, gsl_function *f, double x_lower, double x_upper);
int gsl_root_fsolver_iterate(gsl_root_fsolver *s);
double gsl_root_fsolver_root(const gsl_root_fsolver *s);
double gsl_root_fsolver_x_lower(const gsl_root_fsolver *s);
double gsl_root_fsolver_x_upper(const gsl_root_fsolver *s);
void gsl_root_fsolver_free(gsl_root_fsolver *s);
int gsl_root_test_interval(double x_lower, double x_upper, double epsabs, double epsrel);
#endif
