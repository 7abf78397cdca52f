#ifndef MGP_SHIM_GSL_SPLINE_H
#define MGP_SHIM_GSL_SPLINE_H
#include <gsl/gsl_interp.h>
typedef struct { size_t size; double *x, *y, *c; } gsl_spline;
gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *s, const double *xa, const double *ya, size_t size);
double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a);
double gsl_spline_eval_deriv(const gsl_spline *s, double x, gsl_interp_accel *a);
void gsl_spline_free(gsl_spline *s);
#endif
