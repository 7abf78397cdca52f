#ifndef MGP_SHIM_GSL_ODEIV2_H
#define MGP_SHIM_GSL_ODEIV2_H
#include <stddef.h>
typedef struct {
  int (*function)(double t, const double y[], double dydt[], void *params);
  int (*jacobian)(double t, const double y[], double *dfdy, double dfdt[], void *params);
  size_t dimension;
  void *params;
} gsl_odeiv2_system;
typedef struct { const char *name; } gsl_odeiv2_step_type;
extern const gsl_odeiv2_step_type *gsl_odeiv2_step_rk2;
typedef struct {
  const gsl_odeiv2_system *sys;
  double h, epsabs, epsrel;
  double *k1, *k2, *k3, *ytmp, *y0, *yerr;
} gsl_odeiv2_driver;
gsl_odeiv2_driver *gsl_odeiv2_driver_alloc_y_new(const gsl_odeiv2_system *sys, const gsl_odeiv2_step_type *T,
                                                 double hstart, double epsabs, double epsrel);
int gsl_odeiv2_driver_apply(gsl_odeiv2_driver *d, double *t, double t1, double y[]);
void gsl_odeiv2_driver_free(gsl_odeiv2_driver *d);
#endif
