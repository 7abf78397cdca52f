#ifndef MGP_SHIM_GSL_SPLINE2D_H
#define MGP_SHIM_GSL_SPLINE2D_H
#include <gsl/gsl_interp2d.h>
typedef struct { size_t nx, ny; double *x, *y, *z, *zx, *zy, *zxy; } gsl_spline2d;
gsl_spline2d *gsl_spline2d_alloc(const gsl_interp2d_type *T, size_t nx, size_t ny);
int gsl_spline2d_init(gsl_spline2d *s, const double *xa, const double *ya, const double *za, size_t nx, size_t ny);
double gsl_spline2d_eval(const gsl_spline2d *s, double x, double y, gsl_interp_accel *xa, gsl_interp_accel *ya);
void gsl_spline2d_free(gsl_spline2d *s);
#endif
