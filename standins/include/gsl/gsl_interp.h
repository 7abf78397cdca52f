#ifndef MGP_SHIM_GSL_INTERP_H
#define MGP_SHIM_GSL_INTERP_H
#include <stddef.h>
typedef struct { size_t cache; } gsl_interp_accel;
typedef struct { const char *name; } gsl_interp_type;
extern const gsl_interp_type *gsl_interp_cspline;
gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
#endif
