#ifndef MGP_SHIM_GSL_INTERP2D_H
#define MGP_SHIM_GSL_INTERP2D_H
#include <gsl/gsl_interp.h>
typedef struct { const char *name; } gsl_interp2d_type;
extern const gsl_interp2d_type *gsl_interp2d_bicubic;
#endif
