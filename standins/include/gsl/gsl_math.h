#ifndef MGP_SHIM_GSL_MATH_H
#define MGP_SHIM_GSL_MATH_H
#include <math.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846264338328 /* real gsl_math.h defines it too (-std=c99 hides libm's) */
#endif
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
#endif
