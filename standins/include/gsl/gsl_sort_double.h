#ifndef MGP_SHIM_GSL_SORT_DOUBLE_H
#define MGP_SHIM_GSL_SORT_DOUBLE_H
#include <stddef.h>
void gsl_sort(double *data, size_t stride, size_t n);
#endif
