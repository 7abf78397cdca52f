#ifndef MGP_SHIM_GSL_SF_HYPERG_H
#define MGP_SHIM_GSL_SF_HYPERG_H
double gsl_sf_hyperg_2F1(double a, double b, double c, double x);
#endif
