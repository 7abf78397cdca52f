/* mini-GSL stand-in (TEST INFRASTRUCTURE ONLY, see standins/README.md) */
#ifndef MGP_SHIM_GSL_ERRNO_H
#define MGP_SHIM_GSL_ERRNO_H
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_EDOM = 1, GSL_EMAXITER = 11 };
#endif
