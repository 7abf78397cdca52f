/* mini-GSL stand-in: only gsl_rng_ranlxd1 (the generator 2LPT.c:259 uses). */
#ifndef MGP_SHIM_GSL_RNG_H
#define MGP_SHIM_GSL_RNG_H
typedef struct { const char *name; unsigned int luxury; } gsl_rng_type;
typedef struct {
  const gsl_rng_type *type;
  double xdbl[12];
  double carry;
  unsigned int ir, jr, ir_old, pr;
} gsl_rng;
extern const gsl_rng_type *gsl_rng_ranlxd1;
extern const gsl_rng_type *gsl_rng_ranlxd2;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_set(gsl_rng *r, unsigned long int seed);
double gsl_rng_uniform(gsl_rng *r);
unsigned long int gsl_rng_get(gsl_rng *r);
void gsl_rng_free(gsl_rng *r);
#endif
