/*
 * mini-GSL: the subset of the GNU Scientific Library API the reference calls
 * (SURVEY.md section 8(c)), written from the published algorithms.
 * TEST INFRASTRUCTURE ONLY -- it exists so the unmodified reference sources can be
 * compiled and run as the parity oracle in a container without GSL.
 *
 * What is pinned and what is not:
 *  - gsl_rng_ranlxd1: Luescher's RANLUX double-precision generator as shipped in
 *    GSL's rng/ranlxd.c (48-bit subtract-with-borrow, lags 12/5, luxury p = 202).
 *    Checked in tests/test_oracle_shim.py against the known-answer value in GSL's
 *    own rng/test.c (seed 1, 10000th draw == 0.465248546261094020, i.e.
 *    gsl_rng_get == 1998227290).
 *  - cspline / bicubic: same construction as GSL (natural cubic spline; bicubic
 *    Hermite patch with spline-estimated derivatives); not bit-identical.
 *  - qag / odeiv2-rk2 / brent / 2F1: functionally equivalent replacements that meet
 *    the tolerances the reference asks for; low-order bits differ from real GSL.
 *    These feed only the scalar growth factors / step integrals (cosmo.c), which
 *    cross the C-ABI as numbers, so GPU-vs-oracle parity is unaffected.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <gsl/gsl_errno.h>
#include <gsl/gsl_math.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_roots.h>
#include <gsl/gsl_spline.h>
#include <gsl/gsl_spline2d.h>
#include <gsl/gsl_integration.h>
#include <gsl/gsl_odeiv2.h>
#include <gsl/gsl_sf_hyperg.h>
#include <gsl/gsl_sort_double.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ rng */

static const gsl_rng_type ranlxd1_type = { "ranlxd1", 202 };
static const gsl_rng_type ranlxd2_type = { "ranlxd2", 397 };
const gsl_rng_type *gsl_rng_ranlxd1 = &ranlxd1_type;
const gsl_rng_type *gsl_rng_ranlxd2 = &ranlxd2_type;

static const double one_bit = 1.0 / 281474976710656.0;   /* 2^-48 */

static void ranlxd_increment_state(gsl_rng *s) {
  /* advance the recurrence x[ir] = x[jr] - x[ir] - carry (mod 1) by pr steps */
  unsigned int ir = s->ir, jr = s->jr;
  double carry = s->carry;
  for (unsigned int k = 0; k < s->pr; k++) {
    double y = s->xdbl[jr] - s->xdbl[ir] - carry;
    if (y < 0) { carry = one_bit; y += 1.0; } else { carry = 0.0; }
    s->xdbl[ir] = y;
    ir = (ir + 1) % 12;
    jr = (jr + 1) % 12;
  }
  s->ir = ir; s->ir_old = ir; s->jr = jr; s->carry = carry;
}

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T) {
  gsl_rng *r = (gsl_rng *) calloc(1, sizeof(gsl_rng));
  r->type = T;
  gsl_rng_set(r, 0);
  return r;
}

void gsl_rng_set(gsl_rng *r, unsigned long int s) {
  int xbit[31];
  if (s == 0) s = 1;
  long i = (long) (s & 0x7FFFFFFFUL);
  for (int k = 0; k < 31; k++) { xbit[k] = (int) (i % 2); i /= 2; }
  int ibit = 0, jbit = 18;
  for (int k = 0; k < 12; k++) {
    double x = 0;
    for (int l = 1; l <= 48; l++) {
      double y = (double) ((xbit[ibit] + 1) % 2);
      x += x + y;
      xbit[ibit] = (xbit[ibit] + xbit[jbit]) % 2;
      ibit = (ibit + 1) % 31;
      jbit = (jbit + 1) % 31;
    }
    r->xdbl[k] = one_bit * x;
  }
  r->carry = 0; r->ir = 11; r->jr = 7; r->ir_old = 0; r->pr = r->type->luxury;
}

double gsl_rng_uniform(gsl_rng *r) {
  r->ir = (r->ir + 1) % 12;
  if (r->ir == r->ir_old) ranlxd_increment_state(r);
  return r->xdbl[r->ir];
}

unsigned long int gsl_rng_get(gsl_rng *r) { return (unsigned long int) (gsl_rng_uniform(r) * 4294967296.0); }
void gsl_rng_free(gsl_rng *r) { free(r); }

/* ------------------------------------------------------------------ cubic spline */

static const gsl_interp_type cspline_type = { "cspline" };
const gsl_interp_type *gsl_interp_cspline = &cspline_type;
static const gsl_interp2d_type bicubic_type = { "bicubic" };
const gsl_interp2d_type *gsl_interp2d_bicubic = &bicubic_type;

gsl_interp_accel *gsl_interp_accel_alloc(void) { return (gsl_interp_accel *) calloc(1, sizeof(gsl_interp_accel)); }
void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }

/* second-derivative-like coefficients c[] of the natural cubic spline */
static void cspline_coeffs(const double *x, const double *y, size_t n, double *c) {
  c[0] = 0.0; c[n - 1] = 0.0;
  if (n < 3) return;
  size_t m = n - 2;
  double *diag = (double *) malloc(sizeof(double) * m), *off = (double *) malloc(sizeof(double) * m);
  double *g = (double *) malloc(sizeof(double) * m);
  for (size_t i = 0; i < m; i++) {
    double h_i = x[i + 1] - x[i], h_ip1 = x[i + 2] - x[i + 1];
    double ydiff_i = y[i + 1] - y[i], ydiff_ip1 = y[i + 2] - y[i + 1];
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (ydiff_ip1 / h_ip1 - ydiff_i / h_i);
  }
  /* symmetric tridiagonal solve (Thomas) */
  for (size_t i = 1; i < m; i++) {
    double w = off[i - 1] / diag[i - 1];
    diag[i] -= w * off[i - 1];
    g[i] -= w * g[i - 1];
  }
  c[m] = g[m - 1] / diag[m - 1];
  for (size_t i = m - 1; i-- > 0;) c[i + 1] = (g[i] - off[i] * c[i + 2]) / diag[i];
  free(diag); free(off); free(g);
}

static size_t bsearch_interval(const double *x, size_t n, double v) {
  size_t lo = 0, hi = n - 1;
  while (hi > lo + 1) {
    size_t mid = (lo + hi) / 2;
    if (x[mid] > v) hi = mid; else lo = mid;
  }
  return lo;
}

gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size) {
  (void) T;
  gsl_spline *s = (gsl_spline *) calloc(1, sizeof(gsl_spline));
  s->size = size;
  s->x = (double *) malloc(sizeof(double) * size);
  s->y = (double *) malloc(sizeof(double) * size);
  s->c = (double *) malloc(sizeof(double) * size);
  return s;
}
int gsl_spline_init(gsl_spline *s, const double *xa, const double *ya, size_t size) {
  memcpy(s->x, xa, sizeof(double) * size);
  memcpy(s->y, ya, sizeof(double) * size);
  cspline_coeffs(s->x, s->y, size, s->c);
  return GSL_SUCCESS;
}
static void cspline_bd(const gsl_spline *s, size_t i, double *b, double *d, double *dx) {
  *dx = s->x[i + 1] - s->x[i];
  double dy = s->y[i + 1] - s->y[i];
  *b = dy / *dx - *dx * (s->c[i + 1] + 2.0 * s->c[i]) / 3.0;
  *d = (s->c[i + 1] - s->c[i]) / (3.0 * *dx);
}
double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a) {
  (void) a;
  size_t i = bsearch_interval(s->x, s->size, x);
  double b, d, dx;
  cspline_bd(s, i, &b, &d, &dx);
  double delx = x - s->x[i];
  return s->y[i] + delx * (b + delx * (s->c[i] + delx * d));
}
double gsl_spline_eval_deriv(const gsl_spline *s, double x, gsl_interp_accel *a) {
  (void) a;
  size_t i = bsearch_interval(s->x, s->size, x);
  double b, d, dx;
  cspline_bd(s, i, &b, &d, &dx);
  double delx = x - s->x[i];
  return b + delx * (2.0 * s->c[i] + 3.0 * d * delx);
}
void gsl_spline_free(gsl_spline *s) { free(s->x); free(s->y); free(s->c); free(s); }

/* ------------------------------------------------------------------ bicubic */

gsl_spline2d *gsl_spline2d_alloc(const gsl_interp2d_type *T, size_t nx, size_t ny) {
  (void) T;
  gsl_spline2d *s = (gsl_spline2d *) calloc(1, sizeof(gsl_spline2d));
  s->nx = nx; s->ny = ny;
  s->x = (double *) malloc(sizeof(double) * nx);
  s->y = (double *) malloc(sizeof(double) * ny);
  s->z = (double *) malloc(sizeof(double) * nx * ny);
  s->zx = (double *) malloc(sizeof(double) * nx * ny);
  s->zy = (double *) malloc(sizeof(double) * nx * ny);
  s->zxy = (double *) malloc(sizeof(double) * nx * ny);
  return s;
}

int gsl_spline2d_init(gsl_spline2d *s, const double *xa, const double *ya, const double *za, size_t nx, size_t ny) {
  memcpy(s->x, xa, sizeof(double) * nx);
  memcpy(s->y, ya, sizeof(double) * ny);
  memcpy(s->z, za, sizeof(double) * nx * ny);
  gsl_spline *sp;
  double *tmp;
  /* d/dx along each row j (z index = j * nx + i) */
  sp = gsl_spline_alloc(gsl_interp_cspline, nx);
  tmp = (double *) malloc(sizeof(double) * (nx > ny ? nx : ny));
  for (size_t j = 0; j < ny; j++) {
    for (size_t i = 0; i < nx; i++) tmp[i] = s->z[j * nx + i];
    gsl_spline_init(sp, s->x, tmp, nx);
    for (size_t i = 0; i < nx; i++) s->zx[j * nx + i] = gsl_spline_eval_deriv(sp, s->x[i], NULL);
  }
  gsl_spline_free(sp);
  /* d/dy along each column i */
  sp = gsl_spline_alloc(gsl_interp_cspline, ny);
  for (size_t i = 0; i < nx; i++) {
    for (size_t j = 0; j < ny; j++) tmp[j] = s->z[j * nx + i];
    gsl_spline_init(sp, s->y, tmp, ny);
    for (size_t j = 0; j < ny; j++) s->zy[j * nx + i] = gsl_spline_eval_deriv(sp, s->y[j], NULL);
  }
  gsl_spline_free(sp);
  /* d2/dxdy: d/dx of zy along each row */
  sp = gsl_spline_alloc(gsl_interp_cspline, nx);
  for (size_t j = 0; j < ny; j++) {
    for (size_t i = 0; i < nx; i++) tmp[i] = s->zy[j * nx + i];
    gsl_spline_init(sp, s->x, tmp, nx);
    for (size_t i = 0; i < nx; i++) s->zxy[j * nx + i] = gsl_spline_eval_deriv(sp, s->x[i], NULL);
  }
  gsl_spline_free(sp);
  free(tmp);
  return GSL_SUCCESS;
}

double gsl_spline2d_eval(const gsl_spline2d *s, double x, double y, gsl_interp_accel *xa, gsl_interp_accel *ya) {
  (void) xa; (void) ya;
  const size_t nx = s->nx;
  size_t xi = bsearch_interval(s->x, s->nx, x), yi = bsearch_interval(s->y, s->ny, y);
  double xmin = s->x[xi], xmax = s->x[xi + 1], ymin = s->y[yi], ymax = s->y[yi + 1];
  double dx = xmax - xmin, dy = ymax - ymin;
  double t = (x - xmin) / dx, u = (y - ymin) / dy;
#define Z(a, ii, jj) (s->a[(yi + (jj)) * nx + xi + (ii)])
  /* Hermite basis */
  double h00t = (1 + 2 * t) * (1 - t) * (1 - t), h10t = t * (1 - t) * (1 - t);
  double h01t = t * t * (3 - 2 * t), h11t = t * t * (t - 1);
  double h00u = (1 + 2 * u) * (1 - u) * (1 - u), h10u = u * (1 - u) * (1 - u);
  double h01u = u * u * (3 - 2 * u), h11u = u * u * (u - 1);
  double v = 0;
  v += h00t * h00u * Z(z, 0, 0) + h01t * h00u * Z(z, 1, 0) + h00t * h01u * Z(z, 0, 1) + h01t * h01u * Z(z, 1, 1);
  v += dx * (h10t * h00u * Z(zx, 0, 0) + h11t * h00u * Z(zx, 1, 0) + h10t * h01u * Z(zx, 0, 1) + h11t * h01u * Z(zx, 1, 1));
  v += dy * (h00t * h10u * Z(zy, 0, 0) + h01t * h10u * Z(zy, 1, 0) + h00t * h11u * Z(zy, 0, 1) + h01t * h11u * Z(zy, 1, 1));
  v += dx * dy * (h10t * h10u * Z(zxy, 0, 0) + h11t * h10u * Z(zxy, 1, 0) + h10t * h11u * Z(zxy, 0, 1) + h11t * h11u * Z(zxy, 1, 1));
#undef Z
  return v;
}
void gsl_spline2d_free(gsl_spline2d *s) {
  free(s->x); free(s->y); free(s->z); free(s->zx); free(s->zy); free(s->zxy); free(s);
}

/* ------------------------------------------------------------------ quadrature */

#define NGL 16
static double gl_x[NGL], gl_w[NGL];
static int gl_ready = 0;
static void gl_init(void) {
  /* Gauss-Legendre nodes by Newton iteration on P_n */
  for (int i = 0; i < NGL; i++) {
    double x = cos(M_PI * (i + 0.75) / (NGL + 0.5)), pp = 0;
    for (int it = 0; it < 100; it++) {
      double p0 = 1.0, p1 = x;
      for (int k = 2; k <= NGL; k++) { double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
      pp = NGL * (x * p1 - p0) / (x * x - 1.0);
      double dx = p1 / pp;
      x -= dx;
      if (fabs(dx) < 1e-16) break;
    }
    gl_x[i] = x;
    gl_w[i] = 2.0 / ((1.0 - x * x) * pp * pp);
  }
  gl_ready = 1;
}
static double gl_panel(const gsl_function *f, double a, double b) {
  double c = 0.5 * (a + b), h = 0.5 * (b - a), s = 0;
  for (int i = 0; i < NGL; i++) s += gl_w[i] * GSL_FN_EVAL(f, c + h * gl_x[i]);
  return s * h;
}
static long gl_budget = 0;   /* panels one qag call may still split: GSL's `limit` on the number of subintervals, in this scheme's units */
static double gl_adapt(const gsl_function *f, double a, double b, double whole, double tol, int depth, double *err) {
  double m = 0.5 * (a + b);
  double l = gl_panel(f, a, m), r = gl_panel(f, m, b);
  double e = fabs(l + r - whole);
  /* converged, or at the rounding level of the panel (GSL's qag stops there too: its round-off test; a tolerance that
     keeps halving with depth would otherwise never be met next to an integrable singularity, e.g. the symmetron's
     dphi/dlna ~ 1/sqrt(a - a_ssb), cosmo.c:101-118) */
  if (e <= tol || e <= 4.0e-16 * (fabs(l) + fabs(r)) || depth >= 48 || --gl_budget < 0) { *err += e; return l + r; }
  return gl_adapt(f, a, m, l, 0.5 * tol, depth + 1, err) + gl_adapt(f, m, b, r, 0.5 * tol, depth + 1, err);
}

gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n) {
  gsl_integration_workspace *w = (gsl_integration_workspace *) malloc(sizeof(*w));
  w->limit = n;
  return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace *w) { free(w); }

int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w, double *result, double *abserr) {
  (void) key; (void) w;
  if (!gl_ready) gl_init();
  /* where the integrand is noise (cancellation next to a singular point) no panel converges: stop as GSL does at `limit` */
  gl_budget = 64L * (long) (limit ? limit : 1000);
  if (a == b) { *result = 0; *abserr = 0; return GSL_SUCCESS; }
  /* coarse estimate on 32 panels sets the scale; then converge far below the requested tolerance */
  const int np = 32;
  double est = 0, absest = 0;
  double part[32];
  for (int i = 0; i < np; i++) {
    part[i] = gl_panel(f, a + (b - a) * i / np, a + (b - a) * (i + 1) / np);
    est += part[i]; absest += fabs(part[i]);
  }
  double tol = 1e-4 * (epsabs > epsrel * fabs(est) ? epsabs : epsrel * fabs(est));
  double floor_tol = 1e-14 * absest;
  if (tol < floor_tol) tol = floor_tol;
  double sum = 0, err = 0;
  for (int i = 0; i < np; i++)
    sum += gl_adapt(f, a + (b - a) * i / np, a + (b - a) * (i + 1) / np, part[i], tol / np, 0, &err);
  *result = sum; *abserr = err;
  return GSL_SUCCESS;
}

/* ------------------------------------------------------------------ ODE: embedded RK 2(3) with standard step control */

static const gsl_odeiv2_step_type rk2_type = { "rk2" };
const gsl_odeiv2_step_type *gsl_odeiv2_step_rk2 = &rk2_type;

gsl_odeiv2_driver *gsl_odeiv2_driver_alloc_y_new(const gsl_odeiv2_system *sys, const gsl_odeiv2_step_type *T,
                                                 double hstart, double epsabs, double epsrel) {
  (void) T;
  gsl_odeiv2_driver *d = (gsl_odeiv2_driver *) calloc(1, sizeof(*d));
  size_t n = sys->dimension;
  d->sys = sys; d->h = hstart; d->epsabs = epsabs; d->epsrel = epsrel;
  d->k1 = (double *) malloc(sizeof(double) * n); d->k2 = (double *) malloc(sizeof(double) * n);
  d->k3 = (double *) malloc(sizeof(double) * n); d->ytmp = (double *) malloc(sizeof(double) * n);
  d->y0 = (double *) malloc(sizeof(double) * n); d->yerr = (double *) malloc(sizeof(double) * n);
  return d;
}
void gsl_odeiv2_driver_free(gsl_odeiv2_driver *d) {
  free(d->k1); free(d->k2); free(d->k3); free(d->ytmp); free(d->y0); free(d->yerr); free(d);
}

int gsl_odeiv2_driver_apply(gsl_odeiv2_driver *d, double *t, double t1, double y[]) {
  const gsl_odeiv2_system *sys = d->sys;
  const size_t n = sys->dimension;
  const double sgn = (t1 >= *t) ? 1.0 : -1.0;
  if (d->h * sgn < 0) d->h = -d->h;
  long nsteps = 0;
  while ((t1 - *t) * sgn > 0) {
    double h = d->h;
    int final_step = 0;
    if ((*t + h - t1) * sgn >= 0) { h = t1 - *t; final_step = 1; }
    memcpy(d->y0, y, sizeof(double) * n);
    for (;;) {
      if (sys->function(*t, d->y0, d->k1, sys->params) != GSL_SUCCESS) return GSL_FAILURE;
      for (size_t i = 0; i < n; i++) d->ytmp[i] = d->y0[i] + 0.5 * h * d->k1[i];
      if (sys->function(*t + 0.5 * h, d->ytmp, d->k2, sys->params) != GSL_SUCCESS) return GSL_FAILURE;
      for (size_t i = 0; i < n; i++) d->ytmp[i] = d->y0[i] + h * (-d->k1[i] + 2.0 * d->k2[i]);
      if (sys->function(*t + h, d->ytmp, d->k3, sys->params) != GSL_SUCCESS) return GSL_FAILURE;
      double rmax = 0;
      for (size_t i = 0; i < n; i++) {
        double ksum3 = (d->k1[i] + 4.0 * d->k2[i] + d->k3[i]) / 6.0;
        y[i] = d->y0[i] + h * ksum3;
        d->yerr[i] = h * (d->k2[i] - ksum3);
        double D0 = d->epsabs + d->epsrel * fabs(y[i]);
        double r = fabs(d->yerr[i]) / fabs(D0);
        if (r > rmax) rmax = r;
      }
      if (rmax > 1.1) {
        double r = 0.9 / pow(rmax, 1.0 / 2.0);
        if (r < 0.2) r = 0.2;
        h *= r;
        final_step = 0;
        if (++nsteps > 100000000L) return GSL_FAILURE;
        continue;                       /* reject and retry */
      }
      *t = final_step ? t1 : *t + h;
      if (rmax < 0.5) {
        double r = 0.9 / pow(rmax > 1e-300 ? rmax : 1e-300, 1.0 / 3.0);
        if (r > 5.0) r = 5.0;
        if (r < 1.0) r = 1.0;
        if (!final_step) d->h = h * r;
      } else if (!final_step) {
        d->h = h;
      }
      break;
    }
    if (++nsteps > 100000000L) return GSL_FAILURE;
  }
  return GSL_SUCCESS;
}

/* ------------------------------------------------------------------ root bracketing (interface of the brent solver) */

static const gsl_root_fsolver_type brent_type = { "brent" };
const gsl_root_fsolver_type *gsl_root_fsolver_brent = &brent_type;

gsl_root_fsolver *gsl_root_fsolver_alloc(const gsl_root_fsolver_type *T) {
  gsl_root_fsolver *s = (gsl_root_fsolver *) calloc(1, sizeof(*s));
  s->type = T;
  return s;
}
int gsl_root_fsolver_set(gsl_root_fsolver *s, gsl_function *f, double x_lower, double x_upper) {
  s->function = f; s->x_lower = x_lower; s->x_upper = x_upper;
  s->f_lower = GSL_FN_EVAL(f, x_lower); s->f_upper = GSL_FN_EVAL(f, x_upper);
  s->root = 0.5 * (x_lower + x_upper);
  return GSL_SUCCESS;
}
int gsl_root_fsolver_iterate(gsl_root_fsolver *s) {
  /* regula falsi step safeguarded by bisection (Illinois-style); keeps a valid bracket */
  double a = s->x_lower, b = s->x_upper, fa = s->f_lower, fb = s->f_upper;
  double x = (fa != fb) ? b - fb * (b - a) / (fb - fa) : 0.5 * (a + b);
  if (!(x > a && x < b)) x = 0.5 * (a + b);
  double fx = GSL_FN_EVAL(s->function, x);
  if (fx == 0.0) { s->x_lower = s->x_upper = s->root = x; s->f_lower = s->f_upper = 0; return GSL_SUCCESS; }
  if ((fx < 0) == (fa < 0)) { a = x; fa = fx; } else { b = x; fb = fx; }
  double m = 0.5 * (a + b), fm = GSL_FN_EVAL(s->function, m);
  if ((fm < 0) == (fa < 0)) { a = m; fa = fm; } else { b = m; fb = fm; }
  s->x_lower = a; s->x_upper = b; s->f_lower = fa; s->f_upper = fb;
  s->root = 0.5 * (a + b);
  return GSL_SUCCESS;
}
double gsl_root_fsolver_root(const gsl_root_fsolver *s) { return s->root; }
double gsl_root_fsolver_x_lower(const gsl_root_fsolver *s) { return s->x_lower; }
double gsl_root_fsolver_x_upper(const gsl_root_fsolver *s) { return s->x_upper; }
void gsl_root_fsolver_free(gsl_root_fsolver *s) { free(s); }
int gsl_root_test_interval(double x_lower, double x_upper, double epsabs, double epsrel) {
  double abs_lower = fabs(x_lower), abs_upper = fabs(x_upper), min_abs;
  if ((x_lower > 0 && x_upper > 0) || (x_lower < 0 && x_upper < 0)) min_abs = abs_lower < abs_upper ? abs_lower : abs_upper;
  else min_abs = 0;
  double tolerance = epsabs + epsrel * min_abs;
  return fabs(x_upper - x_lower) < tolerance ? GSL_SUCCESS : GSL_CONTINUE;
}

/* ------------------------------------------------------------------ 2F1 (|x| < 1 Gauss series; only reaches printouts) */

double gsl_sf_hyperg_2F1(double a, double b, double c, double x) {
  if (fabs(x) >= 1.0) {
    /* Pfaff transformation maps x < -1 into (0.5, 1): 2F1(a,b;c;x) = (1-x)^-a 2F1(a,c-b;c;x/(x-1)) */
    if (x < 0) return pow(1.0 - x, -a) * gsl_sf_hyperg_2F1(a, c - b, c, x / (x - 1.0));
    return NAN;
  }
  double term = 1.0, sum = 1.0;
  for (int n = 0; n < 2000000; n++) {
    term *= (a + n) * (b + n) / ((c + n) * (n + 1.0)) * x;
    sum += term;
    if (fabs(term) < 1e-17 * fabs(sum)) break;
  }
  return sum;
}

/* ------------------------------------------------------------------ sort */

static int cmp_double(const void *a, const void *b) {
  double x = *(const double *) a, y = *(const double *) b;
  return (x > y) - (x < y);
}
void gsl_sort(double *data, size_t stride, size_t n) {
  if (stride == 1) { qsort(data, n, sizeof(double), cmp_double); return; }
  double *tmp = (double *) malloc(sizeof(double) * n);
  for (size_t i = 0; i < n; i++) tmp[i] = data[i * stride];
  qsort(tmp, n, sizeof(double), cmp_double);
  for (size_t i = 0; i < n; i++) data[i * stride] = tmp[i];
  free(tmp);
}

/* ------------------------------------------------------------------ symmetric eigen-systems (mm_fof.c:548-556) */
#include <gsl/gsl_eigen.h>

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2) {
  gsl_matrix *m = (gsl_matrix *) malloc(sizeof(gsl_matrix));
  m->size1 = n1; m->size2 = n2; m->data = (double *) calloc(n1 * n2, sizeof(double));
  return m;
}
void gsl_matrix_free(gsl_matrix *m) { if (m) { free(m->data); free(m); } }
void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x) { m->data[i * m->size2 + j] = x; }
double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->size2 + j]; }
gsl_vector *gsl_vector_alloc(size_t n) {
  gsl_vector *v = (gsl_vector *) malloc(sizeof(gsl_vector));
  v->size = n; v->data = (double *) calloc(n, sizeof(double));
  return v;
}
void gsl_vector_free(gsl_vector *v) { if (v) { free(v->data); free(v); } }
double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i]; }
gsl_eigen_symmv_workspace *gsl_eigen_symmv_alloc(size_t n) {
  gsl_eigen_symmv_workspace *w = (gsl_eigen_symmv_workspace *) malloc(sizeof(gsl_eigen_symmv_workspace));
  w->size = n;
  return w;
}
void gsl_eigen_symmv_free(gsl_eigen_symmv_workspace *w) { free(w); }

/* cyclic Jacobi rotations (GSL tridiagonalises and runs implicit QR; eigenvalues agree to rounding, eigenvectors up to
 * sign and, for degenerate eigenvalues, up to a rotation of the eigen-space: what a comparison may rely on) */
int gsl_eigen_symmv(gsl_matrix *A, gsl_vector *eval, gsl_matrix *evec, gsl_eigen_symmv_workspace *w) {
  (void) w;
  const size_t n = A->size1;
  double *a = A->data, *v = evec->data;
  for (size_t i = 0; i < n; i++) for (size_t j = 0; j < n; j++) v[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = 0.0;
    for (size_t p = 0; p < n; p++) for (size_t q = p + 1; q < n; q++) off += a[p * n + q] * a[p * n + q];
    if (off == 0.0) break;
    for (size_t p = 0; p < n; p++)
      for (size_t q = p + 1; q < n; q++) {
        const double apq = a[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (size_t k = 0; k < n; k++) {           /* A <- A J */
          const double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq; a[k * n + q] = s * akp + c * akq;
        }
        for (size_t k = 0; k < n; k++) {           /* A <- J^T A */
          const double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk; a[q * n + k] = s * apk + c * aqk;
        }
        for (size_t k = 0; k < n; k++) {           /* V <- V J */
          const double vkp = v[k * n + p], vkq = v[k * n + q];
          v[k * n + p] = c * vkp - s * vkq; v[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  for (size_t i = 0; i < n; i++) eval->data[i] = a[i * n + i];
  return GSL_SUCCESS;
}
