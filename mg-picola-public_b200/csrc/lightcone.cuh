// Per-particle arithmetic of Drift_Lightcone (lightcone.c:392-471), shared by the kernels of lightcone.cu and by the
// host emulation under tests/host/ (the same functions run particle by particle on the CPU against the reference's rows).
// Every operation is spelled out in the association order the C compiler gives the reference's expressions, with
// round-to-nearest intrinsics on the device so that no multiply-add is contracted: the rows are the reference's floats.
#pragma once

#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define LC_HD __host__ __device__ __forceinline__
#else
#define LC_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define LC_MUL(a, b) __dmul_rn((a), (b))
#define LC_ADD(a, b) __dadd_rn((a), (b))
#define LC_SUB(a, b) __dsub_rn((a), (b))
#define LC_DIV(a, b) __ddiv_rn((a), (b))
#define LC_SQRT(a) __dsqrt_rn((a))
#else
#include <cmath>
#define LC_MUL(a, b) ((a) * (b))
#define LC_ADD(a, b) ((a) + (b))
#define LC_SUB(a, b) ((a) - (b))
#define LC_DIV(a, b) ((a) / (b))
#define LC_SQRT(a) std::sqrt((a))
#endif

namespace mgp {
namespace lc {

// the scalars of one Drift_Lightcone call (lightcone.c:281-317) plus the exit-time splines (326-357)
struct Params {
  double A, AFF, dyyy, da1, da2, dv1, dv2;
  double sV[3];
  double rc_old, rc_new, rc_old2, rc_new2;
  double origin[3];
  double box, boundary, lengthfac, vfac, usecola;
  float boxf;
  int ntab;
  const double *al;        // [ntab] nodes
  const double *y[3];      // [ntab] da1_tab, da2_tab, dyyy_tab
  const double *c[3];      // [ntab] natural-spline coefficients of the three tables
  int nrep;
  const int *rep;          // [nrep][3]
};

struct Particle {
  float pos[3], vel[3], d[3], d2[3];
};

// auxPM.c:649-655 in float (MEMORY_MODE)
LC_HD float wrap_f(float x, float box) {
  while (x >= box) x -= box;
  while (x < 0) x += box;
  if (x == box) x = 0.0f;
  return x;
}

// Delta_Pos of lightcone.c:398-400; true when a component exceeds `boundary` (403: the reference aborts)
LC_HD bool delta_pos(const Params &p, const Particle &q, double dp[3]) {
  bool over = false;
  for (int a = 0; a < 3; a++) {
    dp[a] = LC_ADD(LC_MUL(LC_SUB((double) q.vel[a], p.sV[a]), p.dyyy),
                   LC_MUL(p.usecola, LC_ADD(LC_MUL((double) q.d[a], p.da1), LC_MUL((double) q.d2[a], p.da2))));
    over = over || dp[a] > p.boundary;
  }
  return over;
}

// lightcone.c:418-434: inside the old lightcone at the start of the step and outside the new one at its end
LC_HD bool leaves(const Params &p, const Particle &q, const double dp[3], int i, int j, int k, double &r_old2, double &r_new2) {
  double X = LC_ADD(LC_SUB((double) q.pos[0], p.origin[0]), LC_MUL((double) i, p.box));
  double Y = LC_ADD(LC_SUB((double) q.pos[1], p.origin[1]), LC_MUL((double) j, p.box));
  double Z = LC_ADD(LC_SUB((double) q.pos[2], p.origin[2]), LC_MUL((double) k, p.box));
  r_old2 = LC_ADD(LC_ADD(LC_MUL(X, X), LC_MUL(Y, Y)), LC_MUL(Z, Z));
  if (!(r_old2 <= p.rc_old2)) return false;
  X = LC_ADD(X, dp[0]); Y = LC_ADD(Y, dp[1]); Z = LC_ADD(Z, dp[2]);
  r_new2 = LC_ADD(LC_ADD(LC_MUL(X, X), LC_MUL(Y, Y)), LC_MUL(Z, Z));
  return r_new2 > p.rc_new2;
}

// gsl_interp_cspline's evaluation (natural cubic spline with coefficients c): interval by bisection, then
//   b = dy / dx - dx (c[i+1] + 2 c[i]) / 3,  d = (c[i+1] - c[i]) / (3 dx),  y = y[i] + t (b + t (c[i] + t d))
LC_HD double spline(const double *x, const double *y, const double *c, int n, double v) {
  int lo = 0, hi = n - 1;
  while (hi > lo + 1) {
    const int mid = (lo + hi) / 2;
    if (x[mid] > v) hi = mid; else lo = mid;
  }
  const double dx = LC_SUB(x[lo + 1], x[lo]), dy = LC_SUB(y[lo + 1], y[lo]);
  const double b = LC_SUB(LC_DIV(dy, dx), LC_DIV(LC_MUL(dx, LC_ADD(c[lo + 1], LC_MUL(2.0, c[lo]))), 3.0));
  const double d = LC_DIV(LC_SUB(c[lo + 1], c[lo]), LC_MUL(3.0, dx));
  const double t = LC_SUB(v, x[lo]);
  return LC_ADD(y[lo], LC_MUL(t, LC_ADD(b, LC_MUL(t, LC_ADD(c[lo], LC_MUL(t, d))))));
}

// lightcone.c:436-453: exit time from the two radii, growth increments at that time, the six floats of the row
LC_HD void exit_row(const Params &p, const Particle &q, int i, int j, int k, double r_old2, double r_new2, float row[6]) {
  const double r_old = LC_SQRT(r_old2), r_new = LC_SQRT(r_new2);
  const double AL = LC_ADD(p.A, LC_MUL(LC_SUB(p.AFF, p.A),
                                      LC_DIV(LC_SUB(p.rc_old, r_old), LC_SUB(LC_SUB(r_new, r_old), LC_SUB(p.rc_new, p.rc_old)))));
  const double da1 = spline(p.al, p.y[0], p.c[0], p.ntab, AL);
  const double da2 = spline(p.al, p.y[1], p.c[1], p.ntab, AL);
  const double dyyy = spline(p.al, p.y[2], p.c[2], p.ntab, AL);
  const int ijk[3] = {i, j, k};
  for (int a = 0; a < 3; a++) {
    const double vs = LC_SUB((double) q.vel[a], p.sV[a]);
    const double x = LC_ADD(LC_ADD(LC_ADD((double) q.pos[a], LC_MUL(vs, dyyy)),
                                   LC_MUL(p.usecola, LC_ADD(LC_MUL((double) q.d[a], da1), LC_MUL((double) q.d2[a], da2)))),
                            LC_MUL((double) ijk[a], p.box));
    row[a] = (float) LC_MUL(p.lengthfac, x);
    const double v = LC_ADD(vs, LC_MUL(LC_ADD(LC_MUL((double) q.d[a], p.dv1), LC_MUL((double) q.d2[a], p.dv2)), p.usecola));
    row[3 + a] = (float) LC_MUL(p.vfac, v);
  }
}

// lightcone.c:468-470
LC_HD void advance(const Params &p, Particle &q, const double dp[3]) {
  for (int a = 0; a < 3; a++) q.pos[a] = wrap_f((float) LC_ADD((double) q.pos[a], dp[a]), p.boxf);
}

// One particle of the loop.  WRITE = false: count the images that leave; WRITE = true: also form their rows at
// rows[(offset[r] + slot) * 6] and advance the particle.  `bump(r, out)` is called for EVERY replicate (on the device by all
// lanes of the warp together: one atomic per warp and replicate, the lanes' slots from a ballot; a plain increment in the
// host emulation) and returns the slot of the new row of replicate r when `out` is set.  `valid` = false: a lane beyond
// the last particle that only takes part in the ballots.  Returns true if Delta_Pos exceeds the boundary.
template <bool WRITE, class Bump>
LC_HD bool particle(const Params &p, Particle &q, bool valid, const unsigned long long *offset, float *rows, Bump &&bump) {
  double dp[3];
  const bool over = delta_pos(p, q, dp) && valid;
  for (int r = 0; r < p.nrep; r++) {
    const int i = p.rep[3 * r], j = p.rep[3 * r + 1], k = p.rep[3 * r + 2];
    double ro2 = 0.0, rn2 = 0.0;
    const bool out = valid && leaves(p, q, dp, i, j, k, ro2, rn2);
    const unsigned long long slot = bump(r, out);
    if (WRITE && out) {
      float row[6];
      exit_row(p, q, i, j, k, ro2, rn2, row);
      float *o = rows + (size_t) (offset[r] + slot) * 6;
      for (int a = 0; a < 6; a++) o[a] = row[a];
    }
  }
  if (WRITE && valid) advance(p, q, dp);
  return over;
}

// natural cubic spline through (x, y): the coefficient array gsl_interp_cspline keeps (c[0] = c[n-1] = 0, the interior
// from the symmetric tridiagonal system  h_i c_i + 2 (h_i + h_{i+1}) c_{i+1} + h_{i+1} c_{i+2} = 3 (dy_{i+1} / h_{i+1} - dy_i / h_i))
inline void spline_coeffs(const double *x, const double *y, int n, double *c) {
  c[0] = 0.0; c[n - 1] = 0.0;
  if (n < 3) return;
  const int m = n - 2;
  double *diag = new double[m], *off = new double[m], *g = new double[m];
  for (int i = 0; i < m; i++) {
    const double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
    off[i] = h1;
    diag[i] = 2.0 * (h1 + h0);
    g[i] = 3.0 * ((y[i + 2] - y[i + 1]) / h1 - (y[i + 1] - y[i]) / h0);
  }
  for (int i = 1; i < m; i++) {
    const double w = off[i - 1] / diag[i - 1];
    diag[i] -= w * off[i - 1];
    g[i] -= w * g[i - 1];
  }
  c[m] = g[m - 1] / diag[m - 1];
  for (int i = m - 2; i >= 0; i--) c[i + 1] = (g[i] - off[i] * c[i + 2]) / diag[i];
  delete[] diag; delete[] off; delete[] g;
}

}  // namespace lc
}  // namespace mgp
