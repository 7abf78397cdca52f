// Host emulation of the lightcone kernels (csrc/lightcone.cu): runs the very per-particle function the kernels call
// (csrc/lightcone.cuh), particle by particle on the CPU, counting pass then drift pass exactly as lightcone_drift
// sequences them.  Built with nvcc as a shared library and driven from tests/test_lightcone.py (no kernel launch).
#include <cstring>
#include <vector>

#include "lightcone.cuh"

using namespace mgp;

// scal[]: A, AFF, dyyy, da1, da2, dv1, dv2, sV[3], rc_old, rc_new, origin[3], box, boundary, lengthfac, vfac, usecola
extern "C" int lc_emul(long n, const float *pos, const float *vel, const float *D, const float *D2, const double *scal,
                       int ntab, const double *al, const double *t1, const double *t2, const double *t3, int nrep,
                       const int *rep, unsigned long long *count, float *rows_out, float *newpos) {
  lc::Params p;
  p.A = scal[0]; p.AFF = scal[1]; p.dyyy = scal[2]; p.da1 = scal[3]; p.da2 = scal[4]; p.dv1 = scal[5]; p.dv2 = scal[6];
  for (int a = 0; a < 3; a++) { p.sV[a] = scal[7 + a]; p.origin[a] = scal[12 + a]; }
  p.rc_old = scal[10]; p.rc_new = scal[11]; p.rc_old2 = p.rc_old * p.rc_old; p.rc_new2 = p.rc_new * p.rc_new;
  p.box = scal[15]; p.boxf = (float) scal[15]; p.boundary = scal[16]; p.lengthfac = scal[17]; p.vfac = scal[18]; p.usecola = scal[19];
  std::vector<double> c1(ntab), c2(ntab), c3(ntab);
  lc::spline_coeffs(al, t1, ntab, c1.data()); lc::spline_coeffs(al, t2, ntab, c2.data()); lc::spline_coeffs(al, t3, ntab, c3.data());
  p.ntab = ntab; p.al = al; p.y[0] = t1; p.y[1] = t2; p.y[2] = t3; p.c[0] = c1.data(); p.c[1] = c2.data(); p.c[2] = c3.data();
  p.nrep = nrep; p.rep = rep;
  std::vector<unsigned long long> cnt(nrep > 0 ? nrep : 1, 0ull), off(nrep > 0 ? nrep : 1, 0ull), cur(nrep > 0 ? nrep : 1, 0ull);
  auto load = [&](long i) {
    lc::Particle q;
    for (int a = 0; a < 3; a++) { q.pos[a] = pos[3 * i + a]; q.vel[a] = vel[3 * i + a]; q.d[a] = D[3 * i + a]; q.d2[a] = D2[3 * i + a]; }
    return q;
  };
  bool over = false;
  for (long i = 0; i < n; i++) {
    lc::Particle q = load(i);
    over |= lc::particle<false>(p, q, true, off.data(), nullptr, [&](int r, bool out) { return out ? cnt[r]++ : 0ull; });
  }
  unsigned long long total = 0;
  for (int r = 0; r < nrep; r++) { off[r] = total; total += cnt[r]; count[r] = cnt[r]; }
  if (over) return 1;
  for (long i = 0; i < n; i++) {
    lc::Particle q = load(i);
    lc::particle<true>(p, q, true, off.data(), rows_out, [&](int r, bool out) { return out ? cur[r]++ : 0ull; });
    for (int a = 0; a < 3; a++) newpos[3 * i + a] = q.pos[a];
  }
  for (int r = 0; r < nrep; r++) if (cur[r] != cnt[r]) return 2;
  return 0;
}
