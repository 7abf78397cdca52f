// Per-particle / per-pair / per-halo arithmetic of the FoF halo finder (MatchMaker: mm_main.c:266-338, mm_fof.c:81-186,
// 468-611), shared by the kernels of fof.cu and by the host emulation under tests/host/ (the same functions run on the CPU,
// for several emulated tasks, against the reference's catalogue).  Float and double operations are spelled out in the
// association order the C compiler gives the reference's expressions, with round-to-nearest intrinsics on the device so
// that nothing is contracted into a multiply-add: group membership and halo properties are the reference's to the bit.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "mgpicola.h"

#if defined(__CUDACC__)
#define FOF_HD __host__ __device__ __forceinline__
#else
#define FOF_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define FD_MUL(a, b) __dmul_rn((a), (b))
#define FD_ADD(a, b) __dadd_rn((a), (b))
#define FD_SUB(a, b) __dsub_rn((a), (b))
#define FD_DIV(a, b) __ddiv_rn((a), (b))
#define FD_SQRT(a) __dsqrt_rn((a))
#define FF_MUL(a, b) __fmul_rn((a), (b))
#define FF_ADD(a, b) __fadd_rn((a), (b))
#define FF_SUB(a, b) __fsub_rn((a), (b))
#else
#include <cmath>
#define FD_MUL(a, b) ((a) * (b))
#define FD_ADD(a, b) ((a) + (b))
#define FD_SUB(a, b) ((a) - (b))
#define FD_DIV(a, b) ((a) / (b))
#define FD_SQRT(a) std::sqrt((a))
#define FF_MUL(a, b) ((float) ((float) (a) * (float) (b)))
#define FF_ADD(a, b) ((float) ((float) (a) + (float) (b)))
#define FF_SUB(a, b) ((float) ((float) (a) - (float) (b)))
#endif

namespace mgp {
namespace fof {

// what init_fof (mm_fof.c:81-110) and MatchMaker (mm_main.c:219-240) derive once per call
struct Geometry {
  float edge;          // Param.x_offset: left edge of this task's slab
  float dx_domain;     // width of the slab (distance to the right neighbour's edge, periodic)
  float dfof, d2fof;   // linking length b * mean inter-particle distance, squared (both float)
  float lbox, lbox_half;
  double boxsize;
  // search cells (the library's own: any cell size >= Dfof finds the same pairs as the reference's cells do)
  int ncx, ncy, ncz;
  double icx, icy, icz; // 1 / cell size (double: the cell of a coordinate is exact to 1e-12 cells whatever the mesh)
};

// edge / edge_right: x_offset of this task and of its right neighbour; slab_fraction: Local_nx / Nmesh; max_cells: how
// many search cells the caller can afford (4 bytes each)
inline Geometry geometry(const mgp_fof_config &cfg, int nsample, float edge, float edge_right, double slab_fraction,
                         size_t max_cells) {
  Geometry g;
  g.boxsize = cfg.boxsize;
  g.edge = edge;
  g.dx_domain = edge_right - edge;                                          // mm_main.c:236-237
  if (g.dx_domain < 0.0f) g.dx_domain = (float) ((double) g.dx_domain + cfg.boxsize);
  const double n_part = (double) nsample * (double) nsample * (double) nsample;
  const float ipd = (float) (cfg.boxsize / std::pow(n_part, 1. / 3.));      // init_fof, mm_fof.c:83
  g.dfof = (float) ((double) ipd * cfg.b_fof);
  g.d2fof = g.dfof * g.dfof;
  g.lbox = (float) cfg.boxsize;
  g.lbox_half = (float) (cfg.boxsize / 2);
  // search cells: as close to one linking length as `max_cells` allows (the candidates of a particle are the contents
  // of 27 cells: at Dfof = 0.2 mean distances a cell of one mean distance holds 125 times the particles a cell of Dfof
  // does, and the cores of halos are where the pair tests are), never less than the linking length (with a margin);
  // y and z tile the periodic box, x covers the slab and its strip and clamps what lies beyond
  double cs = (double) g.dfof * 1.001;
  for (;;) {
    g.ncy = g.ncz = std::max(1, (int) std::floor(cfg.boxsize / cs));
    g.ncx = (int) std::floor((cfg.boxsize * slab_fraction + cfg.dx_extra) / cs) + 1;
    if ((double) g.ncx * (double) g.ncy * (double) g.ncz <= (double) max_cells || g.ncy == 1) break;
    cs *= 1.125;
  }
  g.icy = g.icz = (double) g.ncy / cfg.boxsize;
  g.icx = 1.0 / cs;
  return g;
}

// float key -> unsigned that sorts like the float (radix sort of MatchMaker's qsort by x, mm_main.c:360)
FOF_HD unsigned orderable(float x) {
  union { float f; unsigned u; } c;
  c.f = x;
  return (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);
}

// mm_main.c:268-269: x = norm_pos (double) * Pos[0] - edge, stored as float; counted when x <= dx_extra
FOF_HD bool to_left(float pos0, double norm_pos, float edge, double dx_extra) {
  const float x = (float) FD_SUB(FD_MUL(norm_pos, (double) pos0), (double) edge);
  return (double) x <= dx_extra;
}

// mm_main.c:310-334: positions in float (float norm_pos), x relative to the slab edge
FOF_HD void translate_pos(const float pos[3], float norm_pos_f, float edge, float x[3]) {
  x[0] = FF_SUB(FF_MUL(norm_pos_f, pos[0]), edge);
  x[1] = FF_MUL(norm_pos_f, pos[1]);
  x[2] = FF_MUL(norm_pos_f, pos[2]);
}
// mm_main.c:327: v = (float)(norm_vel * (Vel + (D dDdy + D2 dD2dy))) in double (norm_vel a float); 331: without COLA
FOF_HD float translate_vel(float vel, float d, float d2, float norm_vel_f, double dDdy, double dD2dy, bool use_cola) {
  if (!use_cola) return FF_MUL(norm_vel_f, vel);
  return (float) FD_MUL((double) norm_vel_f, FD_ADD((double) vel, FD_ADD(FD_MUL((double) d, dDdy), FD_MUL((double) d2, dD2dy))));
}
// mm_main.c:319 (SCALEDEPENDENT): v = (float)(norm_vel * (Vel + (dDdy + dD2dy))), all float
FOF_HD float translate_vel_sd(float vel, float f1, float f2, float norm_vel_f) {
  return FF_MUL(norm_vel_f, FF_ADD(vel, FF_ADD(f1, f2)));
}

FOF_HD int cell_coord(float x, double inv, int n) {
  int c = (int) ((double) x * inv);
  if (c < 0) c = 0;
  if (c >= n) c = n - 1;
  return c;
}
FOF_HD unsigned cell_of(const Geometry &g, const float x[3]) {
  return ((unsigned) cell_coord(x[0], g.icx, g.ncx) * (unsigned) g.ncy + (unsigned) cell_coord(x[1], g.icy, g.ncy)) * (unsigned) g.ncz +
         (unsigned) cell_coord(x[2], g.icz, g.ncz);
}

// get_neighbors, mm_fof.c:130-146: the friendship test.  x is NOT periodic (slab coordinates), y and z are.
FOF_HD bool linked(const Geometry &g, float x0, float y0, float z0, float x1, float y1, float z1) {
  const float dx = fabsf(FF_SUB(x0, x1));
  if (!(dx <= g.dfof)) return false;
  float dy = fabsf(FF_SUB(y0, y1)), dz = fabsf(FF_SUB(z0, z1));
  if (dy > g.lbox_half) dy = FF_SUB(g.lbox, dy);
  if (dz > g.lbox_half) dz = FF_SUB(g.lbox, dz);
  const float d2 = FF_ADD(FF_ADD(FF_MUL(dx, dx), FF_MUL(dy, dy)), FF_MUL(dz, dz));
  return d2 <= g.d2fof;
}

// ---- union-find: the larger root is hooked under the smaller, so the root of a group is its smallest index whatever
// the order of the unions.  CAS(ptr, expected, desired) returns the old value (atomicCAS on the device).
template <class Load>
FOF_HD unsigned find_root(Load &&load, unsigned i) {
  for (;;) {
    const unsigned p = load(i);
    if (p == i) return i;
    i = p;
  }
}
// the same walk with path halving: a particle on the way is re-pointed at its grandparent.  Safe beside concurrent
// unions: only a particle that is no root (and never will be one again) is written, and only with one of its ancestors.
template <class Load, class Store>
FOF_HD unsigned find_root_halving(Load &&load, Store &&store, unsigned i) {
  for (;;) {
    const unsigned p = load(i);
    if (p == i) return i;
    const unsigned gp = load(p);
    if (gp == p) return p;
    store(i, gp);
    i = gp;
  }
}
// returns the root of the united group as far as this thread knows (an ancestor of both a and b from now on)
template <class Load, class Store, class Cas>
FOF_HD unsigned unite(Load &&load, Store &&store, Cas &&cas, unsigned a, unsigned b) {
  for (;;) {
    a = find_root_halving(load, store, a);
    b = find_root_halving(load, store, b);
    if (a == b) return a;
    if (a < b) { const unsigned t = a; a = b; b = t; }
    if (cas(a, a, b) == a) return b;        // a was still a root: hooked.  Otherwise somebody else hooked it: again
  }
}

// cyclic Jacobi rotations of a symmetric 3 x 3 matrix (row-major a, destroyed): eigenvalues w (unordered), eigenvectors in
// the COLUMNS of v.  Stands where the reference calls gsl_eigen_symmv (mm_fof.c:556), a third-party routine.
FOF_HD void jacobi3(double a[9], double w[3], double v[9]) {
  for (int i = 0; i < 9; i++) v[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = 0.0;
    for (int p = 0; p < 3; p++) for (int q = p + 1; q < 3; q++) off = FD_ADD(off, FD_MUL(a[p * 3 + q], a[p * 3 + q]));
    if (off == 0.0) break;
    for (int p = 0; p < 3; p++)
      for (int q = p + 1; q < 3; q++) {
        const double apq = a[p * 3 + q];
        if (apq == 0.0) continue;
        const double theta = FD_DIV(FD_SUB(a[q * 3 + q], a[p * 3 + p]), FD_MUL(2.0, apq));
        const double t = FD_DIV(theta >= 0.0 ? 1.0 : -1.0, FD_ADD(fabs(theta), FD_SQRT(FD_ADD(FD_MUL(theta, theta), 1.0))));
        const double c = FD_DIV(1.0, FD_SQRT(FD_ADD(FD_MUL(t, t), 1.0))), s = FD_MUL(t, c);
        for (int k = 0; k < 3; k++) {
          const double akp = a[k * 3 + p], akq = a[k * 3 + q];
          a[k * 3 + p] = FD_SUB(FD_MUL(c, akp), FD_MUL(s, akq)); a[k * 3 + q] = FD_ADD(FD_MUL(s, akp), FD_MUL(c, akq));
        }
        for (int k = 0; k < 3; k++) {
          const double apk = a[p * 3 + k], aqk = a[q * 3 + k];
          a[p * 3 + k] = FD_SUB(FD_MUL(c, apk), FD_MUL(s, aqk)); a[q * 3 + k] = FD_ADD(FD_MUL(s, apk), FD_MUL(c, aqk));
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = v[k * 3 + p], vkq = v[k * 3 + q];
          v[k * 3 + p] = FD_SUB(FD_MUL(c, vkp), FD_MUL(s, vkq)); v[k * 3 + q] = FD_ADD(FD_MUL(s, vkp), FD_MUL(c, vkq));
        }
      }
  }
  for (int i = 0; i < 3; i++) w[i] = a[i * 3 + i];
}

// ---- get_halos, mm_fof.c:468-611, for one halo, in three pieces.  Every accumulator of the reference (centre-of-mass sums,
// the 18 sums about the centre) is a sequential sum over the members in increasing index order (= the reference's order:
// sorted by x, the buffer particles last); the ADDENDS of the second pass do not depend on one another.  The host runs
// the pieces member by member (halo_properties below); on the device a warp owns a halo, the lanes form the addends of 32
// members at a time and one lane per accumulator adds them up in order (fof_impl.cuh::PropsStep): the same additions in
// the same order, so the same bits.

// first pass, one position component: the running mean picks the periodic image (490-511), then the sum.
// The reference's test is 2 |xx - xs / j| > L.  j |xx - xs / j| = |xx j - xs| to a few ulps, and a member is either near
// the running mean or a box away from it: only within 2 % of the threshold is the division needed to decide as the
// reference does.
FOF_HD void com_add(double &xs, double xx, int j, double L) {
  if (j > 0) {
    const double dj = (double) j, Lj = FD_MUL(L, dj);
    const double d = fabs(FD_SUB(FD_MUL(xx, dj), xs));
    bool far = d > FD_MUL(0.51, Lj);
    if (!far && !(d < FD_MUL(0.49, Lj))) far = FD_MUL(2.0, fabs(FD_SUB(xx, FD_DIV(xs, dj)))) > L;
    if (far) {
      if (FD_MUL(2.0, xx) > L) xx = FD_SUB(xx, L); else xx = FD_ADD(xx, L);
    }
  }
  xs = FD_ADD(xs, xx);
}

enum { FOF_NACC = 18 };   // second pass: x_rms^2 [3], v_rms^2 [3], inertia tensor [9], angular momentum [3]

// second pass: the 18 addends of one member (519-555)
FOF_HD void member_terms(const float xf[3], const float vf[3], const float xavg[3], const float vavg[3], double L, double t[FOF_NACC]) {
  double dx[3], dv[3];
  for (int ax = 0; ax < 3; ax++) {
    double xx = (double) xf[ax];
    if (FD_MUL(2.0, fabs(FD_SUB(xx, (double) xavg[ax]))) > L) {
      if (FD_MUL(2.0, xx) > L) xx = FD_SUB(xx, L); else xx = FD_ADD(xx, L);
    }
    dx[ax] = FD_SUB(xx, (double) xavg[ax]);
    dv[ax] = FD_SUB((double) vf[ax], (double) vavg[ax]);
  }
  for (int ax = 0; ax < 3; ax++) { t[ax] = FD_MUL(dx[ax], dx[ax]); t[3 + ax] = FD_MUL(dv[ax], dv[ax]); }
  for (int ax = 0; ax < 3; ax++)
    for (int a2 = 0; a2 < 3; a2++) t[6 + a2 + 3 * ax] = FD_MUL(dx[ax], dx[a2]);
  t[15] = FD_SUB(FD_MUL(dx[1], dv[2]), FD_MUL(dx[2], dv[1]));
  t[16] = FD_SUB(FD_MUL(dx[2], dv[0]), FD_MUL(dx[0], dv[2]));
  t[17] = FD_SUB(FD_MUL(dx[0], dv[1]), FD_MUL(dx[1], dv[0]));
}

// the record from the sums (556-611)
FOF_HD void finish_halo(const Geometry &g, int np, double mass_particle, const float xavg[3], const float vavg[3],
                        const double acc[FOF_NACC], mgp_fof_halo &h) {
  const double L = g.boxsize;
  double in[9], w[3], e[9];
  for (int i = 0; i < 9; i++) in[i] = acc[6 + i];
  jacobi3(in, w, e);
  int o[3] = {0, 1, 2};                               // eigenvalues in descending order (compare_evals, 63-68)
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2 - i; j++)
      if (w[o[j]] < w[o[j + 1]]) { const int t = o[j]; o[j] = o[j + 1]; o[j + 1] = t; }
  h.np = np;
  h.m_halo = (float) FD_MUL((double) np, mass_particle);
  if (w[o[0]] <= 0) {
    h.b = 0; h.c = 0;
    for (int k = 0; k < 3; k++) { h.ea[k] = 0; h.eb[k] = 0; h.ec[k] = 0; }
  } else {
    h.b = (float) FD_DIV(w[o[1]], w[o[0]]);
    h.c = (float) FD_DIV(w[o[2]], w[o[0]]);
    for (int k = 0; k < 3; k++) { h.ea[k] = (float) e[k * 3 + o[0]]; h.eb[k] = (float) e[k * 3 + o[1]]; h.ec[k] = (float) e[k * 3 + o[2]]; }
  }
  for (int ax = 0; ax < 3; ax++) {
    h.x_rms[ax] = (float) FD_SQRT(FD_DIV(acc[ax], (double) np));
    h.v_rms[ax] = (float) FD_SQRT(FD_DIV(acc[3 + ax], (double) np));
    h.lam[ax] = (float) acc[15 + ax];
    float xa = xavg[ax];                              // wrap the centre of mass (598-603)
    if (xa < 0) xa = (float) FD_ADD((double) xa, L);
    else if ((double) xa >= L) xa = (float) FD_SUB((double) xa, L);
    h.x_avg[ax] = xa;
    h.v_avg[ax] = vavg[ax];
  }
  h.x_avg[0] = FF_ADD(h.x_avg[0], g.edge);          // 604
}

// one halo, member by member: members ids[0 .. np) of the particle arrays x[3][stride], v[3][stride]
FOF_HD void halo_properties(const Geometry &g, const float *x, const float *v, size_t stride, const unsigned *ids, int np,
                            double mass_particle, mgp_fof_halo &h) {
  const double L = g.boxsize;
  double xs[3] = {0, 0, 0}, vs[3] = {0, 0, 0};
  for (int j = 0; j < np; j++) {
    const unsigned ip = ids[j];
    for (int ax = 0; ax < 3; ax++) {
      com_add(xs[ax], (double) x[ax * stride + ip], j, L);
      vs[ax] = FD_ADD(vs[ax], (double) v[ax * stride + ip]);
    }
  }
  float xavg[3], vavg[3];
  for (int ax = 0; ax < 3; ax++) { xavg[ax] = (float) FD_DIV(xs[ax], (double) np); vavg[ax] = (float) FD_DIV(vs[ax], (double) np); }
  double acc[FOF_NACC];
  for (int a = 0; a < FOF_NACC; a++) acc[a] = 0.0;
  for (int j = 0; j < np; j++) {
    const unsigned ip = ids[j];
    const float xf[3] = {x[ip], x[stride + ip], x[2 * stride + ip]}, vf[3] = {v[ip], v[stride + ip], v[2 * stride + ip]};
    double t[FOF_NACC];
    member_terms(xf, vf, xavg, vavg, L, t);
    for (int a = 0; a < FOF_NACC; a++) acc[a] = FD_ADD(acc[a], t[a]);
  }
  finish_halo(g, np, mass_particle, xavg, vavg, acc, h);
}

}  // namespace fof
}  // namespace mgp
