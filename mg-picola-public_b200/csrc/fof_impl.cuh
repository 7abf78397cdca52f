// The FoF halo finder as ONE sequence of data-parallel steps, written against a small back end (allocate, run a functor
// over n indices, exclusive scan, stable radix sort of pairs, neighbour exchange): fof.cu runs it on the GPU (kernel
// launches, CUB, NCCL), tests/host/fof_emul.cu runs the very same sequence and the very same functors on the CPU for
// several emulated tasks and compares with the reference's catalogue.  See fof.cu for the design and fof.cuh for the
// arithmetic.
#pragma once

#include <algorithm>
#include <vector>

#include "fof.cuh"

namespace mgp {
namespace fof {

// atomics of the functors: CUDA's on the device, plain operations in the (sequential) host emulation
struct Atom {
  static FOF_HD unsigned add(unsigned *p, unsigned v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    const unsigned o = *p; *p = o + v; return o;
#endif
  }
  static FOF_HD void add64(unsigned long long *p, unsigned long long v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
  }
  static FOF_HD unsigned cas(unsigned *p, unsigned expect, unsigned desired) {
#if defined(__CUDA_ARCH__)
    return atomicCAS(p, expect, desired);
#else
    const unsigned o = *p; if (o == expect) *p = desired; return o;
#endif
  }
  static FOF_HD unsigned load(const unsigned *p) { return *(const volatile unsigned *) p; }
  static FOF_HD void store(unsigned *p, unsigned v) { *(volatile unsigned *) p = v; }
};

// the particle store as the library keeps it: 16-byte records, or [3][cap] arrays for the
// scale-dependent per-particle fields
struct Store {
  const float4 *pA, *pB, *pC;
  const float2 *pE;
  const float *f1, *f2;      // scale_dependent: P.dDdy, P.dD2dy as [3][cap] (f2 may be NULL: reads as zero)
  size_t cap;
  int scale_dependent, use_cola;
};

// ---- the steps (one functor per former kernel; `i` is the index a thread handles) ----

struct KeyStep {          // x key of every slab particle, and the strip count (mm_main.c:266-271)
  const float4 *pA; double norm_pos; float norm_pos_f, edge; double dx_extra;
  unsigned *key, *val; unsigned long long *n_toleft;
  FOF_HD void operator()(size_t i) const {
    const float px = pA[i].x;
    key[i] = orderable(FF_SUB(FF_MUL(norm_pos_f, px), edge));
    val[i] = (unsigned) i;
    if (to_left(px, norm_pos, edge, dx_extra)) Atom::add64(n_toleft, 1ull);
  }
};

struct TranslateStep {    // f-th particle of the x order: MatchMaker's position and velocity (mm_main.c:310-338)
  Store s; const unsigned *perm; float norm_pos_f, norm_vel_f, edge; double dDdy, dD2dy;
  float *x, *v; size_t N;
  FOF_HD void operator()(size_t f) const {
    const unsigned i = perm[f];
    const float4 a = s.pA[i], b = s.pB[i];
    const float pos[3] = {a.x, a.y, a.z}, vel[3] = {b.x, b.y, b.z};
    float xo[3], d[3] = {0, 0, 0}, d2[3] = {0, 0, 0};
    translate_pos(pos, norm_pos_f, edge, xo);
    if (s.use_cola) {
      if (s.scale_dependent) {
        for (int ax = 0; ax < 3; ax++) { d[ax] = s.f1[ax * s.cap + i]; d2[ax] = s.f2 ? s.f2[ax * s.cap + i] : 0.0f; }
      } else {
        const float4 cc = s.pC[i];
        const float2 e = s.pE[i];
        d[0] = cc.x; d[1] = cc.y; d[2] = cc.z; d2[0] = cc.w; d2[1] = e.x; d2[2] = e.y;
      }
    }
    for (int ax = 0; ax < 3; ax++) {
      x[ax * N + f] = xo[ax];
      v[ax * N + f] = (s.scale_dependent && s.use_cola) ? translate_vel_sd(vel[ax], d[ax], d2[ax], norm_vel_f)
                                                        : translate_vel(vel[ax], d[ax], d2[ax], norm_vel_f, dDdy, dD2dy, s.use_cola != 0);
    }
  }
};

struct ShiftStep {        // particles received from the right neighbour sit behind its edge (mm_main.c:370-373)
  float *x0_buf; float dx_domain;
  FOF_HD void operator()(size_t k) const { x0_buf[k] = FF_ADD(x0_buf[k], dx_domain); }
};

struct CellCountStep {
  const float *x; size_t N; Geometry g; unsigned *cellid, *rank, *count;
  FOF_HD void operator()(size_t f) const {
    const float p[3] = {x[f], x[N + f], x[2 * N + f]};
    const unsigned c = cell_of(g, p);
    cellid[f] = c;
    rank[f] = Atom::add(&count[c], 1u);
  }
};

struct CellFillStep {     // positions packed per cell, the x-order index in the fourth word
  const float *x; size_t N; const unsigned *cellid, *rank, *start; float4 *packed;
  FOF_HD void operator()(size_t f) const {
    union { unsigned u; float fl; } w;
    w.u = (unsigned) f;
    float4 q;
    q.x = x[f]; q.y = x[N + f]; q.z = x[2 * N + f]; q.w = w.fl;
    packed[(size_t) start[cellid[f]] + rank[f]] = q;
  }
};

struct IotaStep {
  unsigned *parent;
  FOF_HD void operator()(size_t f) const { parent[f] = (unsigned) f; }
};

struct LinkStep {         // every pair of friends, each once: the rest of the own cell and the 13 cells "ahead" of it
  const float4 *packed; const unsigned *start; Geometry g; unsigned *parent;
  FOF_HD static unsigned word(float f) { union { unsigned u; float fl; } w; w.fl = f; return w.u; }
  FOF_HD void operator()(size_t s) const {
    unsigned *par = parent;
    auto load = [par](unsigned i) { return Atom::load(par + i); };
    auto store = [par](unsigned i, unsigned v) { Atom::store(par + i, v); };
    auto cas = [par](unsigned at, unsigned expect, unsigned desired) { return Atom::cas(par + at, expect, desired); };
    const float4 me = packed[s];
    unsigned anc = word(me.w);      // the last known ancestor of this particle: later unions start their walk there
    const int cx = cell_coord(me.x, g.icx, g.ncx), cy = cell_coord(me.y, g.icy, g.ncy), cz = cell_coord(me.z, g.icz, g.ncz);
    const unsigned own = ((unsigned) cx * (unsigned) g.ncy + (unsigned) cy) * (unsigned) g.ncz + (unsigned) cz;
    for (unsigned t = (unsigned) s + 1u, t1 = start[own + 1]; t < t1; t++) {
      const float4 o = packed[t];
      if (linked(g, me.x, me.y, me.z, o.x, o.y, o.z)) anc = unite(load, store, cas, anc, word(o.w));
    }
    for (int k = 14; k < 27; k++) {                                   // (ox, oy, oz) after (0, 0, 0) in lexicographic order
      const int ox = k / 9 - 1, oy = (k / 3) % 3 - 1, oz = k % 3 - 1;
      const int x2 = cx + ox;
      if (x2 >= g.ncx) continue;                                      // x is open
      const int y2 = (cy + oy + g.ncy) % g.ncy, z2 = (cz + oz + g.ncz) % g.ncz;
      const unsigned c2 = ((unsigned) x2 * (unsigned) g.ncy + (unsigned) y2) * (unsigned) g.ncz + (unsigned) z2;
      for (unsigned t = start[c2], t1 = start[c2 + 1]; t < t1; t++) {
        const float4 o = packed[t];
        if (linked(g, me.x, me.y, me.z, o.x, o.y, o.z)) anc = unite(load, store, cas, anc, word(o.w));
      }
    }
  }
};

// root of every particle, size of every group, and whether it holds a particle of this slab (only those seed groups,
// mm_fof.c:371-392: a group of received particles alone belongs to the right neighbour)
struct RootStep {
  const unsigned *parent; size_t n_dom; unsigned *root, *size; unsigned char *dom;
  FOF_HD void operator()(size_t f) const {
    const unsigned *par = parent;
    auto load = [par](unsigned i) { return par[i]; };
    const unsigned r = find_root(load, (unsigned) f);
    root[f] = r;
    Atom::add(&size[r], 1u);
    if (f < n_dom) dom[r] = 1;
  }
};

struct InGroupStep {      // fof_id != 0  <=>  in a group of at least two that this task seeded
  const unsigned *root, *size; const unsigned char *dom; unsigned char *ingroup;
  FOF_HD void operator()(size_t f) const { const unsigned r = root[f]; ingroup[f] = (size[r] >= 2u && dom[r]) ? 1 : 0; }
};

struct MemberStep {       // resolve redundancies (mm_fof.c:425-437): a strip particle the left neighbour grouped is not mine
  size_t n_toleft; const unsigned *root; const unsigned char *ingroup, *back; unsigned char *member; unsigned *gcount;
  FOF_HD void operator()(size_t f) const {
    const bool m = ingroup[f] && !(f < n_toleft && back[f]);
    member[f] = m ? 1 : 0;
    if (m) Atom::add(&gcount[root[f]], 1u);
  }
};

struct MarkStep {         // per root (and one slot past the end for the totals): a halo?  how many members?  per particle: listed?
  size_t N; const unsigned *gcount, *root; const unsigned char *member; unsigned np_min; unsigned *hflag, *hcnt, *sel;
  FOF_HD void operator()(size_t r) const {
    const unsigned g = r < N ? gcount[r] : 0u;
    const bool h = g >= np_min;
    hflag[r] = h ? 1u : 0u;
    hcnt[r] = h ? g : 0u;
    sel[r] = (r < N && member[r] && gcount[root[r]] >= np_min) ? 1u : 0u;
  }
};

struct CompactStep {      // compaction in index order: the k-th listed particle and the number of its halo
  const unsigned *root; const unsigned char *member; const unsigned *gcount; unsigned np_min; const unsigned *pos, *hidx;
  unsigned *key, *val;
  FOF_HD void operator()(size_t f) const {
    const unsigned r = root[f];
    if (member[f] && gcount[r] >= np_min) { key[pos[f]] = hidx[r]; val[pos[f]] = (unsigned) f; }
  }
};

struct SegmentStep {      // where a halo's members start in the list sorted by halo, and how many they are
  const unsigned *hidx, *hoff, *gcount; unsigned np_min; unsigned *seg, *cnt;
  FOF_HD void operator()(size_t r) const {
    const unsigned g = gcount[r];
    if (g >= np_min) { seg[hidx[r]] = hoff[r]; cnt[hidx[r]] = g; }
  }
};

// Halo properties.  Host: one call per halo, member by member.  Device: one WARP per halo (k_fof_warp_step with
// FOF_WARPS_PER_CTA warps per CTA), 32 members at a time: every lane gathers one member (the loads of a chunk are in flight
// together), forms its addends and stages them in shared memory; then one lane per accumulator adds the chunk's addends in
// member order.  What is sequential in the reference -- the chain of additions per accumulator, and in the first pass
// the image choice against the running mean -- stays sequential (18 and 6 chains side by side); everything else runs
// across the lanes.
enum { FOF_WARPS_PER_CTA = 4, FOF_STAGE = FOF_NACC + 1 };

struct PropsStep {
  Geometry g; const float *x, *v; size_t N; const unsigned *members, *seg, *cnt; double mass_particle; mgp_fof_halo *out;
  FOF_HD void operator()(size_t h) const {
    const unsigned *ids = members + seg[h];
    const int np = (int) cnt[h];
    mgp_fof_halo r;
#if defined(__CUDA_ARCH__)
    __shared__ double stage_all[FOF_WARPS_PER_CTA][32 * FOF_STAGE];
    double *stage = stage_all[threadIdx.x >> 5];
    const int lane = (int) (threadIdx.x & 31u);
    const double L = g.boxsize;
    const unsigned full = 0xffffffffu;
    // first pass: lanes 0-2 own the position sums, lanes 3-5 the velocity sums
    double sum = 0.0;
    for (int base = 0; base < np; base += 32) {
      const int m = base + lane, n = min(32, np - base);
      if (m < np) {
        const unsigned ip = ids[m];
        for (int a = 0; a < 3; a++) { stage[lane * FOF_STAGE + a] = (double) x[a * N + ip]; stage[lane * FOF_STAGE + 3 + a] = (double) v[a * N + ip]; }
      }
      __syncwarp(full);
      if (lane < 3) {
        for (int j = 0; j < n; j++) com_add(sum, stage[j * FOF_STAGE + lane], base + j, L);
      } else if (lane < 6) {
        for (int j = 0; j < n; j++) sum = FD_ADD(sum, stage[j * FOF_STAGE + lane]);
      }
      __syncwarp(full);
    }
    const float avg = (float) FD_DIV(sum, (double) np);
    float xavg[3], vavg[3];
    for (int a = 0; a < 3; a++) { xavg[a] = __shfl_sync(full, avg, a); vavg[a] = __shfl_sync(full, avg, 3 + a); }
    // second pass: lane a < 18 owns accumulator a
    double acc1 = 0.0;
    for (int base = 0; base < np; base += 32) {
      const int m = base + lane, n = min(32, np - base);
      if (m < np) {
        const unsigned ip = ids[m];
        const float xf[3] = {x[ip], x[N + ip], x[2 * N + ip]}, vf[3] = {v[ip], v[N + ip], v[2 * N + ip]};
        double t[FOF_NACC];
        member_terms(xf, vf, xavg, vavg, L, t);
        for (int a = 0; a < FOF_NACC; a++) stage[lane * FOF_STAGE + a] = t[a];
      }
      __syncwarp(full);
      if (lane < FOF_NACC)
        for (int j = 0; j < n; j++) acc1 = FD_ADD(acc1, stage[j * FOF_STAGE + lane]);
      __syncwarp(full);
    }
    double acc[FOF_NACC];
    for (int a = 0; a < FOF_NACC; a++) acc[a] = __shfl_sync(full, acc1, a);
    if (lane != 0) return;
    finish_halo(g, np, mass_particle, xavg, vavg, acc, r);
#else
    halo_properties(g, x, v, N, ids, np, mass_particle, r);
#endif
    out[h] = r;
  }
};

struct Task {             // who this task is among the P tasks of the run
  int rank, P, nsample, p_start;     // ThisTask, NTask, Nsample, Local_p_start
  double slab_fraction;              // Local_nx / Nmesh
};

inline void check(bool ok, const char *what);   // provided by the back end's translation unit (throws)

// B: alloc<T>(n), zero(p, bytes), run(n, functor[, block]), run_warp(n, functor) (one warp per index on the device),
// max_cells(n) (search cells the memory allows for n particles), scan(p, n) (exclusive, in place), sort(k0, k1, v0, v1,
// n, bits) (stable, ascending, result in k1 / v1), read32 / read64(p), download(host, dev, bytes),
// gather2(mine[2], all[2 P]), strip_exchange(x, v, N, n_dom, n_toleft, n_buf), flag_exchange(send, n_buf, recv, n_toleft)
template <class B>
void find_halos(B &be, const Store &st, size_t n_dom, const Task &tk, const mgp_fof_config &cfg, std::vector<mgp_fof_halo> &halos) {
  halos.clear();
  const float norm_pos_f = (float) cfg.norm_pos, norm_vel_f = (float) cfg.norm_vel;
  const float edge = (float) ((double) tk.p_start * (cfg.boxsize / (double) tk.nsample));          // mm_main.c:227

  // 1. keys, the strip count, the order by x (MatchMaker's qsort, mm_main.c:360)
  unsigned *key0 = be.template alloc<unsigned>(n_dom), *key1 = be.template alloc<unsigned>(n_dom);
  unsigned *val0 = be.template alloc<unsigned>(n_dom), *perm = be.template alloc<unsigned>(n_dom);
  unsigned long long *d_cnt = be.template alloc<unsigned long long>(1);
  be.zero(d_cnt, sizeof(unsigned long long));
  be.run(n_dom, KeyStep{st.pA, cfg.norm_pos, norm_pos_f, edge, cfg.dx_extra, key0, val0, d_cnt});
  be.sort(key0, key1, val0, perm, n_dom, 32);
  // every task's strip count and first particle plane: the neighbours' are needed
  std::vector<unsigned long long> all(2 * (size_t) tk.P);
  unsigned long long mine[2] = {be.read64(d_cnt), (unsigned long long) tk.p_start};
  be.gather2(mine, all.data());
  const int right = (tk.rank + 1) % tk.P;
  const size_t n_toleft = (size_t) all[2 * (size_t) tk.rank], n_buf = (size_t) all[2 * (size_t) right];
  const float edge_right = (float) ((double) all[2 * (size_t) right + 1] * (cfg.boxsize / (double) tk.nsample));
  const size_t N = n_dom + n_buf;
  const Geometry g = geometry(cfg, tk.nsample, edge, edge_right, tk.slab_fraction, std::min(be.max_cells(N), (size_t) 0xffffff00ull));
  check(N < 0xfffffff0ull, "mgp_fof_find: more than 2^32 particles on one rank");
  const size_t ncell = (size_t) g.ncx * (size_t) g.ncy * (size_t) g.ncz;
  check(ncell < 0xfffffff0ull, "mgp_fof_find: more than 2^32 search cells on one rank");

  // 2. MatchMaker's particles in x order; the strip to the left neighbour, the right neighbour's strip behind mine
  float *x = be.template alloc<float>(3 * N), *v = be.template alloc<float>(3 * N);
  be.run(n_dom, TranslateStep{st, perm, norm_pos_f, norm_vel_f, edge, cfg.dDdy, cfg.dD2dy, x, v, N});
  be.strip_exchange(x, v, N, n_dom, n_toleft, n_buf);
  be.run(n_buf, ShiftStep{x + n_dom, g.dx_domain});

  // 3. search cells
  unsigned *cellid = be.template alloc<unsigned>(N), *rank = be.template alloc<unsigned>(N), *start = be.template alloc<unsigned>(ncell + 1);
  float4 *packed = be.template alloc<float4>(N);
  be.zero(start, (ncell + 1) * sizeof(unsigned));
  be.run(N, CellCountStep{x, N, g, cellid, rank, start});
  be.scan(start, ncell + 1);
  be.run(N, CellFillStep{x, N, cellid, rank, start, packed});

  // 4. friends of friends
  unsigned *parent = be.template alloc<unsigned>(N), *root = cellid /* free again */, *size = be.template alloc<unsigned>(N);
  unsigned *gcount = be.template alloc<unsigned>(N);
  unsigned char *dom = be.template alloc<unsigned char>(N), *ingroup = be.template alloc<unsigned char>(N);
  unsigned char *member = be.template alloc<unsigned char>(N), *back = be.template alloc<unsigned char>(n_toleft);
  be.run(N, IotaStep{parent});
  be.run(N, LinkStep{packed, start, g, parent}, 128);
  be.zero(size, N * sizeof(unsigned));
  be.zero(gcount, N * sizeof(unsigned));
  be.zero(dom, N);
  be.run(N, RootStep{parent, n_dom, root, size, dom});
  be.run(N, InGroupStep{root, size, dom, ingroup});

  // 5. what the left neighbour made of my strip, then who is left in which group
  be.flag_exchange(ingroup + n_dom, n_buf, back, n_toleft);
  be.run(N, MemberStep{n_toleft, root, ingroup, back, member, gcount});
  unsigned *hflag = be.template alloc<unsigned>(N + 1), *hcnt = be.template alloc<unsigned>(N + 1), *sel = be.template alloc<unsigned>(N + 1);
  const unsigned np_min = (unsigned) cfg.np_min;
  be.run(N + 1, MarkStep{N, gcount, root, member, np_min, hflag, hcnt, sel});
  be.scan(hflag, N + 1);        // -> number of a root's halo;            [N] = halos
  be.scan(hcnt, N + 1);         // -> first slot of a root's halo;        [N] = listed particles
  be.scan(sel, N + 1);          // -> slot of a listed particle;          [N] = the same
  const size_t n_halos = be.read32(hflag + N), n_mem = be.read32(hcnt + N);
  check(be.read32(sel + N) == n_mem, "mgp_fof_find: member lists do not add up");
  if (n_halos == 0) return;

  // members in index order (= the reference's order inside a halo), grouped by halo with a stable sort
  unsigned *mkey0 = be.template alloc<unsigned>(n_mem), *mkey1 = be.template alloc<unsigned>(n_mem);
  unsigned *mval0 = be.template alloc<unsigned>(n_mem), *mval1 = be.template alloc<unsigned>(n_mem);
  unsigned *seg = be.template alloc<unsigned>(n_halos), *cnt = be.template alloc<unsigned>(n_halos);
  be.run(N, CompactStep{root, member, gcount, np_min, sel, hflag, mkey0, mval0});
  int bits = 1;
  while (bits < 32 && (1ull << bits) < n_halos) bits++;
  be.sort(mkey0, mkey1, mval0, mval1, n_mem, bits);
  be.run(N, SegmentStep{hflag, hcnt, gcount, np_min, seg, cnt});

  // 6. properties: one thread per halo walks its members in the reference's order
  mgp_fof_halo *d_out = be.template alloc<mgp_fof_halo>(n_halos);
  be.run_warp(n_halos, PropsStep{g, x, v, N, mval1, seg, cnt, 1.0e10 * cfg.mass_part, d_out});
  halos.resize(n_halos);
  be.download(halos.data(), d_out, n_halos * sizeof(mgp_fof_halo));
  std::stable_sort(halos.begin(), halos.end(), [](const mgp_fof_halo &a, const mgp_fof_halo &b) { return a.np > b.np; });   // mm_fof.c:459
}

}  // namespace fof
}  // namespace mgp
