// Scale-dependent growth (-DSCALEDEPENDENT): the per-step Lagrangian displacement fields.
//
// Replaces from_cdisp_store_to_ZA + assign_displacment_field_to_particles (2LPT.c:1539-2005), the
// SCALEDEPENDENT branches of the particle initialisation (main.c:231-304), of Kick (main.c:705-717)
// and of Drift (main.c:760-770).
//
// The reference keeps delta1_k and delta2_k (cdelta_cdm, cdelta_cdm2) from the IC generator and, four
// times per step, multiplies one of them by i k / k^2 * G(|k|, a), transforms the three components
// back, reads them out at the Lagrangian lattice into ZA_D[3][Local_np Nsample^2], subtracts the mean
// and copies ZA_D[coord_q] into the particle that was born at coord_q -- through a request / response
// exchange when that lattice point belongs to another task.
//
// Here:
//  * the growth factor crosses the C ABI as a table over the integer m = |d|^2 (it depends on the mode
//    only through |k| = 2 pi sqrt(m) / Box), filled by the driver with its own growth_*_scaledependent;
//  * the three components are built in the force grids (dead between MtoParticles and the next
//    PtoMesh) and transformed with the batched c2r plan;
//  * there is no ZA_D array: the lattice mean is one reduction pass over the three grids, after which
//    every particle interpolates at its own Lagrangian point (known from its ID,
//    ID = (i Nsample + j) Nsample + k, main.c:263) and stores the mean-subtracted float;
//  * a particle living on another rank than the slab it was born on sends its 8-byte Lagrangian index
//    straight to the birth rank (NVSwitch: no ring), which answers with 12 bytes per field.  The
//    request lists are built once per particle order and reused for the four fields of a step;
//  * the four per-particle fields (D, D2, dDdy, dD2dy) are per-step temporaries -- written here, read
//    by Kick / Drift / Output before the next MoveParticles -- so they live in plain [3][cap] float
//    arrays that the sort never has to permute;
//  * optional merged mode: only D + D2 and dDdy + dD2dy are ever used (main.c:712, 767, 962;
//    compute_pofk.c:324), so one call can build G1 delta1 + norm2 G2 delta2 directly: 6 instead of 12
//    inverse transforms per step.
#include "common.cuh"
#include "klayout.cuh"
#include "reduce.cuh"

#include <cmath>

namespace mgp {

#ifndef MGP_SD_BLOCKS
#define MGP_SD_BLOCKS 6        // resident CTAs per SM the fused Kick / Drift kernels are compiled for (register cap 40)
#endif

// ------------------------------------------------------------------ k-space

// (2LPT.c:1584-1630)  out_a = ( -Im(d) * kvec_a / kmag2 * g , Re(d) * kvec_a / kmag2 * g ),  g = norm * G[m]
// with the IC code's Nyquist convention idx < N/2 ? idx : idx - N.  Two sources are summed in merged mode.
template <typename T>
__global__ void __launch_bounds__(256)
k_sd_field(KL L, const typename Cpx<T>::type *__restrict__ d1, const double *__restrict__ g1, double norm1,
           const typename Cpx<T>::type *__restrict__ d2, const double *__restrict__ g2, double norm2,
           typename Cpx<T>::type *__restrict__ o0, typename Cpx<T>::type *__restrict__ o1,
           typename Cpx<T>::type *__restrict__ o2, double box) {
  typedef typename Cpx<T>::type C;
  const int N = L.N, h = N / 2;
  const double twopi_over_box = 2 * 3.14159265358979323846 / box;
  KLOOP(e, L) {
    int i, j, k;
    kl_decode(L, e, i, j, k);
    const int c0 = i < h ? i : i - N, c1 = j < h ? j : j - N, c2 = k < h ? k : k - N;
    const long long m = (long long) c0 * c0 + (long long) c1 * c1 + (long long) c2 * c2;
    C out[3];
    if (m == 0) {
      out[0].x = out[0].y = out[1].x = out[1].y = out[2].x = out[2].y = (T) 0;
    } else {
      // kvec_a / kmag2 = d_a / (|d|^2 * 2 pi / Box): one division per mode instead of the reference's nine
      // (the quotient differs from kvec[a] / kmag2 of 2LPT.c:1620 in the last bit of a double)
      const double inv = 1.0 / ((double) m * twopi_over_box);
      const double kq[3] = {(double) c0 * inv, (double) c1 * inv, (double) c2 * inv};
      const C s = d1[e];
      const double ga = norm1 * g1[m];
      if (d2 == nullptr) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
          out[a].x = (T) (-(double) s.y * kq[a] * ga);
          out[a].y = (T) ((double) s.x * kq[a] * ga);
        }
      } else {
        const C t = d2[e];
        const double gb = norm2 * g2[m];
#pragma unroll
        for (int a = 0; a < 3; a++) {
          out[a].x = (T) (-(double) s.y * kq[a] * ga + -(double) t.y * kq[a] * gb);
          out[a].y = (T) ((double) s.x * kq[a] * ga + (double) t.x * kq[a] * gb);
        }
      }
    }
    o0[e] = out[0]; o1[e] = out[1]; o2[e] = out[2];
  }
}

// ------------------------------------------------------------------ Lagrangian read-out

// trilinear interpolation of three real grids at the Lagrangian point (n, m, p) of the GLOBAL lattice
// (2LPT.c:1657-1705); (n) must belong to this rank's particle planes
template <typename T>
__device__ __forceinline__ void sd_interp(long long n, int m, int p, int ns, int N, int NZ, int x0, int nx,
                                          const T *__restrict__ g0, const T *__restrict__ g1,
                                          const T *__restrict__ g2, double r[3]) {
  double u = (double) (n * (long long) N) / (double) ns;
  double v = (double) ((long long) m * N) / (double) ns;
  double w = (double) ((long long) p * N) / (double) ns;
  int i = (int) u, j = (int) v, k = (int) w;
  if (i == x0 + nx) i = x0 + nx - 1;
  if (i < x0) i = x0;
  if (j == N) j = N - 1;
  if (k == N) k = N - 1;
  u -= i; v -= j; w -= k;
  i -= x0;
  const int i2 = i + 1;
  int j2 = j + 1, k2 = k + 1;
  if (j2 >= N) j2 -= N;
  if (k2 >= N) k2 -= N;
  const size_t rz = (size_t) 2 * NZ;
  const size_t a = ((size_t) i * N + j) * rz, b = ((size_t) i * N + j2) * rz, cc = ((size_t) i2 * N + j) * rz,
               d = ((size_t) i2 * N + j2) * rz;
  if (u == 0.0 && v == 0.0 && w == 0.0) {      // lattice point on a mesh point (Nmesh % Nsample == 0): f1 = 1, rest 0
    r[0] = (double) g0[a + k]; r[1] = (double) g1[a + k]; r[2] = (double) g2[a + k];
    return;
  }
  const double f1 = (1 - u) * (1 - v) * (1 - w), f2 = (1 - u) * (1 - v) * w, f3 = (1 - u) * v * (1 - w), f4 = (1 - u) * v * w;
  const double f5 = u * (1 - v) * (1 - w), f6 = u * (1 - v) * w, f7 = u * v * (1 - w), f8 = u * v * w;
#define RD(G)                                                                                                     \
  ((double) G[a + k] * f1 + (double) G[a + k2] * f2 + (double) G[b + k] * f3 + (double) G[b + k2] * f4 +        \
   (double) G[cc + k] * f5 + (double) G[cc + k2] * f6 + (double) G[d + k] * f7 + (double) G[d + k2] * f8)
  r[0] = RD(g0); r[1] = RD(g1); r[2] = RD(g2);
#undef RD
}

// sumdis_D over this rank's lattice points (2LPT.c:1699); partial[3 * block + axis]
template <typename T>
__global__ void __launch_bounds__(256)
k_sd_lattice_sum(size_t nloc, int ns, int p0, int N, int NZ, int x0, int nx, const T *__restrict__ g0,
                 const T *__restrict__ g1, const T *__restrict__ g2, double *__restrict__ partial) {
  double s0 = 0, s1 = 0, s2 = 0;
  for (size_t c = blockIdx.x * (size_t) blockDim.x + threadIdx.x; c < nloc; c += (size_t) gridDim.x * blockDim.x) {
    const int p = (int) (c % ns);
    const size_t t = c / ns;
    const int m = (int) (t % ns);
    const long long n = (long long) (t / ns) + p0;
    double r[3];
    sd_interp<T>(n, m, p, ns, N, NZ, x0, nx, g0, g1, g2, r);
    s0 += r[0]; s1 += r[1]; s2 += r[2];
  }
  block_sum3(s0, s1, s2);
  if (threadIdx.x == 0) { partial[3 * blockIdx.x] = s0; partial[3 * blockIdx.x + 1] = s1; partial[3 * blockIdx.x + 2] = s2; }
}

__device__ __forceinline__ unsigned long long particle_id(const float4 a, const float4 b) {
  return ((unsigned long long) __float_as_uint(b.w) << 32) | (unsigned long long) __float_as_uint(a.w);
}

// ZA_D[coord] = dis (float_kind); ZA_D -= sumdis; P.X = ZA_D   (2LPT.c:1700, 1724-1728, 1817-1827)
template <typename T>
__device__ __forceinline__ float sd_value(double dis, double mean) {
  T za = (T) dis;
  za = (T) ((double) za - mean);
  return (float) za;
}

// every particle whose birth plane is local interpolates for itself
// ID32: Nsample^3 < 2^32, the high word of every ID (kept in pB.w) is zero and pB need not be read
template <typename T, int ID32>
__global__ void __launch_bounds__(256)
k_sd_assign(size_t np, const float4 *__restrict__ pA, const float4 *__restrict__ pB, int ns, int p0, int npl, int N,
            int NZ, int x0, int nx, const T *__restrict__ g0, const T *__restrict__ g1, const T *__restrict__ g2,
            double m0, double m1, double m2, float *__restrict__ out, size_t cap) {
  const unsigned long long ns2 = (unsigned long long) ns * ns;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < np; i += (size_t) gridDim.x * blockDim.x) {
    const unsigned long long id = ID32 ? (unsigned long long) __float_as_uint(pA[i].w) : particle_id(pA[i], pB[i]);
    const long long n = (long long) (id / ns2);
    if (n < p0 || n >= p0 + npl) continue;             // born on another rank: answered by k_sd_serve there
    const unsigned rem = (unsigned) (id - (unsigned long long) n * ns2);
    double r[3];
    sd_interp<T>(n, (int) (rem / ns), (int) (rem % ns), ns, N, NZ, x0, nx, g0, g1, g2, r);
    out[i] = sd_value<T>(r[0], m0); out[cap + i] = sd_value<T>(r[1], m1); out[2 * cap + i] = sd_value<T>(r[2], m2);
  }
}

// ------------------------------------------------------------------ remote lattice points (P > 1)

__device__ __forceinline__ int birth_rank(long long n, int ns, int N, int block, int P) {
  const int slab = (int) ((double) (n * (long long) N) / (double) ns);      // initialize_parts, 2LPT.c:130-150
  const int r = slab / block;
  return r < P ? r : P - 1;
}

// PASS 0: count the particles born on every other rank; PASS 1: write their Lagrangian index and their
// local slot into the request list, grouped by birth rank
template <int PASS>
__global__ void __launch_bounds__(256)
k_sd_requests(size_t np, const float4 *__restrict__ pA, const float4 *__restrict__ pB, int ns, int N, int block, int P,
              int me, unsigned *__restrict__ cnt, const unsigned *__restrict__ off, unsigned *__restrict__ cursor,
              unsigned long long *__restrict__ req_id, uint32_t *__restrict__ req_slot) {
  const unsigned lane = threadIdx.x & 31;
  const size_t nround = (np + 31) / 32 * 32;
  const unsigned long long ns2 = (unsigned long long) ns * ns;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < nround; i += (size_t) gridDim.x * blockDim.x) {
    int r = me;
    unsigned long long id = 0;
    if (i < np) { id = particle_id(pA[i], pB[i]); r = birth_rank((long long) (id / ns2), ns, N, block, P); }
    const unsigned away = __ballot_sync(0xffffffffu, r != me);
    if (!away) continue;
    const unsigned peers = __match_any_sync(0xffffffffu, r);
    const int leader = __ffs(peers) - 1;
    if (PASS == 0) {
      if (r != me && (int) lane == leader) atomicAdd(&cnt[r], (unsigned) __popc(peers));
    } else {
      unsigned base = 0;
      if (r != me && (int) lane == leader) base = atomicAdd(&cursor[r], (unsigned) __popc(peers));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (r != me) {
        const size_t s = (size_t) off[r] + base + (unsigned) __popc(peers & ((1u << lane) - 1u));
        req_id[s] = id; req_slot[s] = (uint32_t) i;
      }
    }
  }
}

// the birth rank answers: resp[3 j + a] for request j
template <typename T>
__global__ void __launch_bounds__(256)
k_sd_serve(size_t nreq, const unsigned long long *__restrict__ req_id, int ns, int p0, int npl, int N, int NZ, int x0,
           int nx, const T *__restrict__ g0, const T *__restrict__ g1, const T *__restrict__ g2, double m0, double m1,
           double m2, float *__restrict__ resp) {
  const unsigned long long ns2 = (unsigned long long) ns * ns;
  for (size_t j = blockIdx.x * (size_t) blockDim.x + threadIdx.x; j < nreq; j += (size_t) gridDim.x * blockDim.x) {
    const unsigned long long id = req_id[j];
    long long n = (long long) (id / ns2);
    if (n < p0) n = p0;
    if (n >= p0 + npl) n = p0 + npl - 1;               // cannot happen: the requester computed the same owner
    const unsigned rem = (unsigned) (id % ns2);
    double r[3];
    sd_interp<T>(n, (int) (rem / ns), (int) (rem % ns), ns, N, NZ, x0, nx, g0, g1, g2, r);
    resp[3 * j] = sd_value<T>(r[0], m0); resp[3 * j + 1] = sd_value<T>(r[1], m1); resp[3 * j + 2] = sd_value<T>(r[2], m2);
  }
}

__global__ void __launch_bounds__(256)
k_sd_scatter(size_t nreq, const uint32_t *__restrict__ req_slot, const float *__restrict__ resp, float *__restrict__ out,
             size_t cap) {
  for (size_t j = blockIdx.x * (size_t) blockDim.x + threadIdx.x; j < nreq; j += (size_t) gridDim.x * blockDim.x) {
    const size_t i = req_slot[j];
    out[i] = resp[3 * j]; out[cap + i] = resp[3 * j + 1]; out[2 * cap + i] = resp[3 * j + 2];
  }
}

// cudaFree of a buffer NCCL has used costs up to hundreds of ms (peer mappings are torn down), so the request /
// response buffers start at `floor_bytes` (1/8 of the particle capacity) and double when they have to grow
static void sd_grow(void **p, size_t *have, size_t need, size_t floor_bytes) {
  if (*have >= need && *p) return;
  if (*p) CK(cudaFree(*p));
  size_t n = 2 * need + 1024;
  if (n < floor_bytes) n = floor_bytes;
  CK(cudaMalloc(p, n));
  *have = n;
}

// request lists for the current particle order (valid until the next sort / migration / upload)
static void sd_build_requests(Ctx &c) {
  if (c.P == 1 || c.sd_req_valid) return;
  const int P = c.P, me = c.rank, ns = c.cfg.nsample;
  const int block = (c.N + P - 1) / P;
  const size_t np = c.np;
  if (!c.sd_cnt_dev) {
    CK(cudaMalloc(&c.sd_cnt_dev, (size_t) (P * P + 3 * P) * sizeof(unsigned)));
    CK(cudaMallocHost(&c.sd_cnt_host, (size_t) (P * P + 3 * P) * sizeof(unsigned)));
  }
  unsigned *d_cnt = c.sd_cnt_dev, *d_all = d_cnt + P, *d_off = d_all + P * P, *d_cur = d_off + P;
  CK(cudaMemsetAsync(d_cnt, 0, P * sizeof(unsigned), c.stream));
  if (np)
    k_sd_requests<0><<<grid_for(np, 256), 256, 0, c.stream>>>(np, c.pA, c.pB, ns, c.N, block, P, me, d_cnt, nullptr, nullptr,
                                                            nullptr, nullptr);
  c.launches++;
  CKNCCL(ncclAllGather(d_cnt, d_all, P, ncclUint32, c.comm, c.stream));
  CK(cudaMemcpyAsync(c.sd_cnt_host, d_all, (size_t) P * P * sizeof(unsigned), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  const unsigned *all = c.sd_cnt_host;      // all[s * P + d]: s asks d for that many lattice points
  c.sd_need.assign(P, 0); c.sd_serve.assign(P, 0);
  size_t nneed = 0, nserve = 0;
  for (int q = 0; q < P; q++) {
    c.sd_need[q] = all[me * P + q]; c.sd_serve[q] = all[q * P + me];
    nneed += c.sd_need[q]; nserve += c.sd_serve[q];
  }
  c.sd_nneed = nneed; c.sd_nserve = nserve;
  const size_t fl = c.cap / 8 + 1024;
  sd_grow((void **) &c.sd_req_id, &c.sd_req_id_bytes, nneed * sizeof(unsigned long long), fl * sizeof(unsigned long long));
  sd_grow((void **) &c.sd_req_slot, &c.sd_req_slot_bytes, nneed * sizeof(uint32_t), fl * sizeof(uint32_t));
  sd_grow((void **) &c.sd_srv_id, &c.sd_srv_id_bytes, nserve * sizeof(unsigned long long), fl * sizeof(unsigned long long));
  sd_grow((void **) &c.sd_resp_out, &c.sd_resp_out_bytes, nserve * 3 * sizeof(float), fl * 3 * sizeof(float));
  sd_grow((void **) &c.sd_resp_in, &c.sd_resp_in_bytes, nneed * 3 * sizeof(float), fl * 3 * sizeof(float));
  std::vector<unsigned> off(P, 0);
  for (int q = 1; q < P; q++) off[q] = off[q - 1] + c.sd_need[q - 1];
  CK(cudaMemcpyAsync(d_off, off.data(), P * sizeof(unsigned), cudaMemcpyHostToDevice, c.stream));
  CK(cudaMemsetAsync(d_cur, 0, P * sizeof(unsigned), c.stream));
  if (nneed) {
    k_sd_requests<1><<<grid_for(np, 256), 256, 0, c.stream>>>(np, c.pA, c.pB, ns, c.N, block, P, me, nullptr, d_off, d_cur,
                                                            c.sd_req_id, c.sd_req_slot);
    c.launches++;
  }
  {
    PhaseTimer t(c, PH_COMM);
    CKNCCL(ncclGroupStart());
    size_t so = 0, ro = 0;
    for (int q = 0; q < P; q++) {
      if (c.sd_need[q]) CKNCCL(ncclSend(c.sd_req_id + so, (size_t) c.sd_need[q] * 8, ncclChar, q, c.comm, c.stream));
      if (c.sd_serve[q]) CKNCCL(ncclRecv(c.sd_srv_id + ro, (size_t) c.sd_serve[q] * 8, ncclChar, q, c.comm, c.stream));
      so += c.sd_need[q]; ro += c.sd_serve[q];
    }
    CKNCCL(ncclGroupEnd());
  }
  CK(cudaStreamSynchronize(c.stream));      // `off` is a host temporary
  c.sd_req_valid = true;
}

// ------------------------------------------------------------------ Lagrangian particles before main.c's init loop

// the reference assigns FIELD_D / FIELD_dDdy to P[coord] before positions exist (main.c:231-251): give the
// particle store its Lagrangian IDs so that the same call order works here
__global__ void k_sd_ids(size_t nloc, int ns, int p0, float4 *__restrict__ pA, float4 *__restrict__ pB) {
  for (size_t c = blockIdx.x * (size_t) blockDim.x + threadIdx.x; c < nloc; c += (size_t) gridDim.x * blockDim.x) {
    const unsigned long long id = (unsigned long long) p0 * ns * ns + c;
    pA[c] = make_float4(0.f, 0.f, 0.f, __uint_as_float((unsigned) (id & 0xffffffffull)));
    pB[c] = make_float4(0.f, 0.f, 0.f, __uint_as_float((unsigned) (id >> 32)));
  }
}

static void sd_ensure_particles(Ctx &c) {
  if (c.np != 0) return;
  const int ns = c.cfg.nsample;
  const size_t nloc = (size_t) c.npl * ns * ns;
  REQUIRE(nloc <= c.cap, MGP_ERR_BUFFER, "scale-dependent fields: particle capacity too small");
  if (nloc) k_sd_ids<<<grid_for(nloc, 256), 256, 0, c.stream>>>(nloc, ns, c.p0, c.pA, c.pB);
  c.launches++;
  c.np = nloc;
  c.sorted = false;
  c.bins_valid = false;
  c.drifts_since_sort = 1 << 30;
  c.sd_req_valid = false;
  c.sd_lagrangian_only = true;
}

// ------------------------------------------------------------------ assign_displacment_field_to_particles

template <typename T>
static void sd_assign_t(Ctx &c, int fieldtype, int order, const double *g1, const double *g2) {
  typedef typename Cpx<T>::type C;
  const int N = c.N, ns = c.cfg.nsample;
  const KL L = layout_of(c);
  const size_t mmax = (size_t) 3 * (N / 2) * (N / 2) + 1;
  const bool merged = order == 0;
  const double n3 = (double) N * (double) N * (double) N;
  const double norm2 = -3.0 / 7.0 / n3;                         // 2LPT.c:1556
  sd_ensure_particles(c);

  if (!c.sd_gtab[0]) { CK(cudaMalloc(&c.sd_gtab[0], mmax * sizeof(double))); CK(cudaMalloc(&c.sd_gtab[1], mmax * sizeof(double))); }
  CK(cudaMemcpyAsync(c.sd_gtab[0], g1, mmax * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  if (merged) CK(cudaMemcpyAsync(c.sd_gtab[1], g2, mmax * sizeof(double), cudaMemcpyHostToDevice, c.stream));

  // destination slot (2LPT.c:1817-1827): D / ddDddy -> D (order 1) or D2 (order 2); dDdy / deltaD -> dDdy or dD2dy
  const int pair = (fieldtype == MGP_FIELD_D || fieldtype == MGP_FIELD_DDDDDY) ? 0 : 2;
  const int slot = pair + ((merged || order == 1) ? 0 : 1);
  // merged fields stay resident in their grid block until Kick (pair 0, force grids) / Drift (pair 1, grids 0, 4, 5)
  // read them; separate orders share the force grids and go to the per-particle arrays at once
  static const bool lazy_env = !(getenv("MGP_SD_LAZY") && atoi(getenv("MGP_SD_LAZY")) == 0);
  REQUIRE(!c.forces_live, MGP_ERR_STATE, "scale-dependent fields: the force grids are in use (call mgp_mtoparticles first)");
  // pair 1 would live in grids 0, 4, 5: not while they hold this step's density (RSD multipoles inside PtoMesh)
  const bool lazy = merged && lazy_env && c.aux_block != nullptr && !c.sd_lagrangian_only && !(pair == 2 && c.density_live);
  const int block = (lazy && pair == 2) ? 1 : 0;
  c.sd_res[pair / 2] = -1;                                      // the old field of this pair is dead
  sd_evict_block(c, block);
  C *f[3]; const T *fr[3];
  for (int a = 0; a < 3; a++) { f[a] = (C *) c.grid[block_grid(block, a)]; fr[a] = (const T *) c.grid[block_grid(block, a)]; }
  {
    PhaseTimer t(c, PH_SDFIELD);
    const C *src1 = (const C *) c.sd_delta[(merged || order == 1) ? 0 : 1];
    const double nrm1 = (merged || order == 1) ? 1.0 : norm2;
    k_sd_field<T><<<grid_for(L.total, 256), 256, 0, c.stream>>>(L, src1, c.sd_gtab[0], nrm1,
                                                               merged ? (const C *) c.sd_delta[1] : nullptr, c.sd_gtab[1], norm2,
                                                               f[0], f[1], f[2], c.cfg.box);
    c.launches++;
  }
  fft_c2r_block(c, block);
  halo_fill_block(c, block);                                    // 2LPT.c:1638-1639

  const size_t nloc = (size_t) c.npl * ns * ns;
  double mean[3] = {0, 0, 0};
  // N == Nsample: the lattice sum is the k = 0 mode of the transform, which k_sd_field set to zero; the
  // reference subtracts the FFT's rounding noise (|mean| ~ 1e-17 of the rms).  Skipped in merged mode.
  // N == 2 Nsample: the sub-lattice sum aliases the modes with every component a multiple of Nsample = N/2, which are
  // self-conjugate and therefore real in k-space, so that i k / k^2 times them drops out of the c2r: zero as well.
  // Any other ratio (N = 3 Nsample ...) aliases modes the second-order field does populate: the mean is computed.
  const bool skip_mean = merged && (N == ns || N == 2 * ns);
  if (!skip_mean) {
    PhaseTimer t(c, PH_SDASSIGN);
    const unsigned gp = grid_for(nloc ? nloc : 1, 256, 8);
    reduce_alloc(c, (size_t) gp * 3 + 32);
    double *res = c.d_red + (size_t) gp * 3;
    k_sd_lattice_sum<T><<<gp, 256, 0, c.stream>>>(nloc, ns, c.p0, N, c.NZ, c.x0, c.nx, fr[0], fr[1], fr[2], c.d_red);
    k_final_reduce<<<1, 256, 0, c.stream>>>(c.d_red, (int) gp, 3, 3, 1.0, res);
    c.launches += 2;
    allreduce_sum(c, res, 3);
    CK(cudaMemcpyAsync(c.h_red, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    const double tot = (double) ns * (double) ns * (double) ns;
    for (int a = 0; a < 3; a++) mean[a] = c.h_red[a] / tot;     // sumdis_D /= TotNumPart (2LPT.c:1716-1719)
  }

  float *out = c.sdf[slot];
  if (!lazy) {
    PhaseTimer t(c, PH_SDASSIGN);
    const bool id32 = (double) ns * ns * ns < 4294967296.0;
    if (c.np && id32)
      k_sd_assign<T, 1><<<grid_for(c.np, 256), 256, 0, c.stream>>>(c.np, c.pA, c.pB, ns, c.p0, c.npl, N, c.NZ, c.x0, c.nx, fr[0],
                                                                  fr[1], fr[2], mean[0], mean[1], mean[2], out, c.cap);
    else if (c.np)
      k_sd_assign<T, 0><<<grid_for(c.np, 256), 256, 0, c.stream>>>(c.np, c.pA, c.pB, ns, c.p0, c.npl, N, c.NZ, c.x0, c.nx, fr[0],
                                                                  fr[1], fr[2], mean[0], mean[1], mean[2], out, c.cap);
    c.launches++;
  }
  if (c.P > 1) {
    sd_build_requests(c);
    PhaseTimer t(c, PH_SDASSIGN);
    if (c.sd_nserve) {
      k_sd_serve<T><<<grid_for(c.sd_nserve, 256), 256, 0, c.stream>>>(c.sd_nserve, c.sd_srv_id, ns, c.p0, c.npl, N, c.NZ, c.x0,
                                                                     c.nx, fr[0], fr[1], fr[2], mean[0], mean[1], mean[2],
                                                                     c.sd_resp_out);
      c.launches++;
    }
    {
      PhaseTimer tc(c, PH_COMM);
      CKNCCL(ncclGroupStart());
      size_t so = 0, ro = 0;
      for (int q = 0; q < c.P; q++) {
        if (c.sd_serve[q]) CKNCCL(ncclSend(c.sd_resp_out + 3 * so, (size_t) c.sd_serve[q] * 12, ncclChar, q, c.comm, c.stream));
        if (c.sd_need[q]) CKNCCL(ncclRecv(c.sd_resp_in + 3 * ro, (size_t) c.sd_need[q] * 12, ncclChar, q, c.comm, c.stream));
        so += c.sd_serve[q]; ro += c.sd_need[q];
      }
      CKNCCL(ncclGroupEnd());
    }
    if (c.sd_nneed) {
      k_sd_scatter<<<grid_for(c.sd_nneed, 256), 256, 0, c.stream>>>(c.sd_nneed, c.sd_req_slot, c.sd_resp_in, out, c.cap);
      c.launches++;
    }
  }
  c.sd_zero[slot] = false;
  if (merged) c.sd_zero[slot + 1] = true;                        // the second-order slot reads as 0
  c.sd_set[slot] = true;
  if (merged) c.sd_set[slot + 1] = true;
  if (lazy) {
    c.sd_res[pair / 2] = block;
    for (int a = 0; a < 3; a++) c.sd_res_mean[pair / 2][a] = mean[a];
  }
}

// resident merged field -> the per-particle array of its first-order slot (the remote-born entries are there already)
template <typename T>
static void sd_materialise_t(Ctx &c, int pr) {
  const int block = c.sd_res[pr], ns = c.cfg.nsample;
  const T *fr[3];
  for (int a = 0; a < 3; a++) fr[a] = (const T *) c.grid[block_grid(block, a)];
  const double *m = c.sd_res_mean[pr];
  PhaseTimer t(c, PH_SDASSIGN);
  const bool id32 = (double) ns * ns * ns < 4294967296.0;
  if (c.np && id32)
    k_sd_assign<T, 1><<<grid_for(c.np, 256), 256, 0, c.stream>>>(c.np, c.pA, c.pB, ns, c.p0, c.npl, c.N, c.NZ, c.x0, c.nx, fr[0], fr[1],
                                                                fr[2], m[0], m[1], m[2], c.sdf[2 * pr], c.cap);
  else if (c.np)
    k_sd_assign<T, 0><<<grid_for(c.np, 256), 256, 0, c.stream>>>(c.np, c.pA, c.pB, ns, c.p0, c.npl, c.N, c.NZ, c.x0, c.nx, fr[0], fr[1],
                                                                fr[2], m[0], m[1], m[2], c.sdf[2 * pr], c.cap);
  c.launches++;
  c.sd_res[pr] = -1;
}

void sd_materialise(Ctx &c, int pr) {
  if (pr < 0 || pr > 1 || c.sd_res[pr] < 0) return;
  if (c.gbytes == 4) sd_materialise_t<float>(c, pr); else sd_materialise_t<double>(c, pr);
}

void sd_evict_block(Ctx &c, int block) {
  for (int pr = 0; pr < 2; pr++) if (c.sd_res[pr] == block) sd_materialise(c, pr);
}

void sd_drop(Ctx &c) {
  c.sd_res[0] = c.sd_res[1] = -1;
  for (int s = 0; s < 4; s++) c.sd_set[s] = false;
}

void sd_assign(Ctx &c, int fieldtype, int order, const double *g1, const double *g2, size_t n) {
  REQUIRE(c.cfg.scale_dependent, MGP_ERR_STATE, "scale-dependent fields need mgp_config.scale_dependent = 1");
  REQUIRE(fieldtype >= MGP_FIELD_D && fieldtype <= MGP_FIELD_DELTAD, MGP_ERR_INVALID, "unknown field type");
  REQUIRE(order >= 0 && order <= 2, MGP_ERR_INVALID, "LPT order must be 1 or 2");
  const size_t mmax = (size_t) 3 * (c.N / 2) * (c.N / 2) + 1;
  REQUIRE(g1 != nullptr && n >= mmax && (order != 0 || g2 != nullptr), MGP_ERR_INVALID,
          "growth table must hold 3 (Nmesh/2)^2 + 1 entries");
  REQUIRE(c.sd_have_delta, MGP_ERR_STATE, "scale-dependent fields: no stored delta_k (call mgp_ic_generate first)");
  if (c.gbytes == 4) sd_assign_t<float>(c, fieldtype, order, g1, g2); else sd_assign_t<double>(c, fieldtype, order, g1, g2);
}

// ------------------------------------------------------------------ particle initialisation, SCALEDEPENDENT branch

// main.c:257-304: Vel = dDdy + dD2dy (float add) or 0; Pos = wrap(q + D + D2) with q in double
__global__ void __launch_bounds__(256)
k_sd_init_particles(size_t np, int ns, double box, int use_cola, const float *__restrict__ D, const float *__restrict__ D2,
                    const float *__restrict__ V1, const float *__restrict__ V2, size_t cap, float4 *__restrict__ pA,
                    float4 *__restrict__ pB) {
  const float boxf = (float) box;
  const double dq = box / (double) ns;
  const unsigned long long ns2 = (unsigned long long) ns * ns;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < np; i += (size_t) gridDim.x * blockDim.x) {
    float4 a = pA[i], b = pB[i];
    const unsigned long long id = particle_id(a, b);
    const long long n = (long long) (id / ns2);
    const unsigned rem = (unsigned) (id % ns2);
    const double q[3] = {(double) n * dq, (double) (rem / ns) * dq, (double) (rem % ns) * dq};
    float X[3], V[3];
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      const float d1 = D[ax * cap + i], d2 = D2 ? D2[ax * cap + i] : 0.0f;
      V[ax] = use_cola ? 0.0f : __fadd_rn(V1[ax * cap + i], V2 ? V2[ax * cap + i] : 0.0f);
      float x = (float) __dadd_rn(__dadd_rn(q[ax], (double) d1), (double) d2);
      while (x >= boxf) x -= boxf;
      while (x < 0) x += boxf;
      if (x == boxf) x = 0.0f;
      X[ax] = x;
    }
    a.x = X[0]; a.y = X[1]; a.z = X[2];
    b.x = V[0]; b.y = V[1]; b.z = V[2];
    pA[i] = a; pB[i] = b;
  }
}

void sd_init_particles(Ctx &c) {
  REQUIRE(c.sd_lagrangian_only && c.np, MGP_ERR_STATE, "mgp_init_particles (scale-dependent): assign FIELD_D first");
  REQUIRE(c.sd_set[0] && c.sd_set[1], MGP_ERR_STATE, "mgp_init_particles (scale-dependent): FIELD_D of both orders must be assigned");
  if (!c.cfg.use_cola)
    REQUIRE(c.sd_set[2] && c.sd_set[3], MGP_ERR_STATE, "mgp_init_particles (scale-dependent, no COLA): FIELD_dDdy must be assigned");
  k_sd_init_particles<<<grid_for(c.np, 256), 256, 0, c.stream>>>(c.np, c.cfg.nsample, c.cfg.box, c.cfg.use_cola, c.sdf[0],
                                                               c.sd_zero[1] ? nullptr : c.sdf[1], c.sdf[2],
                                                               c.sd_zero[3] ? nullptr : c.sdf[3], c.cap, c.pA, c.pB);
  c.launches++;
  CK(cudaStreamSynchronize(c.stream));
  c.sd_lagrangian_only = false;
  c.sorted = false;
  c.bins_valid = false;
  c.drifts_since_sort = 1 << 30;
  c.have_disp = false;
}

// ------------------------------------------------------------------ Kick / Drift, SCALEDEPENDENT branches

// main.c:705-717: force = -1.5 Omega Disp - UseCOLA (D + D2) / A   with D + D2 and UseCOLA * (.) in float
__global__ void __launch_bounds__(256)
k_kick_sd(size_t n, float4 *__restrict__ pB, const float *__restrict__ D, const float *__restrict__ D2,
          float *__restrict__ disp, size_t cap, double sDx, double sDy, double sDz, double m15omega, float usecola,
          double A, double dda, double *__restrict__ partial) {
  double s[3] = {0, 0, 0};
  const double sD[3] = {sDx, sDy, sDz};
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    float4 v = pB[i];
    float vel[3] = {v.x, v.y, v.z};
    const float dsp[3] = {disp[i], disp[cap + i], disp[2 * cap + i]};      // all loads before the first store
    const float d1[3] = {D[i], D[cap + i], D[2 * cap + i]};
    const float d2[3] = {D2 ? D2[i] : 0.0f, D2 ? D2[cap + i] : 0.0f, D2 ? D2[2 * cap + i] : 0.0f};
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      const float g = (float) __dsub_rn((double) dsp[ax], sD[ax]);
      disp[ax * cap + i] = g;
      const float dd = __fmul_rn(usecola, __fadd_rn(d1[ax], d2[ax]));
      const double f = __dsub_rn(__dmul_rn(m15omega, (double) g), __ddiv_rn((double) dd, A));
      vel[ax] = (float) __dadd_rn((double) vel[ax], __dmul_rn(f, dda));
      s[ax] += (double) vel[ax];
    }
    v.x = vel[0]; v.y = vel[1]; v.z = vel[2];
    pB[i] = v;
  }
  block_sum3(s[0], s[1], s[2]);
  if (threadIdx.x == 0) { partial[3 * blockIdx.x] = s[0]; partial[3 * blockIdx.x + 1] = s[1]; partial[3 * blockIdx.x + 2] = s[2]; }
}

// per-particle value of a resident merged field: interpolated at the Lagrangian point of the particle's ID when that
// point is on this rank, else the value its birth rank sent (already scattered into `remote`)
struct SdRes {
  const void *g0, *g1, *g2;
  const float *remote;
  double m0, m1, m2;
  int ns, p0, npl, N, NZ, x0, nx;
};

// EXACT: Nmesh is a multiple of Nsample, every lattice point is a mesh point and the read-out is one load per grid
template <typename T, int ID32, int EXACT>
__device__ __forceinline__ void sd_resident_value(const SdRes &R, unsigned idlo, unsigned idhi, size_t i, size_t cap, float D[3]) {
  const unsigned long long ns2 = (unsigned long long) R.ns * R.ns;
  const unsigned long long id = ID32 ? (unsigned long long) idlo : (((unsigned long long) idhi << 32) | idlo);
  const long long n = (long long) (id / ns2);
  if (n < R.p0 || n >= R.p0 + R.npl) {
    D[0] = R.remote[i]; D[1] = R.remote[cap + i]; D[2] = R.remote[2 * cap + i];
    return;
  }
  const unsigned rem = (unsigned) (id - (unsigned long long) n * ns2);
  double r[3];
  if (EXACT) {
    const int f = R.N / R.ns;
    const size_t e = (((size_t) ((int) n * f - R.x0)) * R.N + (size_t) (rem / R.ns) * f) * (size_t) (2 * R.NZ) + (size_t) (rem % R.ns) * f;
    r[0] = (double) ((const T *) R.g0)[e]; r[1] = (double) ((const T *) R.g1)[e]; r[2] = (double) ((const T *) R.g2)[e];
  } else {
    sd_interp<T>(n, (int) (rem / R.ns), (int) (rem % R.ns), R.ns, R.N, R.NZ, R.x0, R.nx, (const T *) R.g0, (const T *) R.g1,
                 (const T *) R.g2, r);
  }
  D[0] = sd_value<T>(r[0], R.m0); D[1] = sd_value<T>(r[1], R.m1); D[2] = sd_value<T>(r[2], R.m2);
}

static SdRes sd_res_of(const Ctx &c, int pr) {
  SdRes R;
  const int block = c.sd_res[pr];
  R.g0 = c.grid[block_grid(block, 0)]; R.g1 = c.grid[block_grid(block, 1)]; R.g2 = c.grid[block_grid(block, 2)];
  R.remote = c.sdf[2 * pr];
  R.m0 = c.sd_res_mean[pr][0]; R.m1 = c.sd_res_mean[pr][1]; R.m2 = c.sd_res_mean[pr][2];
  R.ns = c.cfg.nsample; R.p0 = c.p0; R.npl = c.npl; R.N = c.N; R.NZ = c.NZ; R.x0 = c.x0; R.nx = c.nx;
  return R;
}

// Kick with the D + D2 field read straight from the grids it was transformed in (no per-particle copy)
template <typename T, int ID32, int EXACT, int BL>
__global__ void __launch_bounds__(256, BL)
k_kick_sd_fused(size_t n, const float4 *__restrict__ pA, float4 *__restrict__ pB, SdRes R, float *__restrict__ disp, size_t cap,
                double sDx, double sDy, double sDz, double m15omega, float usecola, double A, double dda,
                double *__restrict__ partial) {
  double s[3] = {0, 0, 0};
  const double sD[3] = {sDx, sDy, sDz};
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    float4 v = pB[i];
    const unsigned idlo = __float_as_uint(pA[i].w);
    const float dsp[3] = {disp[i], disp[cap + i], disp[2 * cap + i]};      // all loads before the first store
    float D[3];
    sd_resident_value<T, ID32, EXACT>(R, idlo, __float_as_uint(v.w), i, cap, D);
    float vel[3] = {v.x, v.y, v.z};
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      const float g = (float) __dsub_rn((double) dsp[ax], sD[ax]);
      disp[ax * cap + i] = g;
      const float dd = __fmul_rn(usecola, __fadd_rn(D[ax], 0.0f));
      const double f = __dsub_rn(__dmul_rn(m15omega, (double) g), __ddiv_rn((double) dd, A));
      vel[ax] = (float) __dadd_rn((double) vel[ax], __dmul_rn(f, dda));
      s[ax] += (double) vel[ax];
    }
    v.x = vel[0]; v.y = vel[1]; v.z = vel[2];
    pB[i] = v;
  }
  block_sum3(s[0], s[1], s[2]);
  if (threadIdx.x == 0) { partial[3 * blockIdx.x] = s[0]; partial[3 * blockIdx.x + 1] = s[1]; partial[3 * blockIdx.x + 2] = s[2]; }
}

template <typename T, int ID32, int EXACT>
__global__ void __launch_bounds__(256, EXACT ? MGP_SD_BLOCKS : 4)
k_drift_sd_fused(size_t n, float4 *__restrict__ pA, const float4 *__restrict__ pB, SdRes R, size_t cap, double sVx, double sVy,
                 double sVz, double dyyy, float usecola, float boxf) {
  const double sV[3] = {sVx, sVy, sVz};
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    float4 p = pA[i];
    const float4 v = pB[i];
    float D[3];
    sd_resident_value<T, ID32, EXACT>(R, __float_as_uint(p.w), __float_as_uint(v.w), i, cap, D);
    float x[3] = {p.x, p.y, p.z};
    const float vel[3] = {v.x, v.y, v.z};
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      float t = (float) __dadd_rn((double) x[ax], __dmul_rn(__dsub_rn((double) vel[ax], sV[ax]), dyyy));
      t = __fadd_rn(t, __fmul_rn(usecola, __fadd_rn(D[ax], 0.0f)));
      while (t >= boxf) t -= boxf;
      while (t < 0) t += boxf;
      if (t == boxf) t = 0.0f;
      x[ax] = t;
    }
    p.x = x[0]; p.y = x[1]; p.z = x[2];
    pA[i] = p;
  }
}

void sd_kick(Ctx &c, double A, double dda, const double sumD[3], double sumV[3]) {
  REQUIRE(c.sd_set[0] && c.sd_set[1], MGP_ERR_STATE, "mgp_kick (scale-dependent): assign FIELD_ddDddy first");
  const size_t n = c.np;
  const unsigned g = grid_for(n, 256, 8);
  reduce_alloc(c, (size_t) g * 3 + 16);
  double *res = c.d_red + (size_t) g * 3;
  if (c.sd_res[0] >= 0) {
    const SdRes R = sd_res_of(c, 0);
    const bool id32 = (double) R.ns * R.ns * R.ns < 4294967296.0;
    const double m15 = -1.5 * c.cfg.omega;
    const float uc = (float) c.cfg.use_cola;
    const bool exact = c.N % R.ns == 0;
    static const int kb = getenv("MGP_KICK_BLOCKS") ? atoi(getenv("MGP_KICK_BLOCKS")) : 4;     // developer knob
#define KICK_FUSED(T, I, E, B) k_kick_sd_fused<T, I, E, B><<<g, 256, 0, c.stream>>>(n, c.pA, c.pB, R, c.disp, c.cap, sumD[0], sumD[1], sumD[2], m15, uc, A, dda, c.d_red)
#define KICK_FUSED_T(T) do { if (id32 && exact) { if (kb == 5) KICK_FUSED(T, 1, 1, 5); else if (kb == 4) KICK_FUSED(T, 1, 1, 4); else KICK_FUSED(T, 1, 1, 6); } \
                             else if (id32) KICK_FUSED(T, 1, 0, 4); else if (exact) KICK_FUSED(T, 0, 1, 5); else KICK_FUSED(T, 0, 0, 4); } while (0)
    if (c.gbytes == 4) KICK_FUSED_T(float); else KICK_FUSED_T(double);
#undef KICK_FUSED_T
#undef KICK_FUSED
  } else
  k_kick_sd<<<g, 256, 0, c.stream>>>(n, c.pB, c.sdf[0], c.sd_zero[1] ? nullptr : c.sdf[1], c.disp, c.cap, sumD[0], sumD[1],
                                     sumD[2], -1.5 * c.cfg.omega, (float) c.cfg.use_cola, A, dda, c.d_red);
  k_final_reduce<<<1, 256, 0, c.stream>>>(c.d_red, (int) g, 3, 3, 1.0, res);
  c.launches += 2;
  allreduce_sum(c, res, 3);
  CK(cudaMemcpyAsync(c.h_red, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  const double tot = (double) c.cfg.nsample * (double) c.cfg.nsample * (double) c.cfg.nsample;
  for (int a = 0; a < 3; a++) sumV[a] = c.h_red[a] / tot;
}

// main.c:760-770: Pos += (Vel - sumxyz) dyyy; Pos = wrap(Pos + UseCOLA (dDdy + dD2dy))   second line all float
__global__ void __launch_bounds__(256)
k_drift_sd(size_t n, float4 *__restrict__ pA, const float4 *__restrict__ pB, const float *__restrict__ V1,
           const float *__restrict__ V2, size_t cap, double sVx, double sVy, double sVz, double dyyy, float usecola,
           float boxf) {
  const double sV[3] = {sVx, sVy, sVz};
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    float4 p = pA[i];
    const float4 v = pB[i];
    float x[3] = {p.x, p.y, p.z};
    const float vel[3] = {v.x, v.y, v.z};
#pragma unroll
    for (int ax = 0; ax < 3; ax++) {
      float t = (float) __dadd_rn((double) x[ax], __dmul_rn(__dsub_rn((double) vel[ax], sV[ax]), dyyy));
      t = __fadd_rn(t, __fmul_rn(usecola, __fadd_rn(V1[ax * cap + i], V2 ? V2[ax * cap + i] : 0.0f)));
      while (t >= boxf) t -= boxf;
      while (t < 0) t += boxf;
      if (t == boxf) t = 0.0f;
      x[ax] = t;
    }
    p.x = x[0]; p.y = x[1]; p.z = x[2];
    pA[i] = p;
  }
}

void sd_drift(Ctx &c, double dyyy, const double sumV[3]) {
  REQUIRE(c.sd_set[2] && c.sd_set[3], MGP_ERR_STATE, "mgp_drift (scale-dependent): assign FIELD_deltaD first");
  const size_t n = c.np;
  if (n && c.sd_res[1] >= 0) {
    const SdRes R = sd_res_of(c, 1);
    const bool id32 = (double) R.ns * R.ns * R.ns < 4294967296.0;
    const float uc = (float) c.cfg.use_cola, bx = (float) c.cfg.box;
    const unsigned g = grid_for(n, 256);
    const bool exact = c.N % R.ns == 0;
#define DRIFT_FUSED(T, I, E) k_drift_sd_fused<T, I, E><<<g, 256, 0, c.stream>>>(n, c.pA, c.pB, R, c.cap, sumV[0], sumV[1], sumV[2], dyyy, uc, bx)
#define DRIFT_FUSED_T(T) do { if (id32 && exact) DRIFT_FUSED(T, 1, 1); else if (id32) DRIFT_FUSED(T, 1, 0); else if (exact) DRIFT_FUSED(T, 0, 1); else DRIFT_FUSED(T, 0, 0); } while (0)
    if (c.gbytes == 4) DRIFT_FUSED_T(float); else DRIFT_FUSED_T(double);
#undef DRIFT_FUSED_T
#undef DRIFT_FUSED
    c.launches++;
  } else if (n) {
    k_drift_sd<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.pA, c.pB, c.sdf[2], c.sd_zero[3] ? nullptr : c.sdf[3], c.cap, sumV[0],
                                                      sumV[1], sumV[2], dyyy, (float) c.cfg.use_cola, (float) c.cfg.box);
    c.launches++;
  }
}

// ------------------------------------------------------------------ host access to the four fields

void sd_alloc(Ctx &c) {
  for (int s = 0; s < 4; s++) {
    CK(cudaMalloc(&c.sdf[s], 3 * c.cap * sizeof(float)));
    CK(cudaMemsetAsync(c.sdf[s], 0, 3 * c.cap * sizeof(float), c.stream));
  }
}

void sd_free(Ctx &c) {
  for (int s = 0; s < 4; s++) cudaFree(c.sdf[s]);
  cudaFree(c.sd_gtab[0]); cudaFree(c.sd_gtab[1]);
  cudaFree(c.sd_cnt_dev); if (c.sd_cnt_host) cudaFreeHost(c.sd_cnt_host);
  cudaFree(c.sd_req_id); cudaFree(c.sd_req_slot); cudaFree(c.sd_srv_id); cudaFree(c.sd_resp_out); cudaFree(c.sd_resp_in);
}

// host [n][3] <-> device [3][cap]
void sd_copy_field(Ctx &c, int slot, float *host, bool to_host) {
  REQUIRE(slot >= 0 && slot < 4 && c.sdf[slot], MGP_ERR_STATE, "scale-dependent field storage missing");
  if (to_host) sd_materialise(c, slot / 2); else c.sd_res[slot / 2] = -1;
  if (to_host && c.sd_zero[slot]) memset(host, 0, c.np * 3 * sizeof(float));
  else copy_soa3(c, c.sdf[slot], c.np, host, to_host);
  if (!to_host) { c.sd_zero[slot] = false; c.sd_set[slot] = true; }
}

}  // namespace mgp
