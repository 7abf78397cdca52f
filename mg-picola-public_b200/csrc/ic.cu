// Initial conditions on the GPU: Gaussian random field + 2LPT displacement fields and the particle
// initialisation loop.  Replaces displacement_fields() (2LPT.c:185-495, 1204-1520, Gaussian branch)
// and main.c:257-309.
//
// Random numbers.  The reference draws every (i, j) row of Fourier modes from its own stream of
// GSL's ranlxd1 generator, seeded from a table that is itself filled from ranlxd1(Seed) in the
// N-GenIC order (2LPT.c:259-271, 344-353).  The generator is restated here with 48-bit integers
// (Luescher's subtract-with-borrow, lags 12 / 5, luxury 202; the GSL double arithmetic is exact, so
// the integer form is bit-identical); the known-answer values of GSL's own test suite pin it
// (tests/test_ic.py).  One GPU thread owns one row: N^2 independent streams.
//
// Memory.  Only the scalar field delta_k is stored (one complex grid); the Zel'dovich field
// psi_a = i k_a / k^2 delta, its six gradients and the second-order field are formed on the fly in
// the three force grids, so the whole 13-FFT pipeline needs one grid more than a time step does.
#include "common.cuh"
#include "klayout.cuh"
#include "reduce.cuh"
#include "ic_modes.cuh"

#include <cmath>

namespace mgp {

// ------------------------------------------------------------------ ranlxd1

struct Ranlxd {
  long long x[12];          // 48-bit fractions
  unsigned ir, jr, ir_old;
  long long carry;
};

__host__ __device__ inline void ranlxd_set(Ranlxd &r, unsigned long s) {
  if (s == 0) s = 1;
  unsigned bits = (unsigned) (s & 0x7FFFFFFFUL);     // xbit[k] = bit k
  int ibit = 0, jbit = 18;
  for (int k = 0; k < 12; k++) {
    long long x = 0;
    for (int l = 1; l <= 48; l++) {
      const unsigned bi = (bits >> ibit) & 1u, bj = (bits >> jbit) & 1u;
      x = 2 * x + (long long) (bi ^ 1u);
      bits = (bits & ~(1u << ibit)) | (((bi + bj) & 1u) << ibit);
      ibit = ibit == 30 ? 0 : ibit + 1;
      jbit = jbit == 30 ? 0 : jbit + 1;
    }
    r.x[k] = x;
  }
  r.carry = 0; r.ir = 11; r.jr = 7; r.ir_old = 0;
}

__host__ __device__ inline double ranlxd_uniform(Ranlxd &r) {
  r.ir = r.ir == 11 ? 0 : r.ir + 1;
  if (r.ir == r.ir_old) {
    unsigned ir = r.ir, jr = r.jr;
    long long carry = r.carry;
    for (int k = 0; k < 202; k++) {                   // luxury level of ranlxd1
      long long y = r.x[jr] - r.x[ir] - carry;
      if (y < 0) { carry = 1; y += (1LL << 48); } else carry = 0;
      r.x[ir] = y;
      ir = ir == 11 ? 0 : ir + 1;
      jr = jr == 11 ? 0 : jr + 1;
    }
    r.ir = ir; r.ir_old = ir; r.jr = jr; r.carry = carry;
  }
  return (double) r.x[r.ir] * (1.0 / 281474976710656.0);
}

// seedtable in the order of 2LPT.c:259-271
static void make_seedtable(unsigned seed, int N, std::vector<unsigned> &t) {
  t.assign((size_t) N * N, 0u);
  Ranlxd r;
  ranlxd_set(r, seed);
  auto draw = [&]() { return (unsigned) (0x7fffffff * ranlxd_uniform(r)); };
  for (int i = 0; i < N / 2; i++) {
    int j;
    for (j = 0; j < i; j++) t[(size_t) i * N + j] = draw();
    for (j = 0; j < i + 1; j++) t[(size_t) j * N + i] = draw();
    for (j = 0; j < i; j++) t[(size_t) (N - 1 - i) * N + j] = draw();
    for (j = 0; j < i + 1; j++) t[(size_t) (N - 1 - j) * N + i] = draw();
    for (j = 0; j < i; j++) t[(size_t) i * N + (N - 1 - j)] = draw();
    for (j = 0; j < i + 1; j++) t[(size_t) j * N + (N - 1 - i)] = draw();
    for (j = 0; j < i; j++) t[(size_t) (N - 1 - i) * N + (N - 1 - j)] = draw();
    for (j = 0; j < i + 1; j++) t[(size_t) (N - 1 - j) * N + (N - 1 - i)] = draw();
  }
}

// ------------------------------------------------------------------ delta_k

__device__ __forceinline__ bool kl_encode(const KL &L, int i, int j, int k, size_t &e) {
  if (!L.transposed) { e = ((size_t) i * L.N + j) * L.NZ + k; return true; }
  const int jl = j - L.j0;
  if (jl < 0 || jl >= L.nyl) return false;
  e = ((size_t) jl * L.NZ + k) * L.N + i;
  return true;
}

// One thread per (i, j) row of modes (2LPT.c:337-495).  power[m] = P(k) at integer |d|^2 = m,
// already multiplied by whatever the driver applies (mg_pofk_ratio, sigma8 ratio; 2LPT.c:388-407).
template <typename T>
__global__ void __launch_bounds__(128)
k_ic_delta(KL L, const unsigned *__restrict__ seedtable, const double *__restrict__ power, int nsample, double box,
           int sphere_mode, int amplitude_fixed, int inverted, typename Cpx<T>::type *__restrict__ dk) {
  typedef typename Cpx<T>::type C;
  const int N = L.N, h = N / 2;
  const long long row = blockIdx.x * (long long) blockDim.x + threadIdx.x;
  if (row >= (long long) N * N) return;
  const int i = (int) (row / N), j = (int) (row % N);
  const int ii = i == 0 ? 0 : N - i, jj = j == 0 ? 0 : N - j;
  if (L.transposed) {     // this rank stores ky in [j0, j0 + nyl): the row itself and its k = 0 conjugate row
    const bool a = j >= L.j0 && j < L.j0 + L.nyl, b = jj >= L.j0 && jj < L.j0 + L.nyl;
    if (!a && !b) return;
  }
  Ranlxd rng;
  ranlxd_set(rng, seedtable[row]);
  const double PI = 3.14159265358979323846;
  const double fac = pow(box, -1.5);
  for (int k = 0; k < h; k++) {
    const double phase = ranlxd_uniform(rng) * 2 * PI;
    double ampl = 0.0;
    if (!amplitude_fixed) {
      do { ampl = ranlxd_uniform(rng); } while (ampl == 0);
    }
    if (i == h || j == h || k == h) continue;
    if (i == 0 && j == 0 && k == 0) continue;
    const int d0 = i < h ? i : i - N, d1 = j < h ? j : j - N, d2 = k;
    const double kv0 = d0 * 2 * PI / box, kv1 = d1 * 2 * PI / box, kv2 = d2 * 2 * PI / box;
    const double kmag2 = kv0 * kv0 + kv1 * kv1 + kv2 * kv2;
    const double kmag = sqrt(kmag2);
    if (sphere_mode == 1) {
      if (kmag * box / (2 * PI) > nsample / 2) continue;
    } else {
      if (fabs(kv0) * box / (2 * PI) > nsample / 2) continue;
      if (fabs(kv1) * box / (2 * PI) > nsample / 2) continue;
      if (fabs(kv2) * box / (2 * PI) > nsample / 2) continue;
    }
    double p_of_k = power[(long long) d0 * d0 + (long long) d1 * d1 + (long long) d2 * d2];
    if (!amplitude_fixed) p_of_k *= -log(ampl);
    double delta = fac * sqrt(p_of_k);
    if (inverted) delta *= -1.0;
    const double re = delta * cos(phase), im = delta * sin(phase);
    size_t e;
    C v, vc;
    v.x = (T) re; v.y = (T) im; vc.x = (T) re; vc.y = (T) -im;
    if (k > 0) {
      if (kl_encode(L, i, j, k, e)) dk[e] = v;
    } else if (i == 0) {
      if (j >= h) continue;
      if (kl_encode(L, i, j, k, e)) dk[e] = v;
      if (kl_encode(L, i, jj, k, e)) dk[e] = vc;       // Psi(i,j,k) = Psi*(-i,-j,-k)
    } else {
      if (i >= h) continue;
      if (kl_encode(L, i, j, k, e)) dk[e] = v;
      if (kl_encode(L, ii, jj, k, e)) dk[e] = vc;
    }
  }
}

// ------------------------------------------------------------------ k-space kernels of the 2LPT pipeline

// the arithmetic per mode is in ic_modes.cuh (shared with the host emulation); `ext`: delta_k comes from external particles
// (READICFROMFILE), whose Nyquist planes carry power and follow AssignDisplacementField's wave-vector convention
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_ic_kernel(KL L, const typename Cpx<T>::type *__restrict__ src, typename Cpx<T>::type *__restrict__ o0,
            typename Cpx<T>::type *__restrict__ o1, typename Cpx<T>::type *__restrict__ o2, double box,
            const double *__restrict__ gtab, double norm, int ext) {
  typedef typename Cpx<T>::type C;
  KLOOP(e, L) {
    int i, j, k;
    kl_decode(L, e, i, j, k);
    C out[3];
    ic_mode<T, C, MODE>(L.N, i, j, k, box, src[e], gtab, norm, ext, out);
    o0[e] = out[0]; o1[e] = out[1]; o2[e] = out[2];
  }
}

// second-order source, two passes over real space (2LPT.c:1282-1290):
//   pass 0: S  = g00 (g11 + g22) + g11 g22         pass 1: S = S - g01^2 - g02^2 - g12^2
template <typename T>
__global__ void k_ic_source(T *__restrict__ S, const T *__restrict__ a, const T *__restrict__ b, const T *__restrict__ cc,
                            size_t n, int pass) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    const T x = a[i], y = b[i], z = cc[i];
    if (pass == 0) S[i] = x * (y + z) + y * z;
    else S[i] = S[i] - x * x - y * y - z * z;
  }
}

// stored second-order scalar: delta2_k = -S_k, zero mode cleared (2LPT.c:1336-1344)
template <typename C>
__global__ void k_ic_neg(const C *__restrict__ s, C *__restrict__ o, size_t n, size_t zero_index) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    C v = s[i];
    v.x = -v.x; v.y = -v.y;
    if (i == zero_index) { v.x = 0; v.y = 0; }
    o[i] = v;
  }
}

// ------------------------------------------------------------------ read-out at the Lagrangian points

// trilinear read-out of three real grids at q = (n + p0, m, p) * Nmesh / Nsample (2LPT.c:1419-1488)
// out[axis * stride + coord] = value * scale (float);  partial[3 * block + axis] = sum (double)
template <typename T>
__global__ void __launch_bounds__(256)
k_ic_readout(size_t nloc, int ns, int p0, int N, int NZ, int x0, int nx, const T *__restrict__ g0, const T *__restrict__ g1,
             const T *__restrict__ g2, double scale, float *__restrict__ out, size_t stride, double *__restrict__ partial) {
  double s0 = 0, s1 = 0, s2 = 0;
  const size_t rz = (size_t) 2 * NZ;
  for (size_t c = blockIdx.x * (size_t) blockDim.x + threadIdx.x; c < nloc; c += (size_t) gridDim.x * blockDim.x) {
    const int p = (int) (c % ns);
    const size_t t = c / ns;
    const int m = (int) (t % ns), n = (int) (t / ns);
    double u = (double) ((long long) (n + p0) * N) / (double) ns;
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
    const double f1 = (1 - u) * (1 - v) * (1 - w), f2 = (1 - u) * (1 - v) * w, f3 = (1 - u) * v * (1 - w), f4 = (1 - u) * v * w;
    const double f5 = u * (1 - v) * (1 - w), f6 = u * (1 - v) * w, f7 = u * v * (1 - w), f8 = u * v * w;
    const size_t a = ((size_t) i * N + j) * rz, b = ((size_t) i * N + j2) * rz, cc = ((size_t) i2 * N + j) * rz,
                 d = ((size_t) i2 * N + j2) * rz;
#define RD(G)                                                                                                     \
  ((double) G[a + k] * f1 + (double) G[a + k2] * f2 + (double) G[b + k] * f3 + (double) G[b + k2] * f4 +        \
   (double) G[cc + k] * f5 + (double) G[cc + k2] * f6 + (double) G[d + k] * f7 + (double) G[d + k2] * f8)
    const double r0 = RD(g0) * scale, r1 = RD(g1) * scale, r2 = RD(g2) * scale;
#undef RD
    out[c] = (float) r0; out[stride + c] = (float) r1; out[2 * stride + c] = (float) r2;
    s0 += r0; s1 += r1; s2 += r2;
  }
  block_sum3(s0, s1, s2);
  if (threadIdx.x == 0) { partial[3 * blockIdx.x] = s0; partial[3 * blockIdx.x + 1] = s1; partial[3 * blockIdx.x + 2] = s2; }
}

// particle initialisation (main.c:257-309): ZA/LPT mean removal (2LPT.c:1501-1508), ID, Vel, Pos
__global__ void __launch_bounds__(256)
k_ic_particles(size_t nloc, int ns, int p0, const float *__restrict__ za, const float *__restrict__ lpt, size_t stride,
               double m1x, double m1y, double m1z, double m2x, double m2y, double m2z, double box, int use_cola, double Di,
               double Di2, double dDdy, double dD2dy, float4 *__restrict__ pA, float4 *__restrict__ pB, float4 *__restrict__ pC,
               float2 *__restrict__ pE) {
  const float boxf = (float) box;
  const double dq = box / (double) ns;
  for (size_t c = blockIdx.x * (size_t) blockDim.x + threadIdx.x; c < nloc; c += (size_t) gridDim.x * blockDim.x) {
    const int k = (int) (c % ns);
    const size_t t = c / ns;
    const int j = (int) (t % ns), i = (int) (t / ns);
    const unsigned long long id = ((unsigned long long) ((long long) (i + p0) * ns + j)) * (unsigned long long) ns + (unsigned long long) k;
    float D[3], D2[3], V[3], X[3];
    const double m1[3] = {m1x, m1y, m1z}, m2[3] = {m2x, m2y, m2z};
    const double q[3] = {(i + p0) * dq, j * dq, k * dq};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      D[a] = (float) ((double) za[a * stride + c] - m1[a]);
      D2[a] = (float) ((double) lpt[a * stride + c] - m2[a]);
      V[a] = use_cola ? 0.0f : (float) ((double) D[a] * dDdy + (double) D2[a] * dD2dy);
      float x = (float) (q[a] + (double) D[a] * Di + (double) D2[a] * Di2);
      while (x >= boxf) x -= boxf;
      while (x < 0) x += boxf;
      if (x == boxf) x = 0.0f;
      X[a] = x;
    }
    pA[c] = make_float4(X[0], X[1], X[2], __uint_as_float((unsigned) (id & 0xffffffffull)));
    pB[c] = make_float4(V[0], V[1], V[2], __uint_as_float((unsigned) (id >> 32)));
    pC[c] = make_float4(D[0], D[1], D[2], D2[0]);
    pE[c] = make_float2(D2[1], D2[2]);
  }
}


// ------------------------------------------------------------------ READICFROMFILE (readICfromfile.c:133-215, 533-778)

// ProcessParticlesSingleFile (readICfromfile.c:133-215): external particles with coordinates in [0, 1), X = pos * Nmesh,
// only those of this rank's slab, CIC with W = (Nmesh/Nsample)^3 onto a grid that started at -1.
template <typename T>
__global__ void __launch_bounds__(256)
k_ic_assign_unit(size_t n, const float *__restrict__ pos, T *__restrict__ grid, int N, int NZ, int nx, int x0, int single_rank,
                 double W, unsigned long long *__restrict__ taken) {
  unsigned long long mine = 0;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    double X = (double) pos[3 * i];
    const int ixx = (int) (X * (double) N) - x0;
    if (ixx >= nx || ixx < 0) continue;
    mine++;
    X *= (double) N;
    const double Y = (double) pos[3 * i + 1] * (double) N, Z = (double) pos[3 * i + 2] * (double) N;
    unsigned ix = (unsigned) X, iy = (unsigned) Y, iz = (unsigned) Z;
    const double dx = X - (double) ix, dz = Z - (double) iz, tx = 1.0 - dx, tz = 1.0 - dz;
    double dy = Y - (double) iy, ty = 1.0 - dy;
    dy *= W; ty *= W;
    ix -= (unsigned) x0;
    if (iy >= (unsigned) N) iy = 0;
    if (iz >= (unsigned) N) iz = 0;
    unsigned ix1 = ix + 1;
    if (single_rank && ix1 == (unsigned) nx) ix1 = 0;          // one rank: the ghost plane is plane 0
    const unsigned iy1 = iy + 1 >= (unsigned) N ? 0 : iy + 1, iz1 = iz + 1 >= (unsigned) N ? 0 : iz + 1;
    const size_t rz = (size_t) 2 * NZ;
    T *r00 = grid + ((size_t) ix * N + iy) * rz, *r01 = grid + ((size_t) ix * N + iy1) * rz;
    T *r10 = grid + ((size_t) ix1 * N + iy) * rz, *r11 = grid + ((size_t) ix1 * N + iy1) * rz;
    atomicAdd(r00 + iz, (T) (tx * ty * tz)); atomicAdd(r00 + iz1, (T) (tx * ty * dz));
    atomicAdd(r01 + iz, (T) (tx * dy * tz)); atomicAdd(r01 + iz1, (T) (tx * dy * dz));
    atomicAdd(r10 + iz, (T) (dx * ty * tz)); atomicAdd(r10 + iz1, (T) (dx * ty * dz));
    atomicAdd(r11 + iz, (T) (dx * dy * tz)); atomicAdd(r11 + iz1, (T) (dx * dy * dz));
  }
  if (mine) atomicAdd(taken, mine);
}

template <typename T>
__global__ void k_ic_fill(T *g, size_t n, T v) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) g[i] = v;
}

// delta_k of the external particles -> the delta_k the 2LPT pipeline starts from: normalisation and growth to z = 0
// (readICfromfile.c:642-646), sharp-k filter at the Nyquist frequency of the particle grid when Nmesh > Nsample (648-680),
// CIC window deconvolution and the LCDM -> MG rescaling of AssignDisplacementField (701-778; rescale[m] at m = |d|^2)
template <typename T>
__global__ void __launch_bounds__(256)
k_ic_delta_ext(KL L, const typename Cpx<T>::type *__restrict__ p3d, const double *__restrict__ rescale, int nsample,
               double normfac, typename Cpx<T>::type *__restrict__ dk) {
  typedef typename Cpx<T>::type C;
  const int N = L.N;
  const double PI = 3.14159265358979323846;
  KLOOP(e, L) {
    int i, j, k;
    kl_decode(L, e, i, j, k);
    const int d0 = i > N / 2 ? i - N : i, d1 = j > N / 2 ? j - N : j, d2 = k;
    const long long m = (long long) d0 * d0 + (long long) d1 * d1 + (long long) d2 * d2;
    C out; out.x = (T) 0; out.y = (T) 0;
    const bool cut = N > nsample && sqrt((double) m) > (double) (nsample / 2);
    if (m != 0 && !cut) {
      double gc = 1.0;
      const int d[3] = {d0, d1, d2};
#pragma unroll
      for (int a = 0; a < 3; a++)
        if (d[a] != 0) gc *= sin((PI * d[a]) / (double) N) / ((PI * d[a]) / (double) N);
      gc = (1.0 / gc) * (1.0 / gc);                        // pow(1.0 / grid_corr, 2.0)
      const C v = p3d[e];
      // density *= normfac in the grid's precision (644-646), then * grid_corr * rescale_fac in double (748-749)
      const T re = (T) ((double) v.x * normfac), im = (T) ((double) v.y * normfac);
      const double rf = rescale[m];
      out.x = (T) (((double) re * gc) * rf); out.y = (T) (((double) im * gc) * rf);     // P3D * grid_corr * rescale_fac, left to right
    }
    dk[e] = out;
  }
}

// ------------------------------------------------------------------ host orchestration

// ext_rescale != nullptr: delta_k comes from the density of external particles (grid 0 holds their transformed density)
template <typename T>
static void ic_generate_t(Ctx &c, const mgp_ic_config *ic, const double *ext_rescale = nullptr, double ext_normfac = 0.0) {
  typedef typename Cpx<T>::type C;
  const int N = c.N, ns = c.cfg.nsample;
  const KL L = layout_of(c);
  const size_t nloc = (size_t) c.npl * ns * ns;
  REQUIRE(nloc <= c.cap, MGP_ERR_BUFFER, "mgp_ic_generate: particle capacity too small");
  const size_t mmax = (size_t) 3 * (N / 2) * (N / 2) + 1;
  const double *table = ext_rescale ? ext_rescale : ic->power_by_k2;
  const int ext = ext_rescale ? 1 : 0;     // AssignDisplacementField's wave-vector convention on the Nyquist planes (ic_modes.cuh)
  REQUIRE(table != nullptr && (ext_rescale || ic->n_power >= mmax), MGP_ERR_INVALID,
          "mgp_ic_generate: power_by_k2 must hold 3 (Nmesh/2)^2 + 1 entries");

  // scratch: delta_k lives in mgarray_one when the model has one, else in a temporary grid
  void *tmp = nullptr;
  C *dk = (C *) c.grid[MGP_GRID_MG_ONE];
  if (!dk) { CK(cudaMalloc(&tmp, c.grid_bytes())); dk = (C *) tmp; }
  void *save4 = c.grid[MGP_GRID_MG_ONE];
  c.grid[MGP_GRID_MG_ONE] = dk;           // so that the FFT helpers can address it by id

  std::vector<unsigned> seeds;
  if (ext_rescale) seeds.assign(1, 0u);
  else if (ic->seedtable) seeds.assign(ic->seedtable, ic->seedtable + (size_t) N * N);
  else make_seedtable(ic->seed, N, seeds);
  unsigned *d_seed = nullptr; double *d_pow = nullptr;
  CK(cudaMalloc(&d_seed, seeds.size() * sizeof(unsigned)));
  CK(cudaMalloc(&d_pow, mmax * sizeof(double)));
  CK(cudaMemcpyAsync(d_seed, seeds.data(), seeds.size() * sizeof(unsigned), cudaMemcpyHostToDevice, c.stream));
  CK(cudaMemcpyAsync(d_pow, table, mmax * sizeof(double), cudaMemcpyHostToDevice, c.stream));

  C *f[3] = {(C *) c.grid[1], (C *) c.grid[2], (C *) c.grid[3]};
  T *fr[3] = {(T *) c.grid[1], (T *) c.grid[2], (T *) c.grid[3]};
  T *S = (T *) c.grid[0];
  const unsigned gk = grid_for(L.total, 256), gr = grid_for(c.grid_vals, 256);

  CK(cudaMemsetAsync(dk, 0, c.grid_bytes(), c.stream));
  if (ext_rescale)
    k_ic_delta_ext<T><<<grid_for(L.total, 256), 256, 0, c.stream>>>(L, (const C *) c.grid[MGP_GRID_DENSITY], d_pow, ns, ext_normfac, dk);
  else
    k_ic_delta<T><<<(unsigned) (((size_t) N * N + 127) / 128), 128, 0, c.stream>>>(L, d_seed, d_pow, ns, c.cfg.box, ic->sphere_mode,
                                                                                 ic->amplitude_fixed, ic->inverted, dk);
  c.launches++;
  // second-order source from the six gradients
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 0) k_ic_kernel<T, 1><<<gk, 256, 0, c.stream>>>(L, dk, f[0], f[1], f[2], c.cfg.box, nullptr, 1.0, ext);
    else k_ic_kernel<T, 2><<<gk, 256, 0, c.stream>>>(L, dk, f[0], f[1], f[2], c.cfg.box, nullptr, 1.0, ext);
    fft_c2r_forces(c);
    k_ic_source<T><<<gr, 256, 0, c.stream>>>(S, fr[0], fr[1], fr[2], c.grid_vals, pass);
    c.launches += 2;
  }
  fft_r2c(c, MGP_GRID_DENSITY);           // S_k, unnormalised like the reference's

  if (c.cfg.scale_dependent) {
    // -DSCALEDEPENDENT: displacement_fields() returns here (2LPT.c:1361-1374) keeping delta1_k and
    // delta2_k = -S_k; the displacements are rebuilt from them with k-dependent growth factors (sd.cu)
    REQUIRE(c.sd_delta[0] && c.sd_delta[1], MGP_ERR_STATE, "scale-dependent storage missing");
    CK(cudaMemcpyAsync(c.sd_delta[0], dk, L.total * sizeof(C), cudaMemcpyDeviceToDevice, c.stream));
    size_t zero_index = (size_t) -1;
    if (!L.transposed || L.j0 == 0) zero_index = 0;      // mode (0,0,0) is element 0 of the rank that owns ky = 0
    k_ic_neg<C><<<gk, 256, 0, c.stream>>>((const C *) c.grid[0], (C *) c.sd_delta[1], L.total, zero_index);
    c.launches++;
    c.sd_have_delta = true;
    c.np = 0;
    c.sd_lagrangian_only = false;
  } else {
    // displacement fields at the Lagrangian points: ZA into disp[], 2LPT into the key/perm scratch
    float *za = c.disp;                      // [3][cap] floats
    float *lpt = (float *) c.pA2;            // the sort's second buffer set is idle here: 4 cap floats
    const unsigned gp = grid_for(nloc, 256, 8);
    reduce_alloc(c, (size_t) gp * 3 + 32);
    double means[6];
    const double n3 = (double) N * (double) N * (double) N;
    for (int order = 1; order <= 2; order++) {
      if (order == 1) k_ic_kernel<T, 0><<<gk, 256, 0, c.stream>>>(L, dk, f[0], f[1], f[2], c.cfg.box, nullptr, 1.0, ext);
      else k_ic_kernel<T, 3><<<gk, 256, 0, c.stream>>>(L, (const C *) c.grid[0], f[0], f[1], f[2], c.cfg.box, nullptr, 1.0, 0);
      fft_c2r_forces(c);
      halo_fill_forces(c);
      double *res = c.d_red + (size_t) gp * 3;
      k_ic_readout<T><<<gp, 256, 0, c.stream>>>(nloc, ns, c.p0, N, c.NZ, c.x0, c.nx, fr[0], fr[1], fr[2],
                                               order == 1 ? 1.0 : (-3.0 / 7.0) / n3, order == 1 ? za : lpt, c.cap, c.d_red);
      k_final_reduce<<<1, 256, 0, c.stream>>>(c.d_red, (int) gp, 3, 3, 1.0, res);
      c.launches += 3;
      allreduce_sum(c, res, 3);
      CK(cudaMemcpyAsync(c.h_red, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      CK(cudaStreamSynchronize(c.stream));
      const double tot = (double) ns * (double) ns * (double) ns;
      for (int a = 0; a < 3; a++) means[3 * (order - 1) + a] = c.h_red[a] / tot;
    }
    for (int a = 0; a < 6; a++) c.ic_means[a] = means[a];
    c.ic_ready = true;
    c.np = 0;
  }
  CK(cudaStreamSynchronize(c.stream));
  c.grid[MGP_GRID_MG_ONE] = save4;
  if (tmp) CK(cudaFree(tmp));
  CK(cudaFree(d_seed)); CK(cudaFree(d_pow));
}

void ic_generate(Ctx &c, const mgp_ic_config *ic) {
  REQUIRE(ic != nullptr, MGP_ERR_INVALID, "mgp_ic_generate: config is NULL");
  if (c.gbytes == 4) ic_generate_t<float>(c, ic); else ic_generate_t<double>(c, ic);
}

// ReadFilesMakeDisplacementField (readICfromfile.c:533-700) in three calls: begin (grid = -1), add (one per particle file),
// finish (halo, r2c, normalisation, filter, deconvolution, then the same 2LPT pipeline as the Gaussian branch)
void ic_particles_begin(Ctx &c) {
  if (c.gbytes == 4) k_ic_fill<float><<<grid_for(c.grid_vals, 256), 256, 0, c.stream>>>((float *) c.grid[0], c.grid_vals, -1.0f);
  else k_ic_fill<double><<<grid_for(c.grid_vals, 256), 256, 0, c.stream>>>((double *) c.grid[0], c.grid_vals, -1.0);
  c.launches++;
  c.ic_ext_taken = 0;
  c.ic_ext_open = true;
  c.density_live = false;
}

void ic_particles_add(Ctx &c, const float *pos01, uint64_t n) {
  REQUIRE(c.ic_ext_open, MGP_ERR_STATE, "mgp_ic_particles_add: call mgp_ic_particles_begin first");
  REQUIRE(pos01 != nullptr || n == 0, MGP_ERR_INVALID, "mgp_ic_particles_add: NULL");
  if (!n) return;
  const double r = (double) c.N / (double) c.cfg.nsample;
  const double W = r * r * r;
  const size_t chunk = (size_t) 1 << 24;
  float *d_pos = nullptr;
  unsigned long long *d_cnt = nullptr, h_cnt = 0;
  CK(cudaMalloc(&d_pos, (n < chunk ? n : chunk) * 3 * sizeof(float)));
  CK(cudaMalloc(&d_cnt, sizeof(unsigned long long)));
  CK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), c.stream));
  for (uint64_t off = 0; off < n; off += chunk) {
    const size_t m = (size_t) ((n - off) < chunk ? (n - off) : chunk);
    CK(cudaMemcpyAsync(d_pos, pos01 + 3 * off, m * 3 * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    if (c.gbytes == 4)
      k_ic_assign_unit<float><<<grid_for(m, 256), 256, 0, c.stream>>>(m, d_pos, (float *) c.grid[0], c.N, c.NZ, c.nx, c.x0, c.P == 1, W, d_cnt);
    else
      k_ic_assign_unit<double><<<grid_for(m, 256), 256, 0, c.stream>>>(m, d_pos, (double *) c.grid[0], c.N, c.NZ, c.nx, c.x0, c.P == 1, W, d_cnt);
    c.launches++;
    CK(cudaStreamSynchronize(c.stream));            // the staging buffer is reused (and pos01 may be pageable)
  }
  CK(cudaMemcpy(&h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost));
  CK(cudaFree(d_pos)); CK(cudaFree(d_cnt));
  c.ic_ext_taken += h_cnt;
}

void ic_particles_finish(Ctx &c, double normfac, const double *rescale_by_k2, size_t n) {
  REQUIRE(c.ic_ext_open, MGP_ERR_STATE, "mgp_ic_particles_finish: call mgp_ic_particles_begin first");
  const size_t mmax = (size_t) 3 * (c.N / 2) * (c.N / 2) + 1;
  REQUIRE(rescale_by_k2 != nullptr && n >= mmax, MGP_ERR_INVALID, "mgp_ic_particles_finish: rescale_by_k2 must hold 3 (Nmesh/2)^2 + 1 entries");
  c.ic_ext_open = false;
  halo_add_density(c, MGP_GRID_DENSITY);            // ghost plane -> right neighbour's plane 0, "+ 1" (readICfromfile.c:630-637)
  fft_r2c(c, MGP_GRID_DENSITY);
  mgp_ic_config ic;
  memset(&ic, 0, sizeof(ic));
  if (c.gbytes == 4) ic_generate_t<float>(c, &ic, rescale_by_k2, normfac); else ic_generate_t<double>(c, &ic, rescale_by_k2, normfac);
}

void ic_init_particles(Ctx &c, double Di, double Di2, double dDdy, double dD2dy) {
  REQUIRE(c.ic_ready, MGP_ERR_STATE, "mgp_init_particles: call mgp_ic_generate first");
  const int ns = c.cfg.nsample;
  const size_t nloc = (size_t) c.npl * ns * ns;
  const double *m = c.ic_means;
  if (nloc)
    k_ic_particles<<<grid_for(nloc, 256), 256, 0, c.stream>>>(nloc, ns, c.p0, c.disp, (const float *) c.pA2, c.cap, m[0], m[1], m[2],
                                                            m[3], m[4], m[5], c.cfg.box, c.cfg.use_cola, Di, Di2, dDdy, dD2dy, c.pA,
                                                            c.pB, c.pC, (float2 *) c.pE);
  c.launches++;
  CK(cudaStreamSynchronize(c.stream));
  c.np = nloc;
  c.sorted = false;
  c.bins_valid = false;
  c.drifts_since_sort = 1 << 30;
  c.have_disp = false;
  c.ic_ready = false;      // disp[] / pA2 scratch is reused from here on
}

void ic_seedtable(unsigned seed, int N, unsigned *out) {
  std::vector<unsigned> t;
  make_seedtable(seed, N, t);
  memcpy(out, t.data(), t.size() * sizeof(unsigned));
}

double ic_ranlxd1_draw(unsigned long seed, long n) {
  Ranlxd r;
  ranlxd_set(r, seed);
  double v = 0;
  for (long i = 0; i < n; i++) v = ranlxd_uniform(r);
  return v;
}

}  // namespace mgp
