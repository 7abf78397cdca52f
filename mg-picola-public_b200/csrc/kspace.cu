// Pointwise k-space and real-space grid kernels: Green's function + gradient (Forces,
// auxPM.c:450-535), the modified-gravity solvers of mg.h (DivideByLaplacian 21-64,
// EffDensitykToPhiofk 74-116, SmoothDensityField 268-309, the screening loops 174-176 and
// 241-250) and the P(k) binning of compute_pofk.c:71-236.
//
// Wave-number conventions are the reference's, not textbook ones: d_x = i > N/2 ? i-N : i (so
// +N/2 at Nyquist), same for y (the reference's explicit "mirror" rows N-j carry d_y = -j), kz in
// [0, N/2]; RK = |d|^2 in integer units; only the (0,0,0) mode is zeroed; no CIC deconvolution in
// the force (grid_corr = 1, auxPM.c:492).
#include "common.cuh"
#include "reduce.cuh"
#include "klayout.cuh"

#include <cmath>

namespace mgp {

// ------------------------------------------------------------------ Forces

// F_a,k = i d_a (-1/RK) delta_k / (Scale N^3)   written as the reference does:
//   dens = (Re*KK, -Im*KK) / N^3 ;  FN_a = (dens[1]*d_a/Scale, dens[0]*d_a/Scale)
template <typename T>
__global__ void __launch_bounds__(256)
k_forces(KL L, const typename Cpx<T>::type *__restrict__ dk, const typename Cpx<T>::type *__restrict__ mgk,
         typename Cpx<T>::type *__restrict__ f1, typename Cpx<T>::type *__restrict__ f2,
         typename Cpx<T>::type *__restrict__ f3, double n3, double scale) {
  typedef typename Cpx<T>::type C;
  KLOOP(e, L) {
    int i, j, k;
    kl_decode(L, e, i, j, k);
    const int d0 = i > L.N / 2 ? i - L.N : i, d1 = j > L.N / 2 ? j - L.N : j, d2 = k;
    C o1, o2, o3;
    if (i == 0 && j == 0 && k == 0) {
      o1.x = o1.y = o2.x = o2.y = o3.x = o3.y = (T) 0;
    } else {
      C v = dk[e];
      double re = (double) v.x, im = (double) v.y;
      if (mgk) { const C m = mgk[e]; re = (double) (T) (re + (double) m.x); im = (double) (T) (im + (double) m.y); }
      const double RK = (double) ((long long) d0 * d0 + (long long) d1 * d1 + (long long) d2 * d2);
      const double KK = -1.0 / RK;
      const double dens0 = (re * KK) / n3, dens1 = (-1.0 * im * KK) / n3;
      o1.x = (T) (dens1 * d0 / scale); o1.y = (T) (dens0 * d0 / scale);
      o2.x = (T) (dens1 * d1 / scale); o2.y = (T) (dens0 * d1 / scale);
      o3.x = (T) (dens1 * d2 / scale); o3.y = (T) (dens0 * d2 / scale);
    }
    f1[e] = o1; f2[e] = o2; f3[e] = o3;
  }
}

void kspace_forces(Ctx &c, bool add_mg) {
  const KL L = layout_of(c);
  const double n3 = (double) c.N * (double) c.N * (double) c.N;   // pow((double)Nmesh,3)
  const double scale = 2. * M_PI / c.cfg.box;
  const unsigned g = grid_for(L.total, 256);
  if (c.gbytes == 4)
    k_forces<float><<<g, 256, 0, c.stream>>>(L, (const float2 *) c.grid[0], add_mg ? (const float2 *) c.grid[5] : nullptr,
                                             (float2 *) c.grid[1], (float2 *) c.grid[2], (float2 *) c.grid[3], n3, scale);
  else
    k_forces<double><<<g, 256, 0, c.stream>>>(L, (const double2 *) c.grid[0], add_mg ? (const double2 *) c.grid[5] : nullptr,
                                              (double2 *) c.grid[1], (double2 *) c.grid[2], (double2 *) c.grid[3], n3, scale);
  c.launches++;
}

// ------------------------------------------------------------------ mg.h k-space kernels

// mode: 0 DivideByLaplacian      out = norm * in * (-1/RK), zero mode 0            (mg.h:21-64)
//       1 EffDensitykToPhiofk    out = norm * in * RK/(RK + massterm2), zero mode 0 (mg.h:74-116)
//       2 SmoothDensityField     out = in * exp(-0.5 (sqrt(RK) kfR)^2) * norm       (mg.h:268-309)
//       3 scale                  out = in * norm
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
k_kmul(KL L, const typename Cpx<T>::type *__restrict__ in, typename Cpx<T>::type *__restrict__ out, double norm,
       double par) {
  typedef typename Cpx<T>::type C;
  KLOOP(e, L) {
    int i, j, k;
    kl_decode(L, e, i, j, k);
    const int d0 = i > L.N / 2 ? L.N - i : i, d1 = j > L.N / 2 ? L.N - j : j;
    const double RK = (double) ((long long) k * k + (long long) d0 * d0 + (long long) d1 * d1);
    const C v = in[e];
    C o;
    if (MODE == 0) {
      if (i == 0 && j == 0 && k == 0) { o.x = o.y = (T) 0; }
      else { const double KK = -1.0 / RK; o.x = (T) (norm * (double) v.x * KK); o.y = (T) (norm * (double) v.y * KK); }
    } else if (MODE == 1) {
      if (i == 0 && j == 0 && k == 0) { o.x = o.y = (T) 0; }
      else { const double KK = RK / (RK + par); o.x = (T) (norm * (double) v.x * KK); o.y = (T) (norm * (double) v.y * KK); }
    } else if (MODE == 2) {
      const double kR = sqrt(RK) * par;
      const double sm = exp(-0.5 * kR * kR) * norm;
      o.x = (T) ((double) v.x * sm); o.y = (T) ((double) v.y * sm);
    } else {
      o.x = (T) ((double) v.x * norm); o.y = (T) ((double) v.y * norm);
    }
    out[e] = o;
  }
}

template <int MODE>
static void kmul(Ctx &c, int src, int dst, double norm, double par) {
  const KL L = layout_of(c);
  const unsigned g = grid_for(L.total, 256);
  if (c.gbytes == 4) k_kmul<float, MODE><<<g, 256, 0, c.stream>>>(L, (const float2 *) c.grid[src], (float2 *) c.grid[dst], norm, par);
  else k_kmul<double, MODE><<<g, 256, 0, c.stream>>>(L, (const double2 *) c.grid[src], (double2 *) c.grid[dst], norm, par);
  c.launches++;
}

void kspace_divide_laplacian(Ctx &c, double normfactor) { kmul<0>(c, MGP_GRID_DENSITY, MGP_GRID_MG_ONE, normfactor, 0.0); }
void kspace_phi_of_k(Ctx &c, int src_grid, double coupling, double massterm2) { kmul<1>(c, src_grid, MGP_GRID_MG_TWO, coupling, massterm2); }
void kspace_smooth(Ctx &c, double rsmooth) {
  const double n3 = (double) c.N * (double) c.N * (double) c.N;
  kmul<2>(c, MGP_GRID_DENSITY, MGP_GRID_MG_ONE, 1.0 / n3, 2.0 * M_PI / c.cfg.box * rsmooth);
}
void kspace_scale(Ctx &c, int grid_id, double f) { kmul<3>(c, grid_id, grid_id, f, 0.0); }
void kspace_scale_to(Ctx &c, int src, int dst, double f) { kmul<3>(c, src, dst, f, 0.0); }

// ------------------------------------------------------------------ real-space kernels (loop over all 2*Total_size values)

template <typename T>
__global__ void k_copy_scale(const T *__restrict__ src, T *__restrict__ dst, size_t n, double s) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    dst[i] = s == 1.0 ? src[i] : (T) ((double) src[i] * s);
}

void real_copy(Ctx &c, int dst, int src, double scale) {
  const size_t n = c.grid_vals;
  if (c.gbytes == 4) k_copy_scale<float><<<grid_for(n, 256), 256, 0, c.stream>>>((const float *) c.grid[src], (float *) c.grid[dst], n, scale);
  else k_copy_scale<double><<<grid_for(n, 256), 256, 0, c.stream>>>((const double *) c.grid[src], (double *) c.grid[dst], n, scale);
  c.launches++;
}

// mgarray_two[j] *= screening_factor_potential(a, mgarray_one[j])   (mg.h:174-176; udf:725-737)
template <typename T>
__global__ void k_screen_potential(const T *__restrict__ phi, T *__restrict__ dens, size_t n, double phicrit) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    const double ph = (double) phi[i];
    double s = 1.0;
    if (!(ph >= 0.0)) { s = fabs(phicrit / ph); if (s > 1.0) s = 1.0; }
    dens[i] = (T) ((double) dens[i] * s);
  }
}

void real_screen_potential(Ctx &c, double phi_crit, bool screening) {
  if (!screening) return;
  const size_t n = c.grid_vals;
  if (c.gbytes == 4) k_screen_potential<float><<<grid_for(n, 256), 256, 0, c.stream>>>((const float *) c.grid[4], (float *) c.grid[5], n, phi_crit);
  else k_screen_potential<double><<<grid_for(n, 256), 256, 0, c.stream>>>((const double *) c.grid[4], (double *) c.grid[5], n, phi_crit);
  c.launches++;
}

// mgarray_two[j] *= coupling * screening_factor_density(a, mgarray_one[j])  (mg.h:241-250; udf:757-766)
// partial[b] = (sum, max, min) of screenfac
template <typename T>
__global__ void __launch_bounds__(256)
k_screen_density(const T *__restrict__ ds, T *__restrict__ dens, size_t n, double coupling, double fac0,
                 double *__restrict__ partial) {
  double sum = 0, mx = 0.0, mn = 1e100;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    const double fac = fac0 * (1.0 + (double) ds[i]);
    double s = 1.0;
    if (!(fac < 1e-5)) s = 2.0 * (sqrt(1.0 + fac) - 1.0) / fac;
    dens[i] = (T) ((double) dens[i] * (coupling * s));
    sum += s; if (s > mx) mx = s; if (s < mn) mn = s;
  }
  __shared__ double sh[3][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sh[0][w] = sum; sh[1][w] = mx; sh[2][w] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < (int) (blockDim.x >> 5); q++) { sum += sh[0][q]; mx = fmax(mx, sh[1][q]); mn = fmin(mn, sh[2][q]); }
    partial[3 * blockIdx.x] = sum; partial[3 * blockIdx.x + 1] = mx; partial[3 * blockIdx.x + 2] = mn;
  }
}

void real_screen_density(Ctx &c, double coupling, double fac0, double stats[3]) {
  const size_t n = c.grid_vals;
  const unsigned g = grid_for(n, 256, 8);
  reduce_alloc(c, (size_t) g * 3 + 16);
  if (c.gbytes == 4) k_screen_density<float><<<g, 256, 0, c.stream>>>((const float *) c.grid[4], (float *) c.grid[5], n, coupling, fac0, c.d_red);
  else k_screen_density<double><<<g, 256, 0, c.stream>>>((const double *) c.grid[4], (double *) c.grid[5], n, coupling, fac0, c.d_red);
  c.launches++;
  if (stats) {
    CK(cudaMemcpyAsync(c.h_red, c.d_red, (size_t) g * 3 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    double sum = 0, mx = 0, mn = 1e100;
    for (unsigned b = 0; b < g; b++) { sum += c.h_red[3 * b]; mx = fmax(mx, c.h_red[3 * b + 1]); mn = fmin(mn, c.h_red[3 * b + 2]); }
    stats[0] = sum / (double) n; stats[1] = mx; stats[2] = mn;     // avg, max, min (mg.h:251-254)
  }
}

// ------------------------------------------------------------------ massive neutrinos (auxPM.c:383-420)

// P3D = cdmfac * P3D + nufac(|k|) * cdelta_cdm for every mode but (0,0,0); nufac[m], m = |d|^2, is
// OmegaNu/Omega * Nmesh^3 * T_nu(k, a) / T_cb(k, 1) filled by the driver from its CAMB splines
template <typename T>
__global__ void __launch_bounds__(256)
k_nu_add(KL L, typename Cpx<T>::type *__restrict__ p3d, const typename Cpx<T>::type *__restrict__ d1,
         const double *__restrict__ nufac, double cdmfac) {
  typedef typename Cpx<T>::type C;
  const int N = L.N;
  KLOOP(e, L) {
    int i, j, k;
    kl_decode(L, e, i, j, k);
    if (i == 0 && j == 0 && k == 0) continue;
    const int d0 = i > N / 2 ? N - i : i, d1i = j > N / 2 ? N - j : j;
    const double nf = nufac[(long long) d0 * d0 + (long long) d1i * d1i + (long long) k * k];
    C v = p3d[e];
    const C w = d1[e];
    v.x = (T) (cdmfac * (double) v.x + nf * (double) w.x);
    v.y = (T) (cdmfac * (double) v.y + nf * (double) w.y);
    p3d[e] = v;
  }
}

void kspace_nu_add(Ctx &c, const double *nufac_host, size_t n, double cdmfac) {
  const size_t mmax = (size_t) 3 * (c.N / 2) * (c.N / 2) + 1;
  REQUIRE(nufac_host != nullptr && n >= mmax, MGP_ERR_INVALID, "neutrino table must hold 3 (Nmesh/2)^2 + 1 entries");
  REQUIRE(c.sd_delta[0] && c.sd_have_delta, MGP_ERR_STATE, "massive neutrinos need the stored delta_cdm(k) (scale_dependent = 1 and mgp_ic_generate)");
  if (!c.nu_tab_d) CK(cudaMalloc(&c.nu_tab_d, mmax * sizeof(double)));
  CK(cudaMemcpyAsync(c.nu_tab_d, nufac_host, mmax * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  const KL L = layout_of(c);
  const unsigned g = grid_for(L.total, 256);
  if (c.gbytes == 4) k_nu_add<float><<<g, 256, 0, c.stream>>>(L, (float2 *) c.grid[0], (const float2 *) c.sd_delta[0], c.nu_tab_d, cdmfac);
  else k_nu_add<double><<<g, 256, 0, c.stream>>>(L, (double2 *) c.grid[0], (const double2 *) c.sd_delta[0], c.nu_tab_d, cdmfac);
  c.launches++;
}

// ------------------------------------------------------------------ P(k)

// The reference's parameter sanitiser (compute_pofk.c:758-805), in integer-k units.
static void adjust_pofk(const Ctx &c, int &nbins, int &bintype, int &shot, double &kmin, double &kmax) {
  nbins = c.pofk.nbins; bintype = c.pofk.bintype; shot = c.pofk.subtract_shotnoise;
  kmin = c.pofk.kmin * c.cfg.box / (2.0 * M_PI);
  kmax = c.pofk.kmax * c.cfg.box / (2.0 * M_PI);
  if (!(bintype == 0 || bintype == 1)) bintype = 0;
  if (!(shot == 0 || shot == 1)) shot = 1;
  if (nbins <= 0) nbins = c.N;
  if (kmax <= kmin) { if (bintype == 0) kmin = 0.0; if (bintype == 1) kmin = 1.0; kmax = (double) c.N; }
  if (kmin < 0.0) { if (bintype == 0) kmin = 0.0; if (bintype == 1) kmin = 1.0; }
  if (bintype == 1 && kmin == 0.0) kmin = 1.0;
  if (kmax > sqrt(3.0) * (double) c.N) kmax = (double) c.N;
}

int pofk_effective_nbins(const Ctx &c) {
  int nbins, bintype, shot; double kmin, kmax;
  adjust_pofk(c, nbins, bintype, shot, kmin, kmax);
  return nbins;
}

// pofk_bin_index (compute_pofk.c:30-46), evaluated on the host for every integer |d|^2
static int bin_index(double kmag, double kmin, double kmax, int nbins, int bintype) {
  if (bintype == 0) return (int) ((kmag - kmin) / (kmax - kmin) * nbins + 0.5);
  if (kmag <= 0.0) return -1;
  return (int) (log(kmag / kmin) / log(kmax / kmin) * nbins + 0.5);
}

// RSD = 1 adds the mu^2 and mu^4 moments of bin_up_RSD_power_spectrum (compute_pofk.c:518-753), mu^2 = kz^2 / k^2
template <typename T, int RSD>
__global__ void __launch_bounds__(256)
k_pofk(KL L, const typename Cpx<T>::type *__restrict__ dk, const int *__restrict__ bin_of_m,
       const double *__restrict__ sinc, int nbins, double norm, double *__restrict__ out) {
  extern __shared__ double sb[];     // [3 (+2)][nbins]: sum P, sum k, sum n (, sum P mu^2, sum P mu^4)
  constexpr int NS = RSD ? 5 : 3;
  for (int b = threadIdx.x; b < NS * nbins; b += blockDim.x) sb[b] = 0.0;
  __syncthreads();
  typedef typename Cpx<T>::type C;
  const int N = L.N;
  // Neighbouring modes fall into the same few bins: the lanes of a warp are summed per distinct bin with shuffles and
  // ONE lane adds the total, instead of 32 lanes fighting over one shared-memory address (atomicAdd on a shared double is
  // a compare-and-swap loop on this part).
  const size_t nround = (L.total + 31) / 32 * 32;
  const unsigned lane = threadIdx.x & 31;
  for (size_t e = blockIdx.x * (size_t) blockDim.x + threadIdx.x; e < nround; e += (size_t) gridDim.x * blockDim.x) {
    int nk = -1;
    double vp = 0.0, vk = 0.0, vn = 0.0, v2 = 0.0, v4 = 0.0;
    if (e < L.total) {
      int i, j, k;
      kl_decode(L, e, i, j, k);
      const int d0 = i > N / 2 ? N - i : i, d1 = j > N / 2 ? N - j : j, d2 = k;
      const long long m = (long long) d0 * d0 + (long long) d1 * d1 + (long long) d2 * d2;
      nk = bin_of_m[m];
      if (nk >= nbins) nk = -1;
      if (nk >= 0) {
        const double gz = (d2 == 0) ? 1.0 : ((2 * d2 == N) ? 2.0 / 3.14159265358979323846 : sinc[d2]);
        const double gc = sinc[d0] * sinc[d1] * gz;
        const double g2 = gc * gc;
        const double corr = 1.0 / (g2 * g2) * norm;          // 1/pow(gx*gy*gz, 4) * fftw_norm_fac
        const C v = dk[e];
        const double p = ((double) v.x * (double) v.x + (double) v.y * (double) v.y) * corr;
        const double w = (d2 == 0 || 2 * d2 == N) ? 1.0 : 2.0;
        const double kmag = sqrt((double) m);
        vp = w * p; vk = w * kmag; vn = w;
        if (RSD) {
          // mu2 = d[2]*d[2]/kmag/kmag; the (0,0,0) mode has mu2 = 0 (compute_pofk.c:601-602)
          const double mu2 = m == 0 ? 0.0 : (double) ((long long) d2 * d2) / kmag / kmag;
          v2 = vp * mu2; v4 = vp * mu2 * mu2;
        }
      }
    }
    unsigned todo = __ballot_sync(0xffffffffu, nk >= 0);
    while (todo) {
      const int b = __shfl_sync(0xffffffffu, nk, __ffs(todo) - 1);
      const bool mine = nk == b;
      double a = mine ? vp : 0.0, bk = mine ? vk : 0.0, cn = mine ? vn : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); bk += __shfl_xor_sync(0xffffffffu, bk, o); cn += __shfl_xor_sync(0xffffffffu, cn, o);
      }
      double r2 = 0.0, r4 = 0.0;
      if (RSD) {
        r2 = mine ? v2 : 0.0; r4 = mine ? v4 : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { r2 += __shfl_xor_sync(0xffffffffu, r2, o); r4 += __shfl_xor_sync(0xffffffffu, r4, o); }
      }
      if (lane == 0) {
        atomicAdd(&sb[b], a); atomicAdd(&sb[nbins + b], bk); atomicAdd(&sb[2 * nbins + b], cn);
        if (RSD) { atomicAdd(&sb[3 * nbins + b], r2); atomicAdd(&sb[4 * nbins + b], r4); }
      }
      todo &= ~__ballot_sync(0xffffffffu, mine);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < NS * nbins; b += blockDim.x)
    if (sb[b] != 0.0) atomicAdd(&out[b], sb[b]);
}

// raw per-bin sums [ns][nbins] (ns = 3, or 5 with the RSD moments) in c.pofk_out_h, all-reduced over the ranks
static void pofk_sums(Ctx &c, int gid, bool rsd, int &nbins, int &bintype, int &shot, double &kmin, double &kmax) {
  REQUIRE(c.pofk_set, MGP_ERR_STATE, "P(k) requested but mgp_set_pofk_config was never called");
  adjust_pofk(c, nbins, bintype, shot, kmin, kmax);
  REQUIRE((size_t) 5 * nbins * sizeof(double) <= 200 * 1024, MGP_ERR_INVALID, "pofk_nbins too large");
  const int N = c.N, h = N / 2;
  const size_t mmax = (size_t) 3 * h * h + 1;
  // bin-of-|d|^2 and sinc tables live on the device until the binning parameters change
  if (!c.pofk_tables_valid || c.pofk_tab_nbins != nbins || c.pofk_tab_bintype != bintype || c.pofk_tab_kmin != kmin ||
      c.pofk_tab_kmax != kmax) {
    std::vector<int> bins(mmax);
    for (size_t m = 0; m < mmax; m++) bins[m] = bin_index(sqrt((double) m), kmin, kmax, nbins, bintype);
    std::vector<double> sinc(h + 1);
    sinc[0] = 1.0;
    for (int d = 1; d <= h; d++) sinc[d] = sin((M_PI * d) / (double) N) / ((M_PI * d) / (double) N);
    if (c.pofk_bins_d) { CK(cudaFree(c.pofk_bins_d)); CK(cudaFree(c.pofk_sinc_d)); CK(cudaFree(c.pofk_out_d)); CK(cudaFreeHost(c.pofk_out_h)); }
    CK(cudaMalloc(&c.pofk_bins_d, mmax * sizeof(int)));
    CK(cudaMalloc(&c.pofk_sinc_d, (h + 1) * sizeof(double)));
    CK(cudaMalloc(&c.pofk_out_d, (size_t) 5 * nbins * sizeof(double)));
    CK(cudaMallocHost(&c.pofk_out_h, (size_t) 5 * nbins * sizeof(double)));
    CK(cudaMemcpyAsync(c.pofk_bins_d, bins.data(), mmax * sizeof(int), cudaMemcpyHostToDevice, c.stream));
    CK(cudaMemcpyAsync(c.pofk_sinc_d, sinc.data(), (h + 1) * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    c.pofk_tables_valid = true;
    c.pofk_tab_nbins = nbins; c.pofk_tab_bintype = bintype; c.pofk_tab_kmin = kmin; c.pofk_tab_kmax = kmax;
  }
  int *d_bins = c.pofk_bins_d; double *d_sinc = c.pofk_sinc_d, *d_out = c.pofk_out_d;
  const int ns = rsd ? 5 : 3;
  CK(cudaMemsetAsync(d_out, 0, (size_t) ns * nbins * sizeof(double), c.stream));
  const KL L = layout_of(c);
  const double n3 = (double) N * (double) N * (double) N;
  const double norm = 1.0 / (n3 * n3);                      // 1/pow(Nmesh,6)
  const size_t sm = (size_t) ns * nbins * sizeof(double);
  const unsigned g = grid_for(L.total, 256, 4);
#define LAUNCH_POFK(T, C2, R)                                                                                         \
  do {                                                                                                                \
    CK(cudaFuncSetAttribute(k_pofk<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));                    \
    k_pofk<T, R><<<g, 256, sm, c.stream>>>(L, (const C2 *) c.grid[gid], d_bins, d_sinc, nbins, norm, d_out);          \
  } while (0)
  if (c.gbytes == 4) { if (rsd) LAUNCH_POFK(float, float2, 1); else LAUNCH_POFK(float, float2, 0); }
  else { if (rsd) LAUNCH_POFK(double, double2, 1); else LAUNCH_POFK(double, double2, 0); }
#undef LAUNCH_POFK
  c.launches++;
  allreduce_sum(c, d_out, ns * nbins);
  CK(cudaMemcpyAsync(c.pofk_out_h, d_out, (size_t) ns * nbins * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
}

void pofk_bin(Ctx &c, int gid, double *pofk, double *kmean, double *nmodes) {
  PhaseTimer t(c, PH_POFK);
  int nbins, bintype, shot; double kmin, kmax;
  pofk_sums(c, gid, false, nbins, bintype, shot, kmin, kmax);
  const double *hout = c.pofk_out_h;
  // normalise and subtract shot noise (compute_pofk.c:230-236)
  const double box3 = pow(c.cfg.box, 3);
  const double shotv = pow(c.cfg.box / (double) c.cfg.nsample, 3);
  for (int b = 0; b < nbins; b++) {
    const double nb = hout[2 * nbins + b];
    double p = 0.0, km = 0.0;
    if (nb > 0) {
      p = (hout[b] / nb) * box3;
      if (shot) p -= shotv;
      km = (hout[nbins + b] / nb) * 2.0 * M_PI / c.cfg.box;
    }
    if (pofk) pofk[b] = p;
    if (kmean) kmean[b] = km;
    if (nmodes) nmodes[b] = nb;
  }
}

// bin_up_RSD_power_spectrum (compute_pofk.c:518-753): out[0..4][nbins] = n, k (bin centre, k_from_index), P0, P2, P4
void pofk_bin_rsd(Ctx &c, int gid, double *out) {
  PhaseTimer t(c, PH_POFK);
  int nbins, bintype, shot; double kmin, kmax;
  pofk_sums(c, gid, true, nbins, bintype, shot, kmin, kmax);
  const double *h = c.pofk_out_h;
  const double box3 = pow(c.cfg.box, 3);
  const double shotv = pow(c.cfg.box / (double) c.cfg.nsample, 3);
  for (int b = 0; b < nbins; b++) {
    const double nb = h[2 * nbins + b];
    double p0 = 0, p2 = 0, p4 = 0, kb = 0;
    if (nb > 0) {
      p0 = (h[b] / nb) * box3; p2 = (h[3 * nbins + b] / nb) * box3; p4 = (h[4 * nbins + b] / nb) * box3;
      p4 = 9.0 * (35.0 * p4 - 30.0 * p2 + 3.0 * p0) / 8.0;     // powers of mu -> Legendre combinations, in this order (708-710)
      p2 = 5.0 * (3.0 * p2 - 1.0 * p0) / 2.0;
      if (shot) { p0 -= shotv; p2 -= shotv; p4 -= shotv; }      // all three, as the reference does (712-716)
      const double kk = bintype == 0 ? kmin + (kmax - kmin) / (double) nbins * b
                                     : exp(log(kmin) + log(kmax / kmin) / (double) nbins * b);   // k_from_index (50-64)
      kb = kk * 2.0 * M_PI / c.cfg.box;
    }
    out[b] = nb; out[nbins + b] = kb; out[2 * nbins + b] = p0; out[3 * nbins + b] = p2; out[4 * nbins + b] = p4;
  }
}

}  // namespace mgp
