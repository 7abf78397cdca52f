// The reference's stand-alone P(k) estimator (SimplePofk/main.cpp) on the particles a context holds: NGP / CIC / TSC
// assignment of counts (main.cpp:59-228), one r2c, per-mode window deconvolution and the tool's integer binning
// (324-424).  main() turns the counts into the density contrast before the transform (513-533: divide by the mean count
// Npart / N^3, subtract 1); the bins exclude k = 0, so here the counts are transformed and the factor (N^3 / Npart)^2
// goes into the normalisation of |d_k|^2.  Post-processing, not part of a COLA step: simple atomic deposits, one rank only (TSC reaches the plane on
// the left, for which a slab keeps no ghost).
#include "common.cuh"
#include "klayout.cuh"

namespace mgp {

// scheme 1 NGP, 2 CIC, 3 TSC.  x = double(pos_float / boxsize) * ngrid (main.cpp:259-261, 62-64): a float divided by a
// double, then the multiplication -- not the Pos * (Nmesh / Box) of PtoMesh.
// slip = 1: TSC exactly as published, i.e. the weight P_z of the three "this y" lines lands on the NEXT z plane
// (main.cpp:189, 201, 213 write izneighN where the comment says 0 = previous); slip = 0: the textbook stencil.
template <typename T, int SCHEME>
__global__ void __launch_bounds__(256)
k_assign_scheme(size_t n, const float4 *__restrict__ pA, T *__restrict__ grid, int N, int NZ, double box, int slip) {
  const size_t rz = (size_t) 2 * NZ;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    const float4 p = pA[i];
    const double X = ((double) p.x / box) * (double) N, Y = ((double) p.y / box) * (double) N, Z = ((double) p.z / box) * (double) N;
    long ix = (long) X, iy = (long) Y, iz = (long) Z;
    const double dx = X - (double) ix, dy = Y - (double) iy, dz = Z - (double) iz;
    if (ix >= N) ix -= N;
    if (iy >= N) iy -= N;
    if (iz >= N) iz -= N;
    if (SCHEME == 1) {
      atomicAdd(grid + ((size_t) ix * N + iy) * rz + iz, (T) 1.0);
    } else if (SCHEME == 2) {
      const long x1 = ix + 1 >= N ? ix + 1 - N : ix + 1, y1 = iy + 1 >= N ? iy + 1 - N : iy + 1, z1 = iz + 1 >= N ? iz + 1 - N : iz + 1;
      const double tx = 1.0 - dx, ty = 1.0 - dy, tz = 1.0 - dz;
      const long xs[2] = {ix, x1}, ys[2] = {iy, y1}, zs[2] = {iz, z1};
      const double wx[2] = {tx, dx}, wy[2] = {ty, dy}, wz[2] = {tz, dz};
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
          for (int c = 0; c < 2; c++) atomicAdd(grid + ((size_t) xs[a] * N + ys[b]) * rz + zs[c], (T) (wx[a] * wy[b] * wz[c]));
    } else {
      const long xn = ix + 1 >= N ? ix + 1 - N : ix + 1, yn = iy + 1 >= N ? iy + 1 - N : iy + 1, zn = iz + 1 >= N ? iz + 1 - N : iz + 1;
      const long xp = ix - 1 < 0 ? ix - 1 + N : ix - 1, yp = iy - 1 < 0 ? iy - 1 + N : iy - 1, zp = iz - 1 < 0 ? iz - 1 + N : iz - 1;
      const long xs[3] = {xp, ix, xn}, ys[3] = {yp, iy, yn};
      const double wx[3] = {0.5 * (0.5 - dx) * (0.5 - dx), 0.75 - dx * dx, 0.5 * (0.5 + dx) * (0.5 + dx)};
      const double wy[3] = {0.5 * (0.5 - dy) * (0.5 - dy), 0.75 - dy * dy, 0.5 * (0.5 + dy) * (0.5 + dy)};
      const double wz[3] = {0.5 * (0.5 - dz) * (0.5 - dz), 0.75 - dz * dz, 0.5 * (0.5 + dz) * (0.5 + dz)};
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
          T *row = grid + ((size_t) xs[a] * N + ys[b]) * rz;
          const double wxy = wx[a] * wy[b];
          atomicAdd(row + ((slip && b == 1) ? zn : zp), (T) (wxy * wz[0]));
          atomicAdd(row + iz, (T) (wxy * wz[1]));
          atomicAdd(row + zn, (T) (wxy * wz[2]));
        }
    }
  }
}

// per-bin sums of |d_k|^2 * norm / window^2 (norm = 1 / N^6 of main.cpp:380 times (N^3 / Npart)^2) and of the mode count over the FULL complex cube (the tool transforms complex
// data): a mode of the half spectrum with 0 < kz < N/2 stands for itself and its conjugate
template <typename T>
__global__ void __launch_bounds__(256)
k_simple_pofk(KL L, const typename Cpx<T>::type *__restrict__ dk, const double *__restrict__ sinc, int power, double norm,
              double *__restrict__ out) {
  extern __shared__ double sb[];     // [2][N]
  const int N = L.N;
  for (int b = threadIdx.x; b < 2 * N; b += blockDim.x) sb[b] = 0.0;
  __syncthreads();
  typedef typename Cpx<T>::type C;
  const size_t nround = (L.total + 31) / 32 * 32;
  const unsigned lane = threadIdx.x & 31;
  for (size_t e = blockIdx.x * (size_t) blockDim.x + threadIdx.x; e < nround; e += (size_t) gridDim.x * blockDim.x) {
    int nk = -1;
    double vp = 0.0, vn = 0.0;
    if (e < L.total) {
      int i, j, k;
      kl_decode(L, e, i, j, k);
      const int d0 = i >= N / 2 ? N - i : i, d1 = j >= N / 2 ? N - j : j, d2 = k;      // |ii|, |jj|, |kk| (main.cpp:382-390)
      const long long m = (long long) d0 * d0 + (long long) d1 * d1 + (long long) d2 * d2;
      const long long kind = (long long) (sqrt((double) m) + 0.5);
      if (kind < N && kind > 0) {
        nk = (int) kind;
        double w = sinc[d0] * sinc[d1] * sinc[d2];
        double wp = w;
        for (int q = 1; q < power; q++) wp *= w;
        const C v = dk[e];
        const double mult = (d2 == 0 || 2 * d2 == N) ? 1.0 : 2.0;
        vp = mult * ((double) v.x * (double) v.x + (double) v.y * (double) v.y) * norm / (wp * wp);
        vn = mult;
      }
    }
    unsigned todo = __ballot_sync(0xffffffffu, nk >= 0);
    while (todo) {
      const int b = __shfl_sync(0xffffffffu, nk, __ffs(todo) - 1);
      const bool mine = nk == b;
      double a = mine ? vp : 0.0, cn = mine ? vn : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); cn += __shfl_xor_sync(0xffffffffu, cn, o); }
      if (lane == 0) { atomicAdd(&sb[b], a); atomicAdd(&sb[N + b], cn); }
      todo &= ~__ballot_sync(0xffffffffu, mine);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < 2 * N; b += blockDim.x)
    if (sb[b] != 0.0) atomicAdd(&out[b], sb[b]);
}

template <typename T>
static void simple_pofk_t(Ctx &c, int scheme, int slip, double *pofk_sum, double *nmodes) {
  const int gid = MGP_GRID_FORCE_X, N = c.N;
  T *grid = (T *) c.grid[gid];
  CK(cudaMemsetAsync(grid, 0, c.grid_bytes(), c.stream));
  if (c.np) {
    const unsigned g = grid_for(c.np, 256);
    if (scheme == 1) k_assign_scheme<T, 1><<<g, 256, 0, c.stream>>>(c.np, c.pA, grid, N, c.NZ, c.cfg.box, slip);
    else if (scheme == 2) k_assign_scheme<T, 2><<<g, 256, 0, c.stream>>>(c.np, c.pA, grid, N, c.NZ, c.cfg.box, slip);
    else k_assign_scheme<T, 3><<<g, 256, 0, c.stream>>>(c.np, c.pA, grid, N, c.NZ, c.cfg.box, slip);
    c.launches++;
  }
  fft_r2c(c, gid);
  // sin(pi d / N) / (pi d / N), d = 0 .. N/2 (window(), main.cpp:324-340)
  std::vector<double> sinc(N / 2 + 1);
  sinc[0] = 1.0;
  for (int d = 1; d <= N / 2; d++) sinc[d] = sin(M_PI / (double) N * d) / (M_PI / (double) N * d);
  double *d_sinc = nullptr, *d_out = nullptr;
  CK(cudaMalloc(&d_sinc, sinc.size() * sizeof(double)));
  CK(cudaMalloc(&d_out, (size_t) 2 * N * sizeof(double)));
  CK(cudaMemcpyAsync(d_sinc, sinc.data(), sinc.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  CK(cudaMemsetAsync(d_out, 0, (size_t) 2 * N * sizeof(double), c.stream));
  const KL L = layout_of(c);
  const double n3 = (double) N * (double) N * (double) N;
  const size_t sm = (size_t) 2 * N * sizeof(double);
  CK(cudaFuncSetAttribute(k_simple_pofk<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
  k_simple_pofk<T><<<grid_for(L.total, 256, 4), 256, sm, c.stream>>>(L, (const typename Cpx<T>::type *) c.grid[gid], d_sinc, scheme,
                                                                     (1.0 / n3) * (1.0 / n3) * (n3 / (double) c.np) * (n3 / (double) c.np), d_out);
  c.launches++;
  std::vector<double> h((size_t) 2 * N);
  CK(cudaMemcpyAsync(h.data(), d_out, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  CK(cudaFree(d_sinc)); CK(cudaFree(d_out));
  for (int b = 0; b < N; b++) { pofk_sum[b] = h[b]; nmodes[b] = h[N + b]; }
}

void simple_pofk(Ctx &c, int scheme, int subtract_shotnoise, int slip, double *pofk, double *nmodes) {
  REQUIRE(scheme >= 1 && scheme <= 3, MGP_ERR_INVALID, "mgp_simple_pofk: scheme must be 1 (NGP), 2 (CIC) or 3 (TSC)");
  REQUIRE(c.np > 0, MGP_ERR_STATE, "mgp_simple_pofk: no particles");
  REQUIRE(c.P == 1 && !c.slab, MGP_ERR_INVALID, "mgp_simple_pofk: one rank only (post-processing; TSC needs the plane on the left)");
  REQUIRE((size_t) 2 * c.N * sizeof(double) <= 200 * 1024, MGP_ERR_INVALID, "mgp_simple_pofk: Nmesh too large for the bin table");
  if (c.cfg.scale_dependent) sd_evict_block(c, 0);            // the work grid is force grid X
  c.forces_live = false;
  if (c.gbytes == 4) simple_pofk_t<float>(c, scheme, slip, pofk, nmodes); else simple_pofk_t<double>(c, scheme, slip, pofk, nmodes);
  for (int b = 0; b < c.N; b++) {
    if (nmodes[b] > 0) pofk[b] /= nmodes[b];                  // main.cpp:412-416
    if (subtract_shotnoise) pofk[b] -= 1.0 / (double) c.np;   // main.cpp:420-424: 1 / npart_tot
  }
}

}  // namespace mgp
