// Host emulation of the fused x-FFT + slab-exchange kernels (csrc/xfft.cuh): runs the very phase functions the
// kernels are made of, thread by thread with a barrier between phases, for P emulated ranks in one process, and
// compares with a direct O(N^2) DFT in long double.  Built with nvcc and executed on the CPU (no kernel launch):
// tests/test_xfft_host.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "xfft.cuh"

using namespace mgp;
using namespace mgp::xf;

template <typename C> std::vector<C> twiddles(const Plan &pl) {
  std::vector<C> tw(pl.twtotal);
  for (int i = 0; i < pl.npass; i++) {
    const int L = pl.R[i] << pl.lgM[i];
    for (int t = 0; t < L; t++) {
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) t / (long double) L;
      tw[pl.twoff[i] + t] = mk<C>((typename RealOf<C>::type) cosl(a), (typename RealOf<C>::type) sinl(a));
    }
  }
  return tw;
}

static double frand() { return (double) rand() / RAND_MAX - 0.5; }

// global field A[x][ky][kz] (x < N, ky < NY, kz < NZ); rank r owns x-planes [r*nxb, (r+1)*nxb) in "real-side" layout
// [xl][NY][NZ] and ky-rows [r*nyl, (r+1)*nyl) in the transposed layout [jl][NZ][N].
template <typename C, int TK>
static int run_case(int N, int P, int NY, int NZ, int nthr, double tol) {
  Plan pl;
  if (!make_plan(N, pl)) { printf("no plan for N=%d\n", N); return 1; }
  const int nxb = N / P, nyl = NY / P;
  std::vector<C> tw = twiddles<C>(pl);
  std::vector<C> smem((size_t) TK * N);
  const int ktiles = (NZ + TK - 1) / TK;
  int bad = 0;
  // ---------------- backward: transposed k-space lines -> owners of x, sign +
  {
    std::vector<std::vector<C>> lines(P), land(P);
    for (int r = 0; r < P; r++) {
      lines[r].resize((size_t) nyl * NZ * N);
      for (auto &v : lines[r]) v = mk<C>((typename RealOf<C>::type) frand(), (typename RealOf<C>::type) frand());
      land[r].assign((size_t) nxb * NY * NZ, mk<C>(777, 777));
    }
    PeerPtrs pp;
    for (int r = 0; r < 16; r++) pp.p[r] = r < P ? land[r].data() : nullptr;
    for (int r = 0; r < P; r++)
      for (int t = 0; t < nyl * ktiles; t++) {
        const int jl = t / ktiles, k0 = (t - jl * ktiles) * TK;
        for (int tid = 0; tid < nthr; tid++) phase_load_lines<TK, C>(smem.data(), lines[r].data(), N, NZ, jl, k0, tid, nthr);
        for (int i = 0; i < pl.npass; i++)
          for (int tid = 0; tid < nthr; tid++) phase_pass<+1, false, TK, C>(smem.data(), pl, tw.data(), i, tid, nthr);
        for (int tid = 0; tid < nthr; tid++) phase_store_owners<TK, C>(smem.data(), pp, pl, nxb, r * nyl, NY, NZ, jl, k0, tid, nthr);
      }
    double emax = 0, vmax = 0;
    for (int r = 0; r < P; r++)
      for (int jl = 0; jl < nyl; jl++)
        for (int k = 0; k < NZ; k++) {
          const C *in = &lines[r][((size_t) jl * NZ + k) * N];
          for (int x = 0; x < N; x++) {
            long double sr = 0, si = 0;
            for (int n = 0; n < N; n++) {
              const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) (((long long) n * x) % N) / N;
              const long double c = cosl(a), s = sinl(a);
              sr += in[n].x * c - in[n].y * s; si += in[n].x * s + in[n].y * c;
            }
            const int o = x / nxb, xl = x % nxb;
            const C got = land[o][((size_t) xl * NY + (r * nyl + jl)) * NZ + k];
            const double e = fmax(fabs((double) (got.x - sr)), fabs((double) (got.y - si)));
            if (e > emax) emax = e;
            const double m = fmax(fabsl(sr), fabsl(si));
            if (m > vmax) vmax = m;
          }
        }
    const double rel = emax / vmax;
    if (!(rel < tol)) { printf("BWD FAIL N=%d P=%d TK=%d nthr=%d rel=%g\n", N, P, TK, nthr, rel); bad++; }
    else printf("bwd ok   N=%4d P=%d TK=%2d nthr=%3d rel=%.2e\n", N, P, TK, nthr, rel);
  }
  // ---------------- forward: owners of x -> transposed k-space lines, sign -
  {
    std::vector<std::vector<C>> src(P), out(P);
    for (int r = 0; r < P; r++) {
      src[r].resize((size_t) nxb * NY * NZ);
      for (auto &v : src[r]) v = mk<C>((typename RealOf<C>::type) frand(), (typename RealOf<C>::type) frand());
      out[r].assign((size_t) nyl * NZ * N, mk<C>(777, 777));
    }
    PeerPtrs pp;
    for (int r = 0; r < 16; r++) pp.p[r] = r < P ? src[r].data() : nullptr;
    for (int r = 0; r < P; r++)
      for (int t = 0; t < nyl * ktiles; t++) {
        const int jl = t / ktiles, k0 = (t - jl * ktiles) * TK;
        for (int tid = 0; tid < nthr; tid++) phase_load_owners<TK, C>(smem.data(), pp, pl, nxb, r * nyl, NY, NZ, jl, k0, tid, nthr);
        for (int i = pl.npass - 1; i >= 0; i--)
          for (int tid = 0; tid < nthr; tid++) phase_pass<-1, true, TK, C>(smem.data(), pl, tw.data(), i, tid, nthr);
        for (int tid = 0; tid < nthr; tid++) phase_store_lines<TK, C>(smem.data(), out[r].data(), N, NZ, jl, k0, tid, nthr);
      }
    double emax = 0, vmax = 0;
    for (int r = 0; r < P; r++)
      for (int jl = 0; jl < nyl; jl++)
        for (int k = 0; k < NZ; k++)
          for (int f = 0; f < N; f++) {
            long double sr = 0, si = 0;
            for (int n = 0; n < N; n++) {
              const C v = src[n / nxb][((size_t) (n % nxb) * NY + (r * nyl + jl)) * NZ + k];
              const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double) (((long long) n * f) % N) / N;
              const long double c = cosl(a), s = sinl(a);
              sr += v.x * c - v.y * s; si += v.x * s + v.y * c;
            }
            const C got = out[r][((size_t) jl * NZ + k) * N + f];
            const double e = fmax(fabs((double) (got.x - sr)), fabs((double) (got.y - si)));
            if (e > emax) emax = e;
            const double m = fmax(fabsl(sr), fabsl(si));
            if (m > vmax) vmax = m;
          }
    const double rel = emax / vmax;
    if (!(rel < tol)) { printf("FWD FAIL N=%d P=%d TK=%d nthr=%d rel=%g\n", N, P, TK, nthr, rel); bad++; }
    else printf("fwd ok   N=%4d P=%d TK=%2d nthr=%3d rel=%.2e\n", N, P, TK, nthr, rel);
  }
  return bad;
}

int main() {
  int bad = 0;
  srand(12345);
  // digit reversal is a permutation and its own inverse map
  for (int N = 16; N <= 4096; N *= 2) {
    Plan pl; make_plan(N, pl);
    int prod = 1; for (int i = 0; i < pl.npass; i++) prod *= pl.R[i];
    if (prod != N) { printf("plan product mismatch N=%d\n", N); bad++; }
    std::vector<int> seen(N, 0);
    for (int p = 0; p < N; p++) {
      const int f = digit_rev(pl, p);
      if (f < 0 || f >= N || seen[f]++) { printf("digit_rev not a permutation N=%d\n", N); bad++; break; }
      if (digit_rev_inv(pl, f) != p) { printf("digit_rev_inv mismatch N=%d p=%d\n", N, p); bad++; break; }
    }
  }
  const double td = 2e-14, tf = 2e-5;
  bad += run_case<double2, 16>(16, 2, 4, 9, 256, td);
  bad += run_case<double2, 8>(32, 1, 2, 3, 64, td);
  bad += run_case<double2, 4>(64, 2, 2, 5, 256, td);
  bad += run_case<double2, 16>(128, 4, 4, 17, 256, td);
  bad += run_case<double2, 16>(256, 8, 8, 9, 256, td);
  bad += run_case<double2, 8>(512, 2, 2, 3, 256, td);
  bad += run_case<double2, 4>(1024, 8, 8, 2, 256, td);
  bad += run_case<double2, 8>(1024, 1, 1, 2, 96, td);
  bad += run_case<double2, 4>(2048, 2, 2, 1, 256, td);
  bad += run_case<float2, 16>(256, 2, 2, 17, 256, tf);
  bad += run_case<float2, 8>(1024, 4, 4, 2, 256, tf);
  bad += run_case<float2, 16>(64, 1, 1, 33, 32, tf);
  printf(bad ? "FAILED (%d)\n" : "ALL OK\n", bad);
  return bad ? 1 : 0;
}
