// Host emulation of the mixed-radix fused x-FFT + slab-exchange kernels (csrc/xfft_mixed.cuh), same scheme as
// xfft_emul.cu: the phase functions the kernels consist of, thread by thread, several emulated ranks, against a direct
// DFT in long double.  tests/test_xfft_host.py builds and runs it.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "xfft_mixed.cuh"

using namespace mgp;
using namespace mgp::xfm;

template <typename C, int N> std::vector<C> twiddles() {
  std::vector<C> tw(plan_twtotal(N));
  for (int i = 0; i < plan_npass(N); i++) {
    const int L = plan_R(N, i) * plan_M(N, i);
    for (int t = 0; t < L; t++) {
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) t / (long double) L;
      tw[plan_twoff(N, i) + t] = mk<C>((typename RealOf<C>::type) cosl(a), (typename RealOf<C>::type) sinl(a));
    }
  }
  return tw;
}

static double frand() { return (double) rand() / RAND_MAX - 0.5; }

template <typename C, int N, int TK, bool FWD> static void run_passes(C *smem, const C *tw, int nthr) {
#define P(I)                                                                                       \
  if constexpr (plan_npass(N) > (I))                                                                \
    for (int tid = 0; tid < nthr; tid++) one_pass<N, (I), FWD, TK, C>(smem, tw, tid, nthr);
  if (FWD) { P(4) P(3) P(2) P(1) P(0) } else { P(0) P(1) P(2) P(3) P(4) }
#undef P
}

template <typename C, int N, int TK>
static int run_case(int P, int NY, int NZ, int nthr, double tol) {
  static_assert(supported(N), "unsupported N");
  const int nxb = N / P, nyl = NY / P;
  std::vector<C> tw = twiddles<C, N>();
  std::vector<long double> CT(N), ST(N);            // the reference DFT's own cos / sin table, long double
  for (int t = 0; t < N; t++) {
    const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) t / (long double) N;
    CT[t] = cosl(a); ST[t] = sinl(a);
  }
  std::vector<C> smem((size_t) TK * N);
  constexpr int A = 128 / (int) sizeof(C);
  const int ktiles = tiles_per_line(NZ, TK, A);
  int bad = 0;
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
        const int jl = t / ktiles, k0 = tile_k0(r * nyl + jl, NZ, t - jl * ktiles, TK, A);
        if (k0 >= NZ || k0 + TK <= 0) continue;
        for (int tid = 0; tid < nthr; tid++) phase_load_lines<N, TK, C>(smem.data(), lines[r].data(), NZ, jl, k0, tid, nthr);
        run_passes<C, N, TK, false>(smem.data(), tw.data(), nthr);
        for (int tid = 0; tid < nthr; tid++) phase_store_owners<N, TK, C>(smem.data(), pp, nxb, r * nyl, NY, NZ, jl, k0, tid, nthr);
      }
    double emax = 0, vmax = 0;
    for (int r = 0; r < P; r++)
      for (int jl = 0; jl < nyl; jl++)
        for (int k = 0; k < NZ; k++) {
          const C *in = &lines[r][((size_t) jl * NZ + k) * N];
          for (int x = 0; x < N; x++) {
            long double sr = 0, si = 0;
            for (int n = 0; n < N; n++) {
              const long double c = CT[((long long) n * x) % N], s = ST[((long long) n * x) % N];
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
        const int jl = t / ktiles, k0 = tile_k0(r * nyl + jl, NZ, t - jl * ktiles, TK, A);
        if (k0 >= NZ || k0 + TK <= 0) continue;
        for (int tid = 0; tid < nthr; tid++) phase_load_owners<N, TK, C>(smem.data(), pp, nxb, r * nyl, NY, NZ, jl, k0, tid, nthr);
        run_passes<C, N, TK, true>(smem.data(), tw.data(), nthr);
        for (int tid = 0; tid < nthr; tid++) phase_store_lines<N, TK, C>(smem.data(), out[r].data(), NZ, jl, k0, tid, nthr);
      }
    double emax = 0, vmax = 0;
    for (int r = 0; r < P; r++)
      for (int jl = 0; jl < nyl; jl++)
        for (int k = 0; k < NZ; k++)
          for (int f = 0; f < N; f++) {
            long double sr = 0, si = 0;
            for (int n = 0; n < N; n++) {
              const C v = src[n / nxb][((size_t) (n % nxb) * NY + (r * nyl + jl)) * NZ + k];
              const long double c = CT[((long long) n * f) % N], s = -ST[((long long) n * f) % N];
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

template <int N> static int check_digits() {
  int prod = 1;
  for (int i = 0; i < plan_npass(N); i++) prod *= plan_R(N, i);
  if (prod != N) { printf("plan product mismatch N=%d (%d)\n", N, prod); return 1; }
  std::vector<int> seen(N, 0);
  for (int p = 0; p < N; p++) {
    const int f = digit_rev<N>(p);
    if (f < 0 || f >= N || seen[f]++) { printf("digit_rev not a permutation N=%d\n", N); return 1; }
    if (digit_rev_inv<N>(f) != p) { printf("digit_rev_inv mismatch N=%d p=%d\n", N, p); return 1; }
  }
  return 0;
}

#define CASE_LIB(C, N, P, NY, NZ, TOL) bad += run_case<C, N, tile_lines(N, sizeof(C))>(P, NY, NZ, 256, TOL)

int main() {
  int bad = 0;
  srand(4321);
  bad += check_digits<320>() + check_digits<400>() + check_digits<384>() + check_digits<640>() + check_digits<800>() +
         check_digits<48>() + check_digits<80>() + check_digits<1280>() + check_digits<256>();
  const double td = 2e-14, tf = 2e-5;
  CASE_LIB(double2, 320, 2, 2, 9, td);       // the 2-GPU weak-scaling mesh
  CASE_LIB(double2, 400, 4, 4, 9, td);       // the 4-GPU weak-scaling mesh
  CASE_LIB(double2, 48, 3, 3, 25, td);
  CASE_LIB(double2, 80, 5, 5, 41, td);
  CASE_LIB(double2, 384, 2, 2, 3, td);
  CASE_LIB(double2, 640, 2, 2, 2, td);
  CASE_LIB(double2, 800, 4, 4, 1, td);
  CASE_LIB(double2, 256, 2, 2, 5, td);       // powers of two work too
  CASE_LIB(float2, 320, 4, 4, 17, tf);
  CASE_LIB(float2, 400, 2, 2, 9, tf);
  bad += run_case<double2, 320, 4>(1, 1, 5, 96, td);
  bad += run_case<double2, 400, 16>(2, 2, 17, 64, td);
  printf(bad ? "FAILED (%d)\n" : "ALL OK\n", bad);
  return bad ? 1 : 0;
}
