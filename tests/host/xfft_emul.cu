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

template <typename C, int LGN> std::vector<C> twiddles() {
  std::vector<C> tw(plan_twtotal(LGN));
  for (int i = 0; i < plan_npass(LGN); i++) {
    const int L = 1 << (plan_lgR(LGN, i) + plan_lgM(LGN, i));
    for (int t = 0; t < L; t++) {
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) t / (long double) L;
      tw[plan_twoff(LGN, i) + t] = mk<C>((typename RealOf<C>::type) cosl(a), (typename RealOf<C>::type) sinl(a));
    }
  }
  return tw;
}

static double frand() { return (double) rand() / RAND_MAX - 0.5; }

// global field A[x][ky][kz] (x < N, ky < NY, kz < NZ); rank r owns x-planes [r*nxb, (r+1)*nxb) in "real-side" layout
// [xl][NY][NZ] and ky-rows [r*nyl, (r+1)*nyl) in the transposed layout [jl][NZ][N].
template <typename C, int LGN, int TK>
static int run_case(int P, int NY, int NZ, int nthr, double tol) {
  constexpr int N = 1 << LGN;
  const int nxb = N / P, nyl = NY / P;
  int lgx = 0; while ((1 << lgx) < nxb) lgx++;
  std::vector<C> tw = twiddles<C, LGN>();
  std::vector<long double> CT(N), ST(N);            // the reference DFT's own cos / sin table, long double
  for (int t = 0; t < N; t++) {
    const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) t / (long double) N;
    CT[t] = cosl(a); ST[t] = sinl(a);
  }
  std::vector<C> smem((size_t) TK * N);
  constexpr int A = 128 / (int) sizeof(C);
  const int ktiles = tiles_per_line(NZ, TK, A);
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
        const int jl = t / ktiles, k0 = tile_k0(r * nyl + jl, NZ, t - jl * ktiles, TK, A);
        if (k0 >= NZ || k0 + TK <= 0) continue;
        for (int tid = 0; tid < nthr; tid++) phase_load_lines<LGN, TK, C>(smem.data(), lines[r].data(), NZ, jl, k0, tid, nthr);
        for (int tid = 0; tid < nthr; tid++) phase_pass<LGN, 0, +1, false, TK, C>(smem.data(), tw.data(), tid, nthr);
        if constexpr (plan_npass(LGN) > 1)
          for (int tid = 0; tid < nthr; tid++) phase_pass<LGN, 1, +1, false, TK, C>(smem.data(), tw.data(), tid, nthr);
        if constexpr (plan_npass(LGN) > 2)
          for (int tid = 0; tid < nthr; tid++) phase_pass<LGN, 2, +1, false, TK, C>(smem.data(), tw.data(), tid, nthr);
        for (int tid = 0; tid < nthr; tid++) phase_store_owners<LGN, TK, C>(smem.data(), pp, lgx, r * nyl, NY, NZ, jl, k0, tid, nthr);
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
        const int jl = t / ktiles, k0 = tile_k0(r * nyl + jl, NZ, t - jl * ktiles, TK, A);
        if (k0 >= NZ || k0 + TK <= 0) continue;
        for (int tid = 0; tid < nthr; tid++) phase_load_owners<LGN, TK, C>(smem.data(), pp, lgx, r * nyl, NY, NZ, jl, k0, tid, nthr);
        if constexpr (plan_npass(LGN) > 2)
          for (int tid = 0; tid < nthr; tid++) phase_pass<LGN, 2, -1, true, TK, C>(smem.data(), tw.data(), tid, nthr);
        if constexpr (plan_npass(LGN) > 1)
          for (int tid = 0; tid < nthr; tid++) phase_pass<LGN, 1, -1, true, TK, C>(smem.data(), tw.data(), tid, nthr);
        for (int tid = 0; tid < nthr; tid++) phase_pass<LGN, 0, -1, true, TK, C>(smem.data(), tw.data(), tid, nthr);
        for (int tid = 0; tid < nthr; tid++) phase_store_lines<LGN, TK, C>(smem.data(), out[r].data(), NZ, jl, k0, tid, nthr);
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

template <int LGN> static int check_digits() {
  constexpr int N = 1 << LGN;
  int prod = 0;
  for (int i = 0; i < plan_npass(LGN); i++) prod += plan_lgR(LGN, i);
  if (prod != LGN) { printf("plan product mismatch N=%d\n", N); return 1; }
  std::vector<int> seen(N, 0);
  for (int p = 0; p < N; p++) {
    const int f = digit_rev<LGN>(p);
    if (f < 0 || f >= N || seen[f]++) { printf("digit_rev not a permutation N=%d\n", N); return 1; }
    if (digit_rev_inv<LGN>(f) != p) { printf("digit_rev_inv mismatch N=%d p=%d\n", N, p); return 1; }
  }
  return 0;
}

// the instance the library launches for this size and precision (tile_lines), 256 threads, plus odd thread counts
#define CASE_LIB(C, LGN, P, NY, NZ, TOL) bad += run_case<C, LGN, tile_lines(LGN, sizeof(C))>(P, NY, NZ, 256, TOL)

int main() {
  int bad = 0;
  srand(12345);
  bad += check_digits<4>() + check_digits<5>() + check_digits<6>() + check_digits<7>() + check_digits<8>() +
         check_digits<9>() + check_digits<10>() + check_digits<11>() + check_digits<12>();
  const double td = 2e-14, tf = 2e-5;
  CASE_LIB(double2, 4, 2, 4, 9, td);
  CASE_LIB(double2, 5, 1, 2, 17, td);
  CASE_LIB(double2, 6, 4, 4, 33, td);
  CASE_LIB(double2, 7, 8, 8, 5, td);
  CASE_LIB(double2, 8, 2, 2, 17, td);
  CASE_LIB(double2, 9, 8, 8, 9, td);
  CASE_LIB(double2, 10, 4, 4, 3, td);
  CASE_LIB(double2, 11, 2, 2, 1, td);
  CASE_LIB(float2, 4, 1, 1, 9, tf);
  CASE_LIB(float2, 7, 2, 2, 17, tf);
  CASE_LIB(float2, 9, 4, 4, 9, tf);
  CASE_LIB(float2, 10, 8, 8, 5, tf);
  CASE_LIB(float2, 12, 2, 2, 1, tf);
  // the dimensions of the 8-GPU runs: Nmesh 512 (NZ = 257) and 1024 (NZ = 513) on 8 ranks, ky = 0 .. 7 (every offset of a
  // row against the 128-byte lines of the owner's buffer)
  CASE_LIB(double2, 9, 8, 8, 257, td);
  CASE_LIB(double2, 10, 8, 8, 513, td);
  // the strong-scaling run of the bench at 512^3 on 4 ranks (2 and 8 ranks have run on the GPUs)
  CASE_LIB(double2, 9, 4, 4, 257, td);
  // the wide-tile instances (xfft_wide.cu)
  bad += run_case<double2, 8, tile_lines_wide(8, 16)>(2, 2, 129, 256, td);
  bad += run_case<double2, 9, tile_lines_wide(9, 16)>(8, 8, 17, 256, td);
  bad += run_case<double2, 10, tile_lines_wide(10, 16)>(4, 4, 9, 256, td);
  bad += run_case<float2, 9, tile_lines_wide(9, 8)>(2, 2, 33, 256, tf);
  // other tile widths and thread counts than the library's
  bad += run_case<double2, 5, 8>(1, 2, 3, 64, td);
  bad += run_case<double2, 6, 4>(2, 2, 5, 96, td);
  bad += run_case<double2, 9, 4>(2, 2, 5, 128, td);
  bad += run_case<double2, 8, 8>(4, 4, 9, 32, td);
  printf(bad ? "FAILED (%d)\n" : "ALL OK\n", bad);
  return bad ? 1 : 0;
}
