// Mixed-radix version of the fused x-transform + slab exchange (xfft.cuh) for mesh sizes N = 2^a 3^b 5^c that are not
// powers of two -- the 2- and 4-GPU weak-scaling sizes 320 = 2^6 5 and 400 = 2^4 5^2.  Same structure as xfft.cuh (tile
// of TK lines in swizzled shared memory, decimation in frequency backward / the transposed graph forward, the digit
// reversal absorbed by the exchange side, runs aligned to 128-byte lines); what changes is that strides are no longer
// powers of two (constant divisions instead of shifts), that the radix set gains 5 and 3, and that the owner of an x-plane
// is x / nxb.  OPT-IN (MGP_XFFT_MIXED=1): verified on the CPU by tests/host/xfft_mixed_emul.cu (thread-by-thread
// emulation against a direct DFT), parity-checked on the GPU (tests/test_slab_fused.py), not yet timed on several ranks; the default for these sizes stays cuFFT's 1-D plan + the
// transpose kernels.
#pragma once

#include "xfft.cuh"

namespace mgp {
namespace xfm {

using xf::cadd; using xf::cmul; using xf::csub; using xf::kBatch; using xf::kThreads; using xf::mk; using xf::peer_load;
using xf::RealOf; using xf::sidx; using xf::tile_k0; using xf::tiles_per_line; using xf::tw_load;

constexpr int kMaxPass = 5;

// ------------------------------------------------------------------ the pass plan of N (compile time)

__host__ __device__ constexpr int next_radix(int rem) {
  if (rem % 16 == 0 && rem != 32) return 16;
  if (rem % 8 == 0) return 8;
  if (rem % 4 == 0) return 4;
  if (rem % 2 == 0) return 2;
  if (rem % 5 == 0) return 5;
  if (rem % 3 == 0) return 3;
  return 0;                                     // not 2^a 3^b 5^c
}
__host__ __device__ constexpr int plan_R(int n, int i) {
  int rem = n, r = 0;
  for (int p = 0; p <= i; p++) {
    if (rem == 1) return 1;
    r = next_radix(rem);
    if (r == 0) return 0;
    rem /= r;
  }
  return r;
}
__host__ __device__ constexpr int plan_M(int n, int i) {      // length of the sub-transforms pass i leaves behind
  int rem = n;
  for (int p = 0; p <= i; p++) { const int r = plan_R(n, p); if (r <= 1) return r == 1 ? 1 : 0; rem /= r; }
  return rem;
}
__host__ __device__ constexpr int plan_npass(int n) {
  int i = 0;
  while (i < kMaxPass + 1 && plan_M(n, i) > 1) i++;
  return i + 1;
}
__host__ __device__ constexpr bool supported(int n) {
  if (n < 16 || n > 4096) return false;
  int rem = n, np = 0;
  while (rem > 1) { const int r = next_radix(rem); if (r == 0) return false; rem /= r; np++; }
  return np <= kMaxPass;
}
__host__ __device__ constexpr int plan_twoff(int n, int i) {   // exp(+2 pi i t / L_i), t < L_i = R_i M_i
  int off = 0;
  for (int p = 0; p < i; p++) off += plan_R(n, p) * plan_M(n, p);
  return off;
}
__host__ __device__ constexpr int plan_twtotal(int n) { return plan_twoff(n, plan_npass(n)); }
__host__ __device__ constexpr int tile_lines(int n, int cbytes) {
  int tk = 16;
  while (tk > 4 && (size_t) tk * n * cbytes > 64 * 1024) tk >>= 1;
  return ((size_t) tk * n * cbytes > 200 * 1024) ? 0 : tk;
}

// position p of the decimation-in-frequency output holds frequency digit_rev(p): p = sum q_i M_i, f = sum q_i prod_{j<i} R_j
// (template recursion over the passes, so that every divisor is a compile-time constant)
__host__ __device__ constexpr int plan_weight(int n, int i) {
  int w = 1;
  for (int p = 0; p < i; p++) w *= plan_R(n, p);
  return w;
}
template <int N, int I, bool END = (I >= plan_npass(N))> struct Digits {
  static XF_HD int rev(int p) { return ((p / plan_M(N, I)) % plan_R(N, I)) * plan_weight(N, I) + Digits<N, I + 1>::rev(p); }
  static XF_HD int inv(int x) { return ((x / plan_weight(N, I)) % plan_R(N, I)) * plan_M(N, I) + Digits<N, I + 1>::inv(x); }
};
template <int N, int I> struct Digits<N, I, true> {
  static XF_HD int rev(int) { return 0; }
  static XF_HD int inv(int) { return 0; }
};
template <int N> XF_HD int digit_rev(int p) { return Digits<N, 0>::rev(p); }
template <int N> XF_HD int digit_rev_inv(int x) { return Digits<N, 0>::inv(x); }

// ------------------------------------------------------------------ radix-R DFTs in registers (natural order in and out)

template <int R, int SIGN, typename C> struct DftM {
  static XF_HD void run(C *v) { xf::Dft<R, SIGN, C>::run(v); }        // 2, 4, 8, 16
};
template <int SIGN, typename C> struct DftM<5, SIGN, C> {
  static XF_HD void run(C *v) {
    typedef typename RealOf<C>::type T;
    const T c1 = (T) 0.30901699437494742410, c2 = (T) -0.80901699437494742410;     // cos(2 pi / 5), cos(4 pi / 5)
    const T s1 = (T) 0.95105651629515357212, s2 = (T) 0.58778525229247312917;      // sin(2 pi / 5), sin(4 pi / 5)
    const C t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
    const C m1 = mk<C>(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
    const C m2 = mk<C>(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
    const C n1 = mk<C>(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    const C n2 = mk<C>(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    // i * SIGN * n = (-SIGN n.y, SIGN n.x)
    const C j1 = SIGN > 0 ? mk<C>(-n1.y, n1.x) : mk<C>(n1.y, -n1.x);
    const C j2 = SIGN > 0 ? mk<C>(-n2.y, n2.x) : mk<C>(n2.y, -n2.x);
    v[0] = mk<C>(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
    v[1] = cadd(m1, j1); v[4] = csub(m1, j1);
    v[2] = cadd(m2, j2); v[3] = csub(m2, j2);
  }
};
template <int SIGN, typename C> struct DftM<3, SIGN, C> {
  static XF_HD void run(C *v) {
    typedef typename RealOf<C>::type T;
    const T s = (T) 0.86602540378443864676;                                         // sin(2 pi / 3); cos = -1/2
    const C t1 = cadd(v[1], v[2]), t2 = csub(v[1], v[2]);
    const C m = mk<C>(v[0].x - (T) 0.5 * t1.x, v[0].y - (T) 0.5 * t1.y);
    const C n = mk<C>(s * t2.x, s * t2.y);
    const C j = SIGN > 0 ? mk<C>(-n.y, n.x) : mk<C>(n.y, -n.x);
    v[0] = cadd(v[0], t1);
    v[1] = cadd(m, j); v[2] = csub(m, j);
  }
};
template <int SIGN, typename C> struct DftM<1, SIGN, C> {
  static XF_HD void run(C *) {}
};

// ------------------------------------------------------------------ passes on the shared-memory tile

template <int N, int I, int SIGN, bool DIT, int TK, typename C>
XF_HD void butterfly(C *s, const C *twtab, int bb, int k) {
  constexpr int R = plan_R(N, I), M = plan_M(N, I);
  const C *tw = twtab + plan_twoff(N, I);
  const int blk = bb / M, b = bb - blk * M;
  const int base = blk * (R * M) + b;
  C v[R];
#pragma unroll
  for (int j = 0; j < R; j++) v[j] = s[sidx<TK>(base + j * M, k)];
  if (DIT && M > 1) {
#pragma unroll
    for (int j = 1; j < R; j++) {
      C t = tw_load(tw + b * j);
      if (SIGN < 0) t.y = -t.y;
      v[j] = cmul(v[j], t);
    }
  }
  DftM<R, SIGN, C>::run(v);
  if (!DIT && M > 1) {
#pragma unroll
    for (int q = 1; q < R; q++) {
      C t = tw_load(tw + b * q);
      if (SIGN < 0) t.y = -t.y;
      v[q] = cmul(v[q], t);
    }
  }
#pragma unroll
  for (int q = 0; q < R; q++) s[sidx<TK>(base + q * M, k)] = v[q];
}

template <int N, int I, int SIGN, bool DIT, int TK, typename C>
XF_HD void phase_pass(C *s, const C *twtab, int tid, int nthr) {
  constexpr int items = TK * (N / plan_R(N, I));
  const int k = tid & (TK - 1);
  for (int w = tid; w < items; w += nthr) butterfly<N, I, SIGN, DIT, TK, C>(s, twtab, w / TK, k);
}

// passes 0 .. NP-1 in decimation-in-frequency order (FWD = false) or NP-1 .. 0 as the transposed graph (FWD = true);
// SYNC() between passes is the caller's barrier (kernel: __syncthreads; host emulation: runs pass by pass itself)
template <int N, int I, bool FWD, int TK, typename C>
XF_HD void one_pass(C *s, const C *twtab, int tid, int nthr) {
  if (FWD) phase_pass<N, I, -1, true, TK, C>(s, twtab, tid, nthr);
  else phase_pass<N, I, +1, false, TK, C>(s, twtab, tid, nthr);
}

// ------------------------------------------------------------------ global side (as xfft.cuh, general N and slab width)

template <int N, int TK, typename C>
XF_HD void phase_load_lines(C *s, const C *__restrict__ in, int NZ, int jl, int k0, int tid, int nthr) {
  constexpr int tot = TK * N;
  const C *src = in + ((long long) jl * NZ + k0) * N;
  for (int e0 = tid; e0 < tot; e0 += kBatch * nthr) {
    C v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int e = e0 + u * nthr;
      v[u] = mk<C>(0, 0);
      if (e < tot && k0 + e / N < NZ && k0 + e / N >= 0) v[u] = src[e];
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int e = e0 + u * nthr;
      if (e < tot) s[sidx<TK>(e % N, e / N)] = v[u];
    }
  }
}

template <int N, int TK, typename C>
XF_HD void phase_store_owners(const C *s, const PeerPtrs &out, int nxb, int y0, int NY, int NZ, int jl, int k0, int tid,
                              int nthr) {
  constexpr int tot = TK * N;
  const int k = tid & (TK - 1);
  if (k0 + k >= NZ || k0 + k < 0) return;
  const size_t col = (size_t) (y0 + jl) * NZ + (k0 + k);
  const size_t xstride = (size_t) NY * NZ;
  for (int e = tid; e < tot; e += nthr) {
    const int p = e / TK;
    const int x = digit_rev<N>(p);
    const int r = x / nxb, xl = x - r * nxb;
    ((C *) out.p[r])[(size_t) xl * xstride + col] = s[sidx<TK>(p, k)];
  }
}

template <int N, int TK, typename C>
XF_HD void phase_load_owners(C *s, const PeerPtrs &in, int nxb, int y0, int NY, int NZ, int jl, int k0, int tid, int nthr) {
  constexpr int tot = TK * N;
  const int k = tid & (TK - 1);
  const bool live = k0 + k < NZ && k0 + k >= 0;
  const size_t col = (size_t) ((long long) (y0 + jl) * NZ + (k0 + k));
  const size_t xstride = (size_t) NY * NZ;
  for (int e0 = tid; e0 < tot; e0 += kBatch * nthr) {
    C v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int e = e0 + u * nthr;
      v[u] = mk<C>(0, 0);
      if (e < tot && live) {
        const int x = e / TK;
        const int r = x / nxb, xl = x - r * nxb;
        v[u] = peer_load((const C *) in.p[r] + (size_t) xl * xstride + col);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int e = e0 + u * nthr;
      if (e < tot) s[sidx<TK>(digit_rev_inv<N>(e / TK), k)] = v[u];
    }
  }
}

template <int N, int TK, typename C>
XF_HD void phase_store_lines(const C *s, C *__restrict__ out, int NZ, int jl, int k0, int tid, int nthr) {
  constexpr int tot = TK * N;
  C *dst = out + ((long long) jl * NZ + k0) * N;
  for (int e = tid; e < tot; e += nthr) {
    const int k = e / N, x = e - k * N;
    if (k0 + k < NZ && k0 + k >= 0) dst[e] = s[sidx<TK>(x, k)];
  }
}

// ------------------------------------------------------------------ kernels

#if defined(__CUDACC__)

#define XFM_PASS(I, FWD)                                                     \
  if constexpr (plan_npass(N) > (I)) {                                        \
    one_pass<N, (I), FWD, TK, C>(s, twtab, threadIdx.x, kThreads);            \
    __syncthreads();                                                          \
  }

template <typename C, int N, int TK>
__global__ void __launch_bounds__(kThreads, 2)
k_xfft_bwd_p2p(const C *__restrict__ in, const __grid_constant__ PeerPtrs out, const C *__restrict__ twtab, int nxb, int y0,
               int NY, int NZ, int nyl) {
  extern __shared__ __align__(16) unsigned char xf_smem[];
  C *s = reinterpret_cast<C *>(xf_smem);
  constexpr int A = 128 / (int) sizeof(C);
  const int ktiles = tiles_per_line(NZ, TK, A), ntiles = nyl * ktiles;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int jl = t / ktiles, k0 = tile_k0(y0 + jl, NZ, t - jl * ktiles, TK, A);
    if (k0 >= NZ || k0 + TK <= 0) continue;
    phase_load_lines<N, TK, C>(s, in, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
    XFM_PASS(0, false) XFM_PASS(1, false) XFM_PASS(2, false) XFM_PASS(3, false) XFM_PASS(4, false)
    phase_store_owners<N, TK, C>(s, out, nxb, y0, NY, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
  }
}

template <typename C, int N, int TK>
__global__ void __launch_bounds__(kThreads, 2)
k_xfft_fwd_p2p(const __grid_constant__ PeerPtrs in, C *__restrict__ out, const C *__restrict__ twtab, int nxb, int y0, int NY,
               int NZ, int nyl) {
  extern __shared__ __align__(16) unsigned char xf_smem[];
  C *s = reinterpret_cast<C *>(xf_smem);
  constexpr int A = 128 / (int) sizeof(C);
  const int ktiles = tiles_per_line(NZ, TK, A), ntiles = nyl * ktiles;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int jl = t / ktiles, k0 = tile_k0(y0 + jl, NZ, t - jl * ktiles, TK, A);
    if (k0 >= NZ || k0 + TK <= 0) continue;
    phase_load_owners<N, TK, C>(s, in, nxb, y0, NY, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
    XFM_PASS(4, true) XFM_PASS(3, true) XFM_PASS(2, true) XFM_PASS(1, true) XFM_PASS(0, true)
    phase_store_lines<N, TK, C>(s, out, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
  }
}

#undef XFM_PASS

#endif  // __CUDACC__

}  // namespace xfm
}  // namespace mgp
