// x-axis FFT of the slab-decomposed transforms, fused with the slab exchange over peer memory.
//
// The FFTW-MPI transforms the reference calls (wrappers.c:26-96) are, on x-slabs, local 2-D (y,z) transforms, a
// global transpose and 1-D transforms along x.  Here the 1-D x-transform and the transpose are ONE kernel:
//   backward (c2r direction): a CTA reads TK lines [ky][kz0..kz0+TK)[x] of the local transposed k-space, transforms them
//     along x in shared memory and stores every x straight into the buffer of the rank that owns it (NVLink stores,
//     runs of TK complex values along kz) -- x-FFT, pack, all-to-all and unpack in one pass over HBM;
//   forward (r2c direction): a CTA pulls the same tile from the owners of the x-planes (NVLink loads), transforms it
//     and writes the lines of the local transposed k-space.
//
// Transform: N = 2^LGN (a template parameter: every radix, stride, twiddle offset and digit shuffle is a compile-time
// constant), in place in shared memory, radix-16/8/4/2 passes held in registers.  Backward runs decimation in
// frequency (natural in, digit-reversed out: the scatter to the owners undoes the permutation for free); forward
// runs the transposed flow graph (decimation in time: digit-reversed in, natural out: the gather from the owners
// applies the permutation for free).  Unnormalised in both directions, like FFTW / cuFFT.
//
// Shared-memory tile: element (x, k) of the TK lines lives at row x, column k ^ (x & (TK-1)).  A butterfly touches
// whole rows (TK lanes = TK columns of one row: conflict-free for every stride); the swizzle makes the x-contiguous
// global side (lanes = consecutive x at fixed k) conflict-free as well.
//
// Everything except the kernels themselves is __host__ __device__, so that tests/host/xfft_emul.cu can run the
// very same phase functions thread by thread on the CPU against a direct DFT.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#define XF_HD __host__ __device__ __forceinline__

namespace mgp {

struct PeerPtrs { void *p[16]; };

namespace xf {

constexpr int kThreads = 256;
constexpr int kMinLgN = 4, kMaxLgN = 12;

// ------------------------------------------------------------------ the pass plan of N = 2^lgn (compile time)
// decimation-in-frequency order: pass i works on sub-transforms of length L_i = R_i * M_i (L_0 = N, L_{i+1} = M_i)

__host__ __device__ constexpr int plan_lgR(int lgn, int i) {
  int rem = lgn, r = 0;
  for (int p = 0; p <= i; p++) {
    r = rem == 5 ? 3 : (rem >= 4 ? 4 : rem);      // 32 = 8 * 4 rather than 16 * 2
    rem -= r;
  }
  return r;
}
__host__ __device__ constexpr int plan_lgM(int lgn, int i) {
  int rem = lgn;
  for (int p = 0; p <= i; p++) rem -= plan_lgR(lgn, p);
  return rem;
}
__host__ __device__ constexpr int plan_npass(int lgn) {
  int n = 0;
  while (plan_lgM(lgn, n) > 0) n++;
  return n + 1;
}
__host__ __device__ constexpr int plan_twoff(int lgn, int i) {      // exp(+2 pi i t / L_i), t < L_i, starts here in the twiddle table
  int off = 0;
  for (int p = 0; p < i; p++) off += 1 << (plan_lgR(lgn, p) + plan_lgM(lgn, p));
  return off;
}
__host__ __device__ constexpr int plan_twtotal(int lgn) { return plan_twoff(lgn, plan_npass(lgn)); }

// lines per tile: at most 64 KB of shared memory per CTA so that two CTAs share an SM (one tile's global traffic
// overlaps the other's butterflies); 0 = the tile does not fit at all
__host__ __device__ constexpr int tile_lines(int lgn, int cbytes) {
  int tk = 16;
  while (tk > 4 && ((size_t) tk << lgn) * cbytes > 64 * 1024) tk >>= 1;
  return (((size_t) tk << lgn) * cbytes > 200 * 1024) ? 0 : tk;
}

// wide tiles for the one-CTA-per-SM configuration of multi-rank runs (at most 128 KB): twice the run length on the
// exchange side (256 bytes at Nmesh 512, 128 at 1024 in double).  Opt-in (MGP_XFFT_WIDE=1, xfft_wide.cu).
__host__ __device__ constexpr int tile_lines_wide(int lgn, int cbytes) {
  int tk = 32;
  while (tk > 4 && ((size_t) tk << lgn) * cbytes > 128 * 1024) tk >>= 1;
  return (((size_t) tk << lgn) * cbytes > 200 * 1024) ? 0 : tk;
}

// k-tiles of a line are laid so that the runs of TK values the exchange side moves start on 128-byte lines of the
// owner's buffer [xl][ky][kz]: row (xl, ky) starts ((ky * NZ) mod A) elements past a line (A = elements per 128 bytes,
// NY * NZ * xl being a multiple of A), so tile kt of that row covers kz in [kt * TK - off, kt * TK - off + TK)
__host__ __device__ constexpr int tiles_per_line(int NZ, int TK, int A) { return (NZ + A - 1 + TK - 1) / TK; }
XF_HD int tile_k0(int ky, int NZ, int kt, int TK, int A) { return kt * TK - (int) (((long long) ky * NZ) & (A - 1)); }

// position p of the decimation-in-frequency output holds frequency digit_rev(p); digit_rev_inv is the inverse map
template <int LGN> XF_HD int digit_rev(int p) {
  constexpr int NP = plan_npass(LGN);
  int f = (p >> plan_lgM(LGN, 0)) & ((1 << plan_lgR(LGN, 0)) - 1);
  if (NP > 1) f |= ((p >> plan_lgM(LGN, 1)) & ((1 << plan_lgR(LGN, 1)) - 1)) << plan_lgR(LGN, 0);
  if (NP > 2) f |= ((p >> plan_lgM(LGN, 2)) & ((1 << plan_lgR(LGN, 2)) - 1)) << (plan_lgR(LGN, 0) + plan_lgR(LGN, 1));
  return f;
}
template <int LGN> XF_HD int digit_rev_inv(int x) {
  constexpr int NP = plan_npass(LGN);
  int p = (x & ((1 << plan_lgR(LGN, 0)) - 1)) << plan_lgM(LGN, 0);
  if (NP > 1) p |= ((x >> plan_lgR(LGN, 0)) & ((1 << plan_lgR(LGN, 1)) - 1)) << plan_lgM(LGN, 1);
  if (NP > 2) p |= ((x >> (plan_lgR(LGN, 0) + plan_lgR(LGN, 1))) & ((1 << plan_lgR(LGN, 2)) - 1)) << plan_lgM(LGN, 2);
  return p;
}
static_assert(plan_npass(kMaxLgN) <= 3, "digit_rev handles three passes");

// ------------------------------------------------------------------ complex helpers

template <typename C> struct RealOf;
template <> struct RealOf<double2> { typedef double type; };
template <> struct RealOf<float2> { typedef float type; };

template <typename C> XF_HD C mk(typename RealOf<C>::type x, typename RealOf<C>::type y) { C r; r.x = x; r.y = y; return r; }
template <typename C> XF_HD C cadd(const C a, const C b) { return mk<C>(a.x + b.x, a.y + b.y); }
template <typename C> XF_HD C csub(const C a, const C b) { return mk<C>(a.x - b.x, a.y - b.y); }
template <typename C> XF_HD C cmul(const C a, const C b) { return mk<C>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// cos / sin of idx * pi / 8, idx = 0..7
XF_HD double c16(int i) {
  switch (i) {
    case 0: return 1.0;
    case 1: return 0.92387953251128675613;
    case 2: return 0.70710678118654752440;
    case 3: return 0.38268343236508977173;
    case 4: return 0.0;
    case 5: return -0.38268343236508977173;
    case 6: return -0.70710678118654752440;
    default: return -0.92387953251128675613;
  }
}
XF_HD double s16(int i) {
  switch (i) {
    case 0: return 0.0;
    case 1: return 0.38268343236508977173;
    case 2: return 0.70710678118654752440;
    case 3: return 0.92387953251128675613;
    case 4: return 1.0;
    case 5: return 0.92387953251128675613;
    case 6: return 0.70710678118654752440;
    default: return 0.38268343236508977173;
  }
}

// a * exp(SIGN * 2 pi i * idx / 16), idx = 0..7 (a compile-time constant once the callers' loops are unrolled)
template <int SIGN, typename C> XF_HD C rot16(const C a, int idx) {
  typedef typename RealOf<C>::type T;
  if (idx == 0) return a;
  if (idx == 4) return SIGN > 0 ? mk<C>(-a.y, a.x) : mk<C>(a.y, -a.x);
  const T wr = (T) c16(idx), wi = (T) (SIGN > 0 ? s16(idx) : -s16(idx));
  return mk<C>(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
}

// R-point DFT in registers: v[q] <- sum_j v[j] exp(SIGN 2 pi i j q / R); natural order in and out
template <int R, int SIGN, typename C> struct Dft {
  static XF_HD void run(C *v) {
    C a[R / 2], b[R / 2];
#pragma unroll
    for (int j = 0; j < R / 2; j++) {
      a[j] = cadd(v[j], v[j + R / 2]);
      b[j] = rot16<SIGN>(csub(v[j], v[j + R / 2]), j * (16 / R));
    }
    Dft<R / 2, SIGN, C>::run(a);
    Dft<R / 2, SIGN, C>::run(b);
#pragma unroll
    for (int q = 0; q < R / 2; q++) { v[2 * q] = a[q]; v[2 * q + 1] = b[q]; }
  }
};
template <int SIGN, typename C> struct Dft<1, SIGN, C> {
  static XF_HD void run(C *) {}
};

// ------------------------------------------------------------------ shared-memory tile

template <int TK> XF_HD int sidx(int x, int k) { return x * TK + (k ^ (x & (TK - 1))); }

template <typename C> XF_HD C tw_load(const C *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// One radix-R butterfly of pass I on line k: inputs x_j = base + j * M.  DIT = false: DFT_R then twiddle
// (decimation in frequency); DIT = true: twiddle then DFT_R (the transposed graph).
template <int LGN, int I, int SIGN, bool DIT, int TK, typename C>
XF_HD void butterfly(C *s, const C *twtab, int bb, int k) {
  constexpr int LGR = plan_lgR(LGN, I), LGM = plan_lgM(LGN, I), R = 1 << LGR, M = 1 << LGM;
  const C *tw = twtab + plan_twoff(LGN, I);
  const int blk = bb >> LGM, b = bb & (M - 1);
  const int base = (blk << (LGR + LGM)) + b;
  C v[R];
  if (M >= TK) {           // x_j & (TK-1) does not depend on j: one column, constant row stride
    const C *p = s + base * TK + (k ^ (base & (TK - 1)));
#pragma unroll
    for (int j = 0; j < R; j++) v[j] = p[j * (M * TK)];
  } else {
#pragma unroll
    for (int j = 0; j < R; j++) v[j] = s[sidx<TK>(base + j * M, k)];
  }
  if (DIT && M > 1) {
#pragma unroll
    for (int j = 1; j < R; j++) {
      C t = tw_load(tw + b * j);
      if (SIGN < 0) t.y = -t.y;
      v[j] = cmul(v[j], t);
    }
  }
  Dft<R, SIGN, C>::run(v);
  if (!DIT && M > 1) {
#pragma unroll
    for (int q = 1; q < R; q++) {
      C t = tw_load(tw + b * q);
      if (SIGN < 0) t.y = -t.y;
      v[q] = cmul(v[q], t);
    }
  }
  if (M >= TK) {
    C *p = s + base * TK + (k ^ (base & (TK - 1)));
#pragma unroll
    for (int q = 0; q < R; q++) p[q * (M * TK)] = v[q];
  } else {
#pragma unroll
    for (int q = 0; q < R; q++) s[sidx<TK>(base + q * M, k)] = v[q];
  }
}

// all butterflies of pass I that belong to thread tid of nthr (nthr a multiple of TK: the line k of a thread is fixed)
template <int LGN, int I, int SIGN, bool DIT, int TK, typename C>
XF_HD void phase_pass(C *s, const C *twtab, int tid, int nthr) {
  constexpr int items = TK << (LGN - plan_lgR(LGN, I));
  const int k = tid & (TK - 1);
  for (int w = tid; w < items; w += nthr) butterfly<LGN, I, SIGN, DIT, TK, C>(s, twtab, w / TK, k);
}

// ------------------------------------------------------------------ global side

constexpr int kBatch = 8;      // independent global loads a thread keeps in flight

template <typename C> XF_HD C peer_load(const C *p) {
#if defined(__CUDA_ARCH__)
  // system-scope relaxed load: the owner wrote this with ordinary stores before the flag barrier
  C v;
  if (sizeof(C) == 16) {
    double a, b;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p));
    v.x = (typename RealOf<C>::type) a; v.y = (typename RealOf<C>::type) b;
  } else {
    float a, b;
    asm volatile("ld.relaxed.sys.global.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "l"(p));
    v.x = (typename RealOf<C>::type) a; v.y = (typename RealOf<C>::type) b;
  }
  return v;
#else
  return *p;
#endif
}

// backward, load: lines [jl][k0 + k][x], x contiguous: lanes run along x
template <int LGN, int TK, typename C>
XF_HD void phase_load_lines(C *s, const C *__restrict__ in, int NZ, int jl, int k0, int tid, int nthr) {
  constexpr int N = 1 << LGN, tot = TK * N;
  const C *src = in + ((long long) jl * NZ + k0) * N;       // k0 may be negative (first tile of a line): guarded below
  for (int e0 = tid; e0 < tot; e0 += kBatch * nthr) {
    C v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int e = e0 + u * nthr;
      v[u] = mk<C>(0, 0);
      if (e < tot && k0 + (e >> LGN) < NZ && k0 + (e >> LGN) >= 0) v[u] = src[e];
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int e = e0 + u * nthr;
      if (e < tot) s[sidx<TK>(e & (N - 1), e >> LGN)] = v[u];
    }
  }
}

// backward, store: position p holds x = digit_rev(p); it goes to rank x >> lg_nxb as [xl][ky = y0 + jl][kz]:
// lanes run along kz (runs of TK complex values)
template <int LGN, int TK, typename C>
XF_HD void phase_store_owners(const C *s, const PeerPtrs &out, int lg_nxb, int y0, int NY, int NZ, int jl, int k0,
                              int tid, int nthr) {
  constexpr int N = 1 << LGN, tot = TK * N;
  const int k = tid & (TK - 1);
  if (k0 + k >= NZ || k0 + k < 0) return;
  const size_t col = (size_t) (y0 + jl) * NZ + (k0 + k);
  const size_t xstride = (size_t) NY * NZ;
  for (int e = tid; e < tot; e += nthr) {
    const int p = e / TK;
    const int x = digit_rev<LGN>(p);
    const int r = x >> lg_nxb, xl = x & ((1 << lg_nxb) - 1);
    ((C *) out.p[r])[(size_t) xl * xstride + col] = s[sidx<TK>(p, k)];
  }
}

// forward, load: element x of line (jl, k0 + k) lives on rank x >> lg_nxb at [xl][ky = y0 + jl][kz]; it is written to
// the position whose digit reversal it is, so that the decimation-in-time passes deliver natural order
template <int LGN, int TK, typename C>
XF_HD void phase_load_owners(C *s, const PeerPtrs &in, int lg_nxb, int y0, int NY, int NZ, int jl, int k0, int tid,
                             int nthr) {
  constexpr int N = 1 << LGN, tot = TK * N;
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
        const int r = x >> lg_nxb, xl = x & ((1 << lg_nxb) - 1);
        v[u] = peer_load((const C *) in.p[r] + (size_t) xl * xstride + col);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; u++) {
      const int e = e0 + u * nthr;
      if (e < tot) s[sidx<TK>(digit_rev_inv<LGN>(e / TK), k)] = v[u];
    }
  }
}

// forward, store: the lines of the local transposed k-space, x contiguous
template <int LGN, int TK, typename C>
XF_HD void phase_store_lines(const C *s, C *__restrict__ out, int NZ, int jl, int k0, int tid, int nthr) {
  constexpr int N = 1 << LGN, tot = TK * N;
  C *dst = out + ((long long) jl * NZ + k0) * N;
  for (int e = tid; e < tot; e += nthr) {
    const int k = e >> LGN, x = e & (N - 1);
    if (k0 + k < NZ && k0 + k >= 0) dst[e] = s[sidx<TK>(x, k)];
  }
}

// ------------------------------------------------------------------ kernels

#if defined(__CUDACC__)

// backward: in = local [nyl][NZ][N] (transposed k-space); out.p[r] = rank r's landing buffer [nxb][NY][NZ]
template <typename C, int LGN, int TK>
__global__ void __launch_bounds__(kThreads, 2)
k_xfft_bwd_p2p(const C *__restrict__ in, const __grid_constant__ PeerPtrs out, const C *__restrict__ twtab, int lg_nxb,
               int y0, int NY, int NZ, int nyl) {
  extern __shared__ __align__(16) unsigned char xf_smem[];
  C *s = reinterpret_cast<C *>(xf_smem);
  constexpr int A = 128 / (int) sizeof(C);
  const int ktiles = tiles_per_line(NZ, TK, A), ntiles = nyl * ktiles;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int jl = t / ktiles, k0 = tile_k0(y0 + jl, NZ, t - jl * ktiles, TK, A);
    if (k0 >= NZ || k0 + TK <= 0) continue;           // empty tile (the offset moved it off the line)
    phase_load_lines<LGN, TK, C>(s, in, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
    phase_pass<LGN, 0, +1, false, TK, C>(s, twtab, threadIdx.x, kThreads);
    __syncthreads();
    if constexpr (plan_npass(LGN) > 1) { phase_pass<LGN, 1, +1, false, TK, C>(s, twtab, threadIdx.x, kThreads); __syncthreads(); }
    if constexpr (plan_npass(LGN) > 2) { phase_pass<LGN, 2, +1, false, TK, C>(s, twtab, threadIdx.x, kThreads); __syncthreads(); }
    phase_store_owners<LGN, TK, C>(s, out, lg_nxb, y0, NY, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
  }
}

// forward: in.p[r] = rank r's [nxb][NY][NZ] (output of its local 2-D r2c); out = local [nyl][NZ][N]
template <typename C, int LGN, int TK>
__global__ void __launch_bounds__(kThreads, 2)
k_xfft_fwd_p2p(const __grid_constant__ PeerPtrs in, C *__restrict__ out, const C *__restrict__ twtab, int lg_nxb, int y0,
               int NY, int NZ, int nyl) {
  extern __shared__ __align__(16) unsigned char xf_smem[];
  C *s = reinterpret_cast<C *>(xf_smem);
  constexpr int A = 128 / (int) sizeof(C);
  const int ktiles = tiles_per_line(NZ, TK, A), ntiles = nyl * ktiles;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int jl = t / ktiles, k0 = tile_k0(y0 + jl, NZ, t - jl * ktiles, TK, A);
    if (k0 >= NZ || k0 + TK <= 0) continue;           // empty tile (the offset moved it off the line)
    phase_load_owners<LGN, TK, C>(s, in, lg_nxb, y0, NY, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
    // the transposed graph runs the passes in reverse order
    if constexpr (plan_npass(LGN) > 2) { phase_pass<LGN, 2, -1, true, TK, C>(s, twtab, threadIdx.x, kThreads); __syncthreads(); }
    if constexpr (plan_npass(LGN) > 1) { phase_pass<LGN, 1, -1, true, TK, C>(s, twtab, threadIdx.x, kThreads); __syncthreads(); }
    phase_pass<LGN, 0, -1, true, TK, C>(s, twtab, threadIdx.x, kThreads);
    __syncthreads();
    phase_store_lines<LGN, TK, C>(s, out, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
  }
}

#endif  // __CUDACC__

}  // namespace xf
}  // namespace mgp
