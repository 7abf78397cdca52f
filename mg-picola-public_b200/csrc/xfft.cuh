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
// Transform: power-of-two N, in-place in shared memory, radix-16/8/4/2 passes held in registers.  Backward runs
// decimation in frequency (natural in, digit-reversed out: the scatter to the owners undoes the permutation for
// free); forward runs the transposed flow graph (decimation in time: digit-reversed in, natural out: the gather from
// the owners applies the permutation for free).  Unnormalised in both directions, like FFTW / cuFFT.
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

constexpr int kMaxPass = 4;

struct Plan {
  int N, npass;
  int R[kMaxPass], lgR[kMaxPass], lgM[kMaxPass];   // decimation-in-frequency order: pass i works on sub-transforms
                                                   // of length L_i = R_i << lgM_i (L_0 = N, L_{i+1} = M_i)
  int twoff[kMaxPass];                             // exp(+2 pi i t / L_i), t < L_i, starts here in the twiddle table
  int twtotal;
};

// false: N is not a power of two in [16, 4096]
inline bool make_plan(int N, Plan &pl) {
  if (N < 16 || N > 4096 || (N & (N - 1))) return false;
  pl.N = N; pl.npass = 0; pl.twtotal = 0;
  int rem = N;
  while (rem > 1) {
    int R;
    if (rem == 32) R = 8;                 // 32 = 8 * 4 rather than 16 * 2
    else if (rem >= 16) R = 16;
    else R = rem;                         // 8, 4 or 2
    int lgR = 0; while ((1 << lgR) < R) lgR++;
    const int M = rem / R;
    int lgM = 0; while ((1 << lgM) < M) lgM++;
    pl.R[pl.npass] = R; pl.lgR[pl.npass] = lgR; pl.lgM[pl.npass] = lgM;
    pl.twoff[pl.npass] = pl.twtotal; pl.twtotal += rem;
    pl.npass++;
    rem = M;
  }
  return true;
}

// position p of the decimation-in-frequency output holds frequency digit_rev(p)
XF_HD int digit_rev(const Plan &pl, int p) {
  int f = 0, sh = 0;
  for (int i = 0; i < pl.npass; i++) {
    f += ((p >> pl.lgM[i]) & (pl.R[i] - 1)) << sh;
    sh += pl.lgR[i];
  }
  return f;
}
XF_HD int digit_rev_inv(const Plan &pl, int x) {
  int p = 0;
  for (int i = 0; i < pl.npass; i++) {
    p += (x & (pl.R[i] - 1)) << pl.lgM[i];
    x >>= pl.lgR[i];
  }
  return p;
}

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

// One radix-R butterfly (work item w of pass i) on the tile.  DIT = false: DFT_R then twiddle (decimation in
// frequency); DIT = true: twiddle then DFT_R (the transposed graph).  tw = table of this pass, exp(+2 pi i t / L).
template <int R, int SIGN, bool DIT, int TK, typename C>
XF_HD void butterfly(C *s, const C *tw, int lgM, int w) {
  const int k = w & (TK - 1), bb = w / TK;
  const int M = 1 << lgM;
  const int blk = bb >> lgM, b = bb & (M - 1);
  const int base = blk * (R << lgM) + b;
  C v[R];
#pragma unroll
  for (int j = 0; j < R; j++) v[j] = s[sidx<TK>(base + (j << lgM), k)];
  if (DIT && lgM > 0) {
#pragma unroll
    for (int j = 1; j < R; j++) {
      C t = tw_load(tw + b * j);
      if (SIGN < 0) t.y = -t.y;
      v[j] = cmul(v[j], t);
    }
  }
  Dft<R, SIGN, C>::run(v);
  if (!DIT && lgM > 0) {
#pragma unroll
    for (int q = 1; q < R; q++) {
      C t = tw_load(tw + b * q);
      if (SIGN < 0) t.y = -t.y;
      v[q] = cmul(v[q], t);
    }
  }
#pragma unroll
  for (int q = 0; q < R; q++) s[sidx<TK>(base + (q << lgM), k)] = v[q];
}

// all work items of pass i that belong to thread tid of nthr
template <int SIGN, bool DIT, int TK, typename C>
XF_HD void phase_pass(C *s, const Plan &pl, const C *twtab, int i, int tid, int nthr) {
  const int R = pl.R[i], lgM = pl.lgM[i];
  const int items = TK * (pl.N >> pl.lgR[i]);
  const C *tw = twtab + pl.twoff[i];
  for (int w = tid; w < items; w += nthr) {
    switch (R) {
      case 16: butterfly<16, SIGN, DIT, TK, C>(s, tw, lgM, w); break;
      case 8: butterfly<8, SIGN, DIT, TK, C>(s, tw, lgM, w); break;
      case 4: butterfly<4, SIGN, DIT, TK, C>(s, tw, lgM, w); break;
      default: butterfly<2, SIGN, DIT, TK, C>(s, tw, lgM, w); break;
    }
  }
}

// ------------------------------------------------------------------ global side, backward (local lines -> owners of x)

// lines [jl][k0 + k][x], x contiguous: lanes run along x
template <int TK, typename C>
XF_HD void phase_load_lines(C *s, const C *__restrict__ in, int N, int NZ, int jl, int k0, int tid, int nthr) {
  const int tot = TK * N;
  for (int e = tid; e < tot; e += nthr) {
    const int k = e / N, x = e - k * N;
    C v = mk<C>(0, 0);
    if (k0 + k < NZ) v = in[((size_t) jl * NZ + (k0 + k)) * N + x];
    s[sidx<TK>(x, k)] = v;
  }
}

// position p holds x = digit_rev(p); it goes to rank x / nxb as [xl][ky = y0 + jl][kz]: lanes run along kz
template <int TK, typename C>
XF_HD void phase_store_owners(const C *s, const PeerPtrs &out, const Plan &pl, int nxb, int y0, int NY, int NZ, int jl,
                              int k0, int tid, int nthr) {
  const int N = pl.N;
  const int tot = TK * N;
  for (int e = tid; e < tot; e += nthr) {
    const int k = e & (TK - 1), p = e / TK;
    if (k0 + k >= NZ) continue;
    const int x = digit_rev(pl, p);
    const int r = x / nxb, xl = x - r * nxb;
    ((C *) out.p[r])[((size_t) xl * NY + (y0 + jl)) * NZ + (k0 + k)] = s[sidx<TK>(p, k)];
  }
}

// ------------------------------------------------------------------ global side, forward (owners of x -> local lines)

template <typename C> XF_HD C peer_load(const C *p) {
#if defined(__CUDA_ARCH__)
  // system-scope relaxed load: the owner wrote this with ordinary stores before the flag barrier
  C v;
  if (sizeof(C) == 16) {
    double a, b;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p) : "memory");
    v.x = (typename RealOf<C>::type) a; v.y = (typename RealOf<C>::type) b;
  } else {
    float a, b;
    asm volatile("ld.relaxed.sys.global.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "l"(p) : "memory");
    v.x = (typename RealOf<C>::type) a; v.y = (typename RealOf<C>::type) b;
  }
  return v;
#else
  return *p;
#endif
}

// element x of line (jl, k0 + k) lives on rank x / nxb at [xl][ky = y0 + jl][kz]; it is written to the position
// whose digit reversal it is, so that the decimation-in-time passes deliver natural order
template <int TK, typename C>
XF_HD void phase_load_owners(C *s, const PeerPtrs &in, const Plan &pl, int nxb, int y0, int NY, int NZ, int jl, int k0,
                             int tid, int nthr) {
  const int N = pl.N;
  const int tot = TK * N;
  for (int e = tid; e < tot; e += nthr) {
    const int k = e & (TK - 1), x = e / TK;
    C v = mk<C>(0, 0);
    if (k0 + k < NZ) {
      const int r = x / nxb, xl = x - r * nxb;
      v = peer_load((const C *) in.p[r] + ((size_t) xl * NY + (y0 + jl)) * NZ + (k0 + k));
    }
    s[sidx<TK>(digit_rev_inv(pl, x), k)] = v;
  }
}

template <int TK, typename C>
XF_HD void phase_store_lines(const C *s, C *__restrict__ out, int N, int NZ, int jl, int k0, int tid, int nthr) {
  const int tot = TK * N;
  for (int e = tid; e < tot; e += nthr) {
    const int k = e / N, x = e - k * N;
    if (k0 + k < NZ) out[((size_t) jl * NZ + (k0 + k)) * N + x] = s[sidx<TK>(x, k)];
  }
}

// ------------------------------------------------------------------ kernels

#if defined(__CUDACC__)

constexpr int kThreads = 256;

// backward: in = local [nyl][NZ][N] (transposed k-space); out.p[r] = rank r's landing buffer [nxb][N][NZ]
template <typename C, int TK>
__global__ void __launch_bounds__(kThreads, 2)
k_xfft_bwd_p2p(const C *__restrict__ in, const __grid_constant__ PeerPtrs out, const __grid_constant__ Plan pl, const C *__restrict__ twtab, int nxb, int y0, int NY,
               int NZ, int nyl) {
  extern __shared__ __align__(16) unsigned char xf_smem[];
  C *s = reinterpret_cast<C *>(xf_smem);
  const int N = pl.N;
  const int ktiles = (NZ + TK - 1) / TK, ntiles = nyl * ktiles;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int jl = t / ktiles, k0 = (t - jl * ktiles) * TK;
    phase_load_lines<TK, C>(s, in, N, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
    for (int i = 0; i < pl.npass; i++) {
      phase_pass<+1, false, TK, C>(s, pl, twtab, i, threadIdx.x, kThreads);
      __syncthreads();
    }
    phase_store_owners<TK, C>(s, out, pl, nxb, y0, NY, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
  }
}

// forward: in.p[r] = rank r's [nxb][N][NZ] (output of its local 2-D r2c); out = local [nyl][NZ][N]
template <typename C, int TK>
__global__ void __launch_bounds__(kThreads, 2)
k_xfft_fwd_p2p(const __grid_constant__ PeerPtrs in, C *__restrict__ out, const __grid_constant__ Plan pl, const C *__restrict__ twtab, int nxb, int y0, int NY,
               int NZ, int nyl) {
  extern __shared__ __align__(16) unsigned char xf_smem[];
  C *s = reinterpret_cast<C *>(xf_smem);
  const int N = pl.N;
  const int ktiles = (NZ + TK - 1) / TK, ntiles = nyl * ktiles;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int jl = t / ktiles, k0 = (t - jl * ktiles) * TK;
    phase_load_owners<TK, C>(s, in, pl, nxb, y0, NY, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
    for (int i = pl.npass - 1; i >= 0; i--) {
      phase_pass<-1, true, TK, C>(s, pl, twtab, i, threadIdx.x, kThreads);
      __syncthreads();
    }
    phase_store_lines<TK, C>(s, out, N, NZ, jl, k0, threadIdx.x, kThreads);
    __syncthreads();
  }
}

#endif  // __CUDACC__

}  // namespace xf
}  // namespace mgp
