// k-space layout helpers shared by the pointwise k-space kernels.
#pragma once
#include "common.cuh"

namespace mgp {

// k-space layout of this rank: single-rank 3-D plans: [kx][ky][kz]; slab transforms (P > 1): transposed [ky_local][kz][kx]
struct KL {
  int N, NZ, transposed, j0, nyl;
  size_t total;
};

static inline KL layout_of(const Ctx &c) {
  KL L;
  L.N = c.N; L.NZ = c.NZ; L.transposed = c.slab; L.j0 = c.y0; L.nyl = c.ny_loc;
  L.total = c.slab ? (size_t) c.ny_loc * c.NZ * c.N : (size_t) c.N * c.N * c.NZ;
  return L;
}

__device__ __forceinline__ void kl_decode(const KL &L, size_t e, int &i, int &j, int &k) {
  if (!L.transposed) {
    k = (int) (e % (size_t) L.NZ);
    const size_t t = e / (size_t) L.NZ;
    j = (int) (t % (size_t) L.N); i = (int) (t / (size_t) L.N);
  } else {
    i = (int) (e % (size_t) L.N);
    const size_t t = e / (size_t) L.N;
    k = (int) (t % (size_t) L.NZ); j = L.j0 + (int) (t / (size_t) L.NZ);
  }
}

template <typename T> struct Cpx;
template <> struct Cpx<float> { typedef float2 type; };
template <> struct Cpx<double> { typedef double2 type; };

#define KLOOP(e, L) \
  for (size_t e = blockIdx.x * (size_t) blockDim.x + threadIdx.x; e < (L).total; e += (size_t) gridDim.x * blockDim.x)


}  // namespace mgp
