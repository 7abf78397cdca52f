// Deterministic block / grid reductions shared by the particle and grid kernels.
#pragma once
#include "common.cuh"

namespace mgp {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums three doubles over the block in a fixed order; result valid in thread 0
__device__ __forceinline__ void block_sum3(double &a, double &b, double &cc) {
  __shared__ double sm_[3][32];
  a = warp_sum(a); b = warp_sum(b); cc = warp_sum(cc);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) { sm_[0][w] = a; sm_[1][w] = b; sm_[2][w] = cc; }
  __syncthreads();
  if (w == 0) {
    a = l < nw ? sm_[0][l] : 0.0; b = l < nw ? sm_[1][l] : 0.0; cc = l < nw ? sm_[2][l] : 0.0;
    a = warp_sum(a); b = warp_sum(b); cc = warp_sum(cc);
  }
}

// out[j] = scale * sum_b partial[b*stride + j]   (single block, fixed order)
static __global__ void k_final_reduce(const double *__restrict__ partial, int nblocks, int stride, int nout,
                                      double scale, double *__restrict__ out) {
  for (int j = 0; j < nout; j++) {
    double s = 0, d1 = 0, d2 = 0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[(size_t) b * stride + j];
    block_sum3(s, d1, d2);
    if (threadIdx.x == 0) out[j] = s * scale;
    __syncthreads();
  }
}

// fft.cu: NCCL sum of n doubles in place on the stream (no-op for a single rank)
void allreduce_sum(Ctx &c, double *dbuf, int n);

}  // namespace mgp
