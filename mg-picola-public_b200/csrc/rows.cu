// The MGP_DEPOSIT_ROWS strategy: per-step bins + tile deposit + tile gather (PtoMesh auxPM.c:292-343, MtoParticles
// auxPM.c:574-634).  Double-precision grids (the reference default); float grids keep the ATOMIC strategy.
//
// Every step the particles are BINNED by (x-plane, y-row, z-chunk of `zc` cells) of the mesh cell they sit in -- an index
// list, the particle records themselves are not moved (36 bytes of traffic per particle against ~150 for a sort):
//     k_bin_rank      bin of every particle + its rank inside the bin (one warp-aggregated atomic per warp and bin)
//     cub scan        bin offsets
//     k_bin_invert    perm[bin_start[bin] + rank] = particle index
//
// Deposit (k_deposit_tiles): a CTA takes a tile of R source rows x one z-chunk of one plane.  The eight cells of a
// particle lie in target rows (sx + {0,1}, sy + {0,1}) and z, z + 1: a shared-memory tile [2][R + 1][zc + 2] of doubles
// holds them all.  One warp owns one source row at a time, even rows first, odd rows after a CTA barrier, so no two warps
// ever touch the same target row; lanes of a batch that hit the same z take turns (match.any hands out the turn numbers).
// Hence plain read-modify-writes only -- no atomics in shared memory (a CAS loop on this part, slower than the global
// reductions) and one visit per particle.  The finished tile leaves through the TMA engine as bulk REDUCTIONS, one
// cp.reduce.async.bulk.add.f64 per tile row (1 KB runs); per particle 2.1 values reach the L2 reduction units instead of
// the 8 of one red.global per cell, which is what bounds the ATOMIC strategy (~3.5e11 f64 adds / s chip-wide).
//
// Gather (k_gather_tiles): same tiling.  The force values a tile's particles can touch are [R + 1][zc + 2] windows of
// planes sx and sx + 1 of the three force grids; bulk copies (cp.async.bulk global -> shared, completion on an mbarrier)
// fetch them while the threads already load their particles, and the 24 reads per particle hit shared memory.
#include "common.cuh"
#include "cic.cuh"
#include "reduce.cuh"
#include "tma.cuh"

#include <cub/device/device_scan.cuh>

namespace mgp {

// ------------------------------------------------------------------ binning

__global__ void __launch_bounds__(256)
k_bin_rank(size_t n, const float4 *__restrict__ pA, double scale, int N, int x0, int nx, int zc, int nc,
           uint32_t *__restrict__ count, uint32_t *__restrict__ bin_of, uint32_t *__restrict__ rank_of) {
  const unsigned lane = threadIdx.x & 31;
  const size_t nround = (n + 31) / 32 * 32;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < nround; i += (size_t) gridDim.x * blockDim.x) {
    unsigned b = 0xffffffffu - lane;                 // idle lanes: distinct sentinels
    if (i < n) {
      const float4 a = pA[i];
      double fl;
      unsigned ix = floor_u32((double) a.x * scale, fl);
      unsigned iy = floor_u32((double) a.y * scale, fl);
      unsigned iz = floor_u32((double) a.z * scale, fl);
      if (iy >= (unsigned) N) iy = 0;
      if (iz >= (unsigned) N) iz = 0;
      if (ix >= (unsigned) N) ix = N - 1;            // cannot happen for float positions in [0, Box)
      int lx = (int) ix - x0;
      if (lx < 0) lx = 0;                            // foreign particles never reach the deposit (MoveParticles ran)
      if (lx >= nx) lx = nx - 1;
      b = ((unsigned) lx * (unsigned) N + iy) * (unsigned) nc + iz / (unsigned) zc;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if ((int) lane == leader && i < n) base = atomicAdd(&count[b], (unsigned) __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < n) {
      bin_of[i] = b;
      rank_of[i] = base + (unsigned) __popc(peers & ((1u << lane) - 1u));
    }
  }
}

__global__ void __launch_bounds__(256)
k_bin_invert(size_t n, const uint32_t *__restrict__ bin_of, const uint32_t *__restrict__ rank_of,
             const uint32_t *__restrict__ start, uint32_t *__restrict__ perm) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    perm[(size_t) start[bin_of[i]] + rank_of[i]] = (uint32_t) i;
}

void rows_alloc(Ctx &c) {
  // z-chunk: the whole row up to 128 cells, else the largest even divisor of Nmesh that is <= 128 (powers of two: 128)
  int zc = c.N;
  if (c.N > 128)
    for (int d = 128; d >= 2; d--)
      if (c.N % d == 0 && d % 2 == 0) { zc = d; break; }
  if (const char *e = getenv("MGP_BIN_ZC")) {       // test knob: several z-chunks on a small mesh
    const int v = atoi(e);
    if (v >= 2 && v % 2 == 0 && c.N % v == 0) zc = v;
  }
  c.bin_zc = zc;
  c.bin_nc = c.N / zc;
  c.nbins = (size_t) c.nx * c.N * c.bin_nc;
  CK(cudaMalloc(&c.bin_perm, c.cap * sizeof(uint32_t)));
  CK(cudaMalloc(&c.bin_start, (c.nbins + 2) * sizeof(uint32_t)));
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, c.bin_start, c.bin_start, (int64_t) (c.nbins + 2), c.stream);
  c.bin_scan_bytes = scan_bytes;
  CK(cudaMalloc(&c.bin_scan_temp, scan_bytes ? scan_bytes : 16));
  c.bins_valid = false;
}

void rows_free(Ctx &c) {
  cudaFree(c.bin_perm); cudaFree(c.bin_start); cudaFree(c.bin_scan_temp);
  c.bin_perm = nullptr; c.bin_start = nullptr; c.bin_scan_temp = nullptr;
}

// index list of the particles grouped by bin, for the current positions and storage order
void rows_bin(Ctx &c) {
  if (c.bins_valid) return;
  PhaseTimer t(c, PH_SORT);
  const size_t n = c.np;
  CK(cudaMemsetAsync(c.bin_start, 0, (c.nbins + 2) * sizeof(uint32_t), c.stream));
  if (n) {
    const double scale = (double) c.N / c.cfg.box;
    k_bin_rank<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.pA, scale, c.N, c.x0, c.nx, c.bin_zc, c.bin_nc, c.bin_start,
                                                     c.key[0], c.perm[0]);
    size_t tb = c.bin_scan_bytes;
    CK(cub::DeviceScan::ExclusiveSum(c.bin_scan_temp, tb, c.bin_start, c.bin_start, (int64_t) (c.nbins + 2), c.stream));
    k_bin_invert<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.key[0], c.perm[0], c.bin_start, c.bin_perm);
    c.launches += 2 + 2;   // + CUB's scan
  }
  c.bins_valid = true;
}

// ------------------------------------------------------------------ deposit

__global__ void k_fill_rows(double *g, size_t n, double v) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) g[i] = v;
}

constexpr int kDepWarps = 16;             // warps per CTA; a tile has 2 * kDepWarps source rows
constexpr int kDepRows = 2 * kDepWarps;

__global__ void __launch_bounds__(kDepWarps * 32)
k_deposit_tiles(const float4 *__restrict__ pA, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ bin_start,
                double *__restrict__ grid, int N, int NZ, int nx, int single_rank, double scale, double W, int zc, int nc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *acc = reinterpret_cast<double *>(smem_raw);          // [2][kDepRows + 1][zw]
  const int rz = 2 * NZ, zw = zc + 2;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t plane = (size_t) (kDepRows + 1) * zw;
  const int nyb = (N + kDepRows - 1) / kDepRows;
  const long long ntiles = (long long) nx * nyb * nc;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int c = (int) (t % nc);
    const long long t2 = t / nc;
    const int yb = (int) (t2 % nyb), sx = (int) (t2 / nyb);
    const int sy0 = yb * kDepRows, rows = min(kDepRows, N - sy0), z0 = c * zc;
    // the (at most two) bins this warp will process
    uint32_t rs[2] = {0, 0}, cnt[2] = {0, 0};
#pragma unroll
    for (int p = 0; p < 2; p++) {
      const int ry = 2 * w + p;
      if (ry < rows) {
        const size_t b = ((size_t) sx * N + sy0 + ry) * nc + c;
        rs[p] = bin_start[b]; cnt[p] = bin_start[b + 1] - rs[p];
      }
    }
    if (!__syncthreads_or((int) (cnt[0] | cnt[1]))) continue;          // empty tile: nothing to add
    for (size_t i = threadIdx.x; i < 2 * plane; i += blockDim.x) acc[i] = 0.0;
    __syncthreads();
#pragma unroll 1
    for (int p = 0; p < 2; p++) {
      const int ry = 2 * w + p;
      const uint32_t total = cnt[p];
      if (total) {
        double *a0 = acc + (size_t) ry * zw, *a1 = a0 + zw;              // plane sx:     rows sy, sy + 1
        double *a2 = a0 + plane, *a3 = a2 + zw;                          // plane sx + 1
        // lane l takes particles l * nb + k: the lanes of a batch are spread over the bin, which keeps same-cell
        // collisions inside a batch rare
        const uint32_t nb = (total + 31) >> 5;
        const uint32_t first = (uint32_t) lane * nb;
        const uint32_t mine = first < total ? min(nb, total - first) : 0u;
        const uint32_t base = rs[p] + first;
        uint32_t i_nxt = 0;
        float4 p_cur = make_float4(0.f, 0.f, 0.f, 0.f), p_nxt = p_cur;
        if (0 < mine) p_cur = pA[perm[base]];
        if (1 < mine) i_nxt = perm[base + 1];
        for (uint32_t k = 0; k < nb; k++) {
          if (k + 1 < mine) p_nxt = pA[i_nxt];
          uint32_t i_nn = 0;
          if (k + 2 < mine) i_nn = perm[base + k + 2];
          const bool live = k < mine;
          int z = -1 - lane;               // idle lanes: distinct, never written
          double w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0, w5 = 0, w6 = 0, w7 = 0;
          if (live) {
            double fl;
            const double X = (double) p_cur.x * scale, Y = (double) p_cur.y * scale, Z = (double) p_cur.z * scale;
            floor_u32(X, fl); const double dx = X - fl, tx = 1.0 - dx;
            floor_u32(Y, fl); const double dy = (Y - fl) * W, ty = (1.0 - (Y - fl)) * W;
            unsigned iz = floor_u32(Z, fl); const double dz = Z - fl, tz = 1.0 - dz;
            const double a = tx * ty, b = tx * dy, cc = dx * ty, d = dx * dy;
            w0 = a * tz; w1 = a * dz; w2 = b * tz; w3 = b * dz; w4 = cc * tz; w5 = cc * dz; w6 = d * tz; w7 = d * dz;
            if (iz >= (unsigned) N) iz = 0;
            z = (int) iz - z0;             // 0 .. zc - 1 by construction of the bins
          }
          const unsigned peers = __match_any_sync(0xffffffffu, z);
          const int turn = __popc(peers & ((1u << lane) - 1u));
          const int turns = (int) __reduce_max_sync(0xffffffffu, (unsigned) __popc(peers));
          for (int r = 0; r < turns; r++) {
            const bool on = live && turn == r;
            if (on) { a0[z] += w0; a1[z] += w2; a2[z] += w4; a3[z] += w6; }
            __syncwarp();
            if (on) { a0[z + 1] += w1; a1[z + 1] += w3; a2[z + 1] += w5; a3[z + 1] += w7; }
            __syncwarp();
          }
          p_cur = p_nxt; i_nxt = i_nn;
        }
      }
      if (p == 1) tma::fence_async_shared();
      __syncthreads();                     // even rows done before the odd rows start; all done before the flush
    }
    // flush: one bulk reduction per tile row (two for the last z-chunk, whose z + 1 column wraps to z = 0)
    if ((int) threadIdx.x < 2 * (rows + 1)) {
      const int pl = (int) threadIdx.x / (rows + 1), ry = (int) threadIdx.x - pl * (rows + 1);
      int gx = sx + pl;
      if (single_rank && gx == nx) gx = 0;               // several ranks: plane nx is the ghost plane
      int gy = sy0 + ry; if (gy >= N) gy -= N;
      double *grow = grid + ((size_t) gx * N + gy) * (size_t) rz;
      const double *srow = acc + (size_t) pl * plane + (size_t) ry * zw;
      if (c + 1 < nc) {
        tma::reduce_add_f64(grow + z0, srow, (unsigned) (zw * sizeof(double)));      // slot zc + 1 is zero
      } else {
        tma::reduce_add_f64(grow + z0, srow, (unsigned) (zc * sizeof(double)));
        tma::reduce_add_f64(grow, srow + zc, (unsigned) (2 * sizeof(double)));
      }
      tma::commit();
      tma::wait_read();
    }
    __syncthreads();
  }
  tma::wait_all();
}

bool deposit_rows_supported(const Ctx &c) { return c.bin_perm != nullptr && c.gbytes == 8; }

void deposit_rows(Ctx &c, int grid_id) {
  rows_bin(c);
  PhaseTimer t(c, PH_PTOMESH);
  double *grid = (double *) c.grid[grid_id];
  const double scale = (double) c.N / c.cfg.box;
  const double r = (double) c.N / (double) c.cfg.nsample;
  const double W = r * r * r;
  const size_t sm = (size_t) 2 * (kDepRows + 1) * (c.bin_zc + 2) * sizeof(double);
  static size_t attr_sm = 0;
  if (sm > attr_sm) {
    CK(cudaFuncSetAttribute(k_deposit_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
    attr_sm = sm;
  }
  k_fill_rows<<<grid_for(c.grid_vals, 256), 256, 0, c.stream>>>(grid, c.grid_vals, -1.0);
  c.launches++;
  if (!c.np) return;
  static int occ = 0;             // host-side query: once (the launch sits right behind a host synchronisation)
  static size_t occ_sm = 0;
  if (!occ || occ_sm != sm) {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_deposit_tiles, kDepWarps * 32, sm));
    if (occ < 1) occ = 1;
    occ_sm = sm;
  }
  const long long ntiles = (long long) c.nx * ((c.N + kDepRows - 1) / kDepRows) * c.bin_nc;
  long long g = (long long) kSMs * occ;
  if (g > ntiles) g = ntiles;
  k_deposit_tiles<<<(unsigned) g, kDepWarps * 32, sm, c.stream>>>(c.pA, c.bin_perm, c.bin_start, grid, c.N, c.NZ, c.nx,
                                                                c.P == 1, scale, W, c.bin_zc, c.bin_nc);
  c.launches++;
}

// ------------------------------------------------------------------ gather

constexpr int kGatWarps = 8;              // warps per CTA = source rows per tile
constexpr int kGatRows = kGatWarps;

// shared tile: [component 3][plane 2][kGatRows + 1][zw] doubles, then one mbarrier
__global__ void __launch_bounds__(kGatWarps * 32)
k_gather_tiles(const float4 *__restrict__ pA, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ bin_start,
               const double *__restrict__ fx, const double *__restrict__ fy, const double *__restrict__ fz,
               float *__restrict__ disp, size_t cap, int N, int NZ, int nx, double scale, int zc, int nc,
               double *__restrict__ partial) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int rz = 2 * NZ, zw = zc + 2;
  const size_t plane = (size_t) (kGatRows + 1) * zw;
  double *tile = reinterpret_cast<double *>(smem_raw);
  uint64_t *bar = reinterpret_cast<uint64_t *>(tile + 6 * plane);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tma::mbar_init(bar, 1); tma::fence_mbar_init(); }
  __syncthreads();
  const int nyb = (N + kGatRows - 1) / kGatRows;
  const long long ntiles = (long long) nx * nyb * nc;
  unsigned parity = 0;
  double sx_ = 0, sy_ = 0, sz_ = 0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int c = (int) (t % nc);
    const long long t2 = t / nc;
    const int yb = (int) (t2 % nyb), sx = (int) (t2 / nyb);
    const int sy0 = yb * kGatRows, rows = min(kGatRows, N - sy0), z0 = c * zc;
    uint32_t rs = 0, total = 0;
    if (w < rows) {
      const size_t b = ((size_t) sx * N + sy0 + w) * nc + c;
      rs = bin_start[b]; total = bin_start[b + 1] - rs;
    }
    if (!__syncthreads_or((int) total)) continue;                      // no particles in this tile
    // windows [rows + 1][zc + 2] of planes sx and sx + 1 (plane nx = the ghost plane = the right neighbour's plane 0) of
    // the three force grids; row N wraps to row 0; the last z-chunk takes its z + 1 column from z = 0
    const int nseg = 6 * (rows + 1);
    const bool lastc = c + 1 == nc;
    if (threadIdx.x == 0) tma::mbar_arrive_expect_tx(bar, (unsigned) nseg * (unsigned) (zw * sizeof(double)));
    if ((int) threadIdx.x < nseg) {
      const int a = (int) threadIdx.x / (2 * (rows + 1)), rem = (int) threadIdx.x - a * 2 * (rows + 1);
      const int pl = rem / (rows + 1), ry = rem - pl * (rows + 1);
      int gy = sy0 + ry; if (gy >= N) gy -= N;
      const double *f = a == 0 ? fx : (a == 1 ? fy : fz);
      const double *grow = f + ((size_t) (sx + pl) * N + gy) * (size_t) rz;
      double *dst = tile + (size_t) (a * 2 + pl) * plane + (size_t) ry * zw;
      if (!lastc) {
        tma::load(dst, grow + z0, (unsigned) (zw * sizeof(double)), bar);
      } else {
        tma::load(dst, grow + z0, (unsigned) (zc * sizeof(double)), bar);
        tma::load(dst + zc, grow, (unsigned) (2 * sizeof(double)), bar);
      }
    }
    // this warp's row: lanes stride over its particles, two at a time; the particle loads overlap the tile load
    bool waited = false;
    for (uint32_t j0 = (uint32_t) lane; j0 < total; j0 += 64) {
      const uint32_t j1 = j0 + 32;
      const bool two = j1 < total;
      const uint32_t i0 = perm[rs + j0], i1 = two ? perm[rs + j1] : 0u;
      const float4 q0 = pA[i0];
      float4 q1 = q0;
      if (two) q1 = pA[i1];
      if (!waited) { tma::mbar_wait(bar, parity); waited = true; }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (u == 1 && !two) break;
        const float4 p = u ? q1 : q0;
        double fl;
        const double X = (double) p.x * scale, Y = (double) p.y * scale, Z = (double) p.z * scale;
        floor_u32(X, fl); const double dx = X - fl, tx = 1.0 - dx;
        floor_u32(Y, fl); const double dy = Y - fl, ty = 1.0 - dy;
        unsigned iz = floor_u32(Z, fl); const double dz = Z - fl, tz = 1.0 - dz;
        if (iz >= (unsigned) N) iz = 0;
        const size_t o00 = (size_t) w * zw + (iz - (unsigned) z0), o01 = o00 + zw;      // plane sx: rows sy, sy + 1
        const size_t o10 = plane + o00, o11 = o10 + zw;                               // plane sx + 1
        const double a = tx * ty, b = tx * dy, cc = dx * ty, d = dx * dy;
        const double w0 = a * tz, w1 = a * dz, w2 = b * tz, w3 = b * dz;
        const double w4 = cc * tz, w5 = cc * dz, w6 = d * tz, w7 = d * dz;
        float g[3];
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
          const double *F = tile + (size_t) ax * 2 * plane;
          g[ax] = (float) (F[o00] * w0 + F[o00 + 1] * w1 + F[o01] * w2 + F[o01 + 1] * w3 + F[o10] * w4 + F[o10 + 1] * w5 +
                           F[o11] * w6 + F[o11 + 1] * w7);
        }
        const size_t i = u ? i1 : i0;
        disp[i] = g[0]; disp[cap + i] = g[1]; disp[2 * cap + i] = g[2];
        sx_ += (double) g[0]; sy_ += (double) g[1]; sz_ += (double) g[2];     // sumDxyz += Disp (float values)
      }
    }
    if (!waited) tma::mbar_wait(bar, parity);     // every thread observes the phase before the barrier is re-armed
    parity ^= 1u;
    __syncthreads();                              // everybody is done with the tile before the next load lands
  }
  block_sum3(sx_, sy_, sz_);
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x] = sx_; partial[3 * blockIdx.x + 1] = sy_; partial[3 * blockIdx.x + 2] = sz_;
  }
}

bool gather_rows_supported(const Ctx &c) { return c.bin_perm != nullptr && c.gbytes == 8; }

// returns the number of per-block partial sums (3 doubles each) left in c.d_red
unsigned gather_rows(Ctx &c) {
  rows_bin(c);
  const size_t sm = (size_t) 6 * (kGatRows + 1) * (c.bin_zc + 2) * sizeof(double) + 16;
  static size_t attr_sm = 0;      // (the kernel also has 768 bytes of static shared memory)
  if (sm > attr_sm) {
    CK(cudaFuncSetAttribute(k_gather_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
    attr_sm = sm;
  }
  static int occ = 0;
  static size_t occ_sm = 0;
  if (!occ || occ_sm != sm) {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gather_tiles, kGatWarps * 32, sm));
    if (occ < 1) occ = 1;
    occ_sm = sm;
  }
  const long long ntiles = (long long) c.nx * ((c.N + kGatRows - 1) / kGatRows) * c.bin_nc;
  long long g = (long long) kSMs * occ;
  if (g > ntiles) g = ntiles;
  reduce_alloc(c, (size_t) g * 3 + 16);
  const double scale = (double) c.N / c.cfg.box;
  k_gather_tiles<<<(unsigned) g, kGatWarps * 32, sm, c.stream>>>(c.pA, c.bin_perm, c.bin_start, (const double *) c.grid[1],
                                                               (const double *) c.grid[2], (const double *) c.grid[3], c.disp,
                                                               c.cap, c.N, c.NZ, c.nx, scale, c.bin_zc, c.bin_nc, c.d_red);
  c.launches++;
  return (unsigned) g;
}

}  // namespace mgp
