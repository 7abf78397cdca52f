// Cloud-in-cell deposit (PtoMesh, auxPM.c:292-343) and trilinear gather (MtoParticles,
// auxPM.c:574-634) on cell-sorted particles.
//
// CIC arithmetic follows the reference exactly: X = (double)Pos * (Nmesh/Box); I = (unsigned)X;
// D = X - I; T = 1 - D; DY,TY *= W with W = (Nmesh/Nsample)^3; y,z wrap to 0 at Nmesh; the x
// neighbour of the last local plane is the ghost plane Local_nx (single rank: wraps to plane 0).
// The grid starts at -1 (auxPM.c:292-293) so that it holds delta = rho/rho_mean - 1.
//
// Three deposit strategies (mgp_config.deposit_mode):
//   ROWSEG (MGP_DEPOSIT_DETERMINISTIC, default): one warp owns one target z-row.  It streams the four
//     source rows that can reach it (particles are contiguous per row after the sort), reduces runs
//     of equal cells with a segmented warp scan and adds the run totals to a shared-memory row
//     accumulator with plain (conflict-free) read-modify-writes.  Every grid value is written exactly
//     once, coalesced, with the -1 baseline folded in: no atomics, no memset, bitwise reproducible.
//   TILE   (MGP_DEPOSIT_TILE): one CTA owns TY source rows of one plane; shared-memory tile of
//     2 x (TY+1) rows, warp-aggregated shared atomics, one global reduction per touched tile value.
//   ATOMIC (MGP_DEPOSIT_ATOMIC): one thread per particle, warp-aggregated global reductions.
#include "common.cuh"
#include "cic.cuh"
#include "reduce.cuh"

namespace mgp {

// ------------------------------------------------------------------ helpers

template <typename T>
__global__ void k_fill(T *g, size_t n, T v) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) g[i] = v;
}

__device__ __forceinline__ double shfl_up_d(double v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }

// ------------------------------------------------------------------ ROWSEG deposit

// One warp per target row (tx, ty).  acc: N doubles of shared memory per warp.
template <typename T>
__global__ void __launch_bounds__(256)
k_deposit_rowseg(const float4 *__restrict__ pA, const uint32_t *__restrict__ row_start, T *__restrict__ grid,
                 int N, int NZ, int nx, int x0, int single_rank, double scale, double W) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double *acc = smem + (size_t) wib * N;
  const long long ntarget = (long long) (single_rank ? nx : nx + 1) * N;
  const long long wstride = (long long) gridDim.x * wpb;
  for (long long tr = (long long) blockIdx.x * wpb + wib; tr < ntarget; tr += wstride) {
    const int tx = (int) (tr / N), ty = (int) (tr - (long long) tx * N);
    for (int z = lane; z < N; z += 32) acc[z] = 0.0;
    __syncwarp();
#pragma unroll 1
    for (int s = 0; s < 4; s++) {
      const int ddx = s >> 1, ddy = s & 1;
      int sx = tx - ddx;
      if (single_rank) { if (sx < 0) sx += nx; }
      else if (sx < 0 || sx >= nx) continue;
      int sy = ty - ddy; if (sy < 0) sy += N;
      const uint32_t rs = row_start[(size_t) sx * N + sy], re = row_start[(size_t) sx * N + sy + 1];
      for (uint32_t b = rs; b < re; b += 32) {
        const uint32_t i = b + lane;
        int z = -1 - lane;           // distinct sentinels for idle lanes
        double v0 = 0.0, v1 = 0.0;
        if (i < re) {
          const Cic q = cic_of(pA[i], scale, (unsigned) N, W);
          const double wxy = (ddx ? q.dx : q.tx) * (ddy ? q.dy : q.ty);
          v0 = wxy * q.tz; v1 = wxy * q.dz;
          z = (int) q.iz;
        }
        // inclusive segmented scan over runs of equal z (runs are contiguous: particles are cell-sorted)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int zz = __shfl_up_sync(0xffffffffu, z, o);
          const double a0 = shfl_up_d(v0, o), a1 = shfl_up_d(v1, o);
          if (lane >= o && zz == z) { v0 += a0; v1 += a1; }
        }
        const int znext = __shfl_down_sync(0xffffffffu, z, 1);
        const bool tail = (z >= 0) && (lane == 31 || znext != z);
        if (tail) acc[z] += v0;
        __syncwarp();
        if (tail) { const int z1 = (z + 1 == N) ? 0 : z + 1; acc[z1] += v1; }
        __syncwarp();
      }
    }
    T *out = grid + ((size_t) tx * N + ty) * (size_t) (2 * NZ);
    for (int z = lane; z < 2 * NZ; z += 32) out[z] = (z < N) ? (T) (acc[z] - 1.0) : (T) (-1.0);
    __syncwarp();
  }
}

// ------------------------------------------------------------------ ATOMIC deposit

template <typename T>
__global__ void __launch_bounds__(256)
k_deposit_atomic(size_t n, const float4 *__restrict__ pA, T *__restrict__ grid, int N, int NZ, int nx, int x0,
                 int single_rank, double scale, double W, int aggregate) {
  const int lane = threadIdx.x & 31;
  const size_t nround = (n + 31) / 32 * 32;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < nround; i += (size_t) gridDim.x * blockDim.x) {
    const bool live = i < n;
    double w[8];
    long long cell = -1 - lane;
    unsigned lx = 0, iy = 0, iz = 0;
    if (live) {
      const Cic q = cic_of(pA[i], scale, (unsigned) N, W);
      lx = q.ix - (unsigned) x0; iy = q.iy; iz = q.iz;
      cell = ((long long) lx * N + iy) * N + iz;
      const double a = q.tx * q.ty, b = q.tx * q.dy, cc = q.dx * q.ty, d = q.dx * q.dy;
      w[0] = a * q.tz; w[1] = a * q.dz; w[2] = b * q.tz; w[3] = b * q.dz;
      w[4] = cc * q.tz; w[5] = cc * q.dz; w[6] = d * q.tz; w[7] = d * q.dz;
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) w[k] = 0.0;
    }
    // warp aggregation: maximal runs of ADJACENT lanes with the same cell are summed by a segmented
    // scan and written once by the last lane of the run (equal cells need not be contiguous overall)
    bool writer = live;
    if (aggregate) {
      const long long prev = __shfl_up_sync(0xffffffffu, cell, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != cell);
      if (heads != 0xffffffffu) {
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));   // first lane of my run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const double a = shfl_up_d(w[k], o);
            if (lane - o >= start) w[k] += a;
          }
        }
        writer = live && (lane == 31 || ((heads >> (lane + 1)) & 1u));
      }
    }
    if (writer) {
      unsigned lx1 = lx + 1;
      if (single_rank && lx1 == (unsigned) nx) lx1 = 0;
      const unsigned iy1 = (iy + 1 == (unsigned) N) ? 0 : iy + 1, iz1 = (iz + 1 == (unsigned) N) ? 0 : iz + 1;
      const size_t rz = (size_t) 2 * NZ;
      T *r00 = grid + ((size_t) lx * N + iy) * rz, *r01 = grid + ((size_t) lx * N + iy1) * rz;
      T *r10 = grid + ((size_t) lx1 * N + iy) * rz, *r11 = grid + ((size_t) lx1 * N + iy1) * rz;
      atomicAdd(r00 + iz, (T) w[0]); atomicAdd(r00 + iz1, (T) w[1]);
      atomicAdd(r01 + iz, (T) w[2]); atomicAdd(r01 + iz1, (T) w[3]);
      atomicAdd(r10 + iz, (T) w[4]); atomicAdd(r10 + iz1, (T) w[5]);
      atomicAdd(r11 + iz, (T) w[6]); atomicAdd(r11 + iz1, (T) w[7]);
    }
  }
}

// ------------------------------------------------------------------ TILE deposit

// CTA per (plane sx, TY consecutive source rows).  Shared tile: [2 planes][TY+1 rows][N] of T.
template <typename T>
__global__ void __launch_bounds__(256)
k_deposit_tile(const float4 *__restrict__ pA, const uint32_t *__restrict__ row_start, T *__restrict__ grid, int N,
               int NZ, int nx, int x0, int single_rank, double scale, double W, int TY) {
  extern __shared__ unsigned char smem_raw[];
  T *tile = reinterpret_cast<T *>(smem_raw);
  const int lane = threadIdx.x & 31;
  const int nyb = (N + TY - 1) / TY;
  const long long ntiles = (long long) nx * nyb;
  const size_t tile_vals = (size_t) 2 * (TY + 1) * N;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int sx = (int) (t / nyb), yb = (int) (t - (long long) sx * nyb);
    const int sy0 = yb * TY, rows = min(TY, N - sy0);
    const uint32_t rs = row_start[(size_t) sx * N + sy0], re = row_start[(size_t) sx * N + sy0 + rows];
    if (rs == re) continue;
    for (size_t k = threadIdx.x; k < tile_vals; k += blockDim.x) tile[k] = (T) 0;
    __syncthreads();
    const uint32_t nround = (re - rs + 31) / 32 * 32;
    for (uint32_t j = threadIdx.x; j < nround; j += blockDim.x) {
      const uint32_t i = rs + j;
      const bool live = i < re;
      double w[8];
      int cell = -1 - lane, ry = 0, iz = 0;
      if (live) {
        const Cic q = cic_of(pA[i], scale, (unsigned) N, W);
        ry = (int) q.iy - sy0; iz = (int) q.iz;
        cell = ry * N + iz;
        const double a = q.tx * q.ty, b = q.tx * q.dy, cc = q.dx * q.ty, d = q.dx * q.dy;
        w[0] = a * q.tz; w[1] = a * q.dz; w[2] = b * q.tz; w[3] = b * q.dz;
        w[4] = cc * q.tz; w[5] = cc * q.dz; w[6] = d * q.tz; w[7] = d * q.dz;
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = 0.0;
      }
      bool writer = live;
      const int prev = __shfl_up_sync(0xffffffffu, cell, 1);
      const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || prev != cell);
      if (heads != 0xffffffffu) {
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));   // first lane of my run
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const double a = shfl_up_d(w[k], o);
            if (lane - o >= start) w[k] += a;
          }
        }
        writer = live && (lane == 31 || ((heads >> (lane + 1)) & 1u));
      }
      if (writer) {
        const int iz1 = (iz + 1 == N) ? 0 : iz + 1;
        T *p0 = tile + (size_t) ry * N, *p1 = p0 + N;                      // plane sx: rows ry, ry+1
        T *q0 = tile + (size_t) (TY + 1 + ry) * N, *q1 = q0 + N;           // plane sx+1
        atomicAdd(p0 + iz, (T) w[0]); atomicAdd(p0 + iz1, (T) w[1]);
        atomicAdd(p1 + iz, (T) w[2]); atomicAdd(p1 + iz1, (T) w[3]);
        atomicAdd(q0 + iz, (T) w[4]); atomicAdd(q0 + iz1, (T) w[5]);
        atomicAdd(q1 + iz, (T) w[6]); atomicAdd(q1 + iz1, (T) w[7]);
      }
    }
    __syncthreads();
    // flush: one global reduction per non-zero tile value
    for (size_t k = threadIdx.x; k < (size_t) 2 * (rows + 1) * N; k += blockDim.x) {
      const int pl = (int) (k / ((size_t) (rows + 1) * N));
      const size_t rem = k - (size_t) pl * (rows + 1) * N;
      const int ry = (int) (rem / N), z = (int) (rem - (size_t) ry * N);
      const T v = tile[((size_t) pl * (TY + 1) + ry) * N + z];
      if (v != (T) 0) {
        int gx = sx + pl;
        if (single_rank && gx == nx) gx = 0;
        int gy = sy0 + ry; if (gy >= N) gy -= N;
        atomicAdd(grid + ((size_t) gx * N + gy) * (size_t) (2 * NZ) + z, v);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ dispatch

template <typename T>
static void deposit_t(Ctx &c, int gid) {
  T *grid = (T *) c.grid[gid];
  const double scale = (double) c.N / c.cfg.box;
  const double r = (double) c.N / (double) c.cfg.nsample;
  const double W = r * r * r;     // pow((double)Nmesh/(double)Nsample, 3)
  const int single = c.P == 1;
  int mode = c.cfg.deposit_mode;
  if (mode == MGP_DEPOSIT_ROWS) mode = MGP_DEPOSIT_ATOMIC;   // only reached when the row tile does not fit (deposit_density)
  if (!c.sorted) mode = MGP_DEPOSIT_ATOMIC;      // TILE / DETERMINISTIC need row_start of the current order
  if (mode == MGP_DEPOSIT_DETERMINISTIC && !c.exact_cell_order) mode = MGP_DEPOSIT_ATOMIC;
  const size_t n = c.np;
  if (mode == MGP_DEPOSIT_DETERMINISTIC) {
    int wpb = 8;
    while (wpb > 1 && (size_t) wpb * c.N * sizeof(double) > 200 * 1024) wpb >>= 1;
    const size_t sm = (size_t) wpb * c.N * sizeof(double);
    REQUIRE(sm <= 227 * 1024, MGP_ERR_INVALID, "deposit: Nmesh too large for the row accumulator");
    CK(cudaFuncSetAttribute(k_deposit_rowseg<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_deposit_rowseg<T>, wpb * 32, sm));
    if (occ < 1) occ = 1;
    const long long ntarget = (long long) (single ? c.nx : c.nx + 1) * c.N;
    long long g = (ntarget + wpb - 1) / wpb;
    if (g > (long long) kSMs * occ) g = (long long) kSMs * occ;
    k_deposit_rowseg<T><<<(unsigned) g, wpb * 32, sm, c.stream>>>(c.pA, c.row_start, grid, c.N, c.NZ, c.nx, c.x0,
                                                                  single, scale, W);
    c.launches++;
    return;
  }
  k_fill<T><<<grid_for(c.grid_vals, 256), 256, 0, c.stream>>>(grid, c.grid_vals, (T) -1.0);
  c.launches++;
  if (n == 0) return;
  if (mode == MGP_DEPOSIT_TILE) {
    int TY = 4;
    while (TY > 1 && (size_t) 2 * (TY + 1) * c.N * sizeof(T) > 96 * 1024) TY--;
    const size_t sm = (size_t) 2 * (TY + 1) * c.N * sizeof(T);
    REQUIRE(sm <= 227 * 1024, MGP_ERR_INVALID, "deposit: Nmesh too large for the shared tile");
    CK(cudaFuncSetAttribute(k_deposit_tile<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm));
    int occ = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_deposit_tile<T>, 256, sm));
    if (occ < 1) occ = 1;
    long long ntiles = (long long) c.nx * ((c.N + TY - 1) / TY);
    long long g = ntiles < (long long) kSMs * occ * 4 ? ntiles : (long long) kSMs * occ * 4;
    k_deposit_tile<T><<<(unsigned) g, 256, sm, c.stream>>>(c.pA, c.row_start, grid, c.N, c.NZ, c.nx, c.x0, single,
                                                           scale, W, TY);
    c.launches++;
    return;
  }
  static const int agg = getenv("MGP_AGG") ? atoi(getenv("MGP_AGG")) : 1;     // developer knob
  k_deposit_atomic<T><<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.pA, grid, c.N, c.NZ, c.nx, c.x0, single, scale, W,
                                                              agg);
  c.launches++;
}

// ------------------------------------------------------------------ redshift-space deposit (PtoMesh_RSD, compute_pofk.c:280-393)

// Line of sight = y (axis 1: y and z swap roles) or z (axis 2).  The velocity shift leaves the x-slab of a
// particle unchanged, which is why the reference skips the x axis (compute_pofk.c:450-451).
//   V = Vel[axis] + COLA LPT velocity;  Z += V * vnorm;  one periodic wrap;  CIC as PtoMesh
// SD = 1: the LPT velocity is P.dDdy + P.dD2dy (float add, compute_pofk.c:321); else D dDdy + D2 dD2dy (double).
template <typename T, int SD>
__global__ void __launch_bounds__(256)
k_deposit_rsd(size_t n, const float4 *__restrict__ pA, const float4 *__restrict__ pB, const float4 *__restrict__ pC,
              const float2 *__restrict__ pE, const float *__restrict__ s1, const float *__restrict__ s2, size_t cap,
              T *__restrict__ grid, int N, int NZ, int nx, int x0, int single_rank, double scale, double W, int axis,
              int usecola, double vnorm, double dDdy, double dD2dy) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    const float4 p = pA[i], v = pB[i];
    const double X = (double) p.x * scale;
    const double Y = (double) (axis == 1 ? p.z : p.y) * scale;
    double Z = (double) (axis == 1 ? p.y : p.z) * scale;
    const float vel = axis == 1 ? v.y : v.z;
    double V;
    if (!usecola) {
      V = (double) vel;
    } else if (SD) {
      const float a = s1[(size_t) axis * cap + i], b = s2 ? s2[(size_t) axis * cap + i] : 0.0f;
      V = (double) __fadd_rn(vel, __fadd_rn(a, b));
    } else {
      const float4 d = pC[i];
      const float2 e = pE[i];
      const double d1 = (double) (axis == 1 ? d.y : d.z), d2 = (double) (axis == 1 ? e.x : e.y);
      V = __dadd_rn((double) vel, __dadd_rn(__dmul_rn(d1, dDdy), __dmul_rn(d2, dD2dy)));
    }
    V *= vnorm;
    Z += V;
    if (Z >= (double) N) Z -= (double) N;
    if (Z < 0) Z += (double) N;
    unsigned ix = (unsigned) X, iy = (unsigned) Y, iz = (unsigned) Z;
    const double dx = X - (double) ix, dz = Z - (double) iz;
    double dy = Y - (double) iy;
    const double tx = 1.0 - dx, tz = 1.0 - dz;
    double ty = 1.0 - dy;
    dy *= W; ty *= W;
    const unsigned lx = ix - (unsigned) x0;
    if (iy >= (unsigned) N) iy = 0;
    if (iz >= (unsigned) N) iz = 0;
    unsigned lx1 = lx + 1;
    if (single_rank && lx1 == (unsigned) nx) lx1 = 0;
    const unsigned iy1 = (iy + 1 == (unsigned) N) ? 0 : iy + 1, iz1 = (iz + 1 == (unsigned) N) ? 0 : iz + 1;
    const size_t rz = (size_t) 2 * NZ;
    T *r00 = grid + ((size_t) lx * N + iy) * rz, *r01 = grid + ((size_t) lx * N + iy1) * rz;
    T *r10 = grid + ((size_t) lx1 * N + iy) * rz, *r11 = grid + ((size_t) lx1 * N + iy1) * rz;
    atomicAdd(r00 + iz, (T) (tx * ty * tz)); atomicAdd(r00 + iz1, (T) (tx * ty * dz));
    atomicAdd(r01 + iz, (T) (tx * dy * tz)); atomicAdd(r01 + iz1, (T) (tx * dy * dz));
    atomicAdd(r10 + iz, (T) (dx * ty * tz)); atomicAdd(r10 + iz1, (T) (dx * ty * dz));
    atomicAdd(r11 + iz, (T) (dx * dy * tz)); atomicAdd(r11 + iz1, (T) (dx * dy * dz));
  }
}

template <typename T>
static void deposit_rsd_t(Ctx &c, int gid, int axis, double vnorm, double dDdy, double dD2dy) {
  T *grid = (T *) c.grid[gid];
  const double scale = (double) c.N / c.cfg.box;
  const double r = (double) c.N / (double) c.cfg.nsample;
  const double W = r * r * r;
  k_fill<T><<<grid_for(c.grid_vals, 256), 256, 0, c.stream>>>(grid, c.grid_vals, (T) -1.0);
  c.launches++;
  if (!c.np) return;
  const unsigned g = grid_for(c.np, 256);
  if (c.cfg.scale_dependent)
    k_deposit_rsd<T, 1><<<g, 256, 0, c.stream>>>(c.np, c.pA, c.pB, nullptr, nullptr, c.sdf[2], c.sd_zero[3] ? nullptr : c.sdf[3], c.cap,
                                                grid, c.N, c.NZ, c.nx, c.x0, c.P == 1, scale, W, axis, c.cfg.use_cola, vnorm, dDdy, dD2dy);
  else
    k_deposit_rsd<T, 0><<<g, 256, 0, c.stream>>>(c.np, c.pA, c.pB, c.pC, (const float2 *) c.pE, nullptr, nullptr, c.cap, grid, c.N,
                                                c.NZ, c.nx, c.x0, c.P == 1, scale, W, axis, c.cfg.use_cola, vnorm, dDdy, dD2dy);
  c.launches++;
}

void deposit_rsd(Ctx &c, int grid_id, int axis, double vnorm, double dDdy, double dD2dy) {
  PhaseTimer t(c, PH_PTOMESH);
  REQUIRE(axis == 1 || axis == 2, MGP_ERR_INVALID, "redshift-space line of sight must be y (1) or z (2)");
  if (c.cfg.scale_dependent && c.cfg.use_cola)
    REQUIRE(c.sd_set[2] && c.sd_set[3], MGP_ERR_STATE, "RSD P(k) (scale-dependent): assign FIELD_dDdy first");
  if (c.gbytes == 4) deposit_rsd_t<float>(c, grid_id, axis, vnorm, dDdy, dD2dy);
  else deposit_rsd_t<double>(c, grid_id, axis, vnorm, dDdy, dD2dy);
}

void deposit_density(Ctx &c, int grid_id) {
  if (c.cfg.deposit_mode == MGP_DEPOSIT_ROWS && deposit_rows_supported(c)) { deposit_rows(c, grid_id); return; }
  if (c.gbytes == 4) deposit_t<float>(c, grid_id); else deposit_t<double>(c, grid_id);
}

// ------------------------------------------------------------------ gather (MtoParticles)

template <typename T>
__global__ void __launch_bounds__(256)
k_gather(size_t n, const float4 *__restrict__ pA, const T *__restrict__ fx, const T *__restrict__ fy,
         const T *__restrict__ fz, float *__restrict__ disp, size_t cap, int N, int NZ, int x0, double scale,
         double *__restrict__ partial) {
  __shared__ double sm[3][8];
  double sx = 0, sy = 0, sz = 0;
  const size_t rz = (size_t) 2 * NZ;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    const Cic q = cic_of(pA[i], scale, (unsigned) N, 1.0);
    const unsigned lx = q.ix - (unsigned) x0, lx1 = lx + 1;     // ghost plane nx holds the right neighbour's plane 0
    const unsigned iy1 = (q.iy + 1 == (unsigned) N) ? 0 : q.iy + 1, iz1 = (q.iz + 1 == (unsigned) N) ? 0 : q.iz + 1;
    const size_t o00 = ((size_t) lx * N + q.iy) * rz, o01 = ((size_t) lx * N + iy1) * rz;
    const size_t o10 = ((size_t) lx1 * N + q.iy) * rz, o11 = ((size_t) lx1 * N + iy1) * rz;
    const double a = q.tx * q.ty, b = q.tx * q.dy, cc = q.dx * q.ty, d = q.dx * q.dy;
    const double w0 = a * q.tz, w1 = a * q.dz, w2 = b * q.tz, w3 = b * q.dz;
    const double w4 = cc * q.tz, w5 = cc * q.dz, w6 = d * q.tz, w7 = d * q.dz;
#define GATHER(F)                                                                                         \
  ((double) F[o00 + q.iz] * w0 + (double) F[o00 + iz1] * w1 + (double) F[o01 + q.iz] * w2 +               \
   (double) F[o01 + iz1] * w3 + (double) F[o10 + q.iz] * w4 + (double) F[o10 + iz1] * w5 +                \
   (double) F[o11 + q.iz] * w6 + (double) F[o11 + iz1] * w7)
    const float gx = (float) GATHER(fx), gy = (float) GATHER(fy), gz = (float) GATHER(fz);
#undef GATHER
    disp[i] = gx; disp[cap + i] = gy; disp[2 * cap + i] = gz;
    sx += (double) gx; sy += (double) gy; sz += (double) gz;    // sumDxyz += Disp (float values)
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sm[0][w] = sx; sm[1][w] = sy; sm[2][w] = sz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, cc = 0;
    for (int k = 0; k < (int) (blockDim.x >> 5); k++) { a += sm[0][k]; b += sm[1][k]; cc += sm[2][k]; }
    partial[3 * blockIdx.x] = a; partial[3 * blockIdx.x + 1] = b; partial[3 * blockIdx.x + 2] = cc;
  }
}

void gather_forces(Ctx &c, double sumD[3]) {
  static const bool rows_env = !(getenv("MGP_GATHER_ROWS") && atoi(getenv("MGP_GATHER_ROWS")) == 0);     // developer knob
  const bool use_rows = rows_env && c.cfg.deposit_mode == MGP_DEPOSIT_ROWS && c.np && gather_rows_supported(c);
  if (use_rows) rows_bin(c);                       // normally still valid from PtoMesh
  PhaseTimer t(c, PH_MTOP);
  const size_t n = c.np;
  unsigned g = grid_for(n, 256, 8);
  reduce_alloc(c, (size_t) g * 3 + 16);
  const double scale = (double) c.N / c.cfg.box;
  if (use_rows)
    g = gather_rows(c);
  else if (c.gbytes == 4)
    k_gather<float><<<g, 256, 0, c.stream>>>(n, c.pA, (const float *) c.grid[1], (const float *) c.grid[2],
                                             (const float *) c.grid[3], c.disp, c.cap, c.N, c.NZ, c.x0, scale, c.d_red);
  else
    k_gather<double><<<g, 256, 0, c.stream>>>(n, c.pA, (const double *) c.grid[1], (const double *) c.grid[2],
                                              (const double *) c.grid[3], c.disp, c.cap, c.N, c.NZ, c.x0, scale, c.d_red);
  double *res = c.d_red + (size_t) g * 3;
  k_final_reduce<<<1, 256, 0, c.stream>>>(c.d_red, (int) g, 3, 3, 1.0, res);
  c.launches += use_rows ? 1 : 2;
  allreduce_sum(c, res, 3);
  CK(cudaMemcpyAsync(c.h_red, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  const double tot = (double) c.cfg.nsample * (double) c.cfg.nsample * (double) c.cfg.nsample;
  for (int a = 0; a < 3; a++) sumD[a] = c.h_red[a] / tot;       // sumDxyz /= TotNumPart (auxPM.c:639)
  c.have_disp = true;
}

}  // namespace mgp
