// Slab ownership and particle migration (MoveParticles, auxPM.c:108-275).
//
// Ownership rule, exactly the reference's (auxPM.c:151-153):
//     X = (int)(Pos[0] * (double)Nmesh / Box);   owner = Slab_to_task[X]
// with Slab_to_task the FFTW-MPI block distribution of 2LPT.c:83-99 (block = ceil(Nmesh / NTask)).
//
// The reference forwards whole particle records hop by hop around the ring until they reach their
// owner (auxPM.c:155-267).  On an NVSwitch box every GPU reaches every other at full bandwidth, so the
// records go straight to the owner in one grouped NCCL send/recv; the resulting ownership is the same.
//
//   1. k_migrate_count   per-destination leaver counts (warp-aggregated atomics)
//   2. ncclAllGather     the P x P count matrix -> every rank knows all send / receive sizes and
//                        can take the "increase Buffer" error decision consistently (no deadlock)
//   3. k_migrate_pack    leavers -> the sort's second buffer set, grouped by destination (SoA)
//   4. grouped ncclSend / ncclRecv: received fields land directly behind the live particles
//   5. the cell / bucket sort that follows recomputes ownership from the positions and drops every
//      particle this rank no longer owns, which also closes the holes the leavers left.
#include "common.cuh"

namespace mgp {

__device__ __forceinline__ int owner_of(float x, double scale, int N, int block, int P) {
  int X = (int) ((double) x * scale);
  if (X < 0) X = 0;
  if (X >= N) X = N - 1;            // cannot happen for float positions in [0, Box)
  int r = X / block;
  return r < P ? r : P - 1;
}

__global__ void __launch_bounds__(256)
k_migrate_count(size_t n, const float4 *__restrict__ pA, double scale, int N, int block, int P, int me,
                unsigned *__restrict__ cnt) {
  const unsigned lane = threadIdx.x & 31;
  const size_t nround = (n + 31) / 32 * 32;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < nround; i += (size_t) gridDim.x * blockDim.x) {
    int dst = me;
    if (i < n) dst = owner_of(pA[i].x, scale, N, block, P);
    const unsigned leaving = __ballot_sync(0xffffffffu, dst != me);
    if (!leaving) continue;
    const unsigned peers = __match_any_sync(0xffffffffu, dst);
    if (dst != me && (int) lane == __ffs(peers) - 1) atomicAdd(&cnt[dst], (unsigned) __popc(peers));
  }
}

// send slot = send_off[dst] + atomic cursor; the order inside a destination segment is irrelevant
// because the receiver sorts.
__global__ void __launch_bounds__(256)
k_migrate_pack(size_t n, const float4 *__restrict__ pA, const float4 *__restrict__ pB, const float4 *__restrict__ pC,
               const float2 *__restrict__ pE, double scale, int N, int block, int P, int me,
               const unsigned *__restrict__ send_off, unsigned *__restrict__ cursor, float4 *__restrict__ sA,
               float4 *__restrict__ sB, float4 *__restrict__ sC, float2 *__restrict__ sE) {
  const unsigned lane = threadIdx.x & 31;
  const size_t nround = (n + 31) / 32 * 32;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < nround; i += (size_t) gridDim.x * blockDim.x) {
    int dst = me;
    float4 a = make_float4(0, 0, 0, 0);
    if (i < n) { a = pA[i]; dst = owner_of(a.x, scale, N, block, P); }
    const unsigned leaving = __ballot_sync(0xffffffffu, dst != me);
    if (!leaving) continue;
    const unsigned peers = __match_any_sync(0xffffffffu, dst);
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (dst != me && (int) lane == leader) base = atomicAdd(&cursor[dst], (unsigned) __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (dst != me) {
      const size_t s = (size_t) send_off[dst] + base + (unsigned) __popc(peers & ((1u << lane) - 1u));
      sA[s] = a; sB[s] = pB[i]; sC[s] = pC[i]; sE[s] = pE[i];
    }
  }
}

// After the exchange the array holds [0, n): the old particles with the leavers still in place, and [n, n + nrecv):
// the arrivals.  Instead of a full sort, the leavers' slots below the new count are refilled with the live particles
// above it (order inside a cell-sorted array is only perturbed at ~1 % of the positions, which the warp-aggregated
// deposit and the gather tolerate; the periodic sort restores it).
__global__ void __launch_bounds__(256)
k_compact_lists(size_t n, size_t ntot, size_t newn, const float4 *__restrict__ pA, double scale, int N, int block, int P, int me,
                unsigned *__restrict__ counters, uint32_t *__restrict__ holes, uint32_t *__restrict__ movers) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < ntot; i += (size_t) gridDim.x * blockDim.x) {
    const bool leaver = i < n && owner_of(pA[i].x, scale, N, block, P) != me;
    if (i < newn) { if (leaver) holes[atomicAdd(&counters[0], 1u)] = (uint32_t) i; }
    else if (!leaver) movers[atomicAdd(&counters[1], 1u)] = (uint32_t) i;
  }
}

__global__ void __launch_bounds__(256)
k_compact_move(const unsigned *__restrict__ counters, const uint32_t *__restrict__ holes, const uint32_t *__restrict__ movers,
               float4 *__restrict__ pA, float4 *__restrict__ pB, float4 *__restrict__ pC, float2 *__restrict__ pE) {
  const unsigned h = counters[0] < counters[1] ? counters[0] : counters[1];
  for (size_t k = blockIdx.x * (size_t) blockDim.x + threadIdx.x; k < h; k += (size_t) gridDim.x * blockDim.x) {
    const uint32_t d = holes[k], s = movers[k];
    pA[d] = pA[s]; pB[d] = pB[s];
    if (pC) { pC[d] = pC[s]; pE[d] = pE[s]; }
  }
}

void particles_migrate(Ctx &c) {
  if (c.P == 1) return;
  const int P = c.P, me = c.rank;
  const size_t n = c.np;
  const double scale = (double) c.N / c.cfg.box;
  const int block = (c.N + P - 1) / P;
  if (!c.mig_dev) {
    CK(cudaMalloc(&c.mig_dev, (size_t) (P * P + 3 * P + 2) * sizeof(unsigned)));
    CK(cudaMallocHost(&c.mig_host, (size_t) (P * P + 3 * P + 2) * sizeof(unsigned)));
  }
  unsigned *d_cnt = c.mig_dev;              // [P]      my per-destination counts
  unsigned *d_all = c.mig_dev + P;          // [P][P]   all ranks' counts (row = sender)
  unsigned *d_off = c.mig_dev + P + P * P;  // [P]      send offsets
  unsigned *d_cur = d_off + P;              // [P]      pack cursors
  CK(cudaMemsetAsync(d_cnt, 0, P * sizeof(unsigned), c.stream));
  if (n) k_migrate_count<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.pA, scale, c.N, block, P, me, d_cnt);
  c.launches++;
  CKNCCL(ncclAllGather(d_cnt, d_all, P, ncclUint32, c.comm, c.stream));
  CK(cudaMemcpyAsync(c.mig_host, d_all, (size_t) P * P * sizeof(unsigned), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  const unsigned *all = c.mig_host;         // all[s * P + d] = particles s sends to d
  // global decisions first (identical on every rank)
  unsigned long long moved_total = 0;
  bool overflow = false;
  for (int r = 0; r < P; r++) {
    unsigned long long in = 0, out = 0;
    for (int q = 0; q < P; q++) { in += all[q * P + r]; out += all[r * P + q]; }
    moved_total += out;
    // every rank knows its own capacity only; the reference aborts all tasks when any overflows
    // (auxPM.c:250-254), here each rank tests itself and the result is combined below
    (void) in;
  }
  unsigned long long nrecv = 0, nsend = 0;
  for (int q = 0; q < P; q++) { nrecv += all[q * P + me]; nsend += all[me * P + q]; }
  overflow = (n + nrecv > c.cap) || (nsend > c.cap);
  // combine the overflow flags so that all ranks fail (or proceed) together
  int *d_flag = c.d_flag;
  c.h_flag[0] = overflow ? 1 : 0;
  CK(cudaMemcpyAsync(d_flag, c.h_flag, sizeof(int), cudaMemcpyHostToDevice, c.stream));
  CKNCCL(ncclAllReduce(d_flag, d_flag, 1, ncclInt, ncclMax, c.comm, c.stream));
  CK(cudaMemcpyAsync(c.h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  if (c.h_flag[0]) {
    char msg[256];
    snprintf(msg, sizeof(msg), "MoveParticles: rank %d would hold %llu particles (+%llu arriving) but has room for %llu: "
             "increase Buffer", me, (unsigned long long) n, nrecv, (unsigned long long) c.cap);
    throw Error(MGP_ERR_BUFFER, msg);
  }
  c.last_moved = moved_total;
  if (moved_total == 0) return;

  std::vector<unsigned> off(2 * P, 0);
  for (int q = 1; q < P; q++) off[q] = off[q - 1] + all[me * P + (q - 1)];
  CK(cudaMemcpyAsync(d_off, off.data(), P * sizeof(unsigned), cudaMemcpyHostToDevice, c.stream));
  CK(cudaMemsetAsync(d_cur, 0, P * sizeof(unsigned), c.stream));
  if (nsend) {
    k_migrate_pack<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.pA, c.pB, c.pC, (const float2 *) c.pE, scale, c.N, block, P, me,
                                                         d_off, d_cur, c.pA2, c.pB2, c.pC2, (float2 *) c.pE2);
    c.launches++;
  }
  {
    PhaseTimer t(c, PH_COMM);
    CKNCCL(ncclGroupStart());
    size_t roff = n;
    for (int q = 0; q < P; q++) {
      if (q == me) continue;
      const size_t ns = all[me * P + q], nr = all[q * P + me];
      if (ns) {
        CKNCCL(ncclSend(c.pA2 + off[q], ns * sizeof(float4), ncclChar, q, c.comm, c.stream));
        CKNCCL(ncclSend(c.pB2 + off[q], ns * sizeof(float4), ncclChar, q, c.comm, c.stream));
        CKNCCL(ncclSend(c.pC2 + off[q], ns * sizeof(float4), ncclChar, q, c.comm, c.stream));
        CKNCCL(ncclSend((float2 *) c.pE2 + off[q], ns * sizeof(float2), ncclChar, q, c.comm, c.stream));
      }
      if (nr) {
        CKNCCL(ncclRecv(c.pA + roff, nr * sizeof(float4), ncclChar, q, c.comm, c.stream));
        CKNCCL(ncclRecv(c.pB + roff, nr * sizeof(float4), ncclChar, q, c.comm, c.stream));
        CKNCCL(ncclRecv(c.pC + roff, nr * sizeof(float4), ncclChar, q, c.comm, c.stream));
        CKNCCL(ncclRecv((float2 *) c.pE + roff, nr * sizeof(float2), ncclChar, q, c.comm, c.stream));
        roff += nr;
      }
    }
    CKNCCL(ncclGroupEnd());
  }
  c.have_disp = false;
  c.sd_req_valid = false;
  const size_t ntot = n + nrecv, newn = n + nrecv - nsend;
  static const bool compact_env = !(getenv("MGP_COMPACT") && atoi(getenv("MGP_COMPACT")) == 0);
  const bool exact_needed = c.cfg.deposit_mode == MGP_DEPOSIT_TILE || c.cfg.deposit_mode == MGP_DEPOSIT_DETERMINISTIC;
  const bool sort_due = !c.cfg.sort_particles || c.drifts_since_sort >= c.cfg.sort_particles;
  if (compact_env && !exact_needed && !sort_due && c.cfg.sort_particles) {
    // hole compaction: leavers' slots below the new count are refilled from above it
    unsigned *d_two = c.mig_dev + P + P * P + 2 * P;          // two spare counters behind the pack cursors
    CK(cudaMemsetAsync(d_two, 0, 2 * sizeof(unsigned), c.stream));
    k_compact_lists<<<grid_for(ntot, 256), 256, 0, c.stream>>>(n, ntot, newn, c.pA, scale, c.N, block, P, me, d_two, c.key[0], c.perm[0]);
    k_compact_move<<<grid_for(nsend ? nsend : 1, 256), 256, 0, c.stream>>>(d_two, c.key[0], c.perm[0], c.pA, c.pB,
                                                                        c.cfg.scale_dependent ? nullptr : c.pC, (float2 *) c.pE);
    c.launches += 2;
    c.np = newn;
    c.np_after_sort = SIZE_MAX;
    c.sorted = false;
    c.bins_valid = false;
    return;
  }
  // the leavers are still in place; the sort that follows drops them (ownership is recomputed there)
  c.np = ntot;
  c.np_after_sort = newn;
  c.sorted = false;
  c.bins_valid = false;
  c.drifts_since_sort = 1 << 30;
}

}  // namespace mgp
