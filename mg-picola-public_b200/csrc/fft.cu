// FFTs, slab halos and scalar all-reduces.
//
// Replaces the seven FFTW-MPI wrappers of wrappers.c:26-96 (plans created once, not per step as
// MEMORY_MODE does, auxPM.c:46-51, 61-71) and the halo MPI_Sendrecv calls of auxPM.c:350-356
// (density ghost plane -> right neighbour) and auxPM.c:546-551 (force plane 0 -> left neighbour).
//
// Single rank: cuFFT 3-D in-place r2c / c2r on the padded layout [x][y][2*(N/2+1)], identical to
// the FFTW in-place layout; unnormalised in both directions like FFTW.
// Several ranks (x-slabs exactly as fftw_mpi_local_size_3d): batched 2-D (y,z) transforms on the
// local planes, a hand-written pack / unpack pair around an NCCL all-to-all, and batched 1-D
// transforms along x.  k-space is kept in the transposed layout [ky_local][kz][kx] -- every
// k-space kernel is pointwise, so FFTW's second transpose back is never needed.
#include "common.cuh"
#include "reduce.cuh"

#include <cmath>

namespace mgp {

typedef long long int lli;
static size_t make_plan_many(cufftHandle *plan, int rank, lli *n, lli *inembed, lli istride, lli idist, lli *onembed,
                             lli ostride, lli odist, cufftType type, lli batch, cudaStream_t st) {
  size_t ws = 0;
  CKFFT(cufftCreate(plan));
  CKFFT(cufftSetAutoAllocation(*plan, 0));
  CKFFT(cufftMakePlanMany64(*plan, rank, n, inembed, istride, idist, onembed, ostride, odist, type, batch, &ws));
  CKFFT(cufftSetStream(*plan, st));
  return ws;
}

// ------------------------------------------------------------------ peer-memory plumbing (P > 1)

static void p2p_teardown(Ctx &c) {
  for (int r = 0; r < 16; r++) {
    if (r != c.rank && c.peer_tbuf[r]) cudaIpcCloseMemHandle(c.peer_tbuf[r]);
    if (r != c.rank && c.peer_flags[r]) cudaIpcCloseMemHandle(c.peer_flags[r]);
    c.peer_tbuf[r] = nullptr; c.peer_flags[r] = nullptr;
  }
  cudaFree(c.sync_flags); c.sync_flags = nullptr;
  c.p2p = false;
}

// Maps every peer's transpose buffer and flag array into this process (cudaIpc over NVLink / NVSwitch).  All ranks
// agree on the outcome: if any mapping fails (no peer access, MGP_P2P=0) everybody keeps the NCCL all-to-all.
static void p2p_setup(Ctx &c) {
  const int P = c.P;
  const char *env = getenv("MGP_P2P");
  int ok = (P <= 16) && !(env && atoi(env) == 0);
  struct Handles { cudaIpcMemHandle_t buf, flags; };
  static_assert(sizeof(Handles) == 128, "cudaIpcMemHandle_t is 64 bytes");
  CK(cudaMalloc(&c.sync_flags, 16 * sizeof(uint32_t)));
  CK(cudaMemset(c.sync_flags, 0, 16 * sizeof(uint32_t)));
  if (P == 1) {                      // MGP_FORCE_SLAB on one rank: the only peer is this rank itself
    c.peer_tbuf[0] = c.tbuf_a; c.peer_flags[0] = c.sync_flags;
    c.p2p = true;
    return;
  }
  Handles mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && cudaIpcGetMemHandle(&mine.buf, c.tbuf_a) != cudaSuccess) ok = 0;
  if (ok && cudaIpcGetMemHandle(&mine.flags, c.sync_flags) != cudaSuccess) ok = 0;
  cudaGetLastError();
  Handles *d_all = nullptr;
  CK(cudaMalloc(&d_all, (size_t) (P + 1) * sizeof(Handles)));
  CK(cudaMemcpy(d_all + P, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  CKNCCL(ncclAllGather(d_all + P, d_all, sizeof(Handles), ncclChar, c.comm, c.stream));
  std::vector<Handles> all(P);
  CK(cudaMemcpyAsync(all.data(), d_all, (size_t) P * sizeof(Handles), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  for (int r = 0; r < P && ok; r++) {
    if (r == c.rank) { c.peer_tbuf[r] = c.tbuf_a; c.peer_flags[r] = c.sync_flags; continue; }
    void *a = nullptr, *b = nullptr;
    if (cudaIpcOpenMemHandle(&a, all[r].buf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
        cudaIpcOpenMemHandle(&b, all[r].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
    c.peer_tbuf[r] = a; c.peer_flags[r] = (uint32_t *) b;
  }
  cudaGetLastError();
  int *d_ok = (int *) d_all;
  CK(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
  CKNCCL(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c.comm, c.stream));
  CK(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  CK(cudaFree(d_all));
  c.p2p = ok != 0;
  if (!c.p2p) { for (int r = 0; r < 16; r++) if (r == c.rank) { c.peer_tbuf[r] = nullptr; c.peer_flags[r] = nullptr; } }
}

// Barrier over peer memory: thread r stores the epoch into rank r's flag array (slot = my rank) and then waits
// until rank r has stored it into mine.  Everything this rank wrote to peer memory in earlier kernels of the stream
// is ordered before the signal by the system-scope release.
__global__ void k_p2p_barrier(PeerPtrs flags, uint32_t *mine, int P, int me, uint32_t epoch) {
  const int r = threadIdx.x;
  if (r < P) {
    __threadfence_system();
    uint32_t *dst = (uint32_t *) flags.p[r] + me;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine + r) : "memory");
    } while ((int32_t) (v - epoch) < 0);
  }
}

static void p2p_barrier(Ctx &c, cudaStream_t st = nullptr) {
  PeerPtrs f;
  for (int r = 0; r < 16; r++) f.p[r] = c.peer_flags[r];
  c.sync_epoch++;
  k_p2p_barrier<<<1, 32, 0, st ? st : c.stream>>>(f, c.sync_flags, c.P, c.rank, c.sync_epoch);
  c.launches++;
}

// forward transpose, fused with the exchange: in = [nx][N (ky)][NZ] on this rank; the element (x, ky, kz) goes to the
// rank that owns ky, into its buffer laid out [nyl][NZ][N (kx)].  32 x 32 (x, kz) tiles through shared memory: reads
// are 512-byte runs along kz, the peer stores 512-byte runs along kx.
template <typename C>
__global__ void k_transpose_fwd_p2p(const C *__restrict__ in, PeerPtrs out, int nxb, int x0, int N, int NZ, int nyl) {
  __shared__ C tile[32][33];
  const int j = blockIdx.z, r = j / nyl, jl = j - r * nyl;
  C *dst = (C *) out.p[r];
  const int xb = blockIdx.x * 32, kb = blockIdx.y * 32;
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int x = xb + q, k = kb + threadIdx.x;
    if (x < nxb && k < NZ) tile[q][threadIdx.x] = in[((size_t) x * N + j) * NZ + k];
  }
  __syncthreads();
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int k = kb + q, x = xb + threadIdx.x;
    if (x < nxb && k < NZ) dst[((size_t) jl * NZ + k) * N + x0 + x] = tile[threadIdx.x][q];
  }
}

// backward: in = [nyl][NZ][N (x)] on this rank; (ky, kz, x) goes to the owner of x, laid out [nxb][N (ky)][NZ]
template <typename C>
__global__ void k_transpose_bwd_p2p(const C *__restrict__ in, PeerPtrs out, int nxb, int y0, int N, int NZ, int nyl) {
  __shared__ C tile[32][33];
  const int jl = blockIdx.z;
  const int xb = blockIdx.x * 32, kb = blockIdx.y * 32;
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int k = kb + q, x = xb + threadIdx.x;
    if (x < N && k < NZ) tile[q][threadIdx.x] = in[((size_t) jl * NZ + k) * N + x];
  }
  __syncthreads();
  for (int q = threadIdx.y; q < 32; q += blockDim.y) {
    const int x = xb + q, k = kb + threadIdx.x;
    if (x < N && k < NZ) {
      const int s = x / nxb, xl = x - s * nxb;
      ((C *) out.p[s])[((size_t) xl * N + (y0 + jl)) * NZ + k] = tile[threadIdx.x][q];
    }
  }
}

// ------------------------------------------------------------------ fused x-transform + exchange (xfft.cuh)

// one switch over the supported Nmesh = 2^lgn; OP is a functor template taking <C, LGN>
#define XF_DISPATCH_LGN(lgn, OP)                                                                             \
  switch (lgn) {                                                                                             \
    case 4: OP(4); break; case 5: OP(5); break; case 6: OP(6); break; case 7: OP(7); break; case 8: OP(8); break; \
    case 9: OP(9); break; case 10: OP(10); break; case 11: OP(11); break; case 12: OP(12); break;             \
    default: throw mgp::Error(MGP_ERR_STATE, "fused x-transform: unsupported Nmesh");                        \
  }

template <typename C, int LGN>
static bool xfft_prepare(Ctx &c) {
  constexpr int TK = xf::tile_lines(LGN, sizeof(C));
  if constexpr (TK == 0) {
    return false;
  } else {
    c.xf_tk = TK;
    c.xf_smem = ((size_t) TK << LGN) * sizeof(C);
    CK(cudaFuncSetAttribute(xf::k_xfft_bwd_p2p<C, LGN, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.xf_smem));
    CK(cudaFuncSetAttribute(xf::k_xfft_fwd_p2p<C, LGN, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.xf_smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, xf::k_xfft_bwd_p2p<C, LGN, TK>, xf::kThreads, c.xf_smem));
    REQUIRE(occ >= 1, MGP_ERR_CUDA, "fused x-transform: kernel does not fit on an SM");
    const long long ntiles = (long long) c.ny_loc * xf::tiles_per_line(c.NZ, TK, 128 / (int) sizeof(C));
    // persistent grid; MGP_XFFT_TRIM CTAs fewer than the machine holds, so that the one-warp flag-barrier kernels of the
    // communication stream find a free slot while an x-transform kernel owns every register of the other SMs
    const char *tr = getenv("MGP_XFFT_TRIM");
    // CTAs per SM: on several ranks the kernel waits on NVLink most of the time; one CTA per SM leaves half of the
    // registers and shared memory to the 2-D cuFFT kernels of the other components, which then overlap the exchange
    // (2 x B200, 512^3: batched inverse transform 3.34 -> 2.88 ms); alone on the GPU two CTAs per SM are faster
    int cps_want = c.P > 1 ? 1 : 2;
    if (const char *cps = getenv("MGP_XFFT_CPS")) cps_want = atoi(cps);
    if (cps_want >= 1 && cps_want < occ) occ = cps_want;
    long long g = (long long) kSMs * occ - (tr ? atoi(tr) : 1);
    if (g < 1) g = 1;
    c.xf_grid = (int) (ntiles < g ? ntiles : g);
    return true;
  }
}

// Picks the kernel instance of this Nmesh (tile = TK lines of N complex values, at most 64 KB so that two CTAs share
// an SM) and uploads the twiddle tables.  Leaves xf_on = false when Nmesh is not a power of two in [16, 4096], the
// slabs are not a power of two wide, or the peer-memory path is off: those cases keep cuFFT's 1-D plan + the
// transpose kernels.
static void xfft_setup(Ctx &c) {
  c.xf_on = false;
  const char *env = getenv("MGP_XFFT");
  if (!c.p2p || (env && atoi(env) == 0)) return;
  int lgn = 0;
  while ((1 << lgn) < c.N) lgn++;
  if ((1 << lgn) != c.N) {
    // mesh sizes with factors 3 and 5 (320, 400, 640, 800): mixed-radix instances, parity-checked on the GPU
    // (tests/test_slab_fused.py) but never timed on several ranks: opt-in
    const char *mx = getenv("MGP_XFFT_MIXED");
    if (mx && atoi(mx) != 0 && xfm_supported(c.N) && xfm_prepare(c)) { c.xf_mixed = true; c.xf_on = true; }
    return;
  }
  if (lgn < xf::kMinLgN || lgn > xf::kMaxLgN) return;
  int lgx = 0;
  while ((1 << lgx) < c.nx) lgx++;
  if ((1 << lgx) != c.nx || c.nx * c.P != c.N) return;
  c.xf_lgn = lgn; c.xf_lgnxb = lgx;
  bool ok = false;
  if (c.gbytes == 4) {
#define OP(L) ok = xfft_prepare<float2, L>(c)
    XF_DISPATCH_LGN(lgn, OP)
#undef OP
  } else {
#define OP(L) ok = xfft_prepare<double2, L>(c)
    XF_DISPATCH_LGN(lgn, OP)
#undef OP
  }
  if (!ok) return;
  const int np = xf::plan_npass(lgn);
  const int total = xf::plan_twtotal(lgn);
  std::vector<double2> tw(total);
  for (int i = 0; i < np; i++) {
    const int L = 1 << (xf::plan_lgR(lgn, i) + xf::plan_lgM(lgn, i));
    for (int t = 0; t < L; t++) {
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) t / (long double) L;
      tw[xf::plan_twoff(lgn, i) + t] = make_double2((double) cosl(a), (double) sinl(a));
    }
  }
  const size_t cb = c.gbytes == 4 ? sizeof(float2) : sizeof(double2);
  CK(cudaMalloc(&c.xf_tw, (size_t) total * cb));
  if (c.gbytes == 4) {
    std::vector<float2> twf(total);
    for (int t = 0; t < total; t++) twf[t] = make_float2((float) tw[t].x, (float) tw[t].y);
    CK(cudaMemcpy(c.xf_tw, twf.data(), (size_t) total * cb, cudaMemcpyHostToDevice));
  } else {
    CK(cudaMemcpy(c.xf_tw, tw.data(), (size_t) total * cb, cudaMemcpyHostToDevice));
  }
  c.xf_on = true;
  {
    // wide tiles (128 KB, twice the run length on the exchange side): on 8 GPUs at 1024^3 the fused kernels move 565 GB/s
    // per direction instead of 374 - 475 (profiles/r02_exchange_8gpu.md) and the step drops from 67.5 to 61.9 ms; default
    // from Nmesh = 1024 on several ranks, MGP_XFFT_WIDE = 0 / 1 overrides
    const char *w = getenv("MGP_XFFT_WIDE");
    const bool want = w ? atoi(w) != 0 : (c.P > 1 && c.N >= 1024);
    if (want && xfw_prepare(c)) c.xf_wide = true;
  }
}

// staging slot `slot` (0..2) of rank r: the second half of its transpose buffer
static inline char *stage_of(Ctx &c, int r, int slot) { return (char *) c.peer_tbuf[r] + (size_t) (3 + slot) * c.grid_bytes(); }

// One strided block per peer through the copy engines (cudaMemcpy2DAsync on the peer mappings), forked from `st` onto
// the copy streams and joined back.  Block of (this rank -> rank r): [nxb x-planes][nyl ky rows][NZ] complex.
//   backward: staging slot (packed [r][xl][jl][kz] by the x-transform kernel) -> r's landing slot [xl][ky0_me + jl][kz]
//   forward : this rank's 2-D transform output `src2d` [xl][ky][kz], rows of r -> r's staging slot 0 at [me][xl][jl][kz]
static void exchange_dma(Ctx &c, cudaStream_t st, bool forward, const void *src2d, int slot) {
  const size_t cb = c.gbytes == 4 ? sizeof(float2) : sizeof(double2);
  const size_t row = (size_t) c.ny_loc * c.NZ * cb;           // one x-plane of one rank's ky rows: contiguous
  const size_t plane = (size_t) c.N * c.NZ * cb;              // one full x-plane
  const size_t blk = (size_t) c.nx * row;
  CK(cudaEventRecord(c.ev_cp_go, st));
  for (int q = 0; q < 4; q++) CK(cudaStreamWaitEvent(c.cp_stream[q], c.ev_cp_go, 0));
  for (int i = 0; i < c.P; i++) {
    const int r = (c.rank + 1 + i) % c.P;                      // staggered: no two ranks start on the same destination
    cudaStream_t cs = c.cp_stream[i & 3];
    if (forward) {
      const char *src = (const char *) src2d + (size_t) r * row;
      char *dst = stage_of(c, r, 0) + (size_t) c.rank * blk;
      CK(cudaMemcpy2DAsync(dst, row, src, plane, row, (size_t) c.nx, cudaMemcpyDefault, cs));
    } else {
      const char *src = stage_of(c, c.rank, slot) + (size_t) r * blk;
      char *dst = (char *) c.peer_tbuf[r] + (size_t) slot * c.grid_bytes() + (size_t) c.rank * row;
      CK(cudaMemcpy2DAsync(dst, plane, src, row, row, (size_t) c.nx, cudaMemcpyDefault, cs));
    }
  }
  for (int q = 0; q < 4; q++) {
    CK(cudaEventRecord(c.ev_cp_done[q], c.cp_stream[q]));
    CK(cudaStreamWaitEvent(st, c.ev_cp_done[q], 0));
  }
}

// backward x-transform of the local transposed k-space `in`, every x stored into slot `slot` of its owner's buffer
template <typename C>
static void xfft_bwd(Ctx &c, const void *in, int slot, cudaStream_t st) {
  PeerPtrs pp;
  // DMA exchange: the "owners" are the per-destination blocks of this rank's staging slot, [r][xl][jl][kz]
  const size_t blk = (size_t) c.nx * c.ny_loc * c.NZ * sizeof(C);
  for (int r = 0; r < 16; r++) {
    if (c.xf_dma) pp.p[r] = r < c.P ? stage_of(c, c.rank, slot) + (size_t) r * blk : nullptr;
    else pp.p[r] = c.peer_tbuf[r] ? (char *) c.peer_tbuf[r] + (size_t) slot * c.grid_bytes() : nullptr;
  }
  const int y0 = c.xf_dma ? 0 : c.y0, NY = c.xf_dma ? c.ny_loc : c.N;
  if (c.xf_mixed) { xfm_bwd(c, in, pp, y0, NY, st); c.launches++; return; }
  if (c.xf_wide) { xfw_bwd(c, in, pp, y0, NY, st); c.launches++; return; }
#define OP(L)                                                                                                   \
  xf::k_xfft_bwd_p2p<C, L, xf::tile_lines(L, sizeof(C)) ? xf::tile_lines(L, sizeof(C)) : 4>                      \
      <<<c.xf_grid, xf::kThreads, c.xf_smem, st>>>((const C *) in, pp, (const C *) c.xf_tw, c.xf_lgnxb, y0, NY, c.NZ, c.ny_loc)
  XF_DISPATCH_LGN(c.xf_lgn, OP)
#undef OP
  CK(cudaGetLastError());
  c.launches++;
}

// forward x-transform: pulls slot 0 of every owner's buffer, writes the local transposed k-space `out`
template <typename C>
static void xfft_fwd(Ctx &c, void *out, cudaStream_t st) {
  PeerPtrs pp;
  const size_t blk = (size_t) c.nx * c.ny_loc * c.NZ * sizeof(C);
  for (int r = 0; r < 16; r++) {
    if (c.xf_dma) pp.p[r] = r < c.P ? stage_of(c, c.rank, 0) + (size_t) r * blk : nullptr;
    else pp.p[r] = c.peer_tbuf[r];
  }
  const int y0 = c.xf_dma ? 0 : c.y0, NY = c.xf_dma ? c.ny_loc : c.N;
  if (c.xf_mixed) { xfm_fwd(c, out, pp, y0, NY, st); c.launches++; return; }
  if (c.xf_wide) { xfw_fwd(c, out, pp, y0, NY, st); c.launches++; return; }
#define OP(L)                                                                                                   \
  xf::k_xfft_fwd_p2p<C, L, xf::tile_lines(L, sizeof(C)) ? xf::tile_lines(L, sizeof(C)) : 4>                      \
      <<<c.xf_grid, xf::kThreads, c.xf_smem, st>>>(pp, (C *) out, (const C *) c.xf_tw, c.xf_lgnxb, y0, NY, c.NZ, c.ny_loc)
  XF_DISPATCH_LGN(c.xf_lgn, OP)
#undef OP
  CK(cudaGetLastError());
  c.launches++;
}

void fft_setup(Ctx &c) {
  const int N = c.N, NZ = c.NZ;
  const bool f32 = c.gbytes == 4;
  size_t ws = 0, w;
  if (!c.slab) {
    lli n[3] = {N, N, N};
    lli rembed[3] = {N, N, 2 * NZ}, cembed[3] = {N, N, NZ};
    const lli rdist = (lli) c.grid_vals, cdist = (lli) (c.grid_vals / 2);
    w = make_plan_many(&c.plan_r2c, 3, n, rembed, 1, rdist, cembed, 1, cdist, f32 ? CUFFT_R2C : CUFFT_D2Z, 1, c.stream);
    ws = w > ws ? w : ws;
    w = make_plan_many(&c.plan_c2r, 3, n, cembed, 1, cdist, rembed, 1, rdist, f32 ? CUFFT_C2R : CUFFT_Z2D, 1, c.stream);
    ws = w > ws ? w : ws;
    // same transform, separate plan handle for out-of-place use (real input grid is preserved)
    w = make_plan_many(&c.plan_r2c_oop, 3, n, rembed, 1, rdist, cembed, 1, cdist, f32 ? CUFFT_R2C : CUFFT_D2Z, 1, c.stream);
    ws = w > ws ? w : ws;
    w = make_plan_many(&c.plan_c2r3, 3, n, cembed, 1, cdist, rembed, 1, rdist, f32 ? CUFFT_C2R : CUFFT_Z2D, 3, c.stream);
    ws = w > ws ? w : ws;
  } else {
    lli n2[2] = {N, N};
    lli rembed[2] = {N, 2 * NZ}, cembed[2] = {N, NZ};
    w = make_plan_many(&c.plan2d_r2c, 2, n2, rembed, 1, N * 2 * NZ, cembed, 1, N * NZ, f32 ? CUFFT_R2C : CUFFT_D2Z, c.nx, c.stream);
    ws = w > ws ? w : ws;
    w = make_plan_many(&c.plan2d_c2r, 2, n2, cembed, 1, N * NZ, rembed, 1, N * 2 * NZ, f32 ? CUFFT_C2R : CUFFT_Z2D, c.nx, c.stream);
    ws = w > ws ? w : ws;
    lli n1[1] = {N};
    w = make_plan_many(&c.plan1d_x, 1, n1, n1, 1, N, n1, 1, N, f32 ? CUFFT_C2C : CUFFT_Z2Z, c.ny_loc * NZ, c.stream);
    ws = w > ws ? w : ws;
    // three landing slots (the batched c2r pipelines its three transposes) + three staging slots when the DMA
    // exchange is selected (MGP_XFFT_DMA=1), one allocation so that a single peer mapping covers both
    {
      const char *dma = getenv("MGP_XFFT_DMA");
      c.xf_dma = dma ? (atoi(dma) != 0) : 0;
    }
    CK(cudaMalloc(&c.tbuf_a, (c.xf_dma ? 6 : 3) * c.grid_bytes()));
    CK(cudaMalloc(&c.tbuf_b, c.grid_bytes()));
    p2p_setup(c);
    xfft_setup(c);
    if (c.xf_on) {
      w = make_plan_many(&c.plan2d_r2c_oop, 2, n2, rembed, 1, N * 2 * NZ, cembed, 1, N * NZ, f32 ? CUFFT_R2C : CUFFT_D2Z, c.nx, c.stream);
      ws = w > ws ? w : ws;
    }
    if (c.p2p) {
      int lo = 0, hi = 0;
      CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CK(cudaStreamCreateWithPriority(&c.comm_stream, cudaStreamNonBlocking, hi));
      CK(cudaEventCreateWithFlags(&c.ev_cp_go, cudaEventDisableTiming));
      for (int q = 0; q < 4; q++) {
        CK(cudaStreamCreateWithFlags(&c.cp_stream[q], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c.ev_cp_done[q], cudaEventDisableTiming));
      }
      for (int a = 0; a < 3; a++) {
        CK(cudaEventCreateWithFlags(&c.ev_fft[a], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c.ev_tr[a], cudaEventDisableTiming));
      }
    }
  }
  if (ws) CK(cudaMalloc(&c.fft_work, ws));     // shared cuFFT work area
  cufftHandle all[8] = {c.plan_r2c, c.plan_c2r, c.plan_c2r3, c.plan2d_r2c, c.plan2d_c2r, c.plan1d_x, c.plan_r2c_oop, c.plan2d_r2c_oop};
  for (cufftHandle h : all)
    if (h && ws) CKFFT(cufftSetWorkArea(h, c.fft_work));
  c.have_plans = true;
}

void fft_teardown(Ctx &c) {
  p2p_teardown(c);
  for (int a = 0; a < 3; a++) { if (c.ev_fft[a]) cudaEventDestroy(c.ev_fft[a]); if (c.ev_tr[a]) cudaEventDestroy(c.ev_tr[a]); }
  if (c.comm_stream) cudaStreamDestroy(c.comm_stream);
  if (c.ev_cp_go) cudaEventDestroy(c.ev_cp_go);
  for (int q = 0; q < 4; q++) { if (c.cp_stream[q]) cudaStreamDestroy(c.cp_stream[q]); if (c.ev_cp_done[q]) cudaEventDestroy(c.ev_cp_done[q]); }
  cufftHandle all[8] = {c.plan_r2c, c.plan_c2r, c.plan_c2r3, c.plan2d_r2c, c.plan2d_c2r, c.plan1d_x, c.plan_r2c_oop, c.plan2d_r2c_oop};
  for (cufftHandle h : all)
    if (h) cufftDestroy(h);
  cudaFree(c.fft_work); cudaFree(c.tbuf_a); cudaFree(c.tbuf_b); cudaFree(c.xf_tw);
}

// ------------------------------------------------------------------ slab transposes (P > 1)

// pack for the forward transpose: in = [nx][N (ky)][NZ] complex; block for rank r holds
// [nx][ny_r][NZ] at complex offset nx * y0_r * NZ  (equal slabs: ny_r = N / P).
template <typename C>
__global__ void k_pack_fwd(const C *__restrict__ in, C *__restrict__ out, int nx, int N, int NZ, int nyb) {
  const size_t tot = (size_t) nx * N * NZ;
  for (size_t e = blockIdx.x * (size_t) blockDim.x + threadIdx.x; e < tot; e += (size_t) gridDim.x * blockDim.x) {
    const int k = (int) (e % NZ);
    const size_t t = e / NZ;
    const int j = (int) (t % N), x = (int) (t / N);
    const int r = j / nyb, jl = j - r * nyb;
    out[((size_t) r * nx + x) * nyb * NZ + (size_t) jl * NZ + k] = in[e];
  }
}

// unpack after the forward all-to-all: in = [s][nx_s][nyl][NZ]  ->  out = [nyl][NZ][N (kx)]
// 32x32 shared-memory tile transpose over (x, kz) for each ky so that both sides coalesce.
template <typename C>
__global__ void k_unpack_fwd(const C *__restrict__ in, C *__restrict__ out, int nxb, int N, int NZ, int nyl) {
  __shared__ C tile[32][33];
  const int jl = blockIdx.z;
  const int x0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = x0 + r, k = k0 + threadIdx.x;
    if (x < N && k < NZ) {
      const int s = x / nxb, xl = x - s * nxb;
      tile[r][threadIdx.x] = in[(((size_t) s * nxb + xl) * nyl + jl) * NZ + k];
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, x = x0 + threadIdx.x;
    if (x < N && k < NZ) out[((size_t) jl * NZ + k) * N + x] = tile[threadIdx.x][r];
  }
}

// pack for the backward transpose: in = [nyl][NZ][N (x)]  ->  out = [s][nx_s][nyl][NZ]
template <typename C>
__global__ void k_pack_bwd(const C *__restrict__ in, C *__restrict__ out, int nxb, int N, int NZ, int nyl) {
  __shared__ C tile[32][33];
  const int jl = blockIdx.z;
  const int x0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int k = k0 + r, x = x0 + threadIdx.x;
    if (x < N && k < NZ) tile[r][threadIdx.x] = in[((size_t) jl * NZ + k) * N + x];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int x = x0 + r, k = k0 + threadIdx.x;
    if (x < N && k < NZ) {
      const int s = x / nxb, xl = x - s * nxb;
      out[(((size_t) s * nxb + xl) * nyl + jl) * NZ + k] = tile[threadIdx.x][r];
    }
  }
}

// unpack after the backward all-to-all: in = [r][nx][ny_r][NZ]  ->  out = [nx][N (ky)][NZ]
template <typename C>
__global__ void k_unpack_bwd(const C *__restrict__ in, C *__restrict__ out, int nx, int N, int NZ, int nyb) {
  const size_t tot = (size_t) nx * N * NZ;
  for (size_t e = blockIdx.x * (size_t) blockDim.x + threadIdx.x; e < tot; e += (size_t) gridDim.x * blockDim.x) {
    const int k = (int) (e % NZ);
    const size_t t = e / NZ;
    const int j = (int) (t % N), x = (int) (t / N);
    const int r = j / nyb, jl = j - r * nyb;
    out[e] = in[((size_t) r * nx + x) * nyb * NZ + (size_t) jl * NZ + k];
  }
}

static void all_to_all(Ctx &c, const void *send, void *recv, size_t block_bytes) {
  PhaseTimer t(c, PH_COMM);
  CKNCCL(ncclGroupStart());
  for (int r = 0; r < c.P; r++) {
    CKNCCL(ncclSend((const char *) send + (size_t) r * block_bytes, block_bytes, ncclChar, r, c.comm, c.stream));
    CKNCCL(ncclRecv((char *) recv + (size_t) r * block_bytes, block_bytes, ncclChar, r, c.comm, c.stream));
  }
  CKNCCL(ncclGroupEnd());
}

template <typename R, typename C>
static void dist_r2c(Ctx &c, void *g) {
  const int N = c.N, NZ = c.NZ, nxb = c.nx, nyl = c.ny_loc;
  if (c.xf_on && c.xf_dma) {
    // 2-D r2c in place; the copy engines deliver every owner's ky rows into its staging slot; the fused kernel
    // gathers the x-lines from the (local) staging slot, transforms them and writes the transposed k-space into g
    if (sizeof(R) == 4) CKFFT(cufftExecR2C(c.plan2d_r2c, (cufftReal *) g, (cufftComplex *) g));
    else CKFFT(cufftExecD2Z(c.plan2d_r2c, (cufftDoubleReal *) g, (cufftDoubleComplex *) g));
    c.launches += 2;
    {
      PhaseTimer t(c, PH_COMM);
      p2p_barrier(c);                       // every rank has consumed its staging slot
      exchange_dma(c, c.stream, true, g, 0);
      p2p_barrier(c);                       // every rank's blocks have landed
    }
    xfft_fwd<C>(c, g, c.stream);
    return;
  }
  if (c.xf_on) {
    // 2-D r2c out of place into this rank's transpose buffer; the fused kernel pulls the x-lines from their owners
    if (sizeof(R) == 4) CKFFT(cufftExecR2C(c.plan2d_r2c_oop, (cufftReal *) g, (cufftComplex *) c.tbuf_a));
    else CKFFT(cufftExecD2Z(c.plan2d_r2c_oop, (cufftDoubleReal *) g, (cufftDoubleComplex *) c.tbuf_a));
    c.launches += 2;
    {
      PhaseTimer t(c, PH_COMM);
      p2p_barrier(c);                       // every rank's 2-D transform is complete and visible
      xfft_fwd<C>(c, g, c.stream);
      p2p_barrier(c);                       // every rank has finished reading this rank's buffer
    }
    return;
  }
  if (sizeof(R) == 4) CKFFT(cufftExecR2C(c.plan2d_r2c, (cufftReal *) g, (cufftComplex *) g));
  else CKFFT(cufftExecD2Z(c.plan2d_r2c, (cufftDoubleReal *) g, (cufftDoubleComplex *) g));
  if (c.p2p) {
    PeerPtrs pp;
    for (int r = 0; r < 16; r++) pp.p[r] = c.peer_tbuf[r];
    {
      PhaseTimer t(c, PH_COMM);
      p2p_barrier(c);                       // every rank has finished reading its transpose buffer
      dim3 gr((nxb + 31) / 32, (NZ + 31) / 32, N), bl(32, 8);
      k_transpose_fwd_p2p<C><<<gr, bl, 0, c.stream>>>((const C *) g, pp, nxb, c.x0, N, NZ, nyl);
      p2p_barrier(c);                       // every rank's stores have landed
    }
    if (sizeof(R) == 4) CKFFT(cufftExecC2C(c.plan1d_x, (cufftComplex *) c.tbuf_a, (cufftComplex *) g, CUFFT_FORWARD));
    else CKFFT(cufftExecZ2Z(c.plan1d_x, (cufftDoubleComplex *) c.tbuf_a, (cufftDoubleComplex *) g, CUFFT_FORWARD));
    c.launches += 3;
    return;
  }
  const size_t tot = (size_t) nxb * N * NZ;
  k_pack_fwd<C><<<grid_for(tot, 256), 256, 0, c.stream>>>((const C *) g, (C *) c.tbuf_a, nxb, N, NZ, nyl);
  all_to_all(c, c.tbuf_a, c.tbuf_b, (size_t) nxb * nyl * NZ * sizeof(C));
  dim3 gr((N + 31) / 32, (NZ + 31) / 32, nyl), bl(32, 8);
  k_unpack_fwd<C><<<gr, bl, 0, c.stream>>>((const C *) c.tbuf_b, (C *) g, nxb, N, NZ, nyl);
  if (sizeof(R) == 4) CKFFT(cufftExecC2C(c.plan1d_x, (cufftComplex *) g, (cufftComplex *) g, CUFFT_FORWARD));
  else CKFFT(cufftExecZ2Z(c.plan1d_x, (cufftDoubleComplex *) g, (cufftDoubleComplex *) g, CUFFT_FORWARD));
  c.launches += 5;
}

template <typename R, typename C>
static void dist_c2r(Ctx &c, void *g) {
  const int N = c.N, NZ = c.NZ, nxb = c.nx, nyl = c.ny_loc;
  if (c.xf_on) {
    if (c.xf_dma) {
      xfft_bwd<C>(c, g, 0, c.stream);       // x-transform + pack into the local staging slot
      PhaseTimer t(c, PH_COMM);
      p2p_barrier(c);                       // every rank is done with its landing slot
      exchange_dma(c, c.stream, false, nullptr, 0);
      p2p_barrier(c);                       // every rank's blocks have landed
    } else {
      PhaseTimer t(c, PH_COMM);
      p2p_barrier(c);                       // every rank is done with its transpose buffer
      xfft_bwd<C>(c, g, 0, c.stream);
      p2p_barrier(c);                       // every rank's stores have landed
    }
    if (sizeof(R) == 4) CKFFT(cufftExecC2R(c.plan2d_c2r, (cufftComplex *) c.tbuf_a, (cufftReal *) g));
    else CKFFT(cufftExecZ2D(c.plan2d_c2r, (cufftDoubleComplex *) c.tbuf_a, (cufftDoubleReal *) g));
    c.launches += 2;
    return;
  }
  if (sizeof(R) == 4) CKFFT(cufftExecC2C(c.plan1d_x, (cufftComplex *) g, (cufftComplex *) g, CUFFT_INVERSE));
  else CKFFT(cufftExecZ2Z(c.plan1d_x, (cufftDoubleComplex *) g, (cufftDoubleComplex *) g, CUFFT_INVERSE));
  if (c.p2p) {
    PeerPtrs pp;
    for (int r = 0; r < 16; r++) pp.p[r] = c.peer_tbuf[r];
    {
      PhaseTimer t(c, PH_COMM);
      p2p_barrier(c);
      dim3 gr2((N + 31) / 32, (NZ + 31) / 32, nyl), bl2(32, 8);
      k_transpose_bwd_p2p<C><<<gr2, bl2, 0, c.stream>>>((const C *) g, pp, nxb, c.y0, N, NZ, nyl);
      p2p_barrier(c);
    }
    if (sizeof(R) == 4) CKFFT(cufftExecC2R(c.plan2d_c2r, (cufftComplex *) c.tbuf_a, (cufftReal *) g));
    else CKFFT(cufftExecZ2D(c.plan2d_c2r, (cufftDoubleComplex *) c.tbuf_a, (cufftDoubleReal *) g));
    c.launches += 3;
    return;
  }
  dim3 gr((N + 31) / 32, (NZ + 31) / 32, nyl), bl(32, 8);
  k_pack_bwd<C><<<gr, bl, 0, c.stream>>>((const C *) g, (C *) c.tbuf_a, nxb, N, NZ, nyl);
  all_to_all(c, c.tbuf_a, c.tbuf_b, (size_t) nxb * nyl * NZ * sizeof(C));
  const size_t tot = (size_t) nxb * N * NZ;
  k_unpack_bwd<C><<<grid_for(tot, 256), 256, 0, c.stream>>>((const C *) c.tbuf_b, (C *) g, nxb, N, NZ, nyl);
  if (sizeof(R) == 4) CKFFT(cufftExecC2R(c.plan2d_c2r, (cufftComplex *) g, (cufftReal *) g));
  else CKFFT(cufftExecZ2D(c.plan2d_c2r, (cufftDoubleComplex *) g, (cufftDoubleReal *) g));
  c.launches += 5;
}

// The three inverse transforms of a vector field (Forces, the displacement fields) as one pipeline: the x-FFTs run on
// the compute stream while the peer-memory transposes of the previous component run on the communication stream, and
// the 2-D c2r of a component starts as soon as all ranks have delivered it (one flag barrier per component):
//   compute: [x-FFT 0][x-FFT 1][x-FFT 2]          [2-D c2r 0][2-D c2r 1][2-D c2r 2]
//   comm   :          [B][T 0][B]   [T 1][B]   [T 2][B]
template <typename R, typename C>
static void dist_c2r3(Ctx &c, int block) {
  const int N = c.N, NZ = c.NZ, nxb = c.nx, nyl = c.ny_loc;
  const size_t gb = c.grid_bytes();
  cudaStream_t S = c.stream, T = c.comm_stream;
  if (c.xf_on && c.xf_dma) {
    // compute stream: the three x-transform + pack kernels back to back, then the 2-D c2r of each component as it lands;
    // communication stream: flag barrier, the copy-engine blocks of component a (while the SMs transform a + 1), barrier
    //   compute: [X 0][X 1][X 2]        [2-D 0][2-D 1][2-D 2]
    //   DMA    :      [B][copy 0][B][copy 1][B][copy 2][B]
    for (int a = 0; a < 3; a++) {
      xfft_bwd<C>(c, c.grid[block_grid(block, a)], a, S);
      CK(cudaEventRecord(c.ev_fft[a], S));
    }
    for (int a = 0; a < 3; a++) {
      CK(cudaStreamWaitEvent(T, c.ev_fft[a], 0));
      if (a == 0) p2p_barrier(c, T);        // every rank is done with all three landing slots
      exchange_dma(c, T, false, nullptr, a);
      p2p_barrier(c, T);                    // component a has landed everywhere
      CK(cudaEventRecord(c.ev_tr[a], T));
    }
    for (int a = 0; a < 3; a++) {
      CK(cudaStreamWaitEvent(S, c.ev_tr[a], 0));
      void *src = (char *) c.tbuf_a + (size_t) a * gb, *g = c.grid[block_grid(block, a)];
      if (sizeof(R) == 4) CKFFT(cufftExecC2R(c.plan2d_c2r, (cufftComplex *) src, (cufftReal *) g));
      else CKFFT(cufftExecZ2D(c.plan2d_c2r, (cufftDoubleComplex *) src, (cufftDoubleReal *) g));
    }
    c.launches += 6;
    return;
  }
  if (c.xf_on) {
    // fused x-transform + exchange of component a on the communication stream, the 2-D c2r of the components that
    // have landed on the compute stream:   comm: [B][X 0][B][X 1][B][X 2][B]     compute: [2-D 0][2-D 1][2-D 2]
    CK(cudaEventRecord(c.ev_fft[0], S));
    CK(cudaStreamWaitEvent(T, c.ev_fft[0], 0));
    p2p_barrier(c, T);                      // every rank is done with all three slots of its transpose buffer
    for (int a = 0; a < 3; a++) {
      xfft_bwd<C>(c, c.grid[block_grid(block, a)], a, T);
      p2p_barrier(c, T);                    // component a has landed everywhere
      CK(cudaEventRecord(c.ev_tr[a], T));
    }
    for (int a = 0; a < 3; a++) {
      CK(cudaStreamWaitEvent(S, c.ev_tr[a], 0));
      void *src = (char *) c.tbuf_a + (size_t) a * gb, *g = c.grid[block_grid(block, a)];
      if (sizeof(R) == 4) CKFFT(cufftExecC2R(c.plan2d_c2r, (cufftComplex *) src, (cufftReal *) g));
      else CKFFT(cufftExecZ2D(c.plan2d_c2r, (cufftDoubleComplex *) src, (cufftDoubleReal *) g));
    }
    c.launches += 6;
    return;
  }
  for (int a = 0; a < 3; a++) {
    void *g = c.grid[block_grid(block, a)];
    if (sizeof(R) == 4) CKFFT(cufftExecC2C(c.plan1d_x, (cufftComplex *) g, (cufftComplex *) g, CUFFT_INVERSE));
    else CKFFT(cufftExecZ2Z(c.plan1d_x, (cufftDoubleComplex *) g, (cufftDoubleComplex *) g, CUFFT_INVERSE));
    CK(cudaEventRecord(c.ev_fft[a], S));
  }
  dim3 gr((N + 31) / 32, (NZ + 31) / 32, nyl), bl(32, 8);
  for (int a = 0; a < 3; a++) {
    CK(cudaStreamWaitEvent(T, c.ev_fft[a], 0));
    if (a == 0) p2p_barrier(c, T);          // every rank is done with all three slots of its transpose buffer
    PeerPtrs pp;
    for (int r = 0; r < 16; r++) pp.p[r] = c.peer_tbuf[r] ? (char *) c.peer_tbuf[r] + (size_t) a * gb : nullptr;
    k_transpose_bwd_p2p<C><<<gr, bl, 0, T>>>((const C *) c.grid[block_grid(block, a)], pp, nxb, c.y0, N, NZ, nyl);
    p2p_barrier(c, T);                      // component a has landed everywhere
    CK(cudaEventRecord(c.ev_tr[a], T));
  }
  for (int a = 0; a < 3; a++) {
    CK(cudaStreamWaitEvent(S, c.ev_tr[a], 0));
    void *src = (char *) c.tbuf_a + (size_t) a * gb, *g = c.grid[block_grid(block, a)];
    if (sizeof(R) == 4) CKFFT(cufftExecC2R(c.plan2d_c2r, (cufftComplex *) src, (cufftReal *) g));
    else CKFFT(cufftExecZ2D(c.plan2d_c2r, (cufftDoubleComplex *) src, (cufftDoubleReal *) g));
  }
  c.launches += 9 + 3;
}

// Developer probe (mgp_debug_time_exchange): device time of ONE piece of the slab exchange, repeated `reps` times back
// to back between two flag barriers, without the transforms around it.  which: 0 flag barrier, 1 fused backward
// (x-transform + push), 2 fused forward (pull + x-transform), 3 backward transpose kernel, 4 forward transpose kernel,
// 5 cudaMemcpyAsync of the remote share of one slab from the next rank's buffer (copy-engine reference), 6 / 7 the
// backward / forward strided copy-engine blocks of the DMA exchange, 8 the batched inverse transform of the force grids
// (the whole pipeline), 9 one local 2-D c2r.
// Overwrites grid 1 / the transpose buffers: call it outside a step.
void fft_c2r_block(Ctx &c, int block);
void fft_debug_exchange(Ctx &c, int which, int reps, float *ms) {
  REQUIRE(c.slab && c.p2p, MGP_ERR_STATE, "exchange probe: needs the peer-memory slab path");
  REQUIRE(which == 0 || which >= 3 || c.xf_on, MGP_ERR_STATE, "exchange probe: fused x-transform is off");
  REQUIRE((which != 6 && which != 7) || c.xf_dma, MGP_ERR_STATE, "exchange probe: the staging slots exist only with MGP_XFFT_DMA=1");
  const int N = c.N, NZ = c.NZ, nxb = c.nx, nyl = c.ny_loc;
  void *g = c.grid[1];
  PeerPtrs pp;
  for (int r = 0; r < 16; r++) pp.p[r] = c.peer_tbuf[r];
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  p2p_barrier(c);
  CK(cudaEventRecord(e0, c.stream));
  for (int it = 0; it < reps; it++) {
    switch (which) {
      case 0: p2p_barrier(c); break;
      case 1: if (c.gbytes == 4) xfft_bwd<float2>(c, g, 0, c.stream); else xfft_bwd<double2>(c, g, 0, c.stream); break;
      case 2: if (c.gbytes == 4) xfft_fwd<float2>(c, g, c.stream); else xfft_fwd<double2>(c, g, c.stream); break;
      case 3: {
        dim3 gr((N + 31) / 32, (NZ + 31) / 32, nyl), bl(32, 8);
        if (c.gbytes == 4) k_transpose_bwd_p2p<float2><<<gr, bl, 0, c.stream>>>((const float2 *) g, pp, nxb, c.y0, N, NZ, nyl);
        else k_transpose_bwd_p2p<double2><<<gr, bl, 0, c.stream>>>((const double2 *) g, pp, nxb, c.y0, N, NZ, nyl);
        break;
      }
      case 4: {
        dim3 gr((nxb + 31) / 32, (NZ + 31) / 32, N), bl(32, 8);
        if (c.gbytes == 4) k_transpose_fwd_p2p<float2><<<gr, bl, 0, c.stream>>>((const float2 *) g, pp, nxb, c.x0, N, NZ, nyl);
        else k_transpose_fwd_p2p<double2><<<gr, bl, 0, c.stream>>>((const double2 *) g, pp, nxb, c.x0, N, NZ, nyl);
        break;
      }
      case 8: fft_c2r_block(c, 0); break;
      case 9:
        if (c.gbytes == 4) CKFFT(cufftExecC2R(c.plan2d_c2r, (cufftComplex *) c.tbuf_a, (cufftReal *) g));
        else CKFFT(cufftExecZ2D(c.plan2d_c2r, (cufftDoubleComplex *) c.tbuf_a, (cufftDoubleReal *) g));
        break;
      case 6: exchange_dma(c, c.stream, false, nullptr, 0); break;
      case 7: exchange_dma(c, c.stream, true, g, 0); break;
      default: {
        const size_t bytes = (size_t) nxb * N * NZ * (c.gbytes == 4 ? sizeof(float2) : sizeof(double2)) / c.P * (c.P - 1);
        if (bytes) CK(cudaMemcpyAsync(c.tbuf_b, c.peer_tbuf[c.right], bytes, cudaMemcpyDefault, c.stream));
        break;
      }
    }
  }
  CK(cudaEventRecord(e1, c.stream));
  p2p_barrier(c);
  CK(cudaStreamSynchronize(c.stream));
  CK(cudaGetLastError());
  CK(cudaEventElapsedTime(ms, e0, e1));
  *ms /= (float) (reps > 0 ? reps : 1);
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

// ------------------------------------------------------------------ public (module) entry points

void fft_r2c(Ctx &c, int gid) {
  PhaseTimer t(c, PH_FFT);
  void *g = c.grid[gid];
  if (c.slab) {
    if (c.gbytes == 4) dist_r2c<float, float2>(c, g); else dist_r2c<double, double2>(c, g);
    return;
  }
  if (c.gbytes == 4) CKFFT(cufftExecR2C(c.plan_r2c, (cufftReal *) g, (cufftComplex *) g));
  else CKFFT(cufftExecD2Z(c.plan_r2c, (cufftDoubleReal *) g, (cufftDoubleComplex *) g));
  c.launches += 3;
}

// single rank only: dst(k) = r2c(src(x)), src is left untouched
void fft_r2c_to(Ctx &c, int src, int dst) {
  PhaseTimer t(c, PH_FFT);
  REQUIRE(!c.slab, MGP_ERR_STATE, "out-of-place r2c is a single-rank path");
  if (c.gbytes == 4) CKFFT(cufftExecR2C(c.plan_r2c_oop, (cufftReal *) c.grid[src], (cufftComplex *) c.grid[dst]));
  else CKFFT(cufftExecD2Z(c.plan_r2c_oop, (cufftDoubleReal *) c.grid[src], (cufftDoubleComplex *) c.grid[dst]));
  c.launches += 3;
}

void fft_c2r(Ctx &c, int gid) {
  PhaseTimer t(c, PH_FFT);
  void *g = c.grid[gid];
  if (c.slab) {
    if (c.gbytes == 4) dist_c2r<float, float2>(c, g); else dist_c2r<double, double2>(c, g);
    return;
  }
  if (c.gbytes == 4) CKFFT(cufftExecC2R(c.plan_c2r, (cufftComplex *) g, (cufftReal *) g));
  else CKFFT(cufftExecZ2D(c.plan_c2r, (cufftDoubleComplex *) g, (cufftDoubleReal *) g));
  c.launches += 3;
}

void fft_c2r_forces(Ctx &c) { fft_c2r_block(c, 0); }

void fft_c2r_block(Ctx &c, int block) {
  REQUIRE(block == 0 || (block == 1 && c.aux_block), MGP_ERR_STATE, "batched c2r: block not allocated");
  if (c.slab && c.p2p) {
    PhaseTimer t(c, PH_FFT);
    if (c.gbytes == 4) dist_c2r3<float, float2>(c, block); else dist_c2r3<double, double2>(c, block);
    return;
  }
  if (c.slab) {
    for (int a = 0; a < 3; a++) fft_c2r(c, block_grid(block, a));
    return;
  }
  PhaseTimer t(c, PH_FFT);
  void *g = block == 0 ? c.force_block : c.aux_block;
  if (c.gbytes == 4) CKFFT(cufftExecC2R(c.plan_c2r3, (cufftComplex *) g, (cufftReal *) g));
  else CKFFT(cufftExecZ2D(c.plan_c2r3, (cufftDoubleComplex *) g, (cufftDoubleReal *) g));
  c.launches += 9;
}

// density[i] += ghost_from_left[i] + 1.0 over plane 0 (auxPM.c:354)
template <typename T>
__global__ void k_halo_add(T *__restrict__ plane0, const T *__restrict__ recv, size_t n) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    plane0[i] += (T) (recv[i] + (T) 1.0);
}

void halo_add_density(Ctx &c, int gid) {
  if (c.P == 1) return;    // single rank: the deposit wraps x itself
  PhaseTimer t(c, PH_COMM);
  char *g = (char *) c.grid[gid];
  const size_t pb = c.plane_bytes();
  CKNCCL(ncclGroupStart());
  CKNCCL(ncclSend(g + (size_t) c.nx * pb, pb, ncclChar, c.right, c.comm, c.stream));
  CKNCCL(ncclRecv(c.halo_recv, pb, ncclChar, c.left, c.comm, c.stream));
  CKNCCL(ncclGroupEnd());
  if (c.gbytes == 4) k_halo_add<float><<<grid_for(c.plane_vals, 256), 256, 0, c.stream>>>((float *) g, (const float *) c.halo_recv, c.plane_vals);
  else k_halo_add<double><<<grid_for(c.plane_vals, 256), 256, 0, c.stream>>>((double *) g, (const double *) c.halo_recv, c.plane_vals);
  c.launches++;
}

void halo_fill_forces(Ctx &c) { halo_fill_block(c, 0); }

void halo_fill_block(Ctx &c, int block) {
  const size_t pb = c.plane_bytes();
  if (c.P == 1) {
    for (int a = 0; a < 3; a++) {
      char *g = (char *) c.grid[block_grid(block, a)];
      CK(cudaMemcpyAsync(g + (size_t) c.nx * pb, g, pb, cudaMemcpyDeviceToDevice, c.stream));
    }
    return;
  }
  PhaseTimer t(c, PH_COMM);
  CKNCCL(ncclGroupStart());
  for (int a = 0; a < 3; a++) {
    char *g = (char *) c.grid[block_grid(block, a)];
    CKNCCL(ncclSend(g, pb, ncclChar, c.left, c.comm, c.stream));
    CKNCCL(ncclRecv(g + (size_t) c.nx * pb, pb, ncclChar, c.right, c.comm, c.stream));
  }
  CKNCCL(ncclGroupEnd());
}

void allreduce_sum(Ctx &c, double *dbuf, int n) {
  if (c.P == 1) return;
  CKNCCL(ncclAllReduce(dbuf, dbuf, (size_t) n, ncclDouble, ncclSum, c.comm, c.stream));
}

}  // namespace mgp
