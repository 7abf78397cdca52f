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

template <typename C, int TK>
static void xfft_prepare(Ctx &c) {
  CK(cudaFuncSetAttribute(xf::k_xfft_bwd_p2p<C, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.xf_smem));
  CK(cudaFuncSetAttribute(xf::k_xfft_fwd_p2p<C, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.xf_smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, xf::k_xfft_bwd_p2p<C, TK>, xf::kThreads, c.xf_smem));
  REQUIRE(occ >= 1, MGP_ERR_CUDA, "fused x-transform: kernel does not fit on an SM");
  const long long ntiles = (long long) c.ny_loc * ((c.NZ + TK - 1) / TK);
  long long g = (long long) kSMs * occ;
  c.xf_grid = (int) (ntiles < g ? ntiles : g);
}

// Chooses the tile (TK lines of N complex values, at most 64 KB so that two CTAs share an SM and one tile's global
// traffic overlaps the other's butterflies) and uploads the twiddle tables.  Leaves xf_on = false when Nmesh is not
// a power of two or the peer-memory path is off: those cases keep cuFFT's 1-D plan + the transpose kernels.
static void xfft_setup(Ctx &c) {
  c.xf_on = false;
  const char *env = getenv("MGP_XFFT");
  if (!c.p2p || (env && atoi(env) == 0)) return;
  if (!xf::make_plan(c.N, c.xf_plan)) return;
  const size_t cb = c.gbytes == 4 ? sizeof(float2) : sizeof(double2);
  int tk = 16;
  while (tk > 4 && (size_t) tk * c.N * cb > 64 * 1024) tk >>= 1;
  if (const char *e = getenv("MGP_XFFT_TK")) { const int v = atoi(e); if (v == 4 || v == 8 || v == 16) tk = v; }
  if ((size_t) tk * c.N * cb > 200 * 1024) return;
  c.xf_tk = tk;
  c.xf_smem = (size_t) tk * c.N * cb;
  const xf::Plan &pl = c.xf_plan;
  std::vector<double2> tw(pl.twtotal);
  for (int i = 0; i < pl.npass; i++) {
    const int L = pl.R[i] << pl.lgM[i];
    for (int t = 0; t < L; t++) {
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) t / (long double) L;
      tw[pl.twoff[i] + t] = make_double2((double) cosl(a), (double) sinl(a));
    }
  }
  CK(cudaMalloc(&c.xf_tw, (size_t) pl.twtotal * cb));
  if (c.gbytes == 4) {
    std::vector<float2> twf(pl.twtotal);
    for (int t = 0; t < pl.twtotal; t++) twf[t] = make_float2((float) tw[t].x, (float) tw[t].y);
    CK(cudaMemcpy(c.xf_tw, twf.data(), (size_t) pl.twtotal * cb, cudaMemcpyHostToDevice));
    if (tk == 4) xfft_prepare<float2, 4>(c); else if (tk == 8) xfft_prepare<float2, 8>(c); else xfft_prepare<float2, 16>(c);
  } else {
    CK(cudaMemcpy(c.xf_tw, tw.data(), (size_t) pl.twtotal * cb, cudaMemcpyHostToDevice));
    if (tk == 4) xfft_prepare<double2, 4>(c); else if (tk == 8) xfft_prepare<double2, 8>(c); else xfft_prepare<double2, 16>(c);
  }
  c.xf_on = true;
}

// backward x-transform of the local transposed k-space `in`, every x stored into slot `slot` of its owner's buffer
template <typename C>
static void xfft_bwd(Ctx &c, const void *in, int slot, cudaStream_t st) {
  PeerPtrs pp;
  for (int r = 0; r < 16; r++) pp.p[r] = c.peer_tbuf[r] ? (char *) c.peer_tbuf[r] + (size_t) slot * c.grid_bytes() : nullptr;
#define XF_LAUNCH(TK)                                                                                              \
  xf::k_xfft_bwd_p2p<C, TK><<<c.xf_grid, xf::kThreads, c.xf_smem, st>>>((const C *) in, pp, c.xf_plan, (const C *) c.xf_tw, \
                                                                        c.nx, c.y0, c.N, c.NZ, c.ny_loc)
  if (c.xf_tk == 4) XF_LAUNCH(4); else if (c.xf_tk == 8) XF_LAUNCH(8); else XF_LAUNCH(16);
#undef XF_LAUNCH
  CK(cudaGetLastError());
  c.launches++;
}

// forward x-transform: pulls slot 0 of every owner's buffer, writes the local transposed k-space `out`
template <typename C>
static void xfft_fwd(Ctx &c, void *out, cudaStream_t st) {
  PeerPtrs pp;
  for (int r = 0; r < 16; r++) pp.p[r] = c.peer_tbuf[r];
#define XF_LAUNCH(TK)                                                                                              \
  xf::k_xfft_fwd_p2p<C, TK><<<c.xf_grid, xf::kThreads, c.xf_smem, st>>>(pp, (C *) out, c.xf_plan, (const C *) c.xf_tw, c.nx, \
                                                                        c.y0, c.N, c.NZ, c.ny_loc)
  if (c.xf_tk == 4) XF_LAUNCH(4); else if (c.xf_tk == 8) XF_LAUNCH(8); else XF_LAUNCH(16);
#undef XF_LAUNCH
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
    CK(cudaMalloc(&c.tbuf_a, 3 * c.grid_bytes()));     // three slots: the batched c2r pipelines its three transposes
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
    {
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
