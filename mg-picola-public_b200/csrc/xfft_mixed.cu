// Kernel instances and launchers of the mixed-radix fused x-transform (xfft_mixed.cuh) for the mesh sizes the weak-scaling
// runs use when Nmesh cannot be a power of two: 320 (2 GPUs), 400 (4 GPUs), and their doubles 640 / 800.  Opt-in
// (MGP_XFFT_MIXED=1, read by fft.cu); everything else about the slab path (peer mappings, barriers, pipelines) is fft.cu's.
#include "common.cuh"
#include "xfft_mixed.cuh"

namespace mgp {

#define XFM_DISPATCH_N(n, OP)                                                              \
  switch (n) {                                                                             \
    case 320: OP(320); break; case 400: OP(400); break; case 640: OP(640); break; case 800: OP(800); break; \
    default: throw mgp::Error(MGP_ERR_STATE, "mixed-radix x-transform: unsupported Nmesh"); \
  }

bool xfm_supported(int n) { return n == 320 || n == 400 || n == 640 || n == 800; }

template <typename C, int N>
static bool prepare(Ctx &c) {
  constexpr int TK = xfm::tile_lines(N, sizeof(C));
  static_assert(TK > 0 && xfm::supported(N), "tile does not fit");
  c.xf_tk = TK;
  c.xf_smem = (size_t) TK * N * sizeof(C);
  CK(cudaFuncSetAttribute(xfm::k_xfft_bwd_p2p<C, N, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.xf_smem));
  CK(cudaFuncSetAttribute(xfm::k_xfft_fwd_p2p<C, N, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.xf_smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, xfm::k_xfft_bwd_p2p<C, N, TK>, xf::kThreads, c.xf_smem));
  REQUIRE(occ >= 1, MGP_ERR_CUDA, "mixed-radix x-transform: kernel does not fit on an SM");
  int cps_want = c.P > 1 ? 1 : 2;                       // see xfft_prepare in fft.cu
  if (const char *cps = getenv("MGP_XFFT_CPS")) cps_want = atoi(cps);
  if (cps_want >= 1 && cps_want < occ) occ = cps_want;
  const char *tr = getenv("MGP_XFFT_TRIM");
  long long g = (long long) kSMs * occ - (tr ? atoi(tr) : 1);
  if (g < 1) g = 1;
  const long long ntiles = (long long) c.ny_loc * xf::tiles_per_line(c.NZ, TK, 128 / (int) sizeof(C));
  c.xf_grid = (int) (ntiles < g ? ntiles : g);
  // twiddle tables of the passes
  const int np = xfm::plan_npass(N), total = xfm::plan_twtotal(N);
  std::vector<C> tw(total);
  for (int i = 0; i < np; i++) {
    const int L = xfm::plan_R(N, i) * xfm::plan_M(N, i);
    for (int t = 0; t < L; t++) {
      const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double) t / (long double) L;
      tw[xfm::plan_twoff(N, i) + t] = xf::mk<C>((typename xf::RealOf<C>::type) cosl(a), (typename xf::RealOf<C>::type) sinl(a));
    }
  }
  CK(cudaMalloc(&c.xf_tw, (size_t) total * sizeof(C)));
  CK(cudaMemcpy(c.xf_tw, tw.data(), (size_t) total * sizeof(C), cudaMemcpyHostToDevice));
  return true;
}

bool xfm_prepare(Ctx &c) {
  if (!xfm_supported(c.N) || c.nx * c.P != c.N) return false;
  bool ok = false;
  if (c.gbytes == 4) {
#define OP(L) ok = prepare<float2, L>(c)
    XFM_DISPATCH_N(c.N, OP)
#undef OP
  } else {
#define OP(L) ok = prepare<double2, L>(c)
    XFM_DISPATCH_N(c.N, OP)
#undef OP
  }
  return ok;
}

template <typename C>
static void bwd_t(Ctx &c, const void *in, const PeerPtrs &pp, int y0, int NY, cudaStream_t st) {
#define OP(L)                                                                                       \
  xfm::k_xfft_bwd_p2p<C, L, xfm::tile_lines(L, sizeof(C))><<<c.xf_grid, xf::kThreads, c.xf_smem, st>>>( \
      (const C *) in, pp, (const C *) c.xf_tw, c.nx, y0, NY, c.NZ, c.ny_loc)
  XFM_DISPATCH_N(c.N, OP)
#undef OP
}
template <typename C>
static void fwd_t(Ctx &c, void *out, const PeerPtrs &pp, int y0, int NY, cudaStream_t st) {
#define OP(L)                                                                                       \
  xfm::k_xfft_fwd_p2p<C, L, xfm::tile_lines(L, sizeof(C))><<<c.xf_grid, xf::kThreads, c.xf_smem, st>>>( \
      pp, (C *) out, (const C *) c.xf_tw, c.nx, y0, NY, c.NZ, c.ny_loc)
  XFM_DISPATCH_N(c.N, OP)
#undef OP
}

void xfm_bwd(Ctx &c, const void *in, const PeerPtrs &pp, int y0, int NY, cudaStream_t st) {
  if (c.gbytes == 4) bwd_t<float2>(c, in, pp, y0, NY, st); else bwd_t<double2>(c, in, pp, y0, NY, st);
  CK(cudaGetLastError());
}
void xfm_fwd(Ctx &c, void *out, const PeerPtrs &pp, int y0, int NY, cudaStream_t st) {
  if (c.gbytes == 4) fwd_t<float2>(c, out, pp, y0, NY, st); else fwd_t<double2>(c, out, pp, y0, NY, st);
  CK(cudaGetLastError());
}

}  // namespace mgp
