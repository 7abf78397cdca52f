// Wide-tile instances of the fused x-transform (xfft.cuh): tiles of up to 128 KB (one CTA per SM, which is how the kernel
// runs on several ranks anyway), i.e. twice the run length of the default instances on the exchange side.  Default from
// Nmesh = 1024 on several ranks (fft.cu; MGP_XFFT_WIDE overrides); parity: tests/test_slab_fused.py, host emulation
// tests/host/xfft_emul.cu.
#include "common.cuh"

namespace mgp {

#define XFW_DISPATCH_LGN(lgn, OP)                                                                            \
  switch (lgn) {                                                                                             \
    case 7: OP(7); break; case 8: OP(8); break; case 9: OP(9); break; case 10: OP(10); break; case 11: OP(11); break; \
    default: throw mgp::Error(MGP_ERR_STATE, "wide-tile x-transform: unsupported Nmesh");                    \
  }

bool xfw_supported(int lgn) { return lgn >= 7 && lgn <= 11; }

template <typename C, int LGN>
static bool prepare(Ctx &c) {
  constexpr int TK = xf::tile_lines_wide(LGN, sizeof(C));
  if constexpr (TK == 0) {
    return false;
  } else {
    c.xf_tk = TK;
    c.xf_smem = ((size_t) TK << LGN) * sizeof(C);
    CK(cudaFuncSetAttribute(xf::k_xfft_bwd_p2p<C, LGN, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.xf_smem));
    CK(cudaFuncSetAttribute(xf::k_xfft_fwd_p2p<C, LGN, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.xf_smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, xf::k_xfft_bwd_p2p<C, LGN, TK>, xf::kThreads, c.xf_smem));
    REQUIRE(occ >= 1, MGP_ERR_CUDA, "wide-tile x-transform: kernel does not fit on an SM");
    int cps_want = c.P > 1 ? 1 : 2;
    if (const char *cps = getenv("MGP_XFFT_CPS")) cps_want = atoi(cps);
    if (cps_want >= 1 && cps_want < occ) occ = cps_want;
    // on several ranks the kernel is NVLink-bound: a quarter of the SMs left entirely to the 2-D cuFFT kernels of the
    // other components of a batched transform costs the exchange 5 % and gains the pipeline 7 % (8 x B200, 1024^3:
    // 8.08 -> 7.52 ms per batch of three, profiles/r02_exchange_8gpu.md)
    const char *tr = getenv("MGP_XFFT_TRIM");
    long long g = (long long) kSMs * occ - (tr ? atoi(tr) : (c.P > 1 ? 38 : 1));
    if (g < 1) g = 1;
    const long long ntiles = (long long) c.ny_loc * xf::tiles_per_line(c.NZ, TK, 128 / (int) sizeof(C));
    c.xf_grid = (int) (ntiles < g ? ntiles : g);
    return true;
  }
}

// replaces the tile geometry xfft_prepare chose (the twiddle tables are the same)
bool xfw_prepare(Ctx &c) {
  if (!xfw_supported(c.xf_lgn)) return false;
  bool ok = false;
  if (c.gbytes == 4) {
#define OP(L) ok = prepare<float2, L>(c)
    XFW_DISPATCH_LGN(c.xf_lgn, OP)
#undef OP
  } else {
#define OP(L) ok = prepare<double2, L>(c)
    XFW_DISPATCH_LGN(c.xf_lgn, OP)
#undef OP
  }
  return ok;
}

template <typename C>
static void bwd_t(Ctx &c, const void *in, const PeerPtrs &pp, int y0, int NY, cudaStream_t st) {
#define OP(L)                                                                                                      \
  xf::k_xfft_bwd_p2p<C, L, xf::tile_lines_wide(L, sizeof(C)) ? xf::tile_lines_wide(L, sizeof(C)) : 4>               \
      <<<c.xf_grid, xf::kThreads, c.xf_smem, st>>>((const C *) in, pp, (const C *) c.xf_tw, c.xf_lgnxb, y0, NY, c.NZ, c.ny_loc)
  XFW_DISPATCH_LGN(c.xf_lgn, OP)
#undef OP
}
template <typename C>
static void fwd_t(Ctx &c, void *out, const PeerPtrs &pp, int y0, int NY, cudaStream_t st) {
#define OP(L)                                                                                                      \
  xf::k_xfft_fwd_p2p<C, L, xf::tile_lines_wide(L, sizeof(C)) ? xf::tile_lines_wide(L, sizeof(C)) : 4>               \
      <<<c.xf_grid, xf::kThreads, c.xf_smem, st>>>(pp, (C *) out, (const C *) c.xf_tw, c.xf_lgnxb, y0, NY, c.NZ, c.ny_loc)
  XFW_DISPATCH_LGN(c.xf_lgn, OP)
#undef OP
}

void xfw_bwd(Ctx &c, const void *in, const PeerPtrs &pp, int y0, int NY, cudaStream_t st) {
  if (c.gbytes == 4) bwd_t<float2>(c, in, pp, y0, NY, st); else bwd_t<double2>(c, in, pp, y0, NY, st);
  CK(cudaGetLastError());
}
void xfw_fwd(Ctx &c, void *out, const PeerPtrs &pp, int y0, int NY, cudaStream_t st) {
  if (c.gbytes == 4) fwd_t<float2>(c, out, pp, y0, NY, st); else fwd_t<double2>(c, out, pp, y0, NY, st);
  CK(cudaGetLastError());
}

}  // namespace mgp
