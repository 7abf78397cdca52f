// Shared declarations of the B200-native COLA particle-mesh library (internal; the public
// surface is include/mgpicola.h).  Everything here is sm_100a-only: no fallbacks.
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mgpicola.h"
#include "xfft.cuh"

namespace mgp {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &m);

#define MGP_STR2(x) #x
#define MGP_STR(x) MGP_STR2(x)
#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      throw mgp::Error(MGP_ERR_CUDA, std::string(__FILE__ ":" MGP_STR(__LINE__) " ") +            \
                                         cudaGetErrorString(e_));                                 \
  } while (0)
#define CKFFT(call)                                                                               \
  do {                                                                                            \
    cufftResult r_ = (call);                                                                      \
    if (r_ != CUFFT_SUCCESS)                                                                      \
      throw mgp::Error(MGP_ERR_CUDA, std::string(__FILE__ ":" MGP_STR(__LINE__) " cuFFT error ") + \
                                         std::to_string((int) r_));                               \
  } while (0)
#define CKNCCL(call)                                                                              \
  do {                                                                                            \
    ncclResult_t r_ = (call);                                                                     \
    if (r_ != ncclSuccess)                                                                        \
      throw mgp::Error(MGP_ERR_CUDA, std::string(__FILE__ ":" MGP_STR(__LINE__) " NCCL: ") +      \
                                         ncclGetErrorString(r_));                                 \
  } while (0)
#define REQUIRE(cond, code, msg)                                                                  \
  do {                                                                                            \
    if (!(cond)) throw mgp::Error(code, msg);                                                     \
  } while (0)

// phases, named after the reference's timer.h sub-categories
enum Phase { PH_MOVE = 0, PH_PTOMESH, PH_FFT, PH_FIFTH, PH_FORCES, PH_MTOP, PH_KICK, PH_DRIFT, PH_POFK, PH_SORT, PH_COMM, PH_SDFIELD, PH_SDASSIGN, PH_COUNT };

constexpr int kSMs = 148;   // B200

struct Ctx {
  mgp_config cfg{};
  int N = 0, NZ = 0;          // Nmesh, Nmesh/2+1
  int P = 1, rank = 0;
  bool slab = false;          // slab-decomposed transforms + transposed k-space: P > 1 (or MGP_FORCE_SLAB=1 on one rank)
  int nx = 0, x0 = 0;         // Local_nx, Local_x_start
  int npl = 0, p0 = 0;        // Local_np, Local_p_start
  int left = 0, right = 0;    // LeftTask / RightTask
  int gbytes = 8;
  size_t plane_vals = 0;      // N * 2*NZ reals per x-plane
  size_t grid_vals = 0;       // (nx+1) * plane_vals
  void *grid[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void *force_block = nullptr;   // FX,FY,FZ contiguous (batched c2r)
  void *halo_send = nullptr, *halo_recv = nullptr;   // 3 planes each

  // FFT
  cufftHandle plan_r2c = 0, plan_c2r = 0, plan_c2r3 = 0, plan_r2c_oop = 0;
  bool have_plans = false;
  // distributed FFT (P > 1)
  cufftHandle plan2d_r2c = 0, plan2d_c2r = 0, plan1d_x = 0;
  int ny_loc = 0, y0 = 0;     // k-space y-slab of this rank
  void *tbuf_a = nullptr, *tbuf_b = nullptr;   // transpose pack / unpack buffers
  // peer-memory transposes (P > 1): every rank maps every other rank's tbuf_a and flag array (cudaIpc) and the
  // transpose kernel stores straight into the owner's buffer over NVLink; barriers are flag exchanges in peer memory
  bool p2p = false;
  void *peer_tbuf[16] = {nullptr};
  uint32_t *sync_flags = nullptr;              // [P] epoch counters, written by the peers
  uint32_t *peer_flags[16] = {nullptr};
  uint32_t sync_epoch = 0;
  cudaStream_t comm_stream = nullptr;          // high-priority stream the transposes of a batched transform run on
  cudaEvent_t ev_fft[3] = {nullptr, nullptr, nullptr}, ev_tr[3] = {nullptr, nullptr, nullptr};
  void *fft_work = nullptr;                    // cuFFT work area shared by all plans
  // x-transform fused with the slab exchange (xfft.cuh): power-of-two Nmesh on the peer-memory path
  bool xf_on = false;
  int xf_lgn = 0, xf_lgnxb = 0;                // log2 Nmesh, log2 Local_nx
  bool xf_wide = false;                        // wide-tile instances (xfft_wide.cu; default from Nmesh = 1024 on several ranks)
  bool xf_mixed = false;                       // Nmesh = 2^a 3^b 5^c instance (xfft_mixed.cu; opt-in MGP_XFFT_MIXED=1)
  void *xf_tw = nullptr;                       // twiddle tables of the passes (complex, grid precision)
  int xf_tk = 0, xf_grid = 0;                  // lines per tile, persistent grid size
  size_t xf_smem = 0;
  cufftHandle plan2d_r2c_oop = 0;              // 2-D r2c from a grid into the transpose buffer (the pull source)
  // exchange engine of the fused path: 1 = the x-transform kernel packs / unpacks a local staging buffer (slots 3-5 of
  // tbuf_a, laid out [rank][x_local][ky_local][kz]) and the copy engines move one strided block per peer over NVLink
  // (cudaMemcpy2DAsync on the peer mappings); 0 = the kernel's own loads / stores go to peer memory
  int xf_dma = 0;
  cudaStream_t cp_stream[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_cp_go = nullptr, ev_cp_done[4] = {nullptr, nullptr, nullptr, nullptr};

  // particles (SoA of 16-byte records; see DESIGN.md "data layout")
  uint64_t np = 0, cap = 0;
  float4 *pA = nullptr;       // pos.xyz , id low  32 bits
  float4 *pB = nullptr;       // vel.xyz , id high 32 bits
  float4 *pC = nullptr;       // D.xyz   , D2.x
  float4 *pE = nullptr;       // (float2 storage) D2.y, D2.z
  float4 *pA2 = nullptr, *pB2 = nullptr, *pC2 = nullptr, *pE2 = nullptr;   // sort destination set (swapped in)
  float *disp = nullptr;      // [3][cap]
  bool have_disp = false;
  bool sorted = false;        // particle order == cell order of current positions
  bool exact_cell_order = false;   // sorted by cell (radix) rather than by bucket
  uint32_t *bucket_start = nullptr;  // (nbuckets + 1) counts -> offsets
  size_t nbuckets = 0;
  int bucket_zshift = 0;
  int drifts_since_sort = 1 << 30;
  size_t np_after_sort = SIZE_MAX;   // set by MoveParticles: live count once the next sort has dropped the leavers
  unsigned *mig_dev = nullptr, *mig_host = nullptr;
  unsigned long long last_moved = 0; // particles exchanged (all ranks) by the last MoveParticles   // forces a sort at the first opportunity after an upload
  uint32_t *key[2] = {nullptr, nullptr};
  uint32_t *perm[2] = {nullptr, nullptr};
  uint32_t *row_start = nullptr;   // (nx*N + 1) offsets into the sorted particle list
  // per-step bins (rows.cu): particle indices grouped by the (x-plane, y-row, z-chunk) of their cell; the records stay where they are
  uint32_t *bin_perm = nullptr;    // [cap]
  uint32_t *bin_start = nullptr;   // [nbins + 2] offsets into bin_perm; bin = (x-plane, y-row, z-chunk)
  int bin_zc = 0, bin_nc = 1;      // cells per z-chunk, chunks per row
  size_t nbins = 0;
  void *bin_scan_temp = nullptr;
  size_t bin_scan_bytes = 0;
  bool bins_valid = false;         // bins describe the current positions and storage order
  void *cub_temp = nullptr;
  size_t cub_temp_bytes = 0;
  int key_zshift = 0;         // cell key drops this many low z bits (0 = full cell sort)
  int key_bits = 32;

  // initial conditions / scale-dependent growth
  double ic_means[6] = {0, 0, 0, 0, 0, 0};
  bool ic_ready = false;
  bool ic_ext_open = false;                  // between mgp_ic_particles_begin and _finish (READICFROMFILE)
  unsigned long long ic_ext_taken = 0;       // external particles that fell into this rank's slab
  void *sd_delta[2] = {nullptr, nullptr};    // delta1_k, delta2_k (cdelta_cdm, cdelta_cdm2; vars.h:272-273)
  bool sd_have_delta = false;
  float *sdf[4] = {nullptr, nullptr, nullptr, nullptr};   // per-particle D, D2, dDdy, dD2dy as [3][cap] (sd.cu)
  bool sd_zero[4] = {false, false, false, false};         // slot reads as 0 (merged mode)
  bool sd_set[4] = {false, false, false, false};          // slot assigned for the current particle order
  bool sd_lagrangian_only = false;                        // particles carry IDs only (before mgp_init_particles)
  double *sd_gtab[2] = {nullptr, nullptr};                // growth tables over |d|^2 on the device
  // merged fields stay in the grids they were transformed in and are read by Kick / Drift directly:
  // sd_res[pair] = block (0: force grids 1-3, 1: grids 0, 4, 5) holding the field of pair 0 (D + D2) / 1 (dDdy + dD2dy)
  void *aux_block = nullptr;                              // grids 0, 4, 5 as one allocation (second batched c2r target)
  int sd_res[2] = {-1, -1};
  bool density_live = false;                              // grid 0 (and 4, 5) hold this step's density / MG fields (PtoMesh .. Forces)
  bool forces_live = false;                               // grids 1-3 hold the force components (Forces .. MtoParticles)
  double sd_res_mean[2][3] = {{0, 0, 0}, {0, 0, 0}};
  // lattice points owned by other ranks (P > 1): request / response lists, built once per particle order
  bool sd_req_valid = false;
  unsigned *sd_cnt_dev = nullptr, *sd_cnt_host = nullptr;
  std::vector<unsigned> sd_need, sd_serve;                // per rank: what I ask for / what I answer
  size_t sd_nneed = 0, sd_nserve = 0;
  unsigned long long *sd_req_id = nullptr, *sd_srv_id = nullptr;
  uint32_t *sd_req_slot = nullptr;
  float *sd_resp_out = nullptr, *sd_resp_in = nullptr;
  size_t sd_req_id_bytes = 0, sd_req_slot_bytes = 0, sd_srv_id_bytes = 0, sd_resp_out_bytes = 0, sd_resp_in_bytes = 0;

  void *stage = nullptr;      // host <-> device particle staging (two chunks)
  size_t stage_bytes = 0;
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;

  double *d_red = nullptr;    // device reduction scratch (doubles)
  double *h_red = nullptr;    // pinned host mirror
  size_t red_cap = 0;
  int *d_flag = nullptr, *h_flag = nullptr;

  // P(k)
  mgp_pofk_config pofk{};
  bool pofk_set = false;
  std::vector<double> step_pofk, step_kmean, step_nmodes;
  bool step_pofk_valid = false;
  std::vector<double> step_pofk_tot, step_kmean_tot, step_nmodes_tot;   // "total" P(k) after the neutrino add
  bool step_pofk_tot_valid = false;
  double *nu_tab_d = nullptr;
  int *pofk_bins_d = nullptr;                 // bin index of every integer |d|^2
  double *pofk_sinc_d = nullptr, *pofk_out_d = nullptr, *pofk_out_h = nullptr;
  bool pofk_tables_valid = false;
  int pofk_tab_nbins = 0, pofk_tab_bintype = 0;
  double pofk_tab_kmin = 0, pofk_tab_kmax = 0;

  // FoF halos of the last mgp_fof_find (fof.cu), by decreasing np
  std::vector<mgp_fof_halo> fof_halos;
  bool fof_valid = false;

  cudaStream_t stream = nullptr;
  ncclComm_t comm = nullptr;

  // instrumentation
  uint64_t launches = 0;
  bool phase_timing = false;
  cudaEvent_t ev[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  int ev_depth = 0;           // phase timers may nest (FFT contains its transposes' Comm)
  double phase_ms[PH_COUNT] = {0};
  uint64_t phase_calls[PH_COUNT] = {0};

  size_t plane_bytes() const { return plane_vals * (size_t) gbytes; }
  size_t grid_bytes() const { return grid_vals * (size_t) gbytes; }
};

struct PhaseTimer {
  Ctx &c; int ph; bool on; int d;
  PhaseTimer(Ctx &c_, int ph_) : c(c_), ph(ph_), on(c_.phase_timing && c_.ev_depth < 4), d(0) {
    if (on) { d = c.ev_depth++; cudaEventRecord(c.ev[d][0], c.stream); }
  }
  ~PhaseTimer() {
    if (on) {
      cudaEventRecord(c.ev[d][1], c.stream);
      cudaEventSynchronize(c.ev[d][1]);
      float ms = 0; cudaEventElapsedTime(&ms, c.ev[d][0], c.ev[d][1]);
      c.phase_ms[ph] += ms; c.phase_calls[ph]++;
      c.ev_depth--;
    }
  }
};

inline unsigned grid_for(size_t n, int block, int max_waves = 32) {
  size_t g = (n + block - 1) / block;
  size_t cap = (size_t) kSMs * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned) g;
}

// ---- module entry points (one .cu each) ----
// particles.cu
void particles_alloc(Ctx &c);
void particles_free(Ctx &c);
void particles_upload(Ctx &c, uint64_t n, const float *pos, const float *vel, const float *D, const float *D2, const uint64_t *id);
void particles_download(Ctx &c, float *pos, float *vel, float *D, float *D2, uint64_t *id);
void particles_sort(Ctx &c);
void particles_snapshot(Ctx &c, double lengthfac, double vfac, const double sumxyz[3], double dDdy, double dD2dy, float *pos,
                        float *vel, uint64_t *id);
void copy_soa3(Ctx &c, float *dev_soa, size_t n, float *host_aos, bool to_host, const double *sub_mean = nullptr);
void particles_kick(Ctx &c, double A, double dda, double ddD, double ddD2, const double sumD[3], double sumV[3]);
void particles_drift(Ctx &c, double dyyy, double dD, double dD2, const double sumV[3]);
void particles_migrate(Ctx &c);
void particles_after_drift(Ctx &c);
// deposit.cu
void deposit_density(Ctx &c, int grid_id);
void gather_forces(Ctx &c, double sumD[3]);
void deposit_rsd(Ctx &c, int grid_id, int axis, double vnorm, double dDdy, double dD2dy);
// rows.cu
void rows_alloc(Ctx &c);
void rows_free(Ctx &c);
void rows_bin(Ctx &c);
void deposit_rows(Ctx &c, int grid_id);
bool deposit_rows_supported(const Ctx &c);
bool gather_rows_supported(const Ctx &c);
unsigned gather_rows(Ctx &c);       // returns the number of per-block partial sums left in c.d_red
// fft.cu
void fft_setup(Ctx &c);
void fft_teardown(Ctx &c);
void fft_r2c(Ctx &c, int grid_id);
void fft_c2r(Ctx &c, int grid_id);
void fft_r2c_to(Ctx &c, int src_grid, int dst_grid);
void fft_c2r_forces(Ctx &c);
void fft_c2r_block(Ctx &c, int block);      // block 0: grids 1, 2, 3; block 1: grids 0, 4, 5 (scale_dependent only)
void halo_fill_block(Ctx &c, int block);
void fft_debug_exchange(Ctx &c, int which, int reps, float *ms);
// xfft_wide.cu
bool xfw_prepare(Ctx &c);
void xfw_bwd(Ctx &c, const void *in, const PeerPtrs &pp, int y0, int NY, cudaStream_t st);
void xfw_fwd(Ctx &c, void *out, const PeerPtrs &pp, int y0, int NY, cudaStream_t st);
// xfft_mixed.cu
bool xfm_supported(int n);
bool xfm_prepare(Ctx &c);
void xfm_bwd(Ctx &c, const void *in, const PeerPtrs &pp, int y0, int NY, cudaStream_t st);
void xfm_fwd(Ctx &c, void *out, const PeerPtrs &pp, int y0, int NY, cudaStream_t st);
inline int block_grid(int block, int a) { return block == 0 ? 1 + a : (a == 0 ? 0 : 3 + a); }
void halo_add_density(Ctx &c, int grid_id);
void halo_fill_forces(Ctx &c);
// kspace.cu
void kspace_forces(Ctx &c, bool add_mg);
void kspace_divide_laplacian(Ctx &c, double normfactor);
void kspace_phi_of_k(Ctx &c, int src_grid, double coupling, double massterm2);
void kspace_smooth(Ctx &c, double rsmooth);
void kspace_scale(Ctx &c, int grid_id, double f);
void kspace_scale_to(Ctx &c, int src, int dst, double f);
void real_copy(Ctx &c, int dst_grid, int src_grid, double scale);
void real_screen_potential(Ctx &c, double phi_crit, bool screening);
void real_screen_density(Ctx &c, double coupling, double fac0, double stats[3]);
void pofk_bin(Ctx &c, int grid_id, double *pofk, double *kmean, double *nmodes);
void pofk_bin_rsd(Ctx &c, int grid_id, double *out5);
void kspace_nu_add(Ctx &c, const double *nufac_host, size_t n, double cdmfac);
int pofk_effective_nbins(const Ctx &c);

// simplepofk.cu
void simple_pofk(Ctx &c, int scheme, int subtract_shotnoise, int slip, double *pofk, double *nmodes);

// fof.cu
void fof_find(Ctx &c, const mgp_fof_config *cfg);
// lightcone.cu
void lightcone_count(Ctx &c, const mgp_lightcone_step *ls, uint64_t *count);
void lightcone_drift(Ctx &c, const mgp_lightcone_step *ls, uint64_t cap, float *block, uint64_t *count);

// ic.cu
void ic_generate(Ctx &c, const mgp_ic_config *ic);
void ic_init_particles(Ctx &c, double Di, double Di2, double dDdy, double dD2dy);
void ic_particles_begin(Ctx &c);
void ic_particles_add(Ctx &c, const float *pos01, uint64_t n);
void ic_particles_finish(Ctx &c, double normfac, const double *rescale_by_k2, size_t n);
void ic_seedtable(unsigned seed, int N, unsigned *out);
double ic_ranlxd1_draw(unsigned long seed, long n);

// sd.cu (scale-dependent growth)
void sd_alloc(Ctx &c);
void sd_free(Ctx &c);
void sd_assign(Ctx &c, int fieldtype, int order, const double *g1, const double *g2, size_t n);
void sd_init_particles(Ctx &c);
void sd_kick(Ctx &c, double A, double dda, const double sumD[3], double sumV[3]);
void sd_drift(Ctx &c, double dyyy, const double sumV[3]);
void sd_copy_field(Ctx &c, int slot, float *host, bool to_host);
void sd_materialise(Ctx &c, int pair);      // resident merged field -> the per-particle arrays
void sd_drop(Ctx &c);                        // a new force evaluation starts: the four fields are dead
void sd_evict_block(Ctx &c, int block);      // about to overwrite a grid block: save a live resident field first

// reductions: sum `n` doubles' worth of per-block partials living in c.d_red into host values
void reduce_alloc(Ctx &c, size_t n);

}  // namespace mgp
