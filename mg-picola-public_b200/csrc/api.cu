// extern "C" surface (include/mgpicola.h) and the per-step orchestration that mirrors
// GetDisplacements (auxPM.c:37-103).
#include "common.cuh"

#include <cmath>
#include <mutex>

namespace mgp {

static thread_local std::string g_last_error;
void set_last_error(const std::string &m) { g_last_error = m; }

static const char *kPhaseNames[PH_COUNT] = {"MoveParticles", "PtoMesh", "FFT", "ComputeFifthForce", "Forces",
                                            "MtoParticles", "Kick", "Drift", "Pofk", "Sort", "Comm", "SDField", "SDAssign"};

static bool needs_mg_arrays(const Ctx &c) { return c.cfg.model == MGP_MODEL_FOFR || c.cfg.model == MGP_MODEL_DGP; }

static void destroy(Ctx *cp);

static Ctx *create(const mgp_config *cfg) {
  REQUIRE(cfg != nullptr, MGP_ERR_INVALID, "mgp_create: cfg is NULL");
  REQUIRE(cfg->nmesh >= 2 && cfg->nmesh % 2 == 0, MGP_ERR_INVALID, "mgp_create: Nmesh must be even and >= 2");
  REQUIRE(cfg->nsample >= 1, MGP_ERR_INVALID, "mgp_create: Nsample must be >= 1");
  REQUIRE(cfg->box > 0, MGP_ERR_INVALID, "mgp_create: Box must be > 0");
  REQUIRE(cfg->grid_bytes == 4 || cfg->grid_bytes == 8, MGP_ERR_INVALID, "mgp_create: grid_bytes must be 4 or 8");
  REQUIRE(cfg->nranks >= 1 && cfg->rank >= 0 && cfg->rank < cfg->nranks, MGP_ERR_INVALID, "mgp_create: bad rank/nranks");
  REQUIRE(cfg->model >= MGP_MODEL_NONE && cfg->model <= MGP_MODEL_GEFF, MGP_ERR_INVALID, "mgp_create: unknown model");
  REQUIRE(cfg->deposit_mode >= 0 && cfg->deposit_mode <= 3, MGP_ERR_INVALID, "mgp_create: unknown deposit_mode");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  REQUIRE(e == cudaSuccess && ndev > 0, MGP_ERR_CUDA,
          "mgp_create: no CUDA device (this library has no CPU fallback)");
  REQUIRE(cfg->device >= 0 && cfg->device < ndev, MGP_ERR_INVALID, "mgp_create: bad device ordinal");
  CK(cudaSetDevice(cfg->device));

  Ctx *cp = new Ctx();
  Ctx &c = *cp;
  try {
    c.cfg = *cfg;
    c.cfg.nccl_unique_id = nullptr;
    if (const char *dm = getenv("MGP_DEPOSIT_MODE")) {            // developer / test knob: every context uses this strategy
      const int m = atoi(dm);
      if (m >= 0 && m <= 3) c.cfg.deposit_mode = m;
    }
    c.N = cfg->nmesh; c.NZ = cfg->nmesh / 2 + 1;
    c.P = cfg->nranks; c.rank = cfg->rank;
    c.gbytes = cfg->grid_bytes;
    {
      const char *fs = getenv("MGP_FORCE_SLAB");     // developer / test knob: the multi-rank transform path on one rank
      c.slab = c.P > 1 || (fs && atoi(fs) != 0);
    }
    if (c.P > 1) {
      REQUIRE(c.N % c.P == 0, MGP_ERR_INVALID, "mgp_create: nranks must divide Nmesh");
      REQUIRE(cfg->nccl_unique_id != nullptr, MGP_ERR_INVALID, "mgp_create: nccl_unique_id required when nranks > 1");
    }
    // slab layout == fftw_mpi_local_size_3d default block distribution (2LPT.c:50)
    const int block = (c.N + c.P - 1) / c.P;
    c.x0 = c.rank * block;
    c.nx = c.x0 >= c.N ? 0 : (c.N - c.x0 < block ? c.N - c.x0 : block);
    REQUIRE(c.nx >= 1, MGP_ERR_INVALID, "mgp_create: rank owns no mesh planes");
    c.ny_loc = c.N / c.P; c.y0 = c.rank * c.ny_loc;
    c.left = (c.rank + c.P - 1) % c.P; c.right = (c.rank + 1) % c.P;
    // particle planes (initialize_parts, 2LPT.c:118-176)
    c.npl = 0; c.p0 = cfg->nsample;
    for (int i = 0; i < cfg->nsample; i++) {
      const int slab = (int) ((double) ((long long) i * c.N) / (double) cfg->nsample);
      if (slab >= c.x0 && slab < c.x0 + c.nx) { c.npl++; if (i < c.p0) c.p0 = i; }
    }
    const double np0 = (double) c.npl * cfg->nsample * (double) cfg->nsample;
    const double buf = (c.P == 1) ? 1.0 : (cfg->buffer >= 1.0 ? cfg->buffer : 1.0);
    c.cap = (uint64_t) ceil(np0 * buf);
    if (c.cap < 32) c.cap = 32;
    REQUIRE(c.cap < ((uint64_t) 1 << 32), MGP_ERR_INVALID, "mgp_create: more than 2^32 particles on one rank: use more ranks");
    c.plane_vals = (size_t) c.N * 2 * c.NZ;
    c.grid_vals = (size_t) (c.nx + 1) * c.plane_vals;

    CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    for (int d = 0; d < 4; d++) { CK(cudaEventCreate(&c.ev[d][0])); CK(cudaEventCreate(&c.ev[d][1])); }

    CK(cudaMalloc(&c.force_block, 3 * c.grid_bytes()));
    for (int a = 0; a < 3; a++) c.grid[1 + a] = (char *) c.force_block + (size_t) a * c.grid_bytes();
    if (cfg->scale_dependent) {
      // density, mgarray_one, mgarray_two contiguous: the second target of the batched inverse transform (sd.cu)
      CK(cudaMalloc(&c.aux_block, 3 * c.grid_bytes()));
      c.grid[0] = c.aux_block;
      c.grid[4] = (char *) c.aux_block + c.grid_bytes();
      c.grid[5] = (char *) c.aux_block + 2 * c.grid_bytes();
    } else {
      CK(cudaMalloc(&c.grid[0], c.grid_bytes()));
      if (needs_mg_arrays(c)) {
        CK(cudaMalloc(&c.grid[4], c.grid_bytes()));
        CK(cudaMalloc(&c.grid[5], c.grid_bytes()));
      }
    }
    if (cfg->scale_dependent) {
      CK(cudaMalloc(&c.sd_delta[0], c.grid_bytes()));
      CK(cudaMalloc(&c.sd_delta[1], c.grid_bytes()));
    }
    CK(cudaMalloc(&c.halo_recv, c.plane_bytes()));
    particles_alloc(c);
    if (cfg->scale_dependent) sd_alloc(c);
    if (c.P > 1) {
      ncclUniqueId id;
      memcpy(&id, cfg->nccl_unique_id, sizeof(id));
      CKNCCL(ncclCommInitRank(&c.comm, c.P, id, c.rank));
    }
    fft_setup(c);
    CK(cudaStreamSynchronize(c.stream));
  } catch (...) {
    destroy(cp);   // frees whatever the half-built context holds (every free below tolerates a null handle)
    throw;
  }
  return cp;
}

static void destroy(Ctx *cp) {
  if (!cp) return;
  Ctx &c = *cp;
  cudaSetDevice(c.cfg.device);
  if (c.stream) cudaStreamSynchronize(c.stream);
  fft_teardown(c);
  particles_free(c);
  sd_free(c);
  if (c.aux_block) cudaFree(c.aux_block);
  else { cudaFree(c.grid[0]); cudaFree(c.grid[4]); cudaFree(c.grid[5]); }
  cudaFree(c.force_block); cudaFree(c.halo_recv);
  cudaFree(c.sd_delta[0]); cudaFree(c.sd_delta[1]);
  cudaFree(c.mig_dev); if (c.mig_host) cudaFreeHost(c.mig_host);
  cudaFree(c.pofk_bins_d); cudaFree(c.pofk_sinc_d); cudaFree(c.pofk_out_d); cudaFree(c.nu_tab_d);
  if (c.pofk_out_h) cudaFreeHost(c.pofk_out_h);
  if (c.comm) ncclCommDestroy(c.comm);
  for (int d = 0; d < 4; d++) { if (c.ev[d][0]) cudaEventDestroy(c.ev[d][0]); if (c.ev[d][1]) cudaEventDestroy(c.ev[d][1]); }
  if (c.stream) cudaStreamDestroy(c.stream);
  cudaGetLastError();      // a teardown of a half-built context may have touched null handles
  delete cp;
}

// ---- per-step path -------------------------------------------------------------------------

// Particle order policy.  sort_particles = 0: never sort; k >= 1: re-sort by mesh cell once at least k
// drifts have happened since the last sort.  The TILE and DETERMINISTIC deposits need the exact cell
// order and therefore sort whenever the order is stale; the ATOMIC deposit and the gather only need
// spatial locality, which the Lagrangian / last-sorted order keeps for several steps.
static void ensure_order(Ctx &c) {
  if (!c.cfg.sort_particles || c.sorted) return;
  const bool need_exact = c.cfg.deposit_mode == MGP_DEPOSIT_TILE || c.cfg.deposit_mode == MGP_DEPOSIT_DETERMINISTIC;
  if (need_exact || c.drifts_since_sort >= c.cfg.sort_particles) particles_sort(c);
}

static void move_particles(Ctx &c) {
  if (c.cfg.scale_dependent) sd_drop(c);
  {
    PhaseTimer t(c, PH_MOVE);
    if (c.P > 1) particles_migrate(c);
  }
  if (c.np_after_sort != SIZE_MAX) particles_sort(c);   // closes the holes the leavers left
  else ensure_order(c);
}

static void ptomesh(Ctx &c, const mgp_step_scalars *s) {
  if (c.cfg.scale_dependent) sd_drop(c);
  ensure_order(c);
  if (c.cfg.deposit_mode == MGP_DEPOSIT_ROWS) rows_bin(c);   // the per-step row bins (timed as Sort, not as PtoMesh)
  if (needs_mg_arrays(c) && !c.slab) {
    // CopyDensityArray (mg.h:381) without the copy: deposit straight into mgarray_two and transform
    // out of place into P3D, which leaves delta(x) in mgarray_two exactly as the reference has it
    { PhaseTimer t(c, PH_PTOMESH); deposit_density(c, MGP_GRID_MG_TWO); }
    fft_r2c_to(c, MGP_GRID_MG_TWO, MGP_GRID_DENSITY);
  } else {
    {
      PhaseTimer t(c, PH_PTOMESH);
      deposit_density(c, MGP_GRID_DENSITY);
      halo_add_density(c, MGP_GRID_DENSITY);
      if (needs_mg_arrays(c)) real_copy(c, MGP_GRID_MG_TWO, MGP_GRID_DENSITY, 1.0);   // CopyDensityArray (mg.h:381)
    }
    fft_r2c(c, MGP_GRID_DENSITY);
  }
  c.density_live = true;
  c.step_pofk_valid = false;
  c.step_pofk_tot_valid = false;
  if (s && s->compute_pofk) {
    const int nb = pofk_effective_nbins(c);
    c.step_pofk.assign(nb, 0.0); c.step_kmean.assign(nb, 0.0); c.step_nmodes.assign(nb, 0.0);
    pofk_bin(c, MGP_GRID_DENSITY, c.step_pofk.data(), c.step_kmean.data(), c.step_nmodes.data());
    c.step_pofk_valid = true;
  }
  if (s && s->nu_by_k2) {                                   // MASSIVE_NEUTRINOS (auxPM.c:383-420)
    { PhaseTimer t(c, PH_PTOMESH); kspace_nu_add(c, s->nu_by_k2, s->n_nu, s->nu_cdmfac); }
    if (s->compute_pofk) {                                  // "total" P(k) (auxPM.c:424-427)
      const int nb = pofk_effective_nbins(c);
      c.step_pofk_tot.assign(nb, 0.0); c.step_kmean_tot.assign(nb, 0.0); c.step_nmodes_tot.assign(nb, 0.0);
      pofk_bin(c, MGP_GRID_DENSITY, c.step_pofk_tot.data(), c.step_kmean_tot.data(), c.step_nmodes_tot.data());
      c.step_pofk_tot_valid = true;
    }
  }
}

static void rsd_power_spectrum(Ctx &c, double vnorm, double dDdy, double dD2dy, double *out_y, double *out_z) {
  REQUIRE(out_y && out_z, MGP_ERR_INVALID, "mgp_compute_rsd_power_spectrum: NULL output");
  if (c.cfg.scale_dependent) { sd_evict_block(c, 0); sd_materialise(c, 1); }   // work grid = force grid X; V needs P.dDdy
  for (int axis = 1; axis <= 2; axis++) {                   // YAXIS then ZAXIS (compute_pofk.c:437-448)
    deposit_rsd(c, MGP_GRID_FORCE_X, axis, vnorm, dDdy, dD2dy);
    halo_add_density(c, MGP_GRID_FORCE_X);
    fft_r2c(c, MGP_GRID_FORCE_X);
    pofk_bin_rsd(c, MGP_GRID_FORCE_X, axis == 1 ? out_y : out_z);
  }
}

static void compute_fifth_force(Ctx &c, const mgp_step_scalars *s) {
  if (c.cfg.model == MGP_MODEL_NONE) return;
  REQUIRE(s != nullptr, MGP_ERR_INVALID, "mgp_compute_fifth_force: step scalars are NULL");
  if (c.cfg.model == MGP_MODEL_GEFF) {
    PhaseTimer t(c, PH_FIFTH);
    kspace_scale(c, MGP_GRID_DENSITY, s->geff);                 // mg.h:127-137
    return;
  }
  if (c.cfg.model == MGP_MODEL_FOFR) {                          // mg.h:147-189
    if (!c.cfg.include_screening) {
      PhaseTimer t(c, PH_FIFTH);
      kspace_phi_of_k(c, MGP_GRID_DENSITY, s->coupling, s->massterm2);
      return;
    }
    const double n3 = pow((double) c.N, 3);
    double normfactor = 1.0 / n3;                               // mg.h:27-28
    normfactor *= 1.5 * c.cfg.omega / s->a * pow(c.cfg.box / 2997.92458 / (2.0 * 3.14159265358979323846), 2);
    { PhaseTimer t(c, PH_FIFTH); kspace_divide_laplacian(c, normfactor); }
    fft_c2r(c, MGP_GRID_MG_ONE);
    { PhaseTimer t(c, PH_FIFTH); real_screen_potential(c, s->phi_crit, true); }
    fft_r2c(c, MGP_GRID_MG_TWO);
    { PhaseTimer t(c, PH_FIFTH); kspace_phi_of_k(c, MGP_GRID_MG_TWO, s->coupling, s->massterm2); }
    return;
  }
  // DGP (mg.h:197-260)
  if (!c.cfg.include_screening) {
    PhaseTimer t(c, PH_FIFTH);
    kspace_scale_to(c, MGP_GRID_DENSITY, MGP_GRID_MG_TWO, s->coupling);   // density aliases P3D (mg.h:214-216)
    return;
  }
  { PhaseTimer t(c, PH_FIFTH); kspace_smooth(c, s->rsmooth); }
  fft_c2r(c, MGP_GRID_MG_ONE);
  { PhaseTimer t(c, PH_FIFTH); real_screen_density(c, s->coupling, s->dgp_fac0, nullptr); }
  fft_r2c(c, MGP_GRID_MG_TWO);
}

static void forces(Ctx &c) {
  { PhaseTimer t(c, PH_FORCES); kspace_forces(c, needs_mg_arrays(c)); }
  c.density_live = false;                  // P3D and mgarray_two are consumed
  fft_c2r_forces(c);
  halo_fill_forces(c);
  c.forces_live = true;
}

}  // namespace mgp

using namespace mgp;

#define API_BEGIN try {
#define API_END                                         \
  }                                                     \
  catch (const mgp::Error &e) {                         \
    mgp::set_last_error(e.what());                      \
    return e.code;                                      \
  }                                                     \
  catch (const std::exception &e) {                     \
    mgp::set_last_error(e.what());                      \
    return MGP_ERR_CUDA;                                \
  }                                                     \
  return MGP_OK;

#define CTX(ctx)                                                         \
  REQUIRE(ctx != nullptr, MGP_ERR_INVALID, "context is NULL");          \
  Ctx &c = *reinterpret_cast<Ctx *>(ctx);                                \
  CK(cudaSetDevice(c.cfg.device));

static void *grid_ptr(Ctx &c, int id) {
  if (id >= 0 && id < 6) return c.grid[id];
  if (id == MGP_GRID_SD_DELTA1) return c.sd_delta[0];
  if (id == MGP_GRID_SD_DELTA2) return c.sd_delta[1];
  return nullptr;
}

extern "C" {

const char *mgp_last_error(void) { return g_last_error.c_str(); }
int mgp_version(void) { return 200; }

int mgp_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int mgp_nccl_unique_id(void *out128) {
  API_BEGIN
  REQUIRE(out128 != nullptr, MGP_ERR_INVALID, "mgp_nccl_unique_id: NULL");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  CKNCCL(ncclGetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  API_END
}

int mgp_create(const mgp_config *cfg, mgp_ctx **out) {
  API_BEGIN
  REQUIRE(out != nullptr, MGP_ERR_INVALID, "mgp_create: out is NULL");
  *out = reinterpret_cast<mgp_ctx *>(create(cfg));
  API_END
}

int mgp_destroy(mgp_ctx *ctx) {
  API_BEGIN
  destroy(reinterpret_cast<Ctx *>(ctx));
  API_END
}

int mgp_get_layout(mgp_ctx *ctx, int *local_nx, int *local_x_start, int *local_np, int *local_p_start, uint64_t *numpart) {
  API_BEGIN
  CTX(ctx);
  if (local_nx) *local_nx = c.nx;
  if (local_x_start) *local_x_start = c.x0;
  if (local_np) *local_np = c.npl;
  if (local_p_start) *local_p_start = c.p0;
  if (numpart) *numpart = c.np;
  API_END
}

int mgp_kspace_layout(mgp_ctx *ctx, int *transposed, int *ky_start, int *ky_local) {
  API_BEGIN
  CTX(ctx);
  if (transposed) *transposed = c.slab ? 1 : 0;
  if (ky_start) *ky_start = c.slab ? c.y0 : 0;
  if (ky_local) *ky_local = c.slab ? c.ny_loc : c.N;
  API_END
}

int mgp_upload_particles(mgp_ctx *ctx, uint64_t n, const float *pos, const float *vel, const float *D, const float *D2,
                         const uint64_t *id) {
  API_BEGIN
  CTX(ctx);
  particles_upload(c, n, pos, vel, D, D2, id);
  if (c.cfg.scale_dependent) {            // P.D / P.D2 live in the per-step field arrays (sd.cu)
    if (D) sd_copy_field(c, 0, const_cast<float *>(D), false);
    if (D2) sd_copy_field(c, 1, const_cast<float *>(D2), false);
  }
  API_END
}

int mgp_download_particles(mgp_ctx *ctx, float *pos, float *vel, float *D, float *D2, uint64_t *id) {
  API_BEGIN
  CTX(ctx);
  if (c.cfg.scale_dependent) {
    particles_download(c, pos, vel, nullptr, nullptr, id);
    if (D) sd_copy_field(c, 0, D, true);
    if (D2) sd_copy_field(c, 1, D2, true);
  } else {
    particles_download(c, pos, vel, D, D2, id);
  }
  API_END
}

int mgp_pack_snapshot(mgp_ctx *ctx, double lengthfac, double velfac_times_fac, const double sumxyz[3], double dDdy, double dD2dy,
                      float *pos, float *vel, uint64_t *id) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(pos && vel && id && sumxyz, MGP_ERR_INVALID, "mgp_pack_snapshot: NULL argument");
  if (c.cfg.scale_dependent && c.cfg.use_cola) {
    sd_materialise(c, 1);                                 // a resident merged field -> P.dDdy / P.dD2dy
    REQUIRE(c.sd_set[2], MGP_ERR_STATE, "mgp_pack_snapshot (scale-dependent): assign FIELD_dDdy first (main.c:824-825)");
  }
  particles_snapshot(c, lengthfac, velfac_times_fac, sumxyz, dDdy, dD2dy, pos, vel, id);
  API_END
}

void *mgp_alloc_host(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

void mgp_free_host(void *p) {
  if (p) cudaFreeHost(p);
}

int mgp_download_disp(mgp_ctx *ctx, float *disp) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(disp != nullptr, MGP_ERR_INVALID, "mgp_download_disp: NULL");
  REQUIRE(c.have_disp, MGP_ERR_STATE, "mgp_download_disp: no displacements available");
  copy_soa3(c, c.disp, c.np, disp, true);
  API_END
}

int mgp_ic_generate(mgp_ctx *ctx, const mgp_ic_config *ic) {
  API_BEGIN
  CTX(ctx);
  if (c.cfg.scale_dependent) sd_drop(c);
  c.density_live = false; c.forces_live = false;
  ic_generate(c, ic);
  API_END
}

int mgp_ic_particles_begin(mgp_ctx *ctx) {
  API_BEGIN
  CTX(ctx);
  if (c.cfg.scale_dependent) sd_drop(c);
  c.density_live = false; c.forces_live = false;
  ic_particles_begin(c);
  API_END
}

int mgp_ic_particles_add(mgp_ctx *ctx, const float *pos01, uint64_t n, uint64_t *taken_total) {
  API_BEGIN
  CTX(ctx);
  ic_particles_add(c, pos01, n);
  if (taken_total) *taken_total = c.ic_ext_taken;
  API_END
}

int mgp_ic_particles_finish(mgp_ctx *ctx, double normfac, const double *rescale_by_k2, size_t n) {
  API_BEGIN
  CTX(ctx);
  ic_particles_finish(c, normfac, rescale_by_k2, n);
  API_END
}

int mgp_ic_download(mgp_ctx *ctx, float *za, float *lpt) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(c.ic_ready, MGP_ERR_STATE, "mgp_ic_download: call mgp_ic_generate first");
  const size_t n = (size_t) c.npl * c.cfg.nsample * c.cfg.nsample;
  for (int f = 0; f < 2; f++) {
    float *dst = f == 0 ? za : lpt;
    if (!dst) continue;
    float *src = f == 0 ? c.disp : (float *) c.pA2;
    copy_soa3(c, src, n, dst, true, &c.ic_means[3 * f]);       // ZA -= sumdis (2LPT.c:1501-1508)
  }
  API_END
}

int mgp_init_particles(mgp_ctx *ctx, double Di, double Di2, double dDdy, double dD2dy) {
  API_BEGIN
  CTX(ctx);
  if (c.cfg.scale_dependent) { sd_materialise(c, 0); sd_materialise(c, 1); sd_init_particles(c); }      // main.c:296-300: the fields already carry their growth factors
  else ic_init_particles(c, Di, Di2, dDdy, dD2dy);
  API_END
}

int mgp_seedtable(unsigned seed, int nmesh, unsigned *out) {
  API_BEGIN
  REQUIRE(out != nullptr && nmesh >= 2 && nmesh % 2 == 0, MGP_ERR_INVALID, "mgp_seedtable: bad arguments");
  ic_seedtable(seed, nmesh, out);
  API_END
}

double mgp_ranlxd1_draw(unsigned long seed, long n) { return ic_ranlxd1_draw(seed, n); }

int mgp_upload_disp(mgp_ctx *ctx, const float *disp) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(disp != nullptr, MGP_ERR_INVALID, "mgp_upload_disp: NULL");
  copy_soa3(c, c.disp, c.np, const_cast<float *>(disp), false);
  c.have_disp = true;
  API_END
}

int mgp_move_particles(mgp_ctx *ctx) {
  API_BEGIN
  CTX(ctx);
  move_particles(c);
  CK(cudaStreamSynchronize(c.stream));
  API_END
}

int mgp_ptomesh(mgp_ctx *ctx, const mgp_step_scalars *s) {
  API_BEGIN
  CTX(ctx);
  ptomesh(c, s);
  CK(cudaStreamSynchronize(c.stream));
  API_END
}

int mgp_compute_fifth_force(mgp_ctx *ctx, const mgp_step_scalars *s) {
  API_BEGIN
  CTX(ctx);
  compute_fifth_force(c, s);
  CK(cudaStreamSynchronize(c.stream));
  API_END
}

int mgp_forces(mgp_ctx *ctx) {
  API_BEGIN
  CTX(ctx);
  forces(c);
  CK(cudaStreamSynchronize(c.stream));
  API_END
}

int mgp_mtoparticles(mgp_ctx *ctx, double sumDxyz[3]) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(sumDxyz != nullptr, MGP_ERR_INVALID, "mgp_mtoparticles: sumDxyz is NULL");
  gather_forces(c, sumDxyz);
  c.forces_live = false;
  API_END
}

int mgp_get_displacements(mgp_ctx *ctx, const mgp_step_scalars *s, double sumDxyz[3]) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(sumDxyz != nullptr, MGP_ERR_INVALID, "mgp_get_displacements: sumDxyz is NULL");
  move_particles(c);
  ptomesh(c, s);
  compute_fifth_force(c, s);
  forces(c);
  gather_forces(c, sumDxyz);
  c.forces_live = false;
  API_END
}

int mgp_kick(mgp_ctx *ctx, double A, double dda, double ddDddy, double ddD2ddy, const double sumDxyz[3], double sumxyz[3]) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(sumDxyz && sumxyz, MGP_ERR_INVALID, "mgp_kick: NULL argument");
  if (c.cfg.scale_dependent) {
    PhaseTimer t(c, PH_KICK);
    REQUIRE(c.have_disp, MGP_ERR_STATE, "mgp_kick: no displacements; call mgp_get_displacements first");
    sd_kick(c, A, dda, sumDxyz, sumxyz);
  } else {
    particles_kick(c, A, dda, ddDddy, ddD2ddy, sumDxyz, sumxyz);
  }
  API_END
}

int mgp_drift(mgp_ctx *ctx, double dyyy, double deltaD, double deltaD2, const double sumxyz[3]) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(sumxyz != nullptr, MGP_ERR_INVALID, "mgp_drift: NULL argument");
  if (c.cfg.scale_dependent) {
    { PhaseTimer t(c, PH_DRIFT); sd_drift(c, dyyy, sumxyz); }
    particles_after_drift(c);
  } else {
    particles_drift(c, dyyy, deltaD, deltaD2, sumxyz);
  }
  CK(cudaStreamSynchronize(c.stream));
  API_END
}

int mgp_fof_find(mgp_ctx *ctx, const mgp_fof_config *cfg, uint64_t *n_halos) {
  API_BEGIN
  CTX(ctx);
  fof_find(c, cfg);
  if (n_halos) *n_halos = c.fof_halos.size();
  API_END
}

int mgp_fof_get(mgp_ctx *ctx, mgp_fof_halo *out) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(c.fof_valid, MGP_ERR_STATE, "mgp_fof_get: call mgp_fof_find first");
  REQUIRE(out != nullptr || c.fof_halos.empty(), MGP_ERR_INVALID, "mgp_fof_get: NULL");
  if (!c.fof_halos.empty()) memcpy(out, c.fof_halos.data(), c.fof_halos.size() * sizeof(mgp_fof_halo));
  API_END
}

int mgp_lightcone_count(mgp_ctx *ctx, const mgp_lightcone_step *ls, uint64_t *count) {
  API_BEGIN
  CTX(ctx);
  lightcone_count(c, ls, count);
  API_END
}

int mgp_drift_lightcone(mgp_ctx *ctx, const mgp_lightcone_step *ls, uint64_t cap, float *block, uint64_t *count) {
  API_BEGIN
  CTX(ctx);
  lightcone_drift(c, ls, cap, block, count);
  API_END
}

int mgp_assign_displacement_field(mgp_ctx *ctx, int fieldtype, int lpt_order, const double *growth_by_k2, size_t n) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(lpt_order == 1 || lpt_order == 2, MGP_ERR_INVALID, "mgp_assign_displacement_field: LPT order must be 1 or 2");
  sd_assign(c, fieldtype, lpt_order, growth_by_k2, nullptr, n);
  CK(cudaStreamSynchronize(c.stream));
  API_END
}

int mgp_assign_displacement_fields_merged(mgp_ctx *ctx, int fieldtype, const double *g1, const double *g2, size_t n) {
  API_BEGIN
  CTX(ctx);
  sd_assign(c, fieldtype, 0, g1, g2, n);
  CK(cudaStreamSynchronize(c.stream));
  API_END
}

int mgp_download_sd_fields(mgp_ctx *ctx, float *dDdy, float *dD2dy) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(c.cfg.scale_dependent, MGP_ERR_STATE, "mgp_download_sd_fields: scale_dependent = 0");
  if (dDdy) sd_copy_field(c, 2, dDdy, true);
  if (dD2dy) sd_copy_field(c, 3, dD2dy, true);
  API_END
}

int mgp_upload_sd_fields(mgp_ctx *ctx, const float *dDdy, const float *dD2dy) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(c.cfg.scale_dependent, MGP_ERR_STATE, "mgp_upload_sd_fields: scale_dependent = 0");
  if (dDdy) sd_copy_field(c, 2, const_cast<float *>(dDdy), false);
  if (dD2dy) sd_copy_field(c, 3, const_cast<float *>(dD2dy), false);
  API_END
}

int mgp_set_pofk_config(mgp_ctx *ctx, const mgp_pofk_config *pc) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(pc != nullptr, MGP_ERR_INVALID, "mgp_set_pofk_config: NULL");
  c.pofk = *pc;
  c.pofk_set = true;
  API_END
}

int mgp_pofk_nbins(mgp_ctx *ctx) {
  if (!ctx) return MGP_ERR_INVALID;
  Ctx &c = *reinterpret_cast<Ctx *>(ctx);
  if (!c.pofk_set) return MGP_ERR_STATE;
  return pofk_effective_nbins(c);
}

int mgp_compute_power_spectrum(mgp_ctx *ctx, double *pofk, double *kmean, double *nmodes) {
  API_BEGIN
  CTX(ctx);
  pofk_bin(c, MGP_GRID_DENSITY, pofk, kmean, nmodes);
  API_END
}

int mgp_get_step_power_spectrum(mgp_ctx *ctx, double *pofk, double *kmean, double *nmodes) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(c.step_pofk_valid, MGP_ERR_STATE, "no in-step P(k) available (set mgp_step_scalars.compute_pofk)");
  const size_t nb = c.step_pofk.size();
  if (pofk) memcpy(pofk, c.step_pofk.data(), nb * sizeof(double));
  if (kmean) memcpy(kmean, c.step_kmean.data(), nb * sizeof(double));
  if (nmodes) memcpy(nmodes, c.step_nmodes.data(), nb * sizeof(double));
  API_END
}

int mgp_get_step_power_spectrum_total(mgp_ctx *ctx, double *pofk, double *kmean, double *nmodes) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(c.step_pofk_tot_valid, MGP_ERR_STATE, "no total-matter P(k) available (set nu_by_k2 and compute_pofk)");
  const size_t nb = c.step_pofk_tot.size();
  if (pofk) memcpy(pofk, c.step_pofk_tot.data(), nb * sizeof(double));
  if (kmean) memcpy(kmean, c.step_kmean_tot.data(), nb * sizeof(double));
  if (nmodes) memcpy(nmodes, c.step_nmodes_tot.data(), nb * sizeof(double));
  API_END
}

int mgp_compute_rsd_power_spectrum(mgp_ctx *ctx, double vnorm, double dDdy, double dD2dy, double *out_y, double *out_z) {
  API_BEGIN
  CTX(ctx);
  rsd_power_spectrum(c, vnorm, dDdy, dD2dy, out_y, out_z);
  API_END
}

int mgp_simple_pofk(mgp_ctx *ctx, int scheme, int subtract_shotnoise, int tsc_as_published, double *pofk, double *nmodes) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(pofk && nmodes, MGP_ERR_INVALID, "mgp_simple_pofk: NULL output");
  simple_pofk(c, scheme, subtract_shotnoise, tsc_as_published, pofk, nmodes);
  API_END
}

size_t mgp_grid_local_values(mgp_ctx *ctx) {
  if (!ctx) return 0;
  return reinterpret_cast<Ctx *>(ctx)->grid_vals;
}

int mgp_download_grid(mgp_ctx *ctx, int grid_id, void *host) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(grid_ptr(c, grid_id), MGP_ERR_INVALID, "mgp_download_grid: grid not allocated");
  CK(cudaMemcpyAsync(host, grid_ptr(c, grid_id), c.grid_bytes(), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  API_END
}

int mgp_upload_grid(mgp_ctx *ctx, int grid_id, const void *host) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(grid_ptr(c, grid_id), MGP_ERR_INVALID, "mgp_upload_grid: grid not allocated");
  CK(cudaMemcpyAsync(grid_ptr(c, grid_id), host, c.grid_bytes(), cudaMemcpyHostToDevice, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  if (grid_id == MGP_GRID_SD_DELTA1 || grid_id == MGP_GRID_SD_DELTA2) c.sd_have_delta = true;
  API_END
}

int mgp_fft_r2c(mgp_ctx *ctx, int grid_id) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(grid_id >= 0 && grid_id < 6 && c.grid[grid_id], MGP_ERR_INVALID, "mgp_fft_r2c: grid not allocated");
  fft_r2c(c, grid_id);
  CK(cudaStreamSynchronize(c.stream));
  API_END
}

int mgp_debug_time_exchange(mgp_ctx *ctx, int which, int reps, float *ms) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(ms != nullptr && reps >= 1, MGP_ERR_INVALID, "mgp_debug_time_exchange: bad arguments");
  fft_debug_exchange(c, which, reps, ms);
  API_END
}

int mgp_fft_c2r(mgp_ctx *ctx, int grid_id) {
  API_BEGIN
  CTX(ctx);
  REQUIRE(grid_id >= 0 && grid_id < 6 && c.grid[grid_id], MGP_ERR_INVALID, "mgp_fft_c2r: grid not allocated");
  fft_c2r(c, grid_id);
  CK(cudaStreamSynchronize(c.stream));
  API_END
}

uint64_t mgp_launch_count(mgp_ctx *ctx, int reset) {
  if (!ctx) return 0;
  Ctx &c = *reinterpret_cast<Ctx *>(ctx);
  const uint64_t v = c.launches;
  if (reset) c.launches = 0;
  return v;
}

int mgp_phase_count(void) { return PH_COUNT; }
const char *mgp_phase_name(int i) { return (i >= 0 && i < PH_COUNT) ? kPhaseNames[i] : ""; }

int mgp_phase_times_ms(mgp_ctx *ctx, double *ms, uint64_t *calls, int reset) {
  API_BEGIN
  CTX(ctx);
  for (int i = 0; i < PH_COUNT; i++) {
    if (ms) ms[i] = c.phase_ms[i];
    if (calls) calls[i] = c.phase_calls[i];
    if (reset) { c.phase_ms[i] = 0; c.phase_calls[i] = 0; }
  }
  API_END
}

int mgp_set_phase_timing(mgp_ctx *ctx, int on) {
  API_BEGIN
  CTX(ctx);
  c.phase_timing = on != 0;
  API_END
}

void *mgp_stream(mgp_ctx *ctx) { return ctx ? (void *) reinterpret_cast<Ctx *>(ctx)->stream : nullptr; }

}  // extern "C"
