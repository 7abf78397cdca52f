/*
 * mgpicola.h -- C ABI of the B200-native COLA particle-mesh force path.
 *
 * This is the drop-in boundary for MG-PICOLA's C driver.  The reference has no plugin
 * mechanism: its "operator API" is a set of argument-less C functions that talk through the
 * globals of src/vars.h.  Every entry point below names the reference function (file:line
 * under the reference's src/) whose work it takes over; adapter/auxPM_cuda.c keeps the
 * reference names and fills these calls from the reference globals.
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success or a negative mgp_status, and
 *    mgp_last_error() gives the text (the driver maps non-zero to FatalError, auxPM.c:668).
 *  - the library owns all device memory (particles, grids, cuFFT plans, NCCL communicator).
 *    Host pointers passed in are read during the call only.
 *  - every call is synchronous with respect to its host-visible outputs.
 *  - there is no CPU fallback: without a CUDA device mgp_create fails.
 *  - one context per process per GPU; with nranks > 1 the x-slab decomposition of
 *    2LPT.c:47-113 is used (rank r owns mesh planes [r*N/P, (r+1)*N/P)).
 */
#ifndef MGPICOLA_H
#define MGPICOLA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgp_ctx mgp_ctx;

enum mgp_status {
  MGP_OK = 0,
  MGP_ERR_INVALID = -1,      /* bad argument / configuration */
  MGP_ERR_CUDA = -2,         /* CUDA, cuFFT or NCCL failure */
  MGP_ERR_BUFFER = -3,       /* particle buffer overflow: "increase Buffer" (auxPM.c:178-199, 250-254) */
  MGP_ERR_STATE = -4         /* call order violated (e.g. Kick before GetDisplacements) */
};

/* modified-gravity model == the reference's compile-time -D choice (Makefile:61-106) */
enum mgp_model {
  MGP_MODEL_NONE = 0,        /* modified_gravity_active = 0 */
  MGP_MODEL_FOFR = 1,        /* FOFRGRAVITY / MBETAMODEL: potential screening, mg.h:147-189 */
  MGP_MODEL_DGP = 2,         /* DGPGRAVITY: density screening + smoothing filter, mg.h:197-260 */
  MGP_MODEL_GEFF = 3         /* BRANSDICKE-like: delta_k *= Geff(a), mg.h:127-137 */
};

enum mgp_deposit_mode {
  MGP_DEPOSIT_ATOMIC = 0,    /* cell-sorted particles, warp-aggregated global reductions */
  MGP_DEPOSIT_TILE = 1,      /* cell-sorted particles, shared-memory tile accumulation */
  MGP_DEPOSIT_DETERMINISTIC = 2, /* cell-sorted, per-cell ordered gather: bitwise reproducible */
  MGP_DEPOSIT_ROWS = 3       /* per-step row bins (index list, records not moved); one warp owns one target row of a
                                shared-memory tile, the tile is written by one bulk copy (TMA); MtoParticles reads the
                                force rows from shared-memory tiles fetched by bulk copies */
};

enum mgp_grid_id {
  MGP_GRID_DENSITY = 0,      /* density / P3D          (vars.h:98-114) */
  MGP_GRID_FORCE_X = 1,      /* N11 / FN11 */
  MGP_GRID_FORCE_Y = 2,      /* N12 / FN12 */
  MGP_GRID_FORCE_Z = 3,      /* N13 / FN13 */
  MGP_GRID_MG_ONE = 4,       /* mgarray_one (vars.h:181-186) */
  MGP_GRID_MG_TWO = 5,       /* mgarray_two */
  MGP_GRID_SD_DELTA1 = 6,    /* cdelta_cdm  (vars.h:272; scale_dependent only, k-space) */
  MGP_GRID_SD_DELTA2 = 7     /* cdelta_cdm2 (vars.h:273) */
};

/* FIELD_* of proto.h:155-158 */
enum mgp_sd_field {
  MGP_FIELD_D = 0,           /* D(k, A)                         -> P.D   / P.D2    */
  MGP_FIELD_DDDY = 1,        /* dD/dy(k, A)                     -> P.dDdy / P.dD2dy */
  MGP_FIELD_DDDDDY = 2,      /* d^2D/dy^2(k, A)                 -> P.D   / P.D2    (2LPT.c:1821) */
  MGP_FIELD_DELTAD = 3       /* D(k, AFF) - D(k, A)             -> P.dDdy / P.dD2dy (2LPT.c:1822) */
};

typedef struct mgp_config {
  int nmesh;                 /* Nmesh */
  int nsample;               /* Nsample */
  double box;                /* Box (Mpc/h) */
  double buffer;             /* Buffer: particle capacity = ceil(NumPart0 * buffer) (main.c:228) */
  double omega;              /* Omega */
  int use_cola;              /* UseCOLA */
  int model;                 /* enum mgp_model */
  int include_screening;     /* include_screening */
  int grid_bytes;            /* 8 = double grids (reference default), 4 = -DSINGLE_PRECISION */
  int deposit_mode;          /* enum mgp_deposit_mode */
  int sort_particles;        /* 0: never; k >= 1: re-sort particles by mesh cell every k-th step (the TILE and
                                DETERMINISTIC deposits always sort; ATOMIC only needs locality) */
  int scale_dependent;       /* -DSCALEDEPENDENT: keep delta1_k / delta2_k and rebuild the displacement fields every
                                step from k-dependent growth factors (2LPT.c:1539-2005) */
  int rank, nranks;          /* slab decomposition: ThisTask / NTask */
  int device;                /* CUDA device ordinal */
  const void *nccl_unique_id;/* 128-byte ncclUniqueId shared by all ranks (NULL if nranks == 1) */
} mgp_config;

/* P(k) binning == the pofk_* parameters (user_defined_functions.h:303-337) */
typedef struct mgp_pofk_config {
  int nbins;
  int bintype;               /* 0 linear, 1 log */
  int subtract_shotnoise;
  double kmin, kmax;         /* h/Mpc */
} mgp_pofk_config;

/* Per-step scalars the reference evaluates per cell / per mode on the host
 * (SURVEY.md section 8(b)); the adapter computes them once per step with the reference's own
 * functions and passes numbers. */
typedef struct mgp_step_scalars {
  double a;                  /* aexp_global */
  /* MGP_MODEL_FOFR (mg.h:21-116, user_defined_functions.h:725-750) */
  double phi_crit;           /* screening_factor_potential's phicrit(a) */
  double coupling;           /* coupling_function(a) */
  double massterm2;          /* a^2 m^2(a) / (2 pi INVERSE_H0_MPCH / Box)^2 (mg.h:80) */
  /* MGP_MODEL_DGP (mg.h:197-309, user_defined_functions.h:757-773) */
  double dgp_fac0;           /* 8/9 Omega (rcH0/beta_DGP(a))^2 */
  double rsmooth;            /* Rsmooth_global (Mpc/h) */
  /* MGP_MODEL_GEFF */
  double geff;               /* GeffoverG(a, 0) */
  int compute_pofk;          /* pofk_compute_every_step: bin P(k) of the CDM density this step */
  /* MASSIVE_NEUTRINOS (auxPM.c:383-420): after the r2c, P3D = nu_cdmfac * P3D + nu_by_k2[m] * cdelta_cdm for every
   * mode but (0,0,0), m = |d|^2.  nu_cdmfac = (Omega - OmegaNu) / Omega; nu_by_k2[m] = OmegaNu / Omega * Nmesh^3 *
   * get_nu_transfer_function(k, a) / get_cdm_baryon_transfer_function(k, 1) at k = 2 pi sqrt(m) / Box.
   * NULL: no neutrinos.  Needs scale_dependent = 1 (the stored cdelta_cdm).  With compute_pofk the "total" spectrum
   * is binned after the add (auxPM.c:424-427) and read with mgp_get_step_power_spectrum_total. */
  const double *nu_by_k2;
  size_t n_nu;
  double nu_cdmfac;
} mgp_step_scalars;

const char *mgp_last_error(void);
int mgp_version(void);
/* CUDA devices visible to this process (0 when there is none): lets a multi-rank driver pick rank % count by default. */
int mgp_device_count(void);
/* fills 128 bytes with a fresh ncclUniqueId (rank 0 calls this and broadcasts it) */
int mgp_nccl_unique_id(void *out128);

/* replaces initialize_ffts + initialize_parts (2LPT.c:47-176) and the per-step malloc/plan churn
 * of MEMORY_MODE (auxPM.c:46-51, 61-71, 80-102): everything is allocated and planned once. */
int mgp_create(const mgp_config *cfg, mgp_ctx **out);
int mgp_destroy(mgp_ctx *ctx);

/* slab layout as the reference globals: Local_nx, Local_x_start, Local_np, Local_p_start, NumPart */
int mgp_get_layout(mgp_ctx *ctx, int *local_nx, int *local_x_start, int *local_np,
                   int *local_p_start, uint64_t *numpart);
/* k-space layout of this rank's grids after mgp_fft_r2c: 0 = [kx][ky][kz] as FFTW's non-transposed in-place r2c
 * (2LPT.c:50, single rank); 1 = [ky_local][kz][kx], the slab-decomposed transforms' transposed output (what
 * FFTW_MPI_TRANSPOSED_OUT would give); *ky_start / *ky_local = the ky rows this rank holds.  Pointers may be NULL. */
int mgp_kspace_layout(mgp_ctx *ctx, int *transposed, int *ky_start, int *ky_local);

/* ---- particle store (struct part_data, vars.h:293-321; GPU-resident SoA) ---- */
/* load n particles of this rank from host arrays laid out [n][3] (D2, id may be NULL) */
int mgp_upload_particles(mgp_ctx *ctx, uint64_t n, const float *pos, const float *vel,
                         const float *D, const float *D2, const uint64_t *id);
/* copy back for Output (main.c:792-1061); any pointer may be NULL; arrays sized NumPart */
int mgp_download_particles(mgp_ctx *ctx, float *pos, float *vel, float *D, float *D2, uint64_t *id);
/* Output(), main.c:915-997: the three GADGET blocks of this rank's particles, formed on the device exactly as Output()
 * forms them on the host and copied into the caller's buffers ([numpart][3] floats twice, [numpart] 64-bit IDs):
 *   pos = (float)(lengthfac * Pos);   vel = (float)(velfac*fac * (Vel - sumxyz + (D dDdy + D2 dD2dy) UseCOLA))
 * (scale_dependent: the float sum P.dDdy + P.dD2dy of the fields assigned last, which Output() makes FIELD_dDdy first,
 * main.c:824-825).  Replaces the download of the whole particle store before an output.  The copies overlap the
 * packing when the buffers are pinned (mgp_alloc_host). */
int mgp_pack_snapshot(mgp_ctx *ctx, double lengthfac, double velfac_times_fac, const double sumxyz[3], double dDdy,
                      double dD2dy, float *pos, float *vel, uint64_t *id);
/* pinned (page-locked, portable) host memory for particle buffers the driver hands to the library; NULL on failure */
void *mgp_alloc_host(size_t bytes);
void mgp_free_host(void *p);

/* Disp[3][NumPart] of MtoParticles (auxPM.c:605-630), as [n][3] */
int mgp_download_disp(mgp_ctx *ctx, float *disp);
/* the inverse: load Disp[3][NumPart] from the host as [n][3] (a driver that keeps its own Disp, tests) */
int mgp_upload_disp(mgp_ctx *ctx, const float *disp);

/* ---- initial conditions (2LPT.c:185-1520 Gaussian branch, main.c:257-309) ---- */
typedef struct mgp_ic_config {
  unsigned seed;             /* Seed */
  int sphere_mode;           /* SphereMode */
  int amplitude_fixed;       /* amplitude_fixed_initial_condition */
  int inverted;              /* inverted_initial_condition */
  /* P(k) [(Mpc/h)^3] for every integer m = |d|^2 in [0, 3 (Nmesh/2)^2], k = 2 pi sqrt(m) / Box: the
   * adapter fills it with PowerSpec(k) [* mg_pofk_ratio(k,1) / sigma8 ratio^2] exactly as
   * 2LPT.c:388-407 evaluates it per mode */
  const double *power_by_k2;
  size_t n_power;
  /* optional Nmesh*Nmesh seed table filled by the driver's own gsl_rng in the order of 2LPT.c:259-271;
   * NULL: the library fills it from `seed` with its ranlxd1 restatement */
  const unsigned *seedtable;
} mgp_ic_config;
/* displacement_fields(): delta_k, the six displacement gradients, the 2LPT source, and the ZA / 2LPT
 * displacements read out at the Lagrangian points of this rank's particle planes (13 FFTs) */
int mgp_ic_generate(mgp_ctx *ctx, const mgp_ic_config *ic);
/* READICFROMFILE: ReadFilesMakeDisplacementField + AssignDisplacementField (readICfromfile.c:133-215, 533-778).  The
 * driver keeps reading the RAMSES / GADGET / ASCII files itself and hands every file's positions over:
 *   begin   density = -1 (580-581)
 *   add     pos01[n][3] floats in [0, 1): the particles of this rank's slab are CIC-assigned with X = pos * Nmesh and
 *           W = (Nmesh/Nsample)^3 (ProcessParticlesSingleFile); *taken_total = how many this rank has taken so far
 *   finish  ghost-plane add, r2c, density *= normfac (the driver passes 1/Nmesh^3 * growth_DLCDM(1) /
 *           growth_DLCDM(a_init), 642-646), sharp-k filter above Nsample/2 when Nmesh > Nsample (648-680), CIC window
 *           deconvolution and rescale_by_k2[m] = sqrt(mg_pofk_ratio(k, 1)) [/ sigma8 ratio] at m = |d|^2 (701-778), then the
 *           2LPT pipeline of mgp_ic_generate on that delta_k.  Continue with mgp_ic_download / mgp_init_particles. */
int mgp_ic_particles_begin(mgp_ctx *ctx);
int mgp_ic_particles_add(mgp_ctx *ctx, const float *pos01, uint64_t n, uint64_t *taken_total);
int mgp_ic_particles_finish(mgp_ctx *ctx, double normfac, const double *rescale_by_k2, size_t n);
/* ZA[3][NumPart] / LPT[3][NumPart] of displacement_fields() (vars.h:265-270) for this rank's Lagrangian
 * particles, mean-subtracted, as [n][3]; valid between mgp_ic_generate and mgp_init_particles.  For a
 * driver that keeps main.c's own initialisation loop (main.c:257-309) on the host. */
int mgp_ic_download(mgp_ctx *ctx, float *za, float *lpt);
/* main.c:257-309: IDs, D = ZA, D2 = LPT, Vel (0 for COLA), Pos = wrap(q + D Di + D2 Di2) */
int mgp_init_particles(mgp_ctx *ctx, double Di, double Di2, double dDdy, double dD2dy);
/* the seed table of 2LPT.c:259-271 (Nmesh*Nmesh unsigned) and the n-th draw of ranlxd1(seed): test hooks
 * that pin the generator against GSL's published known-answer values */
int mgp_seedtable(unsigned seed, int nmesh, unsigned *out);
double mgp_ranlxd1_draw(unsigned long seed, long n);

/* ---- scale-dependent growth (-DSCALEDEPENDENT; 2LPT.c:1539-2005) ---- */
/* assign_displacment_field_to_particles(A, AF, AFF, fieldtype, LPTorder) (2LPT.c:1758): builds
 * i k / k^2 * growth(|k|) * delta^(order)_k, transforms it back, reads it out at the Lagrangian lattice, removes
 * the mean and stores it in the particle that was born there (fetched from the birth rank when it has migrated).
 * growth_by_k2[m], m = |d|^2 in [0, 3 (Nmesh/2)^2], is the factor from_cdisp_store_to_ZA evaluates per mode
 * WITHOUT its normfactor (2LPT.c:1611-1614): growth_X_scaledependent(k, A), or the difference for
 * MGP_FIELD_DELTAD, at k = 2 pi sqrt(m) / Box; the -3/7 / Nmesh^3 of order 2 is applied by the library.
 * The four per-particle fields stay valid until the next mgp_move_particles / mgp_get_displacements.
 * Before mgp_init_particles the call acts on the Lagrangian particles of this rank (main.c:231-251). */
int mgp_assign_displacement_field(mgp_ctx *ctx, int fieldtype, int lpt_order, const double *growth_by_k2, size_t n);
/* both orders in one pass (only D + D2 and dDdy + dD2dy are ever used: main.c:712, 767, 962): the sum goes
 * to the first-order slot, the second-order slot reads as zero.  6 instead of 12 inverse FFTs per step;
 * differs from two separate calls by one float rounding. */
int mgp_assign_displacement_fields_merged(mgp_ctx *ctx, int fieldtype, const double *growth1_by_k2,
                                          const double *growth2_by_k2, size_t n);
/* P.dDdy / P.dD2dy as [n][3] (P.D / P.D2 travel with mgp_upload/download_particles); any pointer may be NULL */
int mgp_download_sd_fields(mgp_ctx *ctx, float *dDdy, float *dD2dy);
int mgp_upload_sd_fields(mgp_ctx *ctx, const float *dDdy, const float *dD2dy);

/* ---- the per-step force path ---- */
/* MoveParticles (auxPM.c:108-275): slab ownership + migration; also (re)sorts by cell */
int mgp_move_particles(mgp_ctx *ctx);
/* PtoMesh (auxPM.c:280-432): CIC deposit, halo add, MG copy, r2c; optional P(k) */
int mgp_ptomesh(mgp_ctx *ctx, const mgp_step_scalars *s);
/* ComputeFifthForce (user_defined_functions.h:687-719 -> mg.h) */
int mgp_compute_fifth_force(mgp_ctx *ctx, const mgp_step_scalars *s);
/* Forces (auxPM.c:437-555): Green's function + gradient, 3 c2r, force halo */
int mgp_forces(mgp_ctx *ctx);
/* MtoParticles (auxPM.c:560-644): trilinear gather; returns sumDxyz (already / TotNumPart) */
int mgp_mtoparticles(mgp_ctx *ctx, double sumDxyz[3]);
/* GetDisplacements (auxPM.c:37-103) = the five calls above */
int mgp_get_displacements(mgp_ctx *ctx, const mgp_step_scalars *s, double sumDxyz[3]);

/* Kick particle loop (main.c:707-739).  dda = Sphi(...), ddDddy/ddD2ddy = growth_ddDddy(A),
 * growth_ddD2ddy(A) stay on the host (ignored when scale_dependent: P.D / P.D2 already hold the weighted fields).  sumDxyz in (0 on the second kick of an output step,
 * main.c:566-569); sumxyz out (mean velocity, / TotNumPart). */
int mgp_kick(mgp_ctx *ctx, double A, double dda, double ddDddy, double ddD2ddy,
             const double sumDxyz[3], double sumxyz[3]);
/* Drift particle loop (main.c:762-783).  dyyy = Sq(...), deltaD = growth_D(AFF)-Di, ... (deltaD, deltaD2 ignored
 * when scale_dependent: P.dDdy / P.dD2dy hold the per-particle increments) */
int mgp_drift(mgp_ctx *ctx, double dyyy, double deltaD, double deltaD2, const double sumxyz[3]);

/* ---- lightcone (-DLIGHTCONE; lightcone.c:265-474) ---- */
/* Drift_Lightcone: the drift of one step, done replicate by replicate so that every particle image that leaves the
 * shrinking lightcone during [A, AFF] is written out at its interpolated exit position.  The scalar part stays on the
 * host exactly as in lightcone.c:294-347 (set_lightcone, flag_replicates and the file writers are host C too); the
 * particle loop (392-471) runs here.  Not available with scale_dependent (the reference refuses that build, Makefile:279). */
typedef struct mgp_lightcone_step {
  double A, AFF;             /* scale factor at the start / end of the drift */
  double dyyy;               /* Sq(A, AFF, AF) (or the StdDA variants, lightcone.c:302-308) */
  double da1, da2;           /* growth_D(AFF) - Di, growth_D2(AFF) - Di2 (310-311) */
  double dv1, dv2;           /* growth_dDdy(AF), growth_dD2dy(AF) (315-316) */
  double sumxyz[3];          /* mean velocity left by Kick */
  double rcomov_old, rcomov_new;   /* Light / Hubble * SphiStd(A, 1), ... (AFF, 1) (286-287) */
  double origin[3];          /* Origin_x, Origin_y, Origin_z */
  double boundary;           /* 20 Mpc/h (283): a larger Delta_Pos component is the reference's FatalError (403-407) */
  double lengthfac;          /* UnitLength_in_cm / 3.085678e24 */
  double velfac_times_fac;   /* UnitVelocity_in_cm_per_s / 1e5 * Hubble / AF */
  /* the exit-time lookup tables of lightcone.c:326-347: ntab nodes AL_tab (increasing) with da1_tab = growth_D(AL) - Di,
   * da2_tab = growth_D2(AL) - Di2, dyyy_tab = Sq(A, AL, AF); the library builds gsl_interp_cspline's natural cubic
   * splines through them (349-357) and evaluates them at every exit time */
  int ntab;
  const double *al_tab, *da1_tab, *da2_tab, *dyyy_tab;
  /* replicates with repflag == 0 (flag_replicates, lightcone.c:62-189), as offsets (i, j, k) in the order of the
   * reference's triple loop (411-413): rep_ijk[3 r .. 3 r + 2] */
  int nrep;
  const int *rep_ijk;
} mgp_lightcone_step;
/* how many particle images leave the lightcone in this step, per listed replicate (count[nrep]); changes nothing */
int mgp_lightcone_count(mgp_ctx *ctx, const mgp_lightcone_step *ls, uint64_t *count);
/* the drift itself: Pos = periodic_wrap(Pos + Delta_Pos) for every particle (468-470) and, for every image that leaves,
 * one row {x, y, z [Mpc/h], vx, vy, vz [km/s]} of six floats at block[(r * cap + slot) * 6] exactly as lightcone.c:447-453
 * forms it -- the layout Output_Lightcone (482-565) reads with blockmaxlen = cap.  count[r] = rows of replicate r.
 * MGP_ERR_BUFFER (nothing moved) when a replicate needs more than cap rows: size cap from mgp_lightcone_count. */
int mgp_drift_lightcone(mgp_ctx *ctx, const mgp_lightcone_step *ls, uint64_t cap, float *block, uint64_t *count);

/* ---- FoF halo finder on the fly (-DMATCHMAKER_HALOFINDER; mm_main.c:129-385, mm_fof.c) ---- */
/* what main.c:834-868 hands to MatchMaker(), as numbers */
typedef struct mgp_fof_config {
  double norm_pos;           /* lengthfac: code positions -> Mpc/h */
  double norm_vel;           /* Hubble / A / A: code velocities -> comoving dx/dt in km/s (main.c:866) */
  double boxsize;            /* Box * lengthfac */
  double dx_extra;           /* mm_dx_extra_mpc * lengthfac: width of the strip shared with the left neighbour */
  double b_fof;              /* mm_linking_length, in units of the mean inter-particle distance */
  int np_min;                /* mm_min_npart_halo */
  double mass_part;          /* particle mass in 1e10 Msun/h (main.c:853) */
  double dDdy, dD2dy;        /* growth_dDdy(A), growth_dD2dy(A): the LPT velocity added back (mm_main.c:301-303, 327);
                                ignored when scale_dependent: P.dDdy / P.dD2dy must then hold FIELD_dDdy (main.c:831-832) */
} mgp_fof_config;
/* FoFHalo of mm_common.h:115-128, field for field (write_halos() of mm_snap_io.c takes an array of these) */
typedef struct mgp_fof_halo {
  int np;
  float m_halo;
  float x_avg[3], x_rms[3], v_avg[3], v_rms[3], lam[3];
  float b, c;
  float ea[3], eb[3], ec[3];
} mgp_fof_halo;
/* picola_to_matchmaker_particles + fof_get_halos for this rank's slab: particles translated and ordered by x, the strip
 * x <= dx_extra handed to the left neighbour, friends-of-friends groups (x open, y and z periodic, linking length
 * b_fof * Box / Nsample), groups the left neighbour also found given up, and the properties of every group with at least
 * np_min members.  *n_halos = halos of this rank; they stay in the context until the next call. */
int mgp_fof_find(mgp_ctx *ctx, const mgp_fof_config *cfg, uint64_t *n_halos);
/* the halos of the last mgp_fof_find, by decreasing np (fof_get_halos returns them in that order, mm_fof.c:459) */
int mgp_fof_get(mgp_ctx *ctx, mgp_fof_halo *out);

/* ---- P(k) (compute_pofk.c:71-271) ---- */
int mgp_set_pofk_config(mgp_ctx *ctx, const mgp_pofk_config *pc);
/* bins the k-space density currently in MGP_GRID_DENSITY; arrays sized mgp_pofk_nbins() */
int mgp_pofk_nbins(mgp_ctx *ctx);
int mgp_compute_power_spectrum(mgp_ctx *ctx, double *pofk, double *kmean, double *nmodes);
/* result of the in-step binning requested through mgp_step_scalars.compute_pofk */
int mgp_get_step_power_spectrum(mgp_ctx *ctx, double *pofk, double *kmean, double *nmodes);

/* "total" (CDM + baryons + neutrinos) spectrum of the last step that had nu_by_k2 and compute_pofk set */
int mgp_get_step_power_spectrum_total(mgp_ctx *ctx, double *pofk, double *kmean, double *nmodes);

/* compute_RSD_powerspectrum (compute_pofk.c:403-512): redshift-space deposits along y and along z (PtoMesh_RSD,
 * 280-393), two r2c, the P0 / P2 / P4 binning of bin_up_RSD_power_spectrum (518-753).  Uses MGP_GRID_FORCE_X as the
 * work grid (free between PtoMesh and Forces, and in Output).
 *   vnorm = (Hubble / A) / (100 A hubble(A)) * Nmesh / Box  (compute_pofk.c:291-292), dDdy / dD2dy = growth_dDdy(A),
 *   growth_dD2dy(A) (ignored when scale_dependent: P.dDdy / P.dD2dy must then hold FIELD_dDdy, i.e. the caller does
 *   what compute_pofk.c:409-428 does around the call).
 * Outputs sized mgp_pofk_nbins(): per line of sight q in {y, z}: n, k (bin centre), P0, P2, P4 as
 * out_y[0..4][nbins], out_z[0..4][nbins]; the file the reference writes holds their mean and |y - z| / sqrt(2). */
int mgp_compute_rsd_power_spectrum(mgp_ctx *ctx, double vnorm, double dDdy, double dD2dy, double *out_y, double *out_z);

/* ---- grid access (tests, write_grid_to_file auxPM.c:757-813) ---- */
/* SimplePofk/main.cpp (the reference's stand-alone P(k) estimator) on the particles the context holds: assignment of
 * counts with scheme 1 = NGP, 2 = CIC, 3 = TSC (main.cpp:59-228; x = double(pos_float / boxsize) * ngrid), normalised to the
 * density contrast as main() does (513-533: a factor (Nmesh^3 / Npart)^2 on every bin, k = 0 is not binned), one transform,
 * |d_k|^2 / N^6 / window^2 with window = prod sinc(pi k_a / N)^scheme (324-340, 393), bins int(|k| + 0.5), 0 < bin < Nmesh
 * (391), mean per bin, optional shot noise 1 / Npart (420-424).  pofk, nmodes: Nmesh doubles each; the tool prints
 * k = (2 i + 1) pi / Box and pofk[i] Box^3 for 1 <= i <= Nmesh / 2.  tsc_as_published = 1 reproduces the tool's TSC as
 * written (the P_z weight of the three "this y" lines lands on the next z plane, main.cpp:189, 201, 213); 0 = the textbook
 * stencil.  One rank only; overwrites force grid X. */
int mgp_simple_pofk(mgp_ctx *ctx, int scheme, int subtract_shotnoise, int tsc_as_published, double *pofk, double *nmodes);

/* local slab incl. ghost plane: (Local_nx+1) * Nmesh * 2*(Nmesh/2+1) values of grid_bytes */
size_t mgp_grid_local_values(mgp_ctx *ctx);
int mgp_download_grid(mgp_ctx *ctx, int grid_id, void *host);
int mgp_upload_grid(mgp_ctx *ctx, int grid_id, const void *host);
/* in-place transforms of one grid (my_fftw_execute on r2c / c2r plans, wrappers.c:42-46) */
int mgp_fft_r2c(mgp_ctx *ctx, int grid_id);
int mgp_fft_c2r(mgp_ctx *ctx, int grid_id);
/* developer probe of the slab exchange (slab-decomposed contexts only): mean device time in ms of `reps` back-to-back
 * executions of one piece -- 0 flag barrier, 1 fused backward x-transform + push, 2 fused pull + forward x-transform,
 * 3 / 4 backward / forward transpose kernel, 5 copy-engine peer copy of the remote share of a slab, 6 / 7 the backward /
 * forward strided copy-engine blocks of the DMA exchange.  Clobbers the force grids and the transpose buffers. */
int mgp_debug_time_exchange(mgp_ctx *ctx, int which, int reps, float *ms);

/* ---- instrumentation ---- */
/* number of kernels this library has launched on ctx since creation / last reset */
uint64_t mgp_launch_count(mgp_ctx *ctx, int reset);
/* CUDA-event time (ms) accumulated per phase since last reset; names mirror timer.h
 * (MoveParticles, PtoMesh, FFT, ComputeFifthForce, Forces, MtoParticles, Kick, Drift, Pofk) */
int mgp_phase_count(void);
const char *mgp_phase_name(int i);
int mgp_phase_times_ms(mgp_ctx *ctx, double *ms, uint64_t *calls, int reset);
/* enable/disable per-phase event timing (adds synchronisation; off by default) */
int mgp_set_phase_timing(mgp_ctx *ctx, int on);
/* the stream all work is issued on (cudaStream_t), for external event timing */
void *mgp_stream(mgp_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
