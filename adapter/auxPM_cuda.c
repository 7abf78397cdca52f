/*
 * auxPM_cuda.c -- the reference-side binding of libmgpicola_cuda.so.
 *
 * Compiled INSTEAD OF the reference's auxPM.c (+mg.h), 2LPT.c and compute_pofk.c, together with the
 * reference's unmodified cosmo.c, power.c, read_param.c, vars.c, timer.c, msg.c, wrappers.c, jbd.c
 * and a main.c that carries the small patch of adapter/main_c.patch (Kick / Drift particle loops
 * and the host copy of the particles before Output).  It keeps the reference's function names and
 * talks to the library through the reference's own globals (vars.h), so main()'s call sites are
 * untouched.  Everything the reference evaluates per cell or per mode on the host is evaluated here
 * ONCE per step with the reference's own functions and handed over as numbers:
 *     coupling_function, mass2_of_a, screening_factor_potential  -> mgp_step_scalars
 *     PowerSpec, mg_pofk_ratio, mg_sigma8_enhancement            -> mgp_ic_config.power_by_k2
 *
 * Particle residency: main.c builds its host array P exactly as before (main.c:257-309, from the
 * ZA / LPT arrays this file fills from the GPU); the first GetDisplacements() moves P to the GPU,
 * where it stays.  mgp_adapter_sync_host() brings it back before Output() reads P.
 */
#include "vars.h"
#include "proto.h"
#include "timer.h"
#include "mgpicola.h"

#include <limits.h>

static mgp_ctx *g_ctx = NULL;
static int g_host_is_newer = 1;      /* host P was (re)built: upload before the next force evaluation */
static int g_host_synced = 0;        /* host P mirrors the device order (between mgp_adapter_sync_host and the next force evaluation) */
static size_t g_host_capacity = 0;

static void ck(int rc, const char *where) {
  if (rc != MGP_OK) {
    printf("%s: %s\n", where, mgp_last_error());
    FatalError((char *) where);
  }
}

/* ------------------------------------------------------------------ small host utilities that lived in the replaced files */

void set_units(void) {                       /* 2LPT.c:35-42 */
  UnitTime_in_s = UnitLength_in_cm / UnitVelocity_in_cm_per_s;
  G = GRAVITY / pow(UnitLength_in_cm, 3) * UnitMass_in_g * pow(UnitTime_in_s, 2);
  Hubble = HUBBLE * UnitTime_in_s;
}

#if (MEMORY_MODE || SINGLE_PRECISION)
float periodic_wrap(float x) {               /* auxPM.c:649-655 */
  while (x >= (float) Box) x -= (float) Box;
  while (x < 0) x += (float) Box;
  if (x == (float) Box) x = 0.0;
  return x;
}
#else
double periodic_wrap(double x) {
  while (x >= Box) x -= Box;
  while (x < 0) x += Box;
  if (x == Box) x = 0.0;
  return x;
}
#endif

void FatalError(char *errmsg) {              /* auxPM.c:668-673 */
  printf("Fatal Error: [%s]\n", errmsg);
  fflush(stdout);
  MPI_Abort(MPI_COMM_WORLD, 1);
  exit(1);
}

size_t my_fwrite(void *ptr, size_t size, size_t nmemb, FILE *stream) {   /* auxPM.c:678-686 */
  size_t nwritten = fwrite(ptr, size, nmemb, stream);
  if (nwritten != nmemb) {
    printf("\nERROR: I/O error (fwrite) on task=%d has occured.\n\n", ThisTask);
    fflush(stdout);
    FatalError((char *) "my_fwrite");
  }
  return nwritten;
}

/* particles never travel through MPI any more (the library migrates them over NCCL) */
void create_MPI_type_for_Particles(MPI_Datatype *t) { *t = MPI_BYTE; }

/* referenced by the ComputeFifthForce dispatcher that user_defined_functions.h compiles into cosmo.o;
 * the library runs the solvers itself inside mgp_get_displacements */
void ComputeFifthForce_PotentialScreening(void) {}
void ComputeFifthForce_DensityScreening(void) {}
void ComputeFifthForce_TimeDepGeffModels(void) {}
void ComputeFifthForce_GradientScreening(void) {}

/* ------------------------------------------------------------------ slab layout (2LPT.c:47-176) */

static int model_id(void) {
  if (!modified_gravity_active) return MGP_MODEL_NONE;
#if defined(FOFRGRAVITY) || defined(MBETAMODEL)
  return MGP_MODEL_FOFR;
#elif defined(DGPGRAVITY)
  return MGP_MODEL_DGP;
#elif defined(BRANSDICKE)
  return MGP_MODEL_GEFF;
#else
  return MGP_MODEL_NONE;
#endif
}

void initialize_ffts(void) {
  mgp_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.nmesh = Nmesh; cfg.nsample = Nsample; cfg.box = Box; cfg.buffer = Buffer; cfg.omega = Omega;
  cfg.use_cola = UseCOLA; cfg.model = model_id(); cfg.include_screening = include_screening;
  cfg.grid_bytes = (int) sizeof(float_kind);
  cfg.deposit_mode = MGP_DEPOSIT_ATOMIC; cfg.sort_particles = 2;
  {
    const char *e = getenv("MGP_DEPOSIT_MODE");
    if (e) cfg.deposit_mode = atoi(e);
    e = getenv("MGP_SORT_INTERVAL");
    if (e) cfg.sort_particles = atoi(e);
  }
#ifdef SCALEDEPENDENT
  cfg.scale_dependent = 1;
#endif
  cfg.rank = ThisTask; cfg.nranks = NTask;
  {
    /* one GPU per rank: rank modulo the visible devices unless MGP_DEVICE names one (ranks of one node; with several
       nodes set MGP_DEVICE to the node-local rank, e.g. from OMPI_COMM_WORLD_LOCAL_RANK) */
    const int ndev = mgp_device_count();
    const char *e = getenv("MGP_DEVICE");
    cfg.device = e ? atoi(e) : (ndev > 0 ? ThisTask % ndev : 0);
    if (NTask > 1 && Nmesh % NTask != 0) {
      if (ThisTask == 0) printf("[mgpicola-cuda] NTask = %d does not divide Nmesh = %d: the CUDA library needs equal slabs\n", NTask, Nmesh);
      MPI_Abort(MPI_COMM_WORLD, 1);
    }
  }
  static char nccl_id[128];
  if (NTask > 1) {
    if (ThisTask == 0) ck(mgp_nccl_unique_id(nccl_id), "mgp_nccl_unique_id");
    MPI_Bcast(nccl_id, 128, MPI_BYTE, 0, MPI_COMM_WORLD);
    cfg.nccl_unique_id = nccl_id;
  }
  ck(mgp_create(&cfg, &g_ctx), "mgp_create");
  int lnx, lx0, lnp, lp0;
  uint64_t np;
  ck(mgp_get_layout(g_ctx, &lnx, &lx0, &lnp, &lp0, &np), "mgp_get_layout");
  Local_nx = lnx; Local_x_start = lx0;
  alloc_slice = Nmesh * (Nmesh / 2 + 1);
  alloc_local = Local_nx * alloc_slice;
  last_slice = Local_nx * alloc_slice;
  Total_size = alloc_local + alloc_slice;

  Local_nx_table = my_malloc(sizeof(int) * NTask);
  int lnx_i = (int) Local_nx;
  MPI_Allgather(&lnx_i, 1, MPI_INT, Local_nx_table, 1, MPI_INT, MPI_COMM_WORLD);
  LeftTask = ThisTask;
  do { LeftTask--; if (LeftTask < 0) LeftTask = NTask - 1; } while (Local_nx_table[LeftTask] == 0);
  RightTask = ThisTask;
  do { RightTask++; if (RightTask >= NTask) RightTask = 0; } while (Local_nx_table[RightTask] == 0);
  Slab_to_task = my_malloc(sizeof(int) * Nmesh);
  int block = (Nmesh + NTask - 1) / NTask;
  for (int i = 0; i < Nmesh; i++) Slab_to_task[i] = i / block < NTask ? i / block : NTask - 1;
  if (ThisTask == 0) printf("\n[mgpicola-cuda] Local_nx = %d, Local_x_start = %d (library version %d)\n", lnx, lx0, mgp_version());
}

void initialize_parts(void) {
  int lnx, lx0, lnp, lp0;
  uint64_t np;
  ck(mgp_get_layout(g_ctx, &lnx, &lx0, &lnp, &lp0, &np), "mgp_get_layout");
  Local_np = lnp; Local_p_start = lp0;
  NumPart = (unsigned int) (Local_np * Nsample * Nsample);
  TotNumPart = ((unsigned long long) Nsample) * ((unsigned long long) Nsample) * ((unsigned long long) Nsample);
  Local_np_table = my_malloc(sizeof(int) * NTask);
  int lnp_i = (int) Local_np;
  MPI_Allgather(&lnp_i, 1, MPI_INT, Local_np_table, 1, MPI_INT, MPI_COMM_WORLD);
  NTaskWithN = 0;
  for (int i = 0; i < NTask; i++) if (Local_np_table[i] > 0) NTaskWithN++;
  Part_to_task = my_malloc(sizeof(int) * Nsample);
  for (int i = 0; i < Nsample; i++) Part_to_task[i] = Slab_to_task[(int) ((double) (i * Nmesh) / (double) Nsample)];
  g_host_capacity = (size_t) ceil(NumPart * Buffer);
  if (ThisTask == 0) printf("Total number of particles = %llu\n\n", TotNumPart);
}

/* ------------------------------------------------------------------ initial conditions (2LPT.c:185-1520) */

/* non-SCALEDEPENDENT builds: main.c:257-309 builds P from ZA / LPT on the host */
static void fetch_za_lpt(void) {
#ifndef SCALEDEPENDENT
  float *za = my_malloc((size_t) NumPart * 3 * sizeof(float)), *lpt = my_malloc((size_t) NumPart * 3 * sizeof(float));
  ck(mgp_ic_download(g_ctx, za, lpt), "mgp_ic_download");
  for (int a = 0; a < 3; a++) {
    ZA[a] = my_malloc(NumPart * sizeof(float));
    LPT[a] = my_malloc(NumPart * sizeof(float));
    for (unsigned int q = 0; q < NumPart; q++) { ZA[a][q] = za[3 * (size_t) q + a]; LPT[a][q] = lpt[3 * (size_t) q + a]; }
  }
  my_free(za); my_free(lpt);
#endif
  g_host_is_newer = 1;
}

void displacement_fields(void) {
  timer_start(_DisplacementFields);
  const int h = Nmesh / 2;
  const size_t nm = (size_t) 3 * h * h + 1;
  double *power = my_malloc(sizeof(double) * nm);
  double s8ratio2 = 1.0;
  if (modified_gravity_active && input_pofk_is_for_lcdm) s8ratio2 = pow(mg_sigma8_enhancement(1.0), 2);   /* 2LPT.c:323-326 */
  power[0] = 0.0;
  for (size_t m = 1; m < nm; m++) {
    const double kmag = 2 * PI / Box * sqrt((double) m);
    double p = PowerSpec(kmag);                                               /* 2LPT.c:388 */
    if (modified_gravity_active && input_pofk_is_for_lcdm) {                  /* 2LPT.c:395-406 */
      p *= mg_pofk_ratio(kmag, 1.0);
      if (!input_sigma8_is_for_lcdm) p /= s8ratio2;
    }
    power[m] = p;
  }
  mgp_ic_config ic;
  memset(&ic, 0, sizeof(ic));
  ic.seed = (unsigned) Seed; ic.sphere_mode = SphereMode;
  ic.amplitude_fixed = amplitude_fixed_initial_condition; ic.inverted = inverted_initial_condition;
  ic.power_by_k2 = power; ic.n_power = nm; ic.seedtable = NULL;
  ck(mgp_ic_generate(g_ctx, &ic), "mgp_ic_generate");
  my_free(power);
  fetch_za_lpt();
  timer_stop(_DisplacementFields);
}

#ifdef READICFROMFILE
/* ------------------------------------------------------------------ initial conditions from particle files
 * ReadFilesMakeDisplacementField (readICfromfile.c:533-699) with the reference's own file readers (readICfromfile.c is
 * compiled unmodified, its ReadFilesMakeDisplacementField renamed out of the way by the Makefile): every file's particles
 * go to mgp_ic_particles_add in place of ProcessParticlesSingleFile, the transform, the normalisation, the sharp-k filter,
 * AssignDisplacementField and the 2LPT pipeline run in mgp_ic_particles_finish.  GADGET files arrive as the floats the
 * reference deposits; RAMSES / ASCII positions are doubles there and are rounded to float here. */
void ReadFilesMakeDisplacementField(void) {
  timer_start(_ReadParticlesFromFile);
  int maxpart = 0;
  if (TypeInputParticleFiles == RAMSESFILE) maxpart = find_maxpart_ramsesfiles(InputParticleFileDir, RamsesOutputNumber, NumInputParticleFiles);
  else if (TypeInputParticleFiles == ASCIIFILE) maxpart = find_maxpart_asciifiles(InputParticleFileDir, InputParticleFilePrefix, NumInputParticleFiles);
  else if (TypeInputParticleFiles == GADGETFILE) maxpart = find_maxpart_gadgetfiles(InputParticleFileDir, InputParticleFilePrefix, NumInputParticleFiles);
  else {
    printf("Error: unknown file-format [%i]\n", TypeInputParticleFiles);
    MPI_Abort(MPI_COMM_WORLD, 1);
    exit(1);
  }
  char *buffer = my_malloc(3 * sizeof(double) * (size_t) (maxpart > 0 ? maxpart : 1));
  float *as_float = NULL;
  if (TypeInputParticleFiles != GADGETFILE) as_float = my_malloc(3 * sizeof(float) * (size_t) (maxpart > 0 ? maxpart : 1));
  ck(mgp_ic_particles_begin(g_ctx), "mgp_ic_particles_begin");
  int npart_read = 0;
  uint64_t taken = 0;
  for (int filenum = 1; filenum <= NumInputParticleFiles; filenum++) {
    int n = 0;
    if (TypeInputParticleFiles == RAMSESFILE) n = read_ramses_file(InputParticleFileDir, RamsesOutputNumber, filenum, buffer, &npart_read);
    else if (TypeInputParticleFiles == ASCIIFILE) n = read_ascii_file(InputParticleFileDir, InputParticleFilePrefix, filenum, buffer, &npart_read);
    else n = read_gadget_file(InputParticleFileDir, InputParticleFilePrefix, filenum - 1, buffer, &npart_read);
    if (ThisTask == 0) printf("Read so far: %i  Part in current file %i\n", npart_read, n);
    const float *pos01 = (const float *) buffer;                       /* GADGET: [x1 y1 z1 x2 ...] floats in [0, 1) */
    if (as_float) {                                                     /* RAMSES, ASCII: [x1 .. xn y1 .. yn z1 .. zn] doubles */
      const double *d = (const double *) buffer;
      for (int i = 0; i < n; i++)
        for (int a = 0; a < 3; a++) as_float[3 * (size_t) i + a] = (float) d[i + (size_t) a * n];
      pos01 = as_float;
    }
    ck(mgp_ic_particles_add(g_ctx, pos01, (uint64_t) n, &taken), "mgp_ic_particles_add");
  }
  my_free(buffer);
  if (as_float) my_free(as_float);
  /* readICfromfile.c:641-643 and 735-741 */
  const double normfac = 1.0 / pow((double) Nmesh, 3) * (growth_DLCDM(1.0) / growth_DLCDM(1.0 / (1.0 + Init_Redshift)));
  const int h = Nmesh / 2;
  const size_t nm = (size_t) 3 * h * h + 1;
  double *rescale = my_malloc(sizeof(double) * nm);
  const double s8 = mg_sigma8_enhancement(1.0);
  rescale[0] = 1.0;
  for (size_t m = 1; m < nm; m++) {
    const double kmag = sqrt((double) m * (2 * PI / Box) * (2 * PI / Box));
    rescale[m] = sqrt(mg_pofk_ratio(kmag, 1.0));
    if (!input_sigma8_is_for_lcdm) rescale[m] /= s8;
  }
  if (ThisTask == 0) printf("Done precomputing delta(k) from particles, now compute displacement-fields\n\n");
  ck(mgp_ic_particles_finish(g_ctx, normfac, rescale, nm), "mgp_ic_particles_finish");
  my_free(rescale);
  fetch_za_lpt();
  timer_stop(_ReadParticlesFromFile);
}
#endif

/* ------------------------------------------------------------------ host <-> device particle copies */

static void upload_host_particles(void) {
  const size_t n = NumPart;
  float *pos = my_malloc(n * 12), *vel = my_malloc(n * 12), *d1 = my_malloc(n * 12), *d2 = my_malloc(n * 12);
  uint64_t *id = my_malloc(n * 8);
  for (size_t i = 0; i < n; i++) {
    for (int a = 0; a < 3; a++) {
      pos[3 * i + a] = P[i].Pos[a]; vel[3 * i + a] = P[i].Vel[a]; d1[3 * i + a] = P[i].D[a]; d2[3 * i + a] = P[i].D2[a];
    }
#ifdef PARTICLE_ID
    id[i] = P[i].ID;
#else
    id[i] = ((uint64_t) Local_p_start * Nsample * Nsample) + i;
#endif
  }
  ck(mgp_upload_particles(g_ctx, n, pos, vel, d1, d2, id), "mgp_upload_particles");
  my_free(pos); my_free(vel); my_free(d1); my_free(d2); my_free(id);
  g_host_is_newer = 0;
  g_host_synced = 0;
}

/* GADGET snapshots are packed on the GPU (mgp_pack_snapshot): Output() never reads P, only NumPart.  The ASCII output of a
 * build without -DGADGET_STYLE and MGP_HOST_SNAPSHOT=1 (the previous behaviour) still bring P back.  The FoF hook of
 * main.c:830-872 no longer needs P either (MatchMaker below). */
static int gpu_snapshot(void) {
#if defined(GADGET_STYLE)
  static int on = -1;
  if (on < 0) { const char *e = getenv("MGP_HOST_SNAPSHOT"); on = !(e && atoi(e) != 0); }
  return on;
#else
  return 0;
#endif
}

/* called (through the main.c patch) before Output() reads P */
void mgp_adapter_sync_host(void) {
  int lnx, lx0, lnp, lp0;
  uint64_t np;
  if (g_host_is_newer) {
    if (!gpu_snapshot()) return;
    upload_host_particles();          /* an output before the first force evaluation: the blocks are packed on the device */
  }
  ck(mgp_get_layout(g_ctx, &lnx, &lx0, &lnp, &lp0, &np), "mgp_get_layout");
  if (gpu_snapshot()) { NumPart = (unsigned int) np; return; }
  if (np > g_host_capacity) FatalError((char *) "mgp_adapter_sync_host: more particles than the host buffer holds; increase Buffer");
  const size_t n = (size_t) np;
  float *pos = my_malloc(n * 12), *vel = my_malloc(n * 12), *d1 = my_malloc(n * 12), *d2 = my_malloc(n * 12);
  uint64_t *id = my_malloc(n * 8);
  ck(mgp_download_particles(g_ctx, pos, vel, d1, d2, id), "mgp_download_particles");
  for (size_t i = 0; i < n; i++) {
    for (int a = 0; a < 3; a++) {
      P[i].Pos[a] = pos[3 * i + a]; P[i].Vel[a] = vel[3 * i + a]; P[i].D[a] = d1[3 * i + a]; P[i].D2[a] = d2[3 * i + a];
    }
#ifdef PARTICLE_ID
    P[i].ID = id[i];
#endif
  }
  NumPart = (unsigned int) n;
  my_free(pos); my_free(vel); my_free(d1); my_free(d2); my_free(id);
#ifdef SCALEDEPENDENT
  {
    float *f1 = my_malloc(n * 12), *f2 = my_malloc(n * 12);
    ck(mgp_download_sd_fields(g_ctx, f1, f2), "mgp_download_sd_fields");
    for (size_t i = 0; i < n; i++)
      for (int a = 0; a < 3; a++) { P[i].dDdy[a] = f1[3 * i + a]; P[i].dD2dy[a] = f2[3 * i + a]; }
    my_free(f1); my_free(f2);
  }
#endif
  g_host_synced = 1;
}

/* called (through the main.c patch) from Output() in place of its three block loops (main.c:936-997) */
void mgp_adapter_write_gadget_blocks(FILE *fp, double lengthfac, double velfac_times_fac, double dDdy, double dD2dy) {
  static float *pos = NULL, *vel = NULL;
  static uint64_t *id = NULL;
  static size_t cap = 0;
  static int pinned = 1;
  if (!gpu_snapshot()) FatalError((char *) "mgp_adapter_write_gadget_blocks: GPU snapshot packing is off in this build");
  if (g_host_is_newer) upload_host_particles();       /* an output before the first force evaluation */
  int lnx, lx0, lnp, lp0;
  uint64_t np;
  ck(mgp_get_layout(g_ctx, &lnx, &lx0, &lnp, &lp0, &np), "mgp_get_layout");
  const size_t n = (size_t) np;
  if (n > cap) {
    if (pos) { if (pinned) { mgp_free_host(pos); mgp_free_host(vel); mgp_free_host(id); } else { free(pos); free(vel); free(id); } }
    cap = n + n / 8 + 1024;
    pos = mgp_alloc_host(cap * 12); vel = mgp_alloc_host(cap * 12); id = mgp_alloc_host(cap * 8);
    pinned = pos && vel && id;
    if (!pinned) {                                     /* no pinned memory to be had: pageable buffers work too */
      if (pos) mgp_free_host(pos);
      if (vel) mgp_free_host(vel);
      if (id) mgp_free_host(id);
      pos = malloc(cap * 12); vel = malloc(cap * 12); id = malloc(cap * 8);
      if (!pos || !vel || !id) FatalError((char *) "mgp_adapter_write_gadget_blocks: out of host memory");
    }
  }
  ck(mgp_pack_snapshot(g_ctx, lengthfac, velfac_times_fac, sumxyz, dDdy, dD2dy, pos, vel, id), "mgp_pack_snapshot");
  int dummy = (int) (sizeof(float) * 3 * n);
  my_fwrite(&dummy, sizeof(dummy), 1, fp);
  if (n) my_fwrite(pos, sizeof(float), 3 * n, fp);
  my_fwrite(&dummy, sizeof(dummy), 1, fp);
  my_fwrite(&dummy, sizeof(dummy), 1, fp);
  if (n) my_fwrite(vel, sizeof(float), 3 * n, fp);
  my_fwrite(&dummy, sizeof(dummy), 1, fp);
#ifdef PARTICLE_ID
  dummy = (int) (sizeof(unsigned long long) * n);
  my_fwrite(&dummy, sizeof(dummy), 1, fp);
  if (n) my_fwrite(id, sizeof(unsigned long long), n, fp);
  my_fwrite(&dummy, sizeof(dummy), 1, fp);
#endif
}

#ifdef SCALEDEPENDENT
/* ------------------------------------------------------------------ scale-dependent displacement fields (2LPT.c:1539-2005) */

/* growth factor of from_cdisp_store_to_ZA (2LPT.c:1611-1614, without normfactor) at every integer m = |d|^2 */
static double *sd_growth_table(double A, double AFF, int fieldtype, int LPTorder, size_t *n_out) {
  const int h = Nmesh / 2;
  const size_t nm = (size_t) 3 * h * h + 1;
  double *g = my_malloc(sizeof(double) * nm);
  double (*fD)(double, double) = LPTorder == 1 ? &growth_D_scaledependent : &growth_D2_scaledependent;
  double (*fdD)(double, double) = LPTorder == 1 ? &growth_dDdy_scaledependent : &growth_dD2dy_scaledependent;
  double (*fddD)(double, double) = LPTorder == 1 ? &growth_ddDddy_scaledependent : &growth_ddD2ddy_scaledependent;
  g[0] = 0.0;
  for (size_t m = 1; m < nm; m++) {
    const double kmag = 2 * PI / Box * sqrt((double) m);
    if (fieldtype == FIELD_D) g[m] = fD(kmag, A);
    else if (fieldtype == FIELD_dDdy) g[m] = fdD(kmag, A);
    else if (fieldtype == FIELD_ddDddy) g[m] = fddD(kmag, A);
    else g[m] = fD(kmag, AFF) - fD(kmag, A);
  }
  *n_out = nm;
  return g;
}

/* host copy of one of the four per-particle fields, when main.c works on the host array P (initialisation, Output) */
static void sd_refresh_host(int fieldtype, int LPTorder) {
  const size_t n = NumPart;
  float *d1 = my_malloc(n * 12), *d2 = my_malloc(n * 12);
  const int first_pair = (fieldtype == FIELD_D || fieldtype == FIELD_ddDddy);
  if (first_pair) ck(mgp_download_particles(g_ctx, NULL, NULL, d1, d2, NULL), "mgp_download_particles");
  else ck(mgp_download_sd_fields(g_ctx, d1, d2), "mgp_download_sd_fields");
  for (size_t i = 0; i < n; i++)
    for (int a = 0; a < 3; a++) {
      if (first_pair) { if (LPTorder == 1) P[i].D[a] = d1[3 * i + a]; else P[i].D2[a] = d2[3 * i + a]; }
      else { if (LPTorder == 1) P[i].dDdy[a] = d1[3 * i + a]; else P[i].dD2dy[a] = d2[3 * i + a]; }
    }
  my_free(d1); my_free(d2);
}

void assign_displacment_field_to_particles(double A, double AF, double AFF, int fieldtype, int LPTorder) {
  (void) AF;
  size_t nm;
  /* MGP_SD_MERGED=1: build D + D2 (dDdy + dD2dy) in one pass when the order-2 call arrives; only their sum is ever
   * used (main.c:712, 767, 962; compute_pofk.c:324).  6 instead of 12 inverse FFTs per step, one float rounding apart. */
  static int merged = -1;
  static double *pending = NULL;
  if (merged < 0) { const char *e = getenv("MGP_SD_MERGED"); merged = e ? atoi(e) : 0; }
  double *g = sd_growth_table(A, AFF, fieldtype, LPTorder, &nm);
  if (merged) {
    if (LPTorder == 1) { if (pending) my_free(pending); pending = g; return; }
    ck(mgp_assign_displacement_fields_merged(g_ctx, fieldtype, pending, g, nm), "mgp_assign_displacement_fields_merged");
    my_free(pending); pending = NULL;
    if (g_host_is_newer || g_host_synced) { sd_refresh_host(fieldtype, 1); sd_refresh_host(fieldtype, 2); }
  } else {
    ck(mgp_assign_displacement_field(g_ctx, fieldtype, LPTorder, g, nm), "mgp_assign_displacement_field");
    if (g_host_is_newer || g_host_synced) sd_refresh_host(fieldtype, LPTorder);
  }
  my_free(g);
}
#endif

/* ------------------------------------------------------------------ P(k) file (compute_pofk.c:48-64, 239-261) */

#ifdef COMPUTE_POFK
static double k_of_bin(int i, double kmin, double kmax, int nbins, int bintype) {
  if (bintype == 0) return (kmin + (kmax - kmin) / (double) nbins * i) * 2.0 * M_PI / Box;
  return exp(log(kmin) + log(kmax / kmin) / (double) nbins * i) * 2.0 * M_PI / Box;
}

static void write_pofk_file(double a, const char *label, int nbins, const double *pofk, const double *kmean, const double *nmodes) {
  /* the sanitised binning parameters, as adjust_pofk_parameters (compute_pofk.c:758-805) leaves them */
  int bintype = pofk_bintype;
  double kmin = pofk_kmin * Box / (2.0 * M_PI), kmax = pofk_kmax * Box / (2.0 * M_PI);
  if (!(bintype == 0 || bintype == 1)) bintype = 0;
  if (kmax <= kmin) { kmin = bintype == 0 ? 0.0 : 1.0; kmax = (double) Nmesh; }
  if (kmin < 0.0) kmin = bintype == 0 ? 0.0 : 1.0;
  if (bintype == 1 && kmin == 0.0) kmin = 1.0;
  if (kmax > sqrt(3.0) * (double) Nmesh) kmax = (double) Nmesh;
  const double znow = 1.0 / a - 1.0;
  if (ThisTask != 0) return;
  char filename[1000];
  const int zint = (int) znow, zfrac = (int) ((znow - zint) * 1000);
  sprintf(filename, "%s/pofk_%s_z%d.%03d_%s.txt", OutputDir, FileBase, zint, zfrac, label);
  printf("Writing power-spectrum to file: [%s]\n", filename);
  FILE *fp = fopen(filename, "w");
  fprintf(fp, "#  k_bin (h/Mpc)        P(k) (Mpc/h)^3     k_mean_bin (h/Mpc)     Delta = k^3P(k)/2pi^2\n");
  for (int i = 1; i < nbins; i++) {
    if (nmodes[i] > 0) {
      const double kb = k_of_bin(i, kmin, kmax, nbins, bintype);
      fprintf(fp, "%10.5f   %10.5f   %10.5f   %10.5f\n", kb, pofk[i], kmean[i], pofk[i] * kb * kb * kb / 2.0 / M_PI / M_PI);
    }
  }
  fclose(fp);
}
#endif

/* compute_pofk.c:403-512 */
void compute_RSD_powerspectrum(double A, int dDdy_set_in_particles) {
#ifdef COMPUTE_POFK
  timer_start(_PofkComputation);
  if (ThisTask == 0) printf("Computing the RSD power-spectrum...\n");
  const double scaleBox = (double) Nmesh / Box;
  const double velfac = (Hubble / A);
  const double vnorm = velfac / (100.0 * A * hubble(A)) * scaleBox;           /* compute_pofk.c:291-292 */
  double dDdy = 0.0, dD2dy = 0.0;
#ifdef SCALEDEPENDENT
  float *save1 = NULL, *save2 = NULL;
  if (UseCOLA && !dDdy_set_in_particles) {                                    /* compute_pofk.c:409-428 */
    int lnx, lx0, lnp, lp0; uint64_t np;
    ck(mgp_get_layout(g_ctx, &lnx, &lx0, &lnp, &lp0, &np), "mgp_get_layout");
    save1 = my_malloc((size_t) np * 12); save2 = my_malloc((size_t) np * 12);
    ck(mgp_download_sd_fields(g_ctx, save1, save2), "mgp_download_sd_fields");
    const double As = aexp_global;
    assign_displacment_field_to_particles(As, As, As, FIELD_dDdy, LPT_ORDER_ONE);
    assign_displacment_field_to_particles(As, As, As, FIELD_dDdy, LPT_ORDER_TWO);
  }
#else
  (void) dDdy_set_in_particles;
  dDdy = growth_dDdy(A); dD2dy = growth_dD2dy(A);
#endif
  static int configured = 0;
  if (!configured) {
    mgp_pofk_config pc = {pofk_nbins, pofk_bintype, pofk_subtract_shotnoise, pofk_kmin, pofk_kmax};
    ck(mgp_set_pofk_config(g_ctx, &pc), "mgp_set_pofk_config");
    configured = 1;
  }
  const int nb = mgp_pofk_nbins(g_ctx);
  double *oy = my_malloc(sizeof(double) * 5 * nb), *oz = my_malloc(sizeof(double) * 5 * nb);
  ck(mgp_compute_rsd_power_spectrum(g_ctx, vnorm, dDdy, dD2dy, oy, oz), "mgp_compute_rsd_power_spectrum");
  if (ThisTask == 0) {                                                       /* compute_pofk.c:453-479 */
    char filename[1000];
    const double znow = 1.0 / A - 1.0;
    const int zint = (int) (znow), zfrac = (int) ((znow - zint) * 1000);
    sprintf(filename, "%s/pofk_RSD_%s_z%d.%03d.txt", OutputDir, FileBase, zint, zfrac);
    printf("Writing RSD power-spectrum to file: [%s]\n", filename);
    FILE *fp = fopen(filename, "w");
    fprintf(fp, "#  k (h/Mpc)      P0 (Mpc/h)^3      P2 (Mpc/h)^3      P4 (Mpc/h)^3     sigma0     sigma2     sigma4\n");
    for (int i = 0; i < nb; i++) {
      if (oy[i] > 0 && oy[nb + i] > 0.0) {
        const double P0 = (oy[2 * nb + i] + oz[2 * nb + i]) / 2.0, P2 = (oy[3 * nb + i] + oz[3 * nb + i]) / 2.0,
                     P4 = (oy[4 * nb + i] + oz[4 * nb + i]) / 2.0;
        fprintf(fp, "%10.5f   %10.5f   %10.5f   %10.5f   %10.5f   %10.5f   %10.5f\n", oy[nb + i], P0, P2, P4,
                fabs(oy[2 * nb + i] - oz[2 * nb + i]) / sqrt(2.0), fabs(oy[3 * nb + i] - oz[3 * nb + i]) / sqrt(2.0),
                fabs(oy[4 * nb + i] - oz[4 * nb + i]) / sqrt(2.0));
      }
    }
    fclose(fp);
  }
  my_free(oy); my_free(oz);
#ifdef SCALEDEPENDENT
  if (save1) {                                                               /* compute_pofk.c:497-508 */
    ck(mgp_upload_sd_fields(g_ctx, save1, save2), "mgp_upload_sd_fields");
    my_free(save1); my_free(save2);
  }
#endif
  timer_stop(_PofkComputation);
#else
  (void) A; (void) dDdy_set_in_particles;
#endif
}

/* ------------------------------------------------------------------ the force path (auxPM.c:37-103) */

void GetDisplacements(void) {
  if (g_host_is_newer) upload_host_particles();
  mgp_step_scalars s;
  memset(&s, 0, sizeof(s));
  s.a = aexp_global;
  s.geff = 1.0;
  if (modified_gravity_active) {
#if defined(FOFRGRAVITY) || defined(MBETAMODEL)
    /* Phi_crit(a): screening_factor_potential returns |Phi_crit / Phi_N| for deeply screened cells (udf:725-750) */
    const double big = -1e30;
    s.phi_crit = include_screening ? screening_factor_potential(aexp_global, big) * fabs(big) : 0.0;
    s.coupling = coupling_function(aexp_global);                                                   /* mg.h:79 */
    s.massterm2 = aexp_global * aexp_global * mass2_of_a(aexp_global) / pow((2.0 * PI) * INVERSE_H0_MPCH / Box, 2);   /* mg.h:80 */
#elif defined(DGPGRAVITY)
    s.coupling = coupling_function(aexp_global);
    s.dgp_fac0 = 8.0 / 9.0 * Omega * pow(rcH0_DGP / beta_DGP(aexp_global), 2);                     /* udf:762 */
    s.rsmooth = Rsmooth_global;
#elif defined(BRANSDICKE)
    s.geff = GeffoverG(aexp_global, 0.0);
#endif
  }
#ifdef MASSIVE_NEUTRINOS
  double *nutab = NULL;
  if (nu_include_massive_neutrinos) {                                         /* auxPM.c:383-420 */
    const int h = Nmesh / 2;
    const size_t nm = (size_t) 3 * h * h + 1;
    const double nufac_tmp = OmegaNu / Omega * (double) (Nmesh * Nmesh * Nmesh);
    nutab = my_malloc(sizeof(double) * nm);
    nutab[0] = 0.0;
    for (size_t m = 1; m < nm; m++) {
      const double kmag = 2.0 * PI / Box * sqrt((double) m);
      nutab[m] = nufac_tmp * get_nu_transfer_function(kmag, aexp_global) / get_cdm_baryon_transfer_function(kmag, 1.0);
    }
    s.nu_by_k2 = nutab; s.n_nu = nm; s.nu_cdmfac = (Omega - OmegaNu) / Omega;
  }
#endif
#ifdef COMPUTE_POFK
  s.compute_pofk = pofk_compute_every_step;
  static int pofk_configured = 0;
  if (!pofk_configured) {
    mgp_pofk_config pc = {pofk_nbins, pofk_bintype, pofk_subtract_shotnoise, pofk_kmin, pofk_kmax};
    ck(mgp_set_pofk_config(g_ctx, &pc), "mgp_set_pofk_config");
    pofk_configured = 1;
  }
#endif
  timer_start(_PtoMesh);
#ifdef COMPUTE_POFK
  /* the reference bins the RSD multipoles inside PtoMesh, between the CDM P(k) and the neutrino add, except in the
   * first step of an output interval (auxPM.c:374-380 and appendix B.1 of SURVEY.md: the stale loop counter makes the
   * test "timeStep_global == 0") */
  const int rsd_now = (pofk_compute_rsd_pofk == 1) && !(timeStep_global == 0);
  if (rsd_now) {
    /* the multipoles need the particles of this step in place but not the force: run the step in two halves */
    ck(mgp_move_particles(g_ctx), "mgp_move_particles");
    ck(mgp_ptomesh(g_ctx, &s), "mgp_ptomesh");
    compute_RSD_powerspectrum(aexp_global, 0);
    ck(mgp_compute_fifth_force(g_ctx, &s), "mgp_compute_fifth_force");
    ck(mgp_forces(g_ctx), "mgp_forces");
    ck(mgp_mtoparticles(g_ctx, sumDxyz), "mgp_mtoparticles");
  } else
#endif
  ck(mgp_get_displacements(g_ctx, &s, sumDxyz), "mgp_get_displacements");
  timer_stop(_PtoMesh);
#ifdef COMPUTE_POFK
  if (pofk_compute_every_step) {
    const int nb = mgp_pofk_nbins(g_ctx);
    double *p = my_malloc(sizeof(double) * nb), *k = my_malloc(sizeof(double) * nb), *n = my_malloc(sizeof(double) * nb);
    ck(mgp_get_step_power_spectrum(g_ctx, p, k, n), "mgp_get_step_power_spectrum");
    write_pofk_file(aexp_global, "CDM", nb, p, k, n);
#ifdef MASSIVE_NEUTRINOS
    if (nu_include_massive_neutrinos) {
      ck(mgp_get_step_power_spectrum_total(g_ctx, p, k, n), "mgp_get_step_power_spectrum_total");
      write_pofk_file(aexp_global, "total", nb, p, k, n);
    }
#endif
    my_free(p); my_free(k); my_free(n);
  }
#endif
#ifdef MASSIVE_NEUTRINOS
  if (nutab) my_free(nutab);
#endif
  /* main.c frees Disp[] after the kick (main.c:551, 574): hand it something to free */
  for (int j = 0; j < 3; j++) Disp[j] = my_malloc(sizeof(float));
  int lnx, lx0, lnp, lp0;
  uint64_t np;
  ck(mgp_get_layout(g_ctx, &lnx, &lx0, &lnp, &lp0, &np), "mgp_get_layout");
  NumPart = (unsigned int) np;                         /* read by main.c:452, 1076 */
}

/* the particle loops of Kick (main.c:707-739) and Drift (main.c:762-783); the scalar part stays in main.c */
void mgp_adapter_kick(double A, double dda, double ddDddy, double ddD2ddy) {
  ck(mgp_kick(g_ctx, A, dda, ddDddy, ddD2ddy, sumDxyz, sumxyz), "mgp_kick");
}

void mgp_adapter_drift(double dyyy, double deltaD, double deltaD2) {
  ck(mgp_drift(g_ctx, dyyy, deltaD, deltaD2, sumxyz), "mgp_drift");
}

#ifdef MATCHMAKER_HALOFINDER
/* ------------------------------------------------------------------ FoF halos on the fly (mm_main.c, mm_fof.c)
 * MatchMaker() keeps its name and its argument (main.c:830-872 is untouched): the parameter block of mm_main.c:144-240 is
 * filled as before because the reference's writers (mm_snap_io.c, mm_msg.c: compiled unmodified) read it; the particle
 * translation, the strip exchange, the friends-of-friends search and the halo properties (picola_to_matchmaker_particles,
 * fof_get_halos) run in the library. */
#include "mm_common.h"

void MatchMaker(struct PicolaToMatchMakerData data) {
  static int initialised = 0;
  if (!initialised) {                                  /* mm_mpi_init: write_halos_root gathers FoFHalo records (mm_snap_io.c:335) */
    int bc[1] = {(int) sizeof(FoFHalo)};
    MPI_Aint off[1] = {0};
    MPI_Datatype ty[1] = {MPI_BYTE};
    MPI_Type_struct(1, bc, off, ty, &HaloMPI);
    MPI_Type_commit(&HaloMPI);
    mm_msg_init();
    initialised = 1;
  }
  mm_msg_printf("\n=========================\n");
  mm_msg_printf("Powering up MatchMaker (CUDA library)\n");
  mm_msg_printf("=========================\n\n");
  Param.output_format = data.output_format; Param.output_pernode = data.output_pernode;
  Param.dx_extra = data.dx_extra; Param.np_min = data.np_min; Param.b_fof = data.b_fof;
  Param.n_part_1d = data.n_part_1d;
  Param.n_part = (lint) data.n_part_1d * (lint) data.n_part_1d * (lint) data.n_part_1d;
  Param.boxsize = data.boxsize; Param.omega_m = data.omega_m; Param.omega_l = data.omega_l; Param.redshift = data.redshift;
  Param.norm_vel = data.norm_vel; Param.norm_pos = data.norm_pos; Param.h = data.HubbleParam;
  Param.NumPart = NumPart; Param.Local_p_start = data.Local_p_start; Param.P = NULL;
  Param.growth_dDdy = data.growth_dDdy; Param.growth_dD2dy = data.growth_dD2dy;
  sprintf(Param.OutputDir, "%s", data.OutputDir);
  sprintf(Param.FileBase, "%s", data.FileBase);
  Param.mp = data.mass_part;
  {
    const double z = Param.redshift;
    const int zint = (int) z, zfrac = (int) ((z - zint) * 1000);
    sprintf(Param.output_prefix, "%s/matchmaker_%s_z%d.%03d", Param.OutputDir, Param.FileBase, zint, zfrac);
  }
  MPI_Comm_rank(MPI_COMM_WORLD, &(Param.i_node));
  MPI_Comm_size(MPI_COMM_WORLD, &(Param.n_nodes));
  Param.i_node_left = (Param.i_node - 1 + Param.n_nodes) % Param.n_nodes;
  Param.i_node_right = (Param.i_node + 1) % Param.n_nodes;
  if (Param.dx_extra >= 0.5 * Param.boxsize / (double) Param.n_nodes) {             /* mm_main.c:214-218 */
    mm_msg_printf("Buffer size might be too big! MatchMaker will not be run!\nReduce it or the number of cores!\n");
    return;
  }
  if (g_host_is_newer) upload_host_particles();
  mgp_fof_config fc;
  memset(&fc, 0, sizeof(fc));
  fc.norm_pos = data.norm_pos; fc.norm_vel = data.norm_vel; fc.boxsize = data.boxsize; fc.dx_extra = data.dx_extra;
  fc.b_fof = data.b_fof; fc.np_min = data.np_min; fc.mass_part = data.mass_part;
#ifndef SCALEDEPENDENT
  {
    const double A = 1.0 / (1.0 + Param.redshift);                                 /* mm_main.c:300-303 */
    fc.dDdy = data.growth_dDdy(A); fc.dD2dy = data.growth_dD2dy(A);
  }
#endif
  uint64_t nh = 0;
  mm_msg_printf("Getting halos\n");
  ck(mgp_fof_find(g_ctx, &fc, &nh), "mgp_fof_find");
  if (sizeof(FoFHalo) != sizeof(mgp_fof_halo)) FatalError((char *) "MatchMaker: FoFHalo and mgp_fof_halo differ");
  FoFHalo *fh = nh ? my_malloc(nh * sizeof(FoFHalo)) : NULL;                        /* write_halos frees it (mm_snap_io.c:300, 337) */
  ck(mgp_fof_get(g_ctx, (mgp_fof_halo *) fh), "mgp_fof_get");
  mm_msg_printf("Writing output\n");
  write_halos((lint) nh, fh);
  mm_msg_printf("=========================\n\n");
}
#endif

#ifdef LIGHTCONE
/* called (through adapter/lightcone_c.patch) from Drift_Lightcone in place of its particle loop (lightcone.c:392-471): the
 * scalars, the exit-time tables and flag_replicates stay in lightcone.c, the rows come back in the layout Output_Lightcone
 * (lightcone.c:482-565, unchanged) writes to the replicate files */
void mgp_adapter_drift_lightcone(double A, double AFF, double dyyy, double da1, double da2, double dv1, double dv2,
                                 double Rcomov_old, double Rcomov_new, double boundary, double lengthfac, double velfac_times_fac,
                                 int ntab, const double *AL_tab, const double *da1_tab, const double *da2_tab, const double *dyyy_tab) {
  static uint64_t cap = 4096;
  if (g_host_is_newer) upload_host_particles();
  const int nall = (Nrep_neg_x + Nrep_pos_x + 1) * (Nrep_neg_y + Nrep_pos_y + 1) * (Nrep_neg_z + Nrep_pos_z + 1);
  int *ijk = my_malloc(sizeof(int) * 3 * (nall > 0 ? nall : 1)), *coords = my_malloc(sizeof(int) * (nall > 0 ? nall : 1));
  int nrep = 0;
  for (int i = -Nrep_neg_x; i <= Nrep_pos_x; i++)                              /* the order of lightcone.c:411-413 */
    for (int j = -Nrep_neg_y; j <= Nrep_pos_y; j++)
      for (int k = -Nrep_neg_z; k <= Nrep_pos_z; k++) {
        const int coord = ((i + Nrep_neg_max[0]) * (Nrep_neg_max[1] + Nrep_pos_max[1] + 1) + (j + Nrep_neg_max[1])) *
                              (Nrep_neg_max[2] + Nrep_pos_max[2] + 1) + (k + Nrep_neg_max[2]);
        if (repflag[coord] == 0) { ijk[3 * nrep] = i; ijk[3 * nrep + 1] = j; ijk[3 * nrep + 2] = k; coords[nrep] = coord; nrep++; }
      }
  mgp_lightcone_step ls;
  memset(&ls, 0, sizeof(ls));
  ls.A = A; ls.AFF = AFF; ls.dyyy = dyyy; ls.da1 = da1; ls.da2 = da2; ls.dv1 = dv1; ls.dv2 = dv2;
  for (int a = 0; a < 3; a++) ls.sumxyz[a] = sumxyz[a];
  ls.rcomov_old = Rcomov_old; ls.rcomov_new = Rcomov_new;
  ls.origin[0] = Origin_x; ls.origin[1] = Origin_y; ls.origin[2] = Origin_z;
  ls.boundary = boundary; ls.lengthfac = lengthfac; ls.velfac_times_fac = velfac_times_fac;
  ls.ntab = ntab; ls.al_tab = AL_tab; ls.da1_tab = da1_tab; ls.da2_tab = da2_tab; ls.dyyy_tab = dyyy_tab;
  ls.nrep = nrep; ls.rep_ijk = ijk;
  uint64_t *count = my_malloc(sizeof(uint64_t) * (nrep > 0 ? nrep : 1));
  float *block = NULL;
  int pinned = 1;
  for (;;) {
    const size_t bytes = sizeof(float) * 6 * cap * (size_t) (nrep > 0 ? nrep : 1);
    pinned = 1;
    block = mgp_alloc_host(bytes);                       /* page-locked: the rows arrive at the full PCIe rate */
    if (!block) { pinned = 0; block = malloc(bytes); }
    if (!block) FatalError((char *) "mgp_adapter_drift_lightcone: out of host memory for the lightcone block");
    const int rc = mgp_drift_lightcone(g_ctx, &ls, cap, block, count);
    if (rc == MGP_OK) break;
    if (pinned) mgp_free_host(block); else free(block);
    if (rc != MGP_ERR_BUFFER) ck(rc, "mgp_drift_lightcone");
    uint64_t most = 0;                                                         /* nothing moved: size the block and repeat */
    for (int r = 0; r < nrep; r++) if (count[r] > most) most = count[r];
    cap = most + most / 8 + 16;
  }
  unsigned int *pc = (unsigned int *) calloc(nall > 0 ? nall : 1, sizeof(unsigned int));
  for (int r = 0; r < nrep; r++) Noutput[coords[r]] += (unsigned int) count[r];
  /* Output_Lightcone indexes the block with 32-bit arithmetic (lightcone.c:563): hand it slices that stay below 2^32 floats */
  const uint64_t lim = 0xffffffffull / (6ull * (uint64_t) (nrep > 0 ? nrep : 1));
  if (cap <= lim) {
    for (int r = 0; r < nrep; r++) pc[r] = (unsigned int) count[r];
    Output_Lightcone(pc, (unsigned int) cap, block);
  } else {
    float *slice = malloc(sizeof(float) * 6 * lim * (size_t) nrep);
    if (!slice) FatalError((char *) "mgp_adapter_drift_lightcone: out of host memory for the output slice");
    for (uint64_t done = 0;; done += lim) {
      int any = 0;
      for (int r = 0; r < nrep; r++) {
        const uint64_t left = count[r] > done ? count[r] - done : 0, take = left < lim ? left : lim;
        pc[r] = (unsigned int) take;
        if (take) { memcpy(slice + 6 * lim * (size_t) r, block + 6 * (cap * (size_t) r + done), sizeof(float) * 6 * take); any = 1; }
      }
      if (!any) break;
      Output_Lightcone(pc, (unsigned int) lim, slice);
    }
    free(slice);
  }
  free(pc);
  if (pinned) mgp_free_host(block); else free(block);
  my_free(count); my_free(ijk); my_free(coords);
}
#endif

void mgp_adapter_finish(void) {
  if (g_ctx) mgp_destroy(g_ctx);
  g_ctx = NULL;
}
