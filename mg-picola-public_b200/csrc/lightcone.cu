// Drift_Lightcone (lightcone.c:265-474) on the GPU: the drift of one step, replicate by replicate, with every particle
// image that leaves the shrinking lightcone written out at its interpolated exit position.
//
// The reference loops particles x active replicates on one core and appends rows to a host block it flushes to files
// whenever the block is full.  Here: a counting pass (one atomic per leaving image, no state change), exclusive offsets
// per replicate on the host (nrep is small), then the drift pass that writes every row at offset[r] + slot into ONE
// packed device buffer of exactly sum(count) rows and advances the particle.  Bytes per particle and pass: the 56-byte
// record read (+ 16 written by the drift pass); the replicate list and the three 1000-node spline tables (48 KB) stay in L1 / L2.
// The arithmetic lives in lightcone.cuh, shared with the host emulation (tests/host/lightcone_emul.cu).
#include "common.cuh"
#include "lightcone.cuh"

namespace mgp {

// Slots of the leaving images: the lanes of a warp vote per replicate, one lane adds the warp's number to the replicate's
// counter and the lanes take consecutive slots behind the value it got back.  (One atomic per image on a few dozen
// addresses serialises in L2: 20 ns each, measured -- profiles/r02r_fof_lc.md.)  The counting pass adds into a per-CTA copy
// of the counters in shared memory first (SMEM_COUNT; nrep * 4 bytes of dynamic shared memory: 32-bit counters, a native ATOMS.ADD).
template <bool WRITE, bool SMEM_COUNT>
__global__ void __launch_bounds__(256)
k_lightcone(size_t n, float4 *__restrict__ pA, const float4 *__restrict__ pB, const float4 *__restrict__ pC,
            const float2 *__restrict__ pE, lc::Params p, unsigned long long *__restrict__ count,
            const unsigned long long *__restrict__ offset, float *__restrict__ rows, int *__restrict__ over_flag) {
  extern __shared__ unsigned int s_count[];            // 32-bit: a native shared-memory atomic (a 64-bit add is a CAS loop there);
  if (SMEM_COUNT) {                                    // a CTA never sees 2^32 particles
    for (int r = threadIdx.x; r < p.nrep; r += blockDim.x) s_count[r] = 0u;
    __syncthreads();
  }
  const unsigned lane = threadIdx.x & 31u;
  bool over = false;
  const size_t n_pad = (n + 31) & ~(size_t) 31;            // the lanes of a warp stay together: they vote in every replicate
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n_pad; i += (size_t) gridDim.x * blockDim.x) {
    const bool valid = i < n;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, d = a;
    float2 e = make_float2(0.f, 0.f);
    if (valid) { a = pA[i]; b = pB[i]; d = pC[i]; e = pE[i]; }
    lc::Particle q;
    q.pos[0] = a.x; q.pos[1] = a.y; q.pos[2] = a.z;
    q.vel[0] = b.x; q.vel[1] = b.y; q.vel[2] = b.z;
    q.d[0] = d.x; q.d[1] = d.y; q.d[2] = d.z;
    q.d2[0] = d.w; q.d2[1] = e.x; q.d2[2] = e.y;
    over |= lc::particle<WRITE>(p, q, valid, offset, rows, [&](int r, bool out) -> unsigned long long {
      const unsigned m = __ballot_sync(0xffffffffu, out);
      if (m == 0u) return 0ull;
      const int leader = __ffs(m) - 1;
      unsigned long long base = 0ull;
      if (SMEM_COUNT) {                                    // counting pass: no slot needed
        if ((int) lane == leader) atomicAdd(&s_count[r], (unsigned) __popc(m));
        return 0ull;
      }
      if ((int) lane == leader) base = atomicAdd(&count[r], (unsigned long long) __popc(m));
      if (!WRITE) return 0ull;
      base = __shfl_sync(0xffffffffu, base, leader);
      return base + (unsigned long long) __popc(m & ((1u << lane) - 1u));
    });
    if (WRITE && valid) {
      a.x = q.pos[0]; a.y = q.pos[1]; a.z = q.pos[2];
      pA[i] = a;
    }
  }
  if (over) *over_flag = 1;
  if (SMEM_COUNT) {
    __syncthreads();
    for (int r = threadIdx.x; r < p.nrep; r += blockDim.x)
      if (s_count[r]) atomicAdd(&count[r], (unsigned long long) s_count[r]);
  }
}

namespace {

// device copies of the tables and the replicate list of one call
struct LcDevice {
  double *tab = nullptr;              // al, y[3], c[3]: 7 * ntab doubles
  int *rep = nullptr;
  unsigned long long *count = nullptr, *offset = nullptr;
  float *rows = nullptr;
  ~LcDevice() {
    if (tab) cudaFree(tab);
    if (rep) cudaFree(rep);
    if (count) cudaFree(count);
    if (rows) cudaFree(rows);
  }
};

void check_step(const Ctx &c, const mgp_lightcone_step *ls) {
  REQUIRE(ls != nullptr, MGP_ERR_INVALID, "lightcone: NULL step");
  REQUIRE(!c.cfg.scale_dependent, MGP_ERR_STATE,
          "lightcone: not available with scale_dependent (the reference refuses LIGHTCONE + SCALEDEPENDENT, Makefile:279)");
  REQUIRE(ls->ntab >= 3 && ls->al_tab && ls->da1_tab && ls->da2_tab && ls->dyyy_tab, MGP_ERR_INVALID,
          "lightcone: the exit-time tables need at least 3 nodes");
  REQUIRE(ls->nrep >= 0 && (ls->nrep == 0 || ls->rep_ijk), MGP_ERR_INVALID, "lightcone: bad replicate list");
  for (int i = 1; i < ls->ntab; i++)
    REQUIRE(ls->al_tab[i] > ls->al_tab[i - 1], MGP_ERR_INVALID, "lightcone: AL_tab must increase (AFF > A)");
}

lc::Params make_params(Ctx &c, const mgp_lightcone_step *ls, LcDevice &dv) {
  const int nt = ls->ntab;
  std::vector<double> h((size_t) 7 * nt);
  const double *src[3] = {ls->da1_tab, ls->da2_tab, ls->dyyy_tab};
  memcpy(h.data(), ls->al_tab, sizeof(double) * nt);
  for (int t = 0; t < 3; t++) {
    memcpy(h.data() + (size_t) (1 + t) * nt, src[t], sizeof(double) * nt);
    lc::spline_coeffs(ls->al_tab, src[t], nt, h.data() + (size_t) (4 + t) * nt);
  }
  CK(cudaMalloc(&dv.tab, h.size() * sizeof(double)));
  CK(cudaMemcpyAsync(dv.tab, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  const int nr = ls->nrep > 0 ? ls->nrep : 1;
  CK(cudaMalloc(&dv.rep, (size_t) 3 * nr * sizeof(int)));
  if (ls->nrep) CK(cudaMemcpyAsync(dv.rep, ls->rep_ijk, (size_t) 3 * ls->nrep * sizeof(int), cudaMemcpyHostToDevice, c.stream));
  CK(cudaMalloc(&dv.count, (size_t) 2 * nr * sizeof(unsigned long long)));
  dv.offset = dv.count + nr;
  CK(cudaStreamSynchronize(c.stream));     // h goes out of scope

  lc::Params p;
  p.A = ls->A; p.AFF = ls->AFF; p.dyyy = ls->dyyy; p.da1 = ls->da1; p.da2 = ls->da2; p.dv1 = ls->dv1; p.dv2 = ls->dv2;
  for (int a = 0; a < 3; a++) { p.sV[a] = ls->sumxyz[a]; p.origin[a] = ls->origin[a]; }
  p.rc_old = ls->rcomov_old; p.rc_new = ls->rcomov_new;
  p.rc_old2 = ls->rcomov_old * ls->rcomov_old; p.rc_new2 = ls->rcomov_new * ls->rcomov_new;
  p.box = c.cfg.box; p.boxf = (float) c.cfg.box; p.boundary = ls->boundary; p.lengthfac = ls->lengthfac;
  p.vfac = ls->velfac_times_fac; p.usecola = (double) c.cfg.use_cola;
  p.ntab = nt; p.al = dv.tab;
  for (int t = 0; t < 3; t++) { p.y[t] = dv.tab + (size_t) (1 + t) * nt; p.c[t] = dv.tab + (size_t) (4 + t) * nt; }
  p.nrep = ls->nrep; p.rep = dv.rep;
  return p;
}

// counting pass; throws the reference's FatalError text when a displacement exceeds the boundary
void count_pass(Ctx &c, const lc::Params &p, LcDevice &dv, std::vector<unsigned long long> &cnt) {
  const int nr = p.nrep > 0 ? p.nrep : 1;
  cnt.assign((size_t) nr, 0ull);
  CK(cudaMemsetAsync(dv.count, 0, (size_t) 2 * nr * sizeof(unsigned long long), c.stream));
  CK(cudaMemsetAsync(c.d_flag, 0, sizeof(int), c.stream));
  if (c.np) {
    if (p.nrep <= 8192)       // 32 KB of counters per CTA at most
      k_lightcone<false, true><<<grid_for(c.np, 256), 256, (size_t) nr * sizeof(unsigned int), c.stream>>>(
          c.np, c.pA, c.pB, c.pC, (const float2 *) c.pE, p, dv.count, dv.offset, nullptr, c.d_flag);
    else
      k_lightcone<false, false><<<grid_for(c.np, 256), 256, 0, c.stream>>>(c.np, c.pA, c.pB, c.pC, (const float2 *) c.pE, p,
                                                                          dv.count, dv.offset, nullptr, c.d_flag);
    CK(cudaGetLastError());
    c.launches++;
  }
  CK(cudaMemcpyAsync(cnt.data(), dv.count, (size_t) nr * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaMemcpyAsync(c.h_flag, c.d_flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  REQUIRE(*c.h_flag == 0, MGP_ERR_INVALID,
          "lightcone.c Particle displacement greater than boundary, increase boundary condition");
}

}  // namespace

void lightcone_count(Ctx &c, const mgp_lightcone_step *ls, uint64_t *count) {
  check_step(c, ls);
  REQUIRE(count != nullptr || ls->nrep == 0, MGP_ERR_INVALID, "mgp_lightcone_count: NULL count");
  LcDevice dv;
  const lc::Params p = make_params(c, ls, dv);
  std::vector<unsigned long long> cnt;
  count_pass(c, p, dv, cnt);
  for (int r = 0; r < ls->nrep; r++) count[r] = cnt[r];
}

void lightcone_drift(Ctx &c, const mgp_lightcone_step *ls, uint64_t cap, float *block, uint64_t *count) {
  PhaseTimer t(c, PH_DRIFT);
  check_step(c, ls);
  REQUIRE(count != nullptr || ls->nrep == 0, MGP_ERR_INVALID, "mgp_drift_lightcone: NULL count");
  LcDevice dv;
  const lc::Params p = make_params(c, ls, dv);
  std::vector<unsigned long long> cnt;
  count_pass(c, p, dv, cnt);
  const int nr = ls->nrep;
  std::vector<unsigned long long> off((size_t) (nr > 0 ? nr : 1), 0ull);
  unsigned long long total = 0, most = 0;
  for (int r = 0; r < nr; r++) {
    off[r] = total; total += cnt[r];
    if (cnt[r] > most) most = cnt[r];
    count[r] = cnt[r];
  }
  REQUIRE(most <= cap, MGP_ERR_BUFFER, "mgp_drift_lightcone: a replicate needs " + std::to_string(most) +
                                            " rows, the block holds " + std::to_string(cap) + " per replicate");
  REQUIRE(total == 0 || block != nullptr, MGP_ERR_INVALID, "mgp_drift_lightcone: NULL block");
  if (total) CK(cudaMalloc(&dv.rows, (size_t) total * 6 * sizeof(float)));
  CK(cudaMemcpyAsync(dv.offset, off.data(), off.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, c.stream));
  CK(cudaMemsetAsync(dv.count, 0, off.size() * sizeof(unsigned long long), c.stream));
  if (c.np) {
    k_lightcone<true, false><<<grid_for(c.np, 256), 256, 0, c.stream>>>(c.np, c.pA, c.pB, c.pC, (const float2 *) c.pE, p,
                                                                       dv.count, dv.offset, dv.rows, c.d_flag);
    CK(cudaGetLastError());
    c.launches++;
  }
  // every replicate's rows straight into its part of the caller's block (full PCIe rate when the block is pinned: mgp_alloc_host)
  for (int r = 0; r < nr; r++)
    if (cnt[r])
      CK(cudaMemcpyAsync(block + (size_t) r * cap * 6, dv.rows + (size_t) off[r] * 6, (size_t) cnt[r] * 6 * sizeof(float),
                         cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  particles_after_drift(c);
}

}  // namespace mgp
