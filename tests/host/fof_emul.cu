// Host emulation of the GPU halo finder (csrc/fof.cu): runs the very sequence of steps (csrc/fof_impl.cuh::find_halos)
// and the very functors the kernels consist of (csrc/fof_impl.cuh, csrc/fof.cuh) on the CPU -- a step is a loop over its
// indices, scans and stable sorts are the standard library's, and the neighbour exchanges of P emulated tasks (one thread
// each) go through shared memory behind a barrier.  Built with nvcc as a shared library and driven from tests/test_fof.py
// (no kernel launch).
#include <pthread.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "fof_impl.cuh"

namespace mgp {
namespace fof {
inline void check(bool ok, const char *what) { if (!ok) throw std::runtime_error(what); }
}  // namespace fof
}  // namespace mgp

using namespace mgp;

namespace {

struct Shared {                     // what the emulated tasks see of each other
  int P;
  pthread_barrier_t bar;
  std::vector<unsigned long long> pair;              // [P][2]
  std::vector<const float *> x, v;                   // strip sources
  std::vector<size_t> n_strip, stride;
  std::vector<const unsigned char *> flags;          // flag sources
  std::vector<size_t> n_flags;
};

struct HostBackend {
  Shared &sh;
  int rank;
  std::vector<void *> owned;
  HostBackend(Shared &s, int r) : sh(s), rank(r) {}
  ~HostBackend() { for (void *q : owned) free(q); }
  template <class T> T *alloc(size_t n) { void *q = malloc((n ? n : 1) * sizeof(T)); owned.push_back(q); return (T *) q; }
  void zero(void *p, size_t bytes) { memset(p, 0, bytes); }
  template <class F> void run(size_t n, F f, int = 256) { for (size_t i = 0; i < n; i++) f(i); }
  template <class F> void run_warp(size_t n, F f) { for (size_t i = 0; i < n; i++) f(i); }
  size_t max_cells(size_t n) {
    if (const char *e = getenv("MGP_FOF_MAX_CELLS")) return (size_t) strtoull(e, nullptr, 10);
    return 160 * n + (1u << 20);
  }
  void scan(unsigned *p, size_t n) { unsigned s = 0; for (size_t i = 0; i < n; i++) { const unsigned t = p[i]; p[i] = s; s += t; } }
  void sort(unsigned *k0, unsigned *k1, unsigned *v0, unsigned *v1, size_t n, int bits) {
    const unsigned mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    std::vector<size_t> o(n);
    std::iota(o.begin(), o.end(), (size_t) 0);
    std::stable_sort(o.begin(), o.end(), [&](size_t a, size_t b) { return (k0[a] & mask) < (k0[b] & mask); });
    for (size_t i = 0; i < n; i++) { k1[i] = k0[o[i]]; v1[i] = v0[o[i]]; }
  }
  void download(void *host, const void *dev, size_t bytes) { memcpy(host, dev, bytes); }
  unsigned read32(const unsigned *p) { return *p; }
  unsigned long long read64(const unsigned long long *p) { return *p; }
  void gather2(const unsigned long long mine[2], unsigned long long *all) {
    sh.pair[2 * rank] = mine[0]; sh.pair[2 * rank + 1] = mine[1];
    pthread_barrier_wait(&sh.bar);
    memcpy(all, sh.pair.data(), sizeof(unsigned long long) * 2 * sh.P);
    pthread_barrier_wait(&sh.bar);
  }
  void strip_exchange(float *x, float *v, size_t N, size_t n_dom, size_t n_toleft, size_t n_buf) {
    const int right = (rank + 1) % sh.P;
    sh.x[rank] = x; sh.v[rank] = v; sh.n_strip[rank] = n_toleft; sh.stride[rank] = N;
    pthread_barrier_wait(&sh.bar);
    if (sh.n_strip[right] != n_buf) throw std::runtime_error("strip sizes disagree");
    for (int a = 0; a < 3; a++) {                    // the right neighbour's first n_buf particles (it may be myself)
      memmove(x + a * N + n_dom, sh.x[right] + a * sh.stride[right], n_buf * sizeof(float));
      memmove(v + a * N + n_dom, sh.v[right] + a * sh.stride[right], n_buf * sizeof(float));
    }
    pthread_barrier_wait(&sh.bar);
  }
  void flag_exchange(const unsigned char *send, size_t n_buf, unsigned char *recv, size_t n_toleft) {
    const int left = (rank - 1 + sh.P) % sh.P;
    sh.flags[rank] = send; sh.n_flags[rank] = n_buf;
    pthread_barrier_wait(&sh.bar);
    if (sh.n_flags[left] != n_toleft) throw std::runtime_error("flag counts disagree");
    memcpy(recv, sh.flags[left], n_toleft);
    pthread_barrier_wait(&sh.bar);
  }
};

struct Job {
  Shared *sh; int rank;
  fof::Store st; size_t n_dom; fof::Task tk; mgp_fof_config cfg;
  std::vector<mgp_fof_halo> halos;
  std::string error;
};

void *worker(void *arg) {
  Job &j = *(Job *) arg;
  try {
    HostBackend be(*j.sh, j.rank);
    fof::find_halos(be, j.st, j.n_dom, j.tk, j.cfg, j.halos);
  } catch (const std::exception &e) {
    j.error = e.what();
    // a task that fails before an exchange would leave the others waiting: this harness is for green paths
    abort();
  }
  return nullptr;
}

}  // namespace

// P tasks; task t holds n[t] particles: pos / vel / D / D2 as [n][3] floats (for scale_dependent D, D2 are P.dDdy, P.dD2dy).
// Output: n_halos[t] and the halos of all tasks one after the other in out (at most out_cap records).  Returns 0, or
// 1 when out is too small.
extern "C" int fof_emul(int P, const long *n, const float *const *pos, const float *const *vel, const float *const *D,
                        const float *const *D2, const int *p_start, const double *slab_fraction, int nsample, int use_cola,
                        int scale_dependent, const mgp_fof_config *cfg, long *n_halos, mgp_fof_halo *out, long out_cap) {
  Shared sh;
  sh.P = P;
  pthread_barrier_init(&sh.bar, nullptr, P);
  sh.pair.resize(2 * P); sh.x.resize(P); sh.v.resize(P); sh.n_strip.resize(P); sh.stride.resize(P); sh.flags.resize(P); sh.n_flags.resize(P);
  std::vector<Job> jobs(P);
  std::vector<std::vector<float4>> pA(P), pB(P), pC(P);
  std::vector<std::vector<float2>> pE(P);
  std::vector<std::vector<float>> f1(P), f2(P);
  for (int t = 0; t < P; t++) {
    const size_t m = (size_t) n[t];
    pA[t].resize(m ? m : 1); pB[t].resize(m ? m : 1); pC[t].resize(m ? m : 1); pE[t].resize(m ? m : 1);
    f1[t].resize(3 * (m ? m : 1)); f2[t].resize(3 * (m ? m : 1));
    for (size_t i = 0; i < m; i++) {
      pA[t][i] = make_float4(pos[t][3 * i], pos[t][3 * i + 1], pos[t][3 * i + 2], 0.f);
      pB[t][i] = make_float4(vel[t][3 * i], vel[t][3 * i + 1], vel[t][3 * i + 2], 0.f);
      pC[t][i] = make_float4(D[t][3 * i], D[t][3 * i + 1], D[t][3 * i + 2], D2[t][3 * i]);
      pE[t][i] = make_float2(D2[t][3 * i + 1], D2[t][3 * i + 2]);
      for (int a = 0; a < 3; a++) { f1[t][a * m + i] = D[t][3 * i + a]; f2[t][a * m + i] = D2[t][3 * i + a]; }
    }
    Job &j = jobs[t];
    j.sh = &sh; j.rank = t; j.n_dom = m; j.cfg = *cfg;
    j.st.pA = pA[t].data(); j.st.pB = pB[t].data(); j.st.pC = pC[t].data(); j.st.pE = pE[t].data();
    j.st.f1 = f1[t].data(); j.st.f2 = f2[t].data(); j.st.cap = m; j.st.scale_dependent = scale_dependent; j.st.use_cola = use_cola;
    j.tk.rank = t; j.tk.P = P; j.tk.nsample = nsample; j.tk.p_start = p_start[t]; j.tk.slab_fraction = slab_fraction[t];
  }
  std::vector<pthread_t> th(P);
  for (int t = 0; t < P; t++) pthread_create(&th[t], nullptr, worker, &jobs[t]);
  for (int t = 0; t < P; t++) pthread_join(th[t], nullptr);
  pthread_barrier_destroy(&sh.bar);
  long at = 0;
  for (int t = 0; t < P; t++) {
    n_halos[t] = (long) jobs[t].halos.size();
    if (at + n_halos[t] > out_cap) return 1;
    if (n_halos[t]) memcpy(out + at, jobs[t].halos.data(), sizeof(mgp_fof_halo) * (size_t) n_halos[t]);
    at += n_halos[t];
  }
  return 0;
}
