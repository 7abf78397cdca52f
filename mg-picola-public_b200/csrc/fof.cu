// FoF halo finder on the fly (MatchMaker: mm_main.c:129-385, mm_fof.c) on the GPU.
//
// The reference copies every particle into a 56-byte record, qsorts the records by x, trades the strip x <= dx_extra with
// the neighbouring tasks, chains friends with a breadth-first search over per-cell id lists on one core, and then walks
// every halo's members for its properties.  Here the particle store stays where it is:
//   1. x keys of the slab's particles, radix sort (the reference's order: it fixes who is "first" in every later sum);
//   2. positions / velocities in MatchMaker's units gathered in that order ([3][N] floats each); the strip goes to the left
//      neighbour over NCCL (a device copy on one rank);
//   3. counting sort into search cells of one mean inter-particle distance, positions packed per cell;
//   4. union-find over the 27 neighbouring cells: the friendship test in the reference's float arithmetic, the larger root
//      hooked under the smaller by atomicCAS -- the root of a group is its smallest index whatever the thread schedule;
//   5. group sizes, the "does the left neighbour own it" flags back over NCCL, members of groups with >= np_min particles
//      compacted by a scan (index order kept) and grouped by halo with a stable radix sort;
//   6. one thread per halo walks its members in the reference's order: the double sums are the reference's to the bit.
// HBM traffic per particle: ~56 B of records read once, ~60 B of keys / cells / union-find arrays written and read, and per
// pair test 16 B from L1 / L2 (a cell's packed positions are contiguous).
// The sequence of steps and the functors are in fof_impl.cuh, the arithmetic in fof.cuh; both are shared with the host
// emulation (tests/host/fof_emul.cu), which runs them on the CPU for several emulated tasks.  This file is the GPU back end.
#include "common.cuh"
#include "fof_impl.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace mgp {

namespace fof {
inline void check(bool ok, const char *what) { REQUIRE(ok, MGP_ERR_INVALID, what); }
}  // namespace fof

namespace {

template <class F>
__global__ void k_fof_step(size_t n, F f) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) f(i);
}

template <class F>
__global__ void k_fof_warp_step(size_t n, F f) {      // one warp per index, every lane calls f(i)
  const size_t w = (blockIdx.x * (size_t) blockDim.x + threadIdx.x) >> 5, nw = ((size_t) gridDim.x * blockDim.x) >> 5;
  for (size_t i = w; i < n; i += nw) f(i);
}

struct DeviceBackend {
  Ctx &c;
  // work arrays are carved out of a few large device allocations (about thirty arrays per call; cudaMalloc and cudaFree
  // of each would cost more than most of the steps)
  std::vector<void *> owned;
  char *cur = nullptr;
  size_t left = 0, chunk;
  explicit DeviceBackend(Ctx &c_) : c(c_), chunk(std::max((size_t) 32 << 20, (size_t) 48 * (size_t) c_.np)) {}
  ~DeviceBackend() { for (void *q : owned) cudaFree(q); }

  template <class T> T *alloc(size_t n) {
    const size_t bytes = (((n ? n : 1) * sizeof(T)) + 255) & ~(size_t) 255;
    if (bytes > left) {
      void *q = nullptr;
      const size_t sz = std::max(bytes, chunk);
      CK(cudaMalloc(&q, sz));
      owned.push_back(q);
      cur = (char *) q;
      left = sz;
    }
    T *r = (T *) cur;
    cur += bytes;
    left -= bytes;
    return r;
  }
  void zero(void *p, size_t bytes) { if (bytes) CK(cudaMemsetAsync(p, 0, bytes, c.stream)); }
  template <class F> void run(size_t n, F f, int block = 256) {
    if (!n) return;
    k_fof_step<F><<<grid_for(n, block, 64), block, 0, c.stream>>>(n, f);
    CK(cudaGetLastError());
    c.launches++;
  }
  template <class F> void run_warp(size_t n, F f) {
    if (!n) return;
    k_fof_warp_step<F><<<grid_for(n * 32, 32 * fof::FOF_WARPS_PER_CTA, 64), 32 * fof::FOF_WARPS_PER_CTA, 0, c.stream>>>(n, f);
    CK(cudaGetLastError());
    c.launches++;
  }
  // search cells for n particles: up to 160 per particle (Dfof = 0.2 mean distances needs 125), within a quarter of the
  // memory that is free once the ~96 bytes per particle of the search itself are set aside
  size_t max_cells(size_t n) {
    if (const char *e = getenv("MGP_FOF_MAX_CELLS")) return (size_t) strtoull(e, nullptr, 10);
    size_t fr = 0, tot = 0;
    CK(cudaMemGetInfo(&fr, &tot));
    const size_t need = 96 * n, avail = fr > need ? fr - need : 0;
    return std::max(n / 8 + 1, std::min(avail / 16, 160 * n + (1u << 20)));
  }
  void scan(unsigned *p, size_t n) {
    size_t tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, p, p, (int64_t) n, c.stream));
    void *tmp = alloc<char>(tb);
    CK(cub::DeviceScan::ExclusiveSum(tmp, tb, p, p, (int64_t) n, c.stream));
    c.launches += 2;
  }
  void sort(unsigned *k0, unsigned *k1, unsigned *v0, unsigned *v1, size_t n, int bits) {
    if (!n) return;
    size_t tb = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, v1, (int64_t) n, 0, bits, c.stream));
    void *tmp = alloc<char>(tb);
    CK(cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, v1, (int64_t) n, 0, bits, c.stream));
    c.launches += 4;
  }
  void download(void *host, const void *dev, size_t bytes) {
    if (bytes) CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
  }
  unsigned read32(const unsigned *p) { unsigned h = 0; download(&h, p, sizeof(h)); return h; }
  unsigned long long read64(const unsigned long long *p) { unsigned long long h = 0; download(&h, p, sizeof(h)); return h; }

  // {strip count, first particle plane} of every rank
  void gather2(const unsigned long long mine[2], unsigned long long *all) {
    if (c.P == 1) { all[0] = mine[0]; all[1] = mine[1]; return; }
    unsigned long long *d = alloc<unsigned long long>(2 + 2 * (size_t) c.P);
    CK(cudaMemcpyAsync(d, mine, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, c.stream));
    CKNCCL(ncclAllGather(d, d + 2, 2, ncclUint64, c.comm, c.stream));
    download(all, d + 2, 2 * (size_t) c.P * sizeof(unsigned long long));
  }
  // my first n_toleft particles to the left neighbour, the right neighbour's strip behind my n_dom particles
  void strip_exchange(float *x, float *v, size_t N, size_t n_dom, size_t n_toleft, size_t n_buf) {
    const int right = (c.rank + 1) % c.P, left = (c.rank - 1 + c.P) % c.P;
    if (c.P == 1) {
      for (int a = 0; a < 3 && n_buf; a++) {
        CK(cudaMemcpyAsync(x + a * N + n_dom, x + a * N, n_buf * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
        CK(cudaMemcpyAsync(v + a * N + n_dom, v + a * N, n_buf * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
      }
      return;
    }
    CKNCCL(ncclGroupStart());
    for (int a = 0; a < 3; a++) {
      if (n_toleft) {
        CKNCCL(ncclSend(x + a * N, n_toleft, ncclFloat, left, c.comm, c.stream));
        CKNCCL(ncclSend(v + a * N, n_toleft, ncclFloat, left, c.comm, c.stream));
      }
      if (n_buf) {
        CKNCCL(ncclRecv(x + a * N + n_dom, n_buf, ncclFloat, right, c.comm, c.stream));
        CKNCCL(ncclRecv(v + a * N + n_dom, n_buf, ncclFloat, right, c.comm, c.stream));
      }
    }
    CKNCCL(ncclGroupEnd());
  }
  // the flags of the particles I received go back to where they came from (mm_fof.c:417-423)
  void flag_exchange(const unsigned char *send, size_t n_buf, unsigned char *recv, size_t n_toleft) {
    const int right = (c.rank + 1) % c.P, left = (c.rank - 1 + c.P) % c.P;
    if (c.P == 1) {
      if (n_buf) CK(cudaMemcpyAsync(recv, send, n_buf, cudaMemcpyDeviceToDevice, c.stream));
      return;
    }
    CKNCCL(ncclGroupStart());
    if (n_buf) CKNCCL(ncclSend(send, n_buf, ncclUint8, right, c.comm, c.stream));
    if (n_toleft) CKNCCL(ncclRecv(recv, n_toleft, ncclUint8, left, c.comm, c.stream));
    CKNCCL(ncclGroupEnd());
  }
};

}  // namespace

void fof_find(Ctx &c, const mgp_fof_config *cfg) {
  REQUIRE(cfg != nullptr, MGP_ERR_INVALID, "mgp_fof_find: NULL configuration");
  REQUIRE(cfg->boxsize > 0 && cfg->b_fof > 0 && cfg->np_min >= 1 && cfg->dx_extra >= 0 && cfg->norm_pos > 0, MGP_ERR_INVALID,
          "mgp_fof_find: bad configuration");
  // the reference does not run the finder when the strip is half a slab or more (mm_main.c:214-218)
  REQUIRE(cfg->dx_extra < 0.5 * cfg->boxsize / (double) c.P, MGP_ERR_INVALID,
          "mgp_fof_find: dx_extra must be less than half a slab (mm_main.c:214)");
  REQUIRE(c.np < 0xfffffff0ull, MGP_ERR_INVALID, "mgp_fof_find: more than 2^32 particles on one rank");
  const bool sd = c.cfg.scale_dependent != 0;
  if (sd && c.cfg.use_cola) {
    sd_materialise(c, 1);
    REQUIRE(c.sd_set[2], MGP_ERR_STATE, "mgp_fof_find (scale-dependent): assign FIELD_dDdy first (main.c:824-832)");
  }
  c.fof_halos.clear();
  c.fof_valid = false;
  fof::Store st;
  st.pA = c.pA; st.pB = c.pB; st.pC = sd ? nullptr : c.pC; st.pE = sd ? nullptr : (const float2 *) c.pE;
  st.f1 = sd ? c.sdf[2] : nullptr; st.f2 = (sd && !c.sd_zero[3]) ? c.sdf[3] : nullptr;
  st.cap = c.cap; st.scale_dependent = sd ? 1 : 0; st.use_cola = c.cfg.use_cola;
  fof::Task tk;
  tk.rank = c.rank; tk.P = c.P; tk.nsample = c.cfg.nsample; tk.p_start = c.p0; tk.slab_fraction = (double) c.nx / (double) c.N;
  DeviceBackend be(c);
  fof::find_halos(be, st, (size_t) c.np, tk, *cfg, c.fof_halos);
  CK(cudaStreamSynchronize(c.stream));
  c.fof_valid = true;
}

}  // namespace mgp
