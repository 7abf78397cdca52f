// Particle store and the streaming particle kernels: cell keys + sort + permute, Kick, Drift.
//
// Reference semantics reproduced (all citations relative to the reference's src/):
//   Kick   main.c:707-739   Disp -= <Disp>; Vel += (-1.5 Omega Disp - UseCOLA (D ddD + D2 ddD2)/A) dda
//   Drift  main.c:762-783   Pos += (Vel - <Vel>) dyyy; Pos = wrap(Pos + UseCOLA (D dD + D2 dD2))
//   wrap   auxPM.c:649-655  float arithmetic
// The reference keeps particle data in float (MEMORY_MODE) and evaluates these expressions in
// double without FMA contraction (gcc, x86-64); the kernels below use __dmul_rn/__dadd_rn in the
// same association order so that the stored floats are bit-identical.
#include "common.cuh"
#include "reduce.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace mgp {

// ------------------------------------------------------------------ allocation

void reduce_alloc(Ctx &c, size_t n) {
  if (n <= c.red_cap) return;
  if (c.d_red) CK(cudaFree(c.d_red));
  if (c.h_red) CK(cudaFreeHost(c.h_red));
  CK(cudaMalloc(&c.d_red, n * sizeof(double)));
  CK(cudaMallocHost(&c.h_red, n * sizeof(double)));
  c.red_cap = n;
}

void particles_alloc(Ctx &c) {
  const size_t cap = c.cap;
  CK(cudaMalloc(&c.pA, cap * sizeof(float4)));
  CK(cudaMalloc(&c.pB, cap * sizeof(float4)));
  CK(cudaMalloc(&c.pC, cap * sizeof(float4)));
  CK(cudaMalloc(&c.pE, cap * sizeof(float2)));
  CK(cudaMalloc(&c.pA2, cap * sizeof(float4)));
  CK(cudaMalloc(&c.pB2, cap * sizeof(float4)));
  CK(cudaMalloc(&c.pC2, cap * sizeof(float4)));
  CK(cudaMalloc(&c.pE2, cap * sizeof(float2)));
  CK(cudaMalloc(&c.disp, 3 * cap * sizeof(float)));
  for (int i = 0; i < 2; i++) {
    CK(cudaMalloc(&c.key[i], cap * sizeof(uint32_t)));
    CK(cudaMalloc(&c.perm[i], cap * sizeof(uint32_t)));
  }
  CK(cudaMalloc(&c.row_start, ((size_t) (c.nx + 1) * c.N + 2) * sizeof(uint32_t)));
  c.cub_temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, c.cub_temp_bytes, c.key[0], c.key[1], c.perm[0], c.perm[1],
                                  (int64_t) cap, 0, 32, c.stream);
  CK(cudaMalloc(&c.cub_temp, c.cub_temp_bytes));
  reduce_alloc(c, 8192);
  rows_alloc(c);
  CK(cudaMalloc(&c.d_flag, 16 * sizeof(int)));
  CK(cudaMallocHost(&c.h_flag, 16 * sizeof(int)));

  // bucket sort (see particles_sort): buckets of 2^bucket_zshift consecutive z-cells of one (x, y) row
  c.bucket_zshift = (c.N % 8 == 0) ? 3 : 0;
  c.nbuckets = (size_t) c.nx * c.N * (size_t) (c.N >> c.bucket_zshift);
  CK(cudaMalloc(&c.bucket_start, (c.nbuckets + 2) * sizeof(uint32_t)));    // + the "no longer mine" bucket + end
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, c.bucket_start, c.bucket_start, (int64_t) (c.nbuckets + 2), c.stream);
  if (scan_bytes > c.cub_temp_bytes) {
    CK(cudaFree(c.cub_temp));
    c.cub_temp_bytes = scan_bytes;
    CK(cudaMalloc(&c.cub_temp, c.cub_temp_bytes));
  }

  // key = row-major cell index of the local slab
  const uint64_t cells = (uint64_t) c.nx * c.N * c.N;
  int b = 0;
  while ((1ull << b) <= cells) b++;      // keys 0 .. cells (cells itself = "no longer mine")
  REQUIRE(b <= 32, MGP_ERR_INVALID, "2^32 or more mesh cells on one rank: use more ranks");
  c.key_zshift = 0;
  c.key_bits = b < 1 ? 1 : b;
}

void particles_free(Ctx &c) {
  cudaFree(c.pA); cudaFree(c.pB); cudaFree(c.pC); cudaFree(c.pE); cudaFree(c.disp);
  cudaFree(c.pA2); cudaFree(c.pB2); cudaFree(c.pC2); cudaFree(c.pE2);
  for (int i = 0; i < 2; i++) { cudaFree(c.key[i]); cudaFree(c.perm[i]); }
  rows_free(c);
  cudaFree(c.row_start); cudaFree(c.cub_temp); cudaFree(c.bucket_start); cudaFree(c.stage);
  for (int i = 0; i < 2; i++) if (c.stage_ev[i]) cudaEventDestroy(c.stage_ev[i]);
  if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
  cudaFree(c.d_red); if (c.h_red) cudaFreeHost(c.h_red);
  cudaFree(c.d_flag); if (c.h_flag) cudaFreeHost(c.h_flag);
}

// ------------------------------------------------------------------ upload / download

__global__ void k_pack(size_t n, size_t off, const float *__restrict__ pos, const float *__restrict__ vel,
                       const float *__restrict__ D, const float *__restrict__ D2,
                       const unsigned long long *__restrict__ id, unsigned long long id_base,
                       float4 *pA, float4 *pB, float4 *pC, float2 *pE) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    unsigned long long pid = id ? id[i] : id_base + i;
    float d2x = D2 ? D2[3 * i] : 0.f, d2y = D2 ? D2[3 * i + 1] : 0.f, d2z = D2 ? D2[3 * i + 2] : 0.f;
    float vx = vel ? vel[3 * i] : 0.f, vy = vel ? vel[3 * i + 1] : 0.f, vz = vel ? vel[3 * i + 2] : 0.f;
    float dx = D ? D[3 * i] : 0.f, dy = D ? D[3 * i + 1] : 0.f, dz = D ? D[3 * i + 2] : 0.f;
    pA[off + i] = make_float4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], __uint_as_float((unsigned) (pid & 0xffffffffull)));
    pB[off + i] = make_float4(vx, vy, vz, __uint_as_float((unsigned) (pid >> 32)));
    pC[off + i] = make_float4(dx, dy, dz, d2x);
    pE[off + i] = make_float2(d2y, d2z);
  }
}

__global__ void k_unpack(size_t n, size_t off, float *pos, float *vel, float *D, float *D2, unsigned long long *id,
                         const float4 *__restrict__ pA, const float4 *__restrict__ pB,
                         const float4 *__restrict__ pC, const float2 *__restrict__ pE) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    float4 a = pA[off + i], b = pB[off + i], cc = pC[off + i];
    float2 e = pE[off + i];
    if (pos) { pos[3 * i] = a.x; pos[3 * i + 1] = a.y; pos[3 * i + 2] = a.z; }
    if (vel) { vel[3 * i] = b.x; vel[3 * i + 1] = b.y; vel[3 * i + 2] = b.z; }
    if (D) { D[3 * i] = cc.x; D[3 * i + 1] = cc.y; D[3 * i + 2] = cc.z; }
    if (D2) { D2[3 * i] = cc.w; D2[3 * i + 1] = e.x; D2[3 * i + 2] = e.y; }
    if (id) id[i] = ((unsigned long long) __float_as_uint(b.w) << 32) | (unsigned long long) __float_as_uint(a.w);
  }
}

static const size_t kChunk = (size_t) 1 << 22;   // particles per staging chunk (two chunks in flight)

// Persistent double-buffered staging area: host arrays [n][3] <-> device SoA records.  With pinned
// host memory every copy is asynchronous, chunk c+1's copy overlaps chunk c's (un)pack kernel and the
// host blocks once, at the end.
static float *stage_buffer(Ctx &c, size_t ch) {
  const size_t need = 2 * ch * (12 * sizeof(float) + sizeof(uint64_t));
  if (c.stage_bytes < need) {
    if (c.stage) CK(cudaFree(c.stage));
    CK(cudaMalloc(&c.stage, need));
    c.stage_bytes = need;
    for (int i = 0; i < 2; i++)
      if (!c.stage_ev[i]) CK(cudaEventCreateWithFlags(&c.stage_ev[i], cudaEventDisableTiming));
    if (!c.copy_stream) CK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
  }
  return (float *) c.stage;
}

void particles_upload(Ctx &c, uint64_t n, const float *pos, const float *vel, const float *D, const float *D2,
                      const uint64_t *id) {
  REQUIRE(n <= c.cap, MGP_ERR_BUFFER, "mgp_upload_particles: more particles than ceil(NumPart*Buffer); increase Buffer");
  REQUIRE(pos != nullptr || n == 0, MGP_ERR_INVALID, "mgp_upload_particles: pos is NULL");
  const size_t ch = n < kChunk ? (n ? n : 1) : kChunk;
  float *base = stage_buffer(c, ch);
  const size_t rec = ch * (12 * sizeof(float) + sizeof(uint64_t));
  cudaEvent_t packed[2] = {c.stage_ev[0], c.stage_ev[1]};
  cudaEvent_t copied;
  CK(cudaEventCreateWithFlags(&copied, cudaEventDisableTiming));
  CK(cudaStreamSynchronize(c.stream));
  int it = 0;
  for (size_t off = 0; off < n; off += ch, it++) {
    const int b = it & 1;
    float *st = (float *) ((char *) base + (size_t) b * rec);
    float *s_pos = st, *s_vel = st + 3 * ch, *s_D = st + 6 * ch, *s_D2 = st + 9 * ch;
    unsigned long long *s_id = (unsigned long long *) (st + 12 * ch);
    const size_t m = (n - off) < ch ? (n - off) : ch;
    if (it >= 2) CK(cudaStreamWaitEvent(c.copy_stream, packed[b], 0));     // buffer b free again
    CK(cudaMemcpyAsync(s_pos, pos + 3 * off, m * 12, cudaMemcpyHostToDevice, c.copy_stream));
    if (vel) CK(cudaMemcpyAsync(s_vel, vel + 3 * off, m * 12, cudaMemcpyHostToDevice, c.copy_stream));
    if (D) CK(cudaMemcpyAsync(s_D, D + 3 * off, m * 12, cudaMemcpyHostToDevice, c.copy_stream));
    if (D2) CK(cudaMemcpyAsync(s_D2, D2 + 3 * off, m * 12, cudaMemcpyHostToDevice, c.copy_stream));
    if (id) CK(cudaMemcpyAsync(s_id, id + off, m * 8, cudaMemcpyHostToDevice, c.copy_stream));
    CK(cudaEventRecord(copied, c.copy_stream));
    CK(cudaStreamWaitEvent(c.stream, copied, 0));
    k_pack<<<grid_for(m, 256), 256, 0, c.stream>>>(m, off, s_pos, vel ? s_vel : nullptr, D ? s_D : nullptr,
                                                   D2 ? s_D2 : nullptr, id ? s_id : nullptr, off, c.pA, c.pB, c.pC,
                                                   (float2 *) c.pE);
    CK(cudaEventRecord(packed[b], c.stream));
    c.launches++;
  }
  CK(cudaStreamSynchronize(c.stream));
  CK(cudaEventDestroy(copied));
  c.np = n;
  c.sorted = false;
  c.bins_valid = false;
  c.drifts_since_sort = 1 << 30;
  c.np_after_sort = SIZE_MAX;
  c.have_disp = false;
  c.sd_req_valid = false;
  c.sd_lagrangian_only = false;
  for (int s = 0; s < 4; s++) c.sd_set[s] = false;
}

void particles_download(Ctx &c, float *pos, float *vel, float *D, float *D2, uint64_t *id) {
  const size_t n = c.np;
  if (!n) return;
  const size_t ch = n < kChunk ? n : kChunk;
  float *base = stage_buffer(c, ch);
  const size_t rec = ch * (12 * sizeof(float) + sizeof(uint64_t));
  cudaEvent_t drained[2] = {c.stage_ev[0], c.stage_ev[1]};
  cudaEvent_t unpacked;
  CK(cudaEventCreateWithFlags(&unpacked, cudaEventDisableTiming));
  int it = 0;
  for (size_t off = 0; off < n; off += ch, it++) {
    const int b = it & 1;
    float *st = (float *) ((char *) base + (size_t) b * rec);
    float *s_pos = st, *s_vel = st + 3 * ch, *s_D = st + 6 * ch, *s_D2 = st + 9 * ch;
    unsigned long long *s_id = (unsigned long long *) (st + 12 * ch);
    const size_t m = (n - off) < ch ? (n - off) : ch;
    if (it >= 2) CK(cudaStreamWaitEvent(c.stream, drained[b], 0));          // buffer b copied out
    k_unpack<<<grid_for(m, 256), 256, 0, c.stream>>>(m, off, pos ? s_pos : nullptr, vel ? s_vel : nullptr,
                                                     D ? s_D : nullptr, D2 ? s_D2 : nullptr, id ? s_id : nullptr,
                                                     c.pA, c.pB, c.pC, (const float2 *) c.pE);
    c.launches++;
    CK(cudaEventRecord(unpacked, c.stream));
    CK(cudaStreamWaitEvent(c.copy_stream, unpacked, 0));
    if (pos) CK(cudaMemcpyAsync(pos + 3 * off, s_pos, m * 12, cudaMemcpyDeviceToHost, c.copy_stream));
    if (vel) CK(cudaMemcpyAsync(vel + 3 * off, s_vel, m * 12, cudaMemcpyDeviceToHost, c.copy_stream));
    if (D) CK(cudaMemcpyAsync(D + 3 * off, s_D, m * 12, cudaMemcpyDeviceToHost, c.copy_stream));
    if (D2) CK(cudaMemcpyAsync(D2 + 3 * off, s_D2, m * 12, cudaMemcpyDeviceToHost, c.copy_stream));
    if (id) CK(cudaMemcpyAsync(id + off, s_id, m * 8, cudaMemcpyDeviceToHost, c.copy_stream));
    CK(cudaEventRecord(drained[b], c.copy_stream));
  }
  CK(cudaStreamSynchronize(c.copy_stream));
  CK(cudaStreamSynchronize(c.stream));
  CK(cudaEventDestroy(unpacked));
}


// ------------------------------------------------------------------ snapshot blocks (Output, main.c:915-997)

// The three GADGET blocks of this rank's particles exactly as Output() forms them:
//   pos[k] = (float)(lengthfac * Pos[k])                                                              (main.c:946)
//   vel[k] = (float)(velfac * fac * (Vel[k] - sumxyz[k] + (D[k] dDdy + D2[k] dD2dy) * UseCOLA))        (main.c:966-967)
//   SCALEDEPENDENT:                 ... + (P.dDdy[k] + P.dD2dy[k]) * UseCOLA   with the float sum      (main.c:962-963)
//   id                                                                                                (main.c:985)
// double arithmetic without contraction, in C's association order, so that the floats are the reference's.
template <int SD>
__global__ void __launch_bounds__(256)
k_snapshot(size_t n, size_t off, const float4 *__restrict__ pA, const float4 *__restrict__ pB, const float4 *__restrict__ pC,
           const float2 *__restrict__ pE, const float *__restrict__ f1, const float *__restrict__ f2, size_t cap,
           double lengthfac, double vfac, double s0, double s1, double s2, double dDdy, double dD2dy, int usecola,
           float *__restrict__ pos, float *__restrict__ vel, unsigned long long *__restrict__ id) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    const size_t j = off + i;
    const float4 a = pA[j], b = pB[j];
    pos[3 * i] = (float) __dmul_rn(lengthfac, (double) a.x);
    pos[3 * i + 1] = (float) __dmul_rn(lengthfac, (double) a.y);
    pos[3 * i + 2] = (float) __dmul_rn(lengthfac, (double) a.z);
    double l0, l1, l2;
    if (SD) {
      const float g0 = f2 ? f2[j] : 0.0f, g1 = f2 ? f2[cap + j] : 0.0f, g2 = f2 ? f2[2 * cap + j] : 0.0f;
      l0 = (double) __fmul_rn(__fadd_rn(f1[j], g0), (float) usecola);
      l1 = (double) __fmul_rn(__fadd_rn(f1[cap + j], g1), (float) usecola);
      l2 = (double) __fmul_rn(__fadd_rn(f1[2 * cap + j], g2), (float) usecola);
    } else {
      const float4 d = pC[j];
      const float2 e = pE[j];
      l0 = __dmul_rn(__dadd_rn(__dmul_rn((double) d.x, dDdy), __dmul_rn((double) d.w, dD2dy)), (double) usecola);
      l1 = __dmul_rn(__dadd_rn(__dmul_rn((double) d.y, dDdy), __dmul_rn((double) e.x, dD2dy)), (double) usecola);
      l2 = __dmul_rn(__dadd_rn(__dmul_rn((double) d.z, dDdy), __dmul_rn((double) e.y, dD2dy)), (double) usecola);
    }
    vel[3 * i] = (float) __dmul_rn(vfac, __dadd_rn(__dsub_rn((double) b.x, s0), l0));
    vel[3 * i + 1] = (float) __dmul_rn(vfac, __dadd_rn(__dsub_rn((double) b.y, s1), l1));
    vel[3 * i + 2] = (float) __dmul_rn(vfac, __dadd_rn(__dsub_rn((double) b.z, s2), l2));
    id[i] = ((unsigned long long) __float_as_uint(b.w) << 32) | (unsigned long long) __float_as_uint(a.w);
  }
}

// Packs chunk by chunk into the staging area; chunk c + 1 is packed while chunk c crosses PCIe (asynchronous when the host
// buffers are pinned, e.g. from mgp_alloc_host).
void particles_snapshot(Ctx &c, double lengthfac, double vfac, const double sumxyz[3], double dDdy, double dD2dy, float *pos,
                        float *vel, uint64_t *id) {
  const size_t n = c.np;
  if (!n) return;
  const size_t ch = n < kChunk ? n : kChunk;
  float *base = stage_buffer(c, ch);
  const size_t rec = ch * (12 * sizeof(float) + sizeof(uint64_t));
  cudaEvent_t drained[2] = {c.stage_ev[0], c.stage_ev[1]};
  cudaEvent_t packed;
  CK(cudaEventCreateWithFlags(&packed, cudaEventDisableTiming));
  const bool sd = c.cfg.scale_dependent != 0;
  int it = 0;
  for (size_t off = 0; off < n; off += ch, it++) {
    const int b = it & 1;
    float *st = (float *) ((char *) base + (size_t) b * rec);
    float *s_pos = st, *s_vel = st + 3 * ch;
    unsigned long long *s_id = (unsigned long long *) (st + 12 * ch);
    const size_t m = (n - off) < ch ? (n - off) : ch;
    if (it >= 2) CK(cudaStreamWaitEvent(c.stream, drained[b], 0));          // buffer b copied out
    if (sd)
      k_snapshot<1><<<grid_for(m, 256), 256, 0, c.stream>>>(m, off, c.pA, c.pB, nullptr, nullptr, c.sdf[2],
                                                           c.sd_zero[3] ? nullptr : c.sdf[3], c.cap, lengthfac, vfac, sumxyz[0],
                                                           sumxyz[1], sumxyz[2], dDdy, dD2dy, c.cfg.use_cola, s_pos, s_vel, s_id);
    else
      k_snapshot<0><<<grid_for(m, 256), 256, 0, c.stream>>>(m, off, c.pA, c.pB, c.pC, (const float2 *) c.pE, nullptr, nullptr, c.cap,
                                                           lengthfac, vfac, sumxyz[0], sumxyz[1], sumxyz[2], dDdy, dD2dy,
                                                           c.cfg.use_cola, s_pos, s_vel, s_id);
    c.launches++;
    CK(cudaEventRecord(packed, c.stream));
    CK(cudaStreamWaitEvent(c.copy_stream, packed, 0));
    CK(cudaMemcpyAsync(pos + 3 * off, s_pos, m * 12, cudaMemcpyDeviceToHost, c.copy_stream));
    CK(cudaMemcpyAsync(vel + 3 * off, s_vel, m * 12, cudaMemcpyDeviceToHost, c.copy_stream));
    CK(cudaMemcpyAsync(id + off, s_id, m * 8, cudaMemcpyDeviceToHost, c.copy_stream));
    CK(cudaEventRecord(drained[b], c.copy_stream));
  }
  CK(cudaStreamSynchronize(c.copy_stream));
  CK(cudaStreamSynchronize(c.stream));
  CK(cudaEventDestroy(packed));
}

// ------------------------------------------------------------------ [3][cap] device arrays <-> [n][3] host arrays

__global__ void k_soa3_to_aos(size_t n, size_t off, const float *__restrict__ soa, size_t cap, float m0, float m1, float m2,
                              double d0, double d1, double d2, int sub, float *__restrict__ aos) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    float x = soa[off + i], y = soa[cap + off + i], z = soa[2 * cap + off + i];
    if (sub) {           // value - mean evaluated in double, stored float (2LPT.c:1501-1508)
      x = (float) ((double) x - d0); y = (float) ((double) y - d1); z = (float) ((double) z - d2);
    }
    aos[3 * i] = x; aos[3 * i + 1] = y; aos[3 * i + 2] = z;
  }
}

__global__ void k_aos_to_soa3(size_t n, size_t off, const float *__restrict__ aos, float *__restrict__ soa, size_t cap) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    soa[off + i] = aos[3 * i]; soa[cap + off + i] = aos[3 * i + 1]; soa[2 * cap + off + i] = aos[3 * i + 2];
  }
}

// The interleaving runs on the device, chunk by chunk through the staging area; the host side is one contiguous copy
// per chunk (no per-element host loop, no pageable temporary).
void copy_soa3(Ctx &c, float *dev_soa, size_t n, float *host_aos, bool to_host, const double *sub_mean) {
  if (!n) return;
  const size_t ch = n < kChunk ? n : kChunk;
  float *st = stage_buffer(c, ch);
  for (size_t off = 0; off < n; off += ch) {
    const size_t m = (n - off) < ch ? (n - off) : ch;
    if (to_host) {
      k_soa3_to_aos<<<grid_for(m, 256), 256, 0, c.stream>>>(m, off, dev_soa, c.cap, 0.f, 0.f, 0.f, sub_mean ? sub_mean[0] : 0.0,
                                                           sub_mean ? sub_mean[1] : 0.0, sub_mean ? sub_mean[2] : 0.0,
                                                           sub_mean != nullptr, st);
      CK(cudaMemcpyAsync(host_aos + 3 * off, st, m * 12, cudaMemcpyDeviceToHost, c.stream));
    } else {
      CK(cudaMemcpyAsync(st, host_aos + 3 * off, m * 12, cudaMemcpyHostToDevice, c.stream));
      k_aos_to_soa3<<<grid_for(m, 256), 256, 0, c.stream>>>(m, off, st, dev_soa, c.cap);
    }
    c.launches++;
    CK(cudaStreamSynchronize(c.stream));       // the staging chunk is reused
  }
}

// ------------------------------------------------------------------ cell keys, sort, permute

// key = ((ix - x0) * N + iy) * N + iz; cell of a position exactly as PtoMesh computes it
// (auxPM.c:298-322): X = (double)Pos * (Nmesh/Box); I = (unsigned)X; I >= Nmesh -> 0 for y, z.
__global__ void k_keys(size_t n, const float4 *__restrict__ pA, double scale, int N, int x0, int nx, int drop_foreign,
                       uint32_t dead_key, uint32_t *__restrict__ key, uint32_t *__restrict__ perm) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    const float4 a = pA[i];
    unsigned ix = (unsigned) ((double) a.x * scale);
    unsigned iy = (unsigned) ((double) a.y * scale);
    unsigned iz = (unsigned) ((double) a.z * scale);
    if (iy >= (unsigned) N) iy = 0;
    if (iz >= (unsigned) N) iz = 0;
    if (ix >= (unsigned) N) ix = N - 1;
    int lx = (int) ix - x0;
    if (drop_foreign && (lx < 0 || lx >= nx)) {
      key[i] = dead_key;          // sent away by MoveParticles: sorts behind every live particle
    } else {
      if (lx < 0) lx = 0;
      if (lx >= nx) lx = nx - 1;
      key[i] = ((uint32_t) lx * (uint32_t) N + iy) * (uint32_t) N + iz;
    }
    perm[i] = (uint32_t) i;
  }
}

template <typename V>
__global__ void k_permute(size_t n, const uint32_t *__restrict__ perm, const V *__restrict__ in, V *__restrict__ out) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    out[i] = in[perm[i]];
}

// row_start[r] = first sorted particle whose row (= key / zc) is >= r, for r in [0, nrows]
__global__ void k_row_start(size_t n, const uint32_t *__restrict__ key, unsigned zc, unsigned nrows,
                            uint32_t *__restrict__ row_start) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i <= n; i += (size_t) gridDim.x * blockDim.x) {
    long long r_prev = (i == 0) ? -1 : (long long) (key[i - 1] / zc);
    long long r = (i == n) ? (long long) nrows : (long long) (key[i] / zc);
    for (long long rr = r_prev + 1; rr <= r; rr++) row_start[rr] = (uint32_t) i;
  }
}

// ---- bucket (counting) sort: the default order for the ATOMIC / TILE deposits and the gather -----
//
// A bucket is a run of 2^zs consecutive z-cells of one (x, y) row (64 bytes of a double grid row).
// Pass 1 ranks every particle inside its bucket with one warp-aggregated atomic per (warp, bucket);
// an exclusive scan of the bucket counts gives the bucket offsets; pass 2 moves all particle
// fields to bucket_start[bucket] + rank in one kernel.  About 150 bytes of traffic per particle against
// about 330 for key generation + 3 radix passes + 4 permutes.  The order inside a bucket is not
// reproducible, so the DETERMINISTIC deposit keeps using the stable radix sort below.
__global__ void __launch_bounds__(256)
k_bucket_rank(size_t n, const float4 *__restrict__ pA, double scale, int N, int x0, int nx, int zs, int drop_foreign,
              unsigned dead_bucket, uint32_t *__restrict__ count, uint32_t *__restrict__ bucket_of,
              uint32_t *__restrict__ rank_of) {
  const unsigned lane = threadIdx.x & 31;
  const size_t nround = (n + 31) / 32 * 32;
  const unsigned nbz = (unsigned) N >> zs;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < nround; i += (size_t) gridDim.x * blockDim.x) {
    unsigned b = 0xffffffffu - lane;                 // idle lanes: distinct sentinels
    if (i < n) {
      const float4 a = pA[i];
      unsigned ix = (unsigned) ((double) a.x * scale);
      unsigned iy = (unsigned) ((double) a.y * scale);
      unsigned iz = (unsigned) ((double) a.z * scale);
      if (iy >= (unsigned) N) iy = 0;
      if (iz >= (unsigned) N) iz = 0;
      if (ix >= (unsigned) N) ix = N - 1;            // cannot happen for float positions in [0, Box)
      int lx = (int) ix - x0;
      if (drop_foreign && (lx < 0 || lx >= nx)) {
        b = dead_bucket;                             // sent away by MoveParticles: sorted behind the live particles
      } else {
        if (lx < 0) lx = 0;
        if (lx >= nx) lx = nx - 1;
        b = ((unsigned) lx * (unsigned) N + iy) * nbz + (iz >> zs);
      }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if ((int) lane == leader && i < n) base = atomicAdd(&count[b], (unsigned) __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < n) {
      bucket_of[i] = b;
      rank_of[i] = base + (unsigned) __popc(peers & ((1u << lane) - 1u));
    }
  }
}

// src_of[bucket_start[b] + rank] = i  (4-byte scatter; the 56-byte payload is then GATHERED, which
// keeps every payload store full-width and coalesced)
__global__ void __launch_bounds__(256)
k_bucket_invert(size_t n, const uint32_t *__restrict__ bucket_of, const uint32_t *__restrict__ rank_of,
                const uint32_t *__restrict__ start, uint32_t *__restrict__ src_of) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    src_of[(size_t) start[bucket_of[i]] + rank_of[i]] = (uint32_t) i;
}

__global__ void __launch_bounds__(256)
k_permute_all(size_t n, const uint32_t *__restrict__ src_of, const float4 *__restrict__ iA, const float4 *__restrict__ iB,
              const float4 *__restrict__ iC, const float2 *__restrict__ iE, float4 *__restrict__ oA,
              float4 *__restrict__ oB, float4 *__restrict__ oC, float2 *__restrict__ oE) {
  for (size_t d = blockIdx.x * (size_t) blockDim.x + threadIdx.x; d < n; d += (size_t) gridDim.x * blockDim.x) {
    const uint32_t i = src_of[d];
    const float4 a = iA[i], b = iB[i];
    oA[d] = a; oB[d] = b;
    if (iC) { const float4 cc = iC[i]; const float2 e = iE[i]; oC[d] = cc; oE[d] = e; }   // unused when scale_dependent
  }
}

// row_start[r] = bucket_start[r * nbz] for r in [0, nrows]
__global__ void k_rows_from_buckets(unsigned nrows, unsigned nbz, const uint32_t *__restrict__ start, uint32_t *__restrict__ row_start) {
  for (size_t r = blockIdx.x * (size_t) blockDim.x + threadIdx.x; r <= nrows; r += (size_t) gridDim.x * blockDim.x)
    row_start[r] = start[(size_t) r * nbz];
}

static void bucket_sort(Ctx &c) {
  const size_t n = c.np;
  const unsigned nrows = (unsigned) (c.nx * c.N);
  const unsigned nbz = (unsigned) c.N >> c.bucket_zshift;
  const double scale = (double) c.N / c.cfg.box;
  REQUIRE(c.pA2 != nullptr, MGP_ERR_STATE, "bucket sort buffers missing");
  CK(cudaMemsetAsync(c.bucket_start, 0, (c.nbuckets + 2) * sizeof(uint32_t), c.stream));
  k_bucket_rank<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.pA, scale, c.N, c.x0, c.nx, c.bucket_zshift, c.P > 1,
                                                      (unsigned) c.nbuckets, c.bucket_start, c.key[0], c.perm[0]);
  size_t tb = c.cub_temp_bytes;
  CK(cub::DeviceScan::ExclusiveSum(c.cub_temp, tb, c.bucket_start, c.bucket_start, (int64_t) (c.nbuckets + 2), c.stream));
  k_bucket_invert<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.key[0], c.perm[0], c.bucket_start, c.perm[1]);
  // scale_dependent: P.D / P.D2 are per-step temporaries kept in the sd field arrays, pC / pE are unused
  k_permute_all<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.perm[1], c.pA, c.pB, c.cfg.scale_dependent ? nullptr : c.pC,
                                                      (const float2 *) c.pE, c.pA2, c.pB2, c.pC2, (float2 *) c.pE2);
  std::swap(c.pA, c.pA2); std::swap(c.pB, c.pB2); std::swap(c.pC, c.pC2); std::swap(c.pE, c.pE2);
  k_rows_from_buckets<<<grid_for((size_t) nrows + 1, 256), 256, 0, c.stream>>>(nrows, nbz, c.bucket_start, c.row_start);
  c.launches += 4 + 2;   // + CUB's scan (2 launches)
}

static void radix_sort(Ctx &c) {
  const size_t n = c.np;
  const unsigned zc = (unsigned) c.N;
  const unsigned nrows = (unsigned) (c.nx * c.N);
  const double scale = (double) c.N / c.cfg.box;
  k_keys<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.pA, scale, c.N, c.x0, c.nx, c.P > 1,
                                                 (uint32_t) ((uint64_t) c.nx * c.N * c.N), c.key[0], c.perm[0]);
  size_t tb = c.cub_temp_bytes;
  CK(cub::DeviceRadixSort::SortPairs(c.cub_temp, tb, c.key[0], c.key[1], c.perm[0], c.perm[1], (int64_t) n, 0,
                                     c.key_bits, c.stream));
  // out-of-place gather into the second buffer set, then swap names
  k_permute_all<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.perm[1], c.pA, c.pB, c.pC, (const float2 *) c.pE, c.pA2, c.pB2,
                                                      c.pC2, (float2 *) c.pE2);
  std::swap(c.pA, c.pA2); std::swap(c.pB, c.pB2); std::swap(c.pC, c.pC2); std::swap(c.pE, c.pE2);
  const size_t nlive = c.np_after_sort != SIZE_MAX ? c.np_after_sort : n;
  k_row_start<<<grid_for(nlive + 1, 256), 256, 0, c.stream>>>(nlive, c.key[1], zc, nrows, c.row_start);
  c.launches += 3 + 4;   // + CUB's histogram / onesweep passes (>= 4 launches for <= 32 bits)
}

void particles_sort(Ctx &c) {
  PhaseTimer t(c, PH_SORT);
  const unsigned nrows = (unsigned) (c.nx * c.N);
  if (c.np == 0) {
    CK(cudaMemsetAsync(c.row_start, 0, ((size_t) nrows + 1) * sizeof(uint32_t), c.stream));
  } else if (c.cfg.deposit_mode == MGP_DEPOSIT_DETERMINISTIC) {
    radix_sort(c);
    c.exact_cell_order = true;
  } else {
    bucket_sort(c);
    c.exact_cell_order = false;
  }
  if (c.np_after_sort != SIZE_MAX) { c.np = c.np_after_sort; c.np_after_sort = SIZE_MAX; }   // leavers dropped
  c.sorted = true;
  c.bins_valid = false;    // the records moved
  c.drifts_since_sort = 0;
  c.have_disp = false;   // Disp was in the old order
  c.sd_req_valid = false;                                  // ... and so were the scale-dependent per-particle fields
  for (int s = 0; s < 4; s++) c.sd_set[s] = false;
}

// ------------------------------------------------------------------ Kick

__global__ void __launch_bounds__(256)
k_kick(size_t n, float4 *__restrict__ pB, const float4 *__restrict__ pC, const float2 *__restrict__ pE,
       float *__restrict__ disp, size_t cap, double sDx, double sDy, double sDz, double m15omega, double usecola,
       double ddD, double ddD2, double A, double dda, double *__restrict__ partial) {
  double sx = 0, sy = 0, sz = 0;
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    float4 v = pB[i];
    const float4 d = pC[i];
    const float2 e = pE[i];
    // Disp[axes][n] -= sumDxyz[axes]   (float storage, double arithmetic)
    const float gx = (float) __dsub_rn((double) disp[i], sDx);
    const float gy = (float) __dsub_rn((double) disp[cap + i], sDy);
    const float gz = (float) __dsub_rn((double) disp[2 * cap + i], sDz);
    disp[i] = gx; disp[cap + i] = gy; disp[2 * cap + i] = gz;
    // force = -1.5*Omega*Disp - UseCOLA*(D*ddDddy + D2*ddD2ddy)/A
    const double fx = __dsub_rn(__dmul_rn(m15omega, (double) gx),
                                __ddiv_rn(__dmul_rn(usecola, __dadd_rn(__dmul_rn((double) d.x, ddD), __dmul_rn((double) d.w, ddD2))), A));
    const double fy = __dsub_rn(__dmul_rn(m15omega, (double) gy),
                                __ddiv_rn(__dmul_rn(usecola, __dadd_rn(__dmul_rn((double) d.y, ddD), __dmul_rn((double) e.x, ddD2))), A));
    const double fz = __dsub_rn(__dmul_rn(m15omega, (double) gz),
                                __ddiv_rn(__dmul_rn(usecola, __dadd_rn(__dmul_rn((double) d.z, ddD), __dmul_rn((double) e.y, ddD2))), A));
    // Vel += force * dda
    v.x = (float) __dadd_rn((double) v.x, __dmul_rn(fx, dda));
    v.y = (float) __dadd_rn((double) v.y, __dmul_rn(fy, dda));
    v.z = (float) __dadd_rn((double) v.z, __dmul_rn(fz, dda));
    pB[i] = v;
    sx += (double) v.x; sy += (double) v.y; sz += (double) v.z;
  }
  block_sum3(sx, sy, sz);
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x] = sx; partial[3 * blockIdx.x + 1] = sy; partial[3 * blockIdx.x + 2] = sz;
  }
}

void particles_kick(Ctx &c, double A, double dda, double ddD, double ddD2, const double sumD[3], double sumV[3]) {
  PhaseTimer t(c, PH_KICK);
  REQUIRE(c.have_disp, MGP_ERR_STATE, "mgp_kick: no displacements; call mgp_get_displacements first");
  const size_t n = c.np;
  const unsigned g = grid_for(n, 256, 8);
  reduce_alloc(c, (size_t) g * 3 + 16);
  double *res = c.d_red + (size_t) g * 3;
  k_kick<<<g, 256, 0, c.stream>>>(n, c.pB, c.pC, (const float2 *) c.pE, c.disp, c.cap, sumD[0], sumD[1], sumD[2],
                                  -1.5 * c.cfg.omega, (double) c.cfg.use_cola, ddD, ddD2, A, dda, c.d_red);
  k_final_reduce<<<1, 256, 0, c.stream>>>(c.d_red, (int) g, 3, 3, 1.0, res);
  c.launches += 2;
  allreduce_sum(c, res, 3);
  CK(cudaMemcpyAsync(c.h_red, res, 3 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  const double tot = (double) c.cfg.nsample * (double) c.cfg.nsample * (double) c.cfg.nsample;
  for (int a = 0; a < 3; a++) sumV[a] = c.h_red[a] / tot;   // sumxyz /= TotNumPart (main.c:738)
}

// ------------------------------------------------------------------ Drift

__device__ inline float periodic_wrap_f(float x, float box) {
  while (x >= box) x -= box;
  while (x < 0) x += box;
  if (x == box) x = 0.0f;
  return x;
}

__global__ void __launch_bounds__(256)
k_drift(size_t n, float4 *__restrict__ pA, const float4 *__restrict__ pB, const float4 *__restrict__ pC,
        const float2 *__restrict__ pE, double sVx, double sVy, double sVz, double dyyy, double usecola, double dD,
        double dD2, float boxf) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
    float4 p = pA[i];
    const float4 v = pB[i];
    const float4 d = pC[i];
    const float2 e = pE[i];
    // Pos += (Vel - sumxyz) * dyyy            (stored back to float)
    float x = (float) __dadd_rn((double) p.x, __dmul_rn(__dsub_rn((double) v.x, sVx), dyyy));
    float y = (float) __dadd_rn((double) p.y, __dmul_rn(__dsub_rn((double) v.y, sVy), dyyy));
    float z = (float) __dadd_rn((double) p.z, __dmul_rn(__dsub_rn((double) v.z, sVz), dyyy));
    // Pos = periodic_wrap(Pos + UseCOLA*(D*deltaD + D2*deltaD2))   (argument converted to float)
    x = periodic_wrap_f((float) __dadd_rn((double) x, __dmul_rn(usecola, __dadd_rn(__dmul_rn((double) d.x, dD), __dmul_rn((double) d.w, dD2)))), boxf);
    y = periodic_wrap_f((float) __dadd_rn((double) y, __dmul_rn(usecola, __dadd_rn(__dmul_rn((double) d.y, dD), __dmul_rn((double) e.x, dD2)))), boxf);
    z = periodic_wrap_f((float) __dadd_rn((double) z, __dmul_rn(usecola, __dadd_rn(__dmul_rn((double) d.z, dD), __dmul_rn((double) e.y, dD2)))), boxf);
    p.x = x; p.y = y; p.z = z;
    pA[i] = p;
  }
}

void particles_drift(Ctx &c, double dyyy, double dD, double dD2, const double sumV[3]) {
  PhaseTimer t(c, PH_DRIFT);
  const size_t n = c.np;
  if (n) {
    k_drift<<<grid_for(n, 256), 256, 0, c.stream>>>(n, c.pA, c.pB, c.pC, (const float2 *) c.pE, sumV[0], sumV[1],
                                                    sumV[2], dyyy, (double) c.cfg.use_cola, dD, dD2, (float) c.cfg.box);
    c.launches++;
  }
  particles_after_drift(c);
}

void particles_after_drift(Ctx &c) {
  c.sorted = false;
  c.bins_valid = false;
  if (c.drifts_since_sort < (1 << 30)) c.drifts_since_sort++;
  c.have_disp = false;
}

}  // namespace mgp
