// Per-mode arithmetic of the 2LPT pipeline in k-space (displacement_fields, 2LPT.c:417-421, 1224-1268, 1310-1355, 1618-1626),
// shared by k_ic_kernel (ic.cu) and by the host emulation under tests/host/ (the same function run mode by mode on the
// CPU against the unmodified reference's ZA / LPT arrays).
//
// The wave-vector conventions are the reference's, Nyquist planes included:
//   * displacement_fields and from_cdisp_store_to_ZA count an index below N/2 up and everything else down, so the Nyquist
//     index N/2 is -N/2 (2LPT.c:1229-1245, 1316-1332).  The generated initial conditions have no power on the Nyquist planes
//     (2LPT.c:361), so for them nothing depends on this.
//   * AssignDisplacementField (READICFROMFILE; readICfromfile.c:717, 733) counts an index ABOVE N/2 down: the Nyquist index
//     is +N/2 in psi = i k / k^2 delta, while the gradients and the second-order displacement keep -N/2.  The density of
//     external particles does have power there (CIC-deposited, then deconvolved), so `ext` = 1 reproduces that.
#pragma once

#if defined(__CUDACC__)
#define ICM_HD __host__ __device__ __forceinline__
#else
#define ICM_HD inline
#endif

namespace mgp {

// MODE 0: psi_a          = (-kp_a/k^2 * d.im,  kp_a/k^2 * d.re)                        (2LPT.c:417-421; readICfromfile.c:744-745)
// MODE 1: psi_a,a        = (-psi_a.im * kv_a,  psi_a.re * kv_a), a = 0,1,2              (2LPT.c:1252-1268)
// MODE 2: psi_0,1 psi_0,2 psi_1,2
// MODE 3: psi2_a         = ( s.im * kv_a / k^2, -s.re * kv_a / k^2)                      (2LPT.c:1348-1355)
// MODE 4: scale-dependent fields: (-d.im * kv_a/k^2 * G[m], d.re * kv_a/k^2 * G[m])      (2LPT.c:1618-1626)
// kv: the Nyquist index counts as -N/2; kp = kv except with ext, where it counts as +N/2.
template <typename T, typename C, int MODE>
ICM_HD void ic_mode(int N, int i, int j, int k, double box, const C s, const double *gtab, double norm, int ext, C out[3]) {
  const int h = N / 2;
  const double PI = 3.14159265358979323846;
  double kv[3];
  kv[0] = (i < h ? i : i - N) * 2 * PI / box;
  kv[1] = (j < h ? j : j - N) * 2 * PI / box;
  kv[2] = (k < h ? k : k - N) * 2 * PI / box;
  double kp[3] = {kv[0], kv[1], kv[2]};
  if (ext) {
    if (i == h) kp[0] = -kv[0];
    if (j == h) kp[1] = -kv[1];
    if (k == h) kp[2] = -kv[2];
  }
  const double kmag2 = kv[0] * kv[0] + kv[1] * kv[1] + kv[2] * kv[2];
  if (!(kmag2 > 0.0)) {
    out[0].x = out[0].y = out[1].x = out[1].y = out[2].x = out[2].y = (T) 0;
  } else if (MODE == 0) {
    for (int a = 0; a < 3; a++) { out[a].x = (T) (-kp[a] / kmag2 * (double) s.y); out[a].y = (T) (kp[a] / kmag2 * (double) s.x); }
  } else if (MODE == 1 || MODE == 2) {
    T pre[3], pim[3];
    for (int a = 0; a < 3; a++) { pre[a] = (T) (-kp[a] / kmag2 * (double) s.y); pim[a] = (T) (kp[a] / kmag2 * (double) s.x); }
    const int A[3] = {0, MODE == 1 ? 1 : 0, MODE == 1 ? 2 : 1}, B[3] = {MODE == 1 ? 0 : 1, MODE == 1 ? 1 : 2, 2};
    for (int q = 0; q < 3; q++) { out[q].x = (T) (-(double) pim[A[q]] * kv[B[q]]); out[q].y = (T) ((double) pre[A[q]] * kv[B[q]]); }
  } else if (MODE == 3) {
    for (int a = 0; a < 3; a++) { out[a].x = (T) ((double) s.y * kv[a] / kmag2); out[a].y = (T) (-(double) s.x * kv[a] / kmag2); }
  } else {
    const int d0 = i < h ? i : i - N, d1 = j < h ? j : j - N, d2 = k < h ? k : k - N;
    const double g = norm * gtab[(long long) d0 * d0 + (long long) d1 * d1 + (long long) d2 * d2];
    for (int a = 0; a < 3; a++) { out[a].x = (T) (-(double) s.y * kv[a] / kmag2 * g); out[a].y = (T) ((double) s.x * kv[a] / kmag2 * g); }
  }
}

}  // namespace mgp
