// Host emulation of k_ic_kernel (csrc/ic.cu): the per-mode arithmetic of the 2LPT pipeline (csrc/ic_modes.cuh) applied to a
// whole half-spectrum [N][N][N/2+1] on the CPU.  Built with nvcc as a shared library and driven from
// tests/test_readic_oracle.py, which strings the modes together with numpy transforms exactly as ic_generate_t does and
// compares the displacements with the unmodified reference's ZA / LPT arrays (no kernel launch).
#include <cstddef>

#include <vector_types.h>

#include "ic_modes.cuh"

using namespace mgp;

template <int MODE>
static void apply(int N, double box, int ext, const double2 *src, const double *gtab, double norm, double2 *o0, double2 *o1, double2 *o2) {
  const int NZ = N / 2 + 1;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++)
      for (int k = 0; k < NZ; k++) {
        const size_t e = ((size_t) i * N + j) * NZ + k;
        double2 out[3];
        ic_mode<double, double2, MODE>(N, i, j, k, box, src[e], gtab, norm, ext, out);
        o0[e] = out[0]; o1[e] = out[1]; o2[e] = out[2];
      }
}

extern "C" int ic_mode_f64(int mode, int N, double box, int ext, const double2 *src, const double *gtab, double norm, double2 *o0,
                           double2 *o1, double2 *o2) {
  switch (mode) {
    case 0: apply<0>(N, box, ext, src, gtab, norm, o0, o1, o2); return 0;
    case 1: apply<1>(N, box, ext, src, gtab, norm, o0, o1, o2); return 0;
    case 2: apply<2>(N, box, ext, src, gtab, norm, o0, o1, o2); return 0;
    case 3: apply<3>(N, box, ext, src, gtab, norm, o0, o1, o2); return 0;
    case 4: apply<4>(N, box, ext, src, gtab, norm, o0, o1, o2); return 0;
  }
  return 1;
}
