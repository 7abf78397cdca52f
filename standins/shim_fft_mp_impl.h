/*
 * Slab-decomposed r2c / c2r on top of the 1-D kernels of shim_fft_impl.h, for the multi-process MPI stand-in.
 * Included by shim_fft_mp.c right after shim_fft_impl.h, once per precision (R / PREFIX as there).
 * TEST INFRASTRUCTURE ONLY.
 *
 * Every rank holds local_n0 x-planes [x][y][2 (n2/2+1)], the FFTW-MPI in-place layout (non-transposed output:
 * k-space keeps the same x-slabs).  z and y transforms are local; for the x transforms every rank copies its planes
 * into one grid in the shared segment, takes an equal share of the (y, kz) columns, and copies its planes back.  The
 * 1-D transforms are the same calls on the same numbers as in the serial stand-in, so results are bit-identical to a
 * one-rank run.
 */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(PREFIX, name)

static void FN(x_pass_shared)(int n0, int n1, int n2, int local_n0, int local_start, R *data, int sign) {
  const int nz = n2 / 2 + 1;
  const size_t plane = (size_t) n1 * nz, ncol = plane;
  const int P = mgp_mp_size(), me = mgp_mp_rank();
  FN(cpx) *S = (FN(cpx) *) mgp_mp_scratch(sizeof(FN(cpx)) * plane * (size_t) n0);
  memcpy(S + (size_t) local_start * plane, data, sizeof(FN(cpx)) * plane * (size_t) local_n0);
  mgp_mp_barrier();
  const size_t c0 = ncol * (size_t) me / (size_t) P, c1 = ncol * (size_t) (me + 1) / (size_t) P;
  FN(plan1d) *px = FN(plan1d_new)(n0);
  FN(cpx) *buf = (FN(cpx) *) malloc(sizeof(FN(cpx)) * (size_t) n0 * 8);
  /* in chunks so that the int column count of fft_cols cannot overflow */
  for (size_t c = c0; c < c1; c += 1 << 20) {
    const size_t n = c1 - c < ((size_t) 1 << 20) ? c1 - c : ((size_t) 1 << 20);
    FN(fft_cols)(px, S + c, plane, (int) n, sign, buf);
  }
  free(buf);
  FN(plan1d_free)(px);
  mgp_mp_barrier();
  memcpy(data, S + (size_t) local_start * plane, sizeof(FN(cpx)) * plane * (size_t) local_n0);
  mgp_mp_barrier();                 /* the shared grid may be reused by the next transform */
}

static void FN(r2c_3d_mp)(int n0, int n1, int n2, int local_n0, int local_start, R *data) {
  const int nz = n2 / 2 + 1;
  FN(plan1d) *pz = FN(plan1d_new)(n2), *py = FN(plan1d_new)(n1);
  int nmax = n1 > n2 ? n1 : n2;
  FN(cpx) *buf = (FN(cpx) *) malloc(sizeof(FN(cpx)) * (size_t) nmax * 8);
  FN(cpx) *cd = (FN(cpx) *) data;
  const size_t nrows = (size_t) local_n0 * n1;
  for (size_t r = 0; r < nrows; r += 2) {          /* z: two real rows per complex FFT (n1 is even: pairs stay in a plane) */
    R *ra = data + r * 2 * nz;
    R *rb = (r + 1 < nrows) ? ra + 2 * nz : NULL;
    for (int k = 0; k < n2; k++) { buf[k].re = ra[k]; buf[k].im = rb ? rb[k] : (R) 0; }
    FN(fft1d)(pz, buf, -1);
    FN(cpx) *ca = (FN(cpx) *) ra, *cb = (FN(cpx) *) rb;
    for (int k = 0; k < nz; k++) {
      const FN(cpx) f = buf[k], g = buf[(n2 - k) % n2];
      ca[k].re = (R) 0.5 * (f.re + g.re);
      ca[k].im = (R) 0.5 * (f.im - g.im);
      if (cb) {
        cb[k].re = (R) 0.5 * (f.im + g.im);
        cb[k].im = (R) 0.5 * (g.re - f.re);
      }
    }
  }
  for (int x = 0; x < local_n0; x++) FN(fft_cols)(py, cd + (size_t) x * n1 * nz, (size_t) nz, nz, -1, buf);
  free(buf);
  FN(plan1d_free)(pz); FN(plan1d_free)(py);
  FN(x_pass_shared)(n0, n1, n2, local_n0, local_start, data, -1);
}

static void FN(c2r_3d_mp)(int n0, int n1, int n2, int local_n0, int local_start, R *data) {
  const int nz = n2 / 2 + 1;
  FN(x_pass_shared)(n0, n1, n2, local_n0, local_start, data, +1);
  FN(plan1d) *pz = FN(plan1d_new)(n2), *py = FN(plan1d_new)(n1);
  int nmax = n1 > n2 ? n1 : n2;
  FN(cpx) *buf = (FN(cpx) *) malloc(sizeof(FN(cpx)) * (size_t) nmax * 8);
  FN(cpx) *cd = (FN(cpx) *) data;
  const size_t nrows = (size_t) local_n0 * n1;
  for (int x = 0; x < local_n0; x++) FN(fft_cols)(py, cd + (size_t) x * n1 * nz, (size_t) nz, nz, +1, buf);
  for (size_t r = 0; r < nrows; r += 2) {
    R *ra = data + r * 2 * nz;
    R *rb = (r + 1 < nrows) ? ra + 2 * nz : NULL;
    const FN(cpx) *ca = (const FN(cpx) *) ra, *cb = (const FN(cpx) *) rb;
    for (int k = 0; k < nz; k++) {
      FN(cpx) a = ca[k], b;
      if (cb) b = cb[k]; else { b.re = 0; b.im = 0; }
      if (k == 0 || 2 * k == n2) { a.im = 0; b.im = 0; }
      buf[k].re = a.re - b.im;
      buf[k].im = a.im + b.re;
      if (k > 0 && 2 * k != n2) {
        buf[n2 - k].re = a.re + b.im;
        buf[n2 - k].im = -a.im + b.re;
      }
    }
    FN(fft1d)(pz, buf, +1);
    for (int k = 0; k < n2; k++) { ra[k] = buf[k].re; if (rb) rb[k] = buf[k].im; }
  }
  free(buf);
  FN(plan1d_free)(pz); FN(plan1d_free)(py);
}

#undef FN
#undef CAT
#undef CAT_
