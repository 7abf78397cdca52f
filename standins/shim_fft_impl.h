/*
 * Precision-generic body of the CPU FFT behind the FFTW stand-in.  Included twice
 * by shim_fft.c with R = double / float.  TEST INFRASTRUCTURE ONLY.
 *
 * 1-D kernel: Stockham autosort radix-2 for powers of two, recursive mixed-radix
 * (naive DFT at prime leaves) otherwise.  3-D r2c/c2r are built from it in the
 * order FFTW and cuFFT use (z first for r2c; z last for c2r), two real rows per
 * complex transform.
 */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(PREFIX, name)

typedef struct { R re, im; } FN(cpx);

typedef struct {
  int n;
  int pow2;
  FN(cpx) *tw;      /* tw[k] = exp(-2 pi i k / n), k < n */
  FN(cpx) *work;    /* n scratch */
} FN(plan1d);

static FN(plan1d) *FN(plan1d_new)(int n) {
  FN(plan1d) *p = (FN(plan1d) *) malloc(sizeof(*p));
  p->n = n;
  p->pow2 = (n & (n - 1)) == 0;
  p->tw = (FN(cpx) *) malloc(sizeof(FN(cpx)) * (size_t) n);
  p->work = (FN(cpx) *) malloc(sizeof(FN(cpx)) * (size_t) n);
  for (int k = 0; k < n; k++) {
    double a = -2.0 * M_PI * (double) k / (double) n;
    p->tw[k].re = (R) cos(a);
    p->tw[k].im = (R) sin(a);
  }
  return p;
}
static void FN(plan1d_free)(FN(plan1d) *p) { free(p->tw); free(p->work); free(p); }

/* Stockham radix-2, forward sign (exp(-i...)); result ends in x. */
static void FN(stockham)(const FN(plan1d) *pl, FN(cpx) *x) {
  const int N = pl->n;
  FN(cpx) *a = x, *b = pl->work;
  int s = 1;
  for (int n = N; n > 1; n >>= 1, s <<= 1) {
    const int m = n >> 1;
    const int tstep = N / n;
    for (int p = 0; p < m; p++) {
      const FN(cpx) w = pl->tw[p * tstep];
      const FN(cpx) *a0 = a + (size_t) s * p, *a1 = a + (size_t) s * (p + m);
      FN(cpx) *b0 = b + (size_t) s * 2 * p, *b1 = b0 + s;
      for (int q = 0; q < s; q++) {
        const R ur = a0[q].re, ui = a0[q].im, vr = a1[q].re, vi = a1[q].im;
        b0[q].re = ur + vr; b0[q].im = ui + vi;
        const R dr = ur - vr, di = ui - vi;
        b1[q].re = dr * w.re - di * w.im;
        b1[q].im = dr * w.im + di * w.re;
      }
    }
    FN(cpx) *t = a; a = b; b = t;
  }
  if (a != x) memcpy(x, a, sizeof(FN(cpx)) * (size_t) N);
}

/* generic recursive mixed radix: out[0..n) = DFT(in[0], in[stride], ...) */
static void FN(mixed)(const FN(plan1d) *pl, int n, int stride, const FN(cpx) *in, FN(cpx) *out) {
  const int N = pl->n;
  if (n == 1) { out[0] = in[0]; return; }
  int p = 2;
  while (n % p) p++;
  const int m = n / p;
  for (int r = 0; r < p; r++) FN(mixed)(pl, m, stride * p, in + (size_t) r * stride, out + (size_t) r * m);
  const int tstep = N / n;
  FN(cpx) t[64];
  FN(cpx) *tt = p <= 64 ? t : (FN(cpx) *) malloc(sizeof(FN(cpx)) * (size_t) p);
  for (int k = 0; k < m; k++) {
    for (int r = 0; r < p; r++) {
      const FN(cpx) w = pl->tw[((size_t) r * k % n) * tstep];
      const FN(cpx) y = out[(size_t) r * m + k];
      tt[r].re = y.re * w.re - y.im * w.im;
      tt[r].im = y.re * w.im + y.im * w.re;
    }
    for (int q = 0; q < p; q++) {
      R sr = 0, si = 0;
      for (int r = 0; r < p; r++) {
        const FN(cpx) w = pl->tw[((size_t) r * q % p) * (N / p)];
        sr += tt[r].re * w.re - tt[r].im * w.im;
        si += tt[r].re * w.im + tt[r].im * w.re;
      }
      out[(size_t) q * m + k].re = sr;
      out[(size_t) q * m + k].im = si;
    }
  }
  if (tt != t) free(tt);
}

/* in-place 1-D transform of a contiguous buffer; sign = -1 forward, +1 backward */
static void FN(fft1d)(const FN(plan1d) *pl, FN(cpx) *x, int sign) {
  const int n = pl->n;
  if (n == 1) return;
  if (sign > 0) for (int i = 0; i < n; i++) x[i].im = -x[i].im;   /* conj trick */
  if (pl->pow2) {
    FN(stockham)(pl, x);
  } else {
    FN(cpx) *tmp = (FN(cpx) *) malloc(sizeof(FN(cpx)) * (size_t) n);
    memcpy(tmp, x, sizeof(FN(cpx)) * (size_t) n);
    FN(mixed)(pl, n, 1, tmp, x);
    free(tmp);
  }
  if (sign > 0) for (int i = 0; i < n; i++) x[i].im = -x[i].im;
}

/* transform along a strided axis for `ncol` adjacent columns (col stride 1) */
static void FN(fft_cols)(const FN(plan1d) *pl, FN(cpx) *base, size_t stride, int ncol, int sign, FN(cpx) *buf) {
  const int n = pl->n;
  for (int c0 = 0; c0 < ncol; c0 += 8) {
    const int nc = ncol - c0 < 8 ? ncol - c0 : 8;
    for (int i = 0; i < n; i++)
      for (int c = 0; c < nc; c++) buf[(size_t) c * n + i] = base[(size_t) i * stride + c0 + c];
    for (int c = 0; c < nc; c++) FN(fft1d)(pl, buf + (size_t) c * n, sign);
    for (int i = 0; i < n; i++)
      for (int c = 0; c < nc; c++) base[(size_t) i * stride + c0 + c] = buf[(size_t) c * n + i];
  }
}

static void FN(r2c_3d)(int n0, int n1, int n2, R *data) {
  const int nz = n2 / 2 + 1;
  FN(plan1d) *pz = FN(plan1d_new)(n2), *py = FN(plan1d_new)(n1), *px = FN(plan1d_new)(n0);
  int nmax = n0 > n1 ? n0 : n1; if (n2 > nmax) nmax = n2;
  FN(cpx) *buf = (FN(cpx) *) malloc(sizeof(FN(cpx)) * (size_t) nmax * 8);
  FN(cpx) *cd = (FN(cpx) *) data;
  const size_t nrows = (size_t) n0 * n1;
  /* z: two real rows per complex FFT */
  for (size_t r = 0; r < nrows; r += 2) {
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
  /* y */
  for (int x = 0; x < n0; x++) FN(fft_cols)(py, cd + (size_t) x * n1 * nz, (size_t) nz, nz, -1, buf);
  /* x */
  FN(fft_cols)(px, cd, (size_t) n1 * nz, n1 * nz, -1, buf);
  free(buf);
  FN(plan1d_free)(pz); FN(plan1d_free)(py); FN(plan1d_free)(px);
}

static void FN(c2r_3d)(int n0, int n1, int n2, R *data) {
  const int nz = n2 / 2 + 1;
  FN(plan1d) *pz = FN(plan1d_new)(n2), *py = FN(plan1d_new)(n1), *px = FN(plan1d_new)(n0);
  int nmax = n0 > n1 ? n0 : n1; if (n2 > nmax) nmax = n2;
  FN(cpx) *buf = (FN(cpx) *) malloc(sizeof(FN(cpx)) * (size_t) nmax * 8);
  FN(cpx) *cd = (FN(cpx) *) data;
  const size_t nrows = (size_t) n0 * n1;
  FN(fft_cols)(px, cd, (size_t) n1 * nz, n1 * nz, +1, buf);
  for (int x = 0; x < n0; x++) FN(fft_cols)(py, cd + (size_t) x * n1 * nz, (size_t) nz, nz, +1, buf);
  /* z: Hermitian-extend two half rows A, B into Z = A + iB, one complex inverse FFT.
     Imaginary parts of the k = 0 and k = n2/2 inputs are dropped (FFTW c2r behaviour). */
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
  FN(plan1d_free)(pz); FN(plan1d_free)(py); FN(plan1d_free)(px);
}

#undef FN
#undef CAT
#undef CAT_
