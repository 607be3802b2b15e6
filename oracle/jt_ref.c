/* jt_ref.c -- TEST/BENCH INFRASTRUCTURE ONLY (never linked or loaded by the product package).
 *
 * Plain-C restatement of the *algorithm* JTransforms runs on the CPU for power-of-two complex transforms, used as
 * the timed "port" CPU baseline (bench.py cpu_baseline / --impl reference) because the reference itself is Java and
 * no JVM exists here.  It follows the structure of the reference, not its hand-unrolled text:
 *   - tables:       utils/CommonUtils.java:372-435 (makewt)             -> make_table()
 *   - 1-D driver:   utils/CommonUtils.java:708-793 (cftfsub/cftbsub): one full radix-4 first pass
 *                   (cftf1st :2844 / cftb1st :3282), then the quarters depth-first (cftrec4 :3872, leaves
 *                   cftleaf :3978), then a serial bit-reversal pass (bitrv2 :824 / bitrv2conj :1650)
 *   - 1-D threads:  utils/CommonUtils.java:3722-3795 (cftrec4_th): 2 tasks above 8192 doubles, 4 above 65536,
 *                   never more than 4; first pass and bit reversal stay serial
 *   - 2-D / 3-D:    fft/DoubleFFT_2D.java:115-213, :3352-3529 and fft/DoubleFFT_3D.java:145-162, :5505-5713,
 *                   :6318-6520: rows / slices dealt round-robin to nthreads tasks, strided axes through a
 *                   gather-4-columns temporary
 * Parity: checked against numpy in tests/test_oracle_cref.py; "parity of timing behaviour" with the JVM is NOT
 * claimed (no JIT, no JLargeArrays pool) -- it is a port, labelled as such.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cpx;

/* plan cache: the reference builds its tables in the constructor, outside the timed region
 * (fft/BenchmarkDoubleFFT.java:126-137 "without constructor") */
static cpx* g_tab[8];
static long g_tab_n[8];
static pthread_mutex_t g_tab_mu = PTHREAD_MUTEX_INITIALIZER;
static cpx* make_table_raw(long n);
static cpx* make_table(long n) {
  pthread_mutex_lock(&g_tab_mu);
  for (int i = 0; i < 8; ++i) if (g_tab[i] && g_tab_n[i] == n) { cpx* w = g_tab[i]; pthread_mutex_unlock(&g_tab_mu); return w; }
  static int next = 0;
  cpx* w = make_table_raw(n);
  if (g_tab[next]) free(g_tab[next]);
  g_tab[next] = w; g_tab_n[next] = n; next = (next + 1) % 8;
  pthread_mutex_unlock(&g_tab_mu);
  return w;
}
static cpx* make_table_raw(long n) {          /* w[k] = exp(-2 pi i k / n), k < n */
  cpx* w = (cpx*)malloc(sizeof(cpx) * (size_t)(n > 1 ? n : 1));
  for (long k = 0; k < n; ++k) {
    double a = -2.0 * M_PI * (double)k / (double)n;
    w[k].re = cos(a); w[k].im = sin(a);
  }
  return w;
}

/* one radix-2^2 decimation-in-frequency pass over a block of m points (quarters q = m/4);
 * sgn = -1 forward, +1 inverse; tw step ts = n/m into the table of the full length */
static void pass4(cpx* a, long m, const cpx* w, long ts, int sgn) {
  const long q = m >> 2;
  for (long j = 0; j < q; ++j) {
    cpx x0 = a[j], x1 = a[j + q], x2 = a[j + 2 * q], x3 = a[j + 3 * q];
    cpx s02 = {x0.re + x2.re, x0.im + x2.im}, d02 = {x0.re - x2.re, x0.im - x2.im};
    cpx s13 = {x1.re + x3.re, x1.im + x3.im}, d13 = {x1.re - x3.re, x1.im - x3.im};
    /* forward: -i*d13 = (d13.im, -d13.re); inverse: +i*d13 = (-d13.im, d13.re) */
    cpx r13 = sgn < 0 ? (cpx){d13.im, -d13.re} : (cpx){-d13.im, d13.re};
    cpx y0 = {s02.re + s13.re, s02.im + s13.im};
    cpx y1 = {s02.re - s13.re, s02.im - s13.im};
    cpx y2 = {d02.re + r13.re, d02.im + r13.im};
    cpx y3 = {d02.re - r13.re, d02.im - r13.im};
    cpx w1 = w[j * ts], w2 = w[2 * j * ts], w3 = w[3 * j * ts];
    if (sgn > 0) { w1.im = -w1.im; w2.im = -w2.im; w3.im = -w3.im; }
    a[j] = y0;
    a[j + q] = (cpx){y1.re * w2.re - y1.im * w2.im, y1.re * w2.im + y1.im * w2.re};
    a[j + 2 * q] = (cpx){y2.re * w1.re - y2.im * w1.im, y2.re * w1.im + y2.im * w1.re};
    a[j + 3 * q] = (cpx){y3.re * w3.re - y3.im * w3.im, y3.re * w3.im + y3.im * w3.re};
  }
}

static void rec4(cpx* a, long m, const cpx* w, long ts, int sgn) {   /* depth-first, like cftrec4 */
  if (m >= 4) {
    pass4(a, m, w, ts, sgn);
    const long q = m >> 2;
    if (q > 1) for (int k = 0; k < 4; ++k) rec4(a + k * q, q, w, ts * 4, sgn);
  } else if (m == 2) {
    cpx x0 = a[0], x1 = a[1];
    a[0] = (cpx){x0.re + x1.re, x0.im + x1.im};
    a[1] = (cpx){x0.re - x1.re, x0.im - x1.im};
  }
}

static void bitrev(cpx* a, long n) {
  for (long i = 0, j = 0; i < n; ++i) {
    if (i < j) { cpx t = a[i]; a[i] = a[j]; a[j] = t; }
    long bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j |= bit;
  }
}

typedef struct { cpx* a; long m; const cpx* w; long ts; int sgn; int k0, k1; } quarter_job;
static void* quarter_main(void* p) {
  quarter_job* j = (quarter_job*)p;
  for (int k = j->k0; k < j->k1; ++k) rec4(j->a + (long)k * j->m, j->m, j->w, j->ts, j->sgn);
  return NULL;
}

static void cfft1d_tab(cpx* a, long n, int sgn, const cpx* w, int nthreads) {
  if (n < 2) return;
  if (n < 4) { rec4(a, n, w, 1, sgn); return; }
  pass4(a, n, w, 1, sgn);                               /* cftf1st / cftb1st: serial full pass */
  const long q = n >> 2;
  if (q > 1) {
    int nt = 1;                                         /* cftrec4_th thresholds on 2n doubles */
    if (nthreads > 1 && 2 * n > 8192) nt = 2;
    if (nthreads >= 4 && 2 * n > 65536) nt = 4;
    if (nt == 1) {
      for (int k = 0; k < 4; ++k) rec4(a + k * q, q, w, 4, sgn);
    } else {
      pthread_t th[4];
      quarter_job jb[4];
      for (int t = 0; t < nt; ++t) {
        jb[t] = (quarter_job){a, q, w, 4, sgn, t * (4 / nt), (t + 1) * (4 / nt)};
        pthread_create(&th[t], NULL, quarter_main, &jb[t]);
      }
      for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
    }
  }
  bitrev(a, n);                                         /* bitrv2 / bitrv2conj: serial */
}

/* DoubleFFT_1D.complexForward / complexInverse (unscaled) for n = 2^k */
int jtref_cfft1d(double* a, long n, int isgn, int nthreads) {
  if (n < 1 || (n & (n - 1))) return 1;
  cpx* w = make_table(n);
  cfft1d_tab((cpx*)a, n, isgn, w, nthreads);
  return 0;
}

/* ---------------------------------------------------------------- N-D drivers */
typedef struct {
  cpx* a; long S, R, C; int sgn, nthreads, tid, phase;
  const cpx *wS, *wR, *wC;
} nd_job;

/* strided axis through the reference's gather-4-columns temporary */
static void strided_fft(cpx* base, long len, long stride, long ncols, const cpx* w, int sgn, cpx* t) {
  for (long c = 0; c < ncols; c += 4) {
    const long nc = ncols - c < 4 ? ncols - c : 4;
    for (long r = 0; r < len; ++r) for (long k = 0; k < nc; ++k) t[k * len + r] = base[r * stride + c + k];
    for (long k = 0; k < nc; ++k) cfft1d_tab(t + k * len, len, sgn, w, 1);
    for (long r = 0; r < len; ++r) for (long k = 0; k < nc; ++k) base[r * stride + c + k] = t[k * len + r];
  }
}

static void* nd_main(void* p) {
  nd_job* j = (nd_job*)p;
  const long S = j->S, R = j->R, C = j->C;
  const long mx = S > R ? S : R;
  cpx* t = (cpx*)malloc(sizeof(cpx) * (size_t)(4 * mx));
  if (j->phase == 0) {            /* xdft3da_subth2: per slice, rows then the row axis inside the slice */
    for (long s = j->tid; s < S; s += j->nthreads) {
      cpx* sl = j->a + s * R * C;
      for (long r = 0; r < R; ++r) cfft1d_tab(sl + r * C, C, j->sgn, j->wC, 1);
      if (R > 1) strided_fft(sl, R, C, C, j->wR, j->sgn, t);
    }
  } else {                        /* cdft3db_subth: per row index, the slice axis */
    for (long r = j->tid; r < R; r += j->nthreads) strided_fft(j->a + r * C, S, R * C, C, j->wS, j->sgn, t);
  }
  free(t);
  return NULL;
}

/* DoubleFFT_3D.complexForward/Inverse (unscaled), power-of-two sizes; S == 1 gives DoubleFFT_2D */
int jtref_cfft3d(double* a, long S, long R, long C, int isgn, int nthreads) {
  if (S < 1 || R < 1 || C < 1 || (S & (S - 1)) || (R & (R - 1)) || (C & (C - 1))) return 1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  cpx *wS = make_table(S), *wR = make_table(R), *wC = make_table(C);
  pthread_t th[256];
  nd_job jb[256];
  for (int phase = 0; phase < (S > 1 ? 2 : 1); ++phase) {
    for (int t = 0; t < nthreads; ++t) {
      jb[t] = (nd_job){(cpx*)a, S, R, C, isgn, nthreads, t, phase, wS, wR, wC};
      pthread_create(&th[t], NULL, nd_main, &jb[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  }
  return 0;
}
