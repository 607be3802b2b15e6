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
 *   - real 2-D:     fft/DoubleFFT_2D.java:820-838: rows rdft (utils/CommonUtils.java:5750 rftfsub after the half-length
 *                   complex transform), complex columns incl. the pseudo column 0 (cdft2d_subth :3352-3529),
 *                   rdft2d_sub (:2544-2574)                              -> jtref_rfft2d()
 *   - DCT/DST/DHT:  dct/DoubleDCT_2D.java:104-183 + ddct2d_subth :625-957 (rows round-robin, columns through the
 *                   gather-4-columns temporary), dst/DoubleDST_2D.java:103, dht/DoubleDHT_2D.java:102-190 + yTransform
 *                   :1288-1309; the 1-D kernel is a half-length complex FFT + pre/post twiddle like Ooura's ddct/dfst
 *                   (utils/CommonUtils.java:5860 dctsub)                 -> jtref_r2r2d()
 *   - Bluestein:    fft/FloatFFT_1D.java bluestein_complex (fft/DoubleFFT_1D.java:1920-2107): per-call ak buffer of
 *                   2*nBluestein floats, chirp multiply, cftbsub, * bk2, cftfsub, chirp multiply -> jtref_bluestein_f32()
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


/* ---------------------------------------------------------------- real transforms (pow2 lines) */
/* DoubleFFT_1D.realForward for n = 2^k >= 4, packed output a[0] = Re X[0], a[1] = Re X[n/2], a[2k], a[2k+1] = X[k]
 * (fft/DoubleFFT_1D.java:524-546): half-length complex transform + the rftfsub split; wh = table of n/2, wn = table of n */
static void rfft1d_tab(double* a, long n, const cpx* wh, const cpx* wn) {
  const long N = n / 2;
  cpx* z = (cpx*)a;
  cfft1d_tab(z, N, -1, wh, 1);
  const cpx z0 = z[0];
  z[0] = (cpx){z0.re + z0.im, z0.re - z0.im};
  for (long k = 1; 2 * k < N; ++k) {
    const cpx p = z[k], q = z[N - k], w = wn[k];
    const cpx ev = {0.5 * (p.re + q.re), 0.5 * (p.im - q.im)}, df = {0.5 * (p.re - q.re), 0.5 * (p.im + q.im)};
    cpx od = {df.re * w.re - df.im * w.im, df.re * w.im + df.im * w.re};
    od = (cpx){od.im, -od.re};
    z[k] = (cpx){ev.re + od.re, ev.im + od.im};
    z[N - k] = (cpx){ev.re - od.re, -(ev.im - od.im)};
  }
  if (N >= 2) z[N / 2].im = -z[N / 2].im;
}

typedef struct { double* a; long R, C; int nthreads, tid, phase, kind; const cpx *wR, *wRh, *wC, *wCh; double f0, f; } r2_job;

static void* rfft2d_main(void* p) {
  r2_job* j = (r2_job*)p;
  const long R = j->R, C = j->C, H = C / 2;
  if (j->phase == 0) {                      /* rows: realForward of every row, rows dealt round-robin */
    for (long r = j->tid; r < R; r += j->nthreads) rfft1d_tab(j->a + r * C, C, j->wCh, j->wC);
  } else {                                  /* complex columns (H of them, pseudo column 0 included), blocks of 4 */
    cpx* t = (cpx*)malloc(sizeof(cpx) * (size_t)(4 * R));
    cpx* z = (cpx*)j->a;
    for (long c = 4 * j->tid; c < H; c += 4 * j->nthreads) {
      const long nc = H - c < 4 ? H - c : 4;
      for (long r = 0; r < R; ++r) for (long k = 0; k < nc; ++k) t[k * R + r] = z[r * H + c + k];
      for (long k = 0; k < nc; ++k) cfft1d_tab(t + k * R, R, -1, j->wR, 1);
      for (long r = 0; r < R; ++r) for (long k = 0; k < nc; ++k) z[r * H + c + k] = t[k * R + r];
    }
    free(t);
  }
  return NULL;
}

static void run_jobs(void* (*fn)(void*), r2_job* proto, int nthreads) {
  pthread_t th[256];
  r2_job jb[256];
  for (int t = 0; t < nthreads; ++t) { jb[t] = *proto; jb[t].tid = t; jb[t].nthreads = nthreads; pthread_create(&th[t], NULL, fn, &jb[t]); }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* DoubleFFT_2D.realForward, power-of-two sizes, packed layout of fft/DoubleFFT_2D.java:794-810 */
int jtref_rfft2d(double* a, long R, long C, int nthreads) {
  if (R < 2 || C < 4 || (R & (R - 1)) || (C & (C - 1))) return 1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  r2_job pr = {a, R, C, nthreads, 0, 0, 0, make_table(R), NULL, make_table(C), make_table(C / 2), 1.0, 1.0};
  run_jobs(rfft2d_main, &pr, nthreads);
  pr.phase = 1;
  run_jobs(rfft2d_main, &pr, nthreads);
  for (long i = 1; i < R / 2; ++i) {          /* rdft2d_sub(1): untangle the two real columns packed in column 0 */
    const long k = R - i;
    double* x = a;
    x[k * C] = 0.5 * (x[i * C] - x[k * C]);
    x[i * C] -= x[k * C];
    x[k * C + 1] = 0.5 * (x[i * C + 1] + x[k * C + 1]);
    x[i * C + 1] -= x[k * C + 1];
  }
  return 0;
}

/* one DCT-II / DST-II / DHT line of n = 2^k >= 4 reals in place through a scratch line v (n doubles):
 * kind 0 DCT, 1 DST, 2 DHT; f0, f = output factors (k = 0, k > 0) */
static void r2r_line(double* x, long n, int kind, double f0, double f, double* v, const cpx* wh, const cpx* wn, const cpx* w4) {
  if (kind == 2) {
    memcpy(v, x, sizeof(double) * (size_t)n);
    rfft1d_tab(v, n, wh, wn);
    x[0] = v[0] * f; x[n / 2] = v[1] * f;
    for (long k = 1; k < n / 2; ++k) { x[k] = (v[2 * k] - v[2 * k + 1]) * f; x[n - k] = (v[2 * k] + v[2 * k + 1]) * f; }
    return;
  }
  /* Makhoul: v[j] = x[2j], v[n-1-j] = x[2j+1]; V = rfft(v); C[k] = Re(e^{-i pi k/2n} V[k]), C[n-k] = -Im(...) */
  for (long j = 0; j < n / 2; ++j) {
    const double e = x[2 * j], o = x[2 * j + 1];
    v[j] = e; v[n - 1 - j] = kind == 1 ? -o : o;      /* DST: odd inputs negated (dst/DoubleDST_1D.java:113-117) */
  }
  rfft1d_tab(v, n, wh, wn);
  double* y = x;
  const long last = n - 1;
#define PUT(k, val) do { if (kind == 1) y[last - (k)] = (val); else y[(k)] = (val); } while (0)   /* DST: reversed */
  PUT(0, v[0] * f0);
  PUT(n / 2, v[1] * 0.70710678118654752440 * f);
  for (long k = 1; k < n / 2; ++k) {
    const cpx w = w4[k];                              /* exp(-i pi k / 2n) = table of 4n at k */
    const double tr = v[2 * k] * w.re - v[2 * k + 1] * w.im, ti = v[2 * k] * w.im + v[2 * k + 1] * w.re;
    PUT(k, tr * f);
    PUT(n - k, -ti * f);
  }
#undef PUT
}

static void* r2r2d_main(void* p) {
  r2_job* j = (r2_job*)p;
  const long R = j->R, C = j->C;
  if (j->phase == 0) {                      /* rows round-robin (ddct2d_subth, dct/DoubleDCT_2D.java:625-957) */
    double* v = (double*)malloc(sizeof(double) * (size_t)C);
    for (long r = j->tid; r < R; r += j->nthreads) r2r_line(j->a + r * C, C, j->kind, j->f0, j->f, v, j->wCh, j->wC, (const cpx*)j->wRh);
    free(v);
  } else {                                  /* columns through the gather-4-columns temporary */
    double* t = (double*)malloc(sizeof(double) * (size_t)(5 * R));
    for (long c = 4 * j->tid; c < C; c += 4 * j->nthreads) {
      const long nc = C - c < 4 ? C - c : 4;
      for (long r = 0; r < R; ++r) for (long k = 0; k < nc; ++k) t[k * R + r] = j->a[r * C + c + k];
      for (long k = 0; k < nc; ++k) r2r_line(t + k * R, R, j->kind, j->f0, j->f, t + 4 * R, j->wCh, j->wC, (const cpx*)j->wRh);
      for (long r = 0; r < R; ++r) for (long k = 0; k < nc; ++k) j->a[r * C + c + k] = t[k * R + r];
    }
    free(t);
  }
  return NULL;
}

static cpx* g_q4[4]; static long g_q4_n[4];
static const cpx* quarter_table(long n) {      /* exp(-i pi k / 2n), k < n */
  pthread_mutex_lock(&g_tab_mu);
  for (int i = 0; i < 4; ++i) if (g_q4[i] && g_q4_n[i] == n) { cpx* w = g_q4[i]; pthread_mutex_unlock(&g_tab_mu); return w; }
  static int next = 0;
  cpx* w = (cpx*)malloc(sizeof(cpx) * (size_t)n);
  for (long k = 0; k < n; ++k) { const double a = -M_PI * (double)k / (2.0 * (double)n); w[k].re = cos(a); w[k].im = sin(a); }
  if (g_q4[next]) free(g_q4[next]);
  g_q4[next] = w; g_q4_n[next] = n; next = (next + 1) % 4;
  pthread_mutex_unlock(&g_tab_mu);
  return w;
}

/* DoubleDCT_2D / DoubleDST_2D .forward(a, scale=true) and DoubleDHT_2D.forward(a), power-of-two sizes.
 * kind 0 DCT, 1 DST, 2 DHT */
int jtref_r2r2d(double* a, long R, long C, int kind, int nthreads) {
  if (R < 4 || C < 4 || (R & (R - 1)) || (C & (C - 1)) || kind < 0 || kind > 2) return 1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  r2_job pr = {a, R, C, nthreads, 0, 0, kind, NULL, NULL, make_table(C), make_table(C / 2), 1.0, 1.0};
  pr.wRh = quarter_table(C);
  if (kind != 2) { pr.f0 = sqrt(1.0 / (double)C); pr.f = sqrt(2.0 / (double)C); }
  run_jobs(r2r2d_main, &pr, nthreads);                 /* rows (length C) */
  r2_job pc = pr;
  pc.phase = 1; pc.wC = make_table(R); pc.wCh = make_table(R / 2); pc.wRh = quarter_table(R);
  if (kind != 2) { pc.f0 = sqrt(1.0 / (double)R); pc.f = sqrt(2.0 / (double)R); }
  run_jobs(r2r2d_main, &pc, nthreads);                 /* columns (length R) */
  if (kind == 2) {                                     /* yTransform (dht/DoubleDHT_2D.java:1288-1309) */
    for (long r = 0; r <= R / 2; ++r) {
      const long mr = (R - r) % R;
      for (long c = 0; c <= C / 2; ++c) {
        const long mc = (C - c) % C;
        const double A = a[r * C + c], B = a[mr * C + c], Cc = a[r * C + mc], D = a[mr * C + mc];
        const double E = 0.5 * ((A + D) - (B + Cc));
        a[r * C + c] = A - E; a[mr * C + c] = B + E; a[r * C + mc] = Cc + E; a[mr * C + mc] = D - E;
      }
    }
  }
  return 0;
}

/* ---------------------------------------------------------------- Bluestein, single precision */
typedef struct { float re, im; } cpxf;
static void pass4f(cpxf* a, long m, const cpxf* w, long ts, int sgn) {
  const long q = m >> 2;
  for (long j = 0; j < q; ++j) {
    cpxf x0 = a[j], x1 = a[j + q], x2 = a[j + 2 * q], x3 = a[j + 3 * q];
    cpxf s02 = {x0.re + x2.re, x0.im + x2.im}, d02 = {x0.re - x2.re, x0.im - x2.im};
    cpxf s13 = {x1.re + x3.re, x1.im + x3.im}, d13 = {x1.re - x3.re, x1.im - x3.im};
    cpxf r13 = sgn < 0 ? (cpxf){d13.im, -d13.re} : (cpxf){-d13.im, d13.re};
    cpxf y0 = {s02.re + s13.re, s02.im + s13.im}, y1 = {s02.re - s13.re, s02.im - s13.im};
    cpxf y2 = {d02.re + r13.re, d02.im + r13.im}, y3 = {d02.re - r13.re, d02.im - r13.im};
    cpxf w1 = w[j * ts], w2 = w[2 * j * ts], w3 = w[3 * j * ts];
    if (sgn > 0) { w1.im = -w1.im; w2.im = -w2.im; w3.im = -w3.im; }
    a[j] = y0;
    a[j + q] = (cpxf){y1.re * w2.re - y1.im * w2.im, y1.re * w2.im + y1.im * w2.re};
    a[j + 2 * q] = (cpxf){y2.re * w1.re - y2.im * w1.im, y2.re * w1.im + y2.im * w1.re};
    a[j + 3 * q] = (cpxf){y3.re * w3.re - y3.im * w3.im, y3.re * w3.im + y3.im * w3.re};
  }
}
static void rec4f(cpxf* a, long m, const cpxf* w, long ts, int sgn) {
  if (m >= 4) {
    pass4f(a, m, w, ts, sgn);
    const long q = m >> 2;
    if (q > 1) for (int k = 0; k < 4; ++k) rec4f(a + k * q, q, w, ts * 4, sgn);
  } else if (m == 2) {
    cpxf x0 = a[0], x1 = a[1];
    a[0] = (cpxf){x0.re + x1.re, x0.im + x1.im};
    a[1] = (cpxf){x0.re - x1.re, x0.im - x1.im};
  }
}
static void bitrevf(cpxf* a, long n) {
  for (long i = 0, j = 0; i < n; ++i) {
    if (i < j) { cpxf t = a[i]; a[i] = a[j]; a[j] = t; }
    long bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j |= bit;
  }
}
static void cfft1df(cpxf* a, long n, int sgn, const cpxf* w) { rec4f(a, n, w, 1, sgn); bitrevf(a, n); }

static struct { long n, M; cpxf *bk1, *bk2, *w; } g_blue;
static void blue_plan(long n) {              /* bluesteini (fft/DoubleFFT_1D.java:1864-1890), tables in double, rounded once */
  if (g_blue.n == n) return;
  long M = 1;
  while (M < 2 * n - 1) M <<= 1;
  free(g_blue.bk1); free(g_blue.bk2); free(g_blue.w);
  g_blue.n = n; g_blue.M = M;
  g_blue.bk1 = (cpxf*)malloc(sizeof(cpxf) * (size_t)n);
  g_blue.bk2 = (cpxf*)malloc(sizeof(cpxf) * (size_t)M);
  g_blue.w = (cpxf*)malloc(sizeof(cpxf) * (size_t)M);
  for (long k = 0; k < M; ++k) { const double a = -2.0 * M_PI * (double)k / (double)M; g_blue.w[k] = (cpxf){(float)cos(a), (float)sin(a)}; }
  cpx* t = (cpx*)calloc((size_t)M, sizeof(cpx));
  cpx* wd = make_table_raw(M);
  for (long i = 0; i < n; ++i) {
    const unsigned long long ph = ((unsigned long long)i * (unsigned long long)i) % (unsigned long long)(2 * n);
    const double a = M_PI * (double)ph / (double)n;
    const cpx b = {cos(a), sin(a)};
    g_blue.bk1[i] = (cpxf){(float)b.re, (float)b.im};
    t[i] = (cpx){b.re / (double)M, b.im / (double)M};
    if (i > 0) t[M - i] = t[i];
  }
  cfft1d_tab(t, M, -1, wd, 1);
  for (long k = 0; k < M; ++k) g_blue.bk2[k] = (cpxf){(float)t[k].re, (float)t[k].im};
  free(t); free(wd);
}
typedef struct { float* a; long n, nb; int nthreads, tid; } blue_job;
static void* blue_main(void* p) {
  blue_job* j = (blue_job*)p;
  const long n = j->n, M = g_blue.M;
  for (long b = j->tid; b < j->nb; b += j->nthreads) {
    cpxf* x = (cpxf*)(j->a + 2 * n * b);
    cpxf* ak = (cpxf*)calloc((size_t)M, sizeof(cpxf));        /* the reference allocates ak per call */
    for (long i = 0; i < n; ++i) {
      const cpxf c = g_blue.bk1[i];
      ak[i] = (cpxf){x[i].re * c.re + x[i].im * c.im, x[i].im * c.re - x[i].re * c.im};
    }
    cfft1df(ak, M, -1, g_blue.w);
    for (long k = 0; k < M; ++k) {
      const cpxf c = g_blue.bk2[k], v = ak[k];
      ak[k] = (cpxf){v.re * c.re - v.im * c.im, v.re * c.im + v.im * c.re};
    }
    cfft1df(ak, M, +1, g_blue.w);
    for (long i = 0; i < n; ++i) {
      const cpxf c = g_blue.bk1[i], v = ak[i];
      x[i] = (cpxf){v.re * c.re + v.im * c.im, v.im * c.re - v.re * c.im};
    }
    free(ak);
  }
  return NULL;
}
/* FloatFFT_1D.complexForward for any n through the Bluestein path, `nb` transforms 2n floats apart; the reference runs
 * them one after the other (each with at most 4 threads) -- here they are dealt to nthreads tasks, one transform per
 * task at a time, which is the faster arrangement on a many-core host */
int jtref_bluestein_f32(float* a, long n, long nb, int nthreads) {
  if (n < 2 || nb < 1) return 1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if (nthreads > nb) nthreads = (int)nb;
  blue_plan(n);
  pthread_t th[256];
  blue_job jb[256];
  for (int t = 0; t < nthreads; ++t) { jb[t] = (blue_job){a, n, nb, nthreads, t}; pthread_create(&th[t], NULL, blue_main, &jb[t]); }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  return 0;
}
