// Host side of the mixed-radix line kernel (jtb_mixed.cuh): factorisation, table, launch geometry.
#include <cmath>
#include <cstdlib>

#include "jtb_engine_impl.cuh"
#include "jtb_mixed.cuh"

namespace jtb {

// n = product of radices from {4, 2, 3, 5, 7, 11, 13}; false when another prime factor remains
static bool mixed_factor(i64 n, int* radix, int* nstages) {
  static const int cand[] = {4, 2, 3, 5, 7, 11, 13};
  int ns = 0;
  // odd radices first: they run with Ns small, where the generic butterflies' twiddle reads are few
  i64 rem = n;
  for (int ci = 2; ci < 7; ++ci)
    while (rem % cand[ci] == 0) { if (ns >= MIX_MAX_STAGES) return false; radix[ns++] = cand[ci]; rem /= cand[ci]; }
  while (rem % 4 == 0) { if (ns >= MIX_MAX_STAGES) return false; radix[ns++] = 4; rem /= 4; }
  while (rem % 2 == 0) { if (ns >= MIX_MAX_STAGES) return false; radix[ns++] = 2; rem /= 2; }
  *nstages = ns;
  return rem == 1 && ns > 0;
}

template <typename T> static bool mixed_small_radices(const MixedParams<T>& p) {
  static const bool off = getenv("JTB_MIXED_NO_BIG") != nullptr;
  if (off) return false;
  for (int s = 0; s < p.nstages; ++s) if (p.radix[s] > 5) return false;
  return true;
}
template <typename T> static int mixed_attr(Engine<T>& e) {
  static bool attr_done[16] = {false, false};
  const int dv = (e.ctx->device & 7) * 2 + (sizeof(T) == 8 ? 0 : 1);
  if (!attr_done[dv]) {
    JTB_CUDA(cudaFuncSetAttribute((fft_mixed_kernel<T, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
    JTB_CUDA(cudaFuncSetAttribute((fft_mixed_kernel<T, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
    attr_done[dv] = true;
  }
  return ST_OK;
}

template <typename T>
int mixed_c2c(Engine<T>& e, cx<T>* a, const Geo& g, i64 nlines, i64 n, bool inverse, bool has_scale, T scale, bool* handled) {
  typedef cx<T> C;
  *handled = false;
  static const bool off = getenv("JTB_NO_MIXED") != nullptr;
  if (off || nlines <= 0 || n < 2 || n > 8192) return ST_OK;
  MixedParams<T> p;
  memset(&p, 0, sizeof p);
  if (!mixed_factor(n, p.radix, &p.nstages)) return ST_OK;
  const size_t line_bytes = 2 * (size_t)(n + 1) * sizeof(C);      // two buffers
  const size_t cap = (size_t)200 * 1024;
  if (line_bytes > cap) return ST_OK;
  const bool strided = g.stride != 1;
  int W;
  if (strided) {
    W = (int)(128 / sizeof(C));
    if (g.c[0] > 1 && g.d[0] == 1) { while (W > 1 && (g.c[0] % W) != 0) W >>= 1; } else W = 1;
  } else {
    W = (int)(4096 / n);
    if (W < 1) W = 1;
    if (W > 16) W = 16;
  }
  while (W > 1 && (size_t)W * line_bytes > cap) W >>= 1;
  if ((i64)W > nlines) W = (int)nlines;
  const size_t smem = (size_t)W * line_bytes;
  JTB_TRY(mixed_attr<T>(e));
  const std::string key = mkkey("mixw", e.pname(), n);
  void* d = e.ctx->table(key);
  if (!d) {
    std::vector<C> h((size_t)n);
    for (i64 j = 0; j < n; ++j) h[(size_t)j] = unit_root<T>(j, n);
    JTB_TRY(e.ctx->put_table(key, h.data(), h.size() * sizeof(C), &d));
  }
  p.in = a; p.out = a; p.gi = g; p.go = g;
  p.nlines = nlines; p.line_base = 0;
  p.n = (int)n; p.W = W; p.wfast = strided && W > 1;
  p.swap_in = inverse; p.swap_out = inverse; p.has_scale = has_scale; p.scale = scale;
  p.wtab = (const C*)d;
  p.m_n = mix_magic((unsigned)n);
  p.logW = 0;
  while ((1 << p.logW) < W) ++p.logW;
  {
    i64 ns = 1;
    for (int s = 0; s < p.nstages; ++s) {
      p.m_nb[s] = mix_magic((unsigned)(n / p.radix[s]));
      p.m_ns[s] = mix_magic((unsigned)ns);
      ns *= p.radix[s];
    }
  }
  const i64 work = (i64)W * n / 4;                       // butterflies of a radix-4 stage
  const bool big = mixed_small_radices(p) && work >= 1024;
  const unsigned threads = big ? 1024u : (work >= 512 ? 512u : (work >= 256 ? 256u : (work >= 128 ? 128u : 64u)));
  const i64 nblk = (nlines + W - 1) / W;
  if (nblk > 0x7fffffffLL) return ST_OK;
  if (big) JTB_LAUNCH((fft_mixed_kernel<T, true>), (unsigned)nblk, threads, smem, e.st, p);
  else JTB_LAUNCH((fft_mixed_kernel<T, false>), (unsigned)nblk, threads, smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}

// ------------------------------------------------------------------------------------------ two-pass mixed radix
// Smooth lengths beyond one CTA (the reference runs them through FFTPACK, fft/DoubleFFT_1D.java:6630-8009; its own
// benchmark sizes 10368 ... 6250000, fft/BenchmarkDoubleFFT.java:56): n = N1*N2, line viewed as [N1][N2]
//   pass 1: N2 strided transforms of length N1 (W adjacent lines per CTA), twiddle W_n^(k1*j2) at the store -> work
//   pass 2: N1 contiguous transforms of length N2, stored transposed (out[k1 + N1*k2]) with k1 fastest -> array
// Two sweeps instead of the 3-4 sweeps at 2-4x the points of the Bluestein route.
namespace {
template <typename T> int mixed_setup(Engine<T>& e, MixedParams<T>& p, i64 n, int W, bool wfast) {
  typedef cx<T> C;
  if (!mixed_factor(n, p.radix, &p.nstages)) return ST_UNSUPPORTED;
  const std::string key = mkkey("mixw", e.pname(), n);
  void* d = e.ctx->table(key);
  if (!d) {
    std::vector<C> h((size_t)n);
    for (i64 j = 0; j < n; ++j) h[(size_t)j] = unit_root<T>(j, n);
    JTB_TRY(e.ctx->put_table(key, h.data(), h.size() * sizeof(C), &d));
  }
  p.wtab = (const C*)d;
  p.n = (int)n; p.W = W; p.wfast = wfast ? 1 : 0;
  p.m_n = mix_magic((unsigned)n);
  p.logW = 0;
  while ((1 << p.logW) < W) ++p.logW;
  i64 ns = 1;
  for (int s = 0; s < p.nstages; ++s) {
    p.m_nb[s] = mix_magic((unsigned)(n / p.radix[s]));
    p.m_ns[s] = mix_magic((unsigned)ns);
    ns *= p.radix[s];
  }
  return ST_OK;
}
template <typename T> int mixed_launch(Engine<T>& e, MixedParams<T>& p) {
  typedef cx<T> C;
  const size_t smem = (size_t)p.W * 2 * (size_t)(p.n + 1) * sizeof(C);
  JTB_TRY(mixed_attr<T>(e));
  const i64 work = (i64)p.W * p.n / 4;
  const bool big = mixed_small_radices(p) && work >= 1024;
  const unsigned threads = big ? 1024u : (work >= 512 ? 512u : (work >= 256 ? 256u : (work >= 128 ? 128u : 64u)));
  const i64 nblk = (p.nlines - p.line_base + p.W - 1) / p.W;
  if (nblk > 0x7fffffffLL) { set_error("too many lines"); return ST_UNSUPPORTED; }
  if (big) JTB_LAUNCH((fft_mixed_kernel<T, true>), (unsigned)nblk, threads, smem, e.st, p);
  else JTB_LAUNCH((fft_mixed_kernel<T, false>), (unsigned)nblk, threads, smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  return ST_OK;
}
// widest power-of-two tile (lines per CTA) a length-n sub-transform can use
template <typename T> int mixed_width(i64 n, int wmax) {
  const size_t cap = (size_t)200 * 1024;
  int W = wmax;
  while (W >= 1 && (size_t)W * 2 * (size_t)(n + 1) * sizeof(cx<T>) > cap) W >>= 1;
  return W;
}
}  // namespace

template <typename T>
int mixed_twopass_contig(Engine<T>& e, cx<T>* a, i64 dist, i64 nlines, i64 n, bool inverse, bool has_scale, T scale,
                         bool* handled) {
  typedef cx<T> C;
  *handled = false;
  static const bool off = getenv("JTB_NO_MIXED") != nullptr || getenv("JTB_NO_MIXED2") != nullptr;
  if (off || nlines <= 0 || n < 64 || n >= (1LL << 31)) return ST_OK;
  {
    // Measured on B200 (profiles/r02_bench_misc.log): where the convolution length M = nextPow2(2n-1) is 2^19 or 2^20 the
    // fused three-pass Bluestein route is still ahead of the generic mixed-radix kernel (165375: 0.027 vs 0.037 ms,
    // 362880: 0.039 vs 0.051 ms); everywhere else the two mixed passes win (6250000: 0.45 vs 1.61 ms).  JTB_MIXED2_ALL=1
    // takes the mixed route regardless.
    static const bool all = getenv("JTB_MIXED2_ALL") != nullptr;
    const i64 M = next_pow2(2 * n - 1);
    if (!all && (M == (1LL << 19) || M == (1LL << 20))) return ST_OK;
  }
  int radix[MIX_MAX_STAGES * 4], nst = 0;
  {
    // full factorisation (more stages than one CTA may run is fine here: each half gets its own plan)
    static const int cand[] = {2, 3, 5, 7, 11, 13};
    i64 rem = n;
    for (int c : cand)
      while (rem % c == 0) { if (nst >= MIX_MAX_STAGES * 4) return ST_OK; radix[nst++] = c; rem /= c; }
    if (rem != 1) return ST_OK;
  }
  // split: enumerate the divisors reachable from the prime multiset; prefer wide tiles, then a balanced split
  const int wmax = (int)(128 / sizeof(C));
  i64 bestN1 = 0;
  double best = -1e300;
  std::vector<i64> divs(1, 1);
  for (int i = 0; i < nst; ++i) {
    const size_t m = divs.size();
    for (size_t j = 0; j < m; ++j) {
      const i64 d = divs[j] * radix[i];
      bool seen = false;
      for (i64 x : divs) if (x == d) { seen = true; break; }
      if (!seen) divs.push_back(d);
    }
  }
  for (i64 N1 : divs) {
    const i64 N2 = n / N1;
    if (N1 < 2 || N2 < 2) continue;
    const int W1 = mixed_width<T>(N1, wmax), W2 = mixed_width<T>(N2, wmax);
    if (W1 < 1 || W2 < 1) continue;
    int r1[MIX_MAX_STAGES], r2[MIX_MAX_STAGES], s1, s2;
    if (!mixed_factor(N1, r1, &s1) || !mixed_factor(N2, r2, &s2)) continue;
    const double score = 1000.0 * (W1 < W2 ? W1 : W2) + 10.0 * (W1 + W2) - std::fabs(std::log((double)N1 / (double)N2));
    if (score > best) { best = score; bestN1 = N1; }
  }
  if (!bestN1) return ST_OK;
  const i64 N1 = bestN1, N2 = n / N1;
  const int W1 = mixed_width<T>(N1, wmax), W2 = mixed_width<T>(N2, wmax);
  // four-step twiddle tables W_n^(L*h), W_n^l with L a power of two near sqrt(n)
  int logL = 0;
  while ((1LL << (2 * logL)) < n) ++logL;
  const i64 L = 1LL << logL, H = (n + L - 1) / L;
  const std::string ka = mkkey("mfsA", e.pname(), n), kb = mkkey("mfsB", e.pname(), n);
  void* da = e.ctx->table(ka);
  void* db = e.ctx->table(kb);
  if (!da || !db) {
    std::vector<C> ha((size_t)H), hb((size_t)L);
    for (i64 h = 0; h < H; ++h) ha[(size_t)h] = unit_root<T>(h * L, n);
    for (i64 l = 0; l < L; ++l) hb[(size_t)l] = unit_root<T>(l, n);
    JTB_TRY(e.ctx->put_table(ka, ha.data(), ha.size() * sizeof(C), &da));
    JTB_TRY(e.ctx->put_table(kb, hb.data(), hb.size() * sizeof(C), &db));
  }
  i64 chunk = (i64)(e.ctx->work_cap / ((size_t)n * sizeof(C)));
  if (chunk < 1) return ST_OK;
  if (chunk > nlines) chunk = nlines;
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_FOURSTEP], (size_t)chunk * (size_t)n * sizeof(C)));
  C* wk = (C*)e.ctx->work[WK_FOURSTEP].p;
  static const bool trace = getenv("JTB_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[jtb] mixed two-pass %s n=%lld = %lld x %lld, W = %d / %d\n", e.pname(), (long long)n, (long long)N1, (long long)N2, W1, W2);
  for (i64 c0 = 0; c0 < nlines; c0 += chunk) {
    const i64 cn = (c0 + chunk < nlines ? c0 + chunk : nlines) - c0;
    MixedParams<T> p;
    memset(&p, 0, sizeof p);
    JTB_TRY(mixed_setup<T>(e, p, N1, W1, true));
    p.in = a + c0 * dist; p.out = wk;
    p.gi = geo_make(N2, 1, dist, N2); p.go = geo_make(N2, 1, n, N2);
    p.nlines = cn * N2; p.line_base = 0;
    p.swap_in = inverse;
    p.tw_mode = 1; p.tw_mod = (int)N2; p.tw_logL = logL; p.twA = (const C*)da; p.twB = (const C*)db;
    JTB_TRY(mixed_launch<T>(e, p));
    MixedParams<T> q;
    memset(&q, 0, sizeof q);
    JTB_TRY(mixed_setup<T>(e, q, N2, W2, false));
    q.in = wk; q.out = a + c0 * dist;
    q.gi = geo_make(N1, N2, n, 1); q.go = geo_make(N1, 1, dist, N1);
    q.nlines = cn * N1; q.line_base = 0;
    q.wfast_out = W2 > 1 ? 1 : 0;
    q.swap_out = inverse; q.has_scale = has_scale; q.scale = scale;
    JTB_TRY(mixed_launch<T>(e, q));
  }
  *handled = true;
  return ST_OK;
}
template int mixed_twopass_contig<double>(Engine<double>&, double2*, i64, i64, i64, bool, bool, double, bool*);
template int mixed_twopass_contig<float>(Engine<float>&, float2*, i64, i64, i64, bool, bool, float, bool*);

template int mixed_c2c<double>(Engine<double>&, double2*, const Geo&, i64, i64, bool, bool, double, bool*);
template int mixed_c2c<float>(Engine<float>&, float2*, const Geo&, i64, i64, bool, bool, float, bool*);

}  // namespace jtb
