// Instantiations and host-side dispatch of fft_fast2_kernel (jtb_fast2.cuh): the two-pass (four-step) transform of
// long contiguous lines, of long strided lines (in place, intermediate in the array's own layout) and the fused
// real-forward row pass.
#include <cstdlib>

#include "jtb_engine_impl.cuh"
#include "jtb_fast2.cuh"

namespace jtb {

namespace {

template <typename T> struct F2Entry {
  int logn, loge, sin, mode, W, threads, smem;
  void (*kern)(const Fast2Params<T>);
  unsigned attr_done;   // bit d set: smem attribute applied on device d
  i64 min_lines;             // use this variant only for launches with at least this many lines (grid fill)
  const cx<T>* twg[16];      // per-device base twiddle table
};
template <typename T, int LOGN, int LOGE, bool SIN, int MODE, int W> F2Entry<T> mk2(i64 min_lines = 0) {
  typedef Sched<LOGN, LOGE> S;
  F2Entry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.sin = SIN; e.mode = MODE; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, SIN, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_fast2_kernel<T, LOGN, LOGE, SIN, MODE, W>;
  e.attr_done = 0;
  e.min_lines = min_lines;
  for (int d = 0; d < 16; ++d) e.twg[d] = nullptr;
  return e;
}
template <typename T> std::vector<F2Entry<T>>& reg2();
template <> std::vector<F2Entry<double>>& reg2<double>() {
  static std::vector<F2Entry<double>> r = {
      // first pass of the two-pass transform (strided lines, twiddle at the store)
      mk2<double, 6, 3, true, FM_TWID, 32>(), mk2<double, 7, 4, true, FM_TWID, 16>(), mk2<double, 8, 4, true, FM_TWID, 8>(),
      mk2<double, 9, 3, true, FM_TWID, 8>(), mk2<double, 10, 4, true, FM_TWID, 8>(),
      // half-width variants for launches that would not fill the GPU otherwise (a single 2^20-point transform is
      // 1024 lines: 128 CTAs at W = 8, 256 at W = 4) -- find2 picks by CTA count
      mk2<double, 9, 3, true, FM_TWID, 4>(), mk2<double, 10, 4, true, FM_TWID, 4>(),
      mk2<double, 9, 3, false, FM_TRANSPOSE, 4>(), mk2<double, 10, 4, false, FM_TRANSPOSE, 4>(),
      // second pass, contiguous lines with transposed store
      mk2<double, 8, 4, false, FM_TRANSPOSE, 8>(), mk2<double, 9, 3, false, FM_TRANSPOSE, 8>(),
      mk2<double, 10, 4, false, FM_TRANSPOSE, 8>(),
      mk2<double, 11, 4, false, FM_TRANSPOSE, 4>(),
      // second pass, strided lines (row permutation only)
      mk2<double, 6, 3, true, FM_PLAIN, 32>(), mk2<double, 7, 4, true, FM_PLAIN, 16>(), mk2<double, 5, 3, true, FM_PLAIN, 32>(),
      // real-forward rows
      mk2<double, 5, 3, false, FM_RFFT, 32>(), mk2<double, 6, 3, false, FM_RFFT, 16>(), mk2<double, 7, 4, false, FM_RFFT, 16>(),
      mk2<double, 8, 4, false, FM_RFFT, 8>(), mk2<double, 9, 3, false, FM_RFFT, 4>(), mk2<double, 10, 3, false, FM_RFFT, 2>(),
      mk2<double, 11, 3, false, FM_RFFT, 1>(), mk2<double, 12, 3, false, FM_RFFT, 1>(), mk2<double, 11, 4, false, FM_RFFT, 2>(),
  };
  return r;
}
template <> std::vector<F2Entry<float>>& reg2<float>() {
  static std::vector<F2Entry<float>> r = {
      mk2<float, 6, 3, true, FM_TWID, 32>(), mk2<float, 7, 4, true, FM_TWID, 16>(),
      mk2<float, 9, 3, true, FM_TWID, 16>(), mk2<float, 10, 4, true, FM_TWID, 16>(),
      mk2<float, 5, 3, true, FM_PLAIN, 32>(), mk2<float, 6, 3, true, FM_PLAIN, 32>(), mk2<float, 7, 4, true, FM_PLAIN, 16>(),
      mk2<float, 9, 3, false, FM_TRANSPOSE, 16>(), mk2<float, 10, 4, false, FM_TRANSPOSE, 16>(),
      mk2<float, 11, 4, false, FM_TRANSPOSE, 8>(),
      mk2<float, 5, 3, false, FM_RFFT, 32>(), mk2<float, 6, 3, false, FM_RFFT, 32>(), mk2<float, 7, 4, false, FM_RFFT, 16>(),
      mk2<float, 8, 4, false, FM_RFFT, 16>(), mk2<float, 9, 3, false, FM_RFFT, 8>(),
      mk2<float, 10, 4, false, FM_RFFT, 4>(), mk2<float, 11, 4, false, FM_RFFT, 2>(), mk2<float, 12, 4, false, FM_RFFT, 1>(),
  };
  return r;
}

// variant for `lines` lines: the widest one (most lines per CTA, longest segments) that still gives every SM two CTAs;
// when none does, the narrowest (most CTAs).  JTB_FS_W forces a width where it exists.
template <typename T> F2Entry<T>* find2(int logn, bool sin, int mode, i64 lines = (1LL << 60)) {
  static const int force_w = getenv("JTB_FS_W") ? atoi(getenv("JTB_FS_W")) : 0;
  F2Entry<T>*wide = nullptr, *narrow = nullptr, *forced = nullptr;
  for (auto& f : reg2<T>()) {
    if (f.logn != logn || (f.sin != 0) != sin || f.mode != mode || lines < f.min_lines) continue;
    if (mode != FM_TWID && mode != FM_TRANSPOSE) return &f;   // the other families list their default first
    if (force_w && f.W == force_w) forced = &f;
    if ((lines + f.W - 1) / f.W >= 2 * 148 && (!wide || f.W > wide->W)) wide = &f;
    if (!narrow || f.W < narrow->W) narrow = &f;
  }
  if (forced) return forced;
  return wide ? wide : narrow;
}

template <typename T> int launch2(Engine<T>& e, F2Entry<T>* f, Fast2Params<T>& p) {
  if (!(f->attr_done & (1u << (e.ctx->device & 31)))) {
    JTB_CUDA(cudaFuncSetAttribute(f->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, f->smem));
    f->attr_done |= 1u << (e.ctx->device & 31);
  }
  const int dv = e.ctx->device & 15;
  if (!f->twg[dv]) JTB_TRY(fast_stage_table<T>(e, f->logn, f->loge, &f->twg[dv]));
  p.twg = f->twg[dv];
  const i64 nblk = (p.nlines + f->W - 1) / f->W;
  if (nblk > 0x7fffffffLL) { set_error("too many lines"); return ST_UNSUPPORTED; }
  static const bool trace = getenv("JTB_TRACE") != nullptr;
  if (trace) fprintf(stderr, "[jtb] fast2 %s logn=%d sin=%d mode=%d W=%d lines=%lld\n", e.pname(), f->logn, f->sin, f->mode, f->W, (long long)p.nlines);
  JTB_LAUNCH(f->kern, (unsigned)nblk, (unsigned)f->threads, (size_t)f->smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  return ST_OK;
}

template <typename T> Fast2Params<T> blank2() {
  Fast2Params<T> p;
  memset(&p, 0, sizeof p);
  p.gmod = 1 << 30;
  p.c0 = 1;
  p.scale = 1;
  return p;
}

const bool g_fast2_off = getenv("JTB_NO_FAST2") != nullptr;

}  // namespace

// two-pass transform of contiguous lines: line l of `in` at l*in_dist -> line l of `out` at l*out_dist
template <typename T>
int fast_fourstep_contig(Engine<T>& e, const cx<T>* in, i64 in_dist, cx<T>* out, i64 out_dist, i64 l0, i64 l1, int logn,
                         bool swap_in, bool swap_out, bool has_scale, T scale, bool* handled) {
  typedef cx<T> C;
  *handled = false;
  if (g_fast2_off || l1 <= l0) return ST_OK;
  // split n = N1 (strided first pass) * N2 (contiguous second pass)
  F2Entry<T>*f1 = nullptr, *f2 = nullptr;
  static const int force_la = getenv("JTB_FS_LA") ? atoi(getenv("JTB_FS_LA")) : 0;   // log2 of the first-pass length
  if (force_la > 0 && force_la < logn) {
    F2Entry<T>* x = find2<T>(force_la, true, FM_TWID, (l1 - l0) << (logn - force_la));
    F2Entry<T>* y = find2<T>(logn - force_la, false, FM_TRANSPOSE, (l1 - l0) << force_la);
    if (x && y) { f1 = x; f2 = y; }
  }
  for (int la = logn / 2; la >= 6 && !f1; --la) {
    for (int s = 0; s < 2 && !f1; ++s) {
      const int a = s == 0 ? la : logn - la;   // try the balanced split first, then its mirror
      F2Entry<T>* x = find2<T>(a, true, FM_TWID, (l1 - l0) << (logn - a));
      F2Entry<T>* y = find2<T>(logn - a, false, FM_TRANSPOSE, (l1 - l0) << a);
      if (x && y) { f1 = x; f2 = y; }
    }
  }
  if (!f1) return ST_OK;
  const i64 n = 1LL << logn, N1 = 1LL << f1->logn, N2 = 1LL << f2->logn;
  if (N2 % f1->W || N1 % f2->W) return ST_OK;
  const C *fsA, *fsB;
  int logL;
  JTB_TRY(e.fs_tables(logn, &fsA, &fsB, &logL));
  i64 chunk = (i64)(e.ctx->work_cap / ((size_t)n * sizeof(C)));
  if (chunk < 1) return ST_OK;
  if (chunk > l1 - l0) chunk = l1 - l0;
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_FOURSTEP], (size_t)chunk * (size_t)n * sizeof(C)));
  C* wk = (C*)e.ctx->work[WK_FOURSTEP].p;
  for (i64 c0 = l0; c0 < l1; c0 += chunk) {
    const i64 c1 = c0 + chunk < l1 ? c0 + chunk : l1;
    Fast2Params<T> p = blank2<T>();
    p.in = in + c0 * in_dist; p.out = wk;
    p.nlines = (c1 - c0) * N2; p.c0 = (int)N2;
    p.in_gdist = in_dist; p.in_cdist = 1; p.in_stride = N2;
    p.out_gdist = n; p.out_cdist = 1; p.out_stride = N2;
    p.swap_in = swap_in;
    p.fsA = fsA; p.fsB = fsB; p.fs_logL = logL; p.tw_src = 0;
    JTB_TRY(launch2(e, f1, p));
    Fast2Params<T> q = blank2<T>();
    q.in = wk; q.out = out + c0 * out_dist;
    q.nlines = (c1 - c0) * N1; q.c0 = (int)N1;
    q.in_gdist = n; q.in_cdist = N2; q.in_stride = 1;
    q.out_gdist = out_dist; q.out_cdist = 1; q.out_stride = N1;
    q.swap_out = swap_out; q.has_scale = has_scale; q.scale = scale;
    JTB_TRY(launch2(e, f2, q));
  }
  *handled = true;
  return ST_OK;
}

// two-pass in-place transform of long strided lines: c0 adjacent lines (distance 1), element stride s, groups d3
template <typename T>
int fast_fourstep_strided(Engine<T>& e, cx<T>* a, const Geo& g, i64 nlines, int logn, bool inverse, bool has_scale,
                          T scale, bool* handled) {
  typedef cx<T> C;
  *handled = false;
  if (g_fast2_off || nlines <= 0) return ST_OK;
  if (!(g.stride > 1 && g.d[0] == 1 && g.c[1] == 1 && g.c[2] == 1 && g.c[0] > 1 && nlines % g.c[0] == 0)) return ST_OK;
  F2Entry<T>*f1 = nullptr, *f2 = nullptr;
  for (int la = (logn + 1) / 2; la <= logn - 5 && !f1; ++la) {
    F2Entry<T>* x = find2<T>(la, true, FM_TWID);
    F2Entry<T>* y = find2<T>(logn - la, true, FM_PLAIN);
    if (x && y) { f1 = x; f2 = y; }
  }
  if (!f1) return ST_OK;
  const i64 n = 1LL << logn, R1 = 1LL << f1->logn, R2 = 1LL << f2->logn, s = g.stride;
  if (g.c[0] % f1->W || g.c[0] % f2->W) return ST_OK;
  if ((n - 1) * s + g.c[0] >= (1LL << 40)) return ST_OK;
  const i64 batches = nlines / g.c[0];
  const i64 ext = geo_extent(g, nlines, n);
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_FOURSTEP], (size_t)ext * sizeof(C)));
  C* wk = (C*)e.ctx->work[WK_FOURSTEP].p;
  const C *fsA, *fsB;
  int logL;
  JTB_TRY(e.fs_tables(logn, &fsA, &fsB, &logL));
  // Column strips small enough that the intermediate written by pass 1 is still in L2 (126 MB) when pass 2
  // reads it: the two passes then cost about one sweep of HBM traffic instead of two.
  // (every strip reuses the same compact [n][cb] work block: see fast_r2r_cols)
  i64 cb = g.c[0];
  {
    const char* ev = getenv("JTB_C2C_STRIP_MB") ? getenv("JTB_C2C_STRIP_MB") : getenv("JTB_STRIP_MB");
    const double strip_mb = ev ? atof(ev) : 0.0;
    const i64 wmax = f1->W > f2->W ? f1->W : f2->W;
    if (strip_mb > 0)
      while (cb > wmax && (cb % 2) == 0 && (double)cb * (double)n * batches * sizeof(C) > strip_mb * 1048576.0) cb /= 2;
    if (cb % f1->W || cb % f2->W) cb = g.c[0];
  }
  const bool compact = cb < g.c[0];
  const i64 ws = compact ? cb : s, wbd = compact ? n * cb : g.d[3];   // row / batch distance inside the work block
  for (i64 cs = 0; cs < g.c[0]; cs += cb) {
    C* wks = compact ? wk : wk + cs;
    // pass 1: lines (c, r2, batch): FFT over r1 (element stride R2*s), twiddle W_n^(k1*r2), a -> work
    Fast2Params<T> p = blank2<T>();
    p.in = a + cs; p.out = wks;
    p.nlines = cb * R2 * batches; p.c0 = (int)cb; p.gmod = (int)R2;
    p.in_gdist = s; p.in_gdist2 = g.d[3]; p.in_cdist = 1; p.in_stride = R2 * s;
    p.out_gdist = ws; p.out_gdist2 = wbd; p.out_cdist = 1; p.out_stride = R2 * ws;
    p.swap_in = inverse;
    p.fsA = fsA; p.fsB = fsB; p.fs_logL = logL; p.tw_src = 1;
    JTB_TRY(launch2(e, f1, p));
    // pass 2: lines (c, k1, batch): FFT over r2 (rows k1*R2 + r2), output element k2 -> row k1 + R1*k2, work -> a
    Fast2Params<T> q = blank2<T>();
    q.in = wks; q.out = a + cs;
    q.nlines = cb * R1 * batches; q.c0 = (int)cb; q.gmod = (int)R1;
    q.in_gdist = R2 * ws; q.in_gdist2 = wbd; q.in_cdist = 1; q.in_stride = ws;
    q.out_gdist = s; q.out_gdist2 = g.d[3]; q.out_cdist = 1; q.out_stride = R1 * s;
    q.swap_out = inverse; q.has_scale = has_scale; q.scale = scale;
    JTB_TRY(launch2(e, f2, q));
  }
  *handled = true;
  return ST_OK;
}

// realForward of contiguous lines of 2N reals (N = 2^logN complex points), in place, JTransforms packing
template <typename T>
int fast_rfft_fwd(Engine<T>& e, cx<T>* a, i64 dist, i64 nlines, int logN, bool* handled) {
  *handled = false;
  if (g_fast2_off || nlines <= 0) return ST_OK;
  F2Entry<T>* f = find2<T>(logN, false, FM_RFFT);
  static const char* ele = getenv("JTB_ROW_LOGE");
  if (ele) for (auto& x : reg2<T>()) if (x.logn == logN && !x.sin && x.mode == FM_RFFT && x.loge == atoi(ele)) { f = &x; break; }
  if (!f) return ST_OK;
  const cx<T>* tw[JTB_MAX_STAGES];
  const cx<T>* rtw;
  JTB_TRY(e.tile_tables(logN, tw, &rtw));
  Fast2Params<T> p = blank2<T>();
  p.in = a; p.out = a; p.nlines = nlines;
  p.in_gdist = p.out_gdist = dist; p.in_stride = p.out_stride = 1;
  p.rtw = rtw;
  {
    static const char* epf = getenv("JTB_ROW_PREFETCH");   // CTAs ahead (0 = off)
    p.prefetch = epf ? atoi(epf) : 0;
  }
  JTB_TRY(launch2(e, f, p));
  *handled = true;
  return ST_OK;
}

// ------------------------------------------------------------------------------------------ three-pass 1-D
namespace {
template <typename T, int LOGN, int LOGE, int W> F2Entry<T> mkbig() {
  typedef Sched<LOGN, LOGE> S;
  F2Entry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.sin = 1; e.mode = FM_PLAIN; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, true, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_fast2_kernel<T, LOGN, LOGE, true, FM_PLAIN, W, PRE_BIGTW>;
  e.attr_done = 0; e.min_lines = 0;
  for (int d = 0; d < 16; ++d) e.twg[d] = nullptr;
  return e;
}
template <typename T> std::vector<F2Entry<T>>& bigreg();
template <> std::vector<F2Entry<double>>& bigreg<double>() {
  static std::vector<F2Entry<double>> r = {mkbig<double, 5, 3, 32>(), mkbig<double, 6, 3, 32>(), mkbig<double, 7, 4, 16>()};
  return r;
}
template <> std::vector<F2Entry<float>>& bigreg<float>() {
  static std::vector<F2Entry<float>> r = {mkbig<float, 5, 3, 32>(), mkbig<float, 6, 3, 32>(), mkbig<float, 7, 4, 16>()};
  return r;
}
}  // namespace

// Contiguous lines of n = N1*N2 points beyond the lean two-pass range, as THREE sweeps: view the line as [N1][N2];
//   A: strided transforms over the first factor R1 of N1 = R1*R2 with the inner twiddle (a -> work 1),
//   B: second factor R2, rows back in natural order k1, times the outer twiddle W_n^(k1*n2) (work 1 -> work 2),
//   C: contiguous transforms of length N2 with the transposed store k1 + N1*k2 (work 2 -> a).
template <typename T>
int fast_threepass_contig(Engine<T>& e, cx<T>* a, i64 dist, i64 l0, i64 l1, int logn, bool inverse, bool has_scale,
                          T scale, bool* handled) {
  typedef cx<T> C;
  *handled = false;
  if (g_fast2_off || l1 <= l0 || getenv("JTB_NO_THREEPASS")) return ST_OK;
  F2Entry<T>*fa = nullptr, *fb = nullptr, *fc = nullptr;
  for (int l2 = 11; l2 >= 8 && !fa; --l2) {
    F2Entry<T>* z = find2<T>(l2, false, FM_TRANSPOSE);
    if (!z) continue;
    const int l1g = logn - l2;
    for (int la = (l1g + 1) / 2; la <= l1g - 5 && !fa; ++la) {
      F2Entry<T>* x = find2<T>(la, true, FM_TWID);
      F2Entry<T>* y = nullptr;
      for (auto& b : bigreg<T>()) if (b.logn == l1g - la) y = &b;
      if (x && y) { fa = x; fb = y; fc = z; }
    }
  }
  if (!fa) return ST_OK;
  const i64 n = 1LL << logn, N2 = 1LL << fc->logn, N1 = n >> fc->logn, R1 = 1LL << fa->logn, R2 = 1LL << fb->logn;
  if (N2 % fa->W || N2 % fb->W || N1 % fc->W || n > (1LL << 30)) return ST_OK;
  const C *inA, *inB, *bgA, *bgB;
  int inL, bgL;
  JTB_TRY(e.fs_tables(ilog2(N1), &inA, &inB, &inL));
  JTB_TRY(e.fs_tables(logn, &bgA, &bgB, &bgL));
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_FOURSTEP], (size_t)n * sizeof(C)));
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_BIG], (size_t)n * sizeof(C)));
  C* w1 = (C*)e.ctx->work[WK_FOURSTEP].p;
  C* w2 = (C*)e.ctx->work[WK_BIG].p;
  for (i64 l = l0; l < l1; ++l) {
    C* base = a + l * dist;
    Fast2Params<T> p = blank2<T>();
    p.in = base; p.out = w1;
    p.nlines = N2 * R2; p.c0 = (int)N2; p.gmod = (int)R2;
    p.in_gdist = N2; p.in_cdist = 1; p.in_stride = R2 * N2;
    p.out_gdist = N2; p.out_cdist = 1; p.out_stride = R2 * N2;
    p.swap_in = inverse;
    p.fsA = inA; p.fsB = inB; p.fs_logL = inL; p.tw_src = 1;
    JTB_TRY(launch2(e, fa, p));
    Fast2Params<T> q = blank2<T>();
    q.in = w1; q.out = w2;
    q.nlines = N2 * R1; q.c0 = (int)N2; q.gmod = (int)R1;
    q.in_gdist = R2 * N2; q.in_cdist = 1; q.in_stride = N2;
    q.out_gdist = N2; q.out_cdist = 1; q.out_stride = R1 * N2;
    q.bigA = bgA; q.bigB = bgB; q.big_logL = bgL;
    JTB_TRY(launch2(e, fb, q));
    Fast2Params<T> r = blank2<T>();
    r.in = w2; r.out = base;
    r.nlines = N1; r.c0 = (int)N1;
    r.in_gdist = n; r.in_cdist = N2; r.in_stride = 1;
    r.out_gdist = dist; r.out_cdist = 1; r.out_stride = N1;
    r.swap_out = inverse; r.has_scale = has_scale; r.scale = scale;
    JTB_TRY(launch2(e, fc, r));
  }
  *handled = true;
  return ST_OK;
}

// ------------------------------------------------------------------------------------------ fused DCT/DST/DHT
namespace {
template <typename T> struct RowEntry {
  int logn, loge, kind, W, threads, smem;
  void (*kern)(const RowR2RParams<T>);
  unsigned attr_done;   // bit d set: smem attribute applied on device d
};
template <typename T, int LOGN, int LOGE, int KIND, int W> RowEntry<T> mkrow() {
  typedef Sched<LOGN, LOGE> S;
  RowEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.kind = KIND; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, false, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_r2r_row_kernel<T, LOGN, LOGE, KIND, W>;
  e.attr_done = 0;
  return e;
}
#define JTB_ROWS(T, LOGN, LOGE, W) mkrow<T, LOGN, LOGE, RK_DCT, W>(), mkrow<T, LOGN, LOGE, RK_DST, W>(), mkrow<T, LOGN, LOGE, RK_DHT, W>()
template <typename T> std::vector<RowEntry<T>>& rowreg();
template <> std::vector<RowEntry<double>>& rowreg<double>() {
  static std::vector<RowEntry<double>> r = {mkrow<double, 12, 4, RK_DHT, 2>(),   // row pairs (fast_dht2d_rows); plain rows skip it (W = 1 first)
                                            JTB_ROWS(double, 12, 4, 1), JTB_ROWS(double, 12, 3, 1), JTB_ROWS(double, 11, 3, 1), JTB_ROWS(double, 11, 4, 2), JTB_ROWS(double, 10, 3, 2),
                                            JTB_ROWS(double, 9, 3, 4), JTB_ROWS(double, 8, 4, 8), JTB_ROWS(double, 7, 4, 16),
                                            JTB_ROWS(double, 6, 3, 16), JTB_ROWS(double, 5, 3, 32)};
  return r;
}
template <> std::vector<RowEntry<float>>& rowreg<float>() {
  static std::vector<RowEntry<float>> r = {mkrow<float, 12, 4, RK_DHT, 2>(), JTB_ROWS(float, 12, 4, 1), JTB_ROWS(float, 11, 4, 2), JTB_ROWS(float, 10, 4, 4),
                                           JTB_ROWS(float, 9, 3, 8),  JTB_ROWS(float, 8, 4, 16), JTB_ROWS(float, 7, 4, 16),
                                           JTB_ROWS(float, 6, 3, 32), JTB_ROWS(float, 5, 3, 32)};
  return r;
}

// permuted first passes of the strided DCT/DST columns
template <typename T> struct PreEntry { int logn, pre; F2Entry<T> e; };
template <typename T, int LOGN, int LOGE, int W, int PRE> PreEntry<T> mkpre() {
  typedef Sched<LOGN, LOGE> S;
  PreEntry<T> x;
  x.logn = LOGN; x.pre = PRE;
  F2Entry<T>& e = x.e;
  e.logn = LOGN; e.loge = LOGE; e.sin = 1; e.mode = FM_TWID; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, true, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_fast2_kernel<T, LOGN, LOGE, true, FM_TWID, W, PRE>;
  e.attr_done = 0; e.min_lines = 0;
  for (int d = 0; d < 16; ++d) e.twg[d] = nullptr;
  return x;
}
template <typename T> std::vector<PreEntry<T>>& prereg();
template <> std::vector<PreEntry<double>>& prereg<double>() {
  static std::vector<PreEntry<double>> r = {mkpre<double, 6, 3, 32, PRE_PERM_DCT>(), mkpre<double, 6, 3, 32, PRE_PERM_DST>(),
                                            mkpre<double, 7, 4, 16, PRE_PERM_DCT>(), mkpre<double, 7, 4, 16, PRE_PERM_DST>()};
  return r;
}
template <> std::vector<PreEntry<float>>& prereg<float>() {
  static std::vector<PreEntry<float>> r = {mkpre<float, 6, 3, 32, PRE_PERM_DCT>(), mkpre<float, 6, 3, 32, PRE_PERM_DST>(),
                                           mkpre<float, 7, 4, 16, PRE_PERM_DCT>(), mkpre<float, 7, 4, 16, PRE_PERM_DST>()};
  return r;
}
}  // namespace

namespace {
template <typename T> struct PairEntry {
  int logn, loge, W, threads, smem;
  void (*kern)(const ColPairParams<T>);
  unsigned attr_done;   // bit d set: smem attribute applied on device d
};
template <typename T, int LOGN, int LOGE, int W> PairEntry<T> mkpair() {
  typedef Sched<LOGN, LOGE> S;
  PairEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.W = W; e.threads = 2 * W * S::TPL;
  e.smem = (int)((FastAddr<T, S, true, 2 * W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_colpair_kernel<T, LOGN, LOGE, W>;
  e.attr_done = 0;
  return e;
}
template <typename T> std::vector<PairEntry<T>>& pairreg();
template <> std::vector<PairEntry<double>>& pairreg<double>() {
  static std::vector<PairEntry<double>> r = {mkpair<double, 6, 3, 32>(), mkpair<double, 7, 4, 16>(), mkpair<double, 5, 3, 32>()};
  return r;
}
template <> std::vector<PairEntry<float>>& pairreg<float>() {
  static std::vector<PairEntry<float>> r = {mkpair<float, 6, 3, 32>(), mkpair<float, 7, 4, 16>(), mkpair<float, 5, 3, 32>()};
  return r;
}

}  // namespace

// forward DCT-II / DST-II / DHT of contiguous real lines (line l at l*dist), in place
template <typename T>
int fast_r2r_rows(Engine<T>& e, T* a, i64 dist, i64 nlines, i64 n, int kind, T f0, T f, bool* handled) {
  *handled = false;
  if (g_fast2_off || nlines <= 0 || !is_pow2(n) || n < 4 || (dist % 2) || ((uintptr_t)a % sizeof(cx<T>))) return ST_OK;
  if (kind != RK_DHT && ((dist % 4) || ((uintptr_t)a % (4 * sizeof(T))))) return ST_OK;   // 4-real vector loads
  const int logN = ilog2(n) - 1;
  RowEntry<T>* r = nullptr;
  static const char* ele = getenv("JTB_ROW_LOGE");   // tuning knob: prefer the variant with this radix
  const int want_loge = ele ? atoi(ele) : 0;
  const auto pair_only = [](const RowEntry<T>& x) { return x.logn == 12 && x.W == 2; };   // registered for fast_dht2d_rows
  for (auto& x : rowreg<T>()) if (x.logn == logN && x.kind == kind && !pair_only(x) && (!want_loge || x.loge == want_loge)) { r = &x; break; }
  if (!r) for (auto& x : rowreg<T>()) if (x.logn == logN && x.kind == kind && !pair_only(x)) { r = &x; break; }
  if (!r) return ST_OK;
  if (!(r->attr_done & (1u << (e.ctx->device & 31)))) {
    JTB_CUDA(cudaFuncSetAttribute(r->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, r->smem));
    r->attr_done |= 1u << (e.ctx->device & 31);
  }
  RowR2RParams<T> p;
  p.a = a; p.nlines = nlines; p.dist = dist; p.f0 = f0; p.f = f; p.pair_rows = 0;
  {
    // L2 prefetch of the rows 74 CTAs ahead: 8192^2 DCT 0.618 -> 0.607 ms (profiles/r02_ab_rowprefetch.log); 0 = off
    static const char* epf = getenv("JTB_ROW_PREFETCH");
    p.prefetch = epf ? atoi(epf) : 74;
  }
  JTB_TRY(fast_stage_table<T>(e, logN, r->loge, &p.twg));
  const cx<T>* tw[JTB_MAX_STAGES];
  JTB_TRY(e.tile_tables(logN, tw, &p.rtw));
  p.dtw = nullptr;
  if (kind != RK_DHT) JTB_TRY(e.dct_table(n, &p.dtw));
  const i64 nblk = (nlines + r->W - 1) / r->W;
  if (nblk > 0x7fffffffLL) return ST_OK;
  JTB_LAUNCH(r->kern, (unsigned)nblk, (unsigned)r->threads, (size_t)r->smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}

// Last pass of DoubleDHT_2D.forward / inverse on one R x n array: the row DHTs with yTransform
// (dht/DoubleDHT_2D.java:1288-1309) folded into the store -- a CTA owns the row pairs (r, R - r), so the separate
// yTransform sweep (one more read + write of the array) disappears.
template <typename T>
int fast_dht2d_rows(Engine<T>& e, T* a, i64 R, i64 n, T f, bool* handled) {
  *handled = false;
  static const bool off = getenv("JTB_NO_DHTFOLD") != nullptr;
  if (g_fast2_off || off || R < 2 || (R % 2) || !is_pow2(n) || n < 4 || ((uintptr_t)a % sizeof(cx<T>))) return ST_OK;
  const int logN = ilog2(n) - 1;
  RowEntry<T>* r = nullptr;
  for (auto& x : rowreg<T>()) if (x.logn == logN && x.kind == RK_DHT && x.W >= 2 && (x.W % 2) == 0) { r = &x; break; }
  if (!r) return ST_OK;
  if (!(r->attr_done & (1u << (e.ctx->device & 31)))) {
    JTB_CUDA(cudaFuncSetAttribute(r->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, r->smem));
    r->attr_done |= 1u << (e.ctx->device & 31);
  }
  RowR2RParams<T> p;
  p.a = a; p.nlines = R; p.dist = n; p.f0 = f; p.f = f; p.pair_rows = R; p.prefetch = 0;
  JTB_TRY(fast_stage_table<T>(e, logN, r->loge, &p.twg));
  const cx<T>* tw[JTB_MAX_STAGES];
  JTB_TRY(e.tile_tables(logN, tw, &p.rtw));
  p.dtw = nullptr;
  const i64 pairs = R / 2, per = r->W / 2;
  const i64 nblk = (pairs + per - 1) / per;
  if (nblk > 0x7fffffffLL) return ST_OK;
  JTB_LAUNCH(r->kern, (unsigned)nblk, (unsigned)r->threads, (size_t)r->smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}
template int fast_dht2d_rows<double>(Engine<double>&, double*, i64, i64, double, bool*);
template int fast_dht2d_rows<float>(Engine<float>&, float*, i64, i64, float, bool*);

// forward DCT-II / DST-II / DHT along the strided axis of length n of `batches` row-major [n][Cn] real arrays
// (batch distance bdist reals): two adjacent real columns = one complex column, two-pass complex FFT with the
// Makhoul row permutation fused into the first pass, then the pair post-pass; column strips sized for L2.
template <typename T>
int fast_r2r_cols(Engine<T>& e, T* a, i64 n, i64 Cn, i64 batches, i64 bdist, int kind, T f0, T f, bool* handled) {
  typedef cx<T> C;
  *handled = false;
  if (g_fast2_off || !is_pow2(n) || (Cn % 2) || (bdist % 2) || ((uintptr_t)a % sizeof(C)) || batches < 1) return ST_OK;
  JTB_TRY(fast_r2r_cols_single<T>(e, a, n, Cn, batches, bdist, kind, false, f0, f, handled));   // n <= 1024: one pass
  if (*handled) return ST_OK;
  const int logn = ilog2(n);
  F2Entry<T>*f1 = nullptr, *f2 = nullptr;
  for (int la = (logn + 1) / 2; la <= logn - 5 && !f1; ++la) {
    F2Entry<T>* x = nullptr;
    if (kind == RK_DHT) x = find2<T>(la, true, FM_TWID);
    else for (auto& pe : prereg<T>()) if (pe.logn == la && pe.pre == (kind == RK_DCT ? PRE_PERM_DCT : PRE_PERM_DST)) x = &pe.e;
    F2Entry<T>* y = find2<T>(logn - la, true, FM_PLAIN);
    if (x && y) { f1 = x; f2 = y; }
  }
  if (!f1) return ST_OK;
  const i64 H = Cn / 2, s = H, R1 = 1LL << f1->logn, R2 = 1LL << f2->logn, bd = bdist / 2;
  if (H % f1->W || H % f2->W) return ST_OK;
  PairEntry<T>* fp = nullptr;
  if (!getenv("JTB_NO_COLPAIR"))
    for (auto& x : pairreg<T>()) if (x.logn == f2->logn && H % x.W == 0) { fp = &x; break; }
  const i64 ext = (batches - 1) * bd + n * s;
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_FOURSTEP], (size_t)ext * sizeof(C)));
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_REAL], (size_t)ext * sizeof(C)));
  C* wk = (C*)e.ctx->work[WK_FOURSTEP].p;
  C* wk2 = (C*)e.ctx->work[WK_REAL].p;
  C* ac = (C*)a;
  const C *fsA, *fsB, *dtw = nullptr;
  int logL;
  JTB_TRY(e.fs_tables(logn, &fsA, &fsB, &logL));
  if (kind != RK_DHT) JTB_TRY(e.dct_table(n, &dtw));
  // Column strips small enough that the intermediate written by the first pass is still in L2 when the second pass
  // reads it; every strip reuses the SAME compact [n][cb] work block, so its dirty lines are overwritten in L2 instead
  // of being written back: the two passes then cost one sweep of HBM traffic (JTB_STRIP_MB = strip size, 0 = off).
  i64 cb = H;
  {
    const char* ev = getenv("JTB_R2R_STRIP_MB") ? getenv("JTB_R2R_STRIP_MB") : getenv("JTB_STRIP_MB");
    const double strip_mb = ev ? atof(ev) : 0.0;
    const i64 wmax = f1->W > f2->W ? f1->W : f2->W;
    if (strip_mb > 0 && fp)
      while (cb > wmax && (cb % 2) == 0 && (double)cb * (double)n * batches * sizeof(C) > strip_mb * 1048576.0) cb /= 2;
    if (cb % f1->W || cb % f2->W || (fp && cb % fp->W)) cb = H;
  }
  const bool compact = cb < H && fp;
  const i64 ws = compact ? cb : s, wbd = compact ? n * cb : bd;   // row / batch distance inside the work block
  for (i64 cs = 0; cs < H; cs += cb) {
    C* wks = compact ? wk : wk + cs;
    Fast2Params<T> p = blank2<T>();
    p.in = ac + cs; p.out = wks;
    p.nlines = cb * R2 * batches; p.c0 = (int)cb; p.gmod = (int)R2;
    p.in_gdist = s; p.in_gdist2 = bd; p.in_cdist = 1; p.in_stride = R2 * s;
    p.out_gdist = ws; p.out_gdist2 = wbd; p.out_cdist = 1; p.out_stride = R2 * ws;
    p.fsA = fsA; p.fsB = fsB; p.fs_logL = logL; p.tw_src = 1;
    p.pre_n = n; p.pre_s = s;
    JTB_TRY(launch2(e, f1, p));
    if (fp) {
      // second pass + pair post-pass in one kernel, writing the final result
      if (!(fp->attr_done & (1u << (e.ctx->device & 31)))) {
        JTB_CUDA(cudaFuncSetAttribute(fp->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fp->smem));
        fp->attr_done |= 1u << (e.ctx->device & 31);
      }
      ColPairParams<T> cp;
      cp.z = wks; cp.out = ac + cs; cp.s = s; cp.bdist = bd; cp.zs = ws; cp.zbdist = wbd;
      cp.R1 = (int)R1; cp.cols = (int)cb; cp.batches = (int)batches; cp.kind = kind; cp.f0 = f0; cp.f = f; cp.dtw = dtw;
      JTB_TRY(fast_stage_table<T>(e, fp->logn, fp->loge, &cp.twg));
      const i64 nblk = (cb / fp->W) * (R1 / 2 + 1) * batches;
      if (nblk > 0x7fffffffLL) { set_error("too many column tiles"); return ST_UNSUPPORTED; }
      JTB_LAUNCH(fp->kern, (unsigned)nblk, (unsigned)fp->threads, (size_t)fp->smem, e.st, cp);
      JTB_CUDA(cudaGetLastError());
      e.ctx->launches++;
      continue;
    }
    Fast2Params<T> q = blank2<T>();
    q.in = wk + cs; q.out = wk2 + cs;
    q.nlines = cb * R1 * batches; q.c0 = (int)cb; q.gmod = (int)R1;
    q.in_gdist = R2 * s; q.in_gdist2 = bd; q.in_cdist = 1; q.in_stride = s;
    q.out_gdist = s; q.out_gdist2 = bd; q.out_cdist = 1; q.out_stride = R1 * s;
    JTB_TRY(launch2(e, f2, q));
    for (i64 b = 0; b < batches; ++b) {
      ColPostParams<T> cp;
      cp.z = wk2 + cs + b * bd; cp.out = ac + cs + b * bd;
      cp.n = n; cp.cols = cb; cp.s = s; cp.kind = kind; cp.f0 = f0; cp.f = f; cp.dtw = dtw;
      unsigned gr, bl;
      grid_for((n / 2 + 1) * cb, &gr, &bl);
      JTB_LAUNCH(k_r2r_colpost<T>, gr, bl, 0, e.st, cp);
      JTB_CUDA(cudaGetLastError());
      e.ctx->launches++;
    }
  }
  *handled = true;
  return ST_OK;
}

// ------------------------------------------------------------------------------------------ fast Bluestein
namespace {
template <typename T, int LOGN, int LOGE, int W, int MODE, int PRE> F2Entry<T> mk2x() {
  typedef Sched<LOGN, LOGE> S;
  F2Entry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.sin = 1; e.mode = MODE; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, true, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_fast2_kernel<T, LOGN, LOGE, true, MODE, W, PRE>;
  e.attr_done = 0; e.min_lines = 0;
  for (int d = 0; d < 16; ++d) e.twg[d] = nullptr;
  return e;
}
template <typename T> struct ConvEntry {
  int logn, loge, W, threads, smem;
  void (*kern)(const ConvParams<T>);
  unsigned attr_done;   // bit d set: smem attribute applied on device d
};
template <typename T, int LOGN, int LOGE, int W> ConvEntry<T> mkconv() {
  typedef Sched<LOGN, LOGE> S;
  ConvEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, false, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_conv_kernel<T, LOGN, LOGE, W>;
  e.attr_done = 0;
  return e;
}
template <typename T> struct BlueReg {
  std::vector<F2Entry<T>> first, last;   // strided passes (chirp in / chirp out)
  std::vector<ConvEntry<T>> mid;
};
template <typename T> BlueReg<T>& bluereg();
template <> BlueReg<double>& bluereg<double>() {
  static BlueReg<double> r = {
      {mk2x<double, 9, 3, 8, FM_TWID, PRE_CHIRP>(), mk2x<double, 10, 3, 4, FM_TWID, PRE_CHIRP>()},
      {mk2x<double, 9, 3, 8, FM_CHIRP_OUT, PRE_NONE>(), mk2x<double, 10, 3, 4, FM_CHIRP_OUT, PRE_NONE>()},
      {mkconv<double, 10, 3, 2>(), mkconv<double, 11, 3, 1>(), mkconv<double, 12, 3, 1>()}};
  return r;
}
template <> BlueReg<float>& bluereg<float>() {
  static BlueReg<float> r = {
      {mk2x<float, 9, 3, 8, FM_TWID, PRE_CHIRP>(), mk2x<float, 10, 4, 16, FM_TWID, PRE_CHIRP>(),
       mk2x<float, 10, 4, 8, FM_TWID, PRE_CHIRP>(), mk2x<float, 10, 3, 8, FM_TWID, PRE_CHIRP>()},
      {mk2x<float, 9, 3, 8, FM_CHIRP_OUT, PRE_NONE>(), mk2x<float, 10, 4, 16, FM_CHIRP_OUT, PRE_NONE>(),
       mk2x<float, 10, 4, 8, FM_CHIRP_OUT, PRE_NONE>(), mk2x<float, 10, 3, 8, FM_CHIRP_OUT, PRE_NONE>()},
      {mkconv<float, 10, 3, 4>(), mkconv<float, 11, 3, 2>(), mkconv<float, 12, 3, 1>(), mkconv<float, 11, 4, 2>(),
       mkconv<float, 12, 4, 1>()}};
  return r;
}

template <typename C> __global__ void k_permute_bk2(const C* bk2, C* out, int logN1, int logN2) {
  const i64 M = 1LL << (logN1 + logN2), N2 = 1LL << logN2;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (i64)gridDim.x * blockDim.x) {
    const i64 k1 = i >> logN2, k2 = i & (N2 - 1);
    out[i] = bk2[k1 + (k2 << logN1)];
  }
}
}  // namespace

// Bluestein chirp-z transform of `nlines` contiguous lines of non-power-of-two length n (line l at l*dist),
// in place, as three passes over an L2-sized work buffer (see fft_conv_kernel).
template <typename T>
int fast_bluestein_contig(Engine<T>& e, cx<T>* a, i64 dist, i64 nlines, i64 n, bool inverse, bool has_scale, T scale,
                          bool* handled) {
  typedef cx<T> C;
  *handled = false;
  if (g_fast2_off || nlines <= 0 || getenv("JTB_NO_FASTBLUE")) return ST_OK;
  const i64 M = next_pow2(2 * n - 1);
  const int logM = ilog2(M);
  BlueReg<T>& br = bluereg<T>();
  F2Entry<T>*fa = nullptr, *fc = nullptr;
  ConvEntry<T>* fb = nullptr;
  // tuning knobs: JTB_BLUE_SPLIT = index of the first-pass variant to start from, JTB_BLUE_LOGE = radix of the middle pass
  static const char* esp = getenv("JTB_BLUE_SPLIT");
  static const char* elg = getenv("JTB_BLUE_LOGE");
  // measured (4096 x n = 1 000 003, float): 512 x 4096 radix-8 136.4 ms, radix-16 middle pass 121.8 ms,
  // 1024 x 2048 split 131.5 ms, both 117.0 ms -> float default; double keeps the 512-point first pass
  const size_t i0 = esp ? (size_t)atoi(esp) % br.first.size() : (sizeof(T) == 4 ? 1 : 0);
  const int want_loge = elg ? atoi(elg) : (sizeof(T) == 4 ? 4 : 3);
  for (size_t ii = 0; ii < br.first.size() && !fa; ++ii) {
    const size_t i = (i0 + ii) % br.first.size();
    for (int pass = 0; pass < 2 && !fa; ++pass)
      for (auto& m : br.mid)
        if (br.first[i].logn + m.logn == logM && (pass == 1 || m.loge == want_loge)) { fa = &br.first[i]; fc = &br.last[i]; fb = &m; break; }
  }
  if (!fa) return ST_OK;
  const i64 N1 = 1LL << fa->logn, N2 = 1LL << fb->logn;
  if (N2 % fa->W) return ST_OK;
  const C *bk1, *bk2;
  i64 M2;
  JTB_TRY(e.blue_tables(n, &bk1, &bk2, &M2));
  const std::string kp = mkkey("bk2p", e.pname(), n, fa->logn);
  C* bk2p = (C*)e.ctx->table(kp);
  if (!bk2p) {
    JTB_CUDA(cudaMalloc((void**)&bk2p, (size_t)M * sizeof(C)));
    unsigned gr, bl;
    grid_for(M, &gr, &bl);
    JTB_LAUNCH(k_permute_bk2<C>, gr, bl, 0, e.st, bk2, bk2p, fa->logn, fb->logn);
    JTB_CUDA(cudaGetLastError());
    e.ctx->launches++;
    e.ctx->adopt_table(kp, bk2p, (size_t)M * sizeof(C));
  }
  const C *fsA, *fsB;
  int logL;
  JTB_TRY(e.fs_tables(logM, &fsA, &fsB, &logL));
  const char* ev = getenv("JTB_BLUE_MB");
  const double mb = ev ? atof(ev) : 1024.0;   // measured: larger chunks win (fewer, longer launches)
  i64 chunk = (i64)(mb * 1048576.0 / ((double)M * sizeof(C)));
  if (chunk < 1) chunk = 1;
  if (chunk > nlines) chunk = nlines;
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_BLUE], (size_t)chunk * (size_t)M * sizeof(C)));
  C* wk = (C*)e.ctx->work[WK_BLUE].p;
  if (!(fb->attr_done & (1u << (e.ctx->device & 31)))) {
    JTB_CUDA(cudaFuncSetAttribute(fb->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fb->smem));
    fb->attr_done |= 1u << (e.ctx->device & 31);
  }
  const cx<T>* twb;
  JTB_TRY(fast_stage_table<T>(e, fb->logn, fb->loge, &twb));
  for (i64 c0 = 0; c0 < nlines; c0 += chunk) {
    const i64 cn = (c0 + chunk < nlines ? c0 + chunk : nlines) - c0;
    Fast2Params<T> p = blank2<T>();
    p.in = a + c0 * dist; p.out = wk;
    p.nlines = cn * N2; p.c0 = (int)N2;
    p.in_gdist = dist; p.in_cdist = 1; p.in_stride = N2;
    p.out_gdist = M; p.out_cdist = 1; p.out_stride = N2;
    p.swap_in = inverse;
    p.fsA = fsA; p.fsB = fsB; p.fs_logL = logL; p.tw_src = 0;
    p.chirp = bk1; p.pre_n = n;
    JTB_TRY(launch2(e, fa, p));
    ConvParams<T> q;
    q.a = wk; q.nlines = cn * N1; q.N1 = (int)N1; q.h = bk2p; q.twg = twb; q.fsA = fsA; q.fsB = fsB; q.fs_logL = logL;
    JTB_LAUNCH(fb->kern, (unsigned)((q.nlines + fb->W - 1) / fb->W), (unsigned)fb->threads, (size_t)fb->smem, e.st, q);
    JTB_CUDA(cudaGetLastError());
    e.ctx->launches++;
    Fast2Params<T> r = blank2<T>();
    r.in = wk; r.out = a + c0 * dist;
    r.nlines = cn * N2; r.c0 = (int)N2;
    r.in_gdist = M; r.in_cdist = 1; r.in_stride = N2;
    r.out_gdist = dist; r.out_cdist = 1; r.out_stride = N2;
    r.swap_in = 1; r.swap_out1 = 1; r.swap_out = inverse;
    r.has_scale = has_scale; r.scale = scale;
    r.chirp = bk1; r.out_n = n;
    JTB_TRY(launch2(e, fc, r));
  }
  *handled = true;
  return ST_OK;
}

#define JTB_INST(T)                                                                                                    \
  template int fast_fourstep_contig<T>(Engine<T>&, const cx<T>*, i64, cx<T>*, i64, i64, i64, int, bool, bool, bool, T, \
                                       bool*);                                                                         \
  template int fast_fourstep_strided<T>(Engine<T>&, cx<T>*, const Geo&, i64, int, bool, bool, T, bool*);               \
  template int fast_rfft_fwd<T>(Engine<T>&, cx<T>*, i64, i64, int, bool*);                                             \
  template int fast_threepass_contig<T>(Engine<T>&, cx<T>*, i64, i64, i64, int, bool, bool, T, bool*);                 \
  template int fast_bluestein_contig<T>(Engine<T>&, cx<T>*, i64, i64, i64, bool, bool, T, bool*);                      \
  template int fast_r2r_rows<T>(Engine<T>&, T*, i64, i64, i64, int, T, T, bool*);                                      \
  template int fast_r2r_cols<T>(Engine<T>&, T*, i64, i64, i64, i64, int, T, T, bool*);
JTB_INST(double)
JTB_INST(float)

}  // namespace jtb
