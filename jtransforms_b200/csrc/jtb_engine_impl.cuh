// Engine<T> method bodies (included by tile_f64.cu / tile_f32.cu, which instantiate them).
//
// Reference roles replaced here (all paths relative to src/main/java/org/jtransforms):
//   tile_tables / fs_tables   <- utils/CommonUtils.java:372-435 (makewt), :502-519 (makect)
//   blue_tables               <- fft/DoubleFFT_1D.java:1864-1890 (bluesteini)
//   c2c_pow2                  <- utils/CommonUtils.java:708-793 (cftfsub / cftbsub drivers)
//   c2c_lines (Bluestein)     <- fft/DoubleFFT_1D.java:1920-2107 (bluestein_complex)
//   real_*_lines              <- fft/DoubleFFT_1D.java:524-561, :946-989 (+ rftfsub/rftbsub)
//   r2r_lines                 <- dct/DoubleDCT_1D.java:169-243,361-434, dst/DoubleDST_1D.java:96-160,
//                                264-325, dht/DoubleDHT_1D.java:94-152
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "jtb_engine.h"

namespace jtb {

template <typename T> const char* Engine<T>::pname() { return sizeof(T) == 8 ? "f64" : "f32"; }
template <typename T> int Engine<T>::max_logn_contig() {
  return g_limit_contig > 0 && g_limit_contig < TileLimits<T>::MAX_LOGN ? g_limit_contig : TileLimits<T>::MAX_LOGN;
}
template <typename T> int Engine<T>::max_logn_strided() {
  const int hw = TileLimits<T>::MAX_LOGN - 2;
  return g_limit_strided > 0 && g_limit_strided < hw ? g_limit_strided : hw;
}

static inline std::string mkkey(const char* what, const char* prec, i64 a, i64 b = 0) {
  char buf[96];
  snprintf(buf, sizeof buf, "%s:%s:%lld:%lld", what, prec, (long long)a, (long long)b);
  return std::string(buf);
}

// exp(-2 pi i num / den) evaluated in long double after exact integer reduction
template <typename T> static inline cx<T> unit_root(i64 num, i64 den) {
  num %= den;
  if (num < 0) num += den;
  const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)num / (long double)den;
  return mk<T>((T)cosl(ang), (T)sinl(ang));
}

template <typename T> int Engine<T>::init_tiles() {
  const int idx = sizeof(T) == 8 ? 0 : 1;
  if (!ctx->tile_init_done[idx]) {
    JTB_CUDA(tile_init_device<T>());
    ctx->tile_init_done[idx] = true;
  }
  return ST_OK;
}

// ---------------------------------------------------------------------------------- tables
template <typename T> int Engine<T>::tile_tables(int logn, const C** tw, const C** rtw) {
  const TileInfo ti = tile_info(logn);
  if (ti.logn != logn) { set_error("no tile kernel for 2^%d", logn); return ST_UNSUPPORTED; }
  i64 ns = 1;
  for (int s = 0; s < JTB_MAX_STAGES; ++s) {
    tw[s] = nullptr;
    if (s >= ti.nstages) continue;
    const i64 R = 1LL << ti.bits[s];
    if (s > 0) {
      const std::string key = mkkey("tw", pname(), logn, s);
      void* d = ctx->table(key);
      if (!d) {
        std::vector<C> h((size_t)((R - 1) * ns));
        for (i64 r = 1; r < R; ++r)
          for (i64 k = 0; k < ns; ++k) h[(size_t)((r - 1) * ns + k)] = unit_root<T>(r * k, ns * R);
        JTB_TRY(ctx->put_table(key, h.data(), h.size() * sizeof(C), &d));
      }
      tw[s] = (const C*)d;
    }
    ns *= R;
  }
  const std::string key = mkkey("rtw", pname(), logn);
  void* d = ctx->table(key);
  if (!d) {
    const i64 N = 1LL << logn;
    std::vector<C> h((size_t)(N / 2 + 1));
    for (i64 k = 0; k <= N / 2; ++k) h[(size_t)k] = unit_root<T>(k, 2 * N);
    JTB_TRY(ctx->put_table(key, h.data(), h.size() * sizeof(C), &d));
  }
  *rtw = (const C*)d;
  return ST_OK;
}

template <typename T> int Engine<T>::fs_tables(int logN, const C** A, const C** B, int* logL) {
  const int lL = logN / 2;
  // hot-path cache in front of the string-keyed table map
  static thread_local struct { const void* ctx; int logn; unsigned long gen; const C* a; const C* b; } memo = {nullptr, 0, 0, nullptr, nullptr};
  if (memo.ctx == (const void*)ctx && memo.logn == logN && memo.gen == ctx->table_gen && !ctx->recorder) {
    *A = memo.a; *B = memo.b; *logL = lL;
    return ST_OK;
  }
  const i64 N = 1LL << logN, L = 1LL << lL, H = N >> lL;
  const std::string ka = mkkey("fsA", pname(), logN), kb = mkkey("fsB", pname(), logN);
  void* da = ctx->table(ka);
  void* db = ctx->table(kb);
  if (!da || !db) {
    std::vector<C> ha((size_t)H), hb((size_t)L);
    for (i64 h = 0; h < H; ++h) ha[(size_t)h] = unit_root<T>(h * L, N);
    for (i64 l = 0; l < L; ++l) hb[(size_t)l] = unit_root<T>(l, N);
    JTB_TRY(ctx->put_table(ka, ha.data(), ha.size() * sizeof(C), &da));
    JTB_TRY(ctx->put_table(kb, hb.data(), hb.size() * sizeof(C), &db));
  }
  *A = (const C*)da; *B = (const C*)db; *logL = lL;
  memo.ctx = (const void*)ctx; memo.logn = logN; memo.gen = ctx->table_gen; memo.a = *A; memo.b = *B;
  return ST_OK;
}

template <typename T> int Engine<T>::dct_table(i64 n, const C** dtw) {
  const std::string key = mkkey("dct", pname(), n);
  void* d = ctx->table(key);
  if (!d) {
    std::vector<C> h((size_t)n);
    for (i64 k = 0; k < n; ++k) h[(size_t)k] = unit_root<T>(k, 4 * n);   // exp(-i pi k / (2n))
    JTB_TRY(ctx->put_table(key, h.data(), h.size() * sizeof(C), &d));
  }
  *dtw = (const C*)d;
  return ST_OK;
}

// chirp tables, built in double and (for float plans) rounded once at the end
int blue_tables_f64(Ctx* ctx, cudaStream_t st, i64 n, const double2** bk1, const double2** bk2, i64* M);

template <typename T> int Engine<T>::blue_tables(i64 n, const C** bk1, const C** bk2, i64* Mout) {
  const std::string k1 = mkkey("bk1", pname(), n), k2 = mkkey("bk2", pname(), n);
  const i64 M = next_pow2(2 * n - 1);
  *Mout = M;
  void* d1 = ctx->table(k1);
  void* d2 = ctx->table(k2);
  if (!d1 || !d2) {
    const double2 *b1, *b2;
    i64 M2;
    JTB_TRY(blue_tables_f64(ctx, st, n, &b1, &b2, &M2));
    if (sizeof(T) == 8) {
      d1 = (void*)b1; d2 = (void*)b2;
    } else {
      JTB_CUDA(cudaMalloc(&d1, (size_t)n * sizeof(C)));
      JTB_CUDA(cudaMalloc(&d2, (size_t)M * sizeof(C)));
      unsigned g, b;
      grid_for(n, &g, &b);
      JTB_LAUNCH(k_cast_c64_c32, g, b, 0, st, b1, (float2*)d1, n);
      grid_for(M, &g, &b);
      JTB_LAUNCH(k_cast_c64_c32, g, b, 0, st, b2, (float2*)d2, M);
      JTB_CUDA(cudaGetLastError());
      ctx->launches += 2;
      ctx->adopt_table(k1, d1, (size_t)n * sizeof(C));
      ctx->adopt_table(k2, d2, (size_t)M * sizeof(C));
    }
  }
  *bk1 = (const C*)d1; *bk2 = (const C*)d2;
  return ST_OK;
}

// ---------------------------------------------------------------------------------- tile launch
template <typename T>
int Engine<T>::tile_call(const C* in, const Geo& gi, C* out, const Geo& go, i64 l0, i64 l1, int logn,
                         TileParams<T>& p) {
  if (l1 <= l0) return ST_OK;
  JTB_TRY(init_tiles());
  const TileInfo ti = tile_info(logn);
  const C* tw[JTB_MAX_STAGES];
  const C* rtw;
  JTB_TRY(tile_tables(logn, tw, &rtw));
  for (int s = 0; s < JTB_MAX_STAGES; ++s) p.tw[s] = tw[s];
  p.rtw = rtw;
  p.in = in; p.out = out; p.gi = gi; p.go = go;
  p.nlines = l1; p.line_base = l0;
  const bool wf = gi.stride != 1;
  const bool staged = p.pro != PRO_DIRECT || p.epi != EPI_DIRECT;
  const size_t line_bytes = (size_t)ti.ld * sizeof(C);
  int W;
  if (wf || p.epi == EPI_REMAP) W = (int)(128 / sizeof(C));      // 128-byte segments across adjacent lines
  else W = 256 / ti.tpl;
  {
    const char* ev = getenv(wf ? "JTB_W_STRIDED" : (p.epi == EPI_REMAP ? "JTB_W_REMAP" : "JTB_W_CONTIG"));
    if (ev && atoi(ev) > 0) W = atoi(ev);
  }
  if (W > ti.maxt / ti.tpl) W = ti.maxt / ti.tpl;
  const bool need_smem = ti.nstages > 1 || staged;
  if (need_smem) while (W > 1 && (size_t)W * line_bytes > (size_t)227 * 1024) W >>= 1;
  if (W < 1) W = 1;
  if ((i64)W > l1 - l0) { W = 1; while ((i64)W * 2 <= l1 - l0 && W * 2 * ti.tpl <= ti.maxt) W *= 2; }
  const size_t smem = need_smem ? (size_t)W * line_bytes : 0;
  if (smem > (size_t)227 * 1024) { set_error("tile 2^%d does not fit shared memory", logn); return ST_UNSUPPORTED; }
  p.W = W; p.wfast = wf ? 1 : 0;
  const i64 nblk = (l1 - l0 + W - 1) / W;
  if (nblk > 0x7fffffffLL) { set_error("too many lines for one launch"); return ST_UNSUPPORTED; }
  JTB_CUDA(launch_tile<T>(logn, p, (unsigned)nblk, (unsigned)(W * ti.tpl), smem, st));
  ctx->launches++;
  return ST_OK;
}

template <typename T> static inline void apply_fuse_in(TileParams<T>& p, const Fuse<T>& f) {
  p.swap_in = f.swap_in; p.swap_in2 = f.swap_in2;
  p.premul = f.premul; p.premul_conj = f.premul_conj;
  p.valid_in = f.valid_in;
}
template <typename T> static inline void apply_fuse_out(TileParams<T>& p, const Fuse<T>& f) {
  p.swap_out1 = f.swap_out1; p.swap_out = f.swap_out;
  p.postmul = f.postmul; p.postmul_conj = f.postmul_conj;
  p.valid_out = f.valid_out;
  p.has_scale = f.has_scale; p.scale = f.scale;
}
template <typename T> static inline TileParams<T> blank_params() {
  TileParams<T> p;
  memset(&p, 0, sizeof p);
  p.valid_in = p.valid_out = -1;
  p.lin_ks = p.lout_ks = 1;
  p.scale = 1;
  return p;
}

// extent (in elements) touched by lines [0, l1) of length n
static inline i64 geo_extent(const Geo& g, i64 l1, i64 n) {
  i64 ext = (n - 1) * g.stride + 1, lines = l1, prod = 1;
  for (int k = 0; k < 3; ++k) { ext += (g.c[k] - 1) * g.d[k]; prod *= g.c[k]; }
  const i64 top = (lines + prod - 1) / prod;
  ext += (top - 1) * g.d[3];
  return ext;
}
static inline bool geo_same(const Geo& a, const Geo& b) {
  for (int k = 0; k < 3; ++k) if (a.c[k] != b.c[k] || a.d[k] != b.d[k]) return false;
  return a.d[3] == b.d[3] && a.stride == b.stride;
}

// ---------------------------------------------------------------------------------- pow2 c2c
template <typename T>
int Engine<T>::c2c_pow2(const C* in, const Geo& gi, C* out, const Geo& go, i64 l0, i64 l1, int logn,
                        const Fuse<T>& f, int pro, int epi) {
  if (l1 <= l0) return ST_OK;
  const bool contig = gi.stride == 1 && go.stride == 1;
  const int lim = contig ? max_logn_contig() : max_logn_strided();
  if (!contig && logn >= 11 && logn <= lim && pro == PRO_DIRECT && epi == EPI_DIRECT && in == out && geo_same(gi, go) &&
      l0 == 0 && f.swap_in == f.swap_out && !f.premul && !f.postmul && f.valid_in < 0 && f.valid_out < 0 && !f.swap_in2 &&
      !f.swap_out1) {
    // long strided lines: two lean passes at HBM speed beat one pass of the general tile kernel (float 4096-point
    // columns: 148 us -> two passes of ~25 us)
    bool handled = false;
    JTB_TRY(fast_fourstep_strided<T>(*this, out, go, l1, logn, f.swap_in, f.has_scale, f.scale, &handled));
    if (handled) return ST_OK;
  }
  if (logn <= lim) {
    TileParams<T> p = blank_params<T>();
    apply_fuse_in(p, f);
    apply_fuse_out(p, f);
    p.pro = pro; p.epi = epi;
    return tile_call(in, gi, out, go, l0, l1, logn, p);
  }
  if (pro != PRO_DIRECT || epi != EPI_DIRECT) { set_error("fused real pass needs a single-tile length"); return ST_UNSUPPORTED; }
  {
    // Beyond the reach of the lean two-pass kernels (2^21) the three-pass composition of lean kernels beats the
    // general two-pass tile path and is the only path above 2^26 (JTB_THREEPASS_MIN: test knob).
    static const char* emin = getenv("JTB_THREEPASS_MIN");
    const int tp_min = emin ? atoi(emin) : 22;
    const bool one_level = gi.c[0] == 1 && gi.c[1] == 1 && gi.c[2] == 1;
    if ((logn > 2 * max_logn_contig() || logn >= tp_min) && contig && in == out && geo_same(gi, go) && one_level &&
        !f.premul && !f.postmul && f.valid_in < 0 && f.valid_out < 0 && !f.swap_in2 && !f.swap_out1 && f.swap_in == f.swap_out)
      return c2c_big_contig(out, go.d[3], l0, l1, logn, f.swap_in != 0, f.has_scale != 0, f.scale);
  }
  {
    // lean two-pass kernels when nothing but the inverse swaps and a scale is fused
    const bool plain = !f.premul && !f.postmul && f.valid_in < 0 && f.valid_out < 0 && !f.swap_in2 && !f.swap_out1;
    const bool one_level_i = gi.c[0] == 1 && gi.c[1] == 1 && gi.c[2] == 1;
    const bool one_level_o = go.c[0] == 1 && go.c[1] == 1 && go.c[2] == 1;
    bool handled = false;
    if (plain && contig && one_level_i && one_level_o) {
      JTB_TRY(fast_fourstep_contig<T>(*this, in, gi.d[3], out, go.d[3], l0, l1, logn, f.swap_in, f.swap_out, f.has_scale,
                                      f.scale, &handled));
    } else if (plain && !contig && in == out && geo_same(gi, go) && l0 == 0 && f.swap_in == f.swap_out) {
      JTB_TRY(fast_fourstep_strided<T>(*this, out, go, l1, logn, f.swap_in, f.has_scale, f.scale, &handled));
    }
    if (handled) return ST_OK;
  }
  if (logn > 2 * max_logn_contig()) {
    set_error("length 2^%d exceeds the two-pass limit 2^%d", logn, 2 * max_logn_contig());
    return ST_UNSUPPORTED;
  }
  if (gi.c[2] != 1 || go.c[2] != 1) { set_error("geometry too deep for the two-pass transform"); return ST_UNSUPPORTED; }

  // n = N1 * N2, input index j = n1*N2 + n2, output index k = k1 + N1*k2
  const int l2 = logn / 2, l1g = logn - l2;
  const i64 n = 1LL << logn, N1 = 1LL << l1g, N2 = 1LL << l2;
  const C *fsA, *fsB;
  int logL;
  JTB_TRY(fs_tables(logn, &fsA, &fsB, &logL));
  const int at_i = gi.stride == 1 ? 0 : 1;
  const int at_o = go.stride == 1 ? 0 : 1;

  const bool mirrored = !contig && in == out && geo_same(gi, go) && f.valid_in < 0 && f.valid_out < 0;
  if (mirrored) {
    // strided in-place lines (column / slice axes of 2-D / 3-D arrays): the intermediate keeps the
    // array's own layout so both passes stay coalesced across adjacent lines.
    if (l0 != 0) { set_error("internal: mirrored two-pass call must start at line 0"); return ST_ARG; }
    const i64 ext = geo_extent(gi, l1, n);
    JTB_TRY(ctx->ensure(ctx->work[WK_FOURSTEP], (size_t)ext * sizeof(C)));
    C* wk = (C*)ctx->work[WK_FOURSTEP].p;
    Geo g1 = geo_insert(gi, at_i, N2, gi.stride);
    g1.stride = N2 * gi.stride;
    TileParams<T> p = blank_params<T>();
    apply_fuse_in(p, f);
    p.lin_ks = N2; p.lin_is = 1; p.lin_level = at_i;
    p.fs_mode = 1; p.fs_level = at_i; p.fsA = fsA; p.fsB = fsB; p.fs_logL = logL;
    JTB_TRY(tile_call(in, g1, wk, g1, 0, l1 * N2, l1g, p));
    Geo g2i = geo_insert(gi, at_i, N1, N2 * gi.stride);   // line k1 starts at row k1*N2, runs over n2
    g2i.stride = gi.stride;
    Geo g2o = geo_insert(go, at_i, N1, go.stride);         // output element k1 + N1*k2
    g2o.stride = N1 * go.stride;
    TileParams<T> q = blank_params<T>();
    apply_fuse_out(q, f);
    q.lout_ks = N1; q.lout_is = 1; q.lout_level = at_i;
    JTB_TRY(tile_call(wk, g2i, out, g2o, 0, l1 * N1, l2, q));
    return ST_OK;
  }

  // dense intermediate work[L][k1][n2], L = line - c0
  i64 gran = 1;
  if (at_i == 1) gran = gi.c[0];
  if (at_o == 1 && go.c[0] > gran) gran = go.c[0];
  if (at_i == 1 && at_o == 1 && gi.c[0] != go.c[0]) { set_error("mismatched strided geometries"); return ST_UNSUPPORTED; }
  i64 chunk = (i64)(ctx->work_cap / ((size_t)n * sizeof(C)));
  chunk -= chunk % gran;
  if (chunk < gran) { set_error("two-pass workspace exceeds the cap"); return ST_OOM; }
  if (gran > 1 && (l0 % gran) != 0) { set_error("internal: unaligned chunk"); return ST_ARG; }
  if (chunk > l1 - l0) chunk = l1 - l0;
  JTB_TRY(ctx->ensure(ctx->work[WK_FOURSTEP], (size_t)chunk * (size_t)n * sizeof(C)));
  for (i64 c0 = l0; c0 < l1; c0 += chunk) {
    const i64 c1 = c0 + chunk < l1 ? c0 + chunk : l1;
    C* wk = (C*)ctx->work[WK_FOURSTEP].p - c0 * n;
    {
      Geo g1 = geo_insert(gi, at_i, N2, gi.stride);
      g1.stride = N2 * gi.stride;
      Geo w1 = geo_work_like(g1, at_i, 1, n);
      w1.stride = N2;
      TileParams<T> p = blank_params<T>();
      apply_fuse_in(p, f);
      p.lin_ks = N2; p.lin_is = 1; p.lin_level = at_i;
      p.fs_mode = 1; p.fs_level = at_i; p.fsA = fsA; p.fsB = fsB; p.fs_logL = logL;
      JTB_TRY(tile_call(in, g1, wk, w1, c0 * N2, c1 * N2, l1g, p));
    }
    {
      Geo g2 = geo_insert(go, at_o, N1, go.stride);
      g2.stride = N1 * go.stride;
      Geo w2 = geo_work_like(g2, at_o, N2, n);
      w2.stride = 1;
      TileParams<T> q = blank_params<T>();
      apply_fuse_out(q, f);
      q.lout_ks = N1; q.lout_is = 1; q.lout_level = at_o;
      q.epi = at_o == 0 ? EPI_REMAP : EPI_DIRECT;
      JTB_TRY(tile_call(wk, w2, out, g2, c0 * N1, c1 * N1, l2, q));
    }
  }
  return ST_OK;
}

template <typename T>
int fast_c2c(Engine<T>& e, cx<T>* a, const Geo& g, i64 nlines, int logn, bool inverse, bool has_scale, T scale,
             bool* handled);   // jtb_fast.cu

// Lines beyond the two-pass limit (2^26 double / 2^28 float; the reference reaches them through LargeArray,
// fft/DoubleFFT_1D.java:280-304): n = N1*N2 with both factors within the two-pass range.  View the line as [N1][N2]:
// (1) N2 adjacent strided transforms of length N1 in place, (2) twiddle W_n^(k1 n2), (3) N1 contiguous transforms of
// length N2, (4) transpose to natural order k1 + N1*k2 through the workspace.  Every step reuses the lean kernels.
template <typename T>
int Engine<T>::c2c_big_contig(C* a, i64 dist, i64 l0, i64 l1, int logn, bool inverse, bool has_scale, T scale) {
  {
    // lean version: three sweeps (strided two-pass sub-transform with the outer twiddle fused, transposing row pass)
    bool handled = false;
    JTB_TRY(fast_threepass_contig<T>(*this, a, dist, l0, l1, logn, inverse, has_scale, scale, &handled));
    if (handled) return ST_OK;
  }
  const int l2 = logn / 2, l1g = logn - l2;
  if (l1g > 2 * max_logn_strided() || l2 > 2 * max_logn_contig() || l2 < 5 || l1g < 5) {
    set_error("length 2^%d exceeds the three-pass limit", logn);
    return ST_UNSUPPORTED;
  }
  const i64 n = 1LL << logn, N1 = 1LL << l1g, N2 = 1LL << l2;
  const C *fsA, *fsB;
  int logL;
  JTB_TRY(fs_tables(logn, &fsA, &fsB, &logL));
  JTB_TRY(ctx->ensure(ctx->work[WK_BIG], (size_t)n * sizeof(C)));
  C* wk = (C*)ctx->work[WK_BIG].p;
  unsigned g, b;
  for (i64 l = l0; l < l1; ++l) {
    C* base = a + l * dist;
    JTB_TRY(c2c_lines(base, geo_make(N2, 1, n, N2), N2, N1, inverse, false, (T)1));
    grid_for(n, &g, &b);
    JTB_LAUNCH(k_big_twiddle<C>, g, b, 0, st, base, N1, l2, fsA, fsB, logL, inverse ? 1 : 0);
    JTB_CUDA(cudaGetLastError());
    ctx->launches++;
    JTB_TRY(c2c_lines(base, geo_contig(N2), N1, N2, inverse, has_scale, scale));
    i64 nt = (N1 / 32) * (N2 / 32);
    if (nt > 148 * 16) nt = 148 * 16;
    JTB_LAUNCH(k_transpose32<C>, (unsigned)nt, dim3(32, 8), (size_t)(32 * 33 * sizeof(C)), st, base, wk, N1, N2);
    JTB_CUDA(cudaGetLastError());
    ctx->launches++;
    JTB_CUDA(cudaMemcpyAsync(base, wk, (size_t)n * sizeof(C), cudaMemcpyDeviceToDevice, st));
  }
  return ST_OK;
}

// ---------------------------------------------------------------------------------- any-length c2c
template <typename T>
int Engine<T>::c2c_lines(C* a, const Geo& g, i64 nlines, i64 n, bool inverse, bool has_scale, T scale) {
  if (nlines <= 0 || n < 1) return ST_OK;
  if (n == 1) {
    if (has_scale && scale != (T)1) { set_error("internal: scaled length-1 transform"); return ST_ARG; }
    return ST_OK;
  }
  if (is_pow2(n)) {
    bool handled = false;
    JTB_TRY(fast_c2c<T>(*this, a, g, nlines, ilog2(n), inverse, has_scale, scale, &handled));
    if (handled) return ST_OK;
    Fuse<T> f;
    f.swap_in = inverse; f.swap_out = inverse;
    f.has_scale = has_scale; f.scale = scale;
    return c2c_pow2(a, g, a, g, 0, nlines, ilog2(n), f);
  }
  // smooth lengths that fit one CTA: native mixed-radix passes (the reference's FFTPACK plan)
  {
    bool handled = false;
    JTB_TRY(mixed_c2c<T>(*this, a, g, nlines, n, inverse, has_scale, scale, &handled));
    if (handled) return ST_OK;
  }
  // Bluestein chirp-z (fft/DoubleFFT_1D.java:1920-2107); inverse = swap . forward . swap
  if (g.stride == 1 && g.c[0] == 1 && g.c[1] == 1 && g.c[2] == 1) {
    bool handled = false;
    // long smooth lengths: two mixed-radix passes (n = N1*N2)
    JTB_TRY(mixed_twopass_contig<T>(*this, a, g.d[3], nlines, n, inverse, has_scale, scale, &handled));
    if (handled) return ST_OK;
    JTB_TRY(fast_bluestein_contig<T>(*this, a, g.d[3], nlines, n, inverse, has_scale, scale, &handled));
    if (handled) return ST_OK;
  }
  const C *bk1, *bk2;
  i64 M;
  JTB_TRY(blue_tables(n, &bk1, &bk2, &M));
  const int logM = ilog2(M);
  const i64 gran = g.stride == 1 ? 1 : g.c[0];
  i64 chunk = (i64)(ctx->work_cap / ((size_t)M * sizeof(C)));
  chunk -= chunk % gran;
  if (chunk < gran) { set_error("Bluestein workspace exceeds the cap"); return ST_OOM; }
  if (chunk > nlines) chunk = nlines;
  JTB_TRY(ctx->ensure(ctx->work[WK_BLUE], (size_t)chunk * (size_t)M * sizeof(C)));
  const Geo gw = geo_contig(M);
  for (i64 c0 = 0; c0 < nlines; c0 += chunk) {
    const i64 c1 = c0 + chunk < nlines ? c0 + chunk : nlines;
    C* wk = (C*)ctx->work[WK_BLUE].p - c0 * M;
    Fuse<T> f1;
    f1.swap_in = inverse;
    f1.premul = bk1; f1.premul_conj = 1; f1.valid_in = n;
    f1.postmul = bk2; f1.swap_out = 1;
    JTB_TRY(c2c_pow2(a, g, wk, gw, c0, c1, logM, f1));
    Fuse<T> f2;
    f2.swap_out1 = 1;                       // undo the swapped-domain inverse
    f2.postmul = bk1; f2.postmul_conj = 1; f2.valid_out = n;
    f2.has_scale = has_scale; f2.scale = scale;
    f2.swap_out = inverse;
    JTB_TRY(c2c_pow2(wk, gw, a, g, c0, c1, logM, f2));
  }
  return ST_OK;
}

// ---------------------------------------------------------------------------------- real lines
template <typename T> static inline bool geo_even(const Geo& g) {
  return g.stride == 1 && (g.d[0] % 2 == 0) && (g.d[1] % 2 == 0) && (g.d[2] % 2 == 0) && (g.d[3] % 2 == 0);
}
static inline Geo geo_halve(const Geo& g) {
  Geo r = g;
  for (int k = 0; k < 4; ++k) r.d[k] = g.d[k] / 2;
  return r;
}

template <typename T> static int r2r_stage(Engine<T>& e, bool pre, R2RParams<T>& p) {
  unsigned gr, bl;
  grid_for((p.nlines - p.line_base) * p.n, &gr, &bl);
  if (pre) JTB_LAUNCH(k_r2r_pre<T>, gr, bl, 0, e.st, p);
  else JTB_LAUNCH(k_r2r_post<T>, gr, bl, 0, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  return ST_OK;
}

// generic pre -> complex FFT -> post pipeline over chunks of lines
template <typename T>
static int staged_lines(Engine<T>& e, T* a, const Geo& g, i64 nlines, i64 n, int pre_mode, int post_mode, int dstflag,
                        bool inverse, T pre_f0, T pre_f, T post_f0, T post_f, const cx<T>* dtw) {
  typedef cx<T> C;
  i64 chunk = (i64)(e.ctx->work_cap / ((size_t)n * sizeof(C)));
  if (chunk < 1) { set_error("line of %lld points exceeds the workspace cap", (long long)n); return ST_OOM; }
  if (chunk > nlines) chunk = nlines;
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_REAL], (size_t)chunk * (size_t)n * sizeof(C)));
  C* wk = (C*)e.ctx->work[WK_REAL].p;
  for (i64 c0 = 0; c0 < nlines; c0 += chunk) {
    const i64 c1 = c0 + chunk < nlines ? c0 + chunk : nlines;
    R2RParams<T> p;
    p.a = a; p.work = wk; p.g = g; p.line_base = c0; p.nlines = c1; p.n = n;
    p.mode = pre_mode; p.dst = dstflag; p.f0 = pre_f0; p.f = pre_f; p.dtw = dtw;
    JTB_TRY(r2r_stage(e, true, p));
    JTB_TRY(e.c2c_lines(wk, geo_contig(n), c1 - c0, n, inverse, false, (T)1));
    p.mode = post_mode; p.f0 = post_f0; p.f = post_f;
    JTB_TRY(r2r_stage(e, false, p));
  }
  return ST_OK;
}

template <typename T> int Engine<T>::real_forward_lines(T* a, const Geo& g, i64 nlines, i64 n) {
  if (n <= 1 || nlines <= 0) return ST_OK;
  if (is_pow2(n) && n >= 4 && ilog2(n) - 1 <= max_logn_contig() && geo_even<T>(g) && ((uintptr_t)a % sizeof(C)) == 0) {
    const Geo gc = geo_halve(g);
    if (gc.c[0] == 1 && gc.c[1] == 1 && gc.c[2] == 1) {
      bool handled = false;
      JTB_TRY(fast_rfft_fwd<T>(*this, (C*)a, gc.d[3], nlines, ilog2(n) - 1, &handled));
      if (handled) return ST_OK;
    }
    Fuse<T> f;
    return c2c_pow2((const C*)a, gc, (C*)a, gc, 0, nlines, ilog2(n) - 1, f, PRO_DIRECT, EPI_RFFT_FWD);
  }
  return staged_lines<T>(*this, a, g, nlines, n, PRE_R2C, POST_PACK, 0, false, (T)1, (T)1, (T)1, (T)1, nullptr);
}

template <typename T> int Engine<T>::real_inverse_lines(T* a, const Geo& g, i64 nlines, i64 n, bool scale) {
  if (n <= 1 || nlines <= 0) return ST_OK;
  if (is_pow2(n) && n >= 4 && ilog2(n) - 1 <= max_logn_contig() && geo_even<T>(g) && ((uintptr_t)a % sizeof(C)) == 0) {
    // unscaled power-of-two result is (n/2) x (fft/DoubleFFT_1D.java:946-967)
    const Geo gc = geo_halve(g);
    if (gc.c[0] == 1 && gc.c[1] == 1 && gc.c[2] == 1) {
      bool handled = false;
      JTB_TRY(fast_rfft_inv<T>(*this, (C*)a, gc.d[3], nlines, ilog2(n) - 1, scale, (T)(1.0 / (double)(n / 2)), &handled));
      if (handled) return ST_OK;
    }
    Fuse<T> f;
    f.swap_out = 1;
    f.has_scale = scale; f.scale = (T)(1.0 / (double)(n / 2));
    return c2c_pow2((const C*)a, gc, (C*)a, gc, 0, nlines, ilog2(n) - 1, f, PRO_RFFT_INV, EPI_DIRECT);
  }
  // generic: Hermitian expansion, unnormalised inverse, real part.  Unscaled: n x for non-power-of-two
  // lengths (:977,983) but (n/2) x for powers of two.
  const T fac = scale ? (T)(1.0 / (double)n) : (is_pow2(n) ? (T)0.5 : (T)1);
  return staged_lines<T>(*this, a, g, nlines, n, PRE_UNPACK_HERM, POST_REAL, 0, true, (T)1, (T)1, fac, fac, nullptr);
}

template <typename T> int Engine<T>::r2r_lines(T* a, const Geo& g, i64 nlines, i64 n, int kind, bool inverse, bool scale) {
  if (n <= 1 || nlines <= 0) return ST_OK;
  const double dn = (double)n;
  // fused forward kernels (contiguous lines, or the column axis of row-major arrays)
  {
    T f0 = 1, f = 1;
    bool fwd_like = !inverse;
    if (kind == 3) { fwd_like = true; f0 = f = (inverse && scale) ? (T)(1.0 / dn) : (T)1; }
    else if (scale) { f0 = (T)std::sqrt(1.0 / dn); f = (T)std::sqrt(2.0 / dn); }
    if (fwd_like && is_pow2(n)) {
      bool handled = false;
      if (g.stride == 1 && g.c[0] == 1 && g.c[1] == 1 && g.c[2] == 1) {
        JTB_TRY(fast_r2r_rows<T>(*this, a, g.d[3], nlines, n, kind, f0, f, &handled));
      } else if (g.stride > 1 && g.d[0] == 1 && g.c[1] == 1 && g.c[2] == 1 && g.c[0] == g.stride &&
                 nlines % g.c[0] == 0 && (nlines == g.c[0] || g.d[3] >= n * g.stride)) {
        JTB_TRY(fast_r2r_cols<T>(*this, a, n, g.c[0], nlines / g.c[0], g.d[3], kind, f0, f, &handled));
      }
      if (handled) return ST_OK;
    }
  }
  if (kind == 3) {   // DHT: forward == inverse up to 1/n (dht/DoubleDHT_1D.java:255-270)
    const T fac = (inverse && scale) ? (T)(1.0 / dn) : (T)1;
    return staged_lines<T>(*this, a, g, nlines, n, PRE_R2C, POST_DHT, 0, false, (T)1, (T)1, fac, fac, nullptr);
  }
  const C* dtw;
  JTB_TRY(dct_table(n, &dtw));
  const int dst = kind == 2 ? 1 : 0;
  const bool p2 = is_pow2(n);
  if (!inverse) {
    // DCT-II.  Unscaled: sum (pow2) / 2 sum (otherwise); scaled: orthonormal (dct/DoubleDCT_1D.java:169-243)
    T f0, f;
    if (scale) { f0 = (T)std::sqrt(1.0 / dn); f = (T)std::sqrt(2.0 / dn); }
    else { f0 = f = p2 ? (T)1 : (T)2; }
    return staged_lines<T>(*this, a, g, nlines, n, PRE_DCT2, POST_DCT2, dst, false, (T)1, (T)1, f0, f, dtw);
  }
  // DCT-III (dct/DoubleDCT_1D.java:361-434)
  T f0, f;
  if (scale) { f0 = (T)std::sqrt(1.0 / dn); f = (T)std::sqrt(2.0 / dn); }
  else if (p2) { f0 = f = (T)1; }
  else { f0 = (T)(0.5 / dn); f = (T)(1.0 / dn); }
  if (p2) {   // fused inverse kernels (jtb_r2r_inv.cuh)
    bool handled = false;
    if (g.stride == 1 && g.c[0] == 1 && g.c[1] == 1 && g.c[2] == 1) {
      JTB_TRY(fast_r2r_rows_inv<T>(*this, a, g.d[3], nlines, n, kind, f0, f, &handled));
    } else if (g.stride > 1 && g.d[0] == 1 && g.c[1] == 1 && g.c[2] == 1 && g.c[0] == g.stride &&
               nlines % g.c[0] == 0 && (nlines == g.c[0] || g.d[3] >= n * g.stride)) {
      JTB_TRY(fast_r2r_cols_inv<T>(*this, a, n, g.c[0], nlines / g.c[0], g.d[3], kind, f0, f, &handled));
    }
    if (handled) return ST_OK;
  }
  return staged_lines<T>(*this, a, g, nlines, n, PRE_DCT3, POST_DCT3, dst, true, f0, f, (T)1, (T)1, dtw);
}

}  // namespace jtb
