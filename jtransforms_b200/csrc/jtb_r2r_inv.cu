// Instantiations and host-side dispatch of the fused inverse DCT/DST kernels (jtb_r2r_inv.cuh).
#include <cstdlib>

#include "jtb_engine_impl.cuh"
#include "jtb_r2r_inv.cuh"

namespace jtb {

namespace {

const bool g_inv_off = getenv("JTB_NO_FASTINV") != nullptr;

template <typename T> struct RowInvEntry {
  int logn, loge, kind, W, threads, smem;
  void (*kern)(const RowR2RParams<T>);
  unsigned attr_done;   // bit d set: smem attribute applied on device d
};
template <typename T, int LOGN, int LOGE, int KIND, int W> RowInvEntry<T> mkrowinv() {
  typedef Sched<LOGN, LOGE> S;
  RowInvEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.kind = KIND; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, false, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_r2r_row_inv_kernel<T, LOGN, LOGE, KIND, W>;
  e.attr_done = 0;
  return e;
}
#define JTB_ROWS_INV(T, LOGN, LOGE, W) mkrowinv<T, LOGN, LOGE, RK_DCT, W>(), mkrowinv<T, LOGN, LOGE, RK_DST, W>()
template <typename T> std::vector<RowInvEntry<T>>& rowinvreg();
template <> std::vector<RowInvEntry<double>>& rowinvreg<double>() {
  static std::vector<RowInvEntry<double>> r = {JTB_ROWS_INV(double, 12, 4, 1), JTB_ROWS_INV(double, 12, 3, 1), JTB_ROWS_INV(double, 11, 3, 1), JTB_ROWS_INV(double, 11, 4, 2),
                                               JTB_ROWS_INV(double, 10, 3, 2), JTB_ROWS_INV(double, 9, 3, 4),
                                               JTB_ROWS_INV(double, 8, 4, 8), JTB_ROWS_INV(double, 7, 4, 16),
                                               JTB_ROWS_INV(double, 6, 3, 16), JTB_ROWS_INV(double, 5, 3, 32)};
  return r;
}
template <> std::vector<RowInvEntry<float>>& rowinvreg<float>() {
  static std::vector<RowInvEntry<float>> r = {JTB_ROWS_INV(float, 12, 4, 1), JTB_ROWS_INV(float, 11, 4, 2),
                                              JTB_ROWS_INV(float, 10, 4, 4), JTB_ROWS_INV(float, 9, 3, 8),
                                              JTB_ROWS_INV(float, 8, 4, 16), JTB_ROWS_INV(float, 7, 4, 16),
                                              JTB_ROWS_INV(float, 6, 3, 32), JTB_ROWS_INV(float, 5, 3, 32)};
  return r;
}

// first pass of the strided inverse (pair pre-pass + FFT over q2 + twiddle)
template <typename T> struct PairInvEntry {
  int logn, loge, W, threads, smem;
  void (*kern)(const ColPairParams<T>, const cx<T>*, const cx<T>*, int);
  unsigned attr_done;
};
template <typename T, int LOGN, int LOGE, int W> PairInvEntry<T> mkpairinv() {
  typedef Sched<LOGN, LOGE> S;
  PairInvEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.W = W; e.threads = 2 * W * S::TPL;
  e.smem = (int)((FastAddr<T, S, true, 2 * W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_colpair_inv_kernel<T, LOGN, LOGE, W>;
  e.attr_done = 0;
  return e;
}
template <typename T> std::vector<PairInvEntry<T>>& pairinvreg();
template <> std::vector<PairInvEntry<double>>& pairinvreg<double>() {
  static std::vector<PairInvEntry<double>> r = {mkpairinv<double, 6, 3, 32>(), mkpairinv<double, 7, 4, 16>(),
                                                mkpairinv<double, 5, 3, 32>()};
  return r;
}
template <> std::vector<PairInvEntry<float>>& pairinvreg<float>() {
  static std::vector<PairInvEntry<float>> r = {mkpairinv<float, 6, 3, 32>(), mkpairinv<float, 7, 4, 16>(),
                                               mkpairinv<float, 5, 3, 32>()};
  return r;
}

// second pass (FFT over q1, un-permuting store)
template <typename T> struct UnpermEntry {
  int logn, loge, pre, W, threads, smem;
  void (*kern)(const Fast2Params<T>);
  unsigned attr_done;
};
template <typename T, int LOGN, int LOGE, int W, int PRE> UnpermEntry<T> mkunperm() {
  typedef Sched<LOGN, LOGE> S;
  UnpermEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.pre = PRE; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, true, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_fast2_kernel<T, LOGN, LOGE, true, FM_PLAIN, W, PRE>;
  e.attr_done = 0;
  return e;
}
template <typename T> std::vector<UnpermEntry<T>>& unpermreg();
template <> std::vector<UnpermEntry<double>>& unpermreg<double>() {
  static std::vector<UnpermEntry<double>> r = {
      mkunperm<double, 6, 3, 32, PRE_UNPERM_DCT>(), mkunperm<double, 6, 3, 32, PRE_UNPERM_DST>(),
      mkunperm<double, 7, 4, 16, PRE_UNPERM_DCT>(), mkunperm<double, 7, 4, 16, PRE_UNPERM_DST>()};
  return r;
}
template <> std::vector<UnpermEntry<float>>& unpermreg<float>() {
  static std::vector<UnpermEntry<float>> r = {
      mkunperm<float, 6, 3, 32, PRE_UNPERM_DCT>(), mkunperm<float, 6, 3, 32, PRE_UNPERM_DST>(),
      mkunperm<float, 7, 4, 16, PRE_UNPERM_DCT>(), mkunperm<float, 7, 4, 16, PRE_UNPERM_DST>()};
  return r;
}

// single-pass strided kernels (forward and inverse)
template <typename T> struct ColEntry {
  int logn, loge, W, inv, threads, smem;
  void (*kern)(const ColR2RParams<T>);
  unsigned attr_done;
};
template <typename T, int LOGN, int LOGE, int W, bool INV> ColEntry<T> mkcol() {
  typedef Sched<LOGN, LOGE> S;
  ColEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.W = W; e.inv = INV; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, true, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_col_r2r_kernel<T, LOGN, LOGE, W, INV>;
  e.attr_done = 0;
  return e;
}
#define JTB_COLS(T, LOGN, LOGE, W) mkcol<T, LOGN, LOGE, W, false>(), mkcol<T, LOGN, LOGE, W, true>()
template <typename T> std::vector<ColEntry<T>>& colreg();
template <> std::vector<ColEntry<double>>& colreg<double>() {
  // per length: widest tile first (needs cols % W == 0), 128-byte rows as the fallback
  static std::vector<ColEntry<double>> r = {JTB_COLS(double, 5, 3, 32), JTB_COLS(double, 5, 3, 8), JTB_COLS(double, 6, 3, 32),
                                            JTB_COLS(double, 6, 3, 8),  JTB_COLS(double, 7, 4, 16), JTB_COLS(double, 7, 4, 8),
                                            JTB_COLS(double, 8, 4, 8),  JTB_COLS(double, 9, 3, 8),  JTB_COLS(double, 10, 4, 8)};
  return r;
}
template <> std::vector<ColEntry<float>>& colreg<float>() {
  static std::vector<ColEntry<float>> r = {JTB_COLS(float, 5, 3, 32), JTB_COLS(float, 6, 3, 16), JTB_COLS(float, 7, 4, 16),
                                           JTB_COLS(float, 8, 4, 16), JTB_COLS(float, 9, 3, 16), JTB_COLS(float, 10, 4, 16)};
  return r;
}

// realInverse rows
template <typename T> struct RfftInvEntry {
  int logn, loge, W, threads, smem;
  void (*kern)(const RfftInvParams<T>);
  unsigned attr_done;
};
template <typename T, int LOGN, int LOGE, int W> RfftInvEntry<T> mkrinv() {
  typedef Sched<LOGN, LOGE> S;
  RfftInvEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, false, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_rfft_inv_row_kernel<T, LOGN, LOGE, W>;
  e.attr_done = 0;
  return e;
}
template <typename T> std::vector<RfftInvEntry<T>>& rinvreg();
template <> std::vector<RfftInvEntry<double>>& rinvreg<double>() {
  static std::vector<RfftInvEntry<double>> r = {mkrinv<double, 5, 3, 32>(), mkrinv<double, 6, 3, 16>(), mkrinv<double, 7, 4, 16>(),
                                                mkrinv<double, 8, 4, 8>(),  mkrinv<double, 9, 3, 4>(),  mkrinv<double, 10, 3, 2>(),
                                                mkrinv<double, 11, 3, 1>(), mkrinv<double, 12, 4, 1>()};
  return r;
}
template <> std::vector<RfftInvEntry<float>>& rinvreg<float>() {
  static std::vector<RfftInvEntry<float>> r = {mkrinv<float, 5, 3, 32>(), mkrinv<float, 6, 3, 32>(), mkrinv<float, 7, 4, 16>(),
                                               mkrinv<float, 8, 4, 16>(), mkrinv<float, 9, 3, 8>(),  mkrinv<float, 10, 4, 4>(),
                                               mkrinv<float, 11, 4, 2>(), mkrinv<float, 12, 4, 1>()};
  return r;
}

template <typename E> int set_smem_once(E* f, int device) {
  if (!(f->attr_done & (1u << (device & 31)))) {
    JTB_CUDA(cudaFuncSetAttribute(f->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, f->smem));
    f->attr_done |= 1u << (device & 31);
  }
  return ST_OK;
}

}  // namespace

// inverse DCT-III / DST-III of contiguous real lines (line l at l*dist), in place; f0/f: input factors of a[0] / a[q>0]
template <typename T>
int fast_r2r_rows_inv(Engine<T>& e, T* a, i64 dist, i64 nlines, i64 n, int kind, T f0, T f, bool* handled) {
  *handled = false;
  if (g_inv_off || nlines <= 0 || !is_pow2(n) || n < 4 || (dist % 2) || ((uintptr_t)a % sizeof(cx<T>))) return ST_OK;
  if (kind != RK_DCT && kind != RK_DST) return ST_OK;
  if ((dist % 4) || ((uintptr_t)a % (4 * sizeof(T)))) return ST_OK;   // 4-real vector stores
  const int logN = ilog2(n) - 1;
  RowInvEntry<T>* r = nullptr;
  static const char* ele = getenv("JTB_ROW_LOGE");   // tuning knob: prefer the variant with this radix
  const int want_loge = ele ? atoi(ele) : 0;
  for (auto& x : rowinvreg<T>()) if (x.logn == logN && x.kind == kind && (!want_loge || x.loge == want_loge)) { r = &x; break; }
  if (!r) for (auto& x : rowinvreg<T>()) if (x.logn == logN && x.kind == kind) { r = &x; break; }
  if (!r) return ST_OK;
  JTB_TRY(set_smem_once(r, e.ctx->device));
  RowR2RParams<T> p;
  p.a = a; p.nlines = nlines; p.dist = dist; p.f0 = f0; p.f = f; p.pair_rows = 0; p.prefetch = 0;
  JTB_TRY(fast_stage_table<T>(e, logN, r->loge, &p.twg));
  const cx<T>* tw[JTB_MAX_STAGES];
  JTB_TRY(e.tile_tables(logN, tw, &p.rtw));
  JTB_TRY(e.dct_table(n, &p.dtw));
  const i64 nblk = (nlines + r->W - 1) / r->W;
  if (nblk > 0x7fffffffLL) return ST_OK;
  JTB_LAUNCH(r->kern, (unsigned)nblk, (unsigned)r->threads, (size_t)r->smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}

// realInverse of contiguous packed lines (N = 2^logN complex slots per line, line l at l*dist complex), in place
template <typename T>
int fast_rfft_inv(Engine<T>& e, cx<T>* a, i64 dist, i64 nlines, int logN, bool has_scale, T scale, bool* handled) {
  *handled = false;
  static const bool off = getenv("JTB_NO_RFFTINV") != nullptr;
  if (g_inv_off || off || nlines <= 0) return ST_OK;
  RfftInvEntry<T>* r = nullptr;
  for (auto& x : rinvreg<T>()) if (x.logn == logN) { r = &x; break; }
  if (!r) return ST_OK;
  JTB_TRY(set_smem_once(r, e.ctx->device));
  RfftInvParams<T> p;
  p.a = a; p.nlines = nlines; p.dist = dist; p.has_scale = has_scale; p.scale = scale;
  JTB_TRY(fast_stage_table<T>(e, logN, r->loge, &p.twg));
  const cx<T>* tw[JTB_MAX_STAGES];
  JTB_TRY(e.tile_tables(logN, tw, &p.rtw));
  const i64 nblk = (nlines + r->W - 1) / r->W;
  if (nblk > 0x7fffffffLL) return ST_OK;
  JTB_LAUNCH(r->kern, (unsigned)nblk, (unsigned)r->threads, (size_t)r->smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}
template int fast_rfft_inv<double>(Engine<double>&, double2*, i64, i64, int, bool, double, bool*);
template int fast_rfft_inv<float>(Engine<float>&, float2*, i64, i64, int, bool, float, bool*);

// forward or inverse DCT/DST (forward DHT) along the strided axis of length n <= 1024 in ONE pass (fft_col_r2r_kernel)
template <typename T>
int fast_r2r_cols_single(Engine<T>& e, T* a, i64 n, i64 Cn, i64 batches, i64 bdist, int kind, bool inverse, T f0, T f,
                         bool* handled) {
  typedef cx<T> C;
  *handled = false;
  if (g_inv_off || !is_pow2(n) || (Cn % 2) || (bdist % 2) || ((uintptr_t)a % sizeof(C)) || batches < 1) return ST_OK;
  if (inverse && kind != RK_DCT && kind != RK_DST) return ST_OK;
  const int logn = ilog2(n);
  const i64 H = Cn / 2;
  ColEntry<T>* f1 = nullptr;
  for (auto& x : colreg<T>())
    if (x.logn == logn && (x.inv != 0) == inverse && H % x.W == 0) { f1 = &x; break; }
  if (!f1) return ST_OK;
  const i64 nblk = (H / f1->W) * batches;
  if (nblk > 0x7fffffffLL) return ST_OK;
  JTB_TRY(set_smem_once(f1, e.ctx->device));
  ColR2RParams<T> p;
  p.a = (C*)a; p.s = H; p.bdist = bdist / 2; p.cols = (int)H; p.batches = (int)batches; p.kind = kind; p.f0 = f0; p.f = f;
  p.dtw = nullptr;
  JTB_TRY(fast_stage_table<T>(e, f1->logn, f1->loge, &p.twg));
  if (kind != RK_DHT) JTB_TRY(e.dct_table(n, &p.dtw));
  JTB_LAUNCH(f1->kern, (unsigned)nblk, (unsigned)f1->threads, (size_t)f1->smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}

// inverse DCT-III / DST-III along the strided axis of length n of `batches` row-major [n][Cn] real arrays (batch
// distance bdist reals): two adjacent real columns travel as one complex column through a two-pass inverse FFT.
template <typename T>
int fast_r2r_cols_inv(Engine<T>& e, T* a, i64 n, i64 Cn, i64 batches, i64 bdist, int kind, T f0, T f, bool* handled) {
  typedef cx<T> C;
  *handled = false;
  if (g_inv_off || !is_pow2(n) || (Cn % 2) || (bdist % 2) || ((uintptr_t)a % sizeof(C)) || batches < 1) return ST_OK;
  if (kind != RK_DCT && kind != RK_DST) return ST_OK;
  JTB_TRY(fast_r2r_cols_single<T>(e, a, n, Cn, batches, bdist, kind, true, f0, f, handled));
  if (*handled) return ST_OK;
  const int logn = ilog2(n);
  PairInvEntry<T>* f1 = nullptr;
  UnpermEntry<T>* f2 = nullptr;
  const int pre = kind == RK_DCT ? PRE_UNPERM_DCT : PRE_UNPERM_DST;
  const i64 H = Cn / 2;
  for (auto& x : pairinvreg<T>()) {
    for (auto& y : unpermreg<T>()) {
      if (x.logn + y.logn != logn || y.pre != pre || (H % x.W) || (H % y.W)) continue;
      if (!f1 || x.logn < f1->logn) { f1 = &x; f2 = &y; }
    }
  }
  if (!f1) return ST_OK;
  const i64 s = H, R2 = 1LL << f1->logn, R1 = 1LL << f2->logn, bd = bdist / 2;
  const i64 ext = (batches - 1) * bd + n * s;
  JTB_TRY(e.ctx->ensure(e.ctx->work[WK_FOURSTEP], (size_t)ext * sizeof(C)));
  C* wk = (C*)e.ctx->work[WK_FOURSTEP].p;
  C* ac = (C*)a;
  const C *fsA, *fsB, *dtw;
  int logL;
  JTB_TRY(e.fs_tables(logn, &fsA, &fsB, &logL));
  JTB_TRY(e.dct_table(n, &dtw));
  JTB_TRY(set_smem_once(f1, e.ctx->device));
  JTB_TRY(set_smem_once(f2, e.ctx->device));
  // pass A: rows q = k1 + R1*q2 -> work rows k1*R2 + m2
  ColPairParams<T> cp;
  cp.z = ac; cp.out = wk; cp.s = s; cp.bdist = bd; cp.zs = s; cp.zbdist = bd;
  cp.R1 = (int)R1; cp.cols = (int)H; cp.batches = (int)batches; cp.kind = kind; cp.f0 = f0; cp.f = f; cp.dtw = dtw;
  JTB_TRY(fast_stage_table<T>(e, f1->logn, f1->loge, &cp.twg));
  const i64 nblk = (H / f1->W) * (R1 / 2 + 1) * batches;
  if (nblk > 0x7fffffffLL) return ST_OK;
  JTB_LAUNCH(f1->kern, (unsigned)nblk, (unsigned)f1->threads, (size_t)f1->smem, e.st, cp, fsA, fsB, logL);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  // pass B: lines (c, m2, batch): FFT over q1 (work rows q1*R2 + m2), output m = m1*R2 + m2 -> un-permuted row of a
  Fast2Params<T> q;
  memset(&q, 0, sizeof q);
  q.scale = 1;
  q.in = wk; q.out = ac;
  q.nlines = H * R2 * batches; q.c0 = (int)H; q.gmod = (int)R2;
  q.in_gdist = s; q.in_gdist2 = bd; q.in_cdist = 1; q.in_stride = R2 * s;
  q.out_gdist2 = bd; q.out_cdist = 1;
  q.pre_n = n; q.pre_s = s;
  q.swap_out = 1;
  JTB_TRY(fast_stage_table<T>(e, f2->logn, f2->loge, &q.twg));
  const i64 nblk2 = q.nlines / f2->W;
  if (nblk2 > 0x7fffffffLL) { set_error("too many lines"); return ST_UNSUPPORTED; }
  JTB_LAUNCH(f2->kern, (unsigned)nblk2, (unsigned)f2->threads, (size_t)f2->smem, e.st, q);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}

template int fast_r2r_rows_inv<double>(Engine<double>&, double*, i64, i64, i64, int, double, double, bool*);
template int fast_r2r_rows_inv<float>(Engine<float>&, float*, i64, i64, i64, int, float, float, bool*);
template int fast_r2r_cols_single<double>(Engine<double>&, double*, i64, i64, i64, i64, int, bool, double, double, bool*);
template int fast_r2r_cols_single<float>(Engine<float>&, float*, i64, i64, i64, i64, int, bool, float, float, bool*);
template int fast_r2r_cols_inv<double>(Engine<double>&, double*, i64, i64, i64, i64, int, double, double, bool*);
template int fast_r2r_cols_inv<float>(Engine<float>&, float*, i64, i64, i64, i64, int, float, float, bool*);

}  // namespace jtb
