// Instantiations and host-side dispatch of fft_fast_kernel (jtb_fast.cuh).
#include <cstdlib>

#include "jtb_engine_impl.cuh"
#include "jtb_fast.cuh"

namespace jtb {

namespace {

template <typename T> struct FastEntry {
  int logn, loge, strided, W, threads, smem, twcount, nstages;
  int bits[JTB_MAX_STAGES];
  void (*kern)(const FastParams<T>);
  unsigned attr_done;   // bit d set: smem attribute applied on device d
};

template <typename T, int LOGN, int LOGE, bool STRIDED, int W> FastEntry<T> make_entry() {
  typedef Sched<LOGN, LOGE> S;
  FastEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.strided = STRIDED; e.W = W; e.threads = W * S::TPL;
  e.twcount = FastTw<S>::COUNT;
  e.smem = (int)((FastAddr<T, S, STRIDED, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.nstages = S::S;
  for (int s = 0; s < JTB_MAX_STAGES; ++s) e.bits[s] = s < S::S ? S::bits(s) : 0;
  e.kern = fft_fast_kernel<T, LOGN, LOGE, STRIDED, W>;
  e.attr_done = 0;
  return e;
}

template <typename T> std::vector<FastEntry<T>>& registry();

template <> std::vector<FastEntry<double>>& registry<double>() {
  static std::vector<FastEntry<double>> r = {
      // first entry of each (logn, layout) is the default; the others are tuning variants (JTB_FAST_W)
      make_entry<double, 9, 3, true, 8>(),  make_entry<double, 9, 3, true, 4>(),   // W = 16 (one 1024-thread CTA per SM): 1.98 -> 2.92 ms at 512^3
      make_entry<double, 9, 3, false, 4>(), make_entry<double, 9, 3, false, 2>(), make_entry<double, 9, 3, false, 1>(),
      make_entry<double, 9, 3, false, 8>(),
      make_entry<double, 6, 3, true, 8>(),  make_entry<double, 6, 3, true, 16>(), make_entry<double, 6, 3, true, 32>(),
      make_entry<double, 6, 3, false, 16>(), make_entry<double, 6, 3, false, 32>(),
      make_entry<double, 10, 4, true, 8>(), make_entry<double, 10, 4, true, 4>(),
      make_entry<double, 10, 4, false, 4>(), make_entry<double, 10, 4, false, 2>(),
      make_entry<double, 12, 4, false, 1>(), make_entry<double, 12, 4, false, 2>(),
      make_entry<double, 13, 4, false, 1>(),
      make_entry<double, 8, 4, true, 8>(), make_entry<double, 8, 4, false, 8>(),
      make_entry<double, 7, 4, true, 16>(), make_entry<double, 7, 4, false, 16>(),
      make_entry<double, 11, 4, false, 2>(), make_entry<double, 11, 4, false, 4>(),
  };
  return r;
}
template <> std::vector<FastEntry<float>>& registry<float>() {
  static std::vector<FastEntry<float>> r = {
      make_entry<float, 9, 3, true, 8>(),  make_entry<float, 9, 3, true, 16>(),   // W = 16 is one 1024-thread CTA per SM: 6 % slower
      make_entry<float, 9, 3, false, 4>(),  make_entry<float, 9, 3, false, 2>(),
      make_entry<float, 10, 4, true, 16>(), make_entry<float, 10, 4, true, 8>(),
      make_entry<float, 10, 4, false, 4>(), make_entry<float, 10, 4, false, 2>(),
      make_entry<float, 11, 4, true, 8>(),
      make_entry<float, 12, 4, false, 1>(), make_entry<float, 13, 4, false, 1>(),
      make_entry<float, 8, 4, true, 16>(), make_entry<float, 8, 4, false, 16>(),
      make_entry<float, 7, 4, true, 16>(), make_entry<float, 7, 4, false, 16>(),
      make_entry<float, 6, 3, true, 16>(), make_entry<float, 6, 3, false, 32>(),
      make_entry<float, 11, 4, false, 2>(), make_entry<float, 11, 4, false, 4>(),
  };
  return r;
}

}  // namespace

// base twiddle table of Sched<logn, loge> (layout: FastTw in jtb_fast.cuh), cached per (precision, logn, loge)
template <typename T> int fast_stage_table(Engine<T>& e, int logn, int loge, const cx<T>** out) {
  typedef cx<T> C;
  const std::string key = mkkey("ftw", e.pname(), logn, loge);
  void* d = e.ctx->table(key);
  if (!d) {
    const int Sn = (logn <= loge) ? 1 : (logn + loge - 1) / loge;
    std::vector<C> h;
    i64 ns = 1;
    for (int s = 0; s < Sn; ++s) {
      const int b = logn / Sn + (s < logn % Sn ? 1 : 0);
      const i64 R = 1LL << b;
      if (s > 0)
        for (int j = 0; j < b; ++j)
          for (i64 k = 0; k < ns; ++k) h.push_back(unit_root<T>((1LL << j) * k, ns * R));
      ns *= R;
    }
    if (h.empty()) h.push_back(mk<T>(1, 0));
    JTB_TRY(e.ctx->put_table(key, h.data(), h.size() * sizeof(C), &d));
  }
  *out = (const C*)d;
  return ST_OK;
}
template int fast_stage_table<double>(Engine<double>&, int, int, const double2**);
template int fast_stage_table<float>(Engine<float>&, int, int, const float2**);

template <typename T>
int fast_c2c_out(Engine<T>& e, cx<T>* a, const Geo& g, cx<T>* out, i64 out_dist, i64 out_stride, i64 nlines, int logn,
                 bool inverse, bool has_scale, T scale, bool* handled);

// Runs the lean kernel when the call is a plain in-place transform on a supported layout.
// *handled = false (and ST_OK) when the caller must use the general tile kernel.
template <typename T>
int fast_c2c(Engine<T>& e, cx<T>* a, const Geo& g, i64 nlines, int logn, bool inverse, bool has_scale, T scale,
             bool* handled) {
  return fast_c2c_out<T>(e, a, g, nullptr, 0, 0, nlines, logn, inverse, has_scale, scale, handled);
}

// `out` != null (strided layouts only): the transformed line groups are stored to out + group*out_dist with element
// stride out_stride instead of in place
template <typename T>
int fast_c2c_out(Engine<T>& e, cx<T>* a, const Geo& g, cx<T>* out, i64 out_dist, i64 out_stride, i64 nlines, int logn,
                 bool inverse, bool has_scale, T scale, bool* handled) {
  *handled = false;
  static const bool off = getenv("JTB_NO_FAST") != nullptr;
  if (off || nlines <= 0) return ST_OK;
  if (((uintptr_t)a % sizeof(cx<T>)) != 0) return ST_OK;
  const i64 n = 1LL << logn;
  bool strided;
  if (g.stride == 1 && g.c[0] == 1 && g.c[1] == 1 && g.c[2] == 1) strided = false;
  else if (g.stride > 1 && g.d[0] == 1 && g.c[1] == 1 && g.c[2] == 1 && g.c[0] > 1 &&
           (n - 1) * g.stride + g.c[0] < 0x7fffffffLL && nlines % g.c[0] == 0) strided = true;
  else return ST_OK;
  if (out && (!strided || out_stride >= 0x7fffffffLL)) return ST_OK;
  if (strided && !out) {   // persistent TMA-fed kernel (jtb_tma.cuh) where a variant exists
    JTB_TRY(fast_tma_c2c<T>(e, a, g, nlines, logn, inverse, has_scale, scale, handled));
    if (*handled) return ST_OK;
  }
  const char* ev = getenv(strided ? "JTB_FAST_WS" : "JTB_FAST_WC");
  const int wwant = ev ? atoi(ev) : 0;
  FastEntry<T>* pick = nullptr;
  for (auto& f : registry<T>()) {
    if (f.logn != logn || (f.strided != 0) != strided) continue;
    if (strided && (g.c[0] % f.W) != 0) continue;
    if (wwant > 0 && f.W != wwant) continue;
    pick = &f;
    break;
  }
  if (!pick) return ST_OK;
  if (!(pick->attr_done & (1u << (e.ctx->device & 31)))) {
    JTB_CUDA(cudaFuncSetAttribute(pick->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pick->smem));
    pick->attr_done |= 1u << (e.ctx->device & 31);
  }
  FastParams<T> p;
  p.a = a; p.nlines = nlines;
  p.line_dist = g.d[3]; p.c0 = (int)g.c[0]; p.stride = (int)g.stride;
  p.inverse = inverse; p.has_scale = has_scale; p.scale = scale;
  JTB_TRY(fast_stage_table<T>(e, pick->logn, pick->loge, &p.twg));
  p.out = out ? out : a; p.out_line_dist = out ? out_dist : g.d[3]; p.out_stride = out ? (int)out_stride : (int)g.stride;
  p.reps = 1;
  p.prefetch = 0;
  if (strided) {
    // lines whose elements are >= 2 MB apart touch one TLB page per element: let a CTA reuse its translations for a
    // few neighbouring column groups (JTB_FAST_REPS overrides)
    const char* er = getenv("JTB_FAST_REPS");
    const i64 bytes_stride = g.stride * (i64)sizeof(cx<T>);
    p.reps = er ? atoi(er) : 1;   // measured: no gain from 2..8 on the 4 MiB-stride pass
    if (p.reps < 1) p.reps = 1;
    while (p.reps > 1 && ((g.c[0] / pick->W) % p.reps) != 0) --p.reps;
    // L2 prefetch of the tile this many CTAs ahead.  Measured on B200, 512^3 double: row-stride (8 KiB) column pass
    // 0.694 -> 0.660 ms at 74..111 (0.645 at 148); slice-stride (4 MiB) pass 0.80 -> 1.10 ms (every prefetched row
    // is its own page) -> on for strides up to 64 KiB only.  JTB_FAST_PREFETCH overrides (0 = off).
    static const char* epf = getenv("JTB_FAST_PREFETCH");
    const int pf_default = (sizeof(T) == 8 && bytes_stride <= (64 << 10)) ? 111 : 0;
    p.prefetch = p.reps == 1 ? (epf ? atoi(epf) : pf_default) : 0;
  }
  i64 nblk = ((nlines + pick->W - 1) / pick->W + p.reps - 1) / p.reps;
  if (strided && e.cta_limit > 0 && nblk > e.cta_limit) {
    const i64 ntiles = (nlines + pick->W - 1) / pick->W;
    p.reps = (int)((ntiles + e.cta_limit - 1) / e.cta_limit);
    p.prefetch = 0;
    nblk = (ntiles + p.reps - 1) / p.reps;
  }
  if (nblk > 0x7fffffffLL) return ST_OK;
  // Measured and removed (profiles/r01_sweep_cluster.log, r01_sweep_raster.log): cluster launch of adjacent column
  // groups, ld.global.L2::256B hints and interleaved CTA rasterisation all slow the strided passes down.
  JTB_LAUNCH(pick->kern, (unsigned)nblk, (unsigned)pick->threads, (size_t)pick->smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}

// ------------------------------------------------------------------------------------------ fused k2 + exchange
namespace {
template <typename T> struct ScatterEntry {
  int logn, W, threads, smem, loge;
  void (*kern)(const ScatterParams<T>);
  unsigned attr_done;   // bit d set: smem attribute applied on device d
};
template <typename T, int LOGN, int LOGE, int W> ScatterEntry<T> make_scatter() {
  typedef Sched<LOGN, LOGE> S;
  ScatterEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, true, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_scatter_kernel<T, LOGN, LOGE, W>;
  e.attr_done = 0;
  return e;
}
template <typename T> std::vector<ScatterEntry<T>>& scatter_registry();
template <> std::vector<ScatterEntry<double>>& scatter_registry<double>() {
  static std::vector<ScatterEntry<double>> r = {make_scatter<double, 9, 3, 8>(), make_scatter<double, 9, 3, 16>(),
                                                make_scatter<double, 6, 3, 8>(), make_scatter<double, 10, 4, 8>(),
                                                make_scatter<double, 7, 4, 16>(), make_scatter<double, 8, 4, 8>()};
  return r;
}
template <> std::vector<ScatterEntry<float>>& scatter_registry<float>() {
  static std::vector<ScatterEntry<float>> r = {make_scatter<float, 9, 3, 16>(), make_scatter<float, 6, 3, 16>(),
                                               make_scatter<float, 7, 4, 16>(), make_scatter<float, 8, 4, 16>(),
                                               make_scatter<float, 10, 4, 16>(), make_scatter<float, 11, 4, 8>()};
  return r;
}
}  // namespace

template <typename T>
int fast_scatter(Engine<T>& e, const cx<T>* a, i64 Ls, i64 R, i64 Cn, int nranks, int rank, void* const* peers,
                 bool inverse, i64 slice_base, bool back, i64 col0, i64 ncols) {
  if (!is_pow2(R) || nranks < 1 || nranks > 8 || !is_pow2(nranks) || R % nranks) {
    set_error("fused exchange needs power-of-two rows and 1, 2, 4 or 8 ranks");
    return ST_UNSUPPORTED;
  }
  {
    // exchange stores issued by the TMA engine instead of the load/store unit (JTB_SCATTER_TMA=1; jtb_tma.cuh)
    static const bool tma = getenv("JTB_SCATTER_TMA") && atoi(getenv("JTB_SCATTER_TMA")) != 0;
    if (tma && !back && slice_base < 0) {
      bool handled = false;
      JTB_TRY(fast_scatter_tma<T>(e, a, Ls, R, Cn, Ls * nranks, nranks, rank, peers, inverse, col0, ncols, &handled));
      if (handled) return ST_OK;
    }
  }
  const int logn = ilog2(R);
  ScatterEntry<T>* pick = nullptr;
  const char* ew = getenv("JTB_SCATTER_W");
  const int wwant = ew ? atoi(ew) : 0;
  for (auto& f : scatter_registry<T>())
    if (f.logn == logn && Cn % f.W == 0 && (wwant <= 0 || f.W == wwant)) { pick = &f; break; }
  if (!pick) { set_error("no fused-exchange kernel for %lld rows x %lld columns", (long long)R, (long long)Cn); return ST_UNSUPPORTED; }
  if (Ls * (Cn / pick->W) > 0x7fffffffLL || Ls * R * Cn >= (1LL << 40)) { set_error("slab too large"); return ST_UNSUPPORTED; }
  if (ncols < 0) { col0 = 0; ncols = Cn; }
  if (col0 < 0 || col0 % pick->W || ncols % pick->W || col0 + ncols > Cn) { set_error("bad column window"); return ST_UNSUPPORTED; }
  if (ncols == 0) return ST_OK;
  if (!(pick->attr_done & (1u << (e.ctx->device & 31)))) {
    JTB_CUDA(cudaFuncSetAttribute(pick->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pick->smem));
    pick->attr_done |= 1u << (e.ctx->device & 31);
  }
  ScatterParams<T> p;
  p.a = a;
  for (int h = 0; h < 8; ++h) p.peer[h] = h < nranks ? (cx<T>*)peers[h] : nullptr;
  p.Ls = (int)Ls; p.C = (int)Cn; p.logRh = ilog2(R / nranks); p.slice0 = slice_base >= 0 ? (int)slice_base : (int)(rank * Ls); p.inverse = inverse;
  if (back) {   // inverse re-slabbing: one [S][Rh*C] block, output row k1 -> [k1 % Ls][rank*Rh + r][c] of its owner
    if (Ls != 1) { set_error("internal: the inverse exchange takes the block as one slice"); return ST_ARG; }
    p.row_base = rank; p.row_ls_mul = 0; p.row_mul = nranks;
  } else {
    p.row_base = (long long)p.slice0 * (R / nranks); p.row_ls_mul = (int)(R / nranks); p.row_mul = 1;
  }
  JTB_TRY(fast_stage_table<T>(e, logn, pick->loge, &p.twg));
  p.col0 = (int)col0; p.groups = (int)(ncols / pick->W);
  const unsigned nblk = (unsigned)(Ls * (ncols / pick->W));
  JTB_LAUNCH(pick->kern, nblk, (unsigned)pick->threads, (size_t)pick->smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  return ST_OK;
}

// ------------------------------------------------------------------------------------------ fused in-slice passes
// rows + columns of `nslices` N x N slices in one cooperative launch (fft_slice2d_kernel); optional fused exchange
template <typename T>
int fast_slice2d(Engine<T>& e, cx<T>* a, i64 nslices, i64 N, bool inverse, bool has_scale, T scale, int nranks, int rank,
                 void* const* peers, bool* handled) {
  *handled = false;
#ifndef JTB_EMU
  // Measured on B200 at 512^3.  Single GPU: the fused kernel moves less DRAM traffic (5.7 GB vs 8.6 GB) but is
  // slower (1.57 ms vs 1.30 ms for the two separate passes: team-barrier stalls) -> off.  With the fused exchange
  // (peers != null) the column phase is NVLink-bound and the row phase hides under it: 8 GPUs 0.537 -> 0.497 ms
  // -> on.  JTB_SLICE2D=0/1 overrides.
  static const char* ev_s2d = getenv("JTB_SLICE2D");
  const bool off = ev_s2d ? atoi(ev_s2d) == 0 : peers == nullptr;
  if (off || nslices < 1 || N != 512 || sizeof(T) != 8 || nslices > 65536) return ST_OK;
  typedef void (*kern_t)(const Slice2DParams<T>);
  kern_t kern = (kern_t)fft_slice2d_kernel<double, 9, 3, 8>;
  typedef Sched<9, 3> S;
  const int threads = 8 * S::TPL;
  const size_t tile = FastAddr<double, S, false, 8>::TILE > FastAddr<double, S, true, 8>::TILE
                          ? FastAddr<double, S, false, 8>::TILE : FastAddr<double, S, true, 8>::TILE;
  const size_t smem = (tile + FastTw<S>::COUNT_SM) * sizeof(double2);
  static bool attr_done[16] = {false};
  static int occ[16] = {0}, sms[16] = {0};
  const int dv = e.ctx->device & 15;
  if (!attr_done[dv]) {
    JTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    JTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[dv], kern, threads, smem));
    JTB_CUDA(cudaDeviceGetAttribute(&sms[dv], cudaDevAttrMultiProcessorCount, e.ctx->device));
    int coop = 0;
    JTB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e.ctx->device));
    if (!coop) occ[dv] = 0;
    attr_done[dv] = true;
  }
  const char* ev = getenv("JTB_TEAM");
  int team = ev ? atoi(ev) : 16;
  if (team < 1 || (N / 8) % team || team > 64) team = 16;
  int grid = occ[dv] * sms[dv];
  grid -= grid % team;
  if ((i64)(grid / team) > nslices) grid = (int)nslices * team;
  if (grid < team) return ST_OK;
  // scratch: counters + error flag
  struct Scratch { int* p; size_t n; };
  static Scratch sc[16] = {};
  if (sc[dv].n < (size_t)nslices + 1) {
    if (sc[dv].p) { JTB_CUDA(cudaDeviceSynchronize()); JTB_CUDA(cudaFree(sc[dv].p)); }
    sc[dv].n = (size_t)nslices + 1 < 1024 ? 1024 : (size_t)nslices + 1;
    JTB_CUDA(cudaMalloc((void**)&sc[dv].p, sc[dv].n * sizeof(int)));
  }
  JTB_CUDA(cudaMemsetAsync(sc[dv].p, 0, ((size_t)nslices + 1) * sizeof(int), e.st));
  Slice2DParams<T> p;
  memset(&p, 0, sizeof p);
  p.a = a; p.nslices = (int)nslices; p.team = team;
  JTB_TRY(e.ctx->ensure_watchdog());
  p.counters = sc[dv].p + 1; p.err = e.ctx->wd_dev;   // time-outs are read back by Ctx::check_watchdog
  JTB_TRY(fast_stage_table<T>(e, 9, 3, &p.twg));
  p.inverse = inverse; p.has_scale = has_scale; p.scale = scale;
  {
    const char* epl = getenv("JTB_SLICE2D_PIPE");
    p.pipelined = epl ? atoi(epl) : 1;
  }
  if (peers) {
    p.scatter = 1; p.logRh = ilog2(N / nranks); p.slice0 = (int)(rank * nslices);
    for (int h = 0; h < 8; ++h) p.peer[h] = h < nranks ? (cx<T>*)peers[h] : nullptr;
  }
  void* args[] = {(void*)&p};
  JTB_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3((unsigned)threads), args, smem, e.st));
  e.ctx->launches++;
  *handled = true;
#endif
  return ST_OK;
}
template int fast_slice2d<double>(Engine<double>&, double2*, i64, i64, bool, bool, double, int, int, void* const*, bool*);
template int fast_slice2d<float>(Engine<float>&, float2*, i64, i64, bool, bool, float, int, int, void* const*, bool*);

// ------------------------------------------------------------------------------------------ fused exchange + k1 pipeline
namespace {
template <typename T> struct PipeEntry {
  int logn, W, threads, smem, loge;
  void (*kern)(const PipeParams<T>);
  unsigned attr_done;
  int occ[32];
};
template <typename T, int LOGN, int LOGE, int W> PipeEntry<T> make_pipe() {
  typedef Sched<LOGN, LOGE> S;
  PipeEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, true, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>));
  e.kern = fft_pipe_kernel<T, LOGN, LOGE, W>;
  e.attr_done = 0;
  for (int d = 0; d < 32; ++d) e.occ[d] = 0;
  return e;
}
template <typename T> std::vector<PipeEntry<T>>& pipe_registry();
template <> std::vector<PipeEntry<double>>& pipe_registry<double>() {
  static std::vector<PipeEntry<double>> r = {make_pipe<double, 9, 3, 8>(), make_pipe<double, 6, 3, 8>(), make_pipe<double, 8, 4, 8>(),
                                             make_pipe<double, 10, 4, 8>()};
  return r;
}
template <> std::vector<PipeEntry<float>>& pipe_registry<float>() {
  static std::vector<PipeEntry<float>> r = {make_pipe<float, 9, 3, 16>(), make_pipe<float, 6, 3, 16>()};
  return r;
}
}  // namespace

template <typename T> bool fast_pipe_has(i64 R, i64 S, i64 Cn, int nranks, int nb) {
#ifdef JTB_EMU
  (void)R; (void)S; (void)Cn; (void)nranks; (void)nb;
  return false;
#else
  if (R != S || !is_pow2(R) || nranks < 2 || nranks > 8 || !is_pow2(nranks) || nb < 1 || nb > 16) return false;
  for (auto& f : pipe_registry<T>())
    if (f.logn == ilog2(R) && Cn % (f.W * nb) == 0) return true;
  return false;
#endif
}
template bool fast_pipe_has<double>(i64, i64, i64, int, int);
template bool fast_pipe_has<float>(i64, i64, i64, int, int);

// One launch: exchange (k2 pass, peer stores) of `nb` column blocks interleaved with the slice-axis pass of the blocks
// that have arrived (fft_pipe_kernel).  Needs R == S (one line length), power-of-two rank counts, zeroed counters
// (1 + nb ints) and flag arrays of 8 x 16 int64 per rank.  *handled = false when the shape has no variant.
template <typename T>
int fast_pipe_exchange(Engine<T>& e, const cx<T>* a, i64 Ls, i64 R, i64 Cn, int nranks, int rank, void* const* peers,
                       void* const* flag_ptrs, long long epoch, int* counters, int nb, bool inverse, bool has_scale, T scale,
                       bool* handled) {
  *handled = false;
#ifndef JTB_EMU
  if (!is_pow2(R) || nranks < 2 || nranks > 8 || !is_pow2(nranks) || R % nranks || Ls * nranks != R || nb < 1 || nb > 16) return ST_OK;
  const int logn = ilog2(R);
  PipeEntry<T>* pick = nullptr;
  for (auto& f : pipe_registry<T>())
    if (f.logn == logn && Cn % (f.W * nb) == 0) { pick = &f; break; }
  if (!pick) return ST_OK;
  const i64 gpb = Cn / nb / pick->W;
  if (2 * (i64)nb * Ls * gpb > 0x7fffffffLL || Ls * R * Cn >= (1LL << 40)) return ST_OK;
  const int dv = e.ctx->device & 31;
  if (!(pick->attr_done & (1u << dv))) {
    JTB_CUDA(cudaFuncSetAttribute(pick->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pick->smem));
    int sms = 0;
    JTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pick->occ[dv], pick->kern, pick->threads, pick->smem));
    JTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e.ctx->device));
    pick->occ[dv] *= sms;
    pick->attr_done |= 1u << dv;
  }
  if (pick->occ[dv] < 1) return ST_OK;
  PipeParams<T> p;
  memset(&p, 0, sizeof p);
  p.a = a;
  for (int h = 0; h < 8; ++h) { p.peer[h] = h < nranks ? (cx<T>*)peers[h] : nullptr; p.flags[h] = h < nranks ? (long long*)flag_ptrs[h] : nullptr; }
  p.recv = (cx<T>*)peers[rank];
  p.counters = counters;
  JTB_TRY(e.ctx->ensure_watchdog());
  p.err = e.ctx->wd_dev;
  JTB_TRY(fast_stage_table<T>(e, logn, pick->loge, &p.twg));
  p.epoch = epoch;
  p.Ls = (int)Ls; p.C = (int)Cn; p.logRh = ilog2(R / nranks); p.P = nranks; p.rank = rank; p.nb = nb; p.gpb = (int)gpb;
  p.inverse = inverse; p.has_scale = has_scale; p.scale = scale;
  i64 grid = pick->occ[dv];
  const i64 total = 2 * (i64)nb * Ls * gpb;
  if (grid > total) grid = total;
  JTB_LAUNCH(pick->kern, (unsigned)grid, (unsigned)pick->threads, (size_t)pick->smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
#endif
  return ST_OK;
}
template int fast_pipe_exchange<double>(Engine<double>&, const double2*, i64, i64, i64, int, int, void* const*, void* const*, long long, int*, int, bool, bool, double, bool*);
template int fast_pipe_exchange<float>(Engine<float>&, const float2*, i64, i64, i64, int, int, void* const*, void* const*, long long, int*, int, bool, bool, float, bool*);

// smallest column window granularity the fused exchange supports for R-point columns (0: no fused kernel)
template <typename T> int fast_scatter_width(i64 R, i64 Cn) {
  if (!is_pow2(R)) return 0;
  const int logn = ilog2(R);
  for (auto& f : scatter_registry<T>())
    if (f.logn == logn && Cn % f.W == 0) return f.W;
  return 0;
}
template int fast_scatter_width<double>(i64, i64);
template int fast_scatter_width<float>(i64, i64);

int peer_barrier(Ctx* ctx, cudaStream_t st, void* const* flag_ptrs, int nranks, int rank, long long epoch, int what) {
  if (nranks < 1 || nranks > 8) { set_error("1..8 ranks"); return ST_ARG; }
  PeerFlags pf;
  for (int h = 0; h < 8; ++h) pf.f[h] = h < nranks ? (long long*)flag_ptrs[h] : nullptr;
  JTB_TRY(ctx->ensure_watchdog());   // per-device flag; a time-out surfaces through Ctx::check_watchdog
  JTB_LAUNCH(peer_barrier_kernel, 1u, 32u, 0, st, pf, nranks, rank, epoch, ctx->wd_dev, what);
  JTB_CUDA(cudaGetLastError());
  ctx->launches++;
  return ST_OK;
}

template int fast_scatter<double>(Engine<double>&, const double2*, i64, i64, i64, int, int, void* const*, bool, i64, bool, i64, i64);
template int fast_scatter<float>(Engine<float>&, const float2*, i64, i64, i64, int, int, void* const*, bool, i64, bool, i64, i64);
// true when a lean strided kernel exists for 2^logn-point lines in groups of c0 adjacent lines
template <typename T> bool fast_has_strided(int logn, i64 c0) {
  for (auto& f : registry<T>())
    if (f.logn == logn && f.strided && c0 % f.W == 0) return true;
  return false;
}
template bool fast_has_strided<double>(int, i64);
template bool fast_has_strided<float>(int, i64);

template int fast_c2c<double>(Engine<double>&, double2*, const Geo&, i64, int, bool, bool, double, bool*);
template int fast_c2c<float>(Engine<float>&, float2*, const Geo&, i64, int, bool, bool, float, bool*);
template int fast_c2c_out<double>(Engine<double>&, double2*, const Geo&, double2*, i64, i64, i64, int, bool, bool, double, bool*);
template int fast_c2c_out<float>(Engine<float>&, float2*, const Geo&, float2*, i64, i64, i64, int, bool, bool, float, bool*);

}  // namespace jtb
