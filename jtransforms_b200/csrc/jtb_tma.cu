// Host dispatch of fft_tma_kernel (jtb_tma.cuh): tensor-map construction and the variant registry.
#include <cstdlib>
#include <cstring>

#include "jtb_engine_impl.cuh"
#include "jtb_tma.cuh"

namespace jtb {

#ifdef JTB_EMU
template <typename T>
int fast_tma_c2c(Engine<T>&, cx<T>*, const Geo&, i64, int, bool, bool, T, bool* handled) {
  *handled = false;
  return ST_OK;
}
#else

namespace {

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_fn get_encoder() {
  static encode_fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (encode_fn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

template <typename T> struct TmaEntry {
  int logn, W, ng, nst, out, threads, smem, loge;
  void (*kern)(const CUtensorMap, const TmaParams<T>);
  unsigned attr_done;
};
template <typename T, int LOGN, int LOGE, int W, int NG, int NST, int OUT> TmaEntry<T> make_tma() {
  typedef Sched<LOGN, LOGE> S;
  TmaEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.W = W; e.ng = NG; e.nst = NST; e.out = OUT;
  e.threads = NG * W * S::TPL;
  e.smem = (int)((NST * FastAddr<T, S, true, W>::TILE + ((FastTw<S>::COUNT_SM + 7) & ~7)) * sizeof(cx<T>) + NG * NST * 8 + 128);
  e.kern = fft_tma_kernel<T, LOGN, LOGE, W, NG, NST, OUT>;
  e.attr_done = 0;
  return e;
}
template <typename T> std::vector<TmaEntry<T>>& tma_registry();
template <> std::vector<TmaEntry<double>>& tma_registry<double>() {
  static std::vector<TmaEntry<double>> r = {
      // first entry of a length is the default; JTB_TMA_NG / JTB_TMA_NST / JTB_TMA_OUT select the others
      // (two compute groups over a three-stage ring faulted on the B200 -- profiles/r02_ab_tma.log -- and were removed)
      make_tma<double, 9, 3, 8, 2, 2, 0>(), make_tma<double, 9, 3, 8, 1, 3, 0>(), make_tma<double, 9, 3, 8, 1, 3, 1>(),
      make_tma<double, 9, 3, 8, 1, 2, 0>(),
  };
  return r;
}
template <> std::vector<TmaEntry<float>>& tma_registry<float>() {
  static std::vector<TmaEntry<float>> r = {
      make_tma<float, 9, 3, 16, 1, 3, 1>(), make_tma<float, 9, 3, 16, 1, 3, 0>(),
  };
  return r;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

// Strided in-place lines through the persistent TMA kernel.  *handled = false when the shape has no variant, the
// driver has no tensor-map encoder, or the path is switched off (JTB_TMA=0).
template <typename T>
int fast_tma_c2c(Engine<T>& e, cx<T>* a, const Geo& g, i64 nlines, int logn, bool inverse, bool has_scale, T scale,
                 bool* handled) {
  *handled = false;
  // Measured on B200, 512^3 strided passes: float 0.49 / 0.69 ms (lean kernel, k2 / k1) -> 0.43 / 0.43 ms; double
  // 0.69 / 0.80 -> 0.74 / 0.93 ms (two 512-thread CTAs per SM already overlap, the extra shared-memory round trip of
  // the staged tile costs more than the prefetch gains) -> default on for float only.  JTB_TMA=0/1 overrides.
  static const int mode = env_int("JTB_TMA", -1);
  const bool on = mode < 0 ? sizeof(T) == 4 : mode != 0;
  if (!on || nlines <= 0) return ST_OK;
  const i64 n = 1LL << logn;
  if (!(g.stride > 1 && g.d[0] == 1 && g.c[1] == 1 && g.c[2] == 1 && g.c[0] > 1 && nlines % g.c[0] == 0)) return ST_OK;
  if (((uintptr_t)a % 16) != 0 || g.stride >= (1LL << 31) || (g.stride * (i64)sizeof(cx<T>)) % 16 || (g.d[3] * (i64)sizeof(cx<T>)) % 16)
    return ST_OK;
  static const int want_ng = env_int("JTB_TMA_NG", 0), want_nst = env_int("JTB_TMA_NST", 0), want_out = env_int("JTB_TMA_OUT", -1);
  TmaEntry<T>* pick = nullptr;
  for (auto& f : tma_registry<T>()) {
    if (f.logn != logn || g.c[0] % f.W) continue;
    if (want_ng > 0 && f.ng != want_ng) continue;
    if (want_nst > 0 && f.nst != want_nst) continue;
    if (want_out >= 0 && f.out != want_out) continue;
    pick = &f;
    break;
  }
  if (!pick) return ST_OK;
  encode_fn enc = get_encoder();
  if (!enc) return ST_OK;
  const i64 ngroups = nlines / g.c[0];
  const i64 ntiles = nlines / pick->W;
  if (ntiles > 0x7fffffffLL || 2 * g.c[0] > 0xffffffffLL || ngroups > 0xffffffffLL) return ST_OK;
  // rank-3 tensor over the real components: x = 2*c0 contiguous reals, y = n rows `stride` complex apart,
  // z = line groups `d[3]` complex apart
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)(2 * g.c[0]), (cuuint64_t)n, (cuuint64_t)ngroups};
  const cuuint64_t strides[2] = {(cuuint64_t)(g.stride * (i64)sizeof(cx<T>)),
                                 (cuuint64_t)((ngroups > 1 ? g.d[3] : g.stride * n) * (i64)sizeof(cx<T>))};
  const cuuint32_t box[3] = {(cuuint32_t)(2 * pick->W), (cuuint32_t)(n > 256 ? 256 : n), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult cr = enc(&map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)a,
                          dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return ST_OK;   // shape outside what a tensor map can describe: the lean kernel takes it
  static int sms[32] = {0};
  const int dv = e.ctx->device & 31;
  if (!(pick->attr_done & (1u << dv))) {
    JTB_CUDA(cudaFuncSetAttribute(pick->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pick->smem));
    pick->attr_done |= 1u << dv;
  }
  if (!sms[dv]) JTB_CUDA(cudaDeviceGetAttribute(&sms[dv], cudaDevAttrMultiProcessorCount, e.ctx->device));
  TmaParams<T> p;
  p.a = a; p.ntiles = (int)ntiles; p.tiles_per_group = (int)(g.c[0] / pick->W);
  p.line_dist = g.d[3]; p.stride = (int)g.stride;
  p.inverse = inverse; p.has_scale = has_scale; p.scale = scale;
  JTB_TRY(fast_stage_table<T>(e, pick->logn, pick->loge, &p.twg));
  i64 grid = sms[dv];
  if (grid > ntiles) grid = ntiles;
  JTB_LAUNCH(pick->kern, (unsigned)grid, (unsigned)pick->threads, (size_t)pick->smem, e.st, map, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}
// ------------------------------------------------------------------------------------------ fused k2 + exchange, TMA stores
namespace {
template <typename T> struct ScatterTmaEntry {
  int logn, W, threads, smem, loge;
  void (*kern)(const PeerMaps, const ScatterParams<T>);
  unsigned attr_done;
};
template <typename T, int LOGN, int LOGE, int W> ScatterTmaEntry<T> make_scatter_tma() {
  typedef Sched<LOGN, LOGE> S;
  ScatterTmaEntry<T> e;
  e.logn = LOGN; e.loge = LOGE; e.W = W; e.threads = W * S::TPL;
  e.smem = (int)((FastAddr<T, S, true, W>::TILE + FastTw<S>::COUNT_SM) * sizeof(cx<T>)) + 128;
  e.kern = fft_scatter_tma_kernel<T, LOGN, LOGE, W>;
  e.attr_done = 0;
  return e;
}
template <typename T> std::vector<ScatterTmaEntry<T>>& scatter_tma_registry();
template <> std::vector<ScatterTmaEntry<double>>& scatter_tma_registry<double>() {
  static std::vector<ScatterTmaEntry<double>> r = {make_scatter_tma<double, 9, 3, 8>(), make_scatter_tma<double, 8, 4, 8>(),
                                                   make_scatter_tma<double, 10, 4, 8>()};
  return r;
}
template <> std::vector<ScatterTmaEntry<float>>& scatter_tma_registry<float>() {
  static std::vector<ScatterTmaEntry<float>> r = {make_scatter_tma<float, 9, 3, 16>(), make_scatter_tma<float, 10, 4, 16>()};
  return r;
}
}  // namespace

// forward re-slabbing only (the inverse exchange interleaves rows of different sources: no rectangular box per peer)
template <typename T>
int fast_scatter_tma(Engine<T>& e, const cx<T>* a, i64 Ls, i64 R, i64 Cn, i64 S, int nranks, int rank, void* const* peers,
                     bool inverse, i64 col0, i64 ncols, bool* handled) {
  *handled = false;
  if (!is_pow2(R) || nranks < 2 || nranks > 8 || !is_pow2(nranks) || R % nranks) return ST_OK;
  const i64 Rh = R / nranks;
  if (Rh > 256 || S * Rh > 0x7fffffffLL) return ST_OK;   // box dimension limit
  const int logn = ilog2(R);
  ScatterTmaEntry<T>* pick = nullptr;
  for (auto& f : scatter_tma_registry<T>())
    if (f.logn == logn && Cn % f.W == 0) { pick = &f; break; }
  if (!pick) return ST_OK;
  if (ncols < 0) { col0 = 0; ncols = Cn; }
  if (col0 < 0 || col0 % pick->W || ncols % pick->W || col0 + ncols > Cn || ncols == 0) return ST_OK;
  if ((Cn * (i64)sizeof(cx<T>)) % 16) return ST_OK;
  encode_fn enc = get_encoder();
  if (!enc) return ST_OK;
  PeerMaps maps;
  memset(&maps, 0, sizeof maps);
  for (int h = 0; h < nranks; ++h) {
    if (((uintptr_t)peers[h]) % 16) return ST_OK;
    const cuuint64_t dims[2] = {(cuuint64_t)(2 * Cn), (cuuint64_t)(S * Rh)};
    const cuuint64_t strides[1] = {(cuuint64_t)(Cn * (i64)sizeof(cx<T>))};
    const cuuint32_t box[2] = {(cuuint32_t)(2 * pick->W), (cuuint32_t)Rh};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&maps.m[h], sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, peers[h],
                            dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return ST_OK;
  }
  const int dv = e.ctx->device & 31;
  if (!(pick->attr_done & (1u << dv))) {
    JTB_CUDA(cudaFuncSetAttribute(pick->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pick->smem));
    pick->attr_done |= 1u << dv;
  }
  ScatterParams<T> p;
  memset(&p, 0, sizeof p);
  p.a = a;
  p.Ls = (int)Ls; p.C = (int)Cn; p.logRh = ilog2(Rh); p.slice0 = (int)(rank * Ls); p.inverse = inverse;
  p.row_base = (long long)p.slice0 * Rh; p.row_ls_mul = (int)Rh; p.row_mul = 1;
  p.col0 = (int)col0; p.groups = (int)(ncols / pick->W);
  JTB_TRY(fast_stage_table<T>(e, logn, pick->loge, &p.twg));
  const i64 nblk = Ls * (ncols / pick->W);
  if (nblk > 0x7fffffffLL) return ST_OK;
  JTB_LAUNCH(pick->kern, (unsigned)nblk, (unsigned)pick->threads, (size_t)pick->smem, e.st, maps, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}
#endif

#ifdef JTB_EMU
template <typename T>
int fast_scatter_tma(Engine<T>&, const cx<T>*, i64, i64, i64, i64, int, int, void* const*, bool, i64, i64, bool* handled) {
  *handled = false;
  return ST_OK;
}
#endif
template int fast_scatter_tma<double>(Engine<double>&, const double2*, i64, i64, i64, i64, int, int, void* const*, bool, i64, i64, bool*);
template int fast_scatter_tma<float>(Engine<float>&, const float2*, i64, i64, i64, i64, int, int, void* const*, bool, i64, i64, bool*);

template int fast_tma_c2c<double>(Engine<double>&, double2*, const Geo&, i64, int, bool, bool, double, bool*);
template int fast_tma_c2c<float>(Engine<float>&, float2*, const Geo&, i64, int, bool, bool, float, bool*);

}  // namespace jtb
