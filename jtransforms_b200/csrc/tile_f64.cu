// fft_tile_kernel instantiations and Engine<double> (DoubleFFT_*, DoubleDCT_*, ...).
#define JTB_TILE_T double
#include "jtb_tile_inst.cuh"
#include "jtb_engine_impl.cuh"

namespace jtb {
template <int LOGN> static void fill_info(TileInfo& t) {
  typedef Sched<LOGN, loge_for(LOGN)> S;
  t.logn = LOGN; t.n = S::N; t.nstages = S::S; t.e = S::E; t.tpl = S::TPL; t.ld = S::LD; t.maxt = S::MAXT;
  for (int s = 0; s < JTB_MAX_STAGES; ++s) t.bits[s] = s < S::S ? S::bits(s) : 0;
}
template <int LOGN> struct InfoDispatch {
  static void get(int logn, TileInfo& t) { if (logn == LOGN) fill_info<LOGN>(t); else InfoDispatch<LOGN - 1>::get(logn, t); }
};
template <> struct InfoDispatch<0> { static void get(int, TileInfo& t) { t.logn = -1; } };
TileInfo tile_info(int logn) { TileInfo t; t.logn = -1; InfoDispatch<14>::get(logn, t); return t; }

// Chirp tables of the Bluestein plan (fft/DoubleFFT_1D.java:1864-1890):
//   bk1[i] = exp(+i pi (i^2 mod 2n) / n),  bk2 = FFT_M(wrap(bk1)) / M,  M = nextPow2(2n - 1)
int blue_tables_f64(Ctx* ctx, cudaStream_t st, i64 n, const double2** bk1, const double2** bk2, i64* Mout) {
  const i64 M = next_pow2(2 * n - 1);
  *Mout = M;
  const std::string k1 = mkkey("bk1", "f64", n), k2 = mkkey("bk2", "f64", n);
  void* d1 = ctx->table(k1);
  void* d2 = ctx->table(k2);
  if (!d1 || !d2) {
    std::vector<double2> h1((size_t)n), hw((size_t)M);
    const double invM = 1.0 / (double)M;
    for (i64 i = 0; i < M; ++i) hw[(size_t)i] = mk<double>(0.0, 0.0);
    for (i64 i = 0; i < n; ++i) {
      const unsigned long long ph = ((unsigned long long)i * (unsigned long long)i) % (unsigned long long)(2 * n);
      double2 w = unit_root<double>((i64)ph, 2 * n);   // exp(-i pi ph / n)
      w.y = -w.y;                                      // exp(+i pi ph / n)
      h1[(size_t)i] = w;
      const double2 ws = mk<double>(w.x * invM, w.y * invM);
      hw[(size_t)i] = ws;
      if (i > 0) hw[(size_t)(M - i)] = ws;
    }
    JTB_TRY(ctx->put_table(k1, h1.data(), h1.size() * sizeof(double2), &d1));
    JTB_TRY(ctx->put_table(k2, hw.data(), hw.size() * sizeof(double2), &d2));
    Engine<double> e(ctx, st);
    Fuse<double> f;
    const Geo g = geo_contig(M);
    JTB_TRY(e.c2c_pow2((const double2*)d2, g, (double2*)d2, g, 0, 1, ilog2(M), f));
  }
  *bk1 = (const double2*)d1; *bk2 = (const double2*)d2;
  return ST_OK;
}

template struct Engine<double>;
}  // namespace jtb
