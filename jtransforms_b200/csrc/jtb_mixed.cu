// Host side of the mixed-radix line kernel (jtb_mixed.cuh): factorisation, table, launch geometry.
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
  static bool attr_done[16] = {false, false};
  const int dv = (e.ctx->device & 7) * 2 + (sizeof(T) == 8 ? 0 : 1);
  if (!attr_done[dv]) {
    JTB_CUDA(cudaFuncSetAttribute(fft_mixed_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap + 1024));
    attr_done[dv] = true;
  }
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
  const unsigned threads = work >= 512 ? 512u : (work >= 256 ? 256u : (work >= 128 ? 128u : 64u));
  const i64 nblk = (nlines + W - 1) / W;
  if (nblk > 0x7fffffffLL) return ST_OK;
  JTB_LAUNCH(fft_mixed_kernel<T>, (unsigned)nblk, threads, smem, e.st, p);
  JTB_CUDA(cudaGetLastError());
  e.ctx->launches++;
  *handled = true;
  return ST_OK;
}

template int mixed_c2c<double>(Engine<double>&, double2*, const Geo&, i64, i64, bool, bool, double, bool*);
template int mixed_c2c<float>(Engine<float>&, float2*, const Geo&, i64, i64, bool, bool, float, bool*);

}  // namespace jtb
