// The tile FFT kernel: one CTA transforms W lines of N = 2^LOGN complex points.
//
// Replaces the reference's cftfsub/cftbsub radix-4 split-radix recursion plus
// bitrv2/bitrv2conj (utils/CommonUtils.java:708-793, :824-2070, :3722-4033) for
// every line that fits one CTA.  Stockham autosort: each thread keeps E = 2^b0
// points in registers, performs register butterflies of radix 2^bits(s) per
// stage and exchanges through (padded) shared memory between stages, so no
// bit-reversal pass exists and the first load / last store touch HBM exactly
// once, coalesced (thread t owns elements t + q*T of its line in EVERY stage).
//
// Prologues/epilogues fuse the neighbouring element-wise steps of the
// reference (real split rftfsub/rftbsub, the four-step twiddle, Bluestein chirp
// multiplies, scale) into the same HBM sweep.
#pragma once
#include "jtb_butterfly.cuh"

namespace jtb {

enum { JTB_MAX_STAGES = 5 };

// stage schedule ---------------------------------------------------------------
template <int LOGN, int LOGE> struct Sched {
  static constexpr int S = (LOGN <= LOGE) ? 1 : (LOGN + LOGE - 1) / LOGE;
  static constexpr int BASE = LOGN / S;
  static constexpr int REM = LOGN % S;
  __host__ __device__ static constexpr int bits(int s) { return BASE + (s < REM ? 1 : 0); }
  static constexpr int LOGE0 = bits(0);          // elements per thread (max radix)
  static constexpr int E = 1 << LOGE0;
  static constexpr int N = 1 << LOGN;
  static constexpr int TPL = N / E;              // threads per line
  __host__ __device__ static constexpr int ns(int s) { int n = 1; for (int i = 0; i < s; ++i) n <<= bits(i); return n; }
  static constexpr int LOGPAD = LOGE0;           // one pad element every E elements
  static constexpr int LD = N + (N >> LOGPAD) + 1;   // padded line length (odd-ish)
  static constexpr int MAXT = (TPL > 512) ? TPL : ((E <= 8) ? 512 : (TPL > 256 ? TPL : 256));
};

enum ProMode { PRO_DIRECT = 0, PRO_RFFT_INV = 1, PRO_BLUE_MID = 2 };
enum EpiMode { EPI_DIRECT = 0, EPI_REMAP = 1, EPI_RFFT_FWD = 2 };

template <typename T> struct TileParams {
  const cx<T>* in;
  cx<T>* out;
  Geo gi, go;
  i64 nlines;
  int W;        // lines per CTA
  int wfast;    // 1: adjacent lines are adjacent in memory (thread index runs over w first)
  // re<->im swaps turn the forward kernel into the inverse DFT (IDFT(x) = swap(DFT(swap(x)))).
  // load:  raw -> swap_in -> premul -> swap_in2 ;  store: swap_out1 -> fs twiddle/postmul/scale -> swap_out
  int swap_in, swap_in2, swap_out1, swap_out;
  i64 line_base;           // first line handled by block 0 (chunked launches)
  int has_scale;
  T scale;
  const cx<T>* tw[JTB_MAX_STAGES];   // per-stage twiddles, layout [(r-1)*Ns + k]
  // four-step twiddle applied at the store: out[k] *= W_big^(idx*k), idx = geometry index of level fs_level
  int fs_mode, fs_level, fs_logL;
  const cx<T>* fsA;   // W_big^(L*m)
  const cx<T>* fsB;   // W_big^(m)
  // Logical element index of element k of a line: e = k*ks + idx*is, idx = geometry index of
  // level *_level in 0..3 (plain lines: ks=1,is=0; four-step passes: the position inside the big transform).
  i64 lin_ks, lin_is;  int lin_level;
  i64 lout_ks, lout_is; int lout_level;
  i64 valid_in;       // PRO_DIRECT: logical elements e >= valid_in read as 0 (zero padding); <0: off
  i64 valid_out;      // EPI_DIRECT/REMAP: logical elements e >= valid_out are not stored; <0: off
  // element-wise chirp/filter multiply fused at load (premul[e]) / store (postmul[e])
  const cx<T>* premul;   int premul_conj;
  const cx<T>* postmul;  int postmul_conj;
  int pro, epi;
  const cx<T>* rtw;   // real split table: exp(-2 pi i k / (2N)), k in [0, N/2]
};

template <typename T, typename S> __device__ __forceinline__ int smem_addr(int i, int w, int W, int wfast) {
  int phys = i + (i >> S::LOGPAD);
  return wfast ? phys * W + w : w * S::LD + phys;
}

// one butterfly stage on the register file ---------------------------------------
template <typename T, typename S, int s> struct Stage {
  static constexpr int LOGR = S::bits(s);
  static constexpr int R = 1 << LOGR;
  static constexpr int NB = S::E / R;     // butterflies per thread in this stage
  static constexpr int NS = S::ns(s);

  __device__ static __forceinline__ void compute(cx<T>* v, int t, const cx<T>* __restrict__ tw) {
#pragma unroll
    for (int m = 0; m < NB; ++m) {
      cx<T> x[R];
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = v[m + r * NB];
      if (s > 0) {
        const int k = (t + m * S::TPL) & (NS - 1);
#pragma unroll
        for (int r = 1; r < R; ++r) x[r] = cmul(x[r], __ldg(tw + (r - 1) * NS + k));
      }
      Bfly<T, R>::run(x);
#pragma unroll
      for (int r = 0; r < R; ++r) v[m + r * NB] = x[r];
    }
  }
  // scatter the stage outputs to their Stockham positions in shared memory
  __device__ static __forceinline__ void scatter(const cx<T>* v, cx<T>* sm, int t, int w, int W, int wfast) {
#pragma unroll
    for (int m = 0; m < NB; ++m) {
      const int jv = t + m * S::TPL;
      const int k = jv & (NS - 1);
      const int j0 = ((jv - k) << LOGR) + k;   // (jv / NS) * NS * R + k
#pragma unroll
      for (int r = 0; r < R; ++r) sm[smem_addr<T, S>(j0 + r * NS, w, W, wfast)] = v[m + r * NB];
    }
  }
};

template <typename T, typename S, int s> struct StageLoop {
  __device__ static __forceinline__ void run(cx<T>* v, cx<T>* sm, int t, int w, int W, int wfast,
                                             const TileParams<T>& p, bool need_sync_first) {
    Stage<T, S, s>::compute(v, t, p.tw[s]);
    if (s + 1 < S::S) {
      if (s > 0 || need_sync_first) __syncthreads();
      Stage<T, S, s>::scatter(v, sm, t, w, W, wfast);
      __syncthreads();
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = sm[smem_addr<T, S>(t + q * S::TPL, w, W, wfast)];
      StageLoop<T, S, (s + 1 < S::S ? s + 1 : s)>::run(v, sm, t, w, W, wfast, p, false);
    }
  }
};

__host__ __device__ __forceinline__ i64 lvl_pick(const i64* v, int l) { return l == 0 ? v[0] : (l == 1 ? v[1] : (l == 2 ? v[2] : v[3])); }

// W_big^(m) from the two half tables
template <typename T> __device__ __forceinline__ cx<T> fs_twiddle(const TileParams<T>& p, i64 m) {
  const i64 lo = m & ((1LL << p.fs_logL) - 1);
  const i64 hi = m >> p.fs_logL;
  return cmul(__ldg(p.fsA + hi), __ldg(p.fsB + lo));
}

template <typename T, int LOGN, int LOGE>
__global__ void __launch_bounds__(Sched<LOGN, LOGE>::MAXT, (Sched<LOGN, LOGE>::MAXT <= 512 ? 2 : 1)) fft_tile_kernel(const TileParams<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  const int W = p.W, wfast = p.wfast;
  const int tid = threadIdx.x;
  int w, t;
  if (wfast) { w = tid % W; t = tid / W; } else { t = tid % S::TPL; w = tid / S::TPL; }
  const i64 line0 = p.line_base + (i64)blockIdx.x * W;
  const i64 line = line0 + w;
  const bool valid = line < p.nlines;
  i64 li[4] = {0, 0, 0, 0};   // level indices of this thread's line
  C v[S::E];
  bool staged = false;

  // ---------------------------------------------------------------- load
  if (p.pro == PRO_DIRECT) {
    const C* src = p.in + (valid ? geo_off(p.gi, line, li) : 0);
    const i64 st = p.gi.stride;
    const i64 ebase = lvl_pick(li, p.lin_level) * p.lin_is;
#pragma unroll
    for (int q = 0; q < S::E; ++q) {
      const int k = t + q * S::TPL;
      const i64 e = ebase + k * p.lin_ks;
      C z = mk<T>(0, 0);
      if (valid && (p.valid_in < 0 || e < p.valid_in)) {
        z = src[k * st];
        if (p.swap_in) z = cswap(z);
        if (p.premul) { C m = __ldg(p.premul + e); z = p.premul_conj ? cmulc(z, m) : cmul(z, m); }
        if (p.swap_in2) z = cswap(z);
      }
      v[q] = z;
    }
  } else if (p.pro == PRO_RFFT_INV) {
    // Packed real half-spectrum X (N complex slots: slot 0 = (Re X[0], Re X[N]), slot k = X[k])
    // -> input Z of the length-N complex inverse FFT whose output is z[j] = x[2j] + i x[2j+1]
    // (the role of rftbsub + the a[0]/a[1] fix-up, fft/DoubleFFT_1D.java:956-962):
    //   Z[k] = (X[k] + conj X[N-k])/2 + i * conj(w^k) * (X[k] - conj X[N-k])/2,  w = exp(-2 pi i/(2N))
    // The inverse runs on the forward kernel in the swapped domain, hence the cswap at the store.
    staged = true;
    const T hf = (T)0.5;
    const int total = W * (S::N / 2 + 1);
    for (int idx = tid; idx < total; idx += blockDim.x) {
      int ww, k;
      if (wfast) { ww = idx % W; k = idx / W; } else { k = idx % (S::N / 2 + 1); ww = idx / (S::N / 2 + 1); }
      const i64 ln = line0 + ww;
      if (ln >= p.nlines) continue;
      const C* src = p.in + geo_off(p.gi, ln);
      const i64 st = p.gi.stride;
      if (k == 0) {
        C a = src[0];              // (Re X[0], Re X[N])
        C z = mk<T>((a.x + a.y) * hf, (a.x - a.y) * hf);
        sm[smem_addr<T, S>(0, ww, W, wfast)] = cswap(z);
      } else if (2 * k == S::N) {
        C a = src[(i64)k * st];
        sm[smem_addr<T, S>(k, ww, W, wfast)] = cswap(mk<T>(a.x, -a.y));
      } else if (2 * k < S::N) {
        C a = src[(i64)k * st];            // X[k]
        C b = src[(i64)(S::N - k) * st];   // X[N-k]
        C wk = __ldg(p.rtw + k);           // exp(-2 pi i k / 2N)
        C ev = mk<T>((a.x + b.x) * hf, (a.y - b.y) * hf);   // (a + conj b)/2
        C df = mk<T>((a.x - b.x) * hf, (a.y + b.y) * hf);   // (a - conj b)/2
        C od = cmulc(df, wk);                 // df * conj(w^k)
        od = mk<T>(-od.y, od.x);              // * i
        C zk = cadd(ev, od);                          // Z[k]
        C zn = mk<T>(ev.x - od.x, -(ev.y - od.y));   // Z[N-k] = conj(ev - od)
        sm[smem_addr<T, S>(k, ww, W, wfast)] = cswap(zk);
        sm[smem_addr<T, S>(S::N - k, ww, W, wfast)] = cswap(zn);
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = sm[smem_addr<T, S>(t + q * S::TPL, w, W, wfast)];
  }

  // ---------------------------------------------------------------- stages
  StageLoop<T, S, 0>::run(v, sm, t, w, W, wfast, p, staged);

  // ---------------------------------------------------------------- store
  if (p.epi == EPI_DIRECT) {
    if (valid) {
      C* dst = p.out + geo_off(p.go, line, li);
      const i64 st = p.go.stride;
      const i64 fidx = lvl_pick(li, p.fs_level);
      const i64 ebase = lvl_pick(li, p.lout_level) * p.lout_is;
#pragma unroll
      for (int q = 0; q < S::E; ++q) {
        const int k = t + q * S::TPL;
        const i64 e = ebase + k * p.lout_ks;
        if (p.valid_out >= 0 && e >= p.valid_out) continue;   // truncated output: neither stored nor looked up (postmul has valid_out entries)
        C z = v[q];
        if (p.swap_out1) z = cswap(z);
        if (p.fs_mode) z = cmul(z, fs_twiddle(p, fidx * k));
        if (p.postmul) { C m = __ldg(p.postmul + e); z = p.postmul_conj ? cmulc(z, m) : cmul(z, m); }
        if (p.has_scale) { z.x *= p.scale; z.y *= p.scale; }
        if (p.swap_out) z = cswap(z);
        dst[k * st] = z;
      }
    }
  } else {
    if (S::S > 1 || staged) __syncthreads();
#pragma unroll
    for (int q = 0; q < S::E; ++q) sm[smem_addr<T, S>(t + q * S::TPL, w, W, wfast)] = v[q];
    __syncthreads();
    if (p.epi == EPI_REMAP) {
      // cooperative store with the line index running fastest (transposed store of the
      // four-step second pass): consecutive threads write consecutive lines' element k
      const int total = W * S::N;
      for (int idx = tid; idx < total; idx += blockDim.x) {
        const int ww = idx % W, k = idx / W;
        const i64 ln = line0 + ww;
        if (ln >= p.nlines) continue;
        i64 lj[4];
        const i64 off = geo_off(p.go, ln, lj);
        const i64 e = lvl_pick(lj, p.lout_level) * p.lout_is + (i64)k * p.lout_ks;
        if (p.valid_out >= 0 && e >= p.valid_out) continue;
        C z = sm[smem_addr<T, S>(k, ww, W, wfast)];
        if (p.swap_out1) z = cswap(z);
        if (p.postmul) { C m = __ldg(p.postmul + e); z = p.postmul_conj ? cmulc(z, m) : cmul(z, m); }
        if (p.has_scale) { z.x *= p.scale; z.y *= p.scale; }
        if (p.swap_out) z = cswap(z);
        p.out[off + (i64)k * p.go.stride] = z;
      }
    } else if (p.epi == EPI_RFFT_FWD) {
      // real split (the role of rftfsub, utils/CommonUtils.java:5750-5776) + JTransforms packing:
      //   X[k] = (Z[k] + conj Z[N-k])/2 - i/2 * w^k (Z[k] - conj Z[N-k]),  w = exp(-2 pi i/(2N))
      //   out[0] = (Re X[0], Re X[N]);  out[k] = X[k], 0 < k < N
      const int total = W * (S::N / 2 + 1);
      for (int idx = tid; idx < total; idx += blockDim.x) {
        int ww, k;
        if (wfast) { ww = idx % W; k = idx / W; } else { k = idx % (S::N / 2 + 1); ww = idx / (S::N / 2 + 1); }
        const i64 ln = line0 + ww;
        if (ln >= p.nlines) continue;
        C* dst = p.out + geo_off(p.go, ln);
        const i64 st = p.go.stride;
        if (k == 0) {
          C z = sm[smem_addr<T, S>(0, ww, W, wfast)];
          C o = mk<T>(z.x + z.y, z.x - z.y);
          if (p.has_scale) { o.x *= p.scale; o.y *= p.scale; }
          dst[0] = o;
        } else if (2 * k == S::N) {
          C z = sm[smem_addr<T, S>(k, ww, W, wfast)];
          C o = mk<T>(z.x, -z.y);
          if (p.has_scale) { o.x *= p.scale; o.y *= p.scale; }
          dst[(i64)k * st] = o;
        } else if (2 * k < S::N) {
          C a = sm[smem_addr<T, S>(k, ww, W, wfast)];
          C b = sm[smem_addr<T, S>(S::N - k, ww, W, wfast)];
          C wk = __ldg(p.rtw + k);
          const T hf = (T)0.5;
          C ev = mk<T>((a.x + b.x) * hf, (a.y - b.y) * hf);   // (a + conj b)/2
          C df = mk<T>((a.x - b.x) * hf, (a.y + b.y) * hf);   // (a - conj b)/2
          C od = cmul(df, wk);
          od = mk<T>(od.y, -od.x);                              // * (-i)
          C xk = cadd(ev, od);
          C xn = mk<T>(ev.x - od.x, -(ev.y - od.y));            // X[N-k] = conj(ev - od)
          if (p.has_scale) { xk.x *= p.scale; xk.y *= p.scale; xn.x *= p.scale; xn.y *= p.scale; }
          dst[(i64)k * st] = xk;
          dst[(i64)(S::N - k) * st] = xn;
        }
      }
    }
  }
}

}  // namespace jtb
