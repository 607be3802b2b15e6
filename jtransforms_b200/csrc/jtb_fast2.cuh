// Out-of-place variants of the lean line FFT kernel with one fused epilogue each:
//   FM_PLAIN      out has the layout class of in (own distances/strides)
//   FM_TWID       FM_PLAIN + four-step twiddle  out[k] *= W_n^(idx*k)  (first pass of the two-pass transform that
//                 replaces cftb1st + cftrec4_th for long lines, utils/CommonUtils.java:3282-3500, :3722-3795)
//   FM_TRANSPOSE  contiguous lines in, transposed (line index fastest) store out, staged through shared memory so it
//                 is made of full segments (second pass of the two-pass transform; replaces bitrv2conj,
//                 utils/CommonUtils.java:1650-2070)
//   FM_RFFT       contiguous lines in, real split + JTransforms packing out (rftfsub + a[0]/a[1] fix-up,
//                 utils/CommonUtils.java:5750-5776, fft/DoubleFFT_1D.java:524-546): realForward of 2N reals
#pragma once
#include "jtb_fast.cuh"

namespace jtb {

enum { FM_PLAIN = 0, FM_TWID = 1, FM_TRANSPOSE = 2, FM_RFFT = 3, FM_CHIRP_OUT = 4 };

template <typename T> struct Fast2Params {
  const cx<T>* in;
  cx<T>* out;
  i64 nlines;
  // line l -> c = l % c0, g = l / c0;  first element at (g % gmod)*gdist + (g / gmod)*gdist2 + c*cdist,
  // element j at + j*stride
  int c0, gmod;
  i64 in_gdist, out_gdist;
  i64 in_gdist2, out_gdist2;
  i64 in_cdist, out_cdist;
  i64 in_stride, out_stride;
  int swap_in, swap_out, has_scale;
  T scale;
  const cx<T>* twg;
  const cx<T>* fsA;   // FM_TWID: W^(L*h)
  const cx<T>* fsB;   //          W^(l)
  int fs_logL, tw_src;           // idx = tw_src ? g % gmod : c
  const cx<T>* rtw;   // FM_RFFT: exp(-2 pi i k / (2N)), k <= N/2
  // PRE != 0 (first pass of a long strided DCT/DST column): element j = r1*gmod + (g % gmod) of the column is
  // read from row perm(j) of the array (Makhoul even/odd permutation), rows are pre_s apart, column length pre_n
  i64 pre_n, pre_s;
  // PRE_CHIRP (first Bluestein pass): element with logical index m = j*in_stride + c*in_cdist is read as
  // in[m] * conj(chirp[m]) for m < pre_n and as 0 beyond (zero padding to the convolution length).
  // FM_CHIRP_OUT (last Bluestein pass): out[m] = v * conj(chirp[m]) for m = j*out_stride + c*out_cdist < out_n.
  const cx<T>* chirp;
  i64 out_n;
  int swap_out1;      // FM_CHIRP_OUT: undo the swapped-domain inverse before the chirp multiply
  // PRE_BIGTW (second pass of the strided sub-transform of a three-pass 1-D transform): the stored element, output row
  // k1 = (g % gmod) + gmod*j of column c, is multiplied by W_n^(k1*c) = bigA[m >> big_logL] * bigB[m & mask]
  const cx<T>* bigA;
  const cx<T>* bigB;
  int big_logL;
  int prefetch;       // FM_RFFT: prefetch.global.L2 of the lines `prefetch` CTAs ahead (0 = off)
};

// PRE_UNPERM_* (second pass of a long strided inverse DCT/DST column, jtb_r2r_inv.cuh) act on the STORE: output
// element m = j*gmod + (g % gmod) goes to row 2m (m < n/2) or 2(n-1-m)+1 of the array, DST negates the odd rows.
enum { PRE_NONE = 0, PRE_PERM_DCT = 1, PRE_PERM_DST = 2, PRE_CHIRP = 3, PRE_UNPERM_DCT = 4, PRE_UNPERM_DST = 5, PRE_BIGTW = 6 };

template <typename T> __device__ __forceinline__ cx<T> fs_tw2(const Fast2Params<T>& p, int m) {
  return cmul(__ldg(p.fsA + (m >> p.fs_logL)), __ldg(p.fsB + (m & ((1 << p.fs_logL) - 1))));
}

template <typename T, int LOGN, int LOGE, bool SIN, int MODE, int W, int PRE = PRE_NONE>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, (Sched<LOGN, LOGE>::E <= 8 ? FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB
                                                                                      : (FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB + 1) / 2))
fft_fast2_kernel(const Fast2Params<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, SIN, W> A;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  const int tid = threadIdx.x;
  int w, t;
  if (SIN) { w = tid % W; t = tid / W; } else { t = tid % S::TPL; w = tid / S::TPL; }
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W * S::TPL) twt[i] = __ldg(p.twg + i);

  const i64 line0 = (i64)blockIdx.x * W;
  const i64 line = line0 + w;
  const bool valid = line < p.nlines;
  const i64 g = line / p.c0;
  const int c = (int)(line - g * p.c0);
  const i64 g_hi = g / p.gmod;
  const int g_lo = (int)(g - g_hi * p.gmod);
#ifndef JTB_EMU
  if (MODE == FM_RFFT && p.prefetch > 0) {
    const i64 pl = line0 + (i64)p.prefetch * W + w;
    if (pl < p.nlines) {
      const char* pb = reinterpret_cast<const char*>(p.in + pl * p.in_gdist);
      for (int i = t * 128; i < (int)(S::N * sizeof(C)); i += S::TPL * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + i));
    }
  }
#endif
  // four-step twiddle of this thread's outputs: W^(idx*t) and the chain step W^(idx*TPL).  Fetched before the data so that
  // the two dependent table reads are in flight under the line loads instead of after the last butterfly.
  C ftw = mk<T>(1, 0), fws = mk<T>(1, 0);
  if (MODE == FM_TWID) {
    const int idx = p.tw_src ? g_lo : c;
    ftw = fs_tw2(p, idx * t);
    fws = fs_tw2(p, idx * S::TPL);
  }
  C v[S::E];
  if (valid && PRE == PRE_CHIRP) {
    const C* src = p.in + g_lo * p.in_gdist + g_hi * p.in_gdist2;
#pragma unroll
    for (int q = 0; q < S::E; ++q) {
      const i64 m = (t + q * S::TPL) * p.in_stride + c * p.in_cdist;
      C z = mk<T>(0, 0);
      if (m < p.pre_n) {
        z = src[m];
        if (p.swap_in) z = cswap(z);
        z = cmulc(z, __ldg(p.chirp + m));
      }
      v[q] = z;
    }
  } else if (valid && (PRE == PRE_PERM_DCT || PRE == PRE_PERM_DST)) {
    const C* src = p.in + g_hi * p.in_gdist2 + c * p.in_cdist;
#pragma unroll
    for (int q = 0; q < S::E; ++q) {
      // pre_n = gmod * N (host): j >= pre_n/2  <=>  t + q*TPL >= N/2  <=>  q >= E/2, a compile-time property of q, so
      // the DST sign folds into the first butterfly and all loads issue back to back
      const i64 j = (i64)(t + q * S::TPL) * p.gmod + g_lo;
      const bool second = q >= S::E / 2;
      const i64 row = second ? 2 * (p.pre_n - 1 - j) + 1 : 2 * j;
      C z = src[row * p.pre_s];
      if (PRE == PRE_PERM_DST && second) { z.x = -z.x; z.y = -z.y; }
      v[q] = z;
    }
  } else if (valid) {
    const C* src = p.in + g_lo * p.in_gdist + g_hi * p.in_gdist2 + c * p.in_cdist;
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = src[(t + q * S::TPL) * p.in_stride];
  } else {
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = mk<T>(0, 0);
  }
  if (p.swap_in && PRE != PRE_CHIRP) {
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
  }
  FastLoop<T, S, 0, SIN, W>::run(v, sm, twt, t, w, p.twg);
  if (MODE == FM_CHIRP_OUT) {
    if (valid) {
      C* dst = p.out + g_lo * p.out_gdist + g_hi * p.out_gdist2;
#pragma unroll
      for (int q = 0; q < S::E; ++q) {
        const i64 m = (t + q * S::TPL) * p.out_stride + c * p.out_cdist;
        if (m < p.out_n) {
          C z = v[q];
          if (p.swap_out1) z = cswap(z);
          z = cmulc(z, __ldg(p.chirp + m));
          if (p.has_scale) { z.x *= p.scale; z.y *= p.scale; }
          if (p.swap_out) z = cswap(z);
          dst[m] = z;
        }
      }
    }
    return;
  }
  if (p.has_scale) {
#pragma unroll
    for (int q = 0; q < S::E; ++q) { v[q].x *= p.scale; v[q].y *= p.scale; }
  }

  if (MODE == FM_PLAIN || MODE == FM_TWID) {
    if (MODE == FM_TWID) {
      C tw = ftw;
      const C ws = fws;
#pragma unroll
      for (int q = 0; q < S::E; ++q) {
        v[q] = cmul(v[q], tw);
        if (q + 1 < S::E) tw = cmul(tw, ws);
      }
    }
    if (PRE == PRE_BIGTW) {
      const i64 mask = (1LL << p.big_logL) - 1;
      const i64 m0 = ((i64)g_lo + (i64)p.gmod * t) * c, ms = (i64)p.gmod * S::TPL * c;
      C tw = cmul(__ldg(p.bigA + (m0 >> p.big_logL)), __ldg(p.bigB + (m0 & mask)));
      const C ws = cmul(__ldg(p.bigA + (ms >> p.big_logL)), __ldg(p.bigB + (ms & mask)));
#pragma unroll
      for (int q = 0; q < S::E; ++q) {
        v[q] = cmul(v[q], tw);
        if (q + 1 < S::E) tw = cmul(tw, ws);
      }
    }
    if (p.swap_out) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
    }
    if (valid && (PRE == PRE_UNPERM_DCT || PRE == PRE_UNPERM_DST)) {
      C* dst = p.out + g_hi * p.out_gdist2 + c * p.out_cdist;
#pragma unroll
      for (int q = 0; q < S::E; ++q) {
        const i64 m = (i64)(t + q * S::TPL) * p.gmod + g_lo;
        const bool second = q >= S::E / 2;          // pre_n = gmod * N
        const i64 row = second ? 2 * (p.pre_n - 1 - m) + 1 : 2 * m;
        C z = v[q];
        if (PRE == PRE_UNPERM_DST && second) { z.x = -z.x; z.y = -z.y; }
        dst[row * p.pre_s] = z;
      }
    } else if (valid) {
      C* dst = p.out + g_lo * p.out_gdist + g_hi * p.out_gdist2 + c * p.out_cdist;
#pragma unroll
      for (int q = 0; q < S::E; ++q) dst[(t + q * S::TPL) * p.out_stride] = v[q];
    }
  } else if (MODE == FM_TRANSPOSE) {
    if (p.swap_out) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
    }
    if (S::S > 1) __syncthreads();
#pragma unroll
    for (int q = 0; q < S::E; ++q) sm[A::at(t + q * S::TPL, w)] = v[q];
    __syncthreads();
    // the W lines of a CTA share g and are adjacent in c (host guarantees W | c0)
    const i64 g0 = line0 / p.c0;
    const int cbase = (int)(line0 - g0 * p.c0);
    const i64 g0_hi = g0 / p.gmod;
    C* dst = p.out + (g0 - g0_hi * p.gmod) * p.out_gdist + g0_hi * p.out_gdist2 + cbase * p.out_cdist;
    const int nl = (p.nlines - line0 < W) ? (int)(p.nlines - line0) : W;
    for (int idx = tid; idx < W * S::N; idx += W * S::TPL) {
      const int ww = idx % W, k = idx / W;
      if (ww < nl) dst[ww * p.out_cdist + k * p.out_stride] = sm[A::at(k, ww)];
    }
  } else if (MODE == FM_RFFT) {
    // Z = FFT_N(x[2j] + i x[2j+1]);  X[k] = (Z[k] + conj Z[N-k])/2 - i/2 w^k (Z[k] - conj Z[N-k]),  w = exp(-2 pi i/(2N))
    // packed: out[0] = (Re X[0], Re X[N]);  out[k] = X[k], 0 < k < N
    // Z[k] for k < N/2 stays in registers; only the upper half is handed over (unpadded: unit-stride accesses)
    if (S::S > 1) __syncthreads();
    C* half = sm + w * S::N;
#pragma unroll
    for (int q = S::E / 2; q < S::E; ++q) half[t + q * S::TPL] = v[q];
    __syncthreads();
    if (valid) {
      C* dst = p.out + g_lo * p.out_gdist + g_hi * p.out_gdist2 + c * p.out_cdist;
      const T hf = (T)0.5;
#pragma unroll
      for (int q = 0; q < S::E / 2; ++q) {
        const int k = t + q * S::TPL;   // 0 .. N/2 - 1
        if (k == 0) {
          const C z = v[0];
          dst[0] = mk<T>(z.x + z.y, z.x - z.y);
          const C zm = v[S::E / 2];
          dst[S::N / 2] = mk<T>(zm.x, -zm.y);
        } else {
          const C a = v[q];
          const C b = half[S::N - k];
          const C wk = __ldg(p.rtw + k);
          const C ev = mk<T>((a.x + b.x) * hf, (a.y - b.y) * hf);
          const C df = mk<T>((a.x - b.x) * hf, (a.y + b.y) * hf);
          C od = cmul(df, wk);
          od = mk<T>(od.y, -od.x);
          dst[k] = cadd(ev, od);
          dst[S::N - k] = mk<T>(ev.x - od.x, -(ev.y - od.y));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Middle pass of the two-pass Bluestein convolution (fft/DoubleFFT_1D.java:1920-2107 does it as cftbsub, a
// multiply loop and cftfsub): per contiguous row k1 of the [N1][N2] work array
//   row FFT (second pass of the forward transform, output k2 <-> frequency k1 + N1 k2)  ->  * bk2p[k1][k2]
//   ->  inverse row FFT (first pass of the inverse transform)  ->  * conj(W_M^(k1 m2))  -> store in place.
// The frequency-domain product never leaves the SM.
template <typename T> struct ConvParams {
  cx<T>* a;              // rows at (g*N1 + k1)*N2, g = transform within the chunk
  i64 nlines;            // rows
  int N1;                // rows per transform
  const cx<T>* h;        // bk2 permuted: h[k1*N2 + k2] = bk2[k1 + N1*k2] (1/M folded in)
  const cx<T>* twg;
  const cx<T>* fsA;
  const cx<T>* fsB;
  int fs_logL;
};

template <typename T, int LOGN, int LOGE, int W>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, (Sched<LOGN, LOGE>::E <= 8 ? FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB
                                                                                      : (FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB + 1) / 2))
fft_conv_kernel(const ConvParams<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, false, W> A;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  const int tid = threadIdx.x;
  const int t = tid % S::TPL, w = tid / S::TPL;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W * S::TPL) twt[i] = __ldg(p.twg + i);
  const i64 line = (i64)blockIdx.x * W + w;
  const bool valid = line < p.nlines;
  const int k1 = (int)(line % p.N1);
  C* row = p.a + line * S::N;
  C v[S::E];
  if (valid) {
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = row[t + q * S::TPL];
  } else {
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = mk<T>(0, 0);
  }
  FastLoop<T, S, 0, false, W>::run(v, sm, twt, t, w, p.twg);
  if (valid) {
    const C* h = p.h + (i64)k1 * S::N;
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = cswap(cmul(v[q], __ldg(h + t + q * S::TPL)));
  }
  if (S::S > 1) __syncthreads();
  FastLoop<T, S, 0, false, W>::run(v, sm, twt, t, w, p.twg);
  if (valid) {
    // conj(W_M^(k1*m2)), m2 = t + q*TPL, as a chain  conj(W^(k1 t)) * conj(W^(k1 TPL))^q
    const int L = (1 << p.fs_logL) - 1;
    const int m0 = k1 * t, ms = k1 * S::TPL;
    C tw = cmul(__ldg(p.fsA + (m0 >> p.fs_logL)), __ldg(p.fsB + (m0 & L)));
    const C ws = cmul(__ldg(p.fsA + (ms >> p.fs_logL)), __ldg(p.fsB + (ms & L)));
#pragma unroll
    for (int q = 0; q < S::E; ++q) {
      row[t + q * S::TPL] = cmulc(cswap(v[q]), tw);
      if (q + 1 < S::E) tw = cmul(tw, ws);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Fused forward DCT-II / DST-II / DHT of contiguous real lines of n = 2N reals (N = 2^LOGN):
//   coalesced load of the line into shared memory -> (Makhoul even/odd permutation for DCT/DST) -> N-point
//   complex FFT of z[j] = v[2j] + i v[2j+1] -> real split V[k] -> twiddle / cas combination -> store.
// Replaces per line: dct/DoubleDCT_1D.java:169-194 (pre-butterfly, rftbsub, cftbsub, dctsub, scale),
// dst/DoubleDST_1D.java:96-160, dht/DoubleDHT_1D.java:94-152.
enum { RK_DCT = 1, RK_DST = 2, RK_DHT = 3 };
// resident CTAs the radix-16 row kernels are compiled for (256-thread CTAs: 3 needs <= 85 registers)
#ifndef JTB_ROW_OCC3
#define JTB_ROW_OCC3 0
#endif
template <int THREADS> struct RowOcc16 { static constexpr int V = (JTB_ROW_OCC3 && THREADS == 256) ? 3 : (FastOcc<THREADS>::MINB + 1) / 2; };

template <typename T> struct RowR2RParams {
  T* a;                 // lines of n reals, line l at l*dist, transformed in place
  i64 nlines, dist;
  T f0, f;              // output factors for index 0 / the others
  const cx<T>* twg;     // stage twiddles
  const cx<T>* rtw;     // exp(-2 pi i k / n), k <= N/2
  const cx<T>* dtw;     // exp(-i pi k / (2n)), k < n   (DCT/DST)
  // RK_DHT, last pass of DoubleDHT_2D.forward (W even): the CTA's lines are the row PAIRS (r, pair_rows - r) of one
  // pair_rows x n array -- lines w = 2j, 2j+1 of a CTA are rows r and R-r; pair 0 is (0, R/2), both self-paired -- and
  // the epilogue applies yTransform (dht/DoubleDHT_2D.java:1288-1309) before the store:
  //   H[r][c] = (T[r][c] + T[R-r][c] + T[r][C-c] - T[R-r][C-c]) / 2.   0: plain rows.
  i64 pair_rows;
  int prefetch;         // > 0: prefetch.global.L2 of the lines the CTA `prefetch` blocks ahead will load (plain rows only)
};

template <typename T, int LOGN, int LOGE, int KIND, int W>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, (Sched<LOGN, LOGE>::E <= 8 ? FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB
                                                                                      : RowOcc16<W * Sched<LOGN, LOGE>::TPL>::V))
fft_r2r_row_kernel(const RowR2RParams<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, false, W> A;
  constexpr int N = S::N, n = 2 * S::N, H = S::E / 2;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  const int tid = threadIdx.x;
  const int t = tid % S::TPL, w = tid / S::TPL;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W * S::TPL) twt[i] = __ldg(p.twg + i);
  const i64 line0 = (i64)blockIdx.x * W;
  bool valid = line0 + w < p.nlines;
  i64 line = line0 + w;
  bool combine = false;          // paired rows: apply yTransform with the partner line w ^ 1
  if (KIND == RK_DHT && W >= 2 && p.pair_rows) {
    const i64 pi = (i64)blockIdx.x * (W / 2) + (w >> 1);
    valid = pi < p.pair_rows / 2;
    line = (w & 1) == 0 ? pi : (pi == 0 ? p.pair_rows / 2 : p.pair_rows - pi);
    combine = pi != 0;
  }
  T* xl = p.a + (valid ? line * p.dist : 0);
#ifndef JTB_EMU
  if (p.prefetch > 0 && !p.pair_rows) {
    // pull the lines of a later CTA into L2 while this one computes: n reals = n*sizeof(T)/128 cache lines per line
    const i64 pl = line0 + (i64)p.prefetch * W + w;
    if (pl < p.nlines) {
      const char* pb = reinterpret_cast<const char*>(p.a + pl * p.dist);
      for (int i = t * 128; i < (int)(n * sizeof(T)); i += S::TPL * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + i));
    }
  }
#endif
  C* half = sm + w * N;          // unpadded line buffer for the half-line hand-overs (conflict-free: unit stride)
  C v[S::E];
  // z[j] = v[2j] + i v[2j+1] of the Makhoul-permuted line v[u] = x[2u], v[n-1-u] = x[2u+1] (DST: odd samples negated):
  //   z[m] = (x[4m], x[4m+2]) and z[N-1-m] = (x[4m+3], x[4m+1]) for m < N/2.
  // A thread reads the 32 contiguous bytes x[4m..4m+3] of ITS m = t + q*TPL (q < E/2), keeps z[m] in registers and
  // hands z[N-1-m] to its owner (thread TPL-1-t) through shared memory: one half-line exchange instead of staging
  // the whole line and gathering it back.  DHT has no permutation: z[j] is one 16-byte load.
  if (KIND == RK_DHT) {
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = valid ? reinterpret_cast<const C*>(xl)[t + q * S::TPL] : mk<T>(0, 0);
  } else {
    T x0[H], x1[H], x2[H], x3[H];
#pragma unroll
    for (int q = 0; q < H; ++q) {
      const int m = t + q * S::TPL;
      if (valid) ld4(xl + 4 * m, x0[q], x1[q], x2[q], x3[q]);
      else x0[q] = x1[q] = x2[q] = x3[q] = 0;
    }
#pragma unroll
    for (int q = 0; q < H; ++q) {
      const int m = t + q * S::TPL;
      v[q] = mk<T>(x0[q], x2[q]);
      half[N - 1 - m] = (KIND == RK_DST) ? mk<T>(-x3[q], -x1[q]) : mk<T>(x3[q], x1[q]);
    }
    __syncthreads();
#pragma unroll
    for (int q = H; q < S::E; ++q) v[q] = half[t + q * S::TPL];
    __syncthreads();
  }
  FastLoop<T, S, 0, false, W>::run(v, sm, twt, t, w, p.twg);
  // real split: V[k] needs Z[k] (own registers, k < N/2) and Z[N-k] (upper half, handed over through shared memory)
  if (S::S > 1) __syncthreads();
#pragma unroll
  for (int q = H; q < S::E; ++q) half[t + q * S::TPL] = v[q];
  __syncthreads();
  if (KIND == RK_DHT && W >= 2 && p.pair_rows) {
    // Row pairs: the outputs of this line stay in registers -- (T[k], T[n-k]) in v[q], (T[N-k], T[N+k]) in v[q+H], both
    // mirrored column pairs (c, C-c) -- are handed to the partner line through shared memory, combined and stored.
    const T hf2 = (T)0.5;
#pragma unroll
    for (int q = 0; q < H; ++q) {
      const int k = t + q * S::TPL;
      if (k == 0) {
        const C z0 = v[0];
        const C zm = v[H];
        v[0] = mk<T>((z0.x + z0.y) * p.f, (z0.x - z0.y) * p.f);        // columns 0 and N: self-mirrored, never combined
        v[H] = mk<T>((zm.x + zm.y) * p.f, (zm.x - zm.y) * p.f);        // columns N/2 and n - N/2 (V[N/2] = conj Z[N/2])
      } else {
        const C a = v[q];
        const C b = half[N - k];
        const C wk = __ldg(p.rtw + k);
        const C ev = mk<T>((a.x + b.x) * hf2, (a.y - b.y) * hf2);
        const C df = mk<T>((a.x - b.x) * hf2, (a.y + b.y) * hf2);
        C od = cmul(df, wk);
        od = mk<T>(od.y, -od.x);
        const C Vk = cadd(ev, od);
        const C Vm = mk<T>(ev.x - od.x, -(ev.y - od.y));
        v[q] = mk<T>((Vk.x - Vk.y) * p.f, (Vk.x + Vk.y) * p.f);          // T[k], T[n-k]
        v[q + H] = mk<T>((Vm.x - Vm.y) * p.f, (Vm.x + Vm.y) * p.f);      // T[N-k], T[N+k]
      }
    }
    __syncthreads();                                                     // every read of `half` is done
#pragma unroll
    for (int q = 0; q < H; ++q) { half[t + q * S::TPL] = v[q]; half[t + (q + H) * S::TPL] = v[q + H]; }
    __syncthreads();
    if (!valid) return;
    const C* other = sm + (w ^ 1) * N;
#pragma unroll
    for (int q = 0; q < H; ++q) {
      const int k = t + q * S::TPL;
      C o1 = v[q], o2 = v[q + H];
      if (combine) {
        const C b1 = other[t + q * S::TPL], b2 = other[t + (q + H) * S::TPL];
        if (k != 0) o1 = mk<T>((o1.x + b1.x + o1.y - b1.y) * hf2, (o1.y + b1.y + o1.x - b1.x) * hf2);
        o2 = mk<T>((o2.x + b2.x + o2.y - b2.y) * hf2, (o2.y + b2.y + o2.x - b2.x) * hf2);
      }
      if (k == 0) {
        xl[0] = o1.x; xl[N] = o1.y;
        xl[N / 2] = o2.x; xl[n - N / 2] = o2.y;
      } else {
        xl[k] = o1.x; xl[n - k] = o1.y;
        xl[N - k] = o2.x; xl[N + k] = o2.y;
      }
    }
    return;
  }
  if (!valid) return;
  T* out = xl;
  const T hf = (T)0.5;
#pragma unroll
  for (int q = 0; q < H; ++q) {
    const int k = t + q * S::TPL;     // 0 .. N/2 - 1 ; pair (k, N - k)
    if (k == 0) {
      const C z0 = v[0];
      const T V0 = z0.x + z0.y, VN = z0.x - z0.y;          // V[0], V[N] (both real)
      const C zm = v[H];                                     // Z[N/2] (thread 0 owns index (E/2)*TPL)
      const C Vh = mk<T>(zm.x, -zm.y);                      // V[N/2] = conj Z[N/2]
      if (KIND == RK_DHT) {
        out[0] = V0 * p.f; out[N] = VN * p.f;
        out[N / 2] = (Vh.x - Vh.y) * p.f; out[n - N / 2] = (Vh.x + Vh.y) * p.f;
      } else {
        const C uh = cmul(Vh, __ldg(p.dtw + N / 2));
        const T cN = VN * (T)0.70710678118654752440084436210485L;
        if (KIND == RK_DCT) {
          out[0] = V0 * p.f0; out[N] = cN * p.f;
          out[N / 2] = uh.x * p.f; out[n - N / 2] = -uh.y * p.f;
        } else {   // DST: output index n-1-k of the DCT result
          out[n - 1] = V0 * p.f0; out[n - 1 - N] = cN * p.f;
          out[n - 1 - N / 2] = uh.x * p.f; out[N / 2 - 1] = -uh.y * p.f;
        }
      }
    } else {
      const C a = v[q];
      const C b = half[N - k];
      // DCT/DST: one table read, exp(-2 pi i k/n) and D(N-k) derived; DHT has no D table
      const C dk = (KIND == RK_DHT) ? mk<T>(0, 0) : __ldg(p.dtw + k);
      const C wk = (KIND == RK_DHT) ? __ldg(p.rtw + k) : dct_tw_pow4(dk);
      const C ev = mk<T>((a.x + b.x) * hf, (a.y - b.y) * hf);
      const C df = mk<T>((a.x - b.x) * hf, (a.y + b.y) * hf);
      C od = cmul(df, wk);
      od = mk<T>(od.y, -od.x);
      const C Vk = cadd(ev, od);                                   // V[k]
      const C Vm = mk<T>(ev.x - od.x, -(ev.y - od.y));             // V[N-k]
      if (KIND == RK_DHT) {
        out[k] = (Vk.x - Vk.y) * p.f;     out[n - k] = (Vk.x + Vk.y) * p.f;
        out[N - k] = (Vm.x - Vm.y) * p.f; out[N + k] = (Vm.x + Vm.y) * p.f;
      } else {
        const C uk = cmul(Vk, dk);
        const C um = cmul(Vm, dct_tw_nmk(dk));
        if (KIND == RK_DCT) {
          out[k] = uk.x * p.f;     out[n - k] = -uk.y * p.f;
          out[N - k] = um.x * p.f; out[N + k] = -um.y * p.f;
        } else {
          out[n - 1 - k] = uk.x * p.f;       out[k - 1] = -uk.y * p.f;
          out[n - 1 - (N - k)] = um.x * p.f; out[N - k - 1] = -um.y * p.f;
        }
      }
    }
  }
}

// Second column pass of the strided forward DCT-II / DST-II / DHT fused with the pair post-pass: a CTA transforms
// the two lines k1 and R1-k1 (length R2, rows k1*R2 + r2) of W adjacent complex columns, so that every output row
// k = k1 + R1*k2 finds its partner n-k = (R1-k1) + R1*(R2-1-k2) in the same CTA (through shared memory) and the
// final real results are stored directly -- no separate sweep for k_r2r_colpost.
template <typename T> struct ColPairParams {
  const cx<T>* z;       // first-pass output, rows zs apart, batch zbdist apart
  cx<T>* out;           // rows s apart, batch bdist apart
  i64 s, bdist, zs, zbdist;
  int R1, cols, batches; // lines per column, complex columns in this launch, arrays
  int kind;
  T f0, f;
  const cx<T>* twg;
  const cx<T>* dtw;
};

template <typename T, int LOGN, int LOGE, int W>
__global__ void __launch_bounds__(2 * W * Sched<LOGN, LOGE>::TPL, 2) fft_colpair_kernel(const ColPairParams<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  constexpr int R2 = S::N, W2 = 2 * W;            // the exchange tile holds 2*W "columns": (u, w)
  typedef FastAddr<T, S, true, W2> A;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  const int tid = threadIdx.x;
  const int wu = tid % W2, t = tid / W2;          // wu = u*W + w
  const int u = wu / W, w = wu - u * W;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W2 * S::TPL) twt[i] = __ldg(p.twg + i);
  // block -> (column group, line pair, batch)
  const int groups = p.cols / W, pairs = p.R1 / 2 + 1;
  int b = blockIdx.x;
  const int cg = b % groups; b /= groups;
  const int pr = b % pairs;
  const int batch = b / pairs;
  const int k1 = u == 0 ? pr : (p.R1 - pr) % p.R1;
  const i64 n = (i64)p.R1 * R2;
  const C* src = p.z + batch * p.zbdist + (i64)k1 * R2 * p.zs + cg * W + w;
  C v[S::E];
#pragma unroll
  for (int q = 0; q < S::E; ++q) v[q] = src[(i64)(t + q * S::TPL) * p.zs];
  FastLoop<T, S, 0, true, W2>::run(v, sm, twt, t, wu, p.twg);
  if (S::S > 1) __syncthreads();
#pragma unroll
  for (int q = 0; q < S::E; ++q) sm[A::at(t + q * S::TPL, wu)] = v[q];
  __syncthreads();
  const bool dup = (u == 1) && (k1 == pr);          // self-paired line handled twice: second copy does not store
  if (dup) return;
  C* dst = p.out + batch * p.bdist + cg * W + w;
  const T hf = (T)0.5;
#pragma unroll
  for (int q = 0; q < S::E; ++q) {
    const int k2 = t + q * S::TPL;
    const i64 k = k1 + (i64)p.R1 * k2;
    // partner Z[n-k]
    int pu, pk2;
    if (k1 == 0) { pu = u; pk2 = (R2 - k2) % R2; }
    else { pu = (p.R1 - k1 == k1) ? u : 1 - u; pk2 = R2 - 1 - k2; }
    C a = v[q];
    C bq = sm[A::at(pk2, pu * W + w)];
    const bool upper = 2 * k > n;
    const i64 kk = upper ? n - k : k;
    if (upper) { const C tmp = a; a = bq; bq = tmp; }
    const C Va = mk<T>((a.x + bq.x) * hf, (a.y - bq.y) * hf);
    const C Vb = mk<T>((a.y + bq.y) * hf, (bq.x - a.x) * hf);
    C o;
    i64 row;
    if (p.kind == RK_DHT) {
      o = upper ? mk<T>((Va.x + Va.y) * p.f, (Vb.x + Vb.y) * p.f) : mk<T>((Va.x - Va.y) * p.f, (Vb.x - Vb.y) * p.f);
      row = k;
    } else {
      const C d = __ldg(p.dtw + kk);
      const C ua = cmul(Va, d), ub = cmul(Vb, d);
      const T fk = kk == 0 ? p.f0 : p.f;
      o = upper ? mk<T>(-ua.y * p.f, -ub.y * p.f) : mk<T>(ua.x * fk, ub.x * fk);
      row = p.kind == RK_DCT ? k : n - 1 - k;
    }
    dst[row * p.s] = o;
  }
}

// Column post-pass of the forward DCT-II / DST-II / DHT along a strided axis of length n where two adjacent real
// columns were transformed as one complex column Z (length-n complex FFT of the permuted rows):
//   Va[k] = (Z[k] + conj Z[n-k])/2, Vb[k] = (Z[k] - conj Z[n-k])/(2i);  DCT: C[k] = Re(d^k V[k]), C[n-k] = -Im(d^k V[k])
// z: [n][ld] complex (rows s apart), out: same shape; DST stores at the reversed row index.
template <typename T> struct ColPostParams {
  const cx<T>* z;
  cx<T>* out;
  i64 n, cols, s;       // rows, complex columns handled, row distance (complex units)
  int kind;
  T f0, f;
  const cx<T>* dtw;     // exp(-i pi k / (2n))
};

template <typename T> __global__ void k_r2r_colpost(const ColPostParams<T> p) {
  typedef cx<T> C;
  const i64 half = p.n / 2;
  const i64 total = (half + 1) * p.cols;
  const T hf = (T)0.5;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const i64 k = idx / p.cols, c = idx - k * p.cols;
    const i64 km = (p.n - k) % p.n;
    const C a = p.z[k * p.s + c];
    const C b = p.z[km * p.s + c];
    // Va = (a + conj b)/2 ; Vb = (a - conj b)/(2i)
    const C Va = mk<T>((a.x + b.x) * hf, (a.y - b.y) * hf);
    const C Vb = mk<T>((a.y + b.y) * hf, (b.x - a.x) * hf);
    C o1, o2;       // rows k and n-k of the result, (column 2c, column 2c+1)
    if (p.kind == RK_DHT) {
      o1 = mk<T>((Va.x - Va.y) * p.f, (Vb.x - Vb.y) * p.f);
      o2 = mk<T>((Va.x + Va.y) * p.f, (Vb.x + Vb.y) * p.f);
      p.out[k * p.s + c] = o1;
      if (km != k) p.out[km * p.s + c] = o2;
    } else {
      const C d = __ldg(p.dtw + k);
      const C ua = cmul(Va, d), ub = cmul(Vb, d);
      const T fk = k == 0 ? p.f0 : p.f;
      o1 = mk<T>(ua.x * fk, ub.x * fk);
      o2 = mk<T>(-ua.y * p.f, -ub.y * p.f);
      if (p.kind == RK_DCT) {
        p.out[k * p.s + c] = o1;
        if (k != 0 && km != k) p.out[km * p.s + c] = o2;
      } else {
        p.out[(p.n - 1 - k) * p.s + c] = o1;
        if (k != 0 && km != k) p.out[(k - 1) * p.s + c] = o2;
      }
    }
  }
}

}  // namespace jtb
