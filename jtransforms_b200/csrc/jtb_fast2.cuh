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

enum { FM_PLAIN = 0, FM_TWID = 1, FM_TRANSPOSE = 2, FM_RFFT = 3 };

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
};

template <typename T> __device__ __forceinline__ cx<T> fs_tw2(const Fast2Params<T>& p, int m) {
  return cmul(__ldg(p.fsA + (m >> p.fs_logL)), __ldg(p.fsB + (m & ((1 << p.fs_logL) - 1))));
}

template <typename T, int LOGN, int LOGE, bool SIN, int MODE, int W>
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
  for (int i = tid; i < FastTw<S>::COUNT; i += W * S::TPL) twt[i] = __ldg(p.twg + i);

  const i64 line0 = (i64)blockIdx.x * W;
  const i64 line = line0 + w;
  const bool valid = line < p.nlines;
  const i64 g = line / p.c0;
  const int c = (int)(line - g * p.c0);
  const i64 g_hi = g / p.gmod;
  const int g_lo = (int)(g - g_hi * p.gmod);
  C v[S::E];
  if (valid) {
    const C* src = p.in + g_lo * p.in_gdist + g_hi * p.in_gdist2 + c * p.in_cdist;
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = src[(t + q * S::TPL) * p.in_stride];
  } else {
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = mk<T>(0, 0);
  }
  if (p.swap_in) {
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
  }
  FastLoop<T, S, 0, SIN, W>::run(v, sm, twt, t, w);
  if (p.has_scale) {
#pragma unroll
    for (int q = 0; q < S::E; ++q) { v[q].x *= p.scale; v[q].y *= p.scale; }
  }

  if (MODE == FM_PLAIN || MODE == FM_TWID) {
    if (MODE == FM_TWID) {
      const int idx = p.tw_src ? g_lo : c;
      C tw = fs_tw2(p, idx * t);
      const C ws = fs_tw2(p, idx * S::TPL);
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
    if (valid) {
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
    if (S::S > 1) __syncthreads();
#pragma unroll
    for (int q = 0; q < S::E; ++q) sm[A::at(t + q * S::TPL, w)] = v[q];
    __syncthreads();
    if (valid) {
      C* dst = p.out + g_lo * p.out_gdist + g_hi * p.out_gdist2 + c * p.out_cdist;
      const T hf = (T)0.5;
#pragma unroll
      for (int q = 0; q < S::E / 2; ++q) {
        const int k = t + q * S::TPL;   // 0 .. N/2 - 1
        if (k == 0) {
          const C z = sm[A::at(0, w)];
          dst[0] = mk<T>(z.x + z.y, z.x - z.y);
          const C zm = sm[A::at(S::N / 2, w)];
          dst[S::N / 2] = mk<T>(zm.x, -zm.y);
        } else {
          const C a = sm[A::at(k, w)];
          const C b = sm[A::at(S::N - k, w)];
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

}  // namespace jtb
