// Mixed-radix line FFT for lengths n = 2^a 3^b 5^c 7^d 11^e 13^f that fit one CTA: the native counterpart of the
// reference's FFTPACK path (cfftf / passf2..passf5 / passfg, fft/DoubleFFT_1D.java:6630-8009), which it selects
// for "smooth" non-power-of-two sizes (fft/DoubleFFT_1D.java:126-146).  Without this kernel those sizes go through
// Bluestein (correct, but 6-8x the work).
//
// Stockham autosort between two shared-memory buffers: stage s has radix R_s and stride Ns = prod_{i<s} R_i;
// butterfly b (0 <= b < n/R_s) reads in[b + q*n/R_s], multiplies by W_{Ns*R_s}^{q*(b mod Ns)}, applies DFT_{R_s} and
// writes out[(b - k)*R_s + k + q*Ns], k = b mod Ns.  W lines per CTA; strided lines put W adjacent lines in one CTA
// with the line index fastest (128-byte segments), exactly like the power-of-two kernels.
#pragma once
#include "jtb_common.cuh"

namespace jtb {

enum { MIX_MAX_STAGES = 12 };

template <typename T> struct MixedParams {
  const cx<T>* in;
  cx<T>* out;
  Geo gi, go;
  i64 nlines, line_base;
  int n, W, wfast;
  int nstages;
  int radix[MIX_MAX_STAGES];
  // division-free indexing: x / d == __umulhi(x, magic(d)) for every x this kernel forms (x * d < 2^32)
  unsigned m_nb[MIX_MAX_STAGES], m_ns[MIX_MAX_STAGES], m_n;
  int logW;            // W is a power of two when wfast
  int swap_in, swap_out, has_scale;
  T scale;
  const cx<T>* wtab;   // exp(-2 pi i j / n), j < n
  // two-pass transform of a long line N = N1*N2 (mixed_twopass_contig): the first pass multiplies output element i of
  // line l by the four-step twiddle W_N^(i * (l mod tw_mod)) = twA[m >> tw_logL] * twB[m & (2^tw_logL - 1)]; the second
  // pass reads contiguous lines but stores with the LINE index fastest (wfast_out: transposed, coalesced store)
  int tw_mode, tw_mod, tw_logL, wfast_out;
  const cx<T>* twA;
  const cx<T>* twB;
};

__host__ __device__ __forceinline__ unsigned mix_magic(unsigned d) { return d <= 1 ? 0u : (unsigned)(0x100000000ULL / d) + 1u; }
__device__ __forceinline__ int mix_div(int x, unsigned magic, int d) {
#ifdef JTB_EMU
  (void)magic; return x / d;
#else
  return d <= 1 ? x : (int)__umulhi((unsigned)x, magic);
#endif
}

template <typename C> __device__ __forceinline__ C mix_w(const C* __restrict__ w, int j) { return __ldg(w + j); }

// DFT of R values held in x[] (forward sign); roots: exp(-2 pi i m / R) = wtab[m * (n / R)]
template <typename T, int R>
__device__ __forceinline__ void dft_small(cx<T>* x, const cx<T>* __restrict__ wtab, int nr /* n / R */) {
  typedef cx<T> C;
  if (R == 2) {
    const C a = x[0], b = x[1];
    x[0] = cadd(a, b); x[1] = csub(a, b);
  } else if (R == 4) {
    const C t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
    const C t2 = cadd(x[1], x[3]), t3 = cmul_mi(csub(x[1], x[3]));
    x[0] = cadd(t0, t2); x[1] = cadd(t1, t3); x[2] = csub(t0, t2); x[3] = csub(t1, t3);
  } else if (R == 3) {
    const T s = (T)0.86602540378443864676372317075294L;   // sin(pi/3)
    const C t = cadd(x[1], x[2]);
    const C m = mk<T>(x[0].x - (T)0.5 * t.x, x[0].y - (T)0.5 * t.y);
    const C d = csub(x[1], x[2]);
    const C r = mk<T>(s * d.y, -s * d.x);                  // -i * s * d
    x[0] = cadd(x[0], t); x[1] = cadd(m, r); x[2] = csub(m, r);
  } else if (R == 5) {
    const T c1 = (T)0.30901699437494742410229341718282L, c2 = (T)-0.80901699437494742410229341718282L;
    const T s1 = (T)0.95105651629515357211643933337938L, s2 = (T)0.58778525229247312916870595463907L;
    const C a1 = cadd(x[1], x[4]), b1 = csub(x[1], x[4]);
    const C a2 = cadd(x[2], x[3]), b2 = csub(x[2], x[3]);
    const C m1 = mk<T>(x[0].x + c1 * a1.x + c2 * a2.x, x[0].y + c1 * a1.y + c2 * a2.y);
    const C m2 = mk<T>(x[0].x + c2 * a1.x + c1 * a2.x, x[0].y + c2 * a1.y + c1 * a2.y);
    // -i * (s1 b1 + s2 b2), -i * (s2 b1 - s1 b2)
    const C r1 = mk<T>(s1 * b1.y + s2 * b2.y, -(s1 * b1.x + s2 * b2.x));
    const C r2 = mk<T>(s2 * b1.y - s1 * b2.y, -(s2 * b1.x - s1 * b2.x));
    x[0] = mk<T>(x[0].x + a1.x + a2.x, x[0].y + a1.y + a2.y);
    x[1] = cadd(m1, r1); x[4] = csub(m1, r1);
    x[2] = cadd(m2, r2); x[3] = csub(m2, r2);
  } else {
    // generic odd radix (7, 11, 13): direct O(R^2) DFT with roots from the table
    C y[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      C acc = x[0];
#pragma unroll
      for (int q = 1; q < R; ++q) acc = cadd(acc, cmul(x[q], mix_w(wtab, ((q * k) % R) * nr)));
      y[k] = acc;
    }
#pragma unroll
    for (int k = 0; k < R; ++k) x[k] = y[k];
  }
}

template <typename T, int R>
__device__ __forceinline__ void mixed_stage(const cx<T>* __restrict__ src, cx<T>* __restrict__ dst, int n, int Ns, int lines,
                                            int ld, const cx<T>* __restrict__ wtab, int tid, int nthreads, int wfast,
                                            int W, int logW, unsigned m_nb, unsigned m_ns) {
  typedef cx<T> C;
  const int nb = n / R;                 // butterflies per line
  const int tstep = n / (Ns * R);       // twiddle index step: W_{Ns R}^{m} = wtab[m * tstep]
  // line-interleaved tiles decode (b, w) with the full tile width: a partial last tile runs its unused columns too
  for (int idx = tid; idx < nb * (wfast ? W : lines); idx += nthreads) {
    int b, w;
    if (wfast) { w = idx & (W - 1); b = idx >> logW; } else { w = mix_div(idx, m_nb, nb); b = idx - w * nb; }
    const int k = b - mix_div(b, m_ns, Ns) * Ns;
    C x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int i = b + q * nb;
      x[q] = src[wfast ? i * W + w : w * ld + i];
    }
    if (Ns > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) x[q] = cmul(x[q], mix_w(wtab, q * k * tstep));
    }
    dft_small<T, R>(x, wtab, nb);
    const int j0 = (b - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const int i = j0 + q * Ns;
      dst[wfast ? i * W + w : w * ld + i] = x[q];
    }
  }
}

// BIG: 1024-thread CTAs (64 registers) for plans whose radices are all <= 5 -- twice the warps to hide the shared-memory
// and twiddle latency of the stage loops; the generic odd radices (7, 11, 13) keep the 512-thread build.
template <typename T, bool BIG>
__global__ void __launch_bounds__(BIG ? 1024 : 512, 1) fft_mixed_kernel(const MixedParams<T> p) {
  typedef cx<T> C;
  JTB_DYN_SMEM(smem_raw);
  const int n = p.n, W = p.W, wfast = p.wfast;
  const int ld = n | 1;            // odd row length: column accesses of the [w][ld] layout are conflict-free
  C* bufA = reinterpret_cast<C*>(smem_raw);
  C* bufB = bufA + (size_t)W * ld;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const i64 line0 = p.line_base + (i64)blockIdx.x * W;
  const int lines = (p.nlines - line0 < W) ? (int)(p.nlines - line0) : W;
  // load (coalesced along the memory-contiguous direction)
  for (int idx = tid; idx < n * W; idx += nthreads) {
    int i, w;
    if (wfast) { w = idx & (W - 1); i = idx >> p.logW; } else { w = mix_div(idx, p.m_n, n); i = idx - w * n; }
    if (w < lines) {
      C z = p.in[geo_off(p.gi, line0 + w) + (i64)i * p.gi.stride];
      if (p.swap_in) z = cswap(z);
      bufA[wfast ? i * W + w : w * ld + i] = z;
    }
  }
  __syncthreads();
  C* src = bufA;
  C* dst = bufB;
  int Ns = 1;
  for (int s = 0; s < p.nstages; ++s) {
    const int R = p.radix[s];
    switch (R) {
      case 2: mixed_stage<T, 2>(src, dst, n, Ns, lines, ld, p.wtab, tid, nthreads, wfast, W, p.logW, p.m_nb[s], p.m_ns[s]); break;
      case 3: mixed_stage<T, 3>(src, dst, n, Ns, lines, ld, p.wtab, tid, nthreads, wfast, W, p.logW, p.m_nb[s], p.m_ns[s]); break;
      case 4: mixed_stage<T, 4>(src, dst, n, Ns, lines, ld, p.wtab, tid, nthreads, wfast, W, p.logW, p.m_nb[s], p.m_ns[s]); break;
      case 5: mixed_stage<T, 5>(src, dst, n, Ns, lines, ld, p.wtab, tid, nthreads, wfast, W, p.logW, p.m_nb[s], p.m_ns[s]); break;
      case 7: if (!BIG) mixed_stage<T, 7>(src, dst, n, Ns, lines, ld, p.wtab, tid, nthreads, wfast, W, p.logW, p.m_nb[s], p.m_ns[s]); break;
      case 11: if (!BIG) mixed_stage<T, 11>(src, dst, n, Ns, lines, ld, p.wtab, tid, nthreads, wfast, W, p.logW, p.m_nb[s], p.m_ns[s]); break;
      default: if (!BIG) mixed_stage<T, 13>(src, dst, n, Ns, lines, ld, p.wtab, tid, nthreads, wfast, W, p.logW, p.m_nb[s], p.m_ns[s]); break;
    }
    __syncthreads();
    Ns *= R;
    C* t = src; src = dst; dst = t;
  }
  const bool wf_out = wfast || p.wfast_out;
  // with the line index fastest (and blockDim a multiple of W) a thread always stores for the same line
  const int jw = p.tw_mode ? (int)((line0 + (tid & (W - 1))) % p.tw_mod) : 0;
  for (int idx = tid; idx < n * W; idx += nthreads) {
    int i, w;
    if (wf_out) { w = idx & (W - 1); i = idx >> p.logW; } else { w = mix_div(idx, p.m_n, n); i = idx - w * n; }
    if (w < lines) {
      C z = src[wfast ? i * W + w : w * ld + i];
      if (p.tw_mode) {
        const int m = i * jw;        // < N1*N2 <= 2^31 (host check)
        z = cmul(z, cmul(__ldg(p.twA + (m >> p.tw_logL)), __ldg(p.twB + (m & ((1 << p.tw_logL) - 1)))));
      }
      if (p.has_scale) { z.x *= p.scale; z.y *= p.scale; }
      if (p.swap_out) z = cswap(z);
      p.out[geo_off(p.go, line0 + w) + (i64)i * p.go.stride] = z;
    }
  }
}

}  // namespace jtb
