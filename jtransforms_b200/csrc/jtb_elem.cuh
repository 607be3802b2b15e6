// Element-wise kernels around the FFT core: generic real<->complex line staging for
// DCT/DST/DHT and non-power-of-two real FFTs, the packed-real untangle passes of the
// 2-D/3-D real transforms, the DHT yTransform and the synthetic input fill.
// All are plain grid-stride HBM-bound kernels (coalesced along the line index that is
// contiguous in memory).
#pragma once
#include "jtb_common.cuh"

namespace jtb {

enum PreMode { PRE_R2C = 0, PRE_UNPACK_HERM = 1, PRE_DCT2 = 2, PRE_DCT3 = 3 };
enum PostMode { POST_PACK = 0, POST_REAL = 1, POST_DCT2 = 2, POST_DCT3 = 3, POST_DHT = 4 };

template <typename T> struct R2RParams {
  T* a;                 // real data (lines described by g, units: real elements)
  cx<T>* work;          // complex staging, line l at work[(l - line_base) * n]
  Geo g;
  i64 line_base, nlines;   // lines [line_base, nlines)
  i64 n;
  int mode;
  int dst;              // DST flavour of the DCT modes (sign alternation / reversal)
  T f0, f;              // factors for index 0 / the others
  const cx<T>* dtw;     // exp(-i pi k / (2n))
};

// real line -> complex staging line ------------------------------------------------------
template <typename T> __global__ void k_r2r_pre(const R2RParams<T> p) {
  const i64 n = p.n;
  const i64 total = (p.nlines - p.line_base) * n;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const i64 lrel = idx / n, q = idx - lrel * n;
    const T* src = p.a + geo_off(p.g, p.line_base + lrel);
    const i64 st = p.g.stride;
    cx<T> z = mk<T>(0, 0);
    switch (p.mode) {
      case PRE_R2C: z.x = src[q * st]; break;
      case PRE_UNPACK_HERM: {
        // packed half spectrum (fft/DoubleFFT_1D.java:436-450) -> full Hermitian spectrum
        const bool upper = 2 * q > n;
        const i64 k = upper ? n - q : q;
        T re, im;
        if (k == 0) { re = src[0]; im = 0; }
        else if ((n & 1) == 0) {
          if (2 * k == n) { re = src[st]; im = 0; }
          else { re = src[2 * k * st]; im = src[(2 * k + 1) * st]; }
        } else {
          re = src[2 * k * st];
          im = (2 * k == n - 1) ? src[st] : src[(2 * k + 1) * st];
        }
        z.x = re; z.y = upper ? -im : im;
      } break;
      case PRE_DCT2: {
        // Makhoul permutation: v[q] = x[2q] (front half), v[n-1-m] = x[2m+1]
        const i64 j = (q < (n + 1) / 2) ? 2 * q : 2 * (n - 1 - q) + 1;
        T x = src[j * st];
        if (p.dst && (j & 1)) x = -x;
        z.x = x;
      } break;
      case PRE_DCT3: {
        // b[q] = g_q * a[q] * exp(+i pi q / 2n)
        const i64 j = p.dst ? n - 1 - q : q;
        const T x = src[j * st] * (q == 0 ? p.f0 : p.f);
        const cx<T> w = __ldg(p.dtw + q);
        z.x = x * w.x; z.y = -x * w.y;
      } break;
    }
    p.work[idx] = z;
  }
}

// complex staging line -> real line -------------------------------------------------------
template <typename T> __global__ void k_r2r_post(const R2RParams<T> p) {
  const i64 n = p.n;
  const i64 total = (p.nlines - p.line_base) * n;
  for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
    const i64 lrel = idx / n, i = idx - lrel * n;
    T* dst = p.a + geo_off(p.g, p.line_base + lrel);
    const cx<T>* wl = p.work + lrel * n;
    const i64 st = p.g.stride;
    T out = 0;
    switch (p.mode) {
      case POST_PACK: {
        if (i == 0) out = wl[0].x;
        else if ((n & 1) == 0) {
          if (i == 1) out = wl[n / 2].x;
          else out = (i & 1) ? wl[(i - 1) / 2].y : wl[i / 2].x;
        } else {
          if (i == 1) out = wl[(n - 1) / 2].y;
          else out = (i & 1) ? wl[(i - 1) / 2].y : wl[i / 2].x;
        }
        out *= p.f;
      } break;
      case POST_REAL: out = wl[i].x * p.f; break;
      case POST_DCT2: {
        const i64 k = p.dst ? n - 1 - i : i;
        const cx<T> w = __ldg(p.dtw + k);
        const cx<T> v = wl[k];
        out = (w.x * v.x - w.y * v.y) * (k == 0 ? p.f0 : p.f);
      } break;
      case POST_DCT3: {
        const i64 m = (i & 1) ? n - 1 - (i - 1) / 2 : i / 2;
        out = wl[m].x;
        if (p.dst && (i & 1)) out = -out;
      } break;
      case POST_DHT: out = (wl[i].x - wl[i].y) * p.f; break;
    }
    dst[i * st] = out;
  }
}

// rdft2d_sub / rdft3d_sub: untangle the (k3 = 0, k3 = C/2) pseudo column ---------------------
// After the complex passes over the other axes, slot (.., 0..1) of every row holds
// G = F0 + i*Fh (F0, Fh: spectra at column 0 and column C/2).  Forward (dir=+1) rewrites the
// pair (P, Q = mirror of P over all leading axes) as
//   P <- (G[P] + conj G[Q]) / 2 = F0[P]        Q <- (-Im, Re) of ... = packed Fh  (see doc table,
// fft/DoubleFFT_2D.java:794-810, :2544-2574; fft/DoubleFFT_3D.java:1298-1328, :6909-7021).
// dir=-1 is the exact inverse (without the 1/2).
template <typename T> __device__ __forceinline__ void untangle_pair(T* pi, T* pj, int dir) {
  const T i0 = pi[0], i1 = pi[1], j0 = pj[0], j1 = pj[1];
  if (dir > 0) {
    const T h = (T)0.5;
    const T nj0 = h * (i0 - j0), nj1 = h * (i1 + j1);
    pj[0] = nj0; pi[0] = i0 - nj0;
    pj[1] = nj1; pi[1] = i1 - nj1;
  } else {
    pi[0] = i0 + j0; pj[0] = i0 - j0;
    pi[1] = i1 + j1; pj[1] = j1 - i1;
  }
}

// 2-D: rows i in [1, R/2) pair with R - i.   a: R x C reals.
template <typename T> __global__ void k_untangle2d(T* a, i64 R, i64 C, int dir) {
  const i64 cnt = R / 2 - 1;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += (i64)gridDim.x * blockDim.x) {
    const i64 i = t + 1, j = R - i;
    untangle_pair(a + i * C, a + j * C, dir);
  }
}

// 3-D: a: S x R x C reals.  Pairs follow fft/DoubleFFT_3D.java:6909-7021:
//   i in [1,S/2): (i,0)<->(S-i,0), (i,R/2)<->(S-i,R/2), (i,k)<->(S-i,R-k), (S-i,k)<->(i,R-k)  for k in [1,R/2)
//   slices 0 and S/2: (s,k)<->(s,R-k) for k in [1,R/2)
// In every pair the first member keeps F0 and the second receives the packed Fh.
template <typename T> __global__ void k_untangle3d(T* a, i64 S, i64 R, i64 C, int dir) {
  // enumerate (s, k) over s in [0,S), k in [0, R/2]; each (s,k) handles pair (s,k) <-> ((S-s)%S, (R-k)%R)
  // restricted so that every unordered pair is visited once with the reference's orientation.
  const i64 RH = R / 2, SH = S / 2;
  const i64 total = S * (RH + 1);
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
    const i64 s = t / (RH + 1), k = t - s * (RH + 1);
    const i64 ms = (S - s) % S, mk = (R - k) % R;
    bool first;   // is (s,k) the member that keeps F0 ?
    if (k == 0 || k == RH) {
      // self-mirrored in k: pair (s,k)<->(S-s,k), first member is s in [1, S/2)
      if (s == 0 || s == SH) continue;
      first = s < SH;
    } else {
      // 0 < k < R/2 always is the first member: (i,k)->(S-i,R-k), (S-i,k)->(i,R-k), (0,k)->(0,R-k), (S/2,k)->(S/2,R-k)
      first = true;
    }
    if (!first) continue;
    untangle_pair(a + (s * R + k) * C, a + (ms * R + mk) * C, dir);
  }
}

// DHT yTransform (dht/DoubleDHT_2D.java:1288-1309): separable cas*cas -> true 2-D DHT
template <typename T> __global__ void k_ytransform2d(T* a, i64 R, i64 C) {
  const i64 RH = R / 2 + 1, CH = C / 2 + 1;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < RH * CH; t += (i64)gridDim.x * blockDim.x) {
    const i64 r = t / CH, c = t - r * CH;
    const i64 mr = (R - r) % R, mc = (C - c) % C;
    const T A = a[r * C + c], B = a[mr * C + c], Cc = a[r * C + mc], D = a[mr * C + mc];
    const T E = ((A + D) - (B + Cc)) * (T)0.5;
    a[r * C + c] = A - E;
    a[mr * C + c] = B + E;
    a[r * C + mc] = Cc + E;
    a[mr * C + mc] = D - E;
  }
}

// 3-D yTransform (dht/DoubleDHT_3D.java:2314-2356)
template <typename T> __global__ void k_ytransform3d(T* a, i64 S, i64 R, i64 C) {
  const i64 SH = S / 2 + 1, RH = R / 2 + 1, CH = C / 2 + 1;
  for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < SH * RH * CH; t += (i64)gridDim.x * blockDim.x) {
    const i64 s = t / (RH * CH), rem = t - s * RH * CH;
    const i64 r = rem / CH, c = rem - r * CH;
    const i64 sC = (S - s) % S, rC = (R - r) % R, cC = (C - c) % C;
    const i64 x1 = (s * R + rC) * C + c, x2 = (s * R + r) * C + cC, x3 = (sC * R + r) * C + c, x4 = (sC * R + rC) * C + cC;
    const i64 x5 = (sC * R + rC) * C + c, x6 = (sC * R + r) * C + cC, x7 = (s * R + r) * C + c, x8 = (s * R + rC) * C + cC;
    const T A = a[x1], B = a[x2], Cv = a[x3], D = a[x4], E = a[x5], F = a[x6], G = a[x7], H = a[x8];
    const T h = (T)0.5;
    a[x7] = (A + B + Cv - D) * h;
    a[x3] = (E + F + G - H) * h;
    a[x1] = (G + H + E - F) * h;
    a[x5] = (Cv + D + A - B) * h;
    a[x2] = (H + G + F - E) * h;
    a[x6] = (D + Cv + B - A) * h;
    a[x8] = (B + A + D - Cv) * h;
    a[x4] = (F + E + H - G) * h;
  }
}


// ---------------------------------------------------------------------------------------------------
// Packed real spectrum -> full Hermitian spectrum (the role of fillSymmetric, fft/DoubleFFT_2D.java:3877-3992,
// fft/DoubleFFT_3D.java:7387-7614, after realForward): every thread produces one complex element of the full array
// from the packed layout (fft/DoubleFFT_2D.java:794-810, fft/DoubleFFT_3D.java:1298-1328) held in `pk`.
// conj_out: realInverseFull of real data = conj of the forward spectrum (times `f`).
template <typename T> __device__ __forceinline__ cx<T> half2d(const T* pk, i64 R, i64 C, i64 r, i64 k) {
  // H(r, k), 0 <= k <= C/2
  const i64 h = C / 2;
  if (k > 0 && k < h) return mk<T>(pk[r * C + 2 * k], pk[r * C + 2 * k + 1]);
  const i64 rm = (R - r) % R;
  if (r == 0 || 2 * r == R) return mk<T>(pk[r * C + (k == 0 ? 0 : 1)], (T)0);
  if (k == 0) return 2 * r < R ? mk<T>(pk[r * C], pk[r * C + 1]) : mk<T>(pk[rm * C], -pk[rm * C + 1]);
  // k == C/2: a[(R-r) C + 1] = Re, a[(R-r) C] = -Im for 0 < r < R/2; conjugate mirror above R/2
  return 2 * r < R ? mk<T>(pk[rm * C + 1], -pk[rm * C]) : mk<T>(pk[r * C + 1], pk[r * C]);
}

template <typename T> __global__ void k_expand_full_1d(const T* pk, cx<T>* out, i64 n, int conj_out, T f) {
  for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (i64)gridDim.x * blockDim.x) {
    const i64 kk = 2 * k > n ? n - k : k;
    cx<T> z;
    if (kk == 0) z = mk<T>(pk[0], (T)0);
    else if (2 * kk == n) z = mk<T>(pk[1], (T)0);
    else z = mk<T>(pk[2 * kk], pk[2 * kk + 1]);
    if ((2 * k > n) != (conj_out != 0)) z.y = -z.y;
    out[k] = mk<T>(z.x * f, z.y * f);
  }
}

template <typename T> __global__ void k_expand_full_2d(const T* pk, cx<T>* out, i64 R, i64 C, int conj_out, T f) {
  const i64 total = R * C, h = C / 2;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
    const i64 r = i / C, c = i - r * C;
    cx<T> z;
    bool cj = conj_out != 0;
    if (c <= h) z = half2d<T>(pk, R, C, r, c);
    else { z = half2d<T>(pk, R, C, (R - r) % R, C - c); cj = !cj; }
    if (cj) z.y = -z.y;
    out[i] = mk<T>(z.x * f, z.y * f);
  }
}

template <typename T> __device__ __forceinline__ cx<T> half3d(const T* pk, i64 S, i64 R, i64 C, i64 k1, i64 k2, i64 k3) {
  // H(k1, k2, k3), 0 <= k3 <= C/2
  const i64 h = C / 2;
  auto at = [&](i64 a, i64 b, i64 c) -> T { return pk[(a * R + b) * C + c]; };
  if (k3 > 0 && k3 < h) return mk<T>(at(k1, k2, 2 * k3), at(k1, k2, 2 * k3 + 1));
  const i64 m1 = (S - k1) % S, m2 = (R - k2) % R;
  const bool k2self = (k2 == 0 || 2 * k2 == R), k1self = (k1 == 0 || 2 * k1 == S);
  if (k3 == 0) {
    if (!k2self) return 2 * k2 < R ? mk<T>(at(k1, k2, 0), at(k1, k2, 1)) : mk<T>(at(m1, m2, 0), -at(m1, m2, 1));
    if (k1self) return mk<T>(at(k1, k2, 0), (T)0);
    return 2 * k1 < S ? mk<T>(at(k1, k2, 0), at(k1, k2, 1)) : mk<T>(at(m1, k2, 0), -at(m1, k2, 1));
  }
  // k3 == C/2
  if (!k2self) {
    // 0 < k2 < R/2: (P[(S-k1)%S][R-k2][1], -P[..][0]); above R/2 the conjugate of the mirror element
    if (2 * k2 < R) return mk<T>(at(m1, R - k2, 1), -at(m1, R - k2, 0));
    return mk<T>(at(k1, k2, 1), at(k1, k2, 0));     // conj H(m1, m2, h) = conj(P[k1][k2][1], -P[k1][k2][0])
  }
  if (k1self) return mk<T>(at(k1, k2, 1), (T)0);
  return 2 * k1 < S ? mk<T>(at(S - k1, k2, 1), -at(S - k1, k2, 0)) : mk<T>(at(k1, k2, 1), at(k1, k2, 0));
}

template <typename T> __global__ void k_expand_full_3d(const T* pk, cx<T>* out, i64 S, i64 R, i64 C, int conj_out, T f) {
  const i64 total = S * R * C, h = C / 2;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
    const i64 k1 = i / (R * C), rem = i - k1 * R * C;
    const i64 k2 = rem / C, k3 = rem - k2 * C;
    cx<T> z;
    bool cj = conj_out != 0;
    if (k3 <= h) z = half3d<T>(pk, S, R, C, k1, k2, k3);
    else { z = half3d<T>(pk, S, R, C, (S - k1) % S, (R - k2) % R, C - k3); cj = !cj; }
    if (cj) z.y = -z.y;
    out[i] = mk<T>(z.x * f, z.y * f);
  }
}


// ---------------------------------------------------------------------------------------------------
// Helpers of the three-pass transform for lines beyond the two-pass limit (jtb_engine_impl.cuh, c2c_big_contig):
// a[k1*N2 + n2] *= W_n^(+-k1*n2) with W_n^m = A[m >> logL] * B[m & (L-1)]  (the twiddle tables of fs_tables)
template <typename C>
__global__ void k_big_twiddle(C* a, i64 N1, int logN2, const C* A, const C* B, int logL, int conj_tw) {
  const i64 total = N1 << logN2, N2m = (1LL << logN2) - 1, Lm = (1LL << logL) - 1;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
    const i64 k1 = i >> logN2, n2 = i & N2m;
    const i64 m = k1 * n2;
    const C w = cmul(__ldg(A + (m >> logL)), __ldg(B + (m & Lm)));
    a[i] = conj_tw ? cmulc(a[i], w) : cmul(a[i], w);
  }
}

// out[c*R + r] = in[r*Cn + c] through a 32 x 32 shared-memory tile (blockDim = 32 x 8); R, Cn multiples of 32
template <typename C> __global__ void k_transpose32(const C* in, C* out, i64 R, i64 Cn) {
  JTB_DYN_SMEM(smem_raw);                      // 32 x 33 elements
  C (*tile)[33] = reinterpret_cast<C (*)[33]>(smem_raw);
  const i64 tiles_c = Cn / 32, ntiles = (R / 32) * tiles_c;
  for (i64 tb = blockIdx.x; tb < ntiles; tb += gridDim.x) {
    const i64 tr = tb / tiles_c, tc = tb - tr * tiles_c;
    for (int j = threadIdx.y; j < 32; j += 8) tile[j][threadIdx.x] = in[(tr * 32 + j) * Cn + tc * 32 + threadIdx.x];
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += 8) out[(tc * 32 + j) * R + tr * 32 + threadIdx.x] = tile[threadIdx.x][j];
    __syncthreads();
  }
}

// counter-based uniform fill: u(i) = (mix64((seed + i) * gamma) >> 11) * 2^-53 (oracle: fill_uniform)
template <typename T> __global__ void k_fill_uniform(T* a, i64 count, unsigned long long seed, T lo, T hi) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x) {
    unsigned long long z = ((unsigned long long)i + seed) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    a[i] = (T)((double)lo + ((double)hi - (double)lo) * u);
  }
}

template <typename T> __global__ void k_scale(T* a, i64 count, T s) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (i64)gridDim.x * blockDim.x) a[i] *= s;
}

__global__ void k_cast_c64_c32(const double2* in, float2* out, i64 count);

}  // namespace jtb
