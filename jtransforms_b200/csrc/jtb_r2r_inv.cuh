// Fused inverse DCT (DCT-III) / DST (DST-III) kernels: the forward kernels of jtb_fast2.cuh run backwards.
//
//   x[i] = v[m(i)],  m(i) = i/2 (i even), n-1-(i-1)/2 (i odd),   v = Re IDFT_n(b),  b[q] = g_q a[q] e^{+i pi q/2n}
// (dct/DoubleDCT_1D.java:361-434: pre-scale, dctsub, cftfsub, rftfsub and the final butterfly loop; DST-III =
// reversed input, DCT-III, odd outputs negated, dst/DoubleDST_1D.java:264-325).  With V[q] = (b[q] + conj b[n-q])/2
// the real sequence v is the inverse real FFT of V; it is computed with ONE complex FFT
//   * rows (contiguous lines): of length N = n/2 on Z[k] = Ze[k] + i Zo[k], Ze = V[k] + conj V[N-k],
//     Zo = (V[k] - conj V[N-k]) e^{+2 pi i k/n}, giving z[j] = v[2j] + i v[2j+1];
//   * columns (strided lines): of length n on Z[q] = V1[q] + i V2[q] for TWO adjacent real columns, giving v1 + i v2.
#pragma once
#include "jtb_fast2.cuh"

namespace jtb {

// ---------------------------------------------------------------------------------------------------
// contiguous real lines of n = 2N reals, in place; KIND = RK_DCT or RK_DST.
// Shared-memory traffic besides the FFT's own exchanges is two half-line hand-overs: Z[N-k] to the thread that owns
// the upper half before the transform, z[N-1-m] to the thread that stores x[4m..4m+3] after it.
template <typename T, int LOGN, int LOGE, int KIND, int W>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, (Sched<LOGN, LOGE>::E <= 8 ? FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB
                                                                                      : RowOcc16<W * Sched<LOGN, LOGE>::TPL>::V))
fft_r2r_row_inv_kernel(const RowR2RParams<T> p) {
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
  const bool valid = line0 + w < p.nlines;
  T* xl = p.a + (valid ? (line0 + w) * p.dist : 0);
  C* half = sm + w * N;          // unpadded line buffer for the half-line hand-overs
  // a'[q] of the (DST: reversed) line, coalesced 8-byte loads straight from global memory
  auto ld = [&](int q) -> T { return valid ? xl[KIND == RK_DST ? n - 1 - q : q] : (T)0; };
  T xk[H], xnk[H], xmk[H], xpk[H];
#pragma unroll
  for (int q = 0; q < H; ++q) {
    const int k = t + q * S::TPL;
    if (k == 0) { xk[q] = ld(0); xnk[q] = ld(N); xmk[q] = ld(N / 2); xpk[q] = ld(n - N / 2); }
    else { xk[q] = ld(k); xnk[q] = ld(n - k); xmk[q] = ld(N - k); xpk[q] = ld(N + k); }
  }
  C v[S::E];
#pragma unroll
  for (int q = 0; q < H; ++q) {
    const int k = t + q * S::TPL;
    C zk, zm;
    if (k == 0) {
      const T a0 = xk[q] * p.f0;
      const T aN = xnk[q] * p.f * (T)0.70710678118654752440084436210485L;
      zk = mk<T>(a0 + aN, a0 - aN);                           // Z[0]: V[0] = b[0], V[N] = Re b[N]
      // Z[N/2]: V[N/2] = (b[N/2] + conj b[n-N/2])/2, Ze = 2 Re V, Zo = -2 Im V   (b[q] = x_q conj(dtw[q]))
      const C dh = __ldg(p.dtw + N / 2), dg = __ldg(p.dtw + (n - N / 2));
      const T x1 = xmk[q] * p.f, x2 = xpk[q] * p.f;
      zm = mk<T>(x1 * dh.x + x2 * dg.x, x1 * dh.y - x2 * dg.y);
    } else {
      const T x0 = xk[q] * p.f, x1 = xnk[q] * p.f, x2 = xmk[q] * p.f, x3 = xpk[q] * p.f;
      // one table read; D(n-k) = -i conj D, D(N-k) = c conj D, D(N+k) = c D, exp(-2 pi i k/n) = D^4
      const C dk = __ldg(p.dtw + k);
      const C dnk = mk<T>(-dk.y, -dk.x), dmk = dct_tw_nmk(dk), dpk = dct_tw_npk(dk);
      // V[k] = (b[k] + conj b[n-k])/2, V[N-k] = (b[N-k] + conj b[N+k])/2
      const C Vk = mk<T>((T)0.5 * (x0 * dk.x + x1 * dnk.x), (T)0.5 * (-x0 * dk.y + x1 * dnk.y));
      const C Vm = mk<T>((T)0.5 * (x2 * dmk.x + x3 * dpk.x), (T)0.5 * (-x2 * dmk.y + x3 * dpk.y));
      const C Ze = mk<T>(Vk.x + Vm.x, Vk.y - Vm.y);             // V[k] + conj V[N-k]
      const C df = mk<T>(Vk.x - Vm.x, Vk.y + Vm.y);             // V[k] - conj V[N-k]
      const C Zo = cmulc(df, dct_tw_pow4(dk));                  // * e^{+2 pi i k/n}
      zk = mk<T>(Ze.x - Zo.y, Ze.y + Zo.x);                     // Z[k]   = Ze + i Zo
      zm = mk<T>(Ze.x + Zo.y, -Ze.y + Zo.x);                    // Z[N-k] = conj Ze + i conj Zo
    }
    // inverse transform with the forward kernel: IDFT(Z) = swap(DFT(swap Z)); Z[k] stays in this thread
    v[q] = cswap(zk);
    half[k == 0 ? N / 2 : N - k] = cswap(zm);
  }
  __syncthreads();
#pragma unroll
  for (int q = H; q < S::E; ++q) v[q] = half[t + q * S::TPL];
  __syncthreads();
  FastLoop<T, S, 0, false, W>::run(v, sm, twt, t, w, p.twg);
  // z[j] = (v[2j], v[2j+1]), j = t + q*TPL.  x[4m..4m+3] = (z[m].x, z[N-1-m].y, z[m].y, z[N-1-m].x): the upper half
  // is handed to the threads that own m < N/2, which store 32 contiguous bytes each (DST: odd outputs negated)
  if (S::S > 1) __syncthreads();
#pragma unroll
  for (int q = H; q < S::E; ++q) half[t + q * S::TPL] = cswap(v[q]);
  __syncthreads();
  if (!valid) return;
#pragma unroll
  for (int q = 0; q < H; ++q) {
    const int m = t + q * S::TPL;
    const C z = cswap(v[q]);
    const C zb = half[N - 1 - m];
    const T sg = (KIND == RK_DST) ? (T)-1 : (T)1;
    st4(xl + 4 * m, z.x, sg * zb.y, z.y, sg * zb.x);
  }
}

// ---------------------------------------------------------------------------------------------------
// First pass of the strided inverse DCT/DST over two-pass lengths n = R1*R2: a CTA holds the two lines k1 and R1-k1
// (rows q = k1 + R1*q2) of W adjacent complex columns (= 2W real columns), builds
//   Z[q] = g c_q/2 ((a1[q] + a2[n-q]) + i (a2[q] - a1[n-q])),  c_q = e^{+i pi q/2n}     (Z[0] = f0 (a1[0] + i a2[0]))
// from the row pair (q, n-q) that meets in the CTA, transforms over q2 (length R2, inverse via the swapped domain),
// multiplies by the four-step twiddle and stores row k1*R2 + m2 of the work array (still swapped).
template <typename T, int LOGN, int LOGE, int W>
__global__ void __launch_bounds__(2 * W * Sched<LOGN, LOGE>::TPL, 2) fft_colpair_inv_kernel(const ColPairParams<T> p, const cx<T>* fsA,
                                                                                            const cx<T>* fsB, int fs_logL) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  constexpr int R2 = S::N, W2 = 2 * W;
  typedef FastAddr<T, S, true, W2> A;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  const int tid = threadIdx.x;
  const int wu = tid % W2, t = tid / W2;
  const int u = wu / W, w = wu - u * W;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W2 * S::TPL) twt[i] = __ldg(p.twg + i);
  const int groups = p.cols / W, pairs = p.R1 / 2 + 1;
  int b = blockIdx.x;
  const int cg = b % groups; b /= groups;
  const int pr = b % pairs;
  const int batch = b / pairs;
  const int k1 = u == 0 ? pr : (p.R1 - pr) % p.R1;
  const i64 n = (i64)p.R1 * R2;
  const C* src = p.z + batch * p.bdist + cg * W + w;
  C v[S::E];
#pragma unroll
  for (int q = 0; q < S::E; ++q) {
    const i64 row = k1 + (i64)p.R1 * (t + q * S::TPL);
    v[q] = src[(p.kind == RK_DST ? n - 1 - row : row) * p.s];
  }
#pragma unroll
  for (int q = 0; q < S::E; ++q) sm[A::at(t + q * S::TPL, wu)] = v[q];
  __syncthreads();
  const T hf = (T)0.5;
#pragma unroll
  for (int q = 0; q < S::E; ++q) {
    const int q2 = t + q * S::TPL;
    const i64 row = k1 + (i64)p.R1 * q2;
    int pu, pq2;
    if (k1 == 0) { pu = u; pq2 = (R2 - q2) % R2; }
    else { pu = (p.R1 - k1 == k1) ? u : 1 - u; pq2 = R2 - 1 - q2; }
    const C a = v[q];
    const C bq = sm[A::at(pq2, pu * W + w)];
    C z;
    if (row == 0) z = mk<T>(a.x * p.f0, a.y * p.f0);
    else {
      const C d = __ldg(p.dtw + row);                           // e^{-i pi q/2n}
      const C s2 = mk<T>((a.x + bq.y) * hf * p.f, (a.y - bq.x) * hf * p.f);
      z = cmulc(s2, d);
    }
    v[q] = cswap(z);
  }
  __syncthreads();
  FastLoop<T, S, 0, true, W2>::run(v, sm, twt, t, wu, p.twg);
  const bool dup = (u == 1) && (k1 == pr);
  if (dup) return;
  {
    const int L = (1 << fs_logL) - 1;
    const int m0 = k1 * t, ms = k1 * S::TPL;
    C tw = cmul(__ldg(fsA + (m0 >> fs_logL)), __ldg(fsB + (m0 & L)));
    const C ws = cmul(__ldg(fsA + (ms >> fs_logL)), __ldg(fsB + (ms & L)));
    C* dst = p.out + batch * p.bdist + (i64)k1 * R2 * p.s + cg * W + w;
#pragma unroll
    for (int q = 0; q < S::E; ++q) {
      dst[(i64)(t + q * S::TPL) * p.s] = cmul(v[q], tw);
      if (q + 1 < S::E) tw = cmul(tw, ws);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Single-pass forward / inverse DCT, DST (and forward DHT) along the strided axis for lengths n = 2^LOGN whose
// W adjacent complex columns (= 2W real columns, rows s complex elements apart) fit one CTA: n <= 1024.
// Forward: Makhoul row permutation at the load, n-point FFT, pair post-pass (row k meets row n-k in shared memory).
// Inverse: pair pre-pass, inverse FFT (swapped domain), un-permuting store.  Replaces the 4-column gather loops of
// ddxt2d_subth (dct/DoubleDCT_2D.java:625-957) for the image sizes between the row kernels and the two-pass path.
template <typename T> struct ColR2RParams {
  cx<T>* a;            // [batches][n][s] complex (= [n][2s] reals), transformed in place
  i64 s, bdist;        // row distance, batch distance (complex units)
  int cols, batches;   // complex columns, arrays
  int kind;
  T f0, f;
  const cx<T>* twg;
  const cx<T>* dtw;    // exp(-i pi k / (2n)), k < n
};

template <typename T, int LOGN, int LOGE, int W, bool INV>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, (Sched<LOGN, LOGE>::E <= 8 ? FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB
                                                                                      : (FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB + 1) / 2))
fft_col_r2r_kernel(const ColR2RParams<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, true, W> A;
  constexpr int n = S::N;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  const int tid = threadIdx.x;
  const int w = tid % W, t = tid / W;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W * S::TPL) twt[i] = __ldg(p.twg + i);
  const int groups = p.cols / W;
  const int batch = blockIdx.x / groups;
  const int cg = blockIdx.x - batch * groups;
  C* base = p.a + (i64)batch * p.bdist + cg * W + w;
  const T hf = (T)0.5;
  C v[S::E];
  if (!INV) {
#pragma unroll
    for (int q = 0; q < S::E; ++q) {
      const int j = t + q * S::TPL;
      int row = j;
      bool second = false;
      if (p.kind != RK_DHT) { second = q >= S::E / 2; row = second ? 2 * (n - 1 - j) + 1 : 2 * j; }   // 2j >= n <=> q >= E/2
      C z = base[(i64)row * p.s];
      if (p.kind == RK_DST && second) { z.x = -z.x; z.y = -z.y; }
      v[q] = z;
    }
    FastLoop<T, S, 0, true, W>::run(v, sm, twt, t, w, p.twg);
    if (S::S > 1) __syncthreads();
#pragma unroll
    for (int q = 0; q < S::E; ++q) sm[A::at(t + q * S::TPL, w)] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < S::E; ++q) {
      const int k = t + q * S::TPL;
      C a = v[q];
      C bq = sm[A::at((n - k) & (n - 1), w)];
      const bool upper = 2 * k > n;
      const int kk = upper ? n - k : k;
      if (upper) { const C tmp = a; a = bq; bq = tmp; }
      const C Va = mk<T>((a.x + bq.x) * hf, (a.y - bq.y) * hf);
      const C Vb = mk<T>((a.y + bq.y) * hf, (bq.x - a.x) * hf);
      C o;
      int row = k;
      if (p.kind == RK_DHT) {
        o = upper ? mk<T>((Va.x + Va.y) * p.f, (Vb.x + Vb.y) * p.f) : mk<T>((Va.x - Va.y) * p.f, (Vb.x - Vb.y) * p.f);
      } else {
        const C d = __ldg(p.dtw + kk);
        const C ua = cmul(Va, d), ub = cmul(Vb, d);
        const T fk = kk == 0 ? p.f0 : p.f;
        o = upper ? mk<T>(-ua.y * p.f, -ub.y * p.f) : mk<T>(ua.x * fk, ub.x * fk);
        if (p.kind == RK_DST) row = n - 1 - k;
      }
      base[(i64)row * p.s] = o;
    }
  } else {
#pragma unroll
    for (int q = 0; q < S::E; ++q) {
      const int row = t + q * S::TPL;
      v[q] = base[(i64)(p.kind == RK_DST ? n - 1 - row : row) * p.s];
    }
#pragma unroll
    for (int q = 0; q < S::E; ++q) sm[A::at(t + q * S::TPL, w)] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < S::E; ++q) {
      const int row = t + q * S::TPL;
      const C a = v[q];
      const C bq = sm[A::at((n - row) & (n - 1), w)];
      C z;
      if (row == 0) z = mk<T>(a.x * p.f0, a.y * p.f0);
      else {
        const C d = __ldg(p.dtw + row);
        const C s2 = mk<T>((a.x + bq.y) * hf * p.f, (a.y - bq.x) * hf * p.f);
        z = cmulc(s2, d);
      }
      v[q] = cswap(z);
    }
    __syncthreads();
    FastLoop<T, S, 0, true, W>::run(v, sm, twt, t, w, p.twg);
#pragma unroll
    for (int q = 0; q < S::E; ++q) {
      const int m = t + q * S::TPL;
      const bool second = q >= S::E / 2;
      const int row = second ? 2 * (n - 1 - m) + 1 : 2 * m;
      C z = cswap(v[q]);
      if (p.kind == RK_DST && second) { z.x = -z.x; z.y = -z.y; }
      base[(i64)row * p.s] = z;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// realInverse of contiguous lines of 2N reals held as the packed half spectrum (N complex slots: slot 0 =
// (Re X[0], Re X[N]), slot k = X[k]), in place: rftbsub + cftbsub of fft/DoubleFFT_1D.java:946-967 in one kernel.
//   Z[k] = (X[k] + conj X[N-k])/2 + i e^{+2 pi i k/2N} (X[k] - conj X[N-k])/2,   z = IDFT_N(Z) = x[2j] + i x[2j+1]
// (unnormalised: N x = (n/2) x, the reference's unscaled power-of-two result).  Thread k builds Z[k] (kept) and
// Z[N-k] (handed to the owner of the upper half); the result is stored straight from registers.
template <typename T> struct RfftInvParams {
  cx<T>* a;
  i64 nlines, dist;     // lines, distance between lines in complex units
  int has_scale;
  T scale;
  const cx<T>* twg;
  const cx<T>* rtw;     // exp(-2 pi i k / 2N), k <= N/2
};

template <typename T, int LOGN, int LOGE, int W>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, (Sched<LOGN, LOGE>::E <= 8 ? FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB
                                                                                      : (FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB + 1) / 2))
fft_rfft_inv_row_kernel(const RfftInvParams<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, false, W> A;
  constexpr int N = S::N, H = S::E / 2;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  const int tid = threadIdx.x;
  const int t = tid % S::TPL, w = tid / S::TPL;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W * S::TPL) twt[i] = __ldg(p.twg + i);
  const i64 line0 = (i64)blockIdx.x * W;
  const bool valid = line0 + w < p.nlines;
  C* cl = p.a + (valid ? (line0 + w) * p.dist : 0);
  C* half = sm + w * N;
  C xa[H], xb[H];
#pragma unroll
  for (int q = 0; q < H; ++q) {
    const int k = t + q * S::TPL;
    xa[q] = valid ? cl[k] : mk<T>(0, 0);
    xb[q] = valid ? cl[k == 0 ? N / 2 : N - k] : mk<T>(0, 0);
  }
  C v[S::E];
  const T hf = (T)0.5;
#pragma unroll
  for (int q = 0; q < H; ++q) {
    const int k = t + q * S::TPL;
    const C a = xa[q], b = xb[q];
    if (k == 0) {
      v[q] = cswap(mk<T>((a.x + a.y) * hf, (a.x - a.y) * hf));      // Z[0] from (Re X[0], Re X[N])
      half[N / 2] = cswap(mk<T>(b.x, -b.y));                         // Z[N/2] = conj X[N/2]
    } else {
      const C wk = __ldg(p.rtw + k);
      const C ev = mk<T>((a.x + b.x) * hf, (a.y - b.y) * hf);        // (a + conj b)/2
      const C df = mk<T>((a.x - b.x) * hf, (a.y + b.y) * hf);        // (a - conj b)/2
      C od = cmulc(df, wk);
      od = mk<T>(-od.y, od.x);                                       // * i
      v[q] = cswap(cadd(ev, od));                                    // Z[k]
      half[N - k] = cswap(mk<T>(ev.x - od.x, -(ev.y - od.y)));       // Z[N-k] = conj(ev - od)
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = H; q < S::E; ++q) v[q] = half[t + q * S::TPL];
  __syncthreads();
  FastLoop<T, S, 0, false, W>::run(v, sm, twt, t, w, p.twg);
  if (!valid) return;
#pragma unroll
  for (int q = 0; q < S::E; ++q) {
    C z = cswap(v[q]);
    if (p.has_scale) { z.x *= p.scale; z.y *= p.scale; }
    cl[t + q * S::TPL] = z;
  }
}

}  // namespace jtb
