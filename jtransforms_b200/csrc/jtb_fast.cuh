// Lean in-place line FFT kernel for the hot power-of-two lengths (the axis passes of DoubleFFT_2D/3D and
// the batched 1-D transforms).  Same Stockham register/shared-memory structure as fft_tile_kernel
// (jtb_tile.cuh) but with everything the hot path does not need stripped: 32-bit in-tile addressing,
// compile-time line layout and lines-per-CTA, per-stage base twiddles w^(k), w^(2k), w^(4k).. staged in
// shared memory with the remaining powers derived by multiplication, no fusion hooks except the
// inverse (re<->im swap) and an output scale.
//
// Replaces, per line: utils/CommonUtils.java:708-793 (cftfsub/cftbsub) + bitrv2/bitrv2conj :824-2070;
// per launch: the row/column/slice loops fft/DoubleFFT_2D.java:3100-3140, :3352-3529 and
// fft/DoubleFFT_3D.java:5505-5713, :6318-6520.
#pragma once
#include "jtb_tile.cuh"

#ifndef JTB_TW_DERIVE
#define JTB_TW_DERIVE 0
#endif

namespace jtb {

template <typename T> struct FastParams {
  cx<T>* a;          // transformed in place
  i64 nlines;
  i64 line_dist;     // contiguous layout: distance between consecutive lines
                     // strided layout: distance between groups of c0 adjacent lines
  int c0;            // strided layout: number of adjacent lines (distance 1) per group
  int stride;        // strided layout: distance between consecutive elements of a line
  int inverse;
  int has_scale;
  T scale;
  const cx<T>* twg;  // base twiddles, layout below (fast_twiddle_count entries)
  int reps;          // strided layout: consecutive groups of W lines handled by one CTA (TLB / launch amortisation)
  int prefetch;      // strided layout: prefetch.global.L2 of the tile `prefetch` CTAs ahead (0 = off)
  // strided layout, out of place: the same line groups stored with another group distance / element stride (the
  // axis-swapping k2 pass of the 3-D transform).  out == a, same distances: in place.
  cx<T>* out;
  i64 out_line_dist;
  int out_stride;
};

// four consecutive reals with one request: 256-bit LDG/STG for double (sm_100a: ld.global.v4.f64, 32-byte aligned),
// 128-bit for float.  A 32-byte-strided pair of 16-byte accesses would touch every sector twice.
__device__ __forceinline__ void ld4(const double* p, double& a, double& b, double& c, double& d) {
#ifdef JTB_EMU
  a = p[0]; b = p[1]; c = p[2]; d = p[3];
#else
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
#endif
}
__device__ __forceinline__ void ld4(const float* p, float& a, float& b, float& c, float& d) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  a = v.x; b = v.y; c = v.z; d = v.w;
}
__device__ __forceinline__ void st4(double* p, double a, double b, double c, double d) {
#ifdef JTB_EMU
  p[0] = a; p[1] = b; p[2] = c; p[3] = d;
#else
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
#endif
}
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  float4 v; v.x = a; v.y = b; v.z = c; v.w = d;
  *reinterpret_cast<float4*>(p) = v;
}

// The DCT/DST twiddles a pair (k, N-k) of a length-n = 2N line needs, all from ONE table read D = exp(-i pi k/2n):
//   D(N-k) = c conj(D), D(N+k) = c D, D(n-k) = -i conj(D), c = exp(-i pi/4);  exp(-2 pi i k/n) = D^4.
template <typename C> __device__ __forceinline__ C dct_tw_nmk(C d) {   // D(N-k)
  typedef decltype(d.x) T;
  const T s = (T)0.70710678118654752440084436210485L;
  C r; r.x = s * (d.x - d.y); r.y = -s * (d.x + d.y); return r;
}
template <typename C> __device__ __forceinline__ C dct_tw_npk(C d) {   // D(N+k)
  typedef decltype(d.x) T;
  const T s = (T)0.70710678118654752440084436210485L;
  C r; r.x = s * (d.x + d.y); r.y = s * (d.y - d.x); return r;
}
template <typename C> __device__ __forceinline__ C dct_tw_pow4(C d) {  // D^4 = exp(-2 pi i k/n)
  C d2; d2.x = (d.x - d.y) * (d.x + d.y); d2.y = 2 * d.x * d.y;
  C d4; d4.x = (d2.x - d2.y) * (d2.x + d2.y); d4.y = 2 * d2.x * d2.y;
  return d4;
}

// base twiddle table: for stage s >= 1 and j < bits(s):  tab[toff(s) + j*ns(s) + k] = exp(-2 pi i 2^j k / (ns(s) 2^bits(s)))
// Only the tables of the early stages (few entries, every entry reused by many threads of the CTA) are staged in
// shared memory; a late stage with ns(s)*bits(s) > SM_LIMIT entries has (almost) no reuse inside one CTA, so its
// twiddles are read straight from global memory (coalesced, L1/L2 resident) instead of being copied per CTA.
template <typename S> struct FastTw {
  static constexpr int SM_LIMIT = 768;
  __host__ __device__ static constexpr int toff(int s) { int o = 0; for (int i = 1; i < s; ++i) o += S::bits(i) * S::ns(i); return o; }
  __host__ __device__ static constexpr bool in_smem(int s) { return S::bits(s) * S::ns(s) <= SM_LIMIT; }
  __host__ __device__ static constexpr int sm_count() { int o = 0; for (int i = 1; i < S::S; ++i) if (in_smem(i)) o = toff(i + 1); return o; }
  static constexpr int COUNT = toff(S::S);
  static constexpr int COUNT_SM = sm_count();
};

template <typename T, typename S, bool STRIDED, int W> struct FastAddr {
  // Bank-conflict rules (32 banks x 4 B; a 16-byte access is served per quarter warp, an 8-byte one per half warp):
  //  * contiguous layout [w][i]: one pad element every 2^PADSH elements.  double2: PADSH = log2(E) (a quarter warp
  //    never straddles a pad); float2: at least 4, otherwise 16 consecutive elements cross a pad and collide.
  //  * line-interleaved layout [i][w]: W*sizeof(complex) >= 128 B needs no padding; 64-byte rows (float2, W = 8)
  //    get one row of padding every 8 rows so that rows i and i+8 (stage-0 scatter) fall in different halves.
  static constexpr int PADSH = (sizeof(T) == 4 && S::LOGPAD < 4) ? 4 : S::LOGPAD;
  static constexpr bool ROWPAD = STRIDED && (W * 2 * sizeof(T) < 128);
  static constexpr int LD = STRIDED ? S::N : S::LD;
  static constexpr int TILE = STRIDED ? (ROWPAD ? S::N * W + (S::N >> 3) * W + W : S::N * W) : S::LD * W;
  __device__ static __forceinline__ int at(int i, int w) {
    if (STRIDED) return ROWPAD ? (i + (i >> 3)) * W + w : i * W + w;
    return w * S::LD + i + (i >> PADSH);
  }
};

template <typename T, typename S, int s, bool STRIDED, int W> struct FastStage {
  static constexpr int LOGR = S::bits(s);
  static constexpr int R = 1 << LOGR;
  static constexpr int NB = S::E / R;
  static constexpr int NS = S::ns(s);
  typedef FastAddr<T, S, STRIDED, W> A;

  __device__ static __forceinline__ void compute(cx<T>* v, int t, const cx<T>* twt, const cx<T>* __restrict__ twg) {
#pragma unroll
    for (int m = 0; m < NB; ++m) {
      cx<T> x[R];
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = v[m + r * NB];
      if (s > 0) {
        const int k = (t + m * S::TPL) & (NS - 1);
        constexpr int TOFF = FastTw<S>::toff(s);
        cx<T> tw[R];
#if JTB_TW_DERIVE
        // one table read per butterfly; w^2k, w^4k, w^8k by squaring (trades 2-3 complex multiplies for the
        // shared-memory / L1 wavefronts of the other base twiddles)
        tw[1] = FastTw<S>::in_smem(s) ? twt[TOFF + k] : __ldg(twg + TOFF + k);
#pragma unroll
        for (int j = 1; j < LOGR; ++j) tw[1 << j] = cmul(tw[1 << (j - 1)], tw[1 << (j - 1)]);
#else
#pragma unroll
        for (int j = 0; j < LOGR; ++j)
          tw[1 << j] = FastTw<S>::in_smem(s) ? twt[TOFF + j * NS + k] : __ldg(twg + TOFF + j * NS + k);
#endif
#pragma unroll
        for (int r = 3; r < R; ++r) {
          // highest set bit h of r; r = 2^h + rest
          const int h = (r >= 8) ? 8 : ((r >= 4) ? 4 : 2);
          if (r != h) tw[r] = cmul(tw[h], tw[r - h]);
        }
#pragma unroll
        for (int r = 1; r < R; ++r) x[r] = cmul(x[r], tw[r]);
      }
      Bfly<T, R>::run(x);
#pragma unroll
      for (int r = 0; r < R; ++r) v[m + r * NB] = x[r];
    }
  }
  __device__ static __forceinline__ void scatter(const cx<T>* v, cx<T>* sm, int t, int w) {
#pragma unroll
    for (int m = 0; m < NB; ++m) {
      const int jv = t + m * S::TPL;
      const int k = jv & (NS - 1);
      const int j0 = ((jv - k) << LOGR) + k;
#pragma unroll
      for (int r = 0; r < R; ++r) sm[A::at(j0 + r * NS, w)] = v[m + r * NB];
    }
  }
};

// Barrier used between the stages: the whole CTA (default), or a named barrier over one compute group of a
// warp-specialised persistent kernel (jtb_tma.cuh), where several groups work on different tiles of one CTA.
struct SyncCta {
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};
template <int THREADS> struct SyncGroup {
  int id;   // named barrier 1..15
  __device__ __forceinline__ void sync() const {
#ifdef JTB_EMU
    __syncthreads();
#else
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(THREADS) : "memory");
#endif
  }
};

template <typename T, typename S, int s, bool STRIDED, int W, typename SY = SyncCta> struct FastLoop {
  typedef FastAddr<T, S, STRIDED, W> A;
  __device__ static __forceinline__ void run(cx<T>* v, cx<T>* sm, const cx<T>* twt, int t, int w, const cx<T>* twg,
                                             const SY sy = SY()) {
    FastStage<T, S, s, STRIDED, W>::compute(v, t, twt, twg);
    if (s + 1 < S::S) {
      if (s > 0) sy.sync();
      FastStage<T, S, s, STRIDED, W>::scatter(v, sm, t, w);
      sy.sync();
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = sm[A::at(t + q * S::TPL, w)];
      FastLoop<T, S, (s + 1 < S::S ? s + 1 : s), STRIDED, W, SY>::run(v, sm, twt, t, w, twg, sy);
    }
  }
};

template <int THREADS> struct FastOcc {   // resident CTAs per SM we compile for
  static constexpr int MINB = THREADS >= 1024 ? 1 : (THREADS >= 512 ? 2 : (THREADS >= 256 ? 4 : (THREADS >= 128 ? 8 : 12)));
};

template <typename T, int LOGN, int LOGE, bool STRIDED, int W>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, (Sched<LOGN, LOGE>::E <= 8 ? FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB
                                                                                      : (FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB + 1) / 2))
fft_fast_kernel(const FastParams<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, STRIDED, W> A;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  const int tid = threadIdx.x;
  int w, t;
  if (STRIDED) { w = tid % W; t = tid / W; } else { t = tid % S::TPL; w = tid / S::TPL; }
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W * S::TPL) twt[i] = __ldg(p.twg + i);

#ifndef JTB_EMU
  if (STRIDED && p.prefetch > 0) {
    // pull the rows of a later tile into L2 while this one is transformed (one 128-byte row segment per thread)
    const i64 pl0 = ((i64)blockIdx.x + p.prefetch) * W;
    if (pl0 < p.nlines) {
      const i64 grp = pl0 / p.c0;
      const C* pb = p.a + grp * p.line_dist + (pl0 - grp * p.c0);
      for (int i = tid; i < S::N; i += W * S::TPL) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + (i64)i * p.stride));
    }
  }
#endif
  for (int rep = 0; rep < (STRIDED ? p.reps : 1); ++rep) {
    const i64 line0 = ((i64)blockIdx.x * (STRIDED ? p.reps : 1) + rep) * W;
    if (line0 >= p.nlines) break;
    C* base;
    C* obase;
    int es, oes;
    bool valid = true;
    if (STRIDED) {
      const i64 grp = line0 / p.c0;
      const int c = (int)(line0 - grp * p.c0);
      base = p.a + grp * p.line_dist + c + w;
      es = p.stride;
      obase = p.out + grp * p.out_line_dist + c + w;
      oes = p.out_stride;
    } else {
      valid = line0 + w < p.nlines;
      base = p.a + (valid ? (line0 + w) * p.line_dist : 0);
      es = 1;
      obase = base; oes = 1;
    }
    C v[S::E];
    if (valid) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = base[(t + q * S::TPL) * es];
    } else {
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = mk<T>(0, 0);
    }
    if (p.inverse) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
    }
    if (rep > 0) __syncthreads();      // the previous group's last shared-memory reads are done
    FastLoop<T, S, 0, STRIDED, W>::run(v, sm, twt, t, w, p.twg);
    if (p.has_scale) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) { v[q].x *= p.scale; v[q].y *= p.scale; }
    }
    if (p.inverse) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
    }
    if (valid) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) obase[(i64)(t + q * S::TPL) * oes] = v[q];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// k2 pass of a slab-decomposed 3-D transform with the re-slabbing all-to-all fused into its stores.
// Rank g holds slices [g*Ls, (g+1)*Ls) as [Ls][R][C].  The kernel transforms the columns (length R,
// stride C) of every local slice and stores output row k2 straight into the receive buffer of the GPU
// that owns it after the exchange: peer h = k2 / Rh gets it at [g*Ls + ls][k2 % Rh][c] of its
// [S][Rh][C] block.  peer[h] are peer-mapped device pointers (NVLink P2P; peer[g] is local memory).
// Replaces the slice-axis gather of cdft3db_subth (fft/DoubleFFT_3D.java:6318-6520) across GPUs.
template <typename T> struct ScatterParams {
  const cx<T>* a;     // local slab [Ls][R][C]
  cx<T>* peer[8];     // receive buffers [S][Rh][C], one per rank
  int Ls, C, logRh;   // local slices, columns, log2(R / nranks)
  int slice0;         // g * Ls
  int inverse;
  const cx<T>* twg;
  // destination row of output row k2 of local slice ls inside the owner's buffer:
  //   row_base + ls*row_ls_mul + (k2 % Rh)*row_mul
  // forward re-slabbing: (slice0 + ls)*Rh + rl  ->  row_base = slice0*Rh, row_ls_mul = Rh, row_mul = 1;
  // inverse re-slabbing (one [S][Rh*C] "slice", owner layout [Ls][R][C]): rl*P + g  ->  row_base = g, row_mul = P
  long long row_base;
  int row_ls_mul, row_mul;
  // column window of this launch (the pipelined exchange sends the slab column block by column block so that the
  // slice-axis pass of a block can start as soon as that block has arrived from every peer): columns
  // [col0, col0 + groups*W) of every local slice
  int col0, groups;
};

template <typename T, int LOGN, int LOGE, int W>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, (Sched<LOGN, LOGE>::E <= 8 ? FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB
                                                                                      : (FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB + 1) / 2))
fft_scatter_kernel(const ScatterParams<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, true, W> A;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  const int tid = threadIdx.x;
  const int w = tid % W, t = tid / W;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W * S::TPL) twt[i] = __ldg(p.twg + i);
  const int groups = p.groups;                      // column groups per slice in this launch's window
  const int ls = blockIdx.x / groups;
  const int c = p.col0 + (blockIdx.x - ls * groups) * W + w;
  const C* src = p.a + (i64)ls * S::N * p.C + c;
  C v[S::E];
#pragma unroll
  for (int q = 0; q < S::E; ++q) v[q] = src[(i64)(t + q * S::TPL) * p.C];
  if (p.inverse) {
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
  }
  FastLoop<T, S, 0, true, W>::run(v, sm, twt, t, w, p.twg);
  if (p.inverse) {
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
  }
  const int Rh = 1 << p.logRh;
  const i64 row0 = p.row_base + (i64)ls * p.row_ls_mul;
#pragma unroll
  for (int q = 0; q < S::E; ++q) {
    const int k2 = t + q * S::TPL;
    const int h = k2 >> p.logRh;
    const int rl = k2 & (Rh - 1);
    p.peer[h][(row0 + (i64)rl * p.row_mul) * p.C + c] = v[q];
  }
}

// ---------------------------------------------------------------------------------------------------
// Both in-slice passes of a 3-D (or batched 2-D) transform in ONE persistent kernel: a team of `team` CTAs owns
// one N x N slice at a time, transforms its rows (phase A), meets at a team barrier (one counter per slice in
// global memory), then transforms its columns (phase B).  The intermediate never has to come back from HBM: a
// slice is 4 MiB (512^2 complex doubles) and all teams together keep well under the 126 MB L2, so phase B reads
// L2 hits and phase A's dirty lines are overwritten before they are evicted.  Replaces xdft3da_subth2
// (fft/DoubleFFT_3D.java:5505-5713).  With `scatter` the phase-B stores are the slab all-to-all (see
// fft_scatter_kernel).  Launched cooperatively so that every CTA of a team is resident.
template <typename T> struct Slice2DParams {
  cx<T>* a;            // [nslices][N][N]
  int nslices, team;
  int* counters;       // nslices arrival counters, zeroed before the launch
  int* err;
  const cx<T>* twg;
  int inverse;
  int has_scale; T scale;
  int scatter, logRh, slice0;
  int pipelined;       // 1: rows of the next slice before the columns of the current one
  cx<T>* peer[8];
};

template <typename T, int LOGN, int LOGE, int W>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, 2) fft_slice2d_kernel(const Slice2DParams<T> p) {
#ifndef JTB_EMU
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, false, W> AC;
  typedef FastAddr<T, S, true, W> AS;
  constexpr int N = S::N;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + (AC::TILE > AS::TILE ? AC::TILE : AS::TILE);
  const int tid = threadIdx.x;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W * S::TPL) twt[i] = __ldg(p.twg + i);
  const int nteams = gridDim.x / p.team;
  const int team = blockIdx.x / p.team, rank = blockIdx.x - team * p.team;
  if (team >= nteams) return;
  const int per = N / p.team;                 // rows (phase A) / columns (phase B) per CTA
  C v[S::E];
  // Software-pipelined over slices: a team transforms the rows of its NEXT slice before it waits for the rows of the
  // current one, so the arrival counter has long been complete when it is polled (no team-barrier stall) and the
  // column phase still finds the slice in L2 (two slices of 4 MiB per team in flight).
  auto rows = [&](int slice) {
    C* sl = p.a + (i64)slice * N * N;
    const int t = tid % S::TPL, w = tid / S::TPL;
    for (int r0 = rank * per; r0 < (rank + 1) * per; r0 += W) {
      C* base = sl + (i64)(r0 + w) * N;
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = base[t + q * S::TPL];
      if (p.inverse) {
#pragma unroll
        for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
      }
      FastLoop<T, S, 0, false, W>::run(v, sm, twt, t, w, p.twg);
#pragma unroll
      for (int q = 0; q < S::E; ++q) base[t + q * S::TPL] = v[q];   // stays in the swapped domain for the columns
      __syncthreads();
    }
    // arrive: this CTA's rows are written (and visible at L2) before the counter moves
    if (tid == 0) {
      __threadfence();
      atomicAdd(p.counters + slice, 1);
    }
  };
  auto wait_rows = [&](int slice) {
    if (tid == 0) {
      const long long t0 = clock64();
      while (*((volatile int*)(p.counters + slice)) < p.team) {
        if (clock64() - t0 > 8000000000LL) { *p.err = 1; break; }
      }
      __threadfence();
    }
    __syncthreads();
  };
  auto cols = [&](int slice) {
    C* sl = p.a + (i64)slice * N * N;
    const int w = tid % W, t = tid / W;
    for (int c0 = rank * per; c0 < (rank + 1) * per; c0 += W) {
      const C* base = sl + c0 + w;
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = __ldcg(base + (i64)(t + q * S::TPL) * N);
      FastLoop<T, S, 0, true, W>::run(v, sm, twt, t, w, p.twg);
      if (p.has_scale) {
#pragma unroll
        for (int q = 0; q < S::E; ++q) { v[q].x *= p.scale; v[q].y *= p.scale; }
      }
      if (p.inverse) {
#pragma unroll
        for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
      }
      if (p.scatter) {
        const int Rh = 1 << p.logRh;
        const i64 row0 = (i64)(p.slice0 + slice) * Rh;
#pragma unroll
        for (int q = 0; q < S::E; ++q) {
          const int k2 = t + q * S::TPL;
          p.peer[k2 >> p.logRh][(row0 + (k2 & (Rh - 1))) * N + c0 + w] = v[q];
        }
      } else {
        C* dst = sl + c0 + w;
#pragma unroll
        for (int q = 0; q < S::E; ++q) dst[(i64)(t + q * S::TPL) * N] = v[q];
      }
      __syncthreads();
    }
  };
  int slice = team;
  if (slice < p.nslices) rows(slice);
  while (slice < p.nslices) {
    const int next = slice + nteams;
    if (next < p.nslices && p.pipelined) rows(next);
    wait_rows(slice);
    cols(slice);
    if (next < p.nslices && !p.pipelined) rows(next);
    slice = next;
  }
#endif
}

// ---------------------------------------------------------------------------------------------------
// Exchange and slice-axis pass of the slab-decomposed 3-D transform in ONE persistent kernel (cube-like shapes, R == S):
// the CTAs pull work items from an atomic queue whose order interleaves
//     scatter tiles of column block j+1   (k2 pass of W columns of one local slice, stores to the owners: NVLink)
//     k1 tiles of column block j          (slice-axis pass of W columns of one received row: local HBM)
// so that NVLink-bound and HBM-bound tiles share the SMs at a fixed ratio instead of two kernels fighting for slots on
// two streams.  The last scatter tile of block j on a rank publishes `epoch` into slot (rank, j) of every peer's flag
// array; a k1 tile of block j polls the P slots (rank h, j) of its own array before it loads.  Scatter tiles never wait,
// and every scatter tile of blocks <= j+1 has been handed out before the first k1 tile of block j, so the spin-waits
// cannot dead-lock however the ranks drift; one stream, no second kernel.  Flag layout: slots 0..7 belong to
// peer_barrier_kernel, slot 8 + 16*rank + block to this kernel (256 int64 per rank).
template <typename T> struct PipeParams {
  const cx<T>* a;        // local slab [Ls][R][C] (rows already transformed)
  cx<T>* peer[8];        // receive buffers [S][Rh][C] of this step, one per rank
  cx<T>* recv;           // this rank's receive buffer (slice-axis pass in place)
  long long* flags[8];   // flag arrays of all ranks
  int* counters;         // [0] work queue, [1 + j] scatter tiles of block j finished on this rank (zeroed before the launch)
  int* err;
  const cx<T>* twg;
  long long epoch;
  int Ls, C, logRh, P, rank, nb, gpb;   // gpb = W-column groups per block
  int inverse, has_scale;
  T scale;
};

template <typename T, int LOGN, int LOGE, int W>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, (Sched<LOGN, LOGE>::E <= 8 ? FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB
                                                                                      : (FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB + 1) / 2))
fft_pipe_kernel(const PipeParams<T> p) {
#ifndef JTB_EMU
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, true, W> A;
  JTB_DYN_SMEM(smem_raw);
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  __shared__ int s_item;
  const int tid = threadIdx.x;
  const int w = tid % W, t = tid / W;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W * S::TPL) twt[i] = __ldg(p.twg + i);
  const int Rh = 1 << p.logRh;
  const int ns = p.Ls * p.gpb;               // tiles per block, both kinds (R == S  =>  Ls * gpb == Rh * gpb)
  const int total = 2 * p.nb * ns;
  const int cc = p.gpb * W;
  // Scatter tiles are counted lazily: a system-scope fence per tile would hold the CTA for an NVLink round trip.  The CTA
  // remembers how many tiles of block `pend_blk` it has stored and flushes -- every thread fences, one thread adds the
  // count, the CTA that completes the block publishes it -- when it draws an item beyond that block's scatter range.
  int pend_blk = -1, pend_cnt = 0;
  auto flush = [&]() {
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      if (atomicAdd(p.counters + 1 + pend_blk, pend_cnt) + pend_cnt == ns) {
        __threadfence_system();
        for (int h = 0; h < p.P; ++h) *((volatile long long*)(p.flags[h] + 8 + p.rank * 16 + pend_blk)) = p.epoch;
        __threadfence_system();
      }
    }
    pend_cnt = 0;
  };
  for (;;) {
    __syncthreads();                         // previous tile's shared memory is free; s_item has been read
    if (tid == 0) s_item = atomicAdd(p.counters, 1);
    __syncthreads();
    const int item = s_item;
    if (pend_cnt > 0 && item >= (pend_blk == 0 ? ns : ns + pend_blk * 2 * ns)) flush();
    if (item >= total) break;
    // decode: phase 0 = scatter block 0; phases 1..nb-1 = scatter block ph interleaved with k1 block ph-1; phase nb = k1 block nb-1
    int blk, idx;
    bool scatter;
    if (item < ns) { scatter = true; blk = 0; idx = item; }
    else if (item >= total - ns) { scatter = false; blk = p.nb - 1; idx = item - (total - ns); }
    else {
      const int r = item - ns, ph = r / (2 * ns), o = r - ph * 2 * ns;
      scatter = (o & 1) == 0; idx = o >> 1; blk = scatter ? ph + 1 : ph;
    }
    const int i1 = idx / p.gpb, cg = idx - i1 * p.gpb;
    const int c = blk * cc + cg * W + w;
    C v[S::E];
    if (scatter) {
      const C* src = p.a + (i64)i1 * S::N * p.C + c;
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = src[(i64)(t + q * S::TPL) * p.C];
    } else {
      if (tid == 0) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (int h = 0; h < p.P; ++h) {
          volatile long long* f = p.flags[p.rank] + 8 + h * 16 + blk;
          while (*f < p.epoch) {
            __nanosleep(200);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 10000000000ULL) { *p.err = 2; break; }
          }
        }
        __threadfence_system();
      }
      __syncthreads();
      const C* src = p.recv + (i64)i1 * p.C + c;          // row i1 of every slice: elements Rh*C apart
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = __ldcg(src + (i64)(t + q * S::TPL) * Rh * p.C);
    }
    if (p.inverse) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
    }
    FastLoop<T, S, 0, true, W>::run(v, sm, twt, t, w, p.twg);
    if (!scatter && p.has_scale) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) { v[q].x *= p.scale; v[q].y *= p.scale; }
    }
    if (p.inverse) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
    }
    if (scatter) {
      const i64 row0 = ((i64)p.rank * p.Ls + i1) * Rh;
#pragma unroll
      for (int q = 0; q < S::E; ++q) {
        const int k2 = t + q * S::TPL;
        p.peer[k2 >> p.logRh][(row0 + (k2 & (Rh - 1))) * p.C + c] = v[q];
      }
      pend_blk = blk;                         // (a CTA never holds tiles of two blocks: the flush above came first)
      ++pend_cnt;
    } else {
      C* dst = p.recv + (i64)i1 * p.C + c;
#pragma unroll
      for (int q = 0; q < S::E; ++q) dst[(i64)(t + q * S::TPL) * Rh * p.C] = v[q];
    }
  }
#endif
}

// cross-GPU barrier: thread h publishes `epoch` into rank h's flag array (slot = my rank) and waits until
// rank h has published it into mine.  flags are peer-mapped int64[nranks] arrays, monotonically increasing.
struct PeerFlags { long long* f[8]; };
// what = 1: publish only (signal), 2: wait only, 3: both (barrier)
static __global__ void peer_barrier_kernel(const PeerFlags flags, int nranks, int rank, long long epoch, int* err, int what) {
#ifndef JTB_EMU
  const int h = threadIdx.x;
  if (h >= nranks) return;
  if (what & 1) {
    __threadfence_system();
    volatile long long* theirs = flags.f[h] + rank;
    *theirs = epoch;
    __threadfence_system();
  }
  if (what & 2) {
    volatile long long* mine = flags.f[rank] + h;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*mine < epoch) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ULL) { *err = 2; break; }   // 10 s: a peer died; give up instead of hanging the GPU
    }
    __threadfence_system();
  }
#endif
}

// host-side description of one instantiation
struct FastInfo {
  int logn, loge, strided, W, threads, smem_bytes, tw_count;
  int nstages, bits[JTB_MAX_STAGES];
};

}  // namespace jtb
