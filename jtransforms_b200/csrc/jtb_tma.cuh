// Persistent, TMA-fed variant of the strided line FFT (the k2 / k1 axis passes of DoubleFFT_2D/3D).
//
// fft_fast_kernel<STRIDED> loads a tile of W adjacent lines (W * sizeof(complex) = 128 bytes per row, N rows a
// stride apart) with synchronous LDGs: while a CTA runs its butterflies nothing of its next tile is in flight, and
// for the 4 MiB-stride k1 pass of a 512^3 array every one of those N row segments is its own DRAM page and TLB
// entry.  Here ONE CTA per SM stays resident and walks over its tiles; the tile rows are fetched by the Tensor
// Memory Accelerator (cp.async.bulk.tensor, one elected thread, completion on an mbarrier) into a ring of NST
// shared-memory stages, so the tiles j+1 .. j+NST-1 stream in from HBM while the NG compute groups (W * TPL threads
// each, named barriers instead of __syncthreads) run the Stockham stages of tile j IN the stage's own shared memory.
// Results leave from registers with the same 128-byte-segment stores as the lean kernel (OUT == 0) or through the
// stage with a bulk tensor store (OUT == 1).
//
// Replaces the slice/column loops of fft/DoubleFFT_3D.java:6318-6520 (cdft3db_subth) and
// fft/DoubleFFT_2D.java:3352-3529 (cdft2d_subth), per line utils/CommonUtils.java:708-793.
#pragma once
#include "jtb_fast.cuh"

#ifndef JTB_EMU
#include <cuda.h>   // CUtensorMap and its enums only; the encoder is fetched with cudaGetDriverEntryPoint (no libcuda link)

namespace jtb {

template <typename T> struct TmaParams {
  cx<T>* a;            // transformed in place
  int ntiles;          // tiles of W adjacent lines
  int tiles_per_group; // c0 / W
  i64 line_dist;       // distance between groups of c0 adjacent lines (complex elements)
  int stride;          // distance between consecutive elements of a line
  int inverse, has_scale;
  T scale;
  const cx<T>* twg;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  // try_wait sleeps in hardware up to a time limit; bounded so that a tensor map that never delivers traps instead
  // of hanging the GPU
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// box {x .. , y .. , z} of a rank-3 tensor -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* src, const CUtensorMap* map, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(x), "r"(y), "r"(z)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void tma_store_2d(const void* src, const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(x), "r"(y)
               : "memory");
}

// rows per TMA box (box dimensions are limited to 256)
template <int N> struct TmaBox { static constexpr int ROWS = N > 256 ? 256 : N; };

template <typename T, int LOGN, int LOGE, int W, int NG, int NST, int OUT>
__global__ void __launch_bounds__(NG * W * Sched<LOGN, LOGE>::TPL, 1)
fft_tma_kernel(const __grid_constant__ CUtensorMap tmap, const TmaParams<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, true, W> A;
  static_assert(!A::ROWPAD, "TMA tiles are dense [row][W]");
  constexpr int GT = W * S::TPL;                       // threads of one compute group
  constexpr unsigned TILE_BYTES = (unsigned)(S::N * W * sizeof(C));
  constexpr int BOXR = TmaBox<S::N>::ROWS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* tiles = reinterpret_cast<C*>(smem_raw);
  C* twt = tiles + NST * A::TILE;
  // One mbarrier per tile slot j % NB, NB = NG * NST: slot b is always consumed by group b % NG, so every waiter sees
  // every phase of its barriers (with one barrier per STAGE a group that runs ahead would find the barrier one phase
  // further than it last saw and mistake the other group's tile for its own).
  constexpr int NB = NG * NST;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(twt + ((FastTw<S>::COUNT_SM + 7) & ~7));
  const int tid = threadIdx.x;
  const int g = tid / GT, gt = tid - g * GT;
  const int w = gt % W, t = gt / W;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += NG * GT) twt[i] = __ldg(p.twg + i);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NB; ++s) mbar_init(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nj = (int)blockIdx.x < p.ntiles ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  auto issue = [&](int j) {   // one thread: fetch tile j of this CTA into stage j % NST
    const int tile = j * (int)gridDim.x + (int)blockIdx.x;
    const int grp = tile / p.tiles_per_group, cg = tile - grp * p.tiles_per_group;
    const int s = j % NST, b = j % NB;
    mbar_expect_tx(full + b, TILE_BYTES);
#pragma unroll
    for (int r0 = 0; r0 < S::N; r0 += BOXR) tma_load_3d(tiles + s * A::TILE + r0 * W, &tmap, cg * W * 2, r0, grp, full + b);
  };
  if (tid == 0)
    for (int j = 0; j < NST && j < nj; ++j) issue(j);

  const SyncGroup<GT> sy{1 + g};
  for (int j = g; j < nj; j += NG) {
    const int s = j % NST;
    C* sm = tiles + s * A::TILE;
    mbar_wait(full + j % NB, (unsigned)((j / NB) & 1));
    C v[S::E];
#pragma unroll
    for (int q = 0; q < S::E; ++q) v[q] = sm[A::at(t + q * S::TPL, w)];
    if (p.inverse) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
    }
    sy.sync();                                       // the stage-0 scatter overwrites what the others just read
    FastLoop<T, S, 0, true, W, SyncGroup<GT>>::run(v, sm, twt, t, w, p.twg, sy);
    if (p.has_scale) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) { v[q].x *= p.scale; v[q].y *= p.scale; }
    }
    if (p.inverse) {
#pragma unroll
      for (int q = 0; q < S::E; ++q) v[q] = cswap(v[q]);
    }
    const int tile = j * (int)gridDim.x + (int)blockIdx.x;
    const int grp = tile / p.tiles_per_group, cg = tile - grp * p.tiles_per_group;
    sy.sync();                                       // every thread of the group is past its last gather
    if (OUT == 0) {
      if (gt == 0 && j + NST < nj) {
        fence_proxy_async();                         // generic-proxy accesses to the stage before the async-proxy refill
        issue(j + NST);
      }
      C* base = p.a + (i64)grp * p.line_dist + cg * W + w;
#pragma unroll
      for (int q = 0; q < S::E; ++q) base[(i64)(t + q * S::TPL) * p.stride] = v[q];
    } else {
      // results back into the stage, then one bulk tensor store; the stage is refilled once the store has read it
#pragma unroll
      for (int q = 0; q < S::E; ++q) sm[A::at(t + q * S::TPL, w)] = v[q];
      fence_proxy_async();
      sy.sync();
      if (gt == 0) {
#pragma unroll
        for (int r0 = 0; r0 < S::N; r0 += BOXR) tma_store_3d(sm + r0 * W, &tmap, cg * W * 2, r0, grp);
        tma_store_commit();
        if (j + NST < nj) {
          tma_store_wait_read<0>();
          issue(j + NST);
        }
      }
    }
  }
  if (OUT == 1 && gt == 0) tma_store_wait<0>();
}

// ---------------------------------------------------------------------------------------------------
// k2 pass of the slab-decomposed 3-D transform with the all-to-all carried by the TMA engine: same transform as
// fft_scatter_kernel (jtb_fast.cuh), but the finished tile goes back into shared memory and leaves as ONE bulk tensor
// store per destination GPU (rows [h*Rh, (h+1)*Rh) of the tile = an Rh x 128-byte box of peer h's receive buffer)
// instead of 16-byte LSU stores from every thread -- the SM's load/store unit no longer carries the NVLink traffic.
// maps.m[h]: rank-2 tensor map of peer h's [S*Rh rows][C columns] receive buffer (peer-mapped memory).
struct PeerMaps { CUtensorMap m[8]; };

template <typename T, int LOGN, int LOGE, int W>
__global__ void __launch_bounds__(W * Sched<LOGN, LOGE>::TPL, (Sched<LOGN, LOGE>::E <= 8 ? FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB
                                                                                      : (FastOcc<W * Sched<LOGN, LOGE>::TPL>::MINB + 1) / 2))
fft_scatter_tma_kernel(const __grid_constant__ PeerMaps maps, const ScatterParams<T> p) {
  typedef Sched<LOGN, LOGE> S;
  typedef cx<T> C;
  typedef FastAddr<T, S, true, W> A;
  static_assert(!A::ROWPAD, "TMA tiles are dense [row][W]");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  C* sm = reinterpret_cast<C*>(smem_raw);
  C* twt = sm + A::TILE;
  const int tid = threadIdx.x;
  const int w = tid % W, t = tid / W;
  for (int i = tid; i < FastTw<S>::COUNT_SM; i += W * S::TPL) twt[i] = __ldg(p.twg + i);
  const int groups = p.groups;
  const int ls = blockIdx.x / groups;
  const int c0 = p.col0 + (blockIdx.x - ls * groups) * W;
  const C* src = p.a + (i64)ls * S::N * p.C + c0 + w;
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
  if (S::S > 1) __syncthreads();                     // the last gather of every thread is done
#pragma unroll
  for (int q = 0; q < S::E; ++q) sm[A::at(t + q * S::TPL, w)] = v[q];
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    const int Rh = 1 << p.logRh, P = S::N >> p.logRh;
    const int row0 = (int)(p.row_base + (long long)ls * p.row_ls_mul);
    for (int h = 0; h < P; ++h) tma_store_2d(sm + (size_t)h * Rh * W, &maps.m[h], 2 * c0, row0);
    tma_store_commit();
    tma_store_wait_read<0>();                        // the tile has been read; the writes complete with the kernel
  }
}

}  // namespace jtb
#endif  // JTB_EMU
