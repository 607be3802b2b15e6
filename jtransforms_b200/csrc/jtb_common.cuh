// Common device/host helpers for libjtb200 (complex arithmetic, line geometry).
#pragma once
#ifdef JTB_EMU_BUILD
#include "emu_cuda.h"
#else
#include <cuda_runtime.h>
#define JTB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define JTB_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif
#include <cstdint>

namespace jtb {

typedef long long i64;

template <typename T> struct CxOf;
template <> struct CxOf<double> { typedef double2 type; };
template <> struct CxOf<float> { typedef float2 type; };
template <typename T> using cx = typename CxOf<T>::type;

template <typename T> __host__ __device__ __forceinline__ cx<T> mk(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <typename C> __host__ __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __host__ __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
// a * b
template <typename C> __host__ __device__ __forceinline__ C cmul(C a, C b) {
  C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// a * conj(b)
template <typename C> __host__ __device__ __forceinline__ C cmulc(C a, C b) {
  C r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r;
}
// a * (-i)
template <typename C> __host__ __device__ __forceinline__ C cmul_mi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }
template <typename C> __host__ __device__ __forceinline__ C cswap(C a) { C r; r.x = a.y; r.y = a.x; return r; }
template <typename C> __host__ __device__ __forceinline__ C cconj(C a) { C r; r.x = a.x; r.y = -a.y; return r; }

// A batch of equally long lines laid out with four index levels (level 3 is unbounded):
//   line l -> i0 = l % c[0], i1 = (l / c[0]) % c[1], i2 = (l / (c[0] c[1])) % c[2], i3 = the rest
//   first element at  sum_k i_k * d[k], element j of the line at  + j * stride
// Units are elements of the kernel's element type (complex or real).
struct Geo {
  i64 c[3];
  i64 d[4];
  i64 stride;
};
__host__ __device__ __forceinline__ i64 geo_off(const Geo& g, i64 line, i64* idx /*[4] or null*/ = nullptr) {
  i64 r = line, off = 0, ii[4];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (g.c[k] > 1) { const i64 q = r / g.c[k]; ii[k] = r - q * g.c[k]; r = q; } else ii[k] = 0;
    off += ii[k] * g.d[k];
  }
  ii[3] = r;
  off += r * g.d[3];
  if (idx) { idx[0] = ii[0]; idx[1] = ii[1]; idx[2] = ii[2]; idx[3] = ii[3]; }
  return off;
}
// lines 0..: offset line*dist, unit element stride
static inline Geo geo_contig(i64 dist) {
  Geo g; g.c[0] = g.c[1] = g.c[2] = 1; g.d[0] = g.d[1] = g.d[2] = 0; g.d[3] = dist; g.stride = 1; return g;
}
// c0 adjacent-ish lines (distance d0) repeated at distance d3, element stride `stride`
static inline Geo geo_make(i64 c0, i64 d0, i64 d3, i64 stride) {
  Geo g; g.c[0] = c0 < 1 ? 1 : c0; g.c[1] = g.c[2] = 1; g.d[0] = d0; g.d[1] = g.d[2] = 0; g.d[3] = d3; g.stride = stride; return g;
}
// insert a new bounded level (count, dist) at position `at` (0..2); the old level 2 must be free (c[2] == 1)
static inline Geo geo_insert(const Geo& g, int at, i64 count, i64 dist) {
  Geo r = g;
  for (int k = 2; k > at; --k) { r.c[k] = g.c[k - 1]; r.d[k] = g.d[k - 1]; }
  r.c[at] = count; r.d[at] = dist;
  return r;
}
// geometry with the level structure of `g` (after geo_insert at `at`) addressing a dense work buffer:
// the inserted level gets distance d_ins, the original line index L (mixed radix over the other levels)
// gets distance line_dist.
static inline Geo geo_work_like(const Geo& g, int at, i64 d_ins, i64 line_dist) {
  Geo r = g;
  i64 mul = line_dist;
  for (int k = 0; k < 3; ++k) {
    if (k == at) { r.d[k] = d_ins; continue; }
    r.d[k] = mul; mul *= g.c[k];
  }
  r.d[3] = mul;
  return r;
}

}  // namespace jtb
