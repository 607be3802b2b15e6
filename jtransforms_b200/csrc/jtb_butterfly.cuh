// Register-resident forward DFT butterflies of radix 2/4/8/16 (sign e^{-2 pi i jk/R}).
// Inputs and outputs are in natural order in a local array; every index is a
// compile-time constant after unrolling so the arrays live in registers.
#pragma once
#include "jtb_common.cuh"

namespace jtb {

template <typename T, int R> struct Bfly;

template <typename T> struct Bfly<T, 1> {
  __host__ __device__ static __forceinline__ void run(cx<T>*) {}
};

template <typename T> struct Bfly<T, 2> {
  __host__ __device__ static __forceinline__ void run(cx<T>* x) {
    cx<T> a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  }
};

template <typename T> struct Bfly<T, 4> {
  __host__ __device__ static __forceinline__ void run(cx<T>* x) {
    cx<T> t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
    cx<T> t2 = cadd(x[1], x[3]), t3 = cmul_mi(csub(x[1], x[3]));
    x[0] = cadd(t0, t2);
    x[1] = cadd(t1, t3);
    x[2] = csub(t0, t2);
    x[3] = csub(t1, t3);
  }
};

template <typename T> struct Bfly<T, 8> {
  __host__ __device__ static __forceinline__ void run(cx<T>* x) {
    const T h = (T)0.70710678118654752440084436210485L;
    cx<T> e[4] = {x[0], x[2], x[4], x[6]};
    cx<T> o[4] = {x[1], x[3], x[5], x[7]};
    Bfly<T, 4>::run(e);
    Bfly<T, 4>::run(o);
    // o[k] *= W8^k
    cx<T> o1 = mk<T>((o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h);
    cx<T> o2 = cmul_mi(o[2]);
    cx<T> o3 = mk<T>((o[3].y - o[3].x) * h, -(o[3].x + o[3].y) * h);
    x[0] = cadd(e[0], o[0]); x[4] = csub(e[0], o[0]);
    x[1] = cadd(e[1], o1);   x[5] = csub(e[1], o1);
    x[2] = cadd(e[2], o2);   x[6] = csub(e[2], o2);
    x[3] = cadd(e[3], o3);   x[7] = csub(e[3], o3);
  }
};

template <typename T> struct Bfly<T, 16> {
  __host__ __device__ static __forceinline__ void run(cx<T>* x) {
    const T h = (T)0.70710678118654752440084436210485L;
    const T c1 = (T)0.92387953251128675612818318939679L;  // cos(pi/8)
    const T s1 = (T)0.38268343236508977172845998403040L;  // sin(pi/8)
    // x[4a+b]: DFT over a for each b
    cx<T> y[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      cx<T> t[4] = {x[b], x[4 + b], x[8 + b], x[12 + b]};
      Bfly<T, 4>::run(t);
#pragma unroll
      for (int c = 0; c < 4; ++c) y[b][c] = t[c];
    }
    // y[b][c] *= W16^{b c}
    y[1][1] = cmul(y[1][1], mk<T>(c1, -s1));
    y[1][2] = mk<T>((y[1][2].x + y[1][2].y) * h, (y[1][2].y - y[1][2].x) * h);   // W16^2 = W8^1
    y[1][3] = cmul(y[1][3], mk<T>(s1, -c1));                                      // W16^3
    y[2][1] = mk<T>((y[2][1].x + y[2][1].y) * h, (y[2][1].y - y[2][1].x) * h);   // W16^2
    y[2][2] = cmul_mi(y[2][2]);                                                   // W16^4
    y[2][3] = mk<T>((y[2][3].y - y[2][3].x) * h, -(y[2][3].x + y[2][3].y) * h);  // W16^6 = W8^3
    y[3][1] = cmul(y[3][1], mk<T>(s1, -c1));                                      // W16^3
    y[3][2] = mk<T>((y[3][2].y - y[3][2].x) * h, -(y[3][2].x + y[3][2].y) * h);  // W16^6
    y[3][3] = cmul(y[3][3], mk<T>(-c1, s1));                                      // W16^9
    // DFT over b for each c: X[c + 4d]
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      cx<T> t[4] = {y[0][c], y[1][c], y[2][c], y[3][c]};
      Bfly<T, 4>::run(t);
#pragma unroll
      for (int d = 0; d < 4; ++d) x[c + 4 * d] = t[d];
    }
  }
};

}  // namespace jtb
