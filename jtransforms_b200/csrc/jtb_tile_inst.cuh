// Explicit instantiation helper: included by tile_f64.cu / tile_f32.cu with JTB_TILE_T defined.
#pragma once
#include "jtb_tile_host.h"

namespace jtb {

template <typename T, int LOGN> struct TileInst {
  static cudaError_t launch(const TileParams<T>& p, unsigned grid, unsigned block, size_t smem, cudaStream_t st) {
    auto kfn = fft_tile_kernel<T, LOGN, loge_for(LOGN)>;
    JTB_LAUNCH(kfn, grid, block, smem, st, p);
    return cudaGetLastError();
  }
  static cudaError_t init() {
    auto kfn = fft_tile_kernel<T, LOGN, loge_for(LOGN)>;
    return cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
};

template <typename T, int LOGN> struct TileDispatch {
  static cudaError_t launch(int logn, const TileParams<T>& p, unsigned g, unsigned b, size_t s, cudaStream_t st) {
    if (logn == LOGN) return TileInst<T, LOGN>::launch(p, g, b, s, st);
    return TileDispatch<T, LOGN - 1>::launch(logn, p, g, b, s, st);
  }
  static cudaError_t init() {
    cudaError_t e = TileInst<T, LOGN>::init();
    if (e != cudaSuccess) return e;
    return TileDispatch<T, LOGN - 1>::init();
  }
};
template <typename T> struct TileDispatch<T, 0> {
  static cudaError_t launch(int, const TileParams<T>&, unsigned, unsigned, size_t, cudaStream_t) {
    return cudaErrorInvalidValue;
  }
  static cudaError_t init() { return cudaSuccess; }
};

template <>
cudaError_t launch_tile<JTB_TILE_T>(int logn, const TileParams<JTB_TILE_T>& p, unsigned grid, unsigned block,
                                    size_t smem_bytes, cudaStream_t stream) {
  return TileDispatch<JTB_TILE_T, TileLimits<JTB_TILE_T>::MAX_LOGN>::launch(logn, p, grid, block, smem_bytes, stream);
}
template <> cudaError_t tile_init_device<JTB_TILE_T>() {
  return TileDispatch<JTB_TILE_T, TileLimits<JTB_TILE_T>::MAX_LOGN>::init();
}

}  // namespace jtb
