// Host-side description and launcher of the tile FFT kernel instantiations.
#pragma once
#include "jtb_tile.cuh"

namespace jtb {

// elements-per-thread exponent used for each line length (tuning knob)
constexpr int loge_for(int /*logn*/) { return 4; }

struct TileInfo {
  int logn, n, nstages, bits[JTB_MAX_STAGES], e, tpl, ld, maxt;
};

template <typename T> struct TileLimits;
template <> struct TileLimits<double> { static constexpr int MAX_LOGN = 13; };
template <> struct TileLimits<float> { static constexpr int MAX_LOGN = 14; };

TileInfo tile_info(int logn);

// launches fft_tile_kernel<T, logn, loge_for(logn)>; smem_bytes may be 0 when no exchange is needed
template <typename T>
cudaError_t launch_tile(int logn, const TileParams<T>& p, unsigned grid, unsigned block, size_t smem_bytes,
                        cudaStream_t stream);
// raises the dynamic shared memory limit of every instantiation (call once per device)
template <typename T> cudaError_t tile_init_device();

}  // namespace jtb
