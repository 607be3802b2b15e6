// fft_tile_kernel instantiations and Engine<float> (FloatFFT_*, FloatDCT_*, ...).
#define JTB_TILE_T float
#include "jtb_tile_inst.cuh"
#include "jtb_engine_impl.cuh"

namespace jtb {
template struct Engine<float>;
}  // namespace jtb
