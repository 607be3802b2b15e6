// Internal interface of the slab decomposition (jtb_slab.cu) used by the plan layer (jtb_capi.cu).
#pragma once
#include "jtb_engine.h"

struct jtb_slab;

namespace jtb {
// one step of a same-process group (see jtb_slab.cu); the caller holds the contexts' mutexes
int slab_group_run(jtb_slab* const* ms, int n, void* const* a, bool back, bool inverse, bool scale, void** results,
                   cudaStream_t const* st);
}  // namespace jtb
