// kernels.h -- host-side declarations shared between translation units.
#pragma once
#include "common.cuh"

namespace socm {
// Repack the nn.Linear weights of the default-arch UNet into the forward / backward tapes and
// the small block (unet_tile.cuh); `packed` has tile::packed_floats(d) floats.
int pack_tape(const socm_unet* net, float* packed, cudaStream_t stream);
}  // namespace socm
