// xde_tile_adaptive_s1.cu -- the 1-stage instantiations of the tiled adaptive solver (AdaptiveHeun).
#include "xde_tile_adaptive.cuh"

namespace xde {
int ad_tile_s1(const AdTileParams &p, cudaStream_t s) { return ad_tile_dispatch<1>(p, s); }
}  // namespace xde
