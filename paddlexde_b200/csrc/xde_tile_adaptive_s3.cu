// xde_tile_adaptive_s3.cu -- the 3-stage instantiations of the tiled adaptive solver (Bosh3).
#include "xde_tile_adaptive.cuh"

namespace xde {
int ad_tile_s3(const AdTileParams &p, cudaStream_t s) { return ad_tile_dispatch<3>(p, s); }
}  // namespace xde
