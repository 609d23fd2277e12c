// xde_tile_adaptive_s2.cu -- the 2-stage instantiations of the tiled adaptive solver (Fehlberg2).
#include "xde_tile_adaptive.cuh"

namespace xde {
int ad_tile_s2(const AdTileParams &p, cudaStream_t s) { return ad_tile_dispatch<2>(p, s); }
}  // namespace xde
