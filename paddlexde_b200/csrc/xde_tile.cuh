// xde_tile.cuh -- the register-tiled FP32 field evaluation shared by the large-state kernels (D = 16 / 32 / 64):
// fixed-grid steppers (xde_tile.cu) and the adaptive Dormand-Prince stepper (xde_tile_adaptive.cu).
// A CTA of 256 threads keeps W1, W2 resident in shared memory and evaluates the MLP for a tile of TM
// trajectories; see xde_tile.cu for the design notes.
#pragma once

#include "xde_common.cuh"

namespace xde {

constexpr int kTileThreads = 256;

__device__ __forceinline__ float pre_rt(int pre, float y) {
  if (pre == XDE_PRE_CUBE) return (y * y) * y;
  if (pre == XDE_PRE_SQUARE) return y * y;
  return y;
}

template <int D, int H, int TM, int R1, int C1, int R2, int C2>
struct TileGeom {
  static constexpr int NRG1 = TM / R1, NCG1 = H / C1;
  static constexpr int NRG2 = TM / R2, NCG2 = D / C2;
  static_assert(NRG1 * NCG1 == kTileThreads, "layer-1 thread grid must cover the CTA");
  static_assert(NRG2 * NCG2 * 2 == kTileThreads, "layer-2 thread grid (x2 hidden parities) must cover the CTA");
  static_assert(R1 == 4 && R2 == 2 && C1 % 4 == 0 && C2 % 4 == 0 && NRG2 % 16 == 0, "operand fetch widths");
  static constexpr int net_floats = D * H + H * D + H + ((D + 3) / 4) * 4;
  // sU (x2: drift / diffusion pre-activations), sH
  static constexpr int act_floats = 2 * D * TM + H * TM;
  static size_t bytes(int nets, int T) { return sizeof(float) * ((size_t)nets * net_floats + act_floats + ((T + 3) / 4) * 4); }
};

template <int D, int H>
__device__ __forceinline__ void load_net(float *s, const xde_mlp_field_t &f) {
  float *w1 = s, *w2 = s + D * H, *b1 = w2 + H * D, *b2 = b1 + H;
  for (int i = threadIdx.x; i < D * H; i += blockDim.x) {
    w1[i] = f.w1[i];
    w2[i] = f.w2[i];
  }
  for (int i = threadIdx.x; i < H; i += blockDim.x) b1[i] = f.b1[i];
  for (int i = threadIdx.x; i < D; i += blockDim.x) b2[i] = f.b2[i];
}

// One field evaluation for the CTA's TM trajectories.  sU holds pre(y) k-major [D][TM]; the result
// F[R2][C2] (rows rg2*R2.., columns cg2*C2..) is returned in registers, identical on both hidden
// parities.  Two __syncthreads: the caller has synchronised after writing sU.
template <int D, int H, int TM, int R1, int C1, int R2, int C2>
__device__ __forceinline__ void tile_eval(const float *__restrict__ net, const float *__restrict__ sU,
                                          float *__restrict__ sH, float (&F)[R2][C2]) {
  using G = TileGeom<D, H, TM, R1, C1, R2, C2>;
  const float *sW1 = net, *sW2 = net + D * H, *sb1 = sW2 + H * D, *sb2 = sb1 + H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- layer 1: Z = U W1, sequential-k fma chain per element ----
  {
    const int rg = tid % G::NRG1, cg = tid / G::NRG1;
    f32x2 acc[R1][C1 / 2];
#pragma unroll
    for (int r = 0; r < R1; ++r)
#pragma unroll
      for (int c = 0; c < C1 / 2; ++c) acc[r][c] = pk1(0.0f);
    const float4 *u4 = reinterpret_cast<const float4 *>(sU + rg * R1);
    const float4 *w4 = reinterpret_cast<const float4 *>(sW1 + cg * C1);
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
      const float4 uv = u4[k * (TM / 4)];
      const float ur[4] = {uv.x, uv.y, uv.z, uv.w};
      f32x2 w[C1 / 2];
#pragma unroll
      for (int q = 0; q < C1 / 4; ++q) {
        const float4 wv = w4[k * (H / 4) + q];
        w[2 * q] = pk(wv.x, wv.y);
        w[2 * q + 1] = pk(wv.z, wv.w);
      }
#pragma unroll
      for (int r = 0; r < R1; ++r)
#pragma unroll
        for (int c = 0; c < C1 / 2; ++c) acc[r][c] = fma2(pk1(ur[r]), w[c], acc[r][c]);
    }
    // bias, tanh, transpose into sH[hidden][trajectory]
#pragma unroll
    for (int c = 0; c < C1 / 2; ++c) {
      const int j = cg * C1 + 2 * c;
      const f32x2 b = pk(sb1[j], sb1[j + 1]);
      float h0[R1], h1[R1];
#pragma unroll
      for (int r = 0; r < R1; ++r) upk(tanh_rat2(add2(acc[r][c], b)), h0[r], h1[r]);
      *reinterpret_cast<float4 *>(sH + j * TM + rg * R1) = make_float4(h0[0], h0[1], h0[2], h0[3]);
      *reinterpret_cast<float4 *>(sH + (j + 1) * TM + rg * R1) = make_float4(h1[0], h1[1], h1[2], h1[3]);
    }
  }
  __syncthreads();
  // ---- layer 2: F = Hh W2; lanes 0-15 of a warp take the even hidden units, lanes 16-31 the odd ----
  {
    const int e = lane >> 4;
    const int p = warp * 16 + (lane & 15);
    const int rg2 = p % G::NRG2, cg2 = p / G::NRG2;
    f32x2 acc[R2][C2 / 2];
#pragma unroll
    for (int r = 0; r < R2; ++r)
#pragma unroll
      for (int c = 0; c < C2 / 2; ++c) acc[r][c] = pk1(0.0f);
    const float2 *h2 = reinterpret_cast<const float2 *>(sH + rg2 * R2);
    const float4 *w4 = reinterpret_cast<const float4 *>(sW2 + cg2 * C2);
#pragma unroll 4
    for (int j = e; j < H; j += 2) {
      const float2 hv = h2[j * (TM / 2)];
      f32x2 w[C2 / 2];
#pragma unroll
      for (int q = 0; q < C2 / 4; ++q) {
        const float4 wv = w4[j * (D / 4) + q];
        w[2 * q] = pk(wv.x, wv.y);
        w[2 * q + 1] = pk(wv.z, wv.w);
      }
#pragma unroll
      for (int c = 0; c < C2 / 2; ++c) {
        acc[0][c] = fma2(pk1(hv.x), w[c], acc[0][c]);
        acc[1][c] = fma2(pk1(hv.y), w[c], acc[1][c]);
      }
    }
#pragma unroll
    for (int r = 0; r < R2; ++r)
#pragma unroll
      for (int c = 0; c < C2 / 2; ++c) {
        const f32x2 other = __shfl_xor_sync(XDE_FULL_MASK, acc[r][c], 16);
        float s0, s1;
        upk(add2(acc[r][c], other), s0, s1);  // even chain + odd chain (commutative: same bits on both halves)
        F[r][2 * c] = s0 + sb2[cg2 * C2 + 2 * c];
        F[r][2 * c + 1] = s1 + sb2[cg2 * C2 + 2 * c + 1];
      }
  }
}

}  // namespace xde
