// xde_tile.cu -- fixed-grid steppers for LARGE states (D = 32 / 64, H = 64 / 128 / 256): cfg3 (ensemble
// RK4, MLP 64-256-64) and cfg4 (Euler-Maruyama, two 32-64-32 nets).
//   * odeint(..., solver=Euler|RK4): FixedSolver.integrate (solver/base_fixed_solver.py:103-144),
//     Euler.step (fixed_solver/euler.py:7-11), RK4.step = 3/8 rule (base_fixed_solver.py:166-197);
//   * sdeint(..., solver=Euler): y1 = y0 + f*dt + g*dW with caller-supplied dW (xde/base_sde.py:44-61).
//
// Design.  At these sizes one trajectory no longer fits a thread, and the field is GEMM-shaped:
//   Z[TM x H] = U[TM x D] W1[D x H],  F[TM x D] = tanh(Z + b1)[TM x H] W2[H x D].
// A CTA (256 threads) keeps BOTH weight matrices resident in shared memory for the whole solve
// (64-256-64: 128 KB) and integrates tiles of TM trajectories over the entire time grid; state and
// RK stages live in registers, only the stage input U and the hidden activations pass through
// shared memory (k-major, so every operand fetch is a conflict-free LDS.128/LDS.64).  The GEMMs are
// register-tiled FP32 with Blackwell's packed FFMA2 (two IEEE FMAs per issue slot): layer 1 gives
// each thread an R1 x C1 block of Z, layer 2 splits the hidden axis between the two half-warps
// (even / odd hidden units -- exactly the two-chain reduction order of the arithmetic
// specification) and joins them with one shuffle.  Results are bit-identical to the oracle.
// Tensor cores (tcgen05, 3xTF32) are the next step for this kernel (DESIGN.md "Next"): they
// cannot be bit-exact against an fp32 oracle, so this FP32 path is also their reference.
#include "xde_tile.cuh"

namespace xde {

struct TileParams {
  xde_mlp_field_t f, g;
  const float *y0, *t_span;
  BmSource bm;
  float *out;
  long long B;
  int T, stride, n_out;
};

// KIND 0: ODE Euler, 1: ODE RK4 (3/8 rule), 2: SDE Euler-Maruyama (increments from the caller's table),
// 3: ODE Midpoint (fixed_solver/midpoint.py:7-18), 4: SDE Euler-Maruyama, increments generated (BmSource),
// 5 / 6: SDE Milstein with table / generated increments (north_star (4); no reference counterpart: the analytic
//        diagonal of the diffusion Jacobian, oracle mlp_eval_diag_jac)
template <int D, int H, int TM, int R1, int C1, int R2, int C2, int KIND>
__global__ void __launch_bounds__(kTileThreads, 1) fixed_tile_kernel(const TileParams p) {
  using G = TileGeom<D, H, TM, R1, C1, R2, C2>;
  extern __shared__ __align__(16) float smem[];
  constexpr bool SDE = (KIND == 2 || KIND == 4 || KIND == 5 || KIND == 6);
  constexpr bool MILSTEIN = (KIND == 5 || KIND == 6);
  constexpr bool GEN = (KIND == 4 || KIND == 6);
  constexpr int NETS = SDE ? 2 : 1;
  float *netf = smem;
  float *netg = smem + G::net_floats;  // only when NETS == 2
  float *sU = smem + NETS * G::net_floats;
  float *sUg = sU + D * TM;
  float *sH = sUg + D * TM;
  float *st = sH + H * TM;
  load_net<D, H>(netf, p.f);
  if (SDE) load_net<D, H>(netg, p.g);
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) st[i] = p.t_span[i];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = lane >> 4;
  const int pidx = warp * 16 + (lane & 15);
  const int rg2 = pidx % G::NRG2, cg2 = pidx / G::NRG2;
  const int c0 = cg2 * C2;
  const float one_third = (float)(1.0 / 3.0);
  const int pref = p.f.pre, preg = p.g.pre;
  const long long n_tiles = (p.B + TM - 1) / TM;

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long b0 = tile * TM + rg2 * R2;  // first of this thread's R2 trajectories
    float y[R2][C2];
#pragma unroll
    for (int r = 0; r < R2; ++r) {
      const long long b = b0 + r;
      const bool ok = b < p.B;
#pragma unroll
      for (int q = 0; q < C2 / 4; ++q) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = *reinterpret_cast<const float4 *>(p.y0 + b * D + c0 + 4 * q);
        y[r][4 * q] = v.x;
        y[r][4 * q + 1] = v.y;
        y[r][4 * q + 2] = v.z;
        y[r][4 * q + 3] = v.w;
        if (ok && r == e) *reinterpret_cast<float4 *>(p.out + b * (long long)p.n_out * D + c0 + 4 * q) = v;
      }
    }
    // publish a stage input: the half-warp with parity e writes row e of its pair
    auto put_u = [&](float *dst, int pre, const float (&v)[R2][C2]) {
#pragma unroll
      for (int c = 0; c < C2; ++c) dst[(c0 + c) * TM + rg2 * R2 + e] = pre_rt(pre, e ? v[1][c] : v[0][c]);
    };
    for (int i = 1; i < p.T; ++i) {
      const float t0 = st[i - 1], t1 = st[i];
      const float dt = t1 - t0;
      float k1[R2][C2];
      // no barrier needed before put_u: every warp has passed the layer-1/layer-2 barrier of the previous
      // evaluation (nobody reads sU any more); the barrier below also fences the previous readers of sH
      put_u(sU, pref, y);
      if (SDE) put_u(sUg, preg, y);
      __syncthreads();
      tile_eval<D, H, TM, R1, C1, R2, C2>(netf, sU, sH, k1);
      if (KIND == 0) {
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) y[r][c] = k1[r][c] * dt + y[r][c];
      } else if (KIND == 3) {
        float k2[R2][C2], yi[R2][C2];
        const float half_dt = 0.5f * dt;
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) yi[r][c] = k1[r][c] * half_dt + y[r][c];
        put_u(sU, pref, yi);
        __syncthreads();
        tile_eval<D, H, TM, R1, C1, R2, C2>(netf, sU, sH, k2);
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) y[r][c] = k2[r][c] * dt + y[r][c];
      } else if (KIND == 1) {
        float k2[R2][C2], k3[R2][C2], k4[R2][C2], yi[R2][C2];
        const float dt13 = dt * one_third;
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) yi[r][c] = k1[r][c] * dt13 + y[r][c];
        put_u(sU, pref, yi);
        __syncthreads();
        tile_eval<D, H, TM, R1, C1, R2, C2>(netf, sU, sH, k2);
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) yi[r][c] = (k1[r][c] - k2[r][c] * one_third) * dt + y[r][c];
        put_u(sU, pref, yi);
        __syncthreads();
        tile_eval<D, H, TM, R1, C1, R2, C2>(netf, sU, sH, k3);
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) yi[r][c] = ((k1[r][c] - k2[r][c]) + k3[r][c]) * dt + y[r][c];
        put_u(sU, pref, yi);
        __syncthreads();
        tile_eval<D, H, TM, R1, C1, R2, C2>(netf, sU, sH, k4);
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) {
            const float a = k1[r][c] * dt + y[r][c];
            const float bb = k2[r][c] * dt + y[r][c];
            const float cc = k3[r][c] * dt + y[r][c];
            const float dd = k4[r][c] * dt + y[r][c];
            y[r][c] = (((a + 3.0f * bb) + 3.0f * cc) + dd) * 0.125f;
          }
      } else {
        float g[R2][C2];
        __syncthreads();  // sH is reused by the second network
        tile_eval<D, H, TM, R1, C1, R2, C2>(netg, sUg, sH, g);
        float gp[R2][C2];
        if (MILSTEIN) {
          // d g_d / d y_d = pre'(y_d) * sum_j ((1 - h_j^2) W1[d][j]) W2[j][d]: the product with W1 is rounded first, then
          // the two-chain reduction over the hidden axis (oracle mlp_eval_diag_jac); sH still holds h of the diffusion net
          const float *gW1 = netg, *gW2 = netg + D * H;
          f32x2 acc[R2][C2 / 2];
#pragma unroll
          for (int r = 0; r < R2; ++r)
#pragma unroll
            for (int c = 0; c < C2 / 2; ++c) acc[r][c] = pk1(0.0f);
          const float2 *h2 = reinterpret_cast<const float2 *>(sH + rg2 * R2);
          const float4 *w4 = reinterpret_cast<const float4 *>(gW2 + c0);
#pragma unroll 2
          for (int j = e; j < H; j += 2) {
            const float2 hv = h2[j * (TM / 2)];
            const float s0 = fmaf(-hv.x, hv.x, 1.0f), s1 = fmaf(-hv.y, hv.y, 1.0f);
            f32x2 w2v[C2 / 2];
#pragma unroll
            for (int q = 0; q < C2 / 4; ++q) {
              const float4 wv = w4[j * (D / 4) + q];
              w2v[2 * q] = pk(wv.x, wv.y);
              w2v[2 * q + 1] = pk(wv.z, wv.w);
            }
#pragma unroll
            for (int c = 0; c < C2 / 2; ++c) {
              const f32x2 w1v = pk(gW1[(c0 + 2 * c) * H + j], gW1[(c0 + 2 * c + 1) * H + j]);
              acc[0][c] = fma2(mul2(pk1(s0), w1v), w2v[c], acc[0][c]);
              acc[1][c] = fma2(mul2(pk1(s1), w1v), w2v[c], acc[1][c]);
            }
          }
#pragma unroll
          for (int r = 0; r < R2; ++r)
#pragma unroll
            for (int c = 0; c < C2 / 2; ++c) {
              const f32x2 other = __shfl_xor_sync(XDE_FULL_MASK, acc[r][c], 16);
              float j0, j1;
              upk(add2(acc[r][c], other), j0, j1);
              const float y0v = y[r][2 * c], y1v = y[r][2 * c + 1];
              gp[r][2 * c] = j0 * (preg == XDE_PRE_CUBE ? 3.0f * (y0v * y0v) : (preg == XDE_PRE_SQUARE ? 2.0f * y0v : 1.0f));
              gp[r][2 * c + 1] = j1 * (preg == XDE_PRE_CUBE ? 3.0f * (y1v * y1v) : (preg == XDE_PRE_SQUARE ? 2.0f * y1v : 1.0f));
            }
        }
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          const long long b = b0 + r;
          const bool ok = b < p.B;
#pragma unroll
          for (int q = 0; q < C2 / 4; ++q) {
            float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) wv = bm_increment4<GEN>(p.bm, i - 1, b, p.B, D, c0 / 4 + q, dt);
            const float w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int z = 0; z < 4; ++z) {
              const int c = 4 * q + z;
              float v = (y[r][c] + k1[r][c] * dt) + g[r][c] * w[z];
              if (MILSTEIN) v = v + ((0.5f * g[r][c]) * gp[r][c]) * (w[z] * w[z] - dt);
              y[r][c] = v;
            }
          }
        }
      }
      // linear_interp at t == t1 is the identity (interpolation/functional/interp_fn.py:4-10)
      if (i % p.stride == 0 || i == p.T - 1) {
        const int row = (i == p.T - 1) ? p.n_out - 1 : i / p.stride;
        const long long b = b0 + e;
        if (b < p.B) {
#pragma unroll
          for (int q = 0; q < C2 / 4; ++q) {
            const float4 v = e ? make_float4(y[1][4 * q], y[1][4 * q + 1], y[1][4 * q + 2], y[1][4 * q + 3])
                               : make_float4(y[0][4 * q], y[0][4 * q + 1], y[0][4 * q + 2], y[0][4 * q + 3]);
            *reinterpret_cast<float4 *>(p.out + (b * (long long)p.n_out + row) * D + c0 + 4 * q) = v;
          }
        }
      }
    }
  }
}

template <int D, int H, int TM, int R1, int C1, int R2, int C2, int KIND>
static int launch_tile(const TileParams &p, cudaStream_t s) {
  using G = TileGeom<D, H, TM, R1, C1, R2, C2>;
  const size_t smem = G::bytes((KIND == 2 || KIND >= 4) ? 2 : 1, p.T);
  XDE_REQUIRE(smem <= 227 * 1024, XDE_E_UNSUPPORTED_FIELD,
              "tiled solver: weights + tiles + grid need %zu bytes of shared memory (> 227 KB)", smem);
  auto kern = fixed_tile_kernel<D, H, TM, R1, C1, R2, C2, KIND>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  XDE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, smem));
  if (per_sm < 1) per_sm = 1;
  long long n_tiles = (p.B + TM - 1) / TM;
  long long grid = (long long)sm_count() * per_sm;  // persistent: a whole number of CTAs per SM
  if (grid > n_tiles) grid = n_tiles;
  kern<<<(unsigned)grid, kTileThreads, smem, s>>>(p);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  return XDE_OK;
}

template <int KIND>
static int tile_dispatch(const TileParams &p, cudaStream_t s) {
  const int D = p.f.d, H = p.f.h;
  if (D == 64 && H == 256) return launch_tile<64, 256, 32, 4, 8, 2, 8, KIND>(p, s);
  if (D == 64 && H == 128) return launch_tile<64, 128, 32, 4, 4, 2, 8, KIND>(p, s);
  if (D == 64 && H == 64) return launch_tile<64, 64, 64, 4, 4, 2, 16, KIND>(p, s);
  if (D == 32 && H == 256) return launch_tile<32, 256, 32, 4, 8, 2, 4, KIND>(p, s);
  if (D == 32 && H == 128) return launch_tile<32, 128, 64, 4, 8, 2, 8, KIND>(p, s);
  if (D == 32 && H == 64) return launch_tile<32, 64, 64, 4, 4, 2, 8, KIND>(p, s);
  if (D == 16 && H == 64) return launch_tile<16, 64, 64, 4, 4, 2, 4, KIND>(p, s);
  set_last_error("tiled solver: no kernel for D=%d H=%d (D in {16,32,64} x H in {64,128,256}; small states D<=8 any H)", D, H);
  return XDE_E_UNSUPPORTED_FIELD;
}

bool tile_covers(int D) { return D >= 16; }

int rk_fixed_tile(int method, const xde_mlp_field_t *f, const float *y0, long long B, const float *t_span,
                  int T, int stride, float *out, cudaStream_t s) {
  TileParams p{};
  p.f = *f;
  p.g = *f;
  p.y0 = y0;
  p.t_span = t_span;
  p.out = out;
  p.B = B;
  p.T = T;
  p.stride = stride;
  p.n_out = (T - 1 + stride - 1) / stride + 1;
  if (method == XDE_FIXED_EULER) return tile_dispatch<0>(p, s);
  if (method == XDE_FIXED_MIDPOINT) return tile_dispatch<3>(p, s);
  return tile_dispatch<1>(p, s);
}

int sde_tile(int scheme, const xde_mlp_field_t *f, const xde_mlp_field_t *g, const float *y0, long long B,
             const float *t_span, int T, const BmSource &bm, int stride, float *out, cudaStream_t s) {
  XDE_REQUIRE(scheme == XDE_SDE_EM || scheme == XDE_SDE_MILSTEIN, XDE_E_BAD_ARG, "unknown SDE scheme %d", scheme);
  XDE_REQUIRE(f->h == g->h, XDE_E_UNSUPPORTED_FIELD, "tiled sde: drift and diffusion must share the hidden width");
  TileParams p{};
  p.f = *f;
  p.g = *g;
  p.y0 = y0;
  p.t_span = t_span;
  p.bm = bm;
  p.out = out;
  p.B = B;
  p.T = T;
  p.stride = stride;
  p.n_out = (T - 1 + stride - 1) / stride + 1;
  if (scheme == XDE_SDE_MILSTEIN) return bm.table ? tile_dispatch<5>(p, s) : tile_dispatch<6>(p, s);
  return bm.table ? tile_dispatch<2>(p, s) : tile_dispatch<4>(p, s);
}

}  // namespace xde
