// xde_sde_adj.cu -- sdeint_adjoint backward for small states (D <= 8), SURVEY 8(f) rank 4.
//
// The reference's SdeintAdjointMethod.backward (functional/sdeint_adjoint.py:57-230) is a copy of the ODE
// adjoint whose `augmented_diffusion` repeats the drift dynamics verbatim (:136-171) on top of the
// uninstantiable BaseSDE: there is no behaviour to reproduce (parity unpinned).  What it is reaching for on
// the solver's fixed grid is implemented: the EXACT adjoint of the Euler-Maruyama recursion
//     y[n+1] = (y[n] + f(y[n]) dt_n) + g(y[n]) * dW_n
// (discretise, then differentiate -- the gradient of what sdeint actually computed):
//     lam[n]     = lam[n+1] + J_f(y[n])^T (lam[n+1] dt_n) + J_g(y[n])^T (lam[n+1] * dW_n) + grad_y[n]
//     g_theta_f += (df/dtheta)(y[n])^T (lam[n+1] dt_n),   g_theta_g += (dg/dtheta)(y[n])^T (lam[n+1] * dW_n)
// with y[n] read back from the stored forward solution and dW from the same source as the forward pass
// (caller's table, or the counter-based generator: nothing has to be stored for the backward pass).
//
// One thread per trajectory, reverse loop over the grid.  Parameter gradients use the adjoint kernel's
// transposed fold (xde_dopri5_adj.cu): every lane parks (h_j, dz_j) of its evaluation as a column of a per-warp
// shared tile, then lane l folds all 32 columns into lane-private accumulators of "its" hidden units
// (l, l+32, ...); fp32 per 32-trajectory tile, fp64 atomics across tiles.  The adjoint state follows the
// oracle's operation order (orc_mlp_vjp) and is bit-exact; gradients agree to fp32 summation order.
#include "xde_common.cuh"

namespace xde {

constexpr int kSaWarps = 4, kSaThreads = kSaWarps * 32;
constexpr int kSaStride = 33;  // padded tile row: writes by lane (column) and reads by unit are conflict-light

struct SdeAdjParams {
  xde_mlp_field_t f, g;
  const float *t_span, *y_all, *grad_y;
  BmSource bm;
  long long B;
  int T;
  double *acc;    // [Pf + Pg]: (gW1, gb1, gW2, gb2) of drift, then of diffusion
  float *adj_y0;  // [B, D] or nullptr
};

__device__ __forceinline__ float pre_rt_(int pre, float y) {
  if (pre == XDE_PRE_CUBE) return (y * y) * y;
  if (pre == XDE_PRE_SQUARE) return y * y;
  return y;
}
__device__ __forceinline__ float pre_grad_rt_(int pre, float y) {
  if (pre == XDE_PRE_CUBE) return 3.0f * (y * y);
  if (pre == XDE_PRE_SQUARE) return 2.0f * y;
  return 1.0f;
}

// field + VJP of one network for this lane's trajectory (cotangent c), the oracle's orc_mlp_vjp operation for
// operation; (h_j, dz_j) go to column `lane` of the warp's tile; returns dy = J^T c
template <int D>
__device__ __forceinline__ void net_vjp(const float *__restrict__ sw, int H, int pre, const float (&y)[D],
                                        const float (&c)[D], float2 *__restrict__ tile, int lane, float (&u)[D],
                                        float (&dy)[D]) {
  constexpr int REC = SmallRec<D>::REC;
  const int NP = SmallRec<D>::pairs(H);
  f32x2 pdu[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    u[k] = pre_rt_(pre, y[k]);
    pdu[k] = pk1(0.0f);
  }
  (void)REC;
  for (int jp = 0; jp < NP; ++jp) {
    f32x2 w1p[D], b1p, w2p[D];
    read_pair_rec<D>(sw, jp, w1p, b1p, w2p);
    f32x2 z = first_layer_seed<D>(u[0], w1p[0]);
#pragma unroll
    for (int k = 1; k < D; ++k) z = fma2(pk1(u[k]), w1p[k], z);
    const f32x2 h = tanh_rat2(add2(z, b1p));
    f32x2 dh = mul2(pk1(c[0]), w2p[0]);
#pragma unroll
    for (int d = 1; d < D; ++d) dh = fma2(pk1(c[d]), w2p[d], dh);
    const f32x2 dz = mul2(dh, one_minus_sq2(h));
#pragma unroll
    for (int k = 0; k < D; ++k) pdu[k] = fma2(dz, w1p[k], pdu[k]);
    float h0, h1, z0, z1;
    upk(h, h0, h1);
    upk(dz, z0, z1);
    tile[(2 * jp) * kSaStride + lane] = make_float2(h0, z0);
    tile[(2 * jp + 1) * kSaStride + lane] = make_float2(h1, z1);  // a padded odd unit is (0, 0)
  }
#pragma unroll
  for (int k = 0; k < D; ++k) {
    float e, o;
    upk(pdu[k], e, o);
    dy[k] = (e + o) * pre_grad_rt_(pre, y[k]);
  }
}

template <int D, int HU>
__global__ void __launch_bounds__(kSaThreads) sde_adj_small_kernel(const SdeAdjParams p) {
  extern __shared__ __align__(16) float smem[];
  const int Hf = p.f.h, Hg = p.g.h;
  const int Hf2 = 2 * SmallRec<D>::pairs(Hf), Hg2 = 2 * SmallRec<D>::pairs(Hg);  // rows incl. the padded unit
  float *swf = smem;
  float *swg = swf + SmallRec<D>::floats(Hf);
  float *st = swg + SmallRec<D>::floats(Hg);
  float *wbase = st + ((p.T + 3) / 4) * 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t per_warp = (size_t)2 * (Hf2 + Hg2) * kSaStride + 32 * 4 * D;  // floats
  float2 *tf = reinterpret_cast<float2 *>(wbase + warp * per_warp);
  float2 *tg = tf + (size_t)Hf2 * kSaStride;
  float *coef = reinterpret_cast<float *>(tg + (size_t)Hg2 * kSaStride);  // [32][4D]: cf | cg | uf | ug
  load_small_field<D>(swf, p.f);
  load_small_field<D>(swg, p.g);
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) st[i] = p.t_span[i];
  __syncthreads();

  const long long Pf = 2LL * D * Hf + Hf + D;
  const long long n_tiles = (p.B + 31) / 32;
  for (long long tile = (long long)blockIdx.x * kSaWarps + warp; tile < n_tiles; tile += (long long)gridDim.x * kSaWarps) {
    const long long b = tile * 32 + lane;
    const bool ok = b < p.B;
    const float *yb = p.y_all + b * (long long)p.T * D;
    const float *gb = p.grad_y + b * (long long)p.T * D;
    float lam[D], sb2f[D], sb2g[D];
#pragma unroll
    for (int e = 0; e < D; ++e) {
      lam[e] = ok ? gb[(long long)(p.T - 1) * D + e] : 0.0f;
      sb2f[e] = sb2g[e] = 0.0f;
    }
    float af[HU][2 * D + 1], ag[HU][2 * D + 1];  // per unit: gb1 | gW1[k] | gW2[d]
#pragma unroll
    for (int uu = 0; uu < HU; ++uu)
#pragma unroll
      for (int q = 0; q < 2 * D + 1; ++q) af[uu][q] = ag[uu][q] = 0.0f;

    for (int n = p.T - 2; n >= 0; --n) {
      const float dt = st[n + 1] - st[n];
      float y[D], cf[D], cg[D], uf[D], ug[D], dyf[D], dyg[D];
#pragma unroll
      for (int e = 0; e < D; ++e) {
        y[e] = ok ? yb[(long long)n * D + e] : 0.0f;
        const float w = ok ? bm_increment1(p.bm, n, b, p.B, D, e, dt) : 0.0f;
        cf[e] = lam[e] * dt;
        cg[e] = lam[e] * w;
        sb2f[e] += cf[e];
        sb2g[e] += cg[e];
      }
      net_vjp<D>(swf, Hf, p.f.pre, y, cf, tf, lane, uf, dyf);
      net_vjp<D>(swg, Hg, p.g.pre, y, cg, tg, lane, ug, dyg);
#pragma unroll
      for (int e = 0; e < D; ++e) {
        coef[lane * 4 * D + e] = cf[e];
        coef[lane * 4 * D + D + e] = cg[e];
        coef[lane * 4 * D + 2 * D + e] = uf[e];
        coef[lane * 4 * D + 3 * D + e] = ug[e];
      }
      __syncwarp();
      // transposed fold: lane l owns hidden units l, l + 32, ... of both networks
#pragma unroll
      for (int uu = 0; uu < HU; ++uu) {
        const int j = lane + 32 * uu;
        if (j < Hf) {
          for (int col = 0; col < 32; ++col) {
            const float2 hz = tf[j * kSaStride + col];
            const float *cc = coef + col * 4 * D;
            af[uu][0] += hz.y;
#pragma unroll
            for (int k = 0; k < D; ++k) af[uu][1 + k] = fmaf(cc[2 * D + k], hz.y, af[uu][1 + k]);
#pragma unroll
            for (int d = 0; d < D; ++d) af[uu][1 + D + d] = fmaf(hz.x, cc[d], af[uu][1 + D + d]);
          }
        }
        if (j < Hg) {
          for (int col = 0; col < 32; ++col) {
            const float2 hz = tg[j * kSaStride + col];
            const float *cc = coef + col * 4 * D;
            ag[uu][0] += hz.y;
#pragma unroll
            for (int k = 0; k < D; ++k) ag[uu][1 + k] = fmaf(cc[3 * D + k], hz.y, ag[uu][1 + k]);
#pragma unroll
            for (int d = 0; d < D; ++d) ag[uu][1 + D + d] = fmaf(hz.x, cc[D + d], ag[uu][1 + D + d]);
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int e = 0; e < D; ++e) lam[e] = ((lam[e] + dyf[e]) + dyg[e]) + (ok ? gb[(long long)n * D + e] : 0.0f);
    }
    if (ok && p.adj_y0)
#pragma unroll
      for (int e = 0; e < D; ++e) p.adj_y0[b * D + e] = lam[e];
    // flush this tile's partial sums: (gW1 [D,H], gb1 [H], gW2 [H,D], gb2 [D]) per network
#pragma unroll
    for (int uu = 0; uu < HU; ++uu) {
      const int j = lane + 32 * uu;
      if (j < Hf) {
        atomicAdd(&p.acc[(long long)D * Hf + j], (double)af[uu][0]);
#pragma unroll
        for (int k = 0; k < D; ++k) atomicAdd(&p.acc[(long long)k * Hf + j], (double)af[uu][1 + k]);
#pragma unroll
        for (int d = 0; d < D; ++d) atomicAdd(&p.acc[(long long)D * Hf + Hf + (long long)j * D + d], (double)af[uu][1 + D + d]);
      }
      if (j < Hg) {
        atomicAdd(&p.acc[Pf + (long long)D * Hg + j], (double)ag[uu][0]);
#pragma unroll
        for (int k = 0; k < D; ++k) atomicAdd(&p.acc[Pf + (long long)k * Hg + j], (double)ag[uu][1 + k]);
#pragma unroll
        for (int d = 0; d < D; ++d)
          atomicAdd(&p.acc[Pf + (long long)D * Hg + Hg + (long long)j * D + d], (double)ag[uu][1 + D + d]);
      }
    }
#pragma unroll
    for (int e = 0; e < D; ++e) {
      double a = (double)sb2f[e], c = (double)sb2g[e];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(XDE_FULL_MASK, a, off);
        c += __shfl_xor_sync(XDE_FULL_MASK, c, off);
      }
      if (lane == 0) {
        atomicAdd(&p.acc[2LL * D * Hf + Hf + e], a);
        atomicAdd(&p.acc[Pf + 2LL * D * Hg + Hg + e], c);
      }
    }
  }
}

__global__ void sde_adj_cast_kernel(const double *__restrict__ acc, float *__restrict__ gf, float *__restrict__ gg,
                                    long long Pf, long long Pg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Pf) gf[i] = (float)acc[i];
  else if (i < Pf + Pg) gg[i - Pf] = (float)acc[i];
}

template <int D, int HU>
static int launch_sde_adj(const SdeAdjParams &p, cudaStream_t s) {
  const int Hf2 = 2 * SmallRec<D>::pairs(p.f.h), Hg2 = 2 * SmallRec<D>::pairs(p.g.h);
  const size_t per_warp = (size_t)2 * (Hf2 + Hg2) * kSaStride + 32 * 4 * D;
  const size_t smem = sizeof(float) * (SmallRec<D>::floats(p.f.h) + SmallRec<D>::floats(p.g.h) + ((p.T + 3) / 4) * 4 +
                                       kSaWarps * per_warp);
  XDE_REQUIRE(smem <= 227 * 1024, XDE_E_UNSUPPORTED_FIELD,
              "sde adjoint: fields (H=%d, %d) + tiles need %zu bytes of shared memory", p.f.h, p.g.h, smem);
  auto kern = sde_adj_small_kernel<D, HU>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  XDE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSaThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const long long n_tiles = (p.B + 31) / 32;
  long long grid = (long long)sm_count() * per_sm;
  const long long want = (n_tiles + kSaWarps - 1) / kSaWarps;
  if (grid > want) grid = want;
  kern<<<(unsigned)grid, kSaThreads, smem, s>>>(p);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  return XDE_OK;
}

template <int D>
static int sde_adj_hu(const SdeAdjParams &p, cudaStream_t s) {
  const int hmax = p.f.h > p.g.h ? p.f.h : p.g.h;
  if (hmax <= 32) return launch_sde_adj<D, 1>(p, s);
  if (hmax <= 64) return launch_sde_adj<D, 2>(p, s);
  if (hmax <= 96) return launch_sde_adj<D, 3>(p, s);
  set_last_error("sde adjoint: hidden width %d has no fused kernel (H <= 96)", hmax);
  return XDE_E_UNSUPPORTED_FIELD;
}

}  // namespace xde

extern "C" XDE_EXPORT int xde_sde_mlp_adjoint_f32(const xde_mlp_field_t *drift, const xde_mlp_field_t *diffusion,
                                                  const float *t_span, int32_t T, const float *y_all,
                                                  const float *grad_y, int64_t B, const float *dW, uint64_t seed,
                                                  int64_t traj_offset, float *out_gdrift, float *out_gdiffusion,
                                                  float *out_adj_y0, void *stream) {
  using namespace xde;
  XDE_REQUIRE(drift && diffusion && t_span && y_all && grad_y && out_gdrift && out_gdiffusion, XDE_E_BAD_ARG,
              "null argument");
  XDE_REQUIRE(B >= 1 && T >= 1, XDE_E_BAD_ARG, "need B>=1, T>=1");
  XDE_REQUIRE(drift->d == diffusion->d, XDE_E_BAD_ARG, "drift and diffusion state dims differ");
  cudaStream_t s = (cudaStream_t)stream;
  SdeAdjParams p{};
  p.f = *drift;
  p.g = *diffusion;
  p.t_span = t_span;
  p.y_all = y_all;
  p.grad_y = grad_y;
  p.bm = BmSource{dW, seed, traj_offset};
  p.B = B;
  p.T = T;
  p.adj_y0 = out_adj_y0;
  const long long Pf = 2LL * drift->d * drift->h + drift->h + drift->d;
  const long long Pg = 2LL * diffusion->d * diffusion->h + diffusion->h + diffusion->d;
  void *acc = nullptr;
  XDE_CUDA_CHECK(scratch_alloc(&acc, sizeof(double) * (size_t)(Pf + Pg), s));
  XDE_CUDA_CHECK(cudaMemsetAsync(acc, 0, sizeof(double) * (size_t)(Pf + Pg), s));
  p.acc = (double *)acc;
  int rc;
  switch (drift->d) {
    case 1: rc = sde_adj_hu<1>(p, s); break;
    case 2: rc = sde_adj_hu<2>(p, s); break;
    case 3: rc = sde_adj_hu<3>(p, s); break;
    case 4: rc = sde_adj_hu<4>(p, s); break;
    case 8: rc = sde_adj_hu<8>(p, s); break;
    default:
      set_last_error("sde adjoint: state dim D=%d has no fused kernel (supported: 1,2,3,4,8)", drift->d);
      rc = XDE_E_UNSUPPORTED_FIELD;
  }
  if (rc == XDE_OK) {
    sde_adj_cast_kernel<<<(unsigned)((Pf + Pg + 255) / 256), 256, 0, s>>>((const double *)acc, out_gdrift,
                                                                         out_gdiffusion, Pf, Pg);
    count_launch();
  }
  XDE_CUDA_CHECK(cudaFreeAsync(acc, s));
  if (rc == XDE_OK) XDE_CUDA_CHECK(cudaGetLastError());
  return rc;
}
