// xde_fixed_small.cu -- fixed-grid steppers for small states (D <= 8), one thread per trajectory:
//   * odeint(..., solver=Euler|RK4|Midpoint): FixedSolver.integrate (solver/base_fixed_solver.py:103-144),
//     Euler.step (fixed_solver/euler.py:7-11), RK4.step = 3/8 rule (base_fixed_solver.py:166-197),
//     BaseODE.fuse = dy*dt + y0 (xde/base_ode.py:58);
//   * sdeint(..., solver=Euler): Euler-Maruyama y1 = y0 + f*dt + g*dW with caller-supplied dW
//     (xde/base_sde.py:44-61, repairs R2/R3); Milstein as an extension.
// The whole grid is integrated in one launch; state and stages stay in registers, the MLP weights
// in shared memory.  Large states (D >= 16) are served by the tiled kernels in xde_tile.cu.
#include "xde_common.cuh"

namespace xde {

constexpr int kFixThreads = 128;

struct FixParams {
  xde_mlp_field_t f, g;
  const float *y0, *t_span;
  BmSource bm;
  float *out;
  long long B;
  int T, stride, n_out, method;
};

template <int D, int PRE, int METHOD>
__global__ void __launch_bounds__(kFixThreads) rk_fixed_small_kernel(const FixParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.f.h;
  float *sw = smem;
  float *st = smem + SmallRec<D>::floats(H);
  load_small_field<D>(sw, p.f);
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) st[i] = p.t_span[i];
  __syncthreads();
  const float one_third = (float)(1.0 / 3.0);
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < p.B;
       b += (long long)gridDim.x * blockDim.x) {
    float y[D];
    float *o = p.out + b * p.n_out * D;
#pragma unroll
    for (int e = 0; e < D; ++e) {
      y[e] = p.y0[b * D + e];
      o[e] = y[e];
    }
    for (int i = 1; i < p.T; ++i) {
      const float t0 = st[i - 1], t1 = st[i];
      const float dt = t1 - t0;
      float k1[D];
      mlp_eval_small<D, PRE>(sw, H, y, k1);
      if (METHOD == XDE_FIXED_EULER) {
#pragma unroll
        for (int e = 0; e < D; ++e) y[e] = k1[e] * dt + y[e];
      } else if (METHOD == XDE_FIXED_MIDPOINT) {
        // fixed_solver/midpoint.py:7-18: y_half = fuse(f(y0), dt/2, y0); y1 = fuse(f(y_half), dt, y0)
        float k2[D], yi[D];
        const float half_dt = 0.5f * dt;
#pragma unroll
        for (int e = 0; e < D; ++e) yi[e] = k1[e] * half_dt + y[e];
        mlp_eval_small<D, PRE>(sw, H, yi, k2);
#pragma unroll
        for (int e = 0; e < D; ++e) y[e] = k2[e] * dt + y[e];
      } else {
        float k2[D], k3[D], k4[D], yi[D];
        const float dt13 = dt * one_third;
#pragma unroll
        for (int e = 0; e < D; ++e) yi[e] = k1[e] * dt13 + y[e];
        mlp_eval_small<D, PRE>(sw, H, yi, k2);
#pragma unroll
        for (int e = 0; e < D; ++e) yi[e] = (k1[e] - k2[e] * one_third) * dt + y[e];
        mlp_eval_small<D, PRE>(sw, H, yi, k3);
#pragma unroll
        for (int e = 0; e < D; ++e) yi[e] = ((k1[e] - k2[e]) + k3[e]) * dt + y[e];
        mlp_eval_small<D, PRE>(sw, H, yi, k4);
#pragma unroll
        for (int e = 0; e < D; ++e) {
          const float a = k1[e] * dt + y[e];
          const float bb = k2[e] * dt + y[e];
          const float c = k3[e] * dt + y[e];
          const float d = k4[e] * dt + y[e];
          y[e] = (((a + 3.0f * bb) + 3.0f * c) + d) * 0.125f;
        }
      }
      // linear_interp at t == t1 is the identity (interpolation/functional/interp_fn.py:4-10)
      if (i % p.stride == 0 || i == p.T - 1) {
        const int row = (i == p.T - 1) ? p.n_out - 1 : i / p.stride;
#pragma unroll
        for (int e = 0; e < D; ++e) o[row * D + e] = y[e];
      }
    }
  }
}

// diffusion value + analytic diagonal of its Jacobian (Milstein)
template <int D, int PRE>
__device__ __forceinline__ void mlp_eval_diag_small(const float *__restrict__ sw, int H, const float (&y)[D],
                                                    float (&g)[D], float (&gp)[D]) {
  constexpr int REC = SmallRec<D>::REC;
  const int NP = SmallRec<D>::pairs(H);
  float u[D];
  f32x2 acc[D], jac[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    u[k] = pre_act<PRE>(y[k]);
    acc[k] = pk1(0.0f);
    jac[k] = pk1(0.0f);
  }
  for (int jp = 0; jp < NP; ++jp) {
    f32x2 w1p[D], b1p, w2p[D];
    read_pair_rec<D>(sw, jp, w1p, b1p, w2p);
    f32x2 z = first_layer_seed<D>(u[0], w1p[0]);
#pragma unroll
    for (int k = 1; k < D; ++k) z = fma2(pk1(u[k]), w1p[k], z);
    const f32x2 h = tanh_rat2(add2(z, b1p));
    const f32x2 s = one_minus_sq2(h);
#pragma unroll
    for (int d = 0; d < D; ++d) {
      acc[d] = fma2(h, w2p[d], acc[d]);
      jac[d] = fma2(mul2(s, w1p[d]), w2p[d], jac[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    float e, o;
    upk(acc[d], e, o);
    g[d] = (e + o) + sw[NP * REC + d];
    upk(jac[d], e, o);
    gp[d] = (e + o) * pre_act_grad<PRE>(y[d]);
  }
}

template <int D, int PREF, int PREG, int SCHEME>
__global__ void __launch_bounds__(kFixThreads) sde_small_kernel(const FixParams p) {
  extern __shared__ __align__(16) float smem[];
  float *swf = smem;
  float *swg = swf + SmallRec<D>::floats(p.f.h);
  float *st = swg + SmallRec<D>::floats(p.g.h);
  load_small_field<D>(swf, p.f);
  load_small_field<D>(swg, p.g);
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) st[i] = p.t_span[i];
  __syncthreads();
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < p.B;
       b += (long long)gridDim.x * blockDim.x) {
    float y[D];
    float *o = p.out + b * p.n_out * D;
#pragma unroll
    for (int e = 0; e < D; ++e) {
      y[e] = p.y0[b * D + e];
      o[e] = y[e];
    }
    for (int i = 1; i < p.T; ++i) {
      const float dt = st[i] - st[i - 1];
      float w[D], f[D], g[D], gp[D];
#pragma unroll
      for (int e = 0; e < D; ++e) w[e] = bm_increment1(p.bm, i - 1, b, p.B, D, e, dt);
      mlp_eval_small<D, PREF>(swf, p.f.h, y, f);
      if (SCHEME == XDE_SDE_MILSTEIN)
        mlp_eval_diag_small<D, PREG>(swg, p.g.h, y, g, gp);
      else
        mlp_eval_small<D, PREG>(swg, p.g.h, y, g);
#pragma unroll
      for (int e = 0; e < D; ++e) {
        float v = (y[e] + f[e] * dt) + g[e] * w[e];
        if (SCHEME == XDE_SDE_MILSTEIN) v = v + ((0.5f * g[e]) * gp[e]) * (w[e] * w[e] - dt);
        y[e] = v;
      }
      if (i % p.stride == 0 || i == p.T - 1) {
        const int row = (i == p.T - 1) ? p.n_out - 1 : i / p.stride;
#pragma unroll
        for (int e = 0; e < D; ++e) o[row * D + e] = y[e];
      }
    }
  }
}

static unsigned grid_for(long long B, int threads) {
  long long want = (B + threads - 1) / threads;
  long long cap = (long long)sm_count() * 8;
  if (want > cap) want = (want + cap - 1) / cap <= 1 ? want : cap;
  return (unsigned)(want < 1 ? 1 : want);
}

template <int D, int PRE, int METHOD>
static int launch_fixed(const FixParams &p, cudaStream_t s) {
  const size_t smem = sizeof(float) * (SmallRec<D>::floats(p.f.h) + p.T);
  XDE_REQUIRE(smem <= 200 * 1024, XDE_E_UNSUPPORTED_FIELD, "fixed solver: field + grid exceed shared memory");
  auto kern = rk_fixed_small_kernel<D, PRE, METHOD>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid_for(p.B, kFixThreads), kFixThreads, smem, s>>>(p);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  return XDE_OK;
}

template <int D, int PRE>
static int fixed_method(const FixParams &p, cudaStream_t s) {
  if (p.method == XDE_FIXED_EULER) return launch_fixed<D, PRE, XDE_FIXED_EULER>(p, s);
  if (p.method == XDE_FIXED_MIDPOINT) return launch_fixed<D, PRE, XDE_FIXED_MIDPOINT>(p, s);
  return launch_fixed<D, PRE, XDE_FIXED_RK4_38>(p, s);
}
template <int D>
static int fixed_pre(const FixParams &p, cudaStream_t s) {
  switch (p.f.pre) {
    case XDE_PRE_ID: return fixed_method<D, XDE_PRE_ID>(p, s);
    case XDE_PRE_SQUARE: return fixed_method<D, XDE_PRE_SQUARE>(p, s);
    case XDE_PRE_CUBE: return fixed_method<D, XDE_PRE_CUBE>(p, s);
  }
  set_last_error("unknown pre-activation %d", p.f.pre);
  return XDE_E_BAD_ARG;
}

template <int D, int PREF, int PREG, int SCHEME>
static int launch_sde(const FixParams &p, cudaStream_t s) {
  const size_t smem = sizeof(float) * (SmallRec<D>::floats(p.f.h) + SmallRec<D>::floats(p.g.h) + p.T);
  XDE_REQUIRE(smem <= 200 * 1024, XDE_E_UNSUPPORTED_FIELD, "sde: fields + grid exceed shared memory");
  auto kern = sde_small_kernel<D, PREF, PREG, SCHEME>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid_for(p.B, kFixThreads), kFixThreads, smem, s>>>(p);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  return XDE_OK;
}
template <int D, int PREF, int PREG>
static int sde_scheme(const FixParams &p, int scheme, cudaStream_t s) {
  return scheme == XDE_SDE_EM ? launch_sde<D, PREF, PREG, XDE_SDE_EM>(p, s)
                              : launch_sde<D, PREF, PREG, XDE_SDE_MILSTEIN>(p, s);
}
template <int D, int PREF>
static int sde_preg(const FixParams &p, int scheme, cudaStream_t s) {
  switch (p.g.pre) {
    case XDE_PRE_ID: return sde_scheme<D, PREF, XDE_PRE_ID>(p, scheme, s);
    case XDE_PRE_SQUARE: return sde_scheme<D, PREF, XDE_PRE_SQUARE>(p, scheme, s);
    case XDE_PRE_CUBE: return sde_scheme<D, PREF, XDE_PRE_CUBE>(p, scheme, s);
  }
  set_last_error("unknown pre-activation %d", p.g.pre);
  return XDE_E_BAD_ARG;
}
template <int D>
static int sde_pref(const FixParams &p, int scheme, cudaStream_t s) {
  switch (p.f.pre) {
    case XDE_PRE_ID: return sde_preg<D, XDE_PRE_ID>(p, scheme, s);
    case XDE_PRE_SQUARE: return sde_preg<D, XDE_PRE_SQUARE>(p, scheme, s);
    case XDE_PRE_CUBE: return sde_preg<D, XDE_PRE_CUBE>(p, scheme, s);
  }
  set_last_error("unknown pre-activation %d", p.f.pre);
  return XDE_E_BAD_ARG;
}

int rk_fixed_small(int method, const xde_mlp_field_t *f, const float *y0, long long B, const float *t_span,
                   int T, int stride, float *out, cudaStream_t s) {
  FixParams p{};
  p.f = *f;
  p.y0 = y0;
  p.t_span = t_span;
  p.out = out;
  p.B = B;
  p.T = T;
  p.stride = stride;
  p.n_out = (T - 1 + stride - 1) / stride + 1;
  p.method = method;
  switch (f->d) {
    case 1: return fixed_pre<1>(p, s);
    case 2: return fixed_pre<2>(p, s);
    case 3: return fixed_pre<3>(p, s);
    case 4: return fixed_pre<4>(p, s);
    case 5: return fixed_pre<5>(p, s);
    case 6: return fixed_pre<6>(p, s);
    case 7: return fixed_pre<7>(p, s);
    case 8: return fixed_pre<8>(p, s);
  }
  set_last_error("fixed solver: state dim D=%d has no fused kernel", f->d);
  return XDE_E_UNSUPPORTED_FIELD;
}

int sde_small(int scheme, const xde_mlp_field_t *f, const xde_mlp_field_t *g, const float *y0, long long B,
              const float *t_span, int T, const BmSource &bm, int stride, float *out, cudaStream_t s) {
  FixParams p{};
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
  switch (f->d) {
    case 1: return sde_pref<1>(p, scheme, s);
    case 2: return sde_pref<2>(p, scheme, s);
    case 3: return sde_pref<3>(p, scheme, s);
    case 4: return sde_pref<4>(p, scheme, s);
    case 5: return sde_pref<5>(p, scheme, s);
    case 6: return sde_pref<6>(p, scheme, s);
    case 7: return sde_pref<7>(p, scheme, s);
    case 8: return sde_pref<8>(p, scheme, s);
  }
  set_last_error("sde: state dim D=%d has no fused kernel", f->d);
  return XDE_E_UNSUPPORTED_FIELD;
}

}  // namespace xde
