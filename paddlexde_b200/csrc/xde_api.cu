// xde_api.cu -- C-ABI entry points that dispatch between the small-state and the tiled kernels.
#include "xde_common.cuh"

namespace xde {
int rk_fixed_small(int method, const xde_mlp_field_t *f, const float *y0, long long B, const float *t_span,
                   int T, int stride, float *out, cudaStream_t s);
int sde_small(int scheme, const xde_mlp_field_t *f, const xde_mlp_field_t *g, const float *y0, long long B,
              const float *t_span, int T, const BmSource &bm, int stride, float *out, cudaStream_t s);
int rk_fixed_tile(int method, const xde_mlp_field_t *f, const float *y0, long long B, const float *t_span,
                  int T, int stride, float *out, cudaStream_t s);
int sde_tile(int scheme, const xde_mlp_field_t *f, const xde_mlp_field_t *g, const float *y0, long long B,
             const float *t_span, int T, const BmSource &bm, int stride, float *out, cudaStream_t s);
bool tile_covers(int D);
int rk_fixed_tc(int method, const xde_mlp_field_t *f, const float *y0, long long B, const float *t_span, int T,
                int stride, float *out, int *status, cudaStream_t s);
int sde_tc(int scheme, const xde_mlp_field_t *f, const xde_mlp_field_t *g, const float *y0, long long B,
           const float *t_span, int T, const BmSource &bm, int stride, float *out, int *status, cudaStream_t s);
// linear_interp (interpolation/functional/interp_fn.py:4-10) of the solver's grid solution at the caller's output
// times: out[b, i] = y[b, i-1] + ((t_out[i] - grid[i-1]) / (grid[i] - grid[i-1])) * (y[b, i] - y[b, i-1]), with the
// two equality shortcuts of the reference; row 0 is copied.  One thread per value.
__global__ void __launch_bounds__(256) fixed_interp_linear_kernel(const float *__restrict__ y, const float *__restrict__ grid,
                                                                  const float *__restrict__ t_out, long long B, int T, int D,
                                                                  float *__restrict__ out) {
  const long long n = B * (long long)T * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int it = (int)((i / D) % T);
    if (it == 0) {
      out[i] = y[i];
      continue;
    }
    const float t0 = grid[it - 1], t1 = grid[it], t = t_out[it];
    const float y0 = y[i - D], y1 = y[i];
    float v;
    if (t == t0) v = y0;
    else if (t == t1) v = y1;
    else v = y0 + __fdiv_rn(t - t0, t1 - t0) * (y1 - y0);
    out[i] = v;
  }
}
}  // namespace xde

extern "C" XDE_EXPORT int xde_fixed_interp_linear_f32(const float *y_grid, const float *grid, const float *t_out, int64_t B,
                                                      int32_t T, int32_t D, float *out, void *stream) {
  using namespace xde;
  XDE_REQUIRE(y_grid && grid && t_out && out, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 1 && D >= 1, XDE_E_BAD_ARG, "need B>=1, T>=1, D>=1");
  const long long n = B * (long long)T * D;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  fixed_interp_linear_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y_grid, grid, t_out, B, T, D, out);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  return XDE_OK;
}

extern "C" XDE_EXPORT int xde_rk_fixed_mlp_f32(int32_t method, const xde_mlp_field_t *field, const float *y0,
                                               int64_t B, const float *t_span, int32_t T, int32_t out_stride_t,
                                               float *out, void *stream) {
  using namespace xde;
  XDE_REQUIRE(field && y0 && t_span && out, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 1 && out_stride_t >= 1, XDE_E_BAD_ARG, "need B>=1, T>=1, out_stride_t>=1");
  XDE_REQUIRE(method == XDE_FIXED_EULER || method == XDE_FIXED_RK4_38 || method == XDE_FIXED_MIDPOINT, XDE_E_BAD_ARG,
              "unknown method %d", method);
  if (tile_covers(field->d)) return rk_fixed_tile(method, field, y0, B, t_span, T, out_stride_t, out, (cudaStream_t)stream);
  return rk_fixed_small(method, field, y0, B, t_span, T, out_stride_t, out, (cudaStream_t)stream);
}

extern "C" XDE_EXPORT int xde_sde_mlp_f32(int32_t scheme, const xde_mlp_field_t *drift,
                                          const xde_mlp_field_t *diffusion, const float *y0, int64_t B,
                                          const float *t_span, int32_t T, const float *dW, int32_t out_stride_t,
                                          float *out, void *stream) {
  using namespace xde;
  XDE_REQUIRE(drift && diffusion && y0 && t_span && dW && out, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 1 && out_stride_t >= 1, XDE_E_BAD_ARG, "need B>=1, T>=1, out_stride_t>=1");
  XDE_REQUIRE(drift->d == diffusion->d, XDE_E_BAD_ARG, "drift and diffusion state dims differ");
  XDE_REQUIRE(scheme == XDE_SDE_EM || scheme == XDE_SDE_MILSTEIN, XDE_E_BAD_ARG, "unknown scheme %d", scheme);
  const BmSource bm{dW, 0ull, 0ll};
  if (tile_covers(drift->d))
    return sde_tile(scheme, drift, diffusion, y0, B, t_span, T, bm, out_stride_t, out, (cudaStream_t)stream);
  return sde_small(scheme, drift, diffusion, y0, B, t_span, T, bm, out_stride_t, out, (cudaStream_t)stream);
}

extern "C" XDE_EXPORT int xde_rk_fixed_mlp_tc_f32(int32_t method, const xde_mlp_field_t *field, const float *y0,
                                                  int64_t B, const float *t_span, int32_t T, int32_t out_stride_t,
                                                  float *out, int32_t *status, void *stream) {
  using namespace xde;
  XDE_REQUIRE(field && y0 && t_span && out, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 1 && out_stride_t >= 1, XDE_E_BAD_ARG, "need B>=1, T>=1, out_stride_t>=1");
  XDE_REQUIRE(method == XDE_FIXED_EULER || method == XDE_FIXED_RK4_38 || method == XDE_FIXED_MIDPOINT, XDE_E_BAD_ARG,
              "unknown method %d", method);
  return rk_fixed_tc(method, field, y0, B, t_span, T, out_stride_t, out, status, (cudaStream_t)stream);
}

extern "C" XDE_EXPORT int xde_sde_mlp_tc_f32(int32_t scheme, const xde_mlp_field_t *drift,
                                             const xde_mlp_field_t *diffusion, const float *y0, int64_t B,
                                             const float *t_span, int32_t T, const float *dW, int32_t out_stride_t,
                                             float *out, int32_t *status, void *stream) {
  using namespace xde;
  XDE_REQUIRE(drift && diffusion && y0 && t_span && dW && out, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 1 && out_stride_t >= 1, XDE_E_BAD_ARG, "need B>=1, T>=1, out_stride_t>=1");
  XDE_REQUIRE(drift->d == diffusion->d, XDE_E_BAD_ARG, "drift and diffusion state dims differ");
  XDE_REQUIRE(scheme == XDE_SDE_EM || scheme == XDE_SDE_MILSTEIN, XDE_E_BAD_ARG, "unknown scheme %d", scheme);
  return sde_tc(scheme, drift, diffusion, y0, B, t_span, T, BmSource{dW, 0ull, 0ll}, out_stride_t, out, status,
                (cudaStream_t)stream);
}

// ---- Brownian increments from the counter-based generator (xde_common.cuh: BmSource) ----------------------
namespace xde {
__global__ void __launch_bounds__(256) brownian_table_kernel(BmSource bm, const float *__restrict__ t_span, int T,
                                                             long long B, int D, float *__restrict__ dW) {
  const int G4 = (D + 3) / 4;
  const long long total = (long long)(T - 1) * B * G4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d4 = (int)(i % G4);
    const long long nb = i / G4;
    const long long b = nb % B;
    const int n = (int)(nb / B);
    const float sq = __fsqrt_rn(fabsf(t_span[n + 1] - t_span[n]));
    const float4 z = bm_normal4(bm.seed, n, b + bm.traj_offset, d4);
    const float v[4] = {z.x * sq, z.y * sq, z.z * sq, z.w * sq};
    float *o = dW + ((long long)n * B + b) * D + 4 * d4;
    for (int e = 0; e < 4 && 4 * d4 + e < D; ++e) o[e] = v[e];
  }
}
}  // namespace xde

extern "C" XDE_EXPORT int xde_brownian_increments_f32(uint64_t seed, int64_t traj_offset, const float *t_span, int32_t T,
                                                      int64_t B, int32_t D, float *dW, void *stream) {
  using namespace xde;
  XDE_REQUIRE(t_span && dW, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(T >= 2 && B >= 1 && D >= 1, XDE_E_BAD_ARG, "need T>=2, B>=1, D>=1");
  const long long total = (long long)(T - 1) * B * ((D + 3) / 4);
  long long grid = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (grid > cap) grid = cap;
  brownian_table_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(BmSource{nullptr, seed, traj_offset}, t_span, T,
                                                                          B, D, dW);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  return XDE_OK;
}

extern "C" XDE_EXPORT int xde_sde_mlp_philox_f32(int32_t scheme, int32_t math, const xde_mlp_field_t *drift,
                                                 const xde_mlp_field_t *diffusion, const float *y0, int64_t B,
                                                 const float *t_span, int32_t T, uint64_t seed, int64_t traj_offset,
                                                 int32_t out_stride_t, float *out, int32_t *status, void *stream) {
  using namespace xde;
  XDE_REQUIRE(drift && diffusion && y0 && t_span && out, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 1 && out_stride_t >= 1, XDE_E_BAD_ARG, "need B>=1, T>=1, out_stride_t>=1");
  XDE_REQUIRE(drift->d == diffusion->d, XDE_E_BAD_ARG, "drift and diffusion state dims differ");
  XDE_REQUIRE(scheme == XDE_SDE_EM || scheme == XDE_SDE_MILSTEIN, XDE_E_BAD_ARG, "unknown scheme %d", scheme);
  XDE_REQUIRE(math == XDE_MATH_FP32 || math == XDE_MATH_TENSOR, XDE_E_BAD_ARG, "unknown math mode %d", math);
  const BmSource bm{nullptr, seed, traj_offset};
  cudaStream_t s = (cudaStream_t)stream;
  if (math == XDE_MATH_TENSOR) return sde_tc(scheme, drift, diffusion, y0, B, t_span, T, bm, out_stride_t, out, status, s);
  if (status) XDE_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int), s));  // the FP32 kernels have no range limit
  if (tile_covers(drift->d)) return sde_tile(scheme, drift, diffusion, y0, B, t_span, T, bm, out_stride_t, out, s);
  return sde_small(scheme, drift, diffusion, y0, B, t_span, T, bm, out_stride_t, out, s);
}
