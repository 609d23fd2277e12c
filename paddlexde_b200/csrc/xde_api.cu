// xde_api.cu -- C-ABI entry points that dispatch between the small-state and the tiled kernels.
#include "xde_common.cuh"

namespace xde {
int rk_fixed_small(int method, const xde_mlp_field_t *f, const float *y0, long long B, const float *t_span,
                   int T, int stride, float *out, cudaStream_t s);
int sde_small(int scheme, const xde_mlp_field_t *f, const xde_mlp_field_t *g, const float *y0, long long B,
              const float *t_span, int T, const float *dW, int stride, float *out, cudaStream_t s);
int rk_fixed_tile(int method, const xde_mlp_field_t *f, const float *y0, long long B, const float *t_span,
                  int T, int stride, float *out, cudaStream_t s);
int sde_tile(int scheme, const xde_mlp_field_t *f, const xde_mlp_field_t *g, const float *y0, long long B,
             const float *t_span, int T, const float *dW, int stride, float *out, cudaStream_t s);
bool tile_covers(int D);
int rk_fixed_tc(int method, const xde_mlp_field_t *f, const float *y0, long long B, const float *t_span, int T,
                int stride, float *out, cudaStream_t s);
int sde_tc(int scheme, const xde_mlp_field_t *f, const xde_mlp_field_t *g, const float *y0, long long B,
           const float *t_span, int T, const float *dW, int stride, float *out, cudaStream_t s);
}  // namespace xde

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
  if (tile_covers(drift->d))
    return sde_tile(scheme, drift, diffusion, y0, B, t_span, T, dW, out_stride_t, out, (cudaStream_t)stream);
  return sde_small(scheme, drift, diffusion, y0, B, t_span, T, dW, out_stride_t, out, (cudaStream_t)stream);
}

extern "C" XDE_EXPORT int xde_rk_fixed_mlp_tc_f32(int32_t method, const xde_mlp_field_t *field, const float *y0,
                                                  int64_t B, const float *t_span, int32_t T, int32_t out_stride_t,
                                                  float *out, void *stream) {
  using namespace xde;
  XDE_REQUIRE(field && y0 && t_span && out, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 1 && out_stride_t >= 1, XDE_E_BAD_ARG, "need B>=1, T>=1, out_stride_t>=1");
  XDE_REQUIRE(method == XDE_FIXED_EULER || method == XDE_FIXED_RK4_38 || method == XDE_FIXED_MIDPOINT, XDE_E_BAD_ARG,
              "unknown method %d", method);
  return rk_fixed_tc(method, field, y0, B, t_span, T, out_stride_t, out, (cudaStream_t)stream);
}

extern "C" XDE_EXPORT int xde_sde_mlp_tc_f32(int32_t scheme, const xde_mlp_field_t *drift,
                                             const xde_mlp_field_t *diffusion, const float *y0, int64_t B,
                                             const float *t_span, int32_t T, const float *dW, int32_t out_stride_t,
                                             float *out, void *stream) {
  using namespace xde;
  XDE_REQUIRE(drift && diffusion && y0 && t_span && dW && out, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 1 && out_stride_t >= 1, XDE_E_BAD_ARG, "need B>=1, T>=1, out_stride_t>=1");
  XDE_REQUIRE(drift->d == diffusion->d, XDE_E_BAD_ARG, "drift and diffusion state dims differ");
  XDE_REQUIRE(scheme == XDE_SDE_EM || scheme == XDE_SDE_MILSTEIN, XDE_E_BAD_ARG, "unknown scheme %d", scheme);
  return sde_tc(scheme, drift, diffusion, y0, B, t_span, T, dW, out_stride_t, out, (cudaStream_t)stream);
}
