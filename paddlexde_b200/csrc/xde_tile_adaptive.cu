// xde_tile_adaptive.cu -- entry of the tiled adaptive solvers (large states) and the Dormand-Prince instantiation;
// the kernel is xde_tile_adaptive.cuh, the other stage counts live in xde_tile_adaptive_s{1,2,3}.cu so that the
// instantiations compile in parallel.
#include "xde_tile_adaptive.cuh"

namespace xde {

int ad_tile_s1(const AdTileParams &p, cudaStream_t s);  // AdaptiveHeun
int ad_tile_s2(const AdTileParams &p, cudaStream_t s);  // Fehlberg2
int ad_tile_s3(const AdTileParams &p, cudaStream_t s);  // Bosh3
static int ad_tile_s6(const AdTileParams &p, cudaStream_t s) { return ad_tile_dispatch<6>(p, s); }  // Dopri5

// forward solve with one controller per trajectory for D >= 16; method = XDE_RK_* (Dopri8's 13 stages do not fit)
int adaptive_rk_tile(int method, const xde_mlp_field_t *field, const float *y0, long long B, const float *t_span, int T,
                     const xde_ctrl_opts_t *opts, const float *step_t, int n_step, const float *jump_t, int n_jump,
                     float *out, xde_stats_t *stats, const xde_attempt_log_t *log, cudaStream_t s) {
  AdTileParams p{};
  XDE_REQUIRE(make_tab(method == XDE_RK_DOPRI5 ? XDE_RK_DOPRI5_TABLE : method, p.tab), XDE_E_BAD_ARG,
              "unknown Runge-Kutta method %d", method);
  p.f = *field;
  p.y0 = y0;
  p.t_span = t_span;
  p.out = out;
  p.B = B;
  p.T = T;
  p.o = *opts;
  p.stats = stats;
  p.log_records = log ? log->records : nullptr;
  p.log_counts = log ? log->counts : nullptr;
  p.log_cap = log ? log->cap : 0;
  p.step_t = step_t;
  p.n_step = n_step;
  p.jump_t = jump_t;
  p.n_jump = n_jump;
  switch (p.tab.S) {
    case 1: return ad_tile_s1(p, s);
    case 2: return ad_tile_s2(p, s);
    case 3: return ad_tile_s3(p, s);
    case 6: return ad_tile_s6(p, s);
  }
  set_last_error("tiled adaptive solver: the %d stages of this tableau do not fit the registers next to the weights "
                 "(large states: Dopri5, Bosh3, Fehlberg2, AdaptiveHeun)", p.tab.S);
  return XDE_E_UNSUPPORTED_FIELD;
}

int dopri5_fwd_tile(const xde_mlp_field_t *field, const float *y0, long long B, const float *t_span, int T,
                    const xde_ctrl_opts_t *opts, float *out, xde_stats_t *stats, const xde_attempt_log_t *log,
                    cudaStream_t s) {
  return adaptive_rk_tile(XDE_RK_DOPRI5, field, y0, B, t_span, T, opts, nullptr, 0, nullptr, 0, out, stats, log, s);
}

}  // namespace xde
