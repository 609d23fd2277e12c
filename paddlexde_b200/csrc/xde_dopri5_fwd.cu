// xde_dopri5_fwd.cu -- odeint(func=MLP, solver=Dopri5) forward, per-trajectory controller.
//
// Replaces the Python hot loop of solver/base_adaptive_solver.py:24-31 +
// solver/base_adaptive_solver_rk.py:116-292 (+ utils/ode_utils.py:28-97) for the fused MLP field.
//
// Design (small state, D <= 8): ONE launch integrates every trajectory over the whole t_span.
//   * one thread owns one trajectory at a time: state, the 7 Dormand-Prince stages, the error
//     controller (t, dt, accept/reject) all live in registers -- no host sync, no global k buffer;
//   * the MLP weights sit in shared memory as one 16-byte-aligned record per hidden unit
//     (broadcast LDS.128), the field is evaluated with explicit fmaf chains + a rational tanh;
//   * the warp advances in lock-step "blocks" of 6 field evaluations (one attempt); a lane that
//     finished its trajectory fetches the next one from a grid-wide queue and spends its block on
//     select_initial_step (2 evaluations) -- when every lane of the warp is initialising the block
//     is cut to 2 evaluations, so uniform workloads pay nothing for the refill;
//   * per-trajectory HBM traffic: y0 in, T output rows out (+ optional attempt log).
#include "xde_common.cuh"

namespace xde {

constexpr int kFwdThreads = 128;

struct FwdParams {
  xde_mlp_field_t field;
  const float *y0;
  const float *t_span;
  float *out;
  long long B;
  int T;
  xde_ctrl_opts_t o;
  xde_stats_t *stats;
  xde_attempt_t *log_records;
  int *log_counts;
  int log_cap;
  unsigned long long *queue;  // next unclaimed trajectory, shared by the whole grid (zeroed)
};

enum { MODE_IDLE = 0, MODE_INIT = 1, MODE_ATTEMPT = 2, MODE_DONE = 3 };

template <int D>
__device__ __forceinline__ float rms_small(const float (&v)[D]) {
  double acc = 0.0;
#pragma unroll
  for (int e = 0; e < D; ++e) acc += (double)(v[e] * v[e]);
  return rms_from_sumsq(acc, (double)D);
}

template <int D, int PRE>
__global__ void __launch_bounds__(kFwdThreads) dopri5_fwd_small_kernel(const FwdParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned long long s_cnt[3];
  __shared__ int s_status;

  const int H = p.field.h;
  float *sw = smem;
  float *st = smem + SmallRec<D>::floats(H);  // t_span copy (solver time: negated when rev)
  load_small_field<D>(sw, p.field);
  // a decreasing t_span is integrated as s = -t with f~(s, y) = -f(-s, y) (repair R5)
  const bool rev = p.t_span[1] < p.t_span[0];
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) st[i] = rev ? -p.t_span[i] : p.t_span[i];

  if (threadIdx.x == 0) {
    s_cnt[0] = s_cnt[1] = s_cnt[2] = 0ull;
    s_status = 0;
  }
  __syncthreads();

  const xde_ctrl_opts_t o = p.o;
  const int lane = threadIdx.x & 31;
  const float fsign = rev ? -1.0f : 1.0f;

  int mode = MODE_IDLE;
  long long traj = -1;
  int i_out = 1, n_steps = 0, n_logged = 0;
  float y0[D], k[7][D];
#pragma unroll
  for (int j = 0; j < 7; ++j)
#pragma unroll
    for (int e = 0; e < D; ++e) k[j][e] = 0.0f;
#pragma unroll
  for (int e = 0; e < D; ++e) y0[e] = 0.0f;
  float t0 = 0.f, dt = 0.f;
  unsigned n_att = 0, n_acc = 0, n_fe = 0;
  int status = 0;

  while (true) {
    // ---- refill idle lanes from the grid-wide queue (warp-aggregated atomic; one queue for the whole grid:
    // with a contiguous chunk per CTA the CTAs of the adjoint finished up to 8 % apart) ----
    {
      const bool need = (mode == MODE_IDLE);
      const unsigned m = __ballot_sync(XDE_FULL_MASK, need);
      if (m) {
        unsigned long long base = 0;
        const int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(p.queue, (unsigned long long)__popc(m));
        base = __shfl_sync(XDE_FULL_MASK, base, leader);
        if (need) {
          const long long cand = (long long)base + __popc(m & ((1u << lane) - 1u));
          if (cand < p.B) {
            traj = cand;
#pragma unroll
            for (int e = 0; e < D; ++e) {
              y0[e] = p.y0[traj * D + e];
              p.out[traj * D + e] = y0[e];  // solution[0] = y0
            }
            t0 = st[0];
            i_out = 1;
            n_steps = 0;
            n_logged = 0;
            mode = MODE_INIT;
          } else {
            mode = MODE_DONE;
          }
        }
      }
    }
    if (__all_sync(XDE_FULL_MASK, mode == MODE_DONE)) break;
    const bool any_attempt = __any_sync(XDE_FULL_MASK, mode == MODE_ATTEMPT);
    const bool att = (mode == MODE_ATTEMPT);
    const bool ini = (mode == MODE_INIT);

    // ---- attempt prologue: assertions of _adaptive_step (base_adaptive_solver_rk.py:200-203) ----
    float t1 = t0 + dt;
    bool live = att;
    if (att) {
      if (!(n_steps < o.max_num_steps)) {
        status = max(status, XDE_ST_MAX_STEPS);
        live = false;
      } else if (!(t0 + dt > t0)) {
        status = max(status, XDE_ST_DT_UNDERFLOW);
        live = false;
      } else {
        bool fin = true;
#pragma unroll
        for (int e = 0; e < D; ++e) fin = fin && (fabsf(y0[e]) < INFINITY);
        if (!fin) {
          status = max(status, XDE_ST_NONFINITE_STATE);
          live = false;
        }
      }
      if (!live) {  // abort this trajectory: remaining outputs are NaN
        for (int i = i_out; i < p.T; ++i)
#pragma unroll
          for (int e = 0; e < D; ++e) p.out[((long long)i * p.B + traj) * D + e] = NAN;
        if (p.log_counts) p.log_counts[traj] = n_logged;
        mode = MODE_IDLE;
      }
    }

    // INIT scratch
    float scale[D], h0 = 0.f, d1 = 0.f;

    // ---- the 6 (or 2) field evaluations of this block ----
    float yin[D], fo[D];
    // evaluation 1: ATTEMPT stage 1 | INIT f0 = f(t0, y0)
#pragma unroll
    for (int e = 0; e < D; ++e) {
      float s = k[0][e] * (DP::beta(0, 0) * dt);
      yin[e] = att ? (y0[e] + s) : y0[e];
    }
    mlp_eval_small<D, PRE>(sw, H, yin, fo);
#pragma unroll
    for (int e = 0; e < D; ++e) fo[e] *= fsign;
    if (ini) {
      // _before_integrate: f0 (base_adaptive_solver_rk.py:83); select_initial_step recomputes the
      // same value (base_adaptive_solver.py:47-48) -- evaluated once here, counted twice in nfe.
      float v[D];
#pragma unroll
      for (int e = 0; e < D; ++e) {
        k[0][e] = fo[e];
        scale[e] = o.atol + fabsf(y0[e]) * o.rtol;
        v[e] = __fdiv_rn(y0[e], scale[e]);
      }
      const float d0 = fabsf(rms_small<D>(v));
#pragma unroll
      for (int e = 0; e < D; ++e) v[e] = __fdiv_rn(fo[e], scale[e]);
      d1 = fabsf(rms_small<D>(v));
      if (d0 < 1e-5f || d1 < 1e-5f)
        h0 = 1e-6f;
      else
        h0 = __fdiv_rn(0.01f * d0, d1);
      h0 = fabsf(h0);
    } else {
#pragma unroll
      for (int e = 0; e < D; ++e) k[1][e] = fo[e];
    }
    // evaluation 2: ATTEMPT stage 2 | INIT probe f(t0 + h0, y0 + f0*h0)
#pragma unroll
    for (int e = 0; e < D; ++e) {
      float s = k[0][e] * (DP::beta(1, 0) * dt);
      s = s + k[1][e] * (DP::beta(1, 1) * dt);
      yin[e] = att ? (y0[e] + s) : (k[0][e] * h0 + y0[e]);
    }
    mlp_eval_small<D, PRE>(sw, H, yin, fo);
#pragma unroll
    for (int e = 0; e < D; ++e) fo[e] *= fsign;
    if (ini) {
      float v[D];
#pragma unroll
      for (int e = 0; e < D; ++e) v[e] = __fdiv_rn(fo[e] - k[0][e], scale[e]);
      const float d2 = fabsf(__fdiv_rn(rms_small<D>(v), h0));
      float h1;
      if (d1 <= 1e-15f && d2 <= 1e-15f) {
        h1 = fmaxf(1e-6f, h0 * 1e-3f);
      } else {
        const float mx = (d2 > d1) ? d2 : d1;
        const float arg = __fdiv_rn(0.01f, mx);
        h1 = (arg > 0.0f && arg < INFINITY) ? root5(arg) : arg;
      }
      h1 = fabsf(h1);
      const float sel = fminf(100.0f * h0, h1);
      dt = (o.first_step == o.first_step) ? o.first_step : sel;
      n_fe += (o.first_step == o.first_step) ? 1u : 3u;
      mode = MODE_ATTEMPT;  // first attempt starts with the next block
    } else {
#pragma unroll
      for (int e = 0; e < D; ++e) k[2][e] = fo[e];
    }

    if (any_attempt) {
      // evaluations 3..6: stages 3..6 (INIT lanes idle along)
#pragma unroll
      for (int i = 2; i < 6; ++i) {
#pragma unroll
        for (int e = 0; e < D; ++e) {
          float s = k[0][e] * (DP::beta(i, 0) * dt);
#pragma unroll
          for (int j = 1; j <= i; ++j) s = s + k[j][e] * (DP::beta(i, j) * dt);
          yin[e] = y0[e] + s;
        }
        mlp_eval_small<D, PRE>(sw, H, yin, fo);
        if (live) {
#pragma unroll
          for (int e = 0; e < D; ++e) k[i + 1][e] = fo[e] * fsign;
        }
      }

      if (live) {
        // y1 = yi of the last stage (FSAL shortcut), f1 = k[6]; error estimate; error ratio
        float v[D];
#pragma unroll
        for (int e = 0; e < D; ++e) {
          float s = k[0][e] * (dt * DP::cerr(0));
#pragma unroll
          for (int j = 1; j < 7; ++j) s = s + k[j][e] * (dt * DP::cerr(j));
          const float tol = o.atol + o.rtol * fmaxf(fabsf(y0[e]), fabsf(yin[e]));
          v[e] = __fdiv_rn(s, tol);
        }
        const float ratio = fabsf(rms_small<D>(v));
        bool accept = (ratio <= 1.0f);
        if (dt > o.max_step) accept = false;
        if (dt <= o.min_step) accept = true;
        n_att++;
        n_fe += 6;
        n_steps++;
        if (p.log_records) {
          if (n_logged < p.log_cap) {
            xde_attempt_t r;
            r.t0 = rev ? -t0 : t0;
            r.dt = rev ? -dt : dt;
            r.ratio = ratio;
            r.accepted = accept ? 1 : 0;
            p.log_records[traj * p.log_cap + n_logged] = r;
          }
          n_logged++;
        }
        const float dt_next = next_step_size(dt, ratio, o);
        if (accept) {
          n_acc++;
          // _interp_fit + interp_fit (base_adaptive_solver_rk.py:286-292, utils/ode_utils.py:28-49);
          // outputs falling inside [t0, t1] are evaluated right away (interp_evaluate :52-77)
          float ce[D], cd[D], cc[D], cb[D], ca[D];
          const float two_dt = 2.0f * dt;
#pragma unroll
          for (int e = 0; e < D; ++e) {
            float s = k[0][e] * (dt * DP::cmid(0));
#pragma unroll
            for (int j = 1; j < 7; ++j) s = s + k[j][e] * (dt * DP::cmid(j));
            const float ym = y0[e] + s;
            const float F0 = k[0][e], F1 = k[6][e], Y0 = y0[e], Y1 = yin[e];
            ca[e] = (two_dt * (F1 - F0) - 8.0f * (Y1 + Y0)) + 16.0f * ym;
            cb[e] = ((dt * (5.0f * F0 - 3.0f * F1) + 18.0f * Y0) + 14.0f * Y1) - 32.0f * ym;
            cc[e] = ((dt * (F1 - 4.0f * F0) - 11.0f * Y0) - 5.0f * Y1) + 16.0f * ym;
            cd[e] = dt * F0;
            ce[e] = Y0;
          }
          while (i_out < p.T && !(st[i_out] > t1)) {
            const float x = __fdiv_rn(st[i_out] - t0, t1 - t0);
#pragma unroll
            for (int e = 0; e < D; ++e) {
              float total = ce[e] + x * cd[e];
              float xp = x * x;
              total = total + xp * cc[e];
              xp = xp * x;
              total = total + xp * cb[e];
              xp = xp * x;
              total = total + xp * ca[e];
              p.out[((long long)i_out * p.B + traj) * D + e] = total;
            }
            i_out++;
            n_steps = 0;
          }
#pragma unroll
          for (int e = 0; e < D; ++e) {
            y0[e] = yin[e];
            k[0][e] = k[6][e];
          }
          t0 = t1;
          if (i_out >= p.T) {
            if (p.log_counts) p.log_counts[traj] = n_logged;
            mode = MODE_IDLE;
          }
        }
        dt = dt_next;
      }
    }
  }

  // ---- stats: warp -> CTA -> global ----
  {
    unsigned a = n_att, b = n_acc, c = n_fe;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(XDE_FULL_MASK, a, off);
      b += __shfl_xor_sync(XDE_FULL_MASK, b, off);
      c += __shfl_xor_sync(XDE_FULL_MASK, c, off);
      status = max(status, __shfl_xor_sync(XDE_FULL_MASK, status, off));
    }
    if (lane == 0) {
      atomicAdd(&s_cnt[0], (unsigned long long)a);
      atomicAdd(&s_cnt[1], (unsigned long long)b);
      atomicAdd(&s_cnt[2], (unsigned long long)c);
      atomicMax(&s_status, status);
    }
    __syncthreads();
    if (threadIdx.x == 0 && p.stats) {
      atomicAdd(&p.stats->n_attempts, s_cnt[0]);
      atomicAdd(&p.stats->n_accepted, s_cnt[1]);
      atomicAdd(&p.stats->nfe, s_cnt[2]);
      atomicMax(&p.stats->status, s_status);
    }
  }
}

template <int D, int PRE>
static int launch_fwd_small(const FwdParams &p, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (SmallRec<D>::floats(p.field.h) + p.T);
  XDE_REQUIRE(smem <= 200 * 1024, XDE_E_UNSUPPORTED_FIELD,
              "dopri5 forward: field (H=%d) + t_span (T=%d) exceed shared memory", p.field.h, p.T);
  auto kern = dopri5_fwd_small_kernel<D, PRE>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  XDE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kFwdThreads, smem));
  if (per_sm < 1) per_sm = 1;
  // persistent grid: a whole number of CTAs per SM, all fed from one trajectory queue
  long long want = (p.B + kFwdThreads - 1) / kFwdThreads;
  long long grid = (long long)sm_count() * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  FwdParams q = p;
  XDE_CUDA_CHECK(scratch_alloc((void **)&q.queue, sizeof(unsigned long long), stream));
  XDE_CUDA_CHECK(cudaMemsetAsync(q.queue, 0, sizeof(unsigned long long), stream));
  kern<<<(unsigned)grid, kFwdThreads, smem, stream>>>(q);
  count_launch();
  cudaError_t le = cudaGetLastError();
  cudaFreeAsync(q.queue, stream);
  XDE_CUDA_CHECK(le);
  return XDE_OK;
}

template <int D>
static int dispatch_pre(const FwdParams &p, cudaStream_t s) {
  switch (p.field.pre) {
    case XDE_PRE_ID: return launch_fwd_small<D, XDE_PRE_ID>(p, s);
    case XDE_PRE_SQUARE: return launch_fwd_small<D, XDE_PRE_SQUARE>(p, s);
    case XDE_PRE_CUBE: return launch_fwd_small<D, XDE_PRE_CUBE>(p, s);
  }
  set_last_error("unknown pre-activation %d", p.field.pre);
  return XDE_E_BAD_ARG;
}

int dopri5_fwd_batch(const xde_mlp_field_t *field, const float *y0, long long B, const float *t_span, int T,
                     const xde_ctrl_opts_t *opts, float *out, xde_stats_t *stats, const xde_attempt_log_t *log,
                     cudaStream_t s);  // xde_dopri5_batch.cu
int dopri5_fwd_tile(const xde_mlp_field_t *field, const float *y0, long long B, const float *t_span, int T,
                    const xde_ctrl_opts_t *opts, float *out, xde_stats_t *stats, const xde_attempt_log_t *log,
                    cudaStream_t s);  // xde_tile_adaptive.cu (D >= 16)

}  // namespace xde

extern "C" XDE_EXPORT int xde_dopri5_mlp_f32(const xde_mlp_field_t *field, const float *y0, int64_t B,
                                  const float *t_span, int32_t T, const xde_ctrl_opts_t *opts,
                                  int32_t controller, float *out, xde_stats_t *stats,
                                  const xde_attempt_log_t *log, void *stream) {
  using namespace xde;
  XDE_REQUIRE(field && y0 && t_span && opts && out, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 2, XDE_E_BAD_ARG, "need B >= 1 and T >= 2 (B=%lld T=%d)", (long long)B, T);
  XDE_REQUIRE(controller == XDE_CTRL_TRAJECTORY || controller == XDE_CTRL_BATCH, XDE_E_BAD_ARG,
              "unknown controller %d", controller);
  cudaStream_t s = (cudaStream_t)stream;
  if (controller == XDE_CTRL_BATCH) {
    if (stats) XDE_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(xde_stats_t), s));
    return dopri5_fwd_batch(field, y0, B, t_span, T, opts, out, stats, log, s);
  }
  // strict monotonicity of t_span is the caller's precondition (the shim checks its host copy);
  // the direction is read on the device so that this call never synchronises.
  FwdParams p{};
  p.field = *field;
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
  if (stats) XDE_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(xde_stats_t), s));
  switch (field->d) {
    case 1: return dispatch_pre<1>(p, s);
    case 2: return dispatch_pre<2>(p, s);
    case 3: return dispatch_pre<3>(p, s);
    case 4: return dispatch_pre<4>(p, s);
    case 5: return dispatch_pre<5>(p, s);
    case 6: return dispatch_pre<6>(p, s);
    case 7: return dispatch_pre<7>(p, s);
    case 8: return dispatch_pre<8>(p, s);
    default:
      if (field->d >= 16) return dopri5_fwd_tile(field, y0, B, t_span, T, opts, out, stats, log, s);
      set_last_error("dopri5 forward: state dim D=%d has no fused kernel (supported: 1..8 and 16,32,64)", field->d);
      return XDE_E_UNSUPPORTED_FIELD;
  }
}
