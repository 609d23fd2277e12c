// xde_dopri5_batch.cu -- odeint(func=MLP, solver=Dopri5) forward with the REFERENCE-FAITHFUL controller:
// one global RMS error norm and one dt for the whole batch (utils/ode_utils.py:8-9,80-82;
// solver/base_adaptive_solver_rk.py:183-284) -- controller = XDE_CTRL_BATCH.
//
// The reference synchronises with the host ~10 times per attempt for this.  Here the
// whole solve is ONE cooperative launch: a persistent grid in which every thread owns a fixed set of
// trajectories (grid-stride), and the controller is replicated -- every thread computes the same
// accept/reject decision from the same reduced number, so t, dt and the step sequence live in
// registers and never leave the device.  Per attempt:
//   phase A  each thread advances its trajectories through the six stages (state and FSAL derivative
//            from the current buffers, L2 resident; stages in registers), stores the tentative
//            (y1, f1, y_mid) and accumulates sum((err/tol)^2) in fp64;
//   reduce   thread -> warp (xor tree) -> CTA (warp order) -> partial[blockIdx]; ONE grid.sync();
//            every CTA re-sums the partials in the same fixed order: the total is bit-identical
//            everywhere and independent of scheduling (deterministic, no atomics);
//   phase B  replicated controller; on accept the buffers are flipped (no copy) and, if requested
//            output times fall into the step, the quartic dense output is evaluated from
//            (y0, f0, y1, f1, y_mid).
// HBM/L2 traffic per trajectory-attempt: read y, f (16 B at D = 2), write y1, f1, y_mid (24 B).
// select_initial_step (solver/base_adaptive_solver.py:33-72) uses the same machinery (two reductions).
#include <cooperative_groups.h>

#include "xde_common.cuh"

namespace cg = cooperative_groups;

namespace xde {

constexpr int kBatchThreads = 128;
constexpr int kBatchWarps = kBatchThreads / 32;

struct BatchParams {
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
  float *ws;        // [5][B*D]: y[2], f[2], y_mid
  double *partial;  // [2 parities][2 slots][gridDim.x]
};

template <int D, int PRE>
__global__ void __launch_bounds__(kBatchThreads) dopri5_fwd_batch_kernel(const BatchParams p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float smem[];
  __shared__ double s_red[2][kBatchWarps];
  __shared__ double s_tot[2];

  const int H = p.field.h;
  float *sw = smem;
  float *st = smem + SmallRec<D>::floats(H);
  load_small_field<D>(sw, p.field);
  const bool rev = p.t_span[1] < p.t_span[0];  // repair R5: s = -t, f~(s, y) = -f(-s, y)
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) st[i] = rev ? -p.t_span[i] : p.t_span[i];
  __syncthreads();

  const xde_ctrl_opts_t o = p.o;
  const float fsign = rev ? -1.0f : 1.0f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  const long long n = p.B * D;
  float *ybuf[2] = {p.ws, p.ws + n};
  float *fbuf[2] = {p.ws + 2 * n, p.ws + 3 * n};
  float *ymid = p.ws + 4 * n;
  const double n_elems = (double)n;
  const bool leader = (gtid == 0);

  // two fp64 sums: per thread -> per CTA (fixed order) -> partial[] ; after grid.sync every CTA adds the
  // partials in the same fixed order, so the totals are identical in every thread of the grid
  auto reduce2 = [&](double a, double b, int par, double &ta, double &tb) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(XDE_FULL_MASK, a, off);
      b += __shfl_xor_sync(XDE_FULL_MASK, b, off);
    }
    if (lane == 0) {
      s_red[0][warp] = a;
      s_red[1][warp] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double sa = 0.0, sb = 0.0;
#pragma unroll
      for (int w = 0; w < kBatchWarps; ++w) {
        sa += s_red[0][w];
        sb += s_red[1][w];
      }
      double *pp = p.partial + (size_t)par * 2 * gridDim.x;
      pp[blockIdx.x] = sa;
      pp[gridDim.x + blockIdx.x] = sb;
    }
    grid.sync();
    if (warp == 0) {
      const double *pp = p.partial + (size_t)par * 2 * gridDim.x;
      double sa = 0.0, sb = 0.0;
      for (unsigned i = lane; i < gridDim.x; i += 32) {
        sa += pp[i];
        sb += pp[gridDim.x + i];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        sa += __shfl_xor_sync(XDE_FULL_MASK, sa, off);
        sb += __shfl_xor_sync(XDE_FULL_MASK, sb, off);
      }
      if (lane == 0) {
        s_tot[0] = sa;
        s_tot[1] = sb;
      }
    }
    __syncthreads();
    ta = s_tot[0];
    tb = s_tot[1];
    __syncthreads();  // s_tot / s_red are reused by the next reduction
  };

  int par = 0, cur = 0, status = 0;
  unsigned long long n_att = 0, n_acc = 0, n_fe = 0;
  int n_logged = 0;

  // ---- _before_integrate (base_adaptive_solver_rk.py:81-114): solution[0] = y0, f0 = f(t0, y0) ----
  const bool has_first = (o.first_step == o.first_step);
  double a0 = 0.0, a1 = 0.0;
  for (long long b = gtid; b < p.B; b += gstride) {
    float y[D], f[D];
#pragma unroll
    for (int e = 0; e < D; ++e) {
      y[e] = p.y0[b * D + e];
      p.out[b * D + e] = y[e];
      ybuf[0][b * D + e] = y[e];
    }
    mlp_eval_small<D, PRE>(sw, H, y, f);
#pragma unroll
    for (int e = 0; e < D; ++e) {
      f[e] *= fsign;
      fbuf[0][b * D + e] = f[e];
      // select_initial_step part 1 (base_adaptive_solver.py:50-57)
      const float sc = o.atol + fabsf(y[e]) * o.rtol;
      const float v0 = __fdiv_rn(y[e], sc), v1 = __fdiv_rn(f[e], sc);
      a0 += (double)(v0 * v0);
      a1 += (double)(v1 * v1);
    }
  }
  float t0 = st[0], dt;
  if (has_first) {
    dt = o.first_step;
    n_fe = 1;
  } else {
    double t_a, t_b;
    reduce2(a0, a1, par, t_a, t_b);
    par ^= 1;
    const float d0 = fabsf((float)sqrt(t_a / n_elems)), d1 = fabsf((float)sqrt(t_b / n_elems));
    float h0;
    if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f; else h0 = __fdiv_rn(0.01f * d0, d1);
    h0 = fabsf(h0);
    // Euler probe f(t0 + h0, y0 + f0*h0) (base_adaptive_solver.py:60-64)
    double a2 = 0.0;
    for (long long b = gtid; b < p.B; b += gstride) {
      float y[D], f0[D], yi[D], f1[D];
#pragma unroll
      for (int e = 0; e < D; ++e) {
        y[e] = ybuf[0][b * D + e];
        f0[e] = fbuf[0][b * D + e];
        yi[e] = f0[e] * h0 + y[e];
      }
      mlp_eval_small<D, PRE>(sw, H, yi, f1);
#pragma unroll
      for (int e = 0; e < D; ++e) {
        const float sc = o.atol + fabsf(y[e]) * o.rtol;
        const float v = __fdiv_rn(f1[e] * fsign - f0[e], sc);
        a2 += (double)(v * v);
      }
    }
    double t_c, t_unused;
    reduce2(a2, 0.0, par, t_c, t_unused);
    par ^= 1;
    const float d2 = fabsf(__fdiv_rn((float)sqrt(t_c / n_elems), h0));
    float h1;
    if (d1 <= 1e-15f && d2 <= 1e-15f) {
      h1 = fmaxf(1e-6f, h0 * 1e-3f);
    } else {
      const float mx = (d2 > d1) ? d2 : d1;
      const float arg = __fdiv_rn(0.01f, mx);
      h1 = (arg > 0.0f && arg < INFINITY) ? root5(arg) : arg;
    }
    h1 = fabsf(h1);
    dt = fminf(100.0f * h0, h1);
    n_fe = 3;
  }

  // ---- AdaptiveSolver.integrate / AdaptiveRKSolver.step (base_adaptive_solver.py:29-30, _rk.py:116-127) ----
  int i_out = 1, n_steps = 0;
  while (i_out < p.T) {
    if (!(n_steps < o.max_num_steps)) {
      status = XDE_ST_MAX_STEPS;
      break;
    }
    if (!(t0 + dt > t0)) {
      status = XDE_ST_DT_UNDERFLOW;
      break;
    }
    const float t1 = t0 + dt;
    // ---------------- phase A: the six stages of every owned trajectory ----------------
    double sq = 0.0, bad = 0.0;
    for (long long b = gtid; b < p.B; b += gstride) {
      float y0[D], k[7][D], yin[D], fo[D];
      bool fin = true;
#pragma unroll
      for (int e = 0; e < D; ++e) {
        y0[e] = ybuf[cur][b * D + e];
        k[0][e] = fbuf[cur][b * D + e];
        fin = fin && (fabsf(y0[e]) < INFINITY);
      }
      if (!fin) bad += 1.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int e = 0; e < D; ++e) {
          float s = k[0][e] * (DP::beta(i, 0) * dt);
#pragma unroll
          for (int j = 1; j <= i; ++j) s = s + k[j][e] * (DP::beta(i, j) * dt);
          yin[e] = y0[e] + s;
        }
        mlp_eval_small<D, PRE>(sw, H, yin, fo);
#pragma unroll
        for (int e = 0; e < D; ++e) k[i + 1][e] = fo[e] * fsign;
      }
#pragma unroll
      for (int e = 0; e < D; ++e) {
        float s = k[0][e] * (dt * DP::cerr(0));
#pragma unroll
        for (int j = 1; j < 7; ++j) s = s + k[j][e] * (dt * DP::cerr(j));
        const float tol = o.atol + o.rtol * fmaxf(fabsf(y0[e]), fabsf(yin[e]));
        const float v = __fdiv_rn(s, tol);
        sq += (double)(v * v);
        float sm = k[0][e] * (dt * DP::cmid(0));
#pragma unroll
        for (int j = 1; j < 7; ++j) sm = sm + k[j][e] * (dt * DP::cmid(j));
        ybuf[cur ^ 1][b * D + e] = yin[e];
        fbuf[cur ^ 1][b * D + e] = k[6][e];
        ymid[b * D + e] = y0[e] + sm;
      }
    }
    double t_sq, t_bad;
    reduce2(sq, bad, par, t_sq, t_bad);
    par ^= 1;
    if (t_bad > 0.0) {  // assert isfinite(y0).all() (base_adaptive_solver_rk.py:201-203)
      status = XDE_ST_NONFINITE_STATE;
      break;
    }
    // ---------------- phase B: the (replicated) controller ----------------
    const float ratio = fabsf((float)sqrt(t_sq / n_elems));
    bool accept = (ratio <= 1.0f);
    if (dt > o.max_step) accept = false;
    if (dt <= o.min_step) accept = true;
    const float dt_next = next_step_size(dt, ratio, o);
    n_att++;
    n_fe += 6;
    n_steps++;
    if (leader && p.log_records && n_logged < p.log_cap) {
      xde_attempt_t r;
      r.t0 = rev ? -t0 : t0;
      r.dt = rev ? -dt : dt;
      r.ratio = ratio;
      r.accepted = accept ? 1 : 0;
      p.log_records[n_logged] = r;
    }
    n_logged++;
    if (accept) {
      n_acc++;
      if (!(st[i_out] > t1)) {
        // _interp_fit + interp_evaluate (utils/ode_utils.py:28-77) for every requested time in (t0, t1]
        const float two_dt = 2.0f * dt;
        for (long long b = gtid; b < p.B; b += gstride) {
          float ce[D], cd[D], cc[D], cb[D], ca[D];
#pragma unroll
          for (int e = 0; e < D; ++e) {
            const float Y0 = ybuf[cur][b * D + e], F0 = fbuf[cur][b * D + e];
            const float Y1 = ybuf[cur ^ 1][b * D + e], F1 = fbuf[cur ^ 1][b * D + e];
            const float ym = ymid[b * D + e];
            ca[e] = (two_dt * (F1 - F0) - 8.0f * (Y1 + Y0)) + 16.0f * ym;
            cb[e] = ((dt * (5.0f * F0 - 3.0f * F1) + 18.0f * Y0) + 14.0f * Y1) - 32.0f * ym;
            cc[e] = ((dt * (F1 - 4.0f * F0) - 11.0f * Y0) - 5.0f * Y1) + 16.0f * ym;
            cd[e] = dt * F0;
            ce[e] = Y0;
          }
          for (int io = i_out; io < p.T && !(st[io] > t1); ++io) {
            const float x = __fdiv_rn(st[io] - t0, t1 - t0);
#pragma unroll
            for (int e = 0; e < D; ++e) {
              float total = ce[e] + x * cd[e];
              float xp = x * x;
              total = total + xp * cc[e];
              xp = xp * x;
              total = total + xp * cb[e];
              xp = xp * x;
              total = total + xp * ca[e];
              p.out[((long long)io * p.B + b) * D + e] = total;
            }
          }
        }
        while (i_out < p.T && !(st[i_out] > t1)) {
          i_out++;
          n_steps = 0;
        }
      }
      cur ^= 1;
      t0 = t1;
    }
    dt = dt_next;
  }
  if (status != 0) {  // the reference raises: the remaining outputs are undefined -> NaN
    for (long long b = gtid; b < p.B; b += gstride)
      for (int i = i_out; i < p.T; ++i)
#pragma unroll
        for (int e = 0; e < D; ++e) p.out[((long long)i * p.B + b) * D + e] = NAN;
  }
  if (leader) {
    if (p.stats) {  // counted per trajectory, like the per-trajectory controller: trajectory-steps
      p.stats->n_attempts = n_att * (unsigned long long)p.B;
      p.stats->n_accepted = n_acc * (unsigned long long)p.B;
      p.stats->nfe = n_fe * (unsigned long long)p.B;
      p.stats->status = status;
    }
    if (p.log_counts) p.log_counts[0] = n_logged;
  }
}

template <int D, int PRE>
static int launch_batch(BatchParams &p, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (SmallRec<D>::floats(p.field.h) + p.T);
  XDE_REQUIRE(smem <= 200 * 1024, XDE_E_UNSUPPORTED_FIELD, "dopri5 (batch controller): field + t_span exceed shared memory");
  auto kern = dopri5_fwd_batch_kernel<D, PRE>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  XDE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBatchThreads, smem));
  XDE_REQUIRE(per_sm >= 1, XDE_E_CUDA, "dopri5 (batch controller): kernel does not fit an SM");
  // cooperative launch: the grid must be co-resident
  long long want = (p.B + kBatchThreads - 1) / kBatchThreads;
  long long grid = (long long)sm_count() * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  const long long n = p.B * D;
  float *ws = nullptr;
  double *partial = nullptr;
  XDE_CUDA_CHECK(scratch_alloc((void **)&ws, sizeof(float) * 5 * n, stream));
  XDE_CUDA_CHECK(scratch_alloc((void **)&partial, sizeof(double) * 4 * grid, stream));
  p.ws = ws;
  p.partial = partial;
  void *args[] = {(void *)&p};
  cudaError_t e = cudaLaunchCooperativeKernel((void *)kern, dim3((unsigned)grid), dim3(kBatchThreads), args, smem, stream);
  count_launch();
  cudaFreeAsync(ws, stream);
  cudaFreeAsync(partial, stream);
  if (e != cudaSuccess) {
    set_last_error("cooperative launch of dopri5_fwd_batch_kernel failed: %s", cudaGetErrorString(e));
    return XDE_E_CUDA;
  }
  return XDE_OK;
}

template <int D>
static int batch_pre(BatchParams &p, cudaStream_t s) {
  switch (p.field.pre) {
    case XDE_PRE_ID: return launch_batch<D, XDE_PRE_ID>(p, s);
    case XDE_PRE_SQUARE: return launch_batch<D, XDE_PRE_SQUARE>(p, s);
    case XDE_PRE_CUBE: return launch_batch<D, XDE_PRE_CUBE>(p, s);
  }
  set_last_error("unknown pre-activation %d", p.field.pre);
  return XDE_E_BAD_ARG;
}

int dopri5_fwd_batch(const xde_mlp_field_t *field, const float *y0, long long B, const float *t_span, int T,
                     const xde_ctrl_opts_t *opts, float *out, xde_stats_t *stats, const xde_attempt_log_t *log,
                     cudaStream_t s) {
  BatchParams p{};
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
  switch (field->d) {
    case 1: return batch_pre<1>(p, s);
    case 2: return batch_pre<2>(p, s);
    case 3: return batch_pre<3>(p, s);
    case 4: return batch_pre<4>(p, s);
    case 5: return batch_pre<5>(p, s);
    case 6: return batch_pre<6>(p, s);
    case 7: return batch_pre<7>(p, s);
    case 8: return batch_pre<8>(p, s);
  }
  set_last_error("dopri5 (batch controller): state dim D=%d has no fused kernel (supported: 1..8)", field->d);
  return XDE_E_UNSUPPORTED_FIELD;
}

}  // namespace xde
