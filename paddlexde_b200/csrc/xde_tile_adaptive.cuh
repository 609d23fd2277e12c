// xde_tile_adaptive.cuh -- odeint(func=MLP, solver=<embedded Runge-Kutta pair>) forward for LARGE states
// (D = 16 / 32 / 64), one controller per trajectory; the tableau is data (RkTab), its stage count S a template
// argument (Dopri5 6, Bosh3 3, Fehlberg2 2, AdaptiveHeun 1: xde_tile_adaptive*.cu instantiate one S each).
//
// Replaces the same Python loop as xde_dopri5_fwd.cu (solver/base_adaptive_solver.py:24-72,
// solver/base_adaptive_solver_rk.py:116-292, utils/ode_utils.py:28-97) when one trajectory no longer fits a
// thread.  The field evaluation is the register-tiled FP32 GEMM pair of xde_tile.cuh (weights resident in
// shared memory, a tile of TM trajectories per CTA); the Dormand-Prince stages, the state and the controller
// of a trajectory live in the registers of the threads that own its columns:
//   * every thread owns R2 = 2 trajectories x C2 columns (both hidden parities hold identical copies, as in
//     the fixed-grid kernel) and carries the controller scalars (t, dt, output index, counters) of its two
//     trajectories redundantly: all threads of a row compute the same decisions from the same numbers;
//   * the RMS norms (error ratio, select_initial_step) are the only cross-thread quantity of a row: the
//     fp32 squares go through a [TM][D+1] shared-memory tile and every thread sums its rows' D squares
//     SEQUENTIALLY in fp64 -- the oracle's order, so accept/reject sequences are bit-identical;
//   * all rows of a tile advance stage by stage together (2 evaluations of select_initial_step, then 6 per
//     attempt); a row that rejects simply repeats with its own smaller dt, a finished row idles until the
//     slowest row of the tile is done (no refill inside a tile: parity first).
// Arithmetic: identical expressions, in identical order, to adaptive_rk_small_kernel (and, with the Dormand-Prince
// table, to dopri5_fwd_small_kernel).
#pragma once

#include "xde_rk_tab.cuh"
#include "xde_tile.cuh"

namespace xde {

struct AdTileParams {
  xde_mlp_field_t f;
  const float *y0, *t_span;
  float *out;  // [T, B, D]
  long long B;
  int T;
  xde_ctrl_opts_t o;
  xde_stats_t *stats;
  xde_attempt_t *log_records;
  int *log_counts;
  int log_cap;
  // step_t / jump_t (base_adaptive_solver_rk.py:94-114): sorted in integration order, filtered to lie at or after
  // t_span[0] (sort_tvals is the shim's), in t_span's own time; may be null / 0
  const float *step_t, *jump_t;
  int n_step, n_jump;
  unsigned long long *queue;  // next unclaimed tile, shared by the whole grid (zeroed)
  RkTab tab;  // the embedded pair (S stages = the kernel's template argument)
};

template <int S, int D, int H, int TM, int R1, int C1, int R2, int C2>
__global__ void __launch_bounds__(kTileThreads, 1) adaptive_tile_kernel(const AdTileParams p) {
  using G = TileGeom<D, H, TM, R1, C1, R2, C2>;
  constexpr int NS = D + 1;  // row stride of the norm tile (odd: the sequential row sums are conflict free)
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned long long s_cnt[3];
  __shared__ int s_status;
  float *net = smem;
  float *sU = smem + G::net_floats;
  float *sH = sU + 2 * D * TM;  // (the second sU slot of TileGeom::act_floats holds the norm tile)
  float *sN = sU + D * TM;
  float *st = sH + H * TM;
  static_assert(TM * (D + 1) <= D * TM + H * TM, "norm tile must fit behind sU");
  load_net<D, H>(net, p.f);
  const bool rev = p.t_span[1] < p.t_span[0];  // decreasing t_span: s = -t, f~ = -f (repair R5)
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) st[i] = rev ? -p.t_span[i] : p.t_span[i];
  if (threadIdx.x == 0) {
    s_cnt[0] = s_cnt[1] = s_cnt[2] = 0ull;
    s_status = 0;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = lane >> 4;
  const int pidx = warp * 16 + (lane & 15);
  const int rg2 = pidx % G::NRG2, cg2 = pidx / G::NRG2;
  const int c0 = cg2 * C2;
  const int pre = p.f.pre;
  const float fsign = rev ? -1.0f : 1.0f;
  const xde_ctrl_opts_t o = p.o;
  const RkTab &tb = p.tab;
  const long long n_tiles = (p.B + TM - 1) / TM;
  const bool writer = (cg2 == 0);  // the thread that speaks for row e of its pair (stats, log, status)
  unsigned long long n_att = 0, n_acc = 0, n_fe = 0;
  int status = 0;

  // publish a stage input: the half-warp with parity e writes row e of its pair
  auto put_u = [&](const float (&v)[R2][C2]) {
#pragma unroll
    for (int c = 0; c < C2; ++c) sU[(c0 + c) * TM + rg2 * R2 + e] = pre_rt(pre, e ? v[1][c] : v[0][c]);
  };
  // rms over the D components of each of this thread's rows; v holds this thread's columns
  auto row_rms = [&](const float (&v)[R2][C2], float (&res)[R2]) {
    __syncthreads();  // previous readers of sN are done
#pragma unroll
    for (int c = 0; c < C2; ++c) {
      const float x = e ? v[1][c] : v[0][c];
      sN[(rg2 * R2 + e) * NS + c0 + c] = x * x;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R2; ++r) {
      const float *row = sN + (rg2 * R2 + r) * NS;
      double acc = 0.0;
#pragma unroll 8
      for (int k = 0; k < D; ++k) acc += (double)row[k];
      res[r] = rms_from_sumsq(acc, (double)D);
    }
  };

  // tiles are handed out by one grid-wide counter: their durations differ (the slowest row of a tile decides)
  __shared__ long long s_tile;
  while (true) {
    __syncthreads();  // the previous tile's readers of s_tile (and of sU / sH) are done
    if (threadIdx.x == 0) s_tile = (long long)atomicAdd(p.queue, 1ull);
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= n_tiles) break;
    const long long b0 = tile * TM + rg2 * R2;  // first of this thread's R2 trajectories
    float y[R2][C2], k[S + 1][R2][C2], yin[R2][C2];
    float t0[R2], dt[R2];
    int i_out[R2], n_steps[R2], n_logged[R2];
    bool done[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) {
      const long long b = b0 + r;
      const bool ok = b < p.B;
      done[r] = !ok;
      t0[r] = st[0];
      dt[r] = 0.0f;
      i_out[r] = 1;
      n_steps[r] = 0;
      n_logged[r] = 0;
#pragma unroll
      for (int q = 0; q < C2 / 4; ++q) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = *reinterpret_cast<const float4 *>(p.y0 + b * D + c0 + 4 * q);
        y[r][4 * q] = v.x;
        y[r][4 * q + 1] = v.y;
        y[r][4 * q + 2] = v.z;
        y[r][4 * q + 3] = v.w;
        if (ok && r == e) *reinterpret_cast<float4 *>(p.out + b * D + c0 + 4 * q) = v;  // solution[0] = y0
      }
    }

    // ================= select_initial_step (solver/base_adaptive_solver.py:33-72) =================
    {
      float scale[R2][C2], v[R2][C2], d0[R2], d1[R2], d2[R2], h0[R2];
      __syncthreads();  // the previous tile's last readers of sU / sH are done
      put_u(y);
      __syncthreads();
      tile_eval<D, H, TM, R1, C1, R2, C2>(net, sU, sH, k[0]);
#pragma unroll
      for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int c = 0; c < C2; ++c) {
          k[0][r][c] *= fsign;
          scale[r][c] = o.atol + fabsf(y[r][c]) * o.rtol;
          v[r][c] = __fdiv_rn(y[r][c], scale[r][c]);
        }
      row_rms(v, d0);
#pragma unroll
      for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int c = 0; c < C2; ++c) v[r][c] = __fdiv_rn(k[0][r][c], scale[r][c]);
      row_rms(v, d1);
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        d0[r] = fabsf(d0[r]);
        d1[r] = fabsf(d1[r]);
        if (d0[r] < 1e-5f || d1[r] < 1e-5f)
          h0[r] = 1e-6f;
        else
          h0[r] = __fdiv_rn(0.01f * d0[r], d1[r]);
        h0[r] = fabsf(h0[r]);
#pragma unroll
        for (int c = 0; c < C2; ++c) yin[r][c] = k[0][r][c] * h0[r] + y[r][c];  // Euler probe (:60)
      }
      put_u(yin);
      __syncthreads();
      tile_eval<D, H, TM, R1, C1, R2, C2>(net, sU, sH, k[1]);
#pragma unroll
      for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int c = 0; c < C2; ++c) v[r][c] = __fdiv_rn(k[1][r][c] * fsign - k[0][r][c], scale[r][c]);
      row_rms(v, d2);
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        const float dd2 = fabsf(__fdiv_rn(d2[r], h0[r]));
        float h1;
        if (d1[r] <= 1e-15f && dd2 <= 1e-15f) {
          h1 = fmaxf(1e-6f, h0[r] * 1e-3f);
        } else {
          const float mx = (dd2 > d1[r]) ? dd2 : d1[r];
          const float arg = __fdiv_rn(0.01f, mx);
          h1 = (arg > 0.0f && arg < INFINITY) ? rootp(arg, tb.order) : arg;
        }
        h1 = fabsf(h1);
        const float sel = fminf(100.0f * h0[r], h1);
        const bool has_first = (o.first_step == o.first_step);
        dt[r] = has_first ? o.first_step : sel;
        if (writer && r == e && !done[r]) n_fe += has_first ? 1u : 3u;
      }
    }

    // next_*_index = min(bisect.bisect(list, t_span[0]), len - 1) (:109-114); values in solver time
    const float gsign = rev ? -1.0f : 1.0f;
    int step_idx[R2], jump_idx[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) {
      step_idx[r] = jump_idx[r] = 0;
      while (step_idx[r] < p.n_step && !(t0[r] < gsign * p.step_t[step_idx[r]])) step_idx[r]++;
      if (step_idx[r] > p.n_step - 1) step_idx[r] = p.n_step - 1;
      while (jump_idx[r] < p.n_jump && !(t0[r] < gsign * p.jump_t[jump_idx[r]])) jump_idx[r]++;
      if (jump_idx[r] > p.n_jump - 1) jump_idx[r] = p.n_jump - 1;
    }

    // ================= adaptive steps until every row of the tile has produced its last output =================
    while (true) {
      // assertions of _adaptive_step (base_adaptive_solver_rk.py:200-203) + max_num_steps (:120-122)
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        if (done[r]) continue;
        int bad = 0;
        if (!(n_steps[r] < o.max_num_steps)) {
          bad = XDE_ST_MAX_STEPS;
        } else if (!(t0[r] + dt[r] > t0[r])) {
          bad = XDE_ST_DT_UNDERFLOW;
        } else {
          // finiteness of the whole row: columns live in other threads -> through the norm tile below
        }
        if (bad) {
          status = max(status, bad);
          done[r] = true;
          if (r == e) {  // remaining outputs are NaN
            for (int i = i_out[r]; i < p.T; ++i)
#pragma unroll
              for (int c = 0; c < C2; ++c) p.out[((long long)i * p.B + b0 + r) * D + c0 + c] = NAN;
            if (writer && p.log_counts) p.log_counts[b0 + r] = n_logged[r];
          }
        }
      }
      // isfinite(y0).all() per row: a non-finite component makes the row's sum of squares non-finite
      {
        float fin[R2];
        row_rms(y, fin);
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          if (done[r]) continue;
          if (!(fabsf(fin[r]) < INFINITY)) {
            // (an overflowing square of a finite state is reported as non-finite as well: |y| > 1.8e19)
            status = max(status, XDE_ST_NONFINITE_STATE);
            done[r] = true;
            if (r == e) {
              for (int i = i_out[r]; i < p.T; ++i)
#pragma unroll
                for (int c = 0; c < C2; ++c) p.out[((long long)i * p.B + b0 + r) * D + c0 + c] = NAN;
              if (writer && p.log_counts) p.log_counts[b0 + r] = n_logged[r];
            }
          }
        }
      }
      bool all_done = true;
#pragma unroll
      for (int r = 0; r < R2; ++r) all_done = all_done && done[r];
      if (__syncthreads_and(all_done)) break;
      // "Make step, respecting prescribed grid points" (:209-224): step_t first, then jump_t
      float t1v[R2];
      bool on_step[R2], on_jump[R2];
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        if (done[r]) dt[r] = 0.0f;  // idle rows re-evaluate their final state (finite, harmless)
        t1v[r] = t0[r] + dt[r];
        on_step[r] = on_jump[r] = false;
        if (done[r]) continue;
        if (p.n_step > 0) {
          const float nt = gsign * p.step_t[step_idx[r]];
          on_step[r] = (t0[r] < nt) && (nt < t0[r] + dt[r]);
          if (on_step[r]) {
            t1v[r] = nt;
            dt[r] = t1v[r] - t0[r];
          }
        }
        if (p.n_jump > 0) {
          const float nj = gsign * p.jump_t[jump_idx[r]];
          on_jump[r] = (t0[r] < nj) && (nj < t0[r] + dt[r]);
          if (on_jump[r]) {
            on_step[r] = false;
            t1v[r] = nj;
            dt[r] = t1v[r] - t0[r];
          }
        }
      }

      // ---- the S stages (base_adaptive_solver_rk.py:129-181): y0 + sum_j k_j (beta_ij dt), products first ----
#pragma unroll
      for (int i = 0; i < S; ++i) {
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) {
            float s = k[0][r][c] * (tb.beta[i][0] * dt[r]);
#pragma unroll
            for (int j = 1; j <= i; ++j) s = s + k[j][r][c] * (tb.beta[i][j] * dt[r]);
            yin[r][c] = y[r][c] + s;
          }
        put_u(yin);
        __syncthreads();
        tile_eval<D, H, TM, R1, C1, R2, C2>(net, sU, sH, k[i + 1]);
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) k[i + 1][r][c] *= fsign;
      }
      if (!tb.fsal) {  // :172-178: y1 from c_sol unless it equals the last stage input; f1 = k[S] either way
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) {
            float s = k[0][r][c] * (dt[r] * tb.csol[0]);
#pragma unroll
            for (int j = 1; j <= S; ++j) s = s + k[j][r][c] * (dt[r] * tb.csol[j]);
            yin[r][c] = y[r][c] + s;
          }
      }

      // ---- error estimate, ratio (ode_utils.py:80-82), accept / reject, next step ----
      float ratio[R2];
      {
        float v[R2][C2];
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) {
            float s = k[0][r][c] * (dt[r] * tb.cerr[0]);
#pragma unroll
            for (int j = 1; j <= S; ++j) s = s + k[j][r][c] * (dt[r] * tb.cerr[j]);
            const float tol = o.atol + o.rtol * fmaxf(fabsf(y[r][c]), fabsf(yin[r][c]));
            v[r][c] = __fdiv_rn(s, tol);
          }
        row_rms(v, ratio);
      }
      bool jumped[R2];
#pragma unroll
      for (int r = 0; r < R2; ++r) jumped[r] = false;
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        if (done[r]) continue;
        const float t1 = t1v[r];
        const float rt = fabsf(ratio[r]);
        bool accept = (rt <= 1.0f);
        if (dt[r] > o.max_step) accept = false;
        if (dt[r] <= o.min_step) accept = true;
        const bool speak = writer && r == e;
        if (speak) {
          n_att++;
          n_fe += (unsigned)S;
        }
        n_steps[r]++;
        if (p.log_records) {
          if (speak && n_logged[r] < p.log_cap) {
            xde_attempt_t rec;
            rec.t0 = rev ? -t0[r] : t0[r];
            rec.dt = rev ? -dt[r] : dt[r];
            rec.ratio = rt;
            rec.accepted = accept ? 1 : 0;
            p.log_records[(b0 + r) * p.log_cap + n_logged[r]] = rec;
          }
          n_logged[r]++;
        }
        // optimal_step_size(dt, ratio, safety, ifactor, dfactor, self.order).clip(min_step, max_step)
        float dt_next;
        if (rt == 0.0f) {
          dt_next = dt[r] * o.ifactor;
        } else {
          const float dfac = (rt < 1.0f) ? 1.0f : o.dfactor;
          const float pw = (rt > 0.0f && rt < INFINITY) ? rootp(rt, tb.order) : rt;
          dt_next = dt[r] * fminf(o.ifactor, fmaxf(__fdiv_rn(o.safety, pw), dfac));
        }
        dt_next = fminf(fmaxf(dt_next, o.min_step), o.max_step);
        if (accept) {
          if (speak) n_acc++;
          // dense output (interp_fit + interp_evaluate, ode_utils.py:28-77) for the outputs inside [t0, t1];
          // written by the parity that owns the row, four columns at a time
          int io_end = i_out[r];
          while (io_end < p.T && !(st[io_end] > t1)) io_end++;
          if (io_end > i_out[r] && r == e) {
            const float two_dt = 2.0f * dt[r];
#pragma unroll
            for (int q = 0; q < C2 / 4; ++q) {
              float ce[4], cd[4], cc[4], cb[4], ca[4];
#pragma unroll
              for (int z = 0; z < 4; ++z) {
                const int c = 4 * q + z;
                float s = k[0][r][c] * (dt[r] * tb.cmid[0]);
#pragma unroll
                for (int j = 1; j <= S; ++j) s = s + k[j][r][c] * (dt[r] * tb.cmid[j]);
                const float ym = y[r][c] + s;
                const float F0 = k[0][r][c], F1 = k[S][r][c], Y0 = y[r][c], Y1 = yin[r][c];
                ca[z] = (two_dt * (F1 - F0) - 8.0f * (Y1 + Y0)) + 16.0f * ym;
                cb[z] = ((dt[r] * (5.0f * F0 - 3.0f * F1) + 18.0f * Y0) + 14.0f * Y1) - 32.0f * ym;
                cc[z] = ((dt[r] * (F1 - 4.0f * F0) - 11.0f * Y0) - 5.0f * Y1) + 16.0f * ym;
                cd[z] = dt[r] * F0;
                ce[z] = Y0;
              }
              for (int io = i_out[r]; io < io_end; ++io) {
                const float x = __fdiv_rn(st[io] - t0[r], t1 - t0[r]);
                float tot[4];
#pragma unroll
                for (int z = 0; z < 4; ++z) {
                  float total = ce[z] + x * cd[z];
                  float xp = x * x;
                  total = total + xp * cc[z];
                  xp = xp * x;
                  total = total + xp * cb[z];
                  xp = xp * x;
                  total = total + xp * ca[z];
                  tot[z] = total;
                }
                *reinterpret_cast<float4 *>(p.out + ((long long)io * p.B + b0 + r) * D + c0 + 4 * q) =
                    make_float4(tot[0], tot[1], tot[2], tot[3]);
              }
            }
          }
          if (io_end > i_out[r]) n_steps[r] = 0;
          i_out[r] = io_end;
#pragma unroll
          for (int c = 0; c < C2; ++c) {
            y[r][c] = yin[r][c];
            k[0][r][c] = k[S][r][c];
          }
          t0[r] = t1;
          if (on_step[r] && step_idx[r] != p.n_step - 1) step_idx[r]++;
          if (on_jump[r]) {  // past a discontinuity: f1 = self.func(t_next, y_next) (:263-273), below
            if (jump_idx[r] != p.n_jump - 1) jump_idx[r]++;
            jumped[r] = true;
            if (speak) n_fe += 1;
          }
          if (i_out[r] >= p.T) {
            if (speak && p.log_counts) p.log_counts[b0 + r] = n_logged[r];
            done[r] = true;
          }
        }
        dt[r] = dt_next;
      }
      if (p.n_jump > 0) {
        bool any_t = false;
#pragma unroll
        for (int r = 0; r < R2; ++r) any_t = any_t || jumped[r];
        if (__syncthreads_or(any_t)) {  // one more evaluation of the tile; rows that did not jump ignore it
          float F[R2][C2];
          put_u(y);
          __syncthreads();
          tile_eval<D, H, TM, R1, C1, R2, C2>(net, sU, sH, F);
#pragma unroll
          for (int r = 0; r < R2; ++r)
#pragma unroll
            for (int c = 0; c < C2; ++c)
              if (jumped[r]) k[0][r][c] = F[r][c] * fsign;
        }
      }
    }
  }

  // ---- stats: CTA -> global ----
  if (n_att | n_acc | n_fe) {
    atomicAdd(&s_cnt[0], n_att);
    atomicAdd(&s_cnt[1], n_acc);
    atomicAdd(&s_cnt[2], n_fe);
  }
  if (status) atomicMax(&s_status, status);
  __syncthreads();
  if (threadIdx.x == 0 && p.stats) {
    atomicAdd(&p.stats->n_attempts, s_cnt[0]);
    atomicAdd(&p.stats->n_accepted, s_cnt[1]);
    atomicAdd(&p.stats->nfe, s_cnt[2]);
    atomicMax(&p.stats->status, s_status);
  }
}

template <int S, int D, int H, int TM, int R1, int C1, int R2, int C2>
static int launch_ad_tile(const AdTileParams &p, cudaStream_t s) {
  using G = TileGeom<D, H, TM, R1, C1, R2, C2>;
  const size_t smem = G::bytes(1, p.T);
  XDE_REQUIRE(smem <= 227 * 1024, XDE_E_UNSUPPORTED_FIELD,
              "tiled adaptive solver: weights + tiles + t_span need %zu bytes of shared memory (> 227 KB)", smem);
  auto kern = adaptive_tile_kernel<S, D, H, TM, R1, C1, R2, C2>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  XDE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, smem));
  if (per_sm < 1) per_sm = 1;
  long long n_tiles = (p.B + TM - 1) / TM;
  long long grid = (long long)sm_count() * per_sm;  // persistent: a whole number of CTAs per SM
  if (grid > n_tiles) grid = n_tiles;
  AdTileParams q = p;
  XDE_CUDA_CHECK(scratch_alloc((void **)&q.queue, sizeof(unsigned long long), s));
  XDE_CUDA_CHECK(cudaMemsetAsync(q.queue, 0, sizeof(unsigned long long), s));
  kern<<<(unsigned)grid, kTileThreads, smem, s>>>(q);
  count_launch();
  cudaError_t le = cudaGetLastError();
  cudaFreeAsync(q.queue, s);
  XDE_CUDA_CHECK(le);
  return XDE_OK;
}

// geometries with R2 x C2 <= 16 state values per thread: the S + 1 stages + state + stage input stay in registers
template <int S>
static int ad_tile_dispatch(const AdTileParams &p, cudaStream_t s) {
  const int D = p.f.d, H = p.f.h;
  if (D == 64 && H == 256) return launch_ad_tile<S, 64, 256, 32, 4, 8, 2, 8>(p, s);
  if (D == 64 && H == 128) return launch_ad_tile<S, 64, 128, 32, 4, 4, 2, 8>(p, s);
  if (D == 32 && H == 256) return launch_ad_tile<S, 32, 256, 32, 4, 8, 2, 4>(p, s);
  if (D == 32 && H == 128) return launch_ad_tile<S, 32, 128, 64, 4, 8, 2, 8>(p, s);
  if (D == 32 && H == 64) return launch_ad_tile<S, 32, 64, 64, 4, 4, 2, 8>(p, s);
  if (D == 16 && H == 64) return launch_ad_tile<S, 16, 64, 64, 4, 4, 2, 4>(p, s);
  set_last_error("tiled adaptive solver: no kernel for D=%d H=%d (D=64: H in {128,256}; D=32: H in {64,128,256}; D=16: "
                 "H=64; small states D in 1..8: any H)", D, H);
  return XDE_E_UNSUPPORTED_FIELD;
}

}  // namespace xde
