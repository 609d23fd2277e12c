// xde_dopri5_adj.cu -- OdeintAdjointMethod.backward for the fused MLP field, per-trajectory controller.
//
// Replaces functional/odeint_adjoint.py:47-167: for every output segment [t_i, t_{i-1}] a fresh
// reverse-time dopri5 solve (select_initial_step + adaptive steps + dense output at t_{i-1}) of the
// augmented state (y, a, g_theta), dynamics augmented_dynamics :89-124, then y <- y_ans[i-1],
// a += grad_y[i-1].  Repairs R4-R6 (flat state, reverse time as s = -t, norm over the flat state).
//
// Design.  The parameter-gradient state g_theta (P = 2DH+H+D values) is the problem: it is an outer
// product per trajectory and stage, summed over the batch, and a rejected step must not contribute.
// Layout that makes this cheap:
//   * a warp works on G trajectory "slots" at once; lane l owns hidden units j = l, l+32, ... :
//     their W1 columns / W2 rows live in registers, so the parameter-gradient accumulators for
//     those units are lane-private (no cross-lane traffic for g_theta at all);
//   * per stage, each slot's (pre(y), a) is broadcast by shuffles, every lane evaluates its hidden
//     units (z, tanh, dh, dz), and the 2D partial sums per slot (f and du) of all G slots are reduced
//     together by ONE multi-value butterfly (2D*G values in ~2D*G shuffles) that leaves each total
//     in the lane that owns that state component -- "owner" lanes keep s0, the 7 stages, t, dt and
//     run the controller for their slot, so there is one controller per trajectory;
//   * g_theta is not integrated as state: because it never feeds back, its value at the segment
//     end is sum_i W_i * k_i^theta with scalar weights W_i known when the attempt starts
//     (dt*c_sol_i, or the dense-output polynomial weights for the last step of a segment); each
//     lane accumulates a tentative per-slot sum and commits it (fp64) only when the slot's
//     controller accepts the step.  The stage-0 term reuses the FSAL evaluation: it is seeded into
//     the next attempt's tentative sum as soon as the next dt is known.
//   * slots advance as independent state machines (INIT: f0, probe; ATTEMPT: stages 1..6), one
//     field evaluation per slot per round; finished slots refill from a per-CTA trajectory queue.
// Supported norm: the adjoint seminorm (functional/odeint_adjoint.py:301-309) -- with one
// controller per trajectory g_theta is a per-trajectory partial integral.
#include "xde_common.cuh"

namespace xde {

constexpr int kAdjThreads = 128;

struct AdjParams {
  xde_mlp_field_t field;
  const float *t_span, *y_ans, *grad_y;
  double *gacc;       // [P] fp64 accumulator (zeroed)
  float *adj_y0;      // [B,D] or null
  long long B;
  int T;
  xde_ctrl_opts_t o;
  xde_stats_t *stats;
  xde_attempt_t *log_records;
  int *log_counts;
  int log_cap;
  long long chunk;
};

struct AdjTables {  // shared-memory coefficient tables (per-lane stage lookup)
  float beta[6][8];  // beta[i][j], stage i+1 input
  float wsol[8], wa[8], wb[8], wc[8];
};

enum { AM_IDLE = 0, AM_INIT = 1, AM_ATTEMPT = 2, AM_DONE = 3 };
// events broadcast from the owner lanes to the hidden-unit role for the theta bookkeeping
enum { EV_NONE = 0, EV_STAGE = 1, EV_ACCEPT_CONT = 2, EV_ACCEPT_END = 3, EV_REJECT = 4, EV_INIT0 = 5, EV_INIT1 = 6 };

template <int N>
struct Log2 { static constexpr int v = 1 + Log2<N / 2>::v; };
template <>
struct Log2<1> { static constexpr int v = 0; };

// Multi-value butterfly: v[0..NV) per lane -> the lane-sum of value (lane >> (5-log2 NV)) in every
// lane.  Arithmetic tree per value: partner distance 16, 8, 4, 2, 1.
template <int NV>
__device__ __forceinline__ float butterfly_reduce(float (&v)[NV], int lane) {
  constexpr int LOG = Log2<NV>::v;
  int n = NV;
#pragma unroll
  for (int step = 0; step < LOG; ++step) {
    const int dist = 16 >> step;
    const bool hi = (lane & dist) != 0;
    n >>= 1;
#pragma unroll
    for (int i = 0; i < NV / 2; ++i) {
      if (i < n) {
        const float keep = hi ? v[i + n] : v[i];
        const float send = hi ? v[i] : v[i + n];
        v[i] = keep + __shfl_xor_sync(XDE_FULL_MASK, send, dist);
      }
    }
  }
  float r = v[0];
#pragma unroll
  for (int step = LOG; step < 5; ++step) r = r + __shfl_xor_sync(XDE_FULL_MASK, r, 16 >> step);
  return r;
}

template <int D, int HPL, int PRE, int G>
__global__ void __launch_bounds__(kAdjThreads) dopri5_adj_kernel(const AdjParams p) {
  constexpr int C = 2 * D;
  constexpr int NV = G * C;
  static_assert(NV == 8 || NV == 16 || NV == 32, "G*2D must be 8, 16 or 32");
  constexpr int SH = 5 - Log2<NV>::v;  // owner duplication: lane l owns value l >> SH
  constexpr int NTH = (2 * D + 1) * HPL;  // lane-private theta values

  extern __shared__ __align__(16) float smem[];
  __shared__ AdjTables tb;
  __shared__ unsigned long long s_next;
  __shared__ unsigned long long s_cnt[3];
  __shared__ int s_status;
  float *st = smem;  // solver times s_i = tsign * t_i

  const int H = p.field.h;
  const int lane = threadIdx.x & 31;
  // direction of the backward sweep: t_span increasing (usual) -> integrate s = -t
  const float tsign = (p.t_span[1] > p.t_span[0]) ? -1.0f : 1.0f;
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) st[i] = tsign * p.t_span[i];
  if (threadIdx.x < 8) {
    const int i = threadIdx.x;
    // dense-output weights of stage i at x in (0,1]:  W_i = dt*(x*[i==0] + x^2*wa + x^3*wb + x^4*wc)
    // (interp_fit/interp_evaluate, utils/ode_utils.py:28-77, expanded in the stage values)
    float cs = 0.f, cm = 0.f;
    if (i < 7) {
      const float csv[7] = {DP::csol(0), DP::csol(1), DP::csol(2), DP::csol(3), DP::csol(4), DP::csol(5), DP::csol(6)};
      const float cmv[7] = {DP::cmid(0), DP::cmid(1), DP::cmid(2), DP::cmid(3), DP::cmid(4), DP::cmid(5), DP::cmid(6)};
      cs = csv[i];
      cm = cmv[i];
    }
    const float d0 = (i == 0) ? 1.f : 0.f, d6 = (i == 6) ? 1.f : 0.f;
    tb.wsol[i] = cs;
    tb.wa[i] = fmaf(16.0f, cm, fmaf(-5.0f, cs, d6 - 4.0f * d0));
    tb.wb[i] = fmaf(-32.0f, cm, fmaf(14.0f, cs, 5.0f * d0 - 3.0f * d6));
    tb.wc[i] = fmaf(16.0f, cm, fmaf(-8.0f, cs, 2.0f * d6 - 2.0f * d0));
  }
  if (threadIdx.x < 48) {
    const int i = threadIdx.x / 8, j = threadIdx.x % 8;
    float b = 0.f;
    const float bt[6][6] = {
        {DP::beta(0, 0), 0, 0, 0, 0, 0},
        {DP::beta(1, 0), DP::beta(1, 1), 0, 0, 0, 0},
        {DP::beta(2, 0), DP::beta(2, 1), DP::beta(2, 2), 0, 0, 0},
        {DP::beta(3, 0), DP::beta(3, 1), DP::beta(3, 2), DP::beta(3, 3), 0, 0},
        {DP::beta(4, 0), DP::beta(4, 1), DP::beta(4, 2), DP::beta(4, 3), DP::beta(4, 4), 0},
        {DP::beta(5, 0), DP::beta(5, 1), DP::beta(5, 2), DP::beta(5, 3), DP::beta(5, 4), DP::beta(5, 5)}};
    if (j < 6) b = bt[i][j];
    tb.beta[i][j] = b;
  }
  const long long c0 = (long long)blockIdx.x * p.chunk;
  const long long c1 = (c0 + p.chunk < p.B) ? c0 + p.chunk : p.B;
  if (threadIdx.x == 0) {
    s_next = (unsigned long long)c0;
    s_cnt[0] = s_cnt[1] = s_cnt[2] = 0ull;
    s_status = 0;
  }
  __syncthreads();

  // ---- hidden-unit role: this lane's units j = lane + 32 q (zero weights for j >= H) ----
  float w1r[D][HPL], b1r[HPL], w2r[HPL][D];
#pragma unroll
  for (int q = 0; q < HPL; ++q) {
    const int j = lane + 32 * q;
    const bool ok = j < H;
    b1r[q] = ok ? p.field.b1[j] : 0.0f;
#pragma unroll
    for (int k = 0; k < D; ++k) w1r[k][q] = ok ? p.field.w1[k * H + j] : 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) w2r[q][d] = ok ? p.field.w2[j * D + d] : 0.0f;
  }
  double acc[NTH];  // committed parameter-gradient sums of this lane's units (fp64)
  float S[G][NTH];  // tentative sums of the attempt in flight, per slot
  float hk[G][HPL], dzk[G][HPL];
#pragma unroll
  for (int i = 0; i < NTH; ++i) acc[i] = 0.0;
#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int i = 0; i < NTH; ++i) S[g][i] = 0.0f;
#pragma unroll
    for (int q = 0; q < HPL; ++q) hk[g][q] = dzk[g][q] = 0.0f;
  }

  // ---- owner role: this lane owns state component `cown` of slot `gown` ----
  const int vown = lane >> SH;
  const int gown = vown / C, cown = vown % C;
  const bool is_y = cown < D;
  const int comp = is_y ? cown : cown - D;
  const bool primary = (lane & ((1 << SH) - 1)) == 0;
  const int slot_lead = (gown * C) << SH;          // lane holding component 0 of my slot
  const int partner_y = (gown * C + comp) << SH;   // lane holding y[comp] of my slot
  const float b2own = is_y ? p.field.b2[comp] : 0.0f;
  const xde_ctrl_opts_t o = p.o;

  int mode = AM_IDLE, stage = 0, seg = 0, n_steps = 0, n_logged = 0;
  long long traj = -1;
  float s0 = 0.f, kk[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float t0 = 0.f, dt = 0.f, te = 0.f, xfin = 0.f;
  bool fin = false;  // the attempt in flight reaches the segment end if accepted
  float scale = 1.f, h0 = 0.f, d1 = 0.f;
  float Sb2 = 0.f;   // tentative gb2 sum (a-component owners)
  double accb2 = 0.0;
  unsigned n_att = 0, n_acc = 0, n_fe = 0;
  int status = 0;

  // theta weight of the eval at `stg` (0..6) for an attempt (dtv, finv, xv); sign folded in
  auto theta_w = [&](int stg, float dtv, bool finv, float xv) -> float {
    float w;
    if (finv) {
      const float x2 = xv * xv, x3 = x2 * xv, x4 = x3 * xv;
      const float lin = (stg == 0) ? xv : 0.0f;
      w = dtv * (((lin + x2 * tb.wa[stg]) + x3 * tb.wb[stg]) + x4 * tb.wc[stg]);
    } else {
      w = dtv * tb.wsol[stg];
    }
    return -tsign * w;  // d g_theta / ds = -tsign * vjp_theta(a)
  };

  while (true) {
    // ================= refill idle slots =================
    {
      const bool need = (mode == AM_IDLE) && (cown == 0) && primary;
      const unsigned m = __ballot_sync(XDE_FULL_MASK, need);
      const unsigned many = __ballot_sync(XDE_FULL_MASK, mode == AM_IDLE);
      if (many) {
        unsigned long long base = 0;
        if (m) {
          const int leader = __ffs(m) - 1;
          if (lane == leader) base = atomicAdd(&s_next, (unsigned long long)__popc(m));
          base = __shfl_sync(XDE_FULL_MASK, base, leader);
        }
        long long cand = (long long)base + __popc(m & ((1u << lane) - 1u));
        cand = __shfl_sync(XDE_FULL_MASK, cand, slot_lead);
        if (mode == AM_IDLE) {
          if (cand < c1) {
            traj = cand;
            seg = p.T - 1;
            const long long src = ((long long)seg * p.B + traj) * D + comp;
            s0 = is_y ? p.y_ans[src] : p.grad_y[src];  // aug_state = [y_ans[-1], grad_y[-1]] (:75-82)
            t0 = st[seg];
            te = st[seg - 1];
            mode = AM_INIT;
            stage = 0;
            n_steps = 0;
            n_logged = 0;
            Sb2 = 0.f;
          } else {
            mode = AM_DONE;
          }
        }
      }
    }
    if (__all_sync(XDE_FULL_MASK, mode == AM_DONE)) break;

    const bool att = (mode == AM_ATTEMPT), ini = (mode == AM_INIT);

    // ================= (1) owner: input of this round's evaluation =================
    float t1 = t0 + dt;
    bool live = att || ini;
    {
      // assertions of _adaptive_step (base_adaptive_solver_rk.py:200-203) + max_num_steps (:120-122),
      // checked when an attempt starts.  Non-finite state is a slot-wide condition (convergent ballot).
      const unsigned nf = __ballot_sync(XDE_FULL_MASK, !(fabsf(s0) < INFINITY));
      const unsigned slot_mask = (((C << SH) == 32) ? 0xffffffffu : ((1u << (C << SH)) - 1u)) << slot_lead;
      if (att && stage == 1) {
        int bad = 0;
        if (!(n_steps < o.max_num_steps)) bad = XDE_ST_MAX_STEPS;
        else if (!(t0 + dt > t0)) bad = XDE_ST_DT_UNDERFLOW;
        else if (nf & slot_mask) bad = XDE_ST_NONFINITE_STATE;
        if (bad) {
          status = max(status, bad);
          if (!is_y && p.adj_y0 && primary) p.adj_y0[traj * D + comp] = NAN;
          if (p.log_counts && cown == 0 && primary) p.log_counts[traj] = n_logged;
          mode = AM_IDLE;
          live = false;
        }
      }
    }
    float yin;
    {
      const int bi = (stage >= 1 && stage <= 6) ? stage - 1 : 0;
      float s = kk[0] * (tb.beta[bi][0] * dt);
#pragma unroll
      for (int j = 1; j < 6; ++j)
        if (j < stage) s = s + kk[j] * (tb.beta[bi][j] * dt);
      const float y_att = s0 + s;
      const float y_ini = (stage == 0) ? s0 : (kk[0] * h0 + s0);
      yin = att ? y_att : y_ini;
    }
    const float bc_own = is_y ? pre_act<PRE>(yin) : yin;  // y owners broadcast u = pre(y), a owners a
    const float ypart = __shfl_sync(XDE_FULL_MASK, yin, partner_y);
    const float dpre = pre_act_grad<PRE>(ypart);

    // ================= (2) field + VJP evaluation, slot by slot (warp-uniform code) =================
    float v[NV];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float u[D], a[D];
#pragma unroll
      for (int k = 0; k < D; ++k) u[k] = __shfl_sync(XDE_FULL_MASK, bc_own, ((g * C + k) << SH));
#pragma unroll
      for (int d = 0; d < D; ++d) a[d] = __shfl_sync(XDE_FULL_MASK, bc_own, ((g * C + D + d) << SH));
      // the INIT probe must not overwrite the retained (h, dz) of the f0 evaluation
      const int keepflag = __shfl_sync(XDE_FULL_MASK, (int)(ini && stage == 1), (g * C) << SH);
      float pf[D], pdu[D];
#pragma unroll
      for (int q = 0; q < HPL; ++q) {
        float z = u[0] * w1r[0][q];
#pragma unroll
        for (int k = 1; k < D; ++k) z = fmaf(u[k], w1r[k][q], z);
        const float h = tanh_rat(z + b1r[q]);
        float dh = a[0] * w2r[q][0];
#pragma unroll
        for (int d = 1; d < D; ++d) dh = fmaf(a[d], w2r[q][d], dh);
        const float t = h * h;
        const float sgrad = 1.0f - t;
        const float dz = dh * sgrad;
#pragma unroll
        for (int d = 0; d < D; ++d) pf[d] = (q == 0) ? h * w2r[q][d] : fmaf(h, w2r[q][d], pf[d]);
#pragma unroll
        for (int k = 0; k < D; ++k) pdu[k] = (q == 0) ? dz * w1r[k][q] : fmaf(dz, w1r[k][q], pdu[k]);
        if (!keepflag) {
          hk[g][q] = h;
          dzk[g][q] = dz;
        }
      }
#pragma unroll
      for (int d = 0; d < D; ++d) {
        v[g * C + d] = pf[d];
        v[g * C + D + d] = pdu[d];
      }
    }
    // ================= (3) reduce: totals land in the owner lanes =================
    const float tot = butterfly_reduce<NV>(v, lane);
    // solver-time dynamics: dy/ds = tsign * f ; da/ds = -tsign * vjp_y(a)
    const float fo = is_y ? tsign * (tot + b2own) : (-tsign) * (tot * dpre);

    // ================= (4) owner: state machine =================
    int ev = EV_NONE;
    float w_eval = 0.f;  // theta weight of the evaluation just done
    float w_seed = 0.f;  // theta weight of stage 0 of the next attempt (seeding)
    float sb2_term = 0.f;
    // rms over the D components of my group (y or a), fp64 accumulation; mixed seminorm =
    // max(|g_t| = 0, rms(y), rms(a)) with Python max semantics (functional/odeint_adjoint.py:304-307)
    auto semi_norm = [&](float val) -> float {
      double sq = (double)(val * val);
#pragma unroll
      for (int off = 1; off < D; off <<= 1) sq += __shfl_xor_sync(XDE_FULL_MASK, sq, off << SH);
      const float mine = rms_from_sumsq(sq, (double)D);
      const float other = __shfl_xor_sync(XDE_FULL_MASK, mine, D << SH);
      const float ny = is_y ? mine : other, na = is_y ? other : mine;
      float best = 0.0f;
      if (ny > best) best = ny;
      if (na > best) best = na;
      return best;
    };

    // norm arguments are formed per lane; the reductions themselves run in convergent code
    float argA = 0.f, argB = 0.f, scale_new = scale;
    const bool is_init0 = live && ini && stage == 0;
    const bool is_init1 = live && ini && stage == 1;
    const bool is_last = live && att && stage == 6;
    if (is_init0) {
      scale_new = o.atol + fabsf(s0) * o.rtol;  // select_initial_step (base_adaptive_solver.py:50)
      argA = __fdiv_rn(s0, scale_new);
      argB = __fdiv_rn(fo, scale_new);
    } else if (is_init1) {
      argA = __fdiv_rn(fo - kk[0], scale);
    } else if (is_last) {
      // error estimate and tolerance (base_adaptive_solver_rk.py:180; ode_utils.py:80-82), k6 = fo
      float e = kk[0] * (dt * DP::cerr(0));
#pragma unroll
      for (int j = 1; j < 6; ++j) e = e + kk[j] * (dt * DP::cerr(j));
      e = e + fo * (dt * DP::cerr(6));
      const float tol = o.atol + o.rtol * fmaxf(fabsf(s0), fabsf(yin));
      argA = __fdiv_rn(e, tol);
    }
    const float nA = semi_norm(argA);
    float nB = 0.f;
    if (__any_sync(XDE_FULL_MASK, is_init0)) nB = semi_norm(argB);

    if (is_init0) {
      // _before_integrate f0 + select_initial_step part 1 (base_adaptive_solver.py:44-57)
      kk[0] = fo;
      scale = scale_new;
      const float d0 = fabsf(nA);
      d1 = fabsf(nB);
      if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f; else h0 = __fdiv_rn(0.01f * d0, d1);
      h0 = fabsf(h0);
      stage = 1;
      ev = EV_INIT0;
    } else if (is_init1) {
      const float d2 = fabsf(__fdiv_rn(nA, h0));
      float h1;
      if (d1 <= 1e-15f && d2 <= 1e-15f) {
        h1 = fmaxf(1e-6f, h0 * 1e-3f);
      } else {
        const float mx = (d2 > d1) ? d2 : d1;
        const float arg = __fdiv_rn(0.01f, mx);
        h1 = (arg > 0.0f && arg < INFINITY) ? root5(arg) : arg;
      }
      h1 = fabsf(h1);
      const bool has_first = (o.first_step == o.first_step);
      dt = has_first ? o.first_step : fminf(100.0f * h0, h1);
      if (primary && cown == 0) n_fe += has_first ? 1u : 3u;
      mode = AM_ATTEMPT;
      stage = 1;
      const float t1n = t0 + dt;
      fin = !(te > t1n);
      xfin = fin ? __fdiv_rn(te - t0, t1n - t0) : 0.f;
      w_seed = theta_w(0, dt, fin, xfin);
      ev = EV_INIT1;
    } else if (live && att) {
      w_eval = theta_w(stage, dt, fin, xfin);
      if (!is_y) sb2_term = w_eval * yin;
      if (stage < 6) {
#pragma unroll
        for (int j = 1; j < 6; ++j)
          if (j == stage) kk[j] = fo;
        stage++;
        ev = EV_STAGE;
      } else {
        kk[6] = fo;
        const float ratio = fabsf(nA);
        bool accept = (ratio <= 1.0f);
        if (dt > o.max_step) accept = false;
        if (dt <= o.min_step) accept = true;
        const float dt_next = next_step_size(dt, ratio, o);
        if (primary && cown == 0) {
          n_att++;
          n_fe += 6;
          if (p.log_records) {
            if (n_logged < p.log_cap) {
              xde_attempt_t r;
              r.t0 = tsign * t0;
              r.dt = tsign * dt;
              r.ratio = ratio;
              r.accepted = accept ? 1 : 0;
              p.log_records[traj * p.log_cap + n_logged] = r;
            }
          }
          if (accept) n_acc++;
        }
        n_logged++;
        n_steps++;
        if (accept) {
          if (fin) {
            // dense output at the segment end (interp_fit + interp_evaluate), then
            // y <- y_ans[i-1], a += grad_y[i-1] (functional/odeint_adjoint.py:153-159)
            float sm = kk[0] * (dt * DP::cmid(0));
#pragma unroll
            for (int j = 1; j < 7; ++j) sm = sm + kk[j] * (dt * DP::cmid(j));
            const float ym = s0 + sm;
            const float F0 = kk[0], F1 = kk[6], Y0 = s0, Y1 = yin;
            const float two_dt = 2.0f * dt;
            const float ca = (two_dt * (F1 - F0) - 8.0f * (Y1 + Y0)) + 16.0f * ym;
            const float cb = ((dt * (5.0f * F0 - 3.0f * F1) + 18.0f * Y0) + 14.0f * Y1) - 32.0f * ym;
            const float cc = ((dt * (F1 - 4.0f * F0) - 11.0f * Y0) - 5.0f * Y1) + 16.0f * ym;
            const float cd = dt * F0;
            const float x = xfin;
            float total = Y0 + x * cd;
            float xp = x * x;
            total = total + xp * cc;
            xp = xp * x;
            total = total + xp * cb;
            xp = xp * x;
            total = total + xp * ca;
            seg -= 1;
            const long long src = ((long long)seg * p.B + traj) * D + comp;
            s0 = is_y ? p.y_ans[src] : (total + p.grad_y[src]);
            n_steps = 0;
            if (seg == 0) {
              if (!is_y && p.adj_y0 && primary) p.adj_y0[traj * D + comp] = s0;
              if (p.log_counts && cown == 0 && primary) p.log_counts[traj] = n_logged;
              mode = AM_IDLE;
            } else {
              t0 = st[seg];
              te = st[seg - 1];
              mode = AM_INIT;
              stage = 0;
            }
            ev = EV_ACCEPT_END;
          } else {
            s0 = yin;
            kk[0] = kk[6];
            t0 = t1;
            dt = dt_next;
            stage = 1;
            const float t1n = t0 + dt;
            fin = !(te > t1n);
            xfin = fin ? __fdiv_rn(te - t0, t1n - t0) : 0.f;
            w_seed = theta_w(0, dt, fin, xfin);
            ev = EV_ACCEPT_CONT;
          }
        } else {
          dt = dt_next;
          stage = 1;
          const float t1n = t0 + dt;
          fin = !(te > t1n);
          xfin = fin ? __fdiv_rn(te - t0, t1n - t0) : 0.f;
          w_seed = theta_w(0, dt, fin, xfin);
          ev = EV_REJECT;
        }
      }
    }
    // gb2: lane-private to the a-component owners (gb2[d] = sum W * a_d)
    if (!is_y) {
      if (ev == EV_STAGE) Sb2 += sb2_term;
      else if (ev == EV_ACCEPT_CONT || ev == EV_ACCEPT_END) {
        accb2 += (double)(Sb2 + sb2_term);
        Sb2 = (ev == EV_ACCEPT_CONT) ? w_seed * s0 : 0.f;
      } else if (ev == EV_REJECT || ev == EV_INIT1) {
        Sb2 = w_seed * s0;
      }
    }
    // start-state broadcast values for re-seeding after INIT / reject
    const float bc_start = is_y ? pre_act<PRE>(s0) : s0;

    // ================= (5) theta bookkeeping, slot by slot (warp-uniform code) =================
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int lead = (g * C) << SH;
      const int evg = __shfl_sync(XDE_FULL_MASK, ev, lead);
      if (evg == EV_NONE || evg == EV_INIT0) continue;
      const float wev = __shfl_sync(XDE_FULL_MASK, w_eval, lead);
      const float wsd = __shfl_sync(XDE_FULL_MASK, w_seed, lead);
      float u[D], a[D];
#pragma unroll
      for (int k = 0; k < D; ++k) u[k] = __shfl_sync(XDE_FULL_MASK, bc_own, ((g * C + k) << SH));
#pragma unroll
      for (int d = 0; d < D; ++d) a[d] = __shfl_sync(XDE_FULL_MASK, bc_own, ((g * C + D + d) << SH));
      if (evg == EV_STAGE || evg == EV_ACCEPT_CONT || evg == EV_ACCEPT_END) {
        // S += W_i * k_i^theta for the evaluation of this round
#pragma unroll
        for (int q = 0; q < HPL; ++q) {
          const float wdz = wev * dzk[g][q], wh = wev * hk[g][q];
#pragma unroll
          for (int k = 0; k < D; ++k) S[g][k * HPL + q] = fmaf(u[k], wdz, S[g][k * HPL + q]);
          S[g][D * HPL + q] += wdz;
#pragma unroll
          for (int d = 0; d < D; ++d) S[g][(D + 1) * HPL + q * D + d] = fmaf(a[d], wh, S[g][(D + 1) * HPL + q * D + d]);
        }
      }
      if (evg == EV_ACCEPT_CONT || evg == EV_ACCEPT_END) {
#pragma unroll
        for (int i = 0; i < NTH; ++i) acc[i] += (double)S[g][i];
      }
      if (evg == EV_ACCEPT_END) {
#pragma unroll
        for (int i = 0; i < NTH; ++i) S[g][i] = 0.0f;
      } else if (evg == EV_ACCEPT_CONT) {
        // the stage-6 point is the next attempt's stage 0 (FSAL): seed with its weight
#pragma unroll
        for (int q = 0; q < HPL; ++q) {
          const float wdz = wsd * dzk[g][q], wh = wsd * hk[g][q];
#pragma unroll
          for (int k = 0; k < D; ++k) S[g][k * HPL + q] = u[k] * wdz;
          S[g][D * HPL + q] = wdz;
#pragma unroll
          for (int d = 0; d < D; ++d) S[g][(D + 1) * HPL + q * D + d] = a[d] * wh;
        }
      } else if (evg == EV_REJECT || evg == EV_INIT1) {
        // stage-0 term at the (unchanged) start state.  INIT1: (h, dz) of the f0 evaluation were
        // retained.  REJECT: they were overwritten by stages 1..6 -> re-evaluate this lane's units.
        float us[D], as[D];
#pragma unroll
        for (int k = 0; k < D; ++k) us[k] = __shfl_sync(XDE_FULL_MASK, bc_start, ((g * C + k) << SH));
#pragma unroll
        for (int d = 0; d < D; ++d) as[d] = __shfl_sync(XDE_FULL_MASK, bc_start, ((g * C + D + d) << SH));
        if (evg == EV_REJECT) {
#pragma unroll
          for (int q = 0; q < HPL; ++q) {
            float z = us[0] * w1r[0][q];
#pragma unroll
            for (int k = 1; k < D; ++k) z = fmaf(us[k], w1r[k][q], z);
            const float h = tanh_rat(z + b1r[q]);
            float dh = as[0] * w2r[q][0];
#pragma unroll
            for (int d = 1; d < D; ++d) dh = fmaf(as[d], w2r[q][d], dh);
            hk[g][q] = h;
            dzk[g][q] = dh * (1.0f - h * h);
          }
        }
#pragma unroll
        for (int q = 0; q < HPL; ++q) {
          const float wdz = wsd * dzk[g][q], wh = wsd * hk[g][q];
#pragma unroll
          for (int k = 0; k < D; ++k) S[g][k * HPL + q] = us[k] * wdz;
          S[g][D * HPL + q] = wdz;
#pragma unroll
          for (int d = 0; d < D; ++d) S[g][(D + 1) * HPL + q * D + d] = as[d] * wh;
        }
      }
    }
  }

  // ================= epilogue: parameter gradients and stats =================
#pragma unroll
  for (int q = 0; q < HPL; ++q) {
    const int j = lane + 32 * q;
    if (j < H) {
#pragma unroll
      for (int k = 0; k < D; ++k) atomicAdd(&p.gacc[k * H + j], acc[k * HPL + q]);
      atomicAdd(&p.gacc[D * H + j], acc[D * HPL + q]);
#pragma unroll
      for (int d = 0; d < D; ++d) atomicAdd(&p.gacc[D * H + H + j * D + d], acc[(D + 1) * HPL + q * D + d]);
    }
  }
  if (!is_y && primary) atomicAdd(&p.gacc[D * H + H + H * D + comp], accb2);
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

__global__ void adj_cast_kernel(const double *__restrict__ a, float *__restrict__ o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = (float)a[i];
}

template <int D, int HPL, int PRE, int G>
static int launch_adj(const AdjParams &p, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (size_t)p.T;
  XDE_REQUIRE(smem <= 160 * 1024, XDE_E_UNSUPPORTED_FIELD, "adjoint: t_span too long for shared memory");
  auto kern = dopri5_adj_kernel<D, HPL, PRE, G>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  XDE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kAdjThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const long long slots_per_cta = (long long)(kAdjThreads / 32) * G;
  long long want = (p.B + slots_per_cta - 1) / slots_per_cta;
  long long grid = (long long)sm_count() * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  AdjParams q = p;
  q.chunk = (p.B + grid - 1) / grid;
  grid = (p.B + q.chunk - 1) / q.chunk;
  kern<<<(unsigned)grid, kAdjThreads, smem, stream>>>(q);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  return XDE_OK;
}

template <int D, int HPL, int G>
static int adj_pre(const AdjParams &p, cudaStream_t s) {
  switch (p.field.pre) {
    case XDE_PRE_ID: return launch_adj<D, HPL, XDE_PRE_ID, G>(p, s);
    case XDE_PRE_SQUARE: return launch_adj<D, HPL, XDE_PRE_SQUARE, G>(p, s);
    case XDE_PRE_CUBE: return launch_adj<D, HPL, XDE_PRE_CUBE, G>(p, s);
  }
  set_last_error("unknown pre-activation %d", p.field.pre);
  return XDE_E_BAD_ARG;
}

}  // namespace xde

extern "C" XDE_EXPORT int xde_dopri5_mlp_adjoint_f32(const xde_mlp_field_t *field, const float *t_span, int32_t T,
                                                     const float *y_ans, const float *grad_y, int64_t B,
                                                     const xde_ctrl_opts_t *opts, int32_t controller,
                                                     int32_t adj_norm, float *out_gparams, float *out_adj_y0,
                                                     xde_stats_t *stats, const xde_attempt_log_t *log,
                                                     void *stream) {
  using namespace xde;
  XDE_REQUIRE(field && t_span && y_ans && grad_y && opts && out_gparams, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 2, XDE_E_BAD_ARG, "need B >= 1 and T >= 2");
  XDE_REQUIRE(controller == XDE_CTRL_TRAJECTORY, XDE_E_UNSUPPORTED_FIELD,
              "adjoint: controller=BATCH is not implemented on the device yet");
  XDE_REQUIRE(adj_norm == XDE_ADJ_NORM_SEMI, XDE_E_UNSUPPORTED_FIELD,
              "adjoint with one controller per trajectory supports the seminorm only "
              "(adjoint_options={'norm': 'seminorm'}); the mixed norm needs controller=BATCH");
  cudaStream_t s = (cudaStream_t)stream;
  const int D = field->d, H = field->h;
  const int P = 2 * D * H + H + D;
  AdjParams p{};
  p.field = *field;
  p.t_span = t_span;
  p.y_ans = y_ans;
  p.grad_y = grad_y;
  p.adj_y0 = out_adj_y0;
  p.B = B;
  p.T = T;
  p.o = *opts;
  p.stats = stats;
  p.log_records = log ? log->records : nullptr;
  p.log_counts = log ? log->counts : nullptr;
  p.log_cap = log ? log->cap : 0;
  XDE_CUDA_CHECK(cudaMallocAsync(&p.gacc, sizeof(double) * P, s));
  XDE_CUDA_CHECK(cudaMemsetAsync(p.gacc, 0, sizeof(double) * P, s));
  if (stats) XDE_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(xde_stats_t), s));
  int rc = XDE_E_UNSUPPORTED_FIELD;
  if (D == 2 && H <= 32) rc = adj_pre<2, 1, 4>(p, s);
  else if (D == 2 && H <= 64) rc = adj_pre<2, 2, 4>(p, s);
  else if (D == 1 && H <= 32) rc = adj_pre<1, 1, 8>(p, s);
  else if (D == 1 && H <= 64) rc = adj_pre<1, 2, 8>(p, s);
  else if (D == 4 && H <= 32) rc = adj_pre<4, 1, 2>(p, s);
  else if (D == 4 && H <= 64) rc = adj_pre<4, 2, 2>(p, s);
  else set_last_error("adjoint: field D=%d H=%d has no fused kernel (D in {1,2,4}, H <= 64)", D, H);
  if (rc == XDE_OK) {
    adj_cast_kernel<<<(P + 255) / 256, 256, 0, s>>>(p.gacc, out_gparams, P);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_last_error("adj_cast_kernel launch failed: %s", cudaGetErrorString(e));
      rc = XDE_E_CUDA;
    }
  }
  cudaFreeAsync(p.gacc, s);
  return rc;
}
