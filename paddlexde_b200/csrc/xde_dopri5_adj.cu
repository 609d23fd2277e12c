// xde_dopri5_adj.cu -- OdeintAdjointMethod.backward for the fused MLP field, per-trajectory controller.
//
// Replaces functional/odeint_adjoint.py:47-167: for every output segment [t_i, t_{i-1}] a fresh
// reverse-time dopri5 solve (select_initial_step + adaptive steps + dense output at t_{i-1}) of the
// augmented state (y, a, g_theta), dynamics augmented_dynamics :89-124, then y <- y_ans[i-1],
// a += grad_y[i-1].  Repairs R4-R6 (flat state, reverse time as s = -t, norm over the flat state).
//
// Design (round 2: warp lock-step; round 1b had every lane as a free-running state machine):
//   * ONE THREAD OWNS ONE TRAJECTORY: (y, a), the 7 Dormand-Prince stages of both, t, dt and the
//     accept/reject controller live in registers.  The lanes of a warp take their decisions independently
//     but walk through the slots of a block together (I-block: f0 + probe of select_initial_step; A-block:
//     stages 1..6 of an attempt or the 6 REPLAY evaluations, see below), one field+VJP evaluation per slot,
//     so every controller path runs converged and once per block.
//   * The parameter-gradient state g_theta (P = 2DH+H+D values) is a per-trajectory outer product
//     summed over the batch, far too large for per-thread registers.  It is never integrated as
//     state: since it does not feed back, its value is sum_i W_i * k_i^theta with scalar weights W_i
//     that are known when an attempt starts (dt*c_sol_i, or the dense-output polynomial weights
//     when the attempt reaches the segment end).  Each round a lane stores the (h_j, dz_j) of its
//     evaluation as one column of a per-warp shared-memory tile [H/2][32] of float4
//     (h_j, h_j+1, dz_j, dz_j+1) and publishes (W*u, W, W*a); then the warp TRANSPOSES roles: lane l
//     owns the hidden-unit pair (2l, 2l+1) and folds all 32 columns into lane-private accumulators
//     (5 packed FFMA2 per pair and trajectory).  No shuffles, no atomics in the loop; tile rows are
//     padded to 33 so both phases are bank-conflict free.
//   * The field + VJP is evaluated two hidden units at a time with Blackwell's packed fp32
//     instructions (FFMA2/FMUL2/FADD2, xde_common.cuh), and XDE_ADJ_U such pairs per loop trip: the
//     weight records of the trip are read first and the tile columns stored last, so that no
//     shared-memory store (a possible alias of the next record, as far as ptxas can tell) separates
//     the independent rational-tanh chains.  A packed instruction holds the FMA pipe for 2 cycles
//     while ALU / XU instructions co-issue (tools/probe_issue.cu): the FMA pipe is the ceiling.
//   * The fold is branch-free: lanes without a hidden-unit pair (H/2 < 32) fold a copy of the last
//     tile row into accumulators nobody reads, so the 32 columns form one basic block and the
//     loads of several columns are in flight at once.
//   * A rejected attempt has already been folded in.  The owning lane then spends one extra block
//     of six evaluations (REPLAY) re-evaluating stages 1..5 of the failed attempt with weights -W_i
//     and the start point with (W_0' - W_0) for the shrunk step; the other lanes keep working.
//     The cost of a rejection is one attempt of that lane only.
//   * The FSAL evaluation is folded with W_6 + W_0(next attempt): the controller runs before the
//     fold of that round.  The f0 column of INIT is kept through the probe evaluation and folded
//     once select_initial_step has produced dt.
// Supported norm: the adjoint seminorm (functional/odeint_adjoint.py:301-309) -- with one
// controller per trajectory g_theta is a per-trajectory partial integral.
#include <type_traits>

#include "xde_common.cuh"

#ifndef XDE_ADJ_U
#define XDE_ADJ_U 5  // hidden-unit pairs evaluated together for D <= 2 (independent tanh chains in flight per lane)
#endif
#ifndef XDE_ADJ_IPOL
#define XDE_ADJ_IPOL 3  // an I-block (2 slots) runs when IPOL * (#lanes starting a segment) >= #lanes in an attempt
#endif
#ifndef XDE_ADJ_CTAS
#define XDE_ADJ_CTAS 3  // __launch_bounds__ minimum; ptxas settles at 166 registers, i.e. 4 resident CTAs per SM
#endif
// U x CTAS swept on B200 after the fold became branch-free (cfg2 forward + adjoint, ms): 2x4 15.8, 3x4 15.6, 4x4 15.1,
// 5x4 15.8, 2x3 15.6, 4x3 15.3, 5x3 15.0 (H = 50: 25 pairs = 5 trips, no tail), 6x3 15.4, 8x3 16.3, 5x2 19.3.

namespace xde {

constexpr int kAdjThreads = 96;  // 3 warps; the per-warp tiles for H = 50 allow 5 CTAs per SM, the registers 4 (XDE_ADJ_CTAS)
constexpr int kAdjWarps = kAdjThreads / 32;
constexpr int kTileStride = 33;  // float4 elements per hidden-unit-pair row (32 lanes + 1 pad)

struct AdjParams {
  xde_mlp_field_t field;
  const float *t_span, *y_ans, *grad_y;
  double *gacc;       // [P] fp64 accumulator (zeroed)
  double *gt_acc;     // [T] fp64 accumulator of grad_t_span (zeroed), or null
  unsigned long long *queue;  // next unclaimed trajectory, shared by the whole grid (zeroed)
  float *adj_y0;      // [B,D] or null
  long long B;
  int T;
  xde_ctrl_opts_t o;
  xde_stats_t *stats;
  xde_attempt_t *log_records;
  int *log_counts;
  int log_cap;
};

struct AdjTables {   // shared-memory coefficient tables (dynamic stage lookup)
  float beta[6][8];  // beta[i][j], input of stage i+1
  float wsol[8], wa[8], wb[8], wc[8];
};

enum { AM_IDLE = 0, AM_INIT = 1, AM_ATTEMPT = 2, AM_REPLAY = 3, AM_DONE = 4 };

template <int D>
struct AdjCoef {  // per-lane fold coefficients: {W*u[0..D), W, W*a[0..D)} padded to float4s
  static constexpr int N = 2 * D + 1;
  static constexpr int STRIDE = ((N + 3) / 4) * 4;
};

template <int D, int HPL>
struct AdjSmem {
  static __host__ __device__ size_t tile_floats(int H) { return (size_t)4 * SmallRec<D>::pairs(H) * kTileStride; }
  static __host__ __device__ size_t warp_floats(int H) { return tile_floats(H) + 32 * AdjCoef<D>::STRIDE; }
  static __host__ __device__ size_t bytes(int H, int T) {
    size_t fl = SmallRec<D>::floats(H) + (size_t)((T + 3) / 4) * 4 + kAdjWarps * warp_floats(H);
    size_t red = sizeof(double) * (size_t)(2 * D * H + H + D);
    size_t b = sizeof(float) * fl;
    return b > red ? b : red;
  }
};

template <int D, int HPL, int PRE>
__global__ void __launch_bounds__(kAdjThreads, (D <= 2 && HPL <= 1) ? XDE_ADJ_CTAS : 1) dopri5_adj_kernel(const AdjParams p) {
  constexpr int C = 2 * D;                 // state components per trajectory: y then a
  constexpr int NTP = (2 * D + 1) * HPL;   // lane-private theta accumulator PAIRS (HPL unit pairs per lane)
  constexpr int NTH = 2 * NTP;
  constexpr int REC = SmallRec<D>::REC;
  constexpr int CST = AdjCoef<D>::STRIDE;

  extern __shared__ __align__(16) float smem[];
  __shared__ AdjTables tb;
  __shared__ unsigned long long s_cnt[3];
  __shared__ int s_status;

  const int H = p.field.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *sw = smem;
  float *st = sw + SmallRec<D>::floats(H);
  float *wbase = st + ((p.T + 3) / 4) * 4 + (size_t)warp * AdjSmem<D, HPL>::warp_floats(H);
  float4 *tile = reinterpret_cast<float4 *>(wbase);                    // [H/2][33] of (h_j, h_j+1, dz_j, dz_j+1)
  const int NP = SmallRec<D>::pairs(H);
  float *coef = wbase + AdjSmem<D, HPL>::tile_floats(H);               // [32][CST]
  const float4 *trow[HPL];  // fold phase: the tile row(s) of this lane's hidden-unit pair(s), clamped to the last row
#pragma unroll
  for (int q = 0; q < HPL; ++q) trow[q] = tile + min(lane + 32 * q, NP - 1) * kTileStride;

  load_small_field<D>(sw, p.field);
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
  if (threadIdx.x == 0) {
    s_cnt[0] = s_cnt[1] = s_cnt[2] = 0ull;
    s_status = 0;
  }
  for (int i = lane; i < 32 * CST; i += 32) coef[i] = 0.0f;
  for (int i = lane; i < NP * kTileStride; i += 32) tile[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  const xde_ctrl_opts_t o = p.o;

  // ---- hidden-unit role (fold phase): this lane's units j = lane + 32 q ----
  double acc[NTH];  // committed parameter-gradient sums of this lane's units (fp64); [2*i + e]: pair i, unit e
  f32x2 Tt[NTP];    // packed fp32 running sums, flushed into acc every few rounds
#pragma unroll
  for (int i = 0; i < NTH; ++i) acc[i] = 0.0;
#pragma unroll
  for (int i = 0; i < NTP; ++i) Tt[i] = pk1(0.0f);
  double gb2acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) gb2acc[d] = 0.0;

  // ---- trajectory role: one controller per thread ----
  int mode = AM_IDLE, seg = 0, n_steps = 0, n_logged = 0;
  long long traj = -1;
  float s0[C], kk[7][C];
#pragma unroll
  for (int e = 0; e < C; ++e) {
    s0[e] = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) kk[j][e] = 0.f;
  }
  float t0 = 0.f, dt = 0.f, te = 0.f, xfin = 0.f;
  bool fin = false;  // the attempt in flight reaches the segment end if accepted
  float dt_old = 0.f, xfin_old = 0.f;
  bool fin_old = false;
  float h0 = 0.f, d1 = 0.f;
  unsigned n_att = 0, n_acc = 0, n_fe = 0;
  int status = 0;
  int since_flush = 0;
  bool col_ok = true;  // this lane's tile column holds finite values only (zero-weight columns may be folded)

  // theta weight of the evaluation at stage `stg` (0..6) of an attempt (dtv, finv, xv); sign folded in
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
  // rms over the y part and over the a part, fp64 accumulation; adjoint seminorm =
  // max(|g_t| = 0, rms(y), rms(a)) with Python max semantics (functional/odeint_adjoint.py:304-307)
  auto semi_norm = [&](const float (&v)[C], float vt) -> float {
    double sy = 0.0, sa = 0.0;
#pragma unroll
    for (int e = 0; e < D; ++e) {
      sy += (double)(v[e] * v[e]);
      sa += (double)(v[D + e] * v[D + e]);
    }
    // max(rms(y), rms(a)) = rms of the larger sum: x -> (float)sqrt(x / D) is monotone, so it commutes with
    // the max and ONE fp64 square root serves both parts (a NaN sum fails both comparisons exactly as a NaN
    // rms failed `>` before).  Mean over D: for a power of two the reciprocal multiply is the exact quotient.
    constexpr bool kPow2 = (D & (D - 1)) == 0;
    double m = 0.0;
    if (sy > m) m = sy;
    if (sa > m) m = sa;
    const float r = kPow2 ? (float)sqrt(m * (1.0 / D)) : rms_from_sumsq(m, (double)D);
    float best = fabsf(vt);  // the g_t slot (0 unless grad_t_span is asked for)
    if (r > best) best = r;
    return best;
  };
  auto plan_attempt = [&]() {  // (t0, dt, te) -> does the attempt reach the segment end, and where
    const float t1n = t0 + dt;
    fin = !(te > t1n);
    xfin = fin ? __fdiv_rn(te - t0, t1n - t0) : 0.f;
  };

  // grad_t_span (functional/odeint_adjoint.py:129-141,161-162): slot 0 of the flat augmented state.  The field
  // ignores t, so its derivative is exactly zero (allow_unused, :116-118): the slot only changes at the output
  // times (aug_state[0] -= dLd_cur_t) and through the rounding of the dense-output evaluation, and it enters the
  // controller through d0 of select_initial_step alone (|g_t / scale|; its f0, f1 and error estimate are 0).
  const bool want_gt = (p.gt_acc != nullptr);
  float gt = 0.f;
  // fp64 sum over the lanes that contribute to the same address (one atomic per warp when they all do)
  auto warp_add = [&](bool mine, double v, int slot_idx) {
    const unsigned m = __ballot_sync(XDE_FULL_MASK, mine);
    if (!m) return;
    const int leader = __ffs(m) - 1;
    const int idx0 = __shfl_sync(XDE_FULL_MASK, slot_idx, leader);
    if (__all_sync(XDE_FULL_MASK, !mine || slot_idx == idx0)) {
      double s = mine ? v : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(XDE_FULL_MASK, s, off);
      if (lane == 0) atomicAdd(&p.gt_acc[idx0], s);
    } else if (mine) {
      atomicAdd(&p.gt_acc[slot_idx], v);
    }
  };

  // ================= warp lock-step schedule =================
  // Every lane is a trajectory with its own controller, but the lanes of a warp walk through the SAME slot of a
  // block together:  I-block = 2 slots (f0, Euler probe of select_initial_step),  A-block = 6 slots (stages 1..6 of
  // an attempt, or the 6 REPLAY evaluations that take a rejected attempt back out of the parameter-gradient
  // fold).  All controller code therefore runs converged, once per block instead of once per evaluation (r1: every
  // lane was its own state machine and each round paid for every controller path that some lane was in; ncu: 12.5 %
  // of the samples in select_initial_step at 8/32 lanes, 8 % in the end-of-attempt controller).  An I-block only
  // helps the lanes that start a segment; the others idle through its 2 slots, while a lane whose I-block is put
  // off idles through a 6-slot A-block: it runs when  2 * (#attempt lanes) <= 6 * (#init lanes).  On a batch whose
  // trajectories take the same number of attempts per segment (cfg2: 96 % take exactly two) the lanes stay in
  // phase and nothing idles.
  while (true) {
    // ---- block boundary: refill idle lanes from the grid-wide queue (warp-aggregated atomic) ----
    // One queue for the whole grid: with one contiguous chunk per CTA the CTAs finished up to 8 % apart (r1s).
    {
      const bool need = (mode == AM_IDLE);
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
            seg = p.T - 1;
            // aug_state = [0, y_ans[-1], grad_y[-1]] (functional/odeint_adjoint.py:75-82)
#pragma unroll
            for (int e = 0; e < D; ++e) {
              const long long src = ((long long)seg * p.B + traj) * D + e;
              s0[e] = p.y_ans[src];
              s0[D + e] = p.grad_y[src];
            }
            gt = 0.f;
            t0 = st[seg];
            te = st[seg - 1];
            mode = AM_INIT;
            n_steps = 0;
            n_logged = 0;
          } else {
            mode = AM_DONE;
          }
        }
      }
    }
    if (__all_sync(XDE_FULL_MASK, mode == AM_DONE)) break;

    const unsigned mI = __ballot_sync(XDE_FULL_MASK, mode == AM_INIT);
    const unsigned mA = __ballot_sync(XDE_FULL_MASK, mode == AM_ATTEMPT || mode == AM_REPLAY);
    const bool run_init = (mI != 0u) && (XDE_ADJ_IPOL * __popc(mI) >= __popc(mA));

#pragma unroll 1
    for (int slot = run_init ? 0 : 2; slot < 8; ++slot) {
      // ================= (1) assertions when an attempt starts =================
      // _adaptive_step (base_adaptive_solver_rk.py:200-203) + max_num_steps (:120-122)
      if (slot == 2) {
        if (mode == AM_ATTEMPT) {
          int bad = 0;
          if (!(n_steps < o.max_num_steps)) {
            bad = XDE_ST_MAX_STEPS;
          } else if (!(t0 + dt > t0)) {
            bad = XDE_ST_DT_UNDERFLOW;
          } else {
            bool finite = true;
#pragma unroll
            for (int e = 0; e < C; ++e) finite = finite && (fabsf(s0[e]) < INFINITY);
            if (!finite) bad = XDE_ST_NONFINITE_STATE;
          }
          if (bad) {
            status = max(status, bad);
            if (p.adj_y0) {
#pragma unroll
              for (int d = 0; d < D; ++d) p.adj_y0[traj * D + d] = NAN;
            }
            if (p.log_counts) p.log_counts[traj] = n_logged;
            mode = AM_IDLE;
          }
        }
        if (!__any_sync(XDE_FULL_MASK, mode == AM_ATTEMPT || mode == AM_REPLAY)) break;
      }
      const bool ini = (slot < 2) && (mode == AM_INIT);
      const bool att = (slot >= 2) && (mode == AM_ATTEMPT);
      const bool rep = (slot >= 2) && (mode == AM_REPLAY);
      const bool live = ini || att || rep;
      // number of k terms of this slot's evaluation point: INIT 0 / 1, ATTEMPT stage 1..6, REPLAY r = 0..5
      const int stage = ini ? slot : (att ? slot - 1 : (rep ? slot - 2 : 0));
      const float t1 = t0 + dt;

      // ================= (2) input of this slot's evaluation =================
      // ATTEMPT stage i: s0 + sum_{j<i} k_j*(beta_ij*dt) (base_adaptive_solver_rk.py:166-168);
      // REPLAY r: the same for stage r of the rejected attempt (dt_old), r = 0 -> the start point;
      // INIT 0: s0; INIT 1: the Euler probe k0*h0 + s0 (base_adaptive_solver.py:60).
      float yin[C];
      {
        const int nst = stage;
        const int row = (nst >= 1) ? nst - 1 : 0;
        const float dtv = rep ? dt_old : dt;
        float cj[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) cj[j] = tb.beta[row][j] * dtv;
        if (ini) cj[0] = h0;
#pragma unroll
        for (int e = 0; e < C; ++e) {
          float s = kk[0][e] * cj[0];
#pragma unroll
          for (int j = 1; j < 6; ++j)
            if (j < nst) s = s + kk[j][e] * cj[j];
          yin[e] = (nst >= 1) ? (s0[e] + s) : s0[e];
        }
      }

      // ================= (3) field + VJP evaluation of this lane's trajectory =================
      // f = tanh(pre(y) W1 + b1) W2 + b2 ; dh = a W2^T ; dz = dh (1 - h^2) ; du = dz W1^T (Appendix B)
      const bool wr = live && (slot != 1);  // the probe must not overwrite the f0 column
      float fo[C];
      {
        float u[D];
        f32x2 accf[D], pdu[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
          u[k] = pre_act<PRE>(yin[k]);
          accf[k] = pk1(0.0f);
          pdu[k] = pk1(0.0f);
        }
        // U hidden-unit pairs per iteration: all weight records are read before and all tile columns are
        // written after the arithmetic, so no shared-memory store sits between two pairs' chains (ptxas
        // cannot prove that the tile and the weight records do not alias and would serialise the pairs);
        // the U rational-tanh chains are independent and interleave.  The second-layer / VJP chains are
        // advanced in pair order: per value the arithmetic is unchanged.
        auto eval_pairs = [&](int jp0, auto ucount) {
          constexpr int U = decltype(ucount)::value;
          f32x2 w1p[U][D], b1p[U], w2p[U][D], h[U], dz[U];
#pragma unroll
          for (int i = 0; i < U; ++i) read_pair_rec<D>(sw, jp0 + i, w1p[i], b1p[i], w2p[i]);
#pragma unroll
          for (int i = 0; i < U; ++i) {
            f32x2 z = first_layer_seed<D>(u[0], w1p[i][0]);
#pragma unroll
            for (int k = 1; k < D; ++k) z = fma2(pk1(u[k]), w1p[i][k], z);
            h[i] = tanh_rat2(add2(z, b1p[i]));
          }
#pragma unroll
          for (int i = 0; i < U; ++i) {
            f32x2 dh = mul2(pk1(yin[D]), w2p[i][0]);
#pragma unroll
            for (int d = 1; d < D; ++d) dh = fma2(pk1(yin[D + d]), w2p[i][d], dh);
            dz[i] = mul2(dh, one_minus_sq2(h[i]));
          }
#pragma unroll
          for (int i = 0; i < U; ++i) {
#pragma unroll
            for (int d = 0; d < D; ++d) accf[d] = fma2(h[i], w2p[i][d], accf[d]);
#pragma unroll
            for (int k = 0; k < D; ++k) pdu[k] = fma2(dz[i], w1p[i][k], pdu[k]);
          }
          if (wr) {
#pragma unroll
            for (int i = 0; i < U; ++i) {
              float h0, h1, z0, z1;
              upk(h[i], h0, h1);
              upk(dz[i], z0, z1);
              tile[(jp0 + i) * kTileStride + lane] = make_float4(h0, h1, z0, z1);
            }
          }
        };
        constexpr int UT = (D <= 2) ? XDE_ADJ_U : 2;  // wider states: two pairs per trip (register budget)
        int jp = 0;
#pragma unroll 1
        for (; jp + UT <= NP; jp += UT) eval_pairs(jp, std::integral_constant<int, UT>());
#pragma unroll 1
        for (; jp < NP; ++jp) eval_pairs(jp, std::integral_constant<int, 1>());
        // solver-time dynamics: dy/ds = tsign * f ; da/ds = -tsign * vjp_y(a)
#pragma unroll
        for (int d = 0; d < D; ++d) {
          float fe, fod, ue, uo;
          upk(accf[d], fe, fod);
          upk(pdu[d], ue, uo);
          fo[d] = tsign * ((fe + fod) + sw[NP * REC + d]);
          fo[D + d] = (-tsign) * ((ue + uo) * pre_act_grad<PRE>(yin[d]));
        }
      }

      if (wr) {
        bool finite = true;
#pragma unroll
        for (int e2 = 0; e2 < C; ++e2) finite = finite && (fabsf(fo[e2]) < INFINITY);
        col_ok = finite;  // fo finite => every (h, dz) of the column is finite
      }

      // ================= (4) controller: one path per slot, entered by all its lanes together =================
      float wacc = 0.f;           // fold weight of this lane's tile column
      bool col_is_start = false;  // the column holds the f0 evaluation at the start point
      float dl_t = 0.f;           // grad_t_span[seg] contribution of this lane (slot 0)
      float gt_done = 0.f;        // final aug_state[0] of a trajectory that finishes in this slot
      bool traj_done = false;
      if (ini) {
        if (slot == 0) {
          // _before_integrate f0 + select_initial_step part 1 (base_adaptive_solver.py:44-57)
          float vt = 0.f;
          if (want_gt) {
            // dLd_cur_t = func(t_i, y_i) . grad_y[i];  aug_state[0] -= dLd_cur_t  (functional/odeint_adjoint.py:135-141);
            // func(t_i, y_i) is the y part of this very evaluation (tsign * fo is exact)
#pragma unroll
            for (int d = 0; d < D; ++d) {
              const float pr = (tsign * fo[d]) * p.grad_y[((long long)seg * p.B + traj) * D + d];
              dl_t = (d == 0) ? pr : dl_t + pr;
            }
            gt = gt - dl_t;
            vt = __fdiv_rn(gt, o.atol + fabsf(gt) * o.rtol);
          }
          float v0[C], v1[C];
#pragma unroll
          for (int e = 0; e < C; ++e) {
            kk[0][e] = fo[e];
            const float sc = o.atol + fabsf(s0[e]) * o.rtol;
            v0[e] = __fdiv_rn(s0[e], sc);
            v1[e] = __fdiv_rn(fo[e], sc);
          }
          const float d0 = fabsf(semi_norm(v0, vt));
          d1 = fabsf(semi_norm(v1, 0.f));
          if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f; else h0 = __fdiv_rn(0.01f * d0, d1);
          h0 = fabsf(h0);
        } else {
          float v[C];
#pragma unroll
          for (int e = 0; e < C; ++e) {
            const float sc = o.atol + fabsf(s0[e]) * o.rtol;
            v[e] = __fdiv_rn(fo[e] - kk[0][e], sc);
          }
          const float d2 = fabsf(__fdiv_rn(semi_norm(v, 0.f), h0));
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
          n_fe += has_first ? 1u : 3u;
          mode = AM_ATTEMPT;
          plan_attempt();
          wacc = theta_w(0, dt, fin, xfin);
          col_is_start = true;
        }
      } else if (att) {
        const float w_eval = theta_w(stage, dt, fin, xfin);
        if (slot < 7) {
#pragma unroll
          for (int j = 1; j < 6; ++j)
            if (j == stage) {
#pragma unroll
              for (int e = 0; e < C; ++e) kk[j][e] = fo[e];
            }
          wacc = w_eval;
        } else {
          // error estimate and ratio (base_adaptive_solver_rk.py:180; ode_utils.py:80-82), k6 = fo
          float v[C];
#pragma unroll
          for (int e = 0; e < C; ++e) {
            kk[6][e] = fo[e];
            float er = kk[0][e] * (dt * DP::cerr(0));
#pragma unroll
            for (int j = 1; j < 7; ++j) er = er + kk[j][e] * (dt * DP::cerr(j));
            const float tol = o.atol + o.rtol * fmaxf(fabsf(s0[e]), fabsf(yin[e]));
            v[e] = __fdiv_rn(er, tol);
          }
          const float ratio = fabsf(semi_norm(v, 0.f));
          bool accept = (ratio <= 1.0f);
          if (dt > o.max_step) accept = false;
          if (dt <= o.min_step) accept = true;
          const float dt_next = next_step_size(dt, ratio, o);
          n_att++;
          n_fe += 6;
          if (p.log_records && n_logged < p.log_cap) {
            xde_attempt_t r;
            r.t0 = tsign * t0;
            r.dt = tsign * dt;
            r.ratio = ratio;
            r.accepted = accept ? 1 : 0;
            p.log_records[traj * p.log_cap + n_logged] = r;
          }
          n_logged++;
          n_steps++;
          if (accept) {
            n_acc++;
            if (fin) {
              // dense output at the segment end (interp_fit + interp_evaluate), then
              // y <- y_ans[i-1], a += grad_y[i-1] (functional/odeint_adjoint.py:153-159)
              seg -= 1;
              const float two_dt = 2.0f * dt;
              const float x = xfin;
              auto dense = [&](float F0, float F1, float Y0, float Y1, float ym) -> float {
                const float ca = (two_dt * (F1 - F0) - 8.0f * (Y1 + Y0)) + 16.0f * ym;
                const float cb = ((dt * (5.0f * F0 - 3.0f * F1) + 18.0f * Y0) + 14.0f * Y1) - 32.0f * ym;
                const float cc = ((dt * (F1 - 4.0f * F0) - 11.0f * Y0) - 5.0f * Y1) + 16.0f * ym;
                const float cd = dt * F0;
                float total = Y0 + x * cd;
                float xp = x * x;
                total = total + xp * cc;
                xp = xp * x;
                total = total + xp * cb;
                xp = xp * x;
                total = total + xp * ca;
                return total;
              };
#pragma unroll
              for (int e = 0; e < C; ++e) {
                float sm = kk[0][e] * (dt * DP::cmid(0));
#pragma unroll
                for (int j = 1; j < 7; ++j) sm = sm + kk[j][e] * (dt * DP::cmid(j));
                const float ym = s0[e] + sm;
                const float total = dense(kk[0][e], kk[6][e], s0[e], yin[e], ym);
                const long long src = ((long long)seg * p.B + traj) * D + (e < D ? e : e - D);
                s0[e] = (e < D) ? p.y_ans[src] : (total + p.grad_y[src]);
              }
              if (want_gt) {
                // the g_t slot: every k_j is 0, so its stage inputs and its midpoint are g_t + 0; the dense-output
                // polynomial is evaluated on it like on any other component (its rounding is part of the result)
                const float zs = 0.0f * (dt * DP::cmid(0));
                const float gy1 = gt + zs;
                gt = dense(0.0f, 0.0f, gt, gy1, gy1);
              }
              n_steps = 0;
              wacc = w_eval;
              if (seg == 0) {
                if (p.adj_y0) {
#pragma unroll
                  for (int d = 0; d < D; ++d) p.adj_y0[traj * D + d] = s0[D + d];
                }
                if (p.log_counts) p.log_counts[traj] = n_logged;
                traj_done = true;
                gt_done = gt;
                mode = AM_IDLE;
              } else {
                t0 = st[seg];
                te = st[seg - 1];
                mode = AM_INIT;
              }
            } else {
#pragma unroll
              for (int e = 0; e < C; ++e) {
                s0[e] = yin[e];
                kk[0][e] = kk[6][e];
              }
              t0 = t1;
              dt = dt_next;
              plan_attempt();
              // FSAL: this evaluation is also stage 0 of the next attempt
              wacc = w_eval + theta_w(0, dt, fin, xfin);
            }
          } else {
            // rejected: everything folded for this attempt is taken back by a REPLAY block
            dt_old = dt;
            fin_old = fin;
            xfin_old = xfin;
            dt = dt_next;
            plan_attempt();
            mode = AM_REPLAY;
            wacc = 0.f;
          }
        }
      } else if (rep) {
        if (stage == 0)
          wacc = theta_w(0, dt, fin, xfin) - theta_w(0, dt_old, fin_old, xfin_old);
        else
          wacc = -theta_w(stage, dt_old, fin_old, xfin_old);
        if (slot == 7) mode = AM_ATTEMPT;
      }
      if (want_gt) {  // converged: batch sums of grad_t_span in fp64
        if (slot == 0) warp_add(ini, (double)dl_t, seg);
        if (slot == 7) warp_add(traj_done, (double)gt_done, 0);
      }
      if (slot == 0) continue;  // nothing to fold: the f0 column waits for the weight select_initial_step produces

      // ================= (5) fold: transpose roles, lane l <- hidden units l, l+32 =================
      {
        // the column's inputs: this slot's evaluation point, or the start point for the kept f0 column
        float cu[D], ca[D];
        const bool fold = (wacc != 0.0f);
#pragma unroll
        for (int d = 0; d < D; ++d) {
          // col_is_start only happens in INIT, where s0 is unchanged since the f0 evaluation; after a
          // segment end s0 already holds the next segment's start, so the stage-6 point is read from yin
          const float yv = col_is_start ? s0[d] : yin[d];
          const float av = col_is_start ? s0[D + d] : yin[D + d];
          cu[d] = fold ? wacc * pre_act<PRE>(yv) : 0.0f;
          ca[d] = fold ? wacc * av : 0.0f;
          gb2acc[d] += (double)ca[d];
        }
        const unsigned fm = __ballot_sync(XDE_FULL_MASK, fold);
        if (fm) {
          float *c = coef + lane * CST;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            c[d] = cu[d];
            c[D + 1 + d] = ca[d];
          }
          c[D] = fold ? wacc : 0.0f;
          __syncwarp();
          auto fold_column = [&](int b) {
            float cb[CST];
            const float4 *c4 = reinterpret_cast<const float4 *>(coef + b * CST);
#pragma unroll
            for (int q = 0; q < CST / 4; ++q) {
              const float4 v = c4[q];
              cb[4 * q] = v.x;
              cb[4 * q + 1] = v.y;
              cb[4 * q + 2] = v.z;
              cb[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int q = 0; q < HPL; ++q) {
              // no `pair < NP` test: a lane without a hidden-unit pair folds a copy of the last row into
              // accumulators the epilogue never reads, so the 32 columns are ONE basic block and ptxas can
              // keep the loads of several columns in flight (with the branch every column exposed a full
              // shared-memory round trip)
              const float4 hv = trow[q][b];
              const f32x2 hp = pk(hv.x, hv.y), dzp = pk(hv.z, hv.w);
#pragma unroll
              for (int k = 0; k < D; ++k) Tt[q * (2 * D + 1) + k] = fma2(pk1(cb[k]), dzp, Tt[q * (2 * D + 1) + k]);
              Tt[q * (2 * D + 1) + D] = fma2(pk1(cb[D]), dzp, Tt[q * (2 * D + 1) + D]);
#pragma unroll
              for (int d = 0; d < D; ++d)
                Tt[q * (2 * D + 1) + D + 1 + d] = fma2(pk1(cb[D + 1 + d]), hp, Tt[q * (2 * D + 1) + D + 1 + d]);
            }
          };
          // Fast path: every column is finite, so columns with zero weight contribute exactly +0 and all 32
          // can be folded by straight-line code (compile-time shared-memory offsets, no mask arithmetic).
          if (__all_sync(XDE_FULL_MASK, col_ok)) {
#pragma unroll
            for (int b = 0; b < 32; ++b) fold_column(b);
          } else {
            unsigned m = fm;
            while (m) {
              const int b = __ffs(m) - 1;
              m &= m - 1;
              fold_column(b);
            }
          }
          __syncwarp();
          if (++since_flush >= 8) {
            since_flush = 0;
#pragma unroll
            for (int i = 0; i < NTP; ++i) {
              float e0, e1;
              upk(Tt[i], e0, e1);
              acc[2 * i] += (double)e0;
              acc[2 * i + 1] += (double)e1;
              Tt[i] = pk1(0.0f);
            }
          }
        }
      }
    }
  }

  // ================= epilogue: parameter gradients and stats =================
#pragma unroll
  for (int i = 0; i < NTP; ++i) {
    float e0, e1;
    upk(Tt[i], e0, e1);
    acc[2 * i] += (double)e0;
    acc[2 * i + 1] += (double)e1;
  }
  __syncthreads();  // every warp is done with its tile: reuse shared memory as the CTA reduction buffer
  double *red = reinterpret_cast<double *>(smem);
  const int P = 2 * D * H + H + D;
  for (int i = threadIdx.x; i < P; i += blockDim.x) red[i] = 0.0;
  __syncthreads();
#pragma unroll
  for (int q = 0; q < HPL; ++q) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = 2 * (lane + 32 * q) + e;  // hidden unit of accumulator half e of pair q
      if (j < H) {
        const int base = q * (2 * D + 1);
#pragma unroll
        for (int k = 0; k < D; ++k) atomicAdd(&red[k * H + j], acc[2 * (base + k) + e]);
        atomicAdd(&red[D * H + j], acc[2 * (base + D) + e]);
#pragma unroll
        for (int d = 0; d < D; ++d) atomicAdd(&red[D * H + H + j * D + d], acc[2 * (base + D + 1 + d) + e]);
      }
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    double v = gb2acc[d];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(XDE_FULL_MASK, v, off);
    if (lane == 0) atomicAdd(&red[D * H + H + H * D + d], v);
  }
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
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) atomicAdd(&p.gacc[i], red[i]);
  if (threadIdx.x == 0 && p.stats) {
    atomicAdd(&p.stats->n_attempts, s_cnt[0]);
    atomicAdd(&p.stats->n_accepted, s_cnt[1]);
    atomicAdd(&p.stats->nfe, s_cnt[2]);
    atomicMax(&p.stats->status, s_status);
  }
}

// gparams [n] and, when asked for, grad_t_span [nt] (fp64 batch sums -> fp32)
__global__ void adj_cast_kernel(const double *__restrict__ a, float *__restrict__ o, int n,
                                const double *__restrict__ at, float *__restrict__ ot, int nt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = (float)a[i];
  if (ot && i < nt) ot[i] = (float)at[i];
}

template <int D, int HPL, int PRE>
static int launch_adj(const AdjParams &p, cudaStream_t stream) {
  const size_t smem = AdjSmem<D, HPL>::bytes(p.field.h, p.T);
  XDE_REQUIRE(smem <= 200 * 1024, XDE_E_UNSUPPORTED_FIELD, "adjoint: field + t_span exceed shared memory");
  auto kern = dopri5_adj_kernel<D, HPL, PRE>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  XDE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kAdjThreads, smem));
  if (per_sm < 1) per_sm = 1;
  // persistent grid: a whole number of CTAs per SM, all fed from one trajectory queue
  long long want = (p.B + kAdjThreads - 1) / kAdjThreads;
  long long grid = (long long)sm_count() * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kAdjThreads, smem, stream>>>(p);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  return XDE_OK;
}

template <int D, int HPL>
static int adj_pre(const AdjParams &p, cudaStream_t s) {
  switch (p.field.pre) {
    case XDE_PRE_ID: return launch_adj<D, HPL, XDE_PRE_ID>(p, s);
    case XDE_PRE_SQUARE: return launch_adj<D, HPL, XDE_PRE_SQUARE>(p, s);
    case XDE_PRE_CUBE: return launch_adj<D, HPL, XDE_PRE_CUBE>(p, s);
  }
  set_last_error("unknown pre-activation %d", p.field.pre);
  return XDE_E_BAD_ARG;
}

template <int D>
static int adj_hpl(const AdjParams &p, cudaStream_t s) {
  const int H = p.field.h;
  if (H <= 64) return adj_pre<D, 1>(p, s);  // HPL = hidden-unit PAIRS per lane in the fold phase
  if constexpr (D <= 2) {
    if (H <= 128) return adj_pre<D, 2>(p, s);
  }
  set_last_error("adjoint: hidden width H=%d has no fused kernel for D=%d (H <= 64; H <= 128 for D <= 2)", H, D);
  return XDE_E_UNSUPPORTED_FIELD;
}

int dopri5_adj_tile(const xde_mlp_field_t *field, const float *t_span, int T, const float *y_ans, const float *grad_y,
                    long long B, const xde_ctrl_opts_t *opts, double *gacc, unsigned long long *queue, float *out_adj_y0,
                    xde_stats_t *stats, const xde_attempt_log_t *log, cudaStream_t s);  // xde_adj_tile.cu

int dopri5_adj_batch(const xde_mlp_field_t *field, const float *t_span, int T, const float *y_ans,
                     const float *grad_y, long long B, const xde_ctrl_opts_t *opts, int adj_norm,
                     float *out_gparams, float *out_adj_y0, xde_stats_t *stats, const xde_attempt_log_t *log,
                     cudaStream_t s);  // xde_dopri5_adj_batch.cu

}  // namespace xde

extern "C" XDE_EXPORT int xde_dopri5_mlp_adjoint_f32(const xde_mlp_field_t *field, const float *t_span, int32_t T,
                                                     const float *y_ans, const float *grad_y, int64_t B,
                                                     const xde_ctrl_opts_t *opts, int32_t controller,
                                                     int32_t adj_norm, float *out_gparams, float *out_adj_y0,
                                                     float *out_grad_t, xde_stats_t *stats,
                                                     const xde_attempt_log_t *log, void *stream) {
  using namespace xde;
  XDE_REQUIRE(field && t_span && y_ans && grad_y && opts && out_gparams, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 2, XDE_E_BAD_ARG, "need B >= 1 and T >= 2");
  XDE_REQUIRE(controller == XDE_CTRL_TRAJECTORY || controller == XDE_CTRL_BATCH, XDE_E_BAD_ARG,
              "unknown controller %d", controller);
  XDE_REQUIRE(adj_norm == XDE_ADJ_NORM_SEMI || adj_norm == XDE_ADJ_NORM_MIXED, XDE_E_BAD_ARG,
              "unknown adjoint norm %d", adj_norm);
  cudaStream_t s = (cudaStream_t)stream;
  if (controller == XDE_CTRL_BATCH) {
    XDE_REQUIRE(out_grad_t == nullptr, XDE_E_UNSUPPORTED_FIELD,
                "grad_t_span is computed by the per-trajectory controller only (controller=TRAJECTORY)");
    if (stats) XDE_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(xde_stats_t), s));
    return dopri5_adj_batch(field, t_span, T, y_ans, grad_y, B, opts, adj_norm, out_gparams, out_adj_y0, stats, log, s);
  }
  XDE_REQUIRE(adj_norm == XDE_ADJ_NORM_SEMI, XDE_E_UNSUPPORTED_FIELD,
              "adjoint with one controller per trajectory supports the seminorm only "
              "(adjoint_options={'norm': 'seminorm'}); the mixed norm couples all trajectories and needs "
              "controller=BATCH");
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
  // [P] parameter gradients | the trajectory queue | [T] grad_t_span
  const size_t n_scratch = (size_t)P + 1 + (out_grad_t ? (size_t)T : 0);
  XDE_CUDA_CHECK(scratch_alloc((void **)&p.gacc, sizeof(double) * n_scratch, s));
  XDE_CUDA_CHECK(cudaMemsetAsync(p.gacc, 0, sizeof(double) * n_scratch, s));
  p.queue = reinterpret_cast<unsigned long long *>(p.gacc + P);
  p.gt_acc = out_grad_t ? p.gacc + P + 1 : nullptr;
  if (stats) XDE_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(xde_stats_t), s));
  int rc = XDE_E_UNSUPPORTED_FIELD;
  switch (D) {
    case 1: rc = adj_hpl<1>(p, s); break;
    case 2: rc = adj_hpl<2>(p, s); break;
    case 3: rc = adj_hpl<3>(p, s); break;
    case 4: rc = adj_hpl<4>(p, s); break;
    case 5: rc = adj_hpl<5>(p, s); break;
    case 6: rc = adj_hpl<6>(p, s); break;
    case 7: rc = adj_hpl<7>(p, s); break;
    case 8: rc = adj_hpl<8>(p, s); break;
    default:  // large states: tiles of trajectories per CTA (xde_adj_tile.cu)
      if (out_grad_t) {
        set_last_error("adjoint: grad_t_span is computed by the small-state kernels only (D <= 8)");
        rc = XDE_E_UNSUPPORTED_FIELD;
      } else {
        rc = dopri5_adj_tile(field, t_span, T, y_ans, grad_y, B, opts, p.gacc, p.queue, out_adj_y0, stats, log, s);
      }
  }
  if (rc == XDE_OK) {
    const int ncast = P > T ? P : T;
    adj_cast_kernel<<<(ncast + 255) / 256, 256, 0, s>>>(p.gacc, out_gparams, P, p.gt_acc, out_grad_t, T);
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
