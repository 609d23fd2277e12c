// xde_adaptive_rk.cu -- odeint(func=MLP, solver=Bosh3 | Fehlberg2 | AdaptiveHeun | Dopri8) forward: the
// reference's other embedded Runge-Kutta pairs (solver/__init__.py:1-6) on ONE table-driven kernel.
//
// Replaces, like xde_dopri5_fwd.cu, the Python hot loop of solver/base_adaptive_solver.py:24-31 +
// solver/base_adaptive_solver_rk.py:116-292 (+ utils/ode_utils.py:28-97); here the tableau is data
// (stage count, FSAL shortcut :172-176, order of the step-size controller), exactly as in the
// reference.  Tableaux: adaptive_solver/bosh3.py:5-27, fehlberg2.py:5-22, adaptive_heun.py:5-27,
// dopri8.py:5-252 (authored in float64, rounded once to fp32: base_adaptive_solver_rk.py:73-79).
//
// One thread owns one trajectory (per-trajectory controller, north star): state, controller and the
// dense-output coefficients in registers, the S+1 stage derivatives in shared memory laid out
// [stage][component][thread] (conflict-free), weights in shared memory as in the Dopri5 kernel.  This is
// the breadth kernel: it shares every arithmetic rule with the tuned Dopri5 kernel
// -- run with the Dormand-Prince tableau (XDE_RK_DOPRI5_TABLE) it reproduces that kernel bit for bit --
// but has no lane refill, so diverging step counts cost warp efficiency.
#include "xde_rk_tab.cuh"

namespace xde {

constexpr int kRkThreads = 128;
struct RkParams {
  xde_mlp_field_t field;
  const float *y0, *t_span;
  float *out;
  long long B;
  int T;
  xde_ctrl_opts_t o;
  xde_stats_t *stats;
  xde_attempt_t *log_records;
  int *log_counts;
  int log_cap;
  // step_t / jump_t (base_adaptive_solver_rk.py:94-114): sorted in integration order, already filtered to
  // lie at or after t_span[0] (sort_tvals is the shim's), in t_span's own time; may be null / 0
  const float *step_t, *jump_t;
  int n_step, n_jump;
  RkTab tab;
};

// ---- tableaux (host, float64) -----------------------------------------------------------------------
namespace tabs {
const double bs_alpha[3] = {1.0 / 2, 3.0 / 4, 1.0};
const double bs_beta[3][3] = {{1.0 / 2}, {0.0, 3.0 / 4}, {2.0 / 9, 1.0 / 3, 4.0 / 9}};
const double bs_csol[4] = {2.0 / 9, 1.0 / 3, 4.0 / 9, 0.0};
const double bs_cerr[4] = {2.0 / 9 - 7.0 / 24, 1.0 / 3 - 1.0 / 4, 4.0 / 9 - 1.0 / 3, -1.0 / 8};
const double bs_cmid[4] = {0.0, 0.5, 0.0, 0.0};

const double fe_alpha[2] = {1.0 / 2, 1.0};
const double fe_beta[2][2] = {{1.0 / 2}, {1.0 / 256, 255.0 / 256}};
const double fe_csol[3] = {1.0 / 512, 255.0 / 256, 1.0 / 512};
const double fe_cerr[3] = {-1.0 / 512, 0, 1.0 / 512};
const double fe_cmid[3] = {0.0, 0.5, 0.0};

const double ah_alpha[1] = {1.0};
const double ah_beta[1][1] = {{1.0}};
const double ah_csol[2] = {0.5, 0.5};
const double ah_cerr[2] = {0.5, -0.5};
const double ah_cmid[2] = {0.5, 0.0};

// Prince & Dormand 8(7)13M, numerators / denominators as printed in adaptive_solver/dopri8.py
const double d8_alpha[13] = {1.0 / 18,          1.0 / 12, 1.0 / 8,
                             5.0 / 16,          3.0 / 8,  59.0 / 400,
                             93.0 / 200,        5490023248.0 / 9719169821.0,
                             13.0 / 20,         1201146811.0 / 1299019798.0,
                             1.0,               1.0,      1.0};
const double d8_beta[13][13] = {
    {1.0 / 18},
    {1.0 / 48, 1.0 / 16},
    {1.0 / 32, 0, 3.0 / 32},
    {5.0 / 16, 0, -75.0 / 64, 75.0 / 64},
    {3.0 / 80, 0, 0, 3.0 / 16, 3.0 / 20},
    {29443841.0 / 614563906, 0, 0, 77736538.0 / 692538347, -28693883.0 / 1125000000, 23124283.0 / 1800000000},
    {16016141.0 / 946692911, 0, 0, 61564180.0 / 158732637, 22789713.0 / 633445777, 545815736.0 / 2771057229,
     -180193667.0 / 1043307555},
    {39632708.0 / 573591083, 0, 0, -433636366.0 / 683701615, -421739975.0 / 2616292301, 100302831.0 / 723423059,
     790204164.0 / 839813087, 800635310.0 / 3783071287},
    {246121993.0 / 1340847787, 0, 0, -37695042795.0 / 15268766246, -309121744.0 / 1061227803,
     -12992083.0 / 490766935, 6005943493.0 / 2108947869, 393006217.0 / 1396673457, 123872331.0 / 1001029789},
    {-1028468189.0 / 846180014, 0, 0, 8478235783.0 / 508512852, 1311729495.0 / 1432422823,
     -10304129995.0 / 1701304382, -48777925059.0 / 3047939560, 15336726248.0 / 1032824649,
     -45442868181.0 / 3398467696, 3065993473.0 / 597172653},
    {185892177.0 / 718116043, 0, 0, -3185094517.0 / 667107341, -477755414.0 / 1098053517,
     -703635378.0 / 230739211, 5731566787.0 / 1027545527, 5232866602.0 / 850066563, -4093664535.0 / 808688257,
     3962137247.0 / 1805957418, 65686358.0 / 487910083},
    {403863854.0 / 491063109, 0, 0, -5068492393.0 / 434740067, -411421997.0 / 543043805, 652783627.0 / 914296604,
     11173962825.0 / 925320556, -13158990841.0 / 6184727034, 3936647629.0 / 1978049680, -160528059.0 / 685178525,
     248638103.0 / 1413531060, 0},
    {14005451.0 / 335480064, 0, 0, 0, 0, -59238493.0 / 1068277825, 181606767.0 / 758867731,
     561292985.0 / 797845732, -1041891430.0 / 1371343529, 760417239.0 / 1151165299, 118820643.0 / 751138087,
     -528747749.0 / 2220607170, 1.0 / 4},
};
const double d8_csol[14] = {14005451.0 / 335480064,      0, 0, 0, 0,
                            -59238493.0 / 1068277825,    181606767.0 / 758867731,
                            561292985.0 / 797845732,     -1041891430.0 / 1371343529,
                            760417239.0 / 1151165299,    118820643.0 / 751138087,
                            -528747749.0 / 2220607170,   1.0 / 4, 0};
// fifth-order (7th for the pair) weights subtracted from c_sol to form c_error (dopri8.py:125-140)
const double d8_clow[14] = {13451932.0 / 455176623,   0, 0, 0, 0,
                            -808719846.0 / 976000145, 1757004468.0 / 5645159321,
                            656045339.0 / 265891186,  -3867574721.0 / 1518517206,
                            465885868.0 / 322736535,  53011238.0 / 667516719,
                            2.0 / 45,                 0, 0};
// dense-output weights at the midpoint: quintics in h (dopri8.py:145-238), evaluated at h = 1/2 and
// divided by 1/h; {column, h^5, h^4, h^3, h^2, h^1, constant}
const double d8_midpoly[10][7] = {
    {0, -6.3448349392860401388, 22.1396504998094068976, -30.0610568289666450593, 19.9990069333683970610,
     -6.6910181737837595697, 1.0},
    {5, -39.6107919852202505218, 116.4422149550342161651, -121.4999627731334642623, 52.2273532792945524050,
     -7.6142658045872677172, 0},
    {6, 20.3761213808791436958, -67.1451318825957197185, 83.1721004639847717481, -46.8919164181093621583,
     10.7281392630428866124, 0},
    {7, 7.3347098826795362023, -16.5672243527496524646, 9.5724507555993664382, -0.1890893225010595467,
     0.5526637063753648783, 0},
    {8, 32.8801774352459155182, -89.9916014847245016028, 87.8406057677205645007, -35.7075975946222072821,
     4.2186562625665153803, 0},
    {9, -10.1588990526426760954, 22.6237489648532849093, -17.4152107770762969005, 6.2736448083240352160,
     -0.6627209125361597559, 0},
    {10, -12.5401268098782561200, 32.2362340167355370113, -28.5903289514790976966, 10.3160881272450748458,
     -1.2636789001135462218, 0},
    {11, 29.5553001484516038033, -82.1020315488359848644, 81.6630950584341412934, -34.7650769866611817349,
     5.4106037898590422230, 0},
    {12, -41.7923486424390588923, 116.2662185791119533462, -114.9375291377009418170, 47.7457971078225540396,
     -7.0321379067945741781, 0},
    {13, 20.3006925822100825485, -53.9020777466385396792, 50.2558364226176017553, -19.0082099341608028453,
     2.3537586759714983486, 0},
};
}  // namespace tabs

static void fill_tab(RkTab &t, int S, int order, const double *alpha, const double *beta, int ldb,
                     const double *csol, const double *cerr, const double *cmid) {
  t = RkTab{};
  t.S = S;
  t.order = order;
  for (int i = 0; i < S; ++i) {
    t.alpha[i] = (float)alpha[i];
    for (int j = 0; j <= i; ++j) t.beta[i][j] = (float)beta[i * ldb + j];
  }
  for (int j = 0; j <= S; ++j) {
    t.csol[j] = (float)csol[j];
    t.cerr[j] = (float)cerr[j];
    t.cmid[j] = (float)cmid[j];
  }
  // base_adaptive_solver_rk.py:172-176, tested on the float64 tableau like the reference
  t.fsal = (csol[S] == 0.0) ? 1 : 0;
  for (int j = 0; j < S; ++j)
    if (csol[j] != beta[(S - 1) * ldb + j]) t.fsal = 0;
}

bool make_tab(int method, RkTab &t) {
  using namespace tabs;
  switch (method) {
    case XDE_RK_DOPRI5_TABLE: {
      // the Dormand-Prince coefficients live once, in xde_common.cuh (struct DP, rounded to fp32 there);
      // c_sol == beta[-1] with a trailing zero by construction (DP::csol), i.e. FSAL
      t = RkTab{};
      t.S = 6;
      t.order = 5;
      t.fsal = 1;
      for (int i = 0; i < 6; ++i) {
        t.alpha[i] = DP::alpha(i);
        for (int j = 0; j <= i; ++j) t.beta[i][j] = DP::beta(i, j);
      }
      for (int j = 0; j < 7; ++j) {
        t.csol[j] = DP::csol(j);
        t.cerr[j] = DP::cerr(j);
        t.cmid[j] = DP::cmid(j);
      }
      return true;
    }
    case XDE_RK_BOSH3:
      fill_tab(t, 3, 3, bs_alpha, &bs_beta[0][0], 3, bs_csol, bs_cerr, bs_cmid);
      return true;
    case XDE_RK_FEHLBERG2:
      fill_tab(t, 2, 2, fe_alpha, &fe_beta[0][0], 2, fe_csol, fe_cerr, fe_cmid);
      return true;
    case XDE_RK_ADAPTIVE_HEUN:
      fill_tab(t, 1, 2, ah_alpha, &ah_beta[0][0], 1, ah_csol, ah_cerr, ah_cmid);
      return true;
    case XDE_RK_DOPRI8: {
      double cerr[14], cmid[14] = {};
      for (int j = 0; j < 14; ++j) cerr[j] = d8_csol[j] - d8_clow[j];
      cerr[12] = 1.0 / 4;  // dopri8.py:139: the last two entries are written out, not differences
      cerr[13] = 0;
      const double h = 0.5;
      for (int r = 0; r < 10; ++r) {
        const double *c = d8_midpoly[r];
        double v = c[1] * (h * h * h * h * h) + c[2] * (h * h * h * h);  // left to right, like the Python source
        v = v + c[3] * (h * h * h);
        v = v + c[4] * (h * h);
        v = v + c[5] * h;
        if (c[6] != 0.0) v = v + c[6];
        cmid[(int)c[0]] = v / (1.0 / h);
      }
      fill_tab(t, 13, 8, d8_alpha, &d8_beta[0][0], 13, d8_csol, cerr, cmid);
      return true;
    }
  }
  return false;
}

// ---- device -------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ float rms_vec(const float (&v)[D]) {
  double acc = 0.0;
#pragma unroll
  for (int e = 0; e < D; ++e) acc += (double)(v[e] * v[e]);
  return rms_from_sumsq(acc, (double)D);
}

template <int D, int PRE>
__global__ void __launch_bounds__(kRkThreads) adaptive_rk_small_kernel(const RkParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned long long s_cnt[3];
  __shared__ int s_status;
  const int H = p.field.h;
  const int S = p.tab.S;
  float *sw = smem;
  float *st = sw + SmallRec<D>::floats(H);       // t_span copy (solver time: negated when rev)
  float *sk = st + ((p.T + 3) / 4) * 4;          // stages [S+1][D][kRkThreads]
  load_small_field<D>(sw, p.field);
  const bool rev = p.t_span[1] < p.t_span[0];    // repair R5: s = -t, f~(s, y) = -f(-s, y)
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) st[i] = rev ? -p.t_span[i] : p.t_span[i];
  if (threadIdx.x == 0) {
    s_cnt[0] = s_cnt[1] = s_cnt[2] = 0ull;
    s_status = 0;
  }
  __syncthreads();

  const xde_ctrl_opts_t o = p.o;
  const RkTab &tb = p.tab;
  const float fsign = rev ? -1.0f : 1.0f;
  const int tid = threadIdx.x;
  auto K = [&](int j, int e) -> float & { return sk[(j * D + e) * kRkThreads + tid]; };
  unsigned long long n_att = 0, n_acc = 0, n_fe = 0;
  int status = 0;

  for (long long traj = (long long)blockIdx.x * blockDim.x + tid; traj < p.B;
       traj += (long long)gridDim.x * blockDim.x) {
    float y0[D], yi[D], fo[D];
#pragma unroll
    for (int e = 0; e < D; ++e) {
      y0[e] = p.y0[traj * D + e];
      p.out[traj * D + e] = y0[e];  // solution[0] = y0
    }
    float t0 = st[0], dt;
    // _before_integrate: f0 (base_adaptive_solver_rk.py:83)
    mlp_eval_small<D, PRE>(sw, H, y0, fo);
#pragma unroll
    for (int e = 0; e < D; ++e) K(0, e) = fo[e] * fsign;
    if (o.first_step == o.first_step) {
      dt = o.first_step;
      n_fe += 1;
    } else {
      // select_initial_step (solver/base_adaptive_solver.py:33-72), order = self.order - 1
      float scale[D], v[D];
#pragma unroll
      for (int e = 0; e < D; ++e) {
        scale[e] = o.atol + fabsf(y0[e]) * o.rtol;
        v[e] = __fdiv_rn(y0[e], scale[e]);
      }
      const float d0 = fabsf(rms_vec<D>(v));
#pragma unroll
      for (int e = 0; e < D; ++e) v[e] = __fdiv_rn(K(0, e), scale[e]);
      const float d1 = fabsf(rms_vec<D>(v));
      float h0 = (d0 < 1e-5f || d1 < 1e-5f) ? 1e-6f : __fdiv_rn(0.01f * d0, d1);
      h0 = fabsf(h0);
#pragma unroll
      for (int e = 0; e < D; ++e) yi[e] = K(0, e) * h0 + y0[e];
      mlp_eval_small<D, PRE>(sw, H, yi, fo);
#pragma unroll
      for (int e = 0; e < D; ++e) v[e] = __fdiv_rn(fo[e] * fsign - K(0, e), scale[e]);
      const float d2 = fabsf(__fdiv_rn(rms_vec<D>(v), h0));
      float h1;
      if (d1 <= 1e-15f && d2 <= 1e-15f) {
        h1 = fmaxf(1e-6f, h0 * 1e-3f);
      } else {
        const float mx = (d2 > d1) ? d2 : d1;
        const float arg = __fdiv_rn(0.01f, mx);
        h1 = (arg > 0.0f && arg < INFINITY) ? rootp(arg, tb.order) : arg;
      }
      h1 = fabsf(h1);
      dt = fminf(100.0f * h0, h1);
      n_fe += 3;
    }

    int i_out = 1, n_steps = 0, n_logged = 0;
    // next_*_index = min(bisect.bisect(list, t_span[0]), len - 1)  (:109-114); values in solver time
    const float gsign = rev ? -1.0f : 1.0f;
    int step_idx = 0, jump_idx = 0;
    while (step_idx < p.n_step && !(t0 < gsign * p.step_t[step_idx])) step_idx++;
    if (step_idx > p.n_step - 1) step_idx = p.n_step - 1;
    while (jump_idx < p.n_jump && !(t0 < gsign * p.jump_t[jump_idx])) jump_idx++;
    if (jump_idx > p.n_jump - 1) jump_idx = p.n_jump - 1;
    while (i_out < p.T) {
      // assertions of step / _adaptive_step (base_adaptive_solver_rk.py:120-122, 200-203)
      int bad = 0;
      if (!(n_steps < o.max_num_steps)) bad = XDE_ST_MAX_STEPS;
      else if (!(t0 + dt > t0)) bad = XDE_ST_DT_UNDERFLOW;
      else {
#pragma unroll
        for (int e = 0; e < D; ++e)
          if (!(fabsf(y0[e]) < INFINITY)) bad = XDE_ST_NONFINITE_STATE;
      }
      if (bad) {  // abort this trajectory: remaining outputs are NaN
        status = max(status, bad);
        for (int i = i_out; i < p.T; ++i)
#pragma unroll
          for (int e = 0; e < D; ++e) p.out[((long long)i * p.B + traj) * D + e] = NAN;
        break;
      }
      float t1 = t0 + dt;
      // "Make step, respecting prescribed grid points" (:209-224): step_t first, then jump_t
      bool on_step = false, on_jump = false;
      if (p.n_step > 0) {
        const float nt = gsign * p.step_t[step_idx];
        on_step = (t0 < nt) && (nt < t0 + dt);
        if (on_step) {
          t1 = nt;
          dt = t1 - t0;
        }
      }
      if (p.n_jump > 0) {
        const float nj = gsign * p.jump_t[jump_idx];
        on_jump = (t0 < nj) && (nj < t0 + dt);
        if (on_jump) {
          on_step = false;
          t1 = nj;
          dt = t1 - t0;
        }
      }
      // _runge_kutta_step (:129-181): stage input y0 + sum_j k_j (beta_ij dt), products first
      for (int i = 0; i < S; ++i) {
#pragma unroll
        for (int e = 0; e < D; ++e) {
          float s = K(0, e) * (tb.beta[i][0] * dt);
          for (int j = 1; j <= i; ++j) s = s + K(j, e) * (tb.beta[i][j] * dt);
          yi[e] = y0[e] + s;
        }
        mlp_eval_small<D, PRE>(sw, H, yi, fo);
#pragma unroll
        for (int e = 0; e < D; ++e) K(i + 1, e) = fo[e] * fsign;
      }
      if (!tb.fsal) {  // :172-178; f1 = k[..., -1] either way
#pragma unroll
        for (int e = 0; e < D; ++e) {
          float s = K(0, e) * (dt * tb.csol[0]);
          for (int j = 1; j <= S; ++j) s = s + K(j, e) * (dt * tb.csol[j]);
          yi[e] = y0[e] + s;
        }
      }
      float v[D];
#pragma unroll
      for (int e = 0; e < D; ++e) {
        float s = K(0, e) * (dt * tb.cerr[0]);
        for (int j = 1; j <= S; ++j) s = s + K(j, e) * (dt * tb.cerr[j]);
        const float tol = o.atol + o.rtol * fmaxf(fabsf(y0[e]), fabsf(yi[e]));
        v[e] = __fdiv_rn(s, tol);
      }
      const float ratio = fabsf(rms_vec<D>(v));
      bool accept = (ratio <= 1.0f);
      if (dt > o.max_step) accept = false;
      if (dt <= o.min_step) accept = true;
      n_att++;
      n_fe += (unsigned)S;
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
      // optimal_step_size(dt, ratio, safety, ifactor, dfactor, self.order).clip(min_step, max_step)
      float dt_next;
      if (ratio == 0.0f) {
        dt_next = dt * o.ifactor;
      } else {
        const float dfac = (ratio < 1.0f) ? 1.0f : o.dfactor;
        const float pw = (ratio > 0.0f && ratio < INFINITY) ? rootp(ratio, tb.order) : ratio;
        dt_next = dt * fminf(o.ifactor, fmaxf(__fdiv_rn(o.safety, pw), dfac));
      }
      dt_next = fminf(fmaxf(dt_next, o.min_step), o.max_step);
      if (accept) {
        n_acc++;
        // _interp_fit + interp_fit (:286-292, utils/ode_utils.py:28-49); outputs inside [t0, t1] right away
        float ce[D], cd[D], cc[D], cb[D], ca[D];
        const float two_dt = 2.0f * dt;
#pragma unroll
        for (int e = 0; e < D; ++e) {
          float s = K(0, e) * (dt * tb.cmid[0]);
          for (int j = 1; j <= S; ++j) s = s + K(j, e) * (dt * tb.cmid[j]);
          const float ym = y0[e] + s;
          const float F0 = K(0, e), F1 = K(S, e), Y0 = y0[e], Y1 = yi[e];
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
          y0[e] = yi[e];
          K(0, e) = K(S, e);
        }
        t0 = t1;
        if (on_step && step_idx != p.n_step - 1) step_idx++;
        if (on_jump) {  // past a discontinuity: f1 = self.func(t_next, y_next)  (:263-273)
          if (jump_idx != p.n_jump - 1) jump_idx++;
          mlp_eval_small<D, PRE>(sw, H, y0, fo);
#pragma unroll
          for (int e = 0; e < D; ++e) K(0, e) = fo[e] * fsign;
          n_fe += 1;
        }
      }
      dt = dt_next;
    }
    if (p.log_counts) p.log_counts[traj] = n_logged;
  }

  // ---- stats: CTA -> global ----
  atomicAdd(&s_cnt[0], n_att);
  atomicAdd(&s_cnt[1], n_acc);
  atomicAdd(&s_cnt[2], n_fe);
  atomicMax(&s_status, status);
  __syncthreads();
  if (threadIdx.x == 0 && p.stats) {
    atomicAdd(&p.stats->n_attempts, s_cnt[0]);
    atomicAdd(&p.stats->n_accepted, s_cnt[1]);
    atomicAdd(&p.stats->nfe, s_cnt[2]);
    atomicMax(&p.stats->status, s_status);
  }
}

template <int D, int PRE>
static int launch_rk(const RkParams &p, cudaStream_t stream) {
  const size_t smem = sizeof(float) * ((size_t)SmallRec<D>::floats(p.field.h) + ((p.T + 3) / 4) * 4 +
                                       (size_t)(p.tab.S + 1) * D * kRkThreads);
  XDE_REQUIRE(smem <= 200 * 1024, XDE_E_UNSUPPORTED_FIELD,
              "adaptive RK: field (H=%d) + t_span (T=%d) + %d stages exceed shared memory", p.field.h, p.T,
              p.tab.S + 1);
  auto kern = adaptive_rk_small_kernel<D, PRE>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  XDE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRkThreads, smem));
  if (per_sm < 1) per_sm = 1;
  long long want = (p.B + kRkThreads - 1) / kRkThreads;
  long long grid = (long long)sm_count() * per_sm;  // persistent: a whole number of CTAs per SM
  if (grid > want) grid = want;
  kern<<<(unsigned)grid, kRkThreads, smem, stream>>>(p);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  return XDE_OK;
}

template <int D>
static int rk_dispatch_pre(const RkParams &p, cudaStream_t s) {
  switch (p.field.pre) {
    case XDE_PRE_ID: return launch_rk<D, XDE_PRE_ID>(p, s);
    case XDE_PRE_SQUARE: return launch_rk<D, XDE_PRE_SQUARE>(p, s);
    case XDE_PRE_CUBE: return launch_rk<D, XDE_PRE_CUBE>(p, s);
  }
  set_last_error("unknown pre-activation %d", p.field.pre);
  return XDE_E_BAD_ARG;
}

}  // namespace xde

namespace xde {
int adaptive_rk_tile(int method, const xde_mlp_field_t *field, const float *y0, long long B, const float *t_span, int T,
                     const xde_ctrl_opts_t *opts, const float *step_t, int n_step, const float *jump_t, int n_jump,
                     float *out, xde_stats_t *stats, const xde_attempt_log_t *log, cudaStream_t s);  // xde_tile_adaptive.cu
}

extern "C" XDE_EXPORT int xde_adaptive_rk_mlp_grid_f32(int32_t method, const xde_mlp_field_t *field,
                                                       const float *y0, int64_t B, const float *t_span, int32_t T,
                                                       const xde_ctrl_opts_t *opts, int32_t controller,
                                                       const float *step_t, int32_t n_step, const float *jump_t,
                                                       int32_t n_jump, float *out, xde_stats_t *stats,
                                                       const xde_attempt_log_t *log, void *stream) {
  using namespace xde;
  const bool grid_pts = (n_step > 0 || n_jump > 0);
  if (method == XDE_RK_DOPRI5 && !grid_pts)  // the tuned kernels (both controllers)
    return xde_dopri5_mlp_f32(field, y0, B, t_span, T, opts, controller, out, stats, log, stream);
  if (method == XDE_RK_DOPRI5) method = XDE_RK_DOPRI5_TABLE;  // same arithmetic, table-driven
  XDE_REQUIRE(field && y0 && t_span && opts && out, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(B >= 1 && T >= 2, XDE_E_BAD_ARG, "need B >= 1 and T >= 2 (B=%lld T=%d)", (long long)B, T);
  XDE_REQUIRE(n_step >= 0 && n_jump >= 0 && (n_step == 0 || step_t) && (n_jump == 0 || jump_t), XDE_E_BAD_ARG,
              "step_t / jump_t: negative count or null pointer");
  if (field->d >= 16) {  // large states: the register-tiled kernels (xde_tile_adaptive.cuh)
    XDE_REQUIRE(controller == XDE_CTRL_TRAJECTORY, XDE_E_UNSUPPORTED_FIELD,
                "large states (D >= 16) have the per-trajectory controller only");
    if (stats) XDE_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(xde_stats_t), (cudaStream_t)stream));
    return adaptive_rk_tile(method, field, y0, B, t_span, T, opts, step_t, n_step, jump_t, n_jump, out, stats, log,
                            (cudaStream_t)stream);
  }
  RkParams p{};
  XDE_REQUIRE(make_tab(method, p.tab), XDE_E_BAD_ARG, "unknown Runge-Kutta method %d", method);
  XDE_REQUIRE(controller == XDE_CTRL_TRAJECTORY, XDE_E_UNSUPPORTED_FIELD,
              "the table-driven adaptive kernel (other tableaux, step_t / jump_t) has the per-trajectory "
              "controller only; controller='batch' is fused for plain Dopri5");
  cudaStream_t s = (cudaStream_t)stream;
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
  p.step_t = step_t;
  p.n_step = n_step;
  p.jump_t = jump_t;
  p.n_jump = n_jump;
  if (stats) XDE_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(xde_stats_t), s));
  switch (field->d) {
    case 1: return rk_dispatch_pre<1>(p, s);
    case 2: return rk_dispatch_pre<2>(p, s);
    case 3: return rk_dispatch_pre<3>(p, s);
    case 4: return rk_dispatch_pre<4>(p, s);
    case 5: return rk_dispatch_pre<5>(p, s);
    case 6: return rk_dispatch_pre<6>(p, s);
    case 7: return rk_dispatch_pre<7>(p, s);
    case 8: return rk_dispatch_pre<8>(p, s);
    default:
      set_last_error("adaptive RK: state dim D=%d has no fused kernel (supported: 1..8)", field->d);
      return XDE_E_UNSUPPORTED_FIELD;
  }
}

extern "C" XDE_EXPORT int xde_adaptive_rk_mlp_f32(int32_t method, const xde_mlp_field_t *field, const float *y0,
                                                  int64_t B, const float *t_span, int32_t T,
                                                  const xde_ctrl_opts_t *opts, int32_t controller, float *out,
                                                  xde_stats_t *stats, const xde_attempt_log_t *log, void *stream) {
  return xde_adaptive_rk_mlp_grid_f32(method, field, y0, B, t_span, T, opts, controller, nullptr, 0, nullptr, 0, out,
                                      stats, log, stream);
}
