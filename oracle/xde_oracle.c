/*
 * xde_oracle.c -- CPU restatement of the PaddleXDE integration hot path (see xde_oracle.h).
 * TEST INFRASTRUCTURE ONLY; never linked into the product.
 *
 * Arithmetic: fp32, round-to-nearest-even, no flush-to-zero, no implicit contraction (build with
 * -ffp-contract=off); fused multiply-adds appear only as explicit fmaf().  Reductions that the
 * reference leaves order-unspecified (RMS norm mean, batch sums) accumulate in fp64 so that their
 * fp32 result does not depend on the summation order.  DESIGN.md "Arithmetic specification" is the
 * normative text; the CUDA kernels implement the same specification independently.
 *
 * Paths in comments are relative to /root/reference.
 */
#include "xde_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* Dormand-Prince tableau: solver/adaptive_solver/dopri5.py:5-55, authored in float64 and cast  */
/* once to the state dtype (solver/base_adaptive_solver_rk.py:73-79).                           */
/* ------------------------------------------------------------------------------------------ */
static const double DP_ALPHA64[6] = {1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
static const double DP_BETA64[6][6] = {
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84},
};
static const double DP_CERR64[7] = {
    35.0 / 384 - 1951.0 / 21600,
    0,
    500.0 / 1113 - 22642.0 / 50085,
    125.0 / 192 - 451.0 / 720,
    -2187.0 / 6784 - -12231.0 / 42400,
    11.0 / 84 - 649.0 / 6300,
    -1.0 / 60.0,
};
static const double DP_CMID64[7] = {
    6025192743.0 / 30085553152.0 / 2,   0,
    51252292925.0 / 65400821598.0 / 2,  -2691868925.0 / 45128329728.0 / 2,
    187940372067.0 / 1594534317056.0 / 2, -1776094331.0 / 19743644256.0 / 2,
    11237099.0 / 235043384.0 / 2,
};

/* ------------------------------------------------------------------------------------------ */
/* The other embedded Runge-Kutta tableaux of solver/__init__.py:1-6, same driver:  */
/* Bosh3 adaptive_solver/bosh3.py:5-27, Fehlberg2 fehlberg2.py:5-22, AdaptiveHeun                 */
/* adaptive_heun.py:5-27, Dopri8 dopri8.py:5-252.  Authored in float64, cast once to fp32.        */
/* ------------------------------------------------------------------------------------------ */
#define ORC_MAX_STAGES 13 /* len(alpha) of Dopri8; k has one more column */
typedef struct {
  int S;     /* len(tableau.alpha): field evaluations per attempt */
  int order; /* solver.order: exponent of optimal_step_size, and of select_initial_step (order-1+1) */
  int fsal;  /* c_sol[-1] == 0 and c_sol[:-1] == beta[-1]  (base_adaptive_solver_rk.py:172-176) */
  float alpha[ORC_MAX_STAGES], beta[ORC_MAX_STAGES][ORC_MAX_STAGES];
  float csol[ORC_MAX_STAGES + 1], cerr[ORC_MAX_STAGES + 1], cmid[ORC_MAX_STAGES + 1];
} rk_tab_t;

static const double BS_ALPHA64[3] = {1.0 / 2, 3.0 / 4, 1.0};
static const double BS_BETA64[3][3] = {{1.0 / 2, 0, 0}, {0.0, 3.0 / 4, 0}, {2.0 / 9, 1.0 / 3, 4.0 / 9}};
static const double BS_CSOL64[4] = {2.0 / 9, 1.0 / 3, 4.0 / 9, 0.0};
static const double BS_CERR64[4] = {2.0 / 9 - 7.0 / 24, 1.0 / 3 - 1.0 / 4, 4.0 / 9 - 1.0 / 3, -1.0 / 8};
static const double BS_CMID64[4] = {0.0, 0.5, 0.0, 0.0};

static const double FE_ALPHA64[2] = {1.0 / 2, 1.0};
static const double FE_BETA64[2][2] = {{1.0 / 2, 0}, {1.0 / 256, 255.0 / 256}};
static const double FE_CSOL64[3] = {1.0 / 512, 255.0 / 256, 1.0 / 512};
static const double FE_CERR64[3] = {-1.0 / 512, 0, 1.0 / 512};
static const double FE_CMID64[3] = {0.0, 0.5, 0.0};

static const double AH_ALPHA64[1] = {1.0};
static const double AH_BETA64[1][1] = {{1.0}};
static const double AH_CSOL64[2] = {0.5, 0.5};
static const double AH_CERR64[2] = {0.5, -0.5};
static const double AH_CMID64[2] = {0.5, 0.0};

static const double D8_ALPHA64[13] = {1.0 / 18, 1.0 / 12, 1.0 / 8, 5.0 / 16, 3.0 / 8, 59.0 / 400, 93.0 / 200,
                                      5490023248.0 / 9719169821.0, 13.0 / 20, 1201146811.0 / 1299019798.0,
                                      1.0, 1.0, 1.0};
static const double D8_BETA64[13][13] = {
    {1.0 / 18},
    {1.0 / 48, 1.0 / 16},
    {1.0 / 32, 0, 3.0 / 32},
    {5.0 / 16, 0, -75.0 / 64, 75.0 / 64},
    {3.0 / 80, 0, 0, 3.0 / 16, 3.0 / 20},
    {29443841.0 / 614563906, 0, 0, 77736538.0 / 692538347, -28693883.0 / 1125000000, 23124283.0 / 1800000000},
    {16016141.0 / 946692911, 0, 0, 61564180.0 / 158732637, 22789713.0 / 633445777, 545815736.0 / 2771057229,
     -180193667.0 / 1043307555},
    {39632708.0 / 573591083, 0, 0, -433636366.0 / 683701615, -421739975.0 / 2616292301,
     100302831.0 / 723423059, 790204164.0 / 839813087, 800635310.0 / 3783071287},
    {246121993.0 / 1340847787, 0, 0, -37695042795.0 / 15268766246, -309121744.0 / 1061227803,
     -12992083.0 / 490766935, 6005943493.0 / 2108947869, 393006217.0 / 1396673457,
     123872331.0 / 1001029789},
    {-1028468189.0 / 846180014, 0, 0, 8478235783.0 / 508512852, 1311729495.0 / 1432422823,
     -10304129995.0 / 1701304382, -48777925059.0 / 3047939560, 15336726248.0 / 1032824649,
     -45442868181.0 / 3398467696, 3065993473.0 / 597172653},
    {185892177.0 / 718116043, 0, 0, -3185094517.0 / 667107341, -477755414.0 / 1098053517,
     -703635378.0 / 230739211, 5731566787.0 / 1027545527, 5232866602.0 / 850066563,
     -4093664535.0 / 808688257, 3962137247.0 / 1805957418, 65686358.0 / 487910083},
    {403863854.0 / 491063109, 0, 0, -5068492393.0 / 434740067, -411421997.0 / 543043805,
     652783627.0 / 914296604, 11173962825.0 / 925320556, -13158990841.0 / 6184727034,
     3936647629.0 / 1978049680, -160528059.0 / 685178525, 248638103.0 / 1413531060, 0},
    {14005451.0 / 335480064, 0, 0, 0, 0, -59238493.0 / 1068277825, 181606767.0 / 758867731,
     561292985.0 / 797845732, -1041891430.0 / 1371343529, 760417239.0 / 1151165299,
     118820643.0 / 751138087, -528747749.0 / 2220607170, 1.0 / 4},
};
static const double D8_CSOL64[14] = {14005451.0 / 335480064, 0, 0, 0, 0, -59238493.0 / 1068277825,
                                     181606767.0 / 758867731, 561292985.0 / 797845732,
                                     -1041891430.0 / 1371343529, 760417239.0 / 1151165299,
                                     118820643.0 / 751138087, -528747749.0 / 2220607170, 1.0 / 4, 0};
static const double D8_CERR64[14] = {14005451.0 / 335480064 - 13451932.0 / 455176623, 0, 0, 0, 0,
                                     -59238493.0 / 1068277825 - -808719846.0 / 976000145,
                                     181606767.0 / 758867731 - 1757004468.0 / 5645159321,
                                     561292985.0 / 797845732 - 656045339.0 / 265891186,
                                     -1041891430.0 / 1371343529 - -3867574721.0 / 1518517206,
                                     760417239.0 / 1151165299 - 465885868.0 / 322736535,
                                     118820643.0 / 751138087 - 53011238.0 / 667516719,
                                     -528747749.0 / 2220607170 - 2.0 / 45, 1.0 / 4, 0};
/* dopri8.py:149-238: C_mid[j] = (quintic in h, without / with the constant 1) / (1/h) at h = 1/2;
 * rows: column j, coefficients of h^5 .. h^1, constant */
static const double D8_MIDPOLY[10][7] = {
    {0, -6.3448349392860401388, 22.1396504998094068976, -30.0610568289666450593, 19.9990069333683970610,
     -6.6910181737837595697, 1.0},
    {5, -39.6107919852202505218, 116.4422149550342161651, -121.4999627731334642623, 52.2273532792945524050,
     -7.6142658045872677172, 0.0},
    {6, 20.3761213808791436958, -67.1451318825957197185, 83.1721004639847717481, -46.8919164181093621583,
     10.7281392630428866124, 0.0},
    {7, 7.3347098826795362023, -16.5672243527496524646, 9.5724507555993664382, -0.1890893225010595467,
     0.5526637063753648783, 0.0},
    {8, 32.8801774352459155182, -89.9916014847245016028, 87.8406057677205645007, -35.7075975946222072821,
     4.2186562625665153803, 0.0},
    {9, -10.1588990526426760954, 22.6237489648532849093, -17.4152107770762969005, 6.2736448083240352160,
     -0.6627209125361597559, 0.0},
    {10, -12.5401268098782561200, 32.2362340167355370113, -28.5903289514790976966, 10.3160881272450748458,
     -1.2636789001135462218, 0.0},
    {11, 29.5553001484516038033, -82.1020315488359848644, 81.6630950584341412934, -34.7650769866611817349,
     5.4106037898590422230, 0.0},
    {12, -41.7923486424390588923, 116.2662185791119533462, -114.9375291377009418170, 47.7457971078225540396,
     -7.0321379067945741781, 0.0},
    {13, 20.3006925822100825485, -53.9020777466385396792, 50.2558364226176017553, -19.0082099341608028453,
     2.3537586759714983486, 0.0},
};

static void tab_fill(rk_tab_t *t, int S, int order, const double *alpha, const double *beta, int ldb,
                     const double *csol, const double *cerr, const double *cmid) {
  memset(t, 0, sizeof(*t));
  t->S = S;
  t->order = order;
  for (int i = 0; i < S; ++i) {
    t->alpha[i] = (float)alpha[i];
    for (int j = 0; j <= i; ++j) t->beta[i][j] = (float)beta[i * ldb + j];
  }
  for (int j = 0; j <= S; ++j) {
    t->csol[j] = (float)csol[j];
    t->cerr[j] = (float)cerr[j];
    t->cmid[j] = (float)cmid[j];
  }
  /* the FSAL test of base_adaptive_solver_rk.py:172-176, on the float64 tableau like the reference */
  t->fsal = (csol[S] == 0.0);
  for (int j = 0; j < S; ++j)
    if (csol[j] != beta[(S - 1) * ldb + j]) t->fsal = 0;
}

static int rk_tab_init(rk_tab_t *t, int method) {
  switch (method) {
    case ORC_RK_DOPRI5: {
      double csol[7];
      for (int j = 0; j < 6; ++j) csol[j] = DP_BETA64[5][j];
      csol[6] = 0.0;
      tab_fill(t, 6, 5, DP_ALPHA64, &DP_BETA64[0][0], 6, csol, DP_CERR64, DP_CMID64);
      return 0;
    }
    case ORC_RK_BOSH3:
      tab_fill(t, 3, 3, BS_ALPHA64, &BS_BETA64[0][0], 3, BS_CSOL64, BS_CERR64, BS_CMID64);
      return 0;
    case ORC_RK_FEHLBERG2:
      tab_fill(t, 2, 2, FE_ALPHA64, &FE_BETA64[0][0], 2, FE_CSOL64, FE_CERR64, FE_CMID64);
      return 0;
    case ORC_RK_ADAPTIVE_HEUN:
      tab_fill(t, 1, 2, AH_ALPHA64, &AH_BETA64[0][0], 1, AH_CSOL64, AH_CERR64, AH_CMID64);
      return 0;
    case ORC_RK_DOPRI8: {
      double cmid[14] = {0};
      const double h = 0.5;
      for (int r = 0; r < 10; ++r) {
        const double *c = D8_MIDPOLY[r];
        /* Python: c5*h**5 + c4*h**4 + c3*h**3 + c2*h**2 + c1*h (+ 1.0), left to right, then / (1/h) */
        double v = c[1] * (h * h * h * h * h) + c[2] * (h * h * h * h);
        v = v + c[3] * (h * h * h);
        v = v + c[4] * (h * h);
        v = v + c[5] * h;
        if (c[6] != 0.0) v = v + c[6];
        cmid[(int)c[0]] = v / (1.0 / h);
      }
      tab_fill(t, 13, 8, D8_ALPHA64, &D8_BETA64[0][0], 13, D8_CSOL64, D8_CERR64, cmid);
      return 0;
    }
  }
  return -1;
}

void orc_default_opts(orc_opts_t *o) {
  /* solver/base_adaptive_solver_rk.py:32-49; functional/odeint.py:14-15 */
  o->rtol = 1e-7f;
  o->atol = 1e-9f;
  o->min_step = 0.0f;
  o->max_step = INFINITY;
  o->first_step = NAN;
  o->safety = 0.9f;
  o->ifactor = 10.0f;
  o->dfactor = 0.2f;
  o->max_num_steps = 2147483647;
}

/* ------------------------------------------------------------------------------------------ */
/* Scalar primitives                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* paddle.tanh in fp32 (example/ode_demo.py:23).  Defined here as the 13/6 rational minimax
 * (the approximation Eigen -- Paddle's CPU elementwise backend -- uses for float tanh):
 * <= 5 ulp over the whole range, built only from correctly rounded +,*,fma,/ so that CPU and
 * GPU agree bit for bit.  NaN propagates. */
float orc_tanhf(float a) {
  const float c = 7.90531110763549805f;
  if (!(a == a)) return a;
  float x = fminf(fmaxf(a, -c), c);
  float x2 = x * x;
  float p = fmaf(x2, -2.76076847742355e-16f, 2.00018790482477e-13f);
  p = fmaf(x2, p, -8.60467152213735e-11f);
  p = fmaf(x2, p, 5.12229709037114e-08f);
  p = fmaf(x2, p, 1.48572235717979e-05f);
  p = fmaf(x2, p, 6.37261928875436e-04f);
  p = fmaf(x2, p, 4.89352455891786e-03f);
  p = x * p;
  float q = fmaf(x2, 1.19825839466702e-06f, 1.18534705686654e-04f);
  q = fmaf(x2, q, 2.26843463243900e-03f);
  q = fmaf(x2, q, 4.89352518554385e-03f);
  float r = p / q;
  return (fabsf(a) < 0.0004f) ? a : r;
}

/* error_ratio ** (1/order) with order = 5 (utils/ode_utils.py:92-95) and the 1/(order+1) power
 * of select_initial_step with order = 4 (solver/base_adaptive_solver.py:70).  Deterministic
 * Newton iteration for the fifth root: integer seed + 4 iterations of x <- (4x + r/x^4)/5.
 * Caller guarantees r finite and > 0. */
float orc_root5f(float r) {
  union {
    float f;
    uint32_t u;
  } v;
  v.f = r;
  v.u = v.u / 5u + 0x32CCCCCCu;
  float x = v.f;
  for (int it = 0; it < 4; ++it) {
    float x2 = x * x;
    float x4 = x2 * x2;
    float q = r / x4;
    x = fmaf(4.0f, x, q) * 0.2f;
  }
  return x;
}

/* r ** (1/p) for the orders of the supported tableaux (same deterministic style as orc_root5f):
 * p = 2: sqrtf (IEEE); p = 8: sqrtf three times; p = 3: integer seed + 4 Newton steps
 * x <- (2x + r/x^2)/3; p = 5: orc_root5f.  Caller guarantees r finite and > 0. */
float orc_rootpf(float r, int32_t p) {
  if (p == 5) return orc_root5f(r);
  if (p == 2) return sqrtf(r);
  if (p == 8) return sqrtf(sqrtf(sqrtf(r)));
  if (p == 3) {
    union {
      float f;
      uint32_t u;
    } v;
    v.f = r;
    v.u = v.u / 3u + 0x2A555555u;
    float x = v.f;
    for (int it = 0; it < 4; ++it) {
      float q = r / (x * x);
      x = fmaf(2.0f, x, q) * (float)(1.0 / 3.0);
    }
    return x;
  }
  return powf(r, 1.0f / (float)p);
}

static inline float pre_act(int pre, float y) {
  /* paddle `y**3` / `y**2` (example/ode_demo.py:33, example/sde_demo.py:183) */
  if (pre == ORC_PRE_CUBE) return (y * y) * y;
  if (pre == ORC_PRE_SQUARE) return y * y;
  return y;
}
static inline float pre_act_grad(int pre, float y) {
  /* pow grad: factor * x^(factor-1) */
  if (pre == ORC_PRE_CUBE) return 3.0f * (y * y);
  if (pre == ORC_PRE_SQUARE) return 2.0f * y;
  return 1.0f;
}

#define ORC_MAX_D 256
#define ORC_MAX_H 1024

/* Reduction over the hidden axis j = 0..H-1 in the order the arithmetic specification fixes
 * (DESIGN.md "Arithmetic specification"): two interleaved fma chains -- one over the even hidden
 * units, one over the odd ones, each started from +0 -- added at the end.  (A matmul's summation
 * order is implementation-defined in Paddle; this is the order the B200 kernels use, where one
 * packed FFMA2 instruction advances both chains.)  term(j) = a[j] * b[j*stride]. */
static float chain2_dot(const float *a, const float *b, int stride, int H) {
  float acc_e = 0.0f, acc_o = 0.0f;
  for (int j = 0; j < H; j += 2) {
    acc_e = fmaf(a[j], b[(size_t)j * stride], acc_e);
    if (j + 1 < H) acc_o = fmaf(a[j + 1], b[(size_t)(j + 1) * stride], acc_o);
  }
  return acc_e + acc_o;
}

/* f = tanh(pre(y) @ W1 + b1) @ W2 + b2.  nn.Linear = matmul then bias add; the first matmul is a
 * sequential-k fma chain (first term a plain product), the second one uses chain2_dot. */
void orc_mlp_eval(const orc_mlp_t *m, const float *y, float *f, float *hbuf) {
  const int D = m->d, H = m->h;
  float u[ORC_MAX_D], hloc[ORC_MAX_H];
  float *h = hbuf ? hbuf : hloc;
  for (int k = 0; k < D; ++k) u[k] = pre_act(m->pre, y[k]);
  for (int j = 0; j < H; ++j) {
    float acc = u[0] * m->w1[j];
    for (int k = 1; k < D; ++k) acc = fmaf(u[k], m->w1[(size_t)k * H + j], acc);
    h[j] = orc_tanhf(acc + m->b1[j]);
  }
  for (int d = 0; d < D; ++d) f[d] = chain2_dot(h, m->w2 + d, D, H) + m->b2[d];
}

void orc_mlp_eval_batch(const orc_mlp_t *m, const float *y, int64_t B, float *f) {
  for (int64_t b = 0; b < B; ++b) orc_mlp_eval(m, y + b * m->d, f + b * m->d, NULL);
}

/* paddle.autograd.grad(f, (y, *params), grad_outputs=c)  (functional/odeint_adjoint.py:108-114);
 * SURVEY Appendix B.  dh = c W2^T ; dz = dh*(1-h*h) ; du = dz W1^T ; dy = du * pre'(y);
 * gW2 = h^T c ; gb2 = c ; gW1 = u^T dz ; gb1 = dz. */
void orc_mlp_vjp(const orc_mlp_t *m, const float *y, const float *c, float *f, float *dy, float *gw1,
                 float *gb1, float *gw2, float *gb2) {
  const int D = m->d, H = m->h;
  float u[ORC_MAX_D], h[ORC_MAX_H], dz[ORC_MAX_H];
  for (int k = 0; k < D; ++k) u[k] = pre_act(m->pre, y[k]);
  orc_mlp_eval(m, y, f, h);
  for (int j = 0; j < H; ++j) {
    float acc = c[0] * m->w2[(size_t)j * D];
    for (int d = 1; d < D; ++d) acc = fmaf(c[d], m->w2[(size_t)j * D + d], acc);
    float s = fmaf(-h[j], h[j], 1.0f); /* tanh' = 1 - h^2, one rounding (fused, as Paddle's GPU tanh_grad) */
    dz[j] = acc * s;
  }
  for (int k = 0; k < D; ++k)
    dy[k] = chain2_dot(dz, m->w1 + (size_t)k * H, 1, H) * pre_act_grad(m->pre, y[k]);
  if (gw1)
    for (int k = 0; k < D; ++k)
      for (int j = 0; j < H; ++j) gw1[(size_t)k * H + j] += u[k] * dz[j];
  if (gb1)
    for (int j = 0; j < H; ++j) gb1[j] += dz[j];
  if (gw2)
    for (int j = 0; j < H; ++j)
      for (int d = 0; d < D; ++d) gw2[(size_t)j * D + d] += h[j] * c[d];
  if (gb2)
    for (int d = 0; d < D; ++d) gb2[d] += c[d];
}

/* ------------------------------------------------------------------------------------------ */
/* Generic adaptive driver on a flat fp32 state (AdaptiveRKSolver, Dopri5)                      */
/* ------------------------------------------------------------------------------------------ */
typedef void (*rhs_fn)(void *ctx, float t, const float *y, float *dy);
typedef float (*norm_fn)(void *ctx, const float *v);

typedef struct {
  int64_t n;
  rhs_fn rhs;
  norm_fn norm;
  void *ctx;
  const orc_opts_t *o;
  int rev; /* repair R5: integrate s = -t with f~(s,y) = -f(-s,y) */
  rk_tab_t tab;
  /* step_t / jump_t (base_adaptive_solver_rk.py:94-114): sorted, >= t0, in solver time; may be empty */
  const float *step_t, *jump_t;
  int n_step, n_jump, step_idx, jump_idx;
  /* _RungeKuttaState (solver/base_adaptive_solver_rk.py:22-24) */
  float *y1, *f1, *coef[5];
  float t0, t1, dt;
  /* work */
  float *k; /* [S+1][n], S <= 13 */
  float *yi, *err, *v, *ymid, *ynew;
  orc_stats_t *st;
  orc_attempt_t *log;
  int64_t log_cap, *log_len;
} drv_t;

static int drv_alloc(drv_t *d, int64_t n) {
  size_t nn = (size_t)n;
  memset(d, 0, sizeof(*d));
  d->n = n;
  const size_t K = ORC_MAX_STAGES + 1;
  float *blk = (float *)malloc(sizeof(float) * nn * (2 + 5 + K + 5));
  if (!blk) return -1;
  d->y1 = blk;
  d->f1 = blk + nn;
  for (int i = 0; i < 5; ++i) d->coef[i] = blk + nn * (2 + i);
  d->k = blk + nn * 7;
  d->yi = blk + nn * (7 + K);
  d->err = blk + nn * (8 + K);
  d->v = blk + nn * (9 + K);
  d->ymid = blk + nn * (10 + K);
  d->ynew = blk + nn * (11 + K);
  rk_tab_init(&d->tab, ORC_RK_DOPRI5); /* callers of other tableaux overwrite d->tab */
  return 0;
}
static void drv_free(drv_t *d) { free(d->y1); }

static void drv_rhs(drv_t *d, float s, const float *y, float *dy) {
  if (d->rev) {
    d->rhs(d->ctx, -s, y, dy);
    for (int64_t e = 0; e < d->n; ++e) dy[e] = -dy[e];
  } else {
    d->rhs(d->ctx, s, y, dy);
  }
  if (d->st) d->st->nfe++;
}

/* solver/base_adaptive_solver.py:33-72 (Hairer II.4), called with order = self.order - 1
 * (solver/base_adaptive_solver_rk.py:85-87), i.e. the exponent is 1/self.order. */
static float drv_select_initial_step(drv_t *d, float t0, const float *y0) {
  const orc_opts_t *o = d->o;
  const int64_t n = d->n;
  float *f0 = d->k;         /* scratch: k[0] */
  float *scale = d->k + n;  /* k[1] */
  float *f1 = d->k + 2 * n; /* k[2] */
  drv_rhs(d, t0, y0, f0);   /* `f0 = self.move(t0, 0, y0)` :47-48 (recomputed, same value) */
  for (int64_t e = 0; e < n; ++e) scale[e] = o->atol + fabsf(y0[e]) * o->rtol;
  for (int64_t e = 0; e < n; ++e) d->v[e] = y0[e] / scale[e];
  float d0 = fabsf(d->norm(d->ctx, d->v));
  for (int64_t e = 0; e < n; ++e) d->v[e] = f0[e] / scale[e];
  float d1 = fabsf(d->norm(d->ctx, d->v));
  float h0;
  if (d0 < 1e-5f || d1 < 1e-5f)
    h0 = 1e-6f;
  else
    h0 = (0.01f * d0) / d1;
  h0 = fabsf(h0);
  for (int64_t e = 0; e < n; ++e) d->yi[e] = f0[e] * h0 + y0[e]; /* fuse: dy*dt + y0 */
  drv_rhs(d, t0 + h0, d->yi, f1);
  for (int64_t e = 0; e < n; ++e) d->v[e] = (f1[e] - f0[e]) / scale[e];
  float d2 = fabsf(d->norm(d->ctx, d->v) / h0);
  float h1;
  if (d1 <= 1e-15f && d2 <= 1e-15f) {
    h1 = fmaxf(1e-6f, h0 * 1e-3f);
  } else {
    float mx = (d2 > d1) ? d2 : d1; /* python max(d1, d2) */
    float arg = 0.01f / mx;
    h1 = (arg > 0.0f && arg < INFINITY) ? orc_rootpf(arg, d->tab.order) : arg;
  }
  h1 = fabsf(h1);
  return fminf(100.0f * h0, h1);
}

/* solver/base_adaptive_solver_rk.py:81-114 */
static void drv_before_integrate(drv_t *d, float t0, const float *y0) {
  const int64_t n = d->n;
  memcpy(d->y1, y0, sizeof(float) * n);
  drv_rhs(d, t0, y0, d->f1);
  float first;
  if (d->o->first_step == d->o->first_step)
    first = d->o->first_step;
  else
    first = drv_select_initial_step(d, t0, y0);
  d->t0 = t0;
  d->t1 = t0;
  d->dt = first;
  for (int i = 0; i < 5; ++i) memcpy(d->coef[i], y0, sizeof(float) * n);
  /* next_*_index = min(bisect.bisect(list, t_span[0]), len - 1)  (:109-114) */
  d->step_idx = d->jump_idx = 0;
  while (d->step_idx < d->n_step && !(t0 < d->step_t[d->step_idx])) d->step_idx++;
  if (d->step_idx > d->n_step - 1) d->step_idx = d->n_step - 1;
  while (d->jump_idx < d->n_jump && !(t0 < d->jump_t[d->jump_idx])) d->jump_idx++;
  if (d->jump_idx > d->n_jump - 1) d->jump_idx = d->n_jump - 1;
}

/* solver/base_adaptive_solver_rk.py:183-284 with _runge_kutta_step :129-181, compute_error_ratio
 * utils/ode_utils.py:80-82, optimal_step_size :85-97, _interp_fit :286-292 + interp_fit :28-49 */
static int drv_adaptive_step(drv_t *d) {
  const orc_opts_t *o = d->o;
  const rk_tab_t *tb = &d->tab;
  const int S = tb->S;
  const int64_t n = d->n;
  float *y0 = d->y1, *f0 = d->f1;
  const float t0 = d->t1;
  float dt = d->dt;
  float t1 = t0 + dt;
  if (!(t0 + dt > t0)) return ORC_DT_UNDERFLOW;
  for (int64_t e = 0; e < n; ++e)
    if (!isfinite(y0[e])) return ORC_NONFINITE_STATE;
  /* "Make step, respecting prescribed grid points" (:209-224): all the step_t handling, then all the
   * jump_t handling */
  int on_step = 0, on_jump = 0;
  if (d->n_step > 0) {
    const float nt = d->step_t[d->step_idx];
    on_step = (t0 < nt) && (nt < t0 + dt);
    if (on_step) {
      t1 = nt;
      dt = t1 - t0;
    }
  }
  if (d->n_jump > 0) {
    const float nj = d->jump_t[d->jump_idx];
    on_jump = (t0 < nj) && (nj < t0 + dt);
    if (on_jump) {
      on_step = 0;
      t1 = nj;
      dt = t1 - t0;
    }
  }

  float *k = d->k;
  memcpy(k, f0, sizeof(float) * n);
  for (int i = 0; i < S; ++i) {
    float ti = (tb->alpha[i] == 1.0f) ? t1 : t0 + tb->alpha[i] * dt;
    float bd[ORC_MAX_STAGES];
    for (int j = 0; j <= i; ++j) bd[j] = tb->beta[i][j] * dt;
    for (int64_t e = 0; e < n; ++e) {
      float s = k[e] * bd[0];
      for (int j = 1; j <= i; ++j) s = s + k[(size_t)j * n + e] * bd[j];
      d->yi[e] = y0[e] + s;
    }
    drv_rhs(d, ti, d->yi, k + (size_t)(i + 1) * n);
  }
  /* FSAL shortcut (true for Dormand-Prince, Bosh3): y1 = yi; otherwise y1 = y0 + sum k (dt c_sol)
   * (:172-179); f1 = k[...,-1] either way */
  float *y1 = d->yi;
  float *f1 = k + (size_t)S * n;
  if (!tb->fsal) {
    float cs[ORC_MAX_STAGES + 1];
    for (int j = 0; j <= S; ++j) cs[j] = dt * tb->csol[j];
    for (int64_t e = 0; e < n; ++e) {
      float s = k[e] * cs[0];
      for (int j = 1; j <= S; ++j) s = s + k[(size_t)j * n + e] * cs[j];
      y1[e] = y0[e] + s;
    }
  }
  float ce[ORC_MAX_STAGES + 1];
  for (int j = 0; j <= S; ++j) ce[j] = dt * tb->cerr[j];
  for (int64_t e = 0; e < n; ++e) {
    float s = k[e] * ce[0];
    for (int j = 1; j <= S; ++j) s = s + k[(size_t)j * n + e] * ce[j];
    d->err[e] = s;
  }
  for (int64_t e = 0; e < n; ++e) {
    float tol = o->atol + o->rtol * fmaxf(fabsf(y0[e]), fabsf(y1[e]));
    d->v[e] = d->err[e] / tol;
  }
  float ratio = fabsf(d->norm(d->ctx, d->v));
  int accept = (ratio <= 1.0f);
  if (dt > o->max_step) accept = 0;
  if (dt <= o->min_step) accept = 1;

  if (d->st) {
    d->st->n_attempts++;
    d->st->n_accepted += accept;
    float m1 = fabsf(ratio - 1.0f);
    if (m1 < d->st->min_abs_ratio_m1) d->st->min_abs_ratio_m1 = m1;
  }
  if (d->log && d->log_len && *d->log_len < d->log_cap) {
    orc_attempt_t *r = &d->log[*d->log_len];
    r->t0 = d->rev ? -t0 : t0;
    r->dt = d->rev ? -dt : dt;
    r->ratio = ratio;
    r->accepted = accept;
    (*d->log_len)++;
  }

  if (accept) {
    float cm[ORC_MAX_STAGES + 1];
    for (int j = 0; j <= S; ++j) cm[j] = dt * tb->cmid[j];
    const float two_dt = 2.0f * dt;
    for (int64_t e = 0; e < n; ++e) {
      float s = k[e] * cm[0];
      for (int j = 1; j <= S; ++j) s = s + k[(size_t)j * n + e] * cm[j];
      float ym = y0[e] + s;
      float F0 = k[e], F1 = f1[e], Y0 = y0[e], Y1 = y1[e];
      float a = (two_dt * (F1 - F0) - 8.0f * (Y1 + Y0)) + 16.0f * ym;
      float b = ((dt * (5.0f * F0 - 3.0f * F1) + 18.0f * Y0) + 14.0f * Y1) - 32.0f * ym;
      float c = ((dt * (F1 - 4.0f * F0) - 11.0f * Y0) - 5.0f * Y1) + 16.0f * ym;
      float dd = dt * F0;
      d->coef[0][e] = Y0;
      d->coef[1][e] = dd;
      d->coef[2][e] = c;
      d->coef[3][e] = b;
      d->coef[4][e] = a;
    }
    memcpy(d->ynew, y1, sizeof(float) * n);
    memcpy(d->y1, d->ynew, sizeof(float) * n);
    memcpy(d->f1, f1, sizeof(float) * n);
    if (on_step && d->step_idx != d->n_step - 1) d->step_idx++;
    if (on_jump) {
      if (d->jump_idx != d->n_jump - 1) d->jump_idx++;
      drv_rhs(d, t1, d->y1, d->f1); /* f1 = self.func(t_next, y_next) past a discontinuity (:269-273) */
    }
  }
  /* optimal_step_size(dt, error_ratio, safety, ifactor, dfactor, self.order) */
  float dt_next;
  if (ratio == 0.0f) {
    dt_next = dt * o->ifactor;
  } else {
    float dfac = (ratio < 1.0f) ? 1.0f : o->dfactor;
    float p = (ratio > 0.0f && ratio < INFINITY) ? orc_rootpf(ratio, tb->order) : ratio;
    float factor = fminf(o->ifactor, fmaxf(o->safety / p, dfac));
    dt_next = dt * factor;
  }
  dt_next = fminf(fmaxf(dt_next, o->min_step), o->max_step);
  d->t0 = t0;
  d->t1 = accept ? t1 : t0;
  d->dt = dt_next;
  return ORC_OK;
}

/* AdaptiveRKSolver.step (:116-127) + interp_evaluate (utils/ode_utils.py:52-77) */
static int drv_step(drv_t *d, float next_t, float *out) {
  int64_t n_steps = 0;
  while (next_t > d->t1) {
    if (!(n_steps < (int64_t)d->o->max_num_steps)) return ORC_MAX_STEPS;
    int st = drv_adaptive_step(d);
    if (st) return st;
    n_steps++;
  }
  if (!(d->t0 <= next_t && next_t <= d->t1)) return ORC_INTERP_RANGE;
  float x = (next_t - d->t0) / (d->t1 - d->t0);
  for (int64_t e = 0; e < d->n; ++e) {
    float total = d->coef[0][e] + x * d->coef[1][e];
    float xp = x;
    for (int c = 2; c < 5; ++c) {
      xp = xp * x;
      total = total + xp * d->coef[c][e];
    }
    out[e] = total;
  }
  return ORC_OK;
}

/* AdaptiveSolver.integrate (solver/base_adaptive_solver.py:24-31); out [T][n].
 * A decreasing t_span is integrated in reversed time (repair R5). */
static int drv_integrate(drv_t *d, const float *y0, const float *t_span, int T, float *out) {
  if (T < 2) return ORC_BAD_ARG;
  d->rev = (t_span[1] < t_span[0]);
  for (int i = 1; i < T; ++i) {
    if (d->rev ? !(t_span[i] < t_span[i - 1]) : !(t_span[i] > t_span[i - 1])) return ORC_BAD_ARG;
  }
  memcpy(out, y0, sizeof(float) * d->n);
  float s0 = d->rev ? -t_span[0] : t_span[0];
  drv_before_integrate(d, s0, y0);
  for (int i = 1; i < T; ++i) {
    float s = d->rev ? -t_span[i] : t_span[i];
    int st = drv_step(d, s, out + (size_t)i * d->n);
    if (st) return st;
  }
  return ORC_OK;
}

/* _rms_norm (utils/ode_utils.py:8-9): squares in fp32, mean and sqrt in fp64, result fp32 */
static float rms_f64(const float *v, int64_t n) {
  double acc = 0.0;
  for (int64_t e = 0; e < n; ++e) {
    float q = v[e] * v[e];
    acc += (double)q;
  }
  return (float)sqrt(acc / (double)n);
}

static void stats_reset(orc_stats_t *s) {
  s->n_attempts = s->n_accepted = s->nfe = 0;
  s->status = ORC_OK;
  s->min_abs_ratio_m1 = INFINITY;
}

/* ------------------------------------------------------------------------------------------ */
/* Forward: odeint(MLP, Dopri5)                                                                 */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const orc_mlp_t *m;
  int64_t B; /* trajectories in this flat state */
} fwd_ctx_t;

static void fwd_rhs(void *ctx, float t, const float *y, float *dy) {
  (void)t; /* the field ignores t (example/ode_demo.py:32) */
  fwd_ctx_t *c = (fwd_ctx_t *)ctx;
  orc_mlp_eval_batch(c->m, y, c->B, dy);
}
static float fwd_norm(void *ctx, const float *v) {
  fwd_ctx_t *c = (fwd_ctx_t *)ctx;
  return rms_f64(v, c->B * c->m->d);
}

int orc_dopri5_mlp(const orc_mlp_t *m, const float *y0, int64_t B, const float *t_span, int32_t T,
                   const orc_opts_t *opts, int32_t controller, float *out, orc_stats_t *stats,
                   orc_attempt_t *log, int64_t log_cap, int64_t log_traj, int64_t *log_len,
                   int32_t nthreads) {
  return orc_adaptive_rk_mlp(ORC_RK_DOPRI5, m, y0, B, t_span, T, opts, controller, out, stats, log, log_cap,
                             log_traj, log_len, nthreads);
}

int orc_adaptive_rk_mlp(int32_t method, const orc_mlp_t *m, const float *y0, int64_t B, const float *t_span,
                        int32_t T, const orc_opts_t *opts, int32_t controller, float *out,
                        orc_stats_t *stats, orc_attempt_t *log, int64_t log_cap, int64_t log_traj,
                        int64_t *log_len, int32_t nthreads) {
  return orc_adaptive_rk_mlp_grid(method, m, y0, B, t_span, T, opts, controller, NULL, 0, NULL, 0, out, stats,
                                  log, log_cap, log_traj, log_len, nthreads);
}

int orc_adaptive_rk_mlp_grid(int32_t method, const orc_mlp_t *m, const float *y0, int64_t B,
                             const float *t_span, int32_t T, const orc_opts_t *opts, int32_t controller,
                             const float *step_t, int32_t n_step, const float *jump_t, int32_t n_jump,
                             float *out, orc_stats_t *stats, orc_attempt_t *log, int64_t log_cap,
                             int64_t log_traj, int64_t *log_len, int32_t nthreads) {
  const int D = m->d;
  rk_tab_t tab;
  if (D > ORC_MAX_D || m->h > ORC_MAX_H || T < 2 || B < 1 || rk_tab_init(&tab, method)) return ORC_BAD_ARG;
  if (log_len) *log_len = 0;
  if (controller == ORC_CTRL_BATCH) {
    drv_t d;
    if (drv_alloc(&d, B * D)) return ORC_BAD_ARG;
    d.tab = tab;
    d.step_t = step_t, d.n_step = n_step, d.jump_t = jump_t, d.n_jump = n_jump;
    fwd_ctx_t c = {m, B};
    orc_stats_t st;
    stats_reset(&st);
    d.rhs = fwd_rhs;
    d.norm = fwd_norm;
    d.ctx = &c;
    d.o = opts;
    d.st = &st;
    d.log = log;
    d.log_cap = log_cap;
    d.log_len = log_len;
    int rc = drv_integrate(&d, y0, t_span, T, out); /* out [T][B*D] is already time-major */
    st.status = rc;
    if (stats) stats[0] = st;
    drv_free(&d);
    return rc;
  }
  int worst = ORC_OK;
#ifdef _OPENMP
  if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
  {
    drv_t d;
    float *tmp = (float *)malloc(sizeof(float) * (size_t)T * D);
    int ok = (drv_alloc(&d, D) == 0) && tmp;
    if (ok) {
      d.tab = tab;
      d.step_t = step_t, d.n_step = n_step, d.jump_t = jump_t, d.n_jump = n_jump;
    }
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
    for (int64_t b = 0; b < B; ++b) {
      if (!ok) continue;
      fwd_ctx_t c = {m, 1};
      orc_stats_t st;
      stats_reset(&st);
      d.rhs = fwd_rhs;
      d.norm = fwd_norm;
      d.ctx = &c;
      d.o = opts;
      d.st = &st;
      int want_log = (log && b == log_traj);
      d.log = want_log ? log : NULL;
      d.log_cap = log_cap;
      d.log_len = want_log ? log_len : NULL;
      int rc = drv_integrate(&d, y0 + b * D, t_span, T, tmp);
      st.status = rc;
      for (int i = 0; i < T; ++i)
        memcpy(out + ((size_t)i * B + b) * D, tmp + (size_t)i * D, sizeof(float) * D);
      if (stats) stats[b] = st;
      if (rc) {
#ifdef _OPENMP
#pragma omp critical
#endif
        { if (rc > worst) worst = rc; }
      }
    }
    if (ok) drv_free(&d);
    free(tmp);
  }
  return worst;
}

/* ------------------------------------------------------------------------------------------ */
/* Fixed grid: Euler / RK4 (3/8 rule)                                                           */
/* ------------------------------------------------------------------------------------------ */
int orc_fixed_mlp(int32_t method, const orc_mlp_t *m, const float *y0, int64_t B,
                  const float *t_span, int32_t T, float *out, int32_t nthreads) {
  const int D = m->d;
  if (D > ORC_MAX_D || m->h > ORC_MAX_H || T < 1 || B < 1) return ORC_BAD_ARG;
  const float one_third = (float)(1.0 / 3.0);
#ifdef _OPENMP
  if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
  for (int64_t b = 0; b < B; ++b) {
    float y[ORC_MAX_D], yi[ORC_MAX_D], k1[ORC_MAX_D], k2[ORC_MAX_D], k3[ORC_MAX_D], k4[ORC_MAX_D];
    float *o = out + (size_t)b * T * D;
    memcpy(y, y0 + b * D, sizeof(float) * D);
    memcpy(o, y, sizeof(float) * D);
    for (int i = 1; i < T; ++i) {
      const float t0 = t_span[i - 1], t1 = t_span[i];
      const float dt = t1 - t0;
      if (method == ORC_FIXED_EULER) {
        /* fixed_solver/euler.py:7-11 ; BaseODE.fuse xde/base_ode.py:58 */
        orc_mlp_eval(m, y, k1, NULL);
        for (int e = 0; e < D; ++e) y[e] = k1[e] * dt + y[e];
      } else if (method == ORC_FIXED_MIDPOINT) {
        /* fixed_solver/midpoint.py:7-18: y_half = fuse(f(y0), dt/2, y0); y1 = fuse(f(y_half), dt, y0) */
        const float half_dt = 0.5f * dt;
        orc_mlp_eval(m, y, k1, NULL);
        for (int e = 0; e < D; ++e) yi[e] = k1[e] * half_dt + y[e];
        orc_mlp_eval(m, yi, k2, NULL);
        for (int e = 0; e < D; ++e) y[e] = k2[e] * dt + y[e];
      } else {
        /* rk4_alt_step_func solver/base_fixed_solver.py:166-197 */
        const float dt13 = dt * one_third;
        orc_mlp_eval(m, y, k1, NULL);
        for (int e = 0; e < D; ++e) yi[e] = k1[e] * dt13 + y[e];
        orc_mlp_eval(m, yi, k2, NULL);
        for (int e = 0; e < D; ++e) yi[e] = (k1[e] - k2[e] * one_third) * dt + y[e];
        orc_mlp_eval(m, yi, k3, NULL);
        for (int e = 0; e < D; ++e) yi[e] = ((k1[e] - k2[e]) + k3[e]) * dt + y[e];
        orc_mlp_eval(m, yi, k4, NULL);
        for (int e = 0; e < D; ++e) {
          float a = k1[e] * dt + y[e];
          float bb = k2[e] * dt + y[e];
          float c = k3[e] * dt + y[e];
          float dd = k4[e] * dt + y[e];
          y[e] = (((a + 3.0f * bb) + 3.0f * c) + dd) * 0.125f;
        }
      }
      /* linear_interp with t == t1 returns y1 (interpolation/functional/interp_fn.py:4-10) */
      memcpy(o + (size_t)i * D, y, sizeof(float) * D);
    }
  }
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* Adjoint: OdeintAdjointMethod.backward with flat augmented state (repairs R4-R6)              */
/* layout: [g_t | y (Bm*D) | a (Bm*D) | gW1 (D*H) | gb1 (H) | gW2 (H*D) | gb2 (D)]              */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const orc_mlp_t *m;
  int64_t Bm;
  int adj_norm;
} adj_ctx_t;

static int64_t adj_nparams(const orc_mlp_t *m) {
  return (int64_t)m->d * m->h + m->h + (int64_t)m->h * m->d + m->d;
}

/* ---- order-independent batch sums of the parameter-gradient dynamics (round 2) ----------------------------
 * The reference sums the per-trajectory outer products over the batch inside paddle.autograd.grad, in whatever order
 * the backend's reduction takes.  At the reference's default tolerances the error estimate of the g_theta part is of
 * the size of the rounding noise of that sum, so the accept/reject sequence of the mixed norm depends on the
 * summation ORDER.  To make it reproducible on any machine (CPU oracle, any GPU grid) the sum is specified as:
 *   1. blocks of 32 consecutive trajectories (b = 32m .. 32m+31, the last block short): a sequential fp32 fma chain
 *      from 0 in trajectory order, X = fma(c_b, v_b, X)
 *        gW1[k][j]: c = u_k, v = dz_j;   gb1[j]: c = 1, v = dz_j;   gW2[j][d]: c = cot_d, v = h_j;
 *      gb2[d]: no chain, the addends are the values cot_d themselves;
 *   2. the block values are added EXACTLY in a 128-bit two's-complement fixed-point accumulator with LSB 2^-59
 *      (an addend is truncated toward zero to that grid: exact for anything above 2^-35, at most 1.7e-18 absolute
 *      below; an addend that is not finite or >= 2^40 in magnitude makes the result NaN; 2^28 addends fit) -- integer addition is associative, so every
 *      summation tree gives the same bits;
 *   3. the total is rounded ONCE to fp32 (round to nearest even).
 * With one trajectory (the per-trajectory controller) this is the plain product, as before. */
typedef struct {
  unsigned __int128 v; /* two's complement */
  int bad;
} fx128_t;

static void fx_add_float(fx128_t *a, float x) {
  uint32_t bits;
  memcpy(&bits, &x, 4);
  const int E = (int)((bits >> 23) & 0xffu);
  const uint32_t M = bits & 0x7fffffu;
  if (E == 255) {
    a->bad = 1;
    return;
  }
  const uint32_t mant = E ? (M | 0x800000u) : M;
  const int shift = (E ? E - 150 : -149) + 59;
  unsigned __int128 mag;
  if (shift >= 0) {
    if (shift > 75) { /* |x| >= 2^40 */
      a->bad = 1;
      return;
    }
    mag = (unsigned __int128)mant << shift;
  } else {
    mag = (-shift >= 32) ? 0 : (unsigned __int128)(mant >> (-shift));
  }
  a->v += (bits >> 31) ? (unsigned __int128)0 - mag : mag;
}

static float fx_to_float(const fx128_t *a) {
  if (a->bad) return NAN;
  unsigned __int128 v = a->v;
  const int neg = (int)(v >> 127);
  unsigned __int128 mag = neg ? (unsigned __int128)0 - v : v;
  if (mag == 0) return 0.0f;
  int p = 127;
  while (!((mag >> p) & 1)) --p;
  float r;
  if (p <= 23) {
    r = ldexpf((float)(uint32_t)mag, -59);
  } else {
    const int sh = p - 23;
    uint32_t top = (uint32_t)(mag >> sh);
    const unsigned __int128 rem = mag & ((((unsigned __int128)1) << sh) - 1);
    const unsigned __int128 half = ((unsigned __int128)1) << (sh - 1);
    if (rem > half || (rem == half && (top & 1u))) top++;
    r = ldexpf((float)top, sh - 59);
  }
  return neg ? -r : r;
}

/* test hook: the exact sum of n fp32 addends, rounded once */
float orc_fx_sum(const float *x, int64_t n) {
  fx128_t a = {0, 0};
  for (int64_t i = 0; i < n; ++i) fx_add_float(&a, x[i]);
  return fx_to_float(&a);
}

/* paddle.autograd.grad(f(y), (y, *params), grad_outputs=c) for a BATCH of Bm trajectories: per-trajectory f and vjp_y,
 * parameter gradients summed over the batch by the order-independent specification above (one trajectory: the plain
 * products).  y, c, f, dy: [Bm, D]; gparams: [P] = (gW1, gb1, gW2, gb2), overwritten. */
void orc_mlp_vjp_batch(const orc_mlp_t *m, const float *y, const float *c, int64_t Bm, float *f, float *dy,
                       float *gparams) {
  const int D = m->d, H = m->h;
  const int64_t P = adj_nparams(m);
  float *gw1 = gparams;
  float *gb1 = gw1 + (size_t)D * H;
  float *gw2 = gb1 + H;
  float *gb2 = gw2 + (size_t)H * D;
  memset(gparams, 0, sizeof(float) * P);
  if (Bm == 1) { /* one trajectory: every sum has one term */
    orc_mlp_vjp(m, y, c, f, dy, gw1, gb1, gw2, gb2);
    return;
  }
  fx128_t *tot = (fx128_t *)calloc((size_t)P, sizeof(fx128_t));
  float *X = (float *)malloc(sizeof(float) * (size_t)P);
  float u[ORC_MAX_D], h[ORC_MAX_H], dz[ORC_MAX_H];
  for (int64_t b0 = 0; b0 < Bm; b0 += 32) {
    const int64_t b1 = (b0 + 32 < Bm) ? b0 + 32 : Bm;
    float *xw1 = X, *xb1 = X + (size_t)D * H, *xw2 = xb1 + H;
    for (int64_t i = 0; i < P; ++i) X[i] = 0.0f;
    for (int64_t b = b0; b < b1; ++b) {
      const float *yb = y + b * D, *cot = c + b * D;
      /* f, dy exactly as orc_mlp_vjp; the parameter terms enter the block chains */
      for (int k = 0; k < D; ++k) u[k] = pre_act(m->pre, yb[k]);
      orc_mlp_eval(m, yb, f + b * D, h);
      for (int j = 0; j < H; ++j) {
        float acc = cot[0] * m->w2[(size_t)j * D];
        for (int d = 1; d < D; ++d) acc = fmaf(cot[d], m->w2[(size_t)j * D + d], acc);
        float sd = fmaf(-h[j], h[j], 1.0f);
        dz[j] = acc * sd;
      }
      for (int k = 0; k < D; ++k)
        dy[b * D + k] = chain2_dot(dz, m->w1 + (size_t)k * H, 1, H) * pre_act_grad(m->pre, yb[k]);
      for (int k = 0; k < D; ++k)
        for (int j = 0; j < H; ++j) xw1[(size_t)k * H + j] = fmaf(u[k], dz[j], xw1[(size_t)k * H + j]);
      for (int j = 0; j < H; ++j) xb1[j] = fmaf(1.0f, dz[j], xb1[j]);
      for (int j = 0; j < H; ++j)
        for (int d = 0; d < D; ++d) xw2[(size_t)j * D + d] = fmaf(cot[d], h[j], xw2[(size_t)j * D + d]);
      for (int d = 0; d < D; ++d) fx_add_float(&tot[(size_t)D * H + H + (size_t)H * D + d], cot[d]);
    }
    for (int64_t i = 0; i < (int64_t)D * H + H + (int64_t)H * D; ++i) fx_add_float(&tot[i], X[i]);
  }
  for (int64_t i = 0; i < P; ++i) gparams[i] = fx_to_float(&tot[i]);
  free(X);
  free(tot);
}

/* augmented_dynamics (functional/odeint_adjoint.py:89-124) */
static void adj_rhs(void *ctx, float t, const float *Y, float *dY) {
  (void)t;
  adj_ctx_t *c = (adj_ctx_t *)ctx;
  const orc_mlp_t *m = c->m;
  const int D = m->d;
  const int64_t Bm = c->Bm;
  const float *y = Y + 1, *a = Y + 1 + Bm * D;
  dY[0] = 0.0f; /* vjp_t: the field ignores t -> allow_unused zero (:116-118) */
  float *cot = (float *)malloc(sizeof(float) * (size_t)Bm * D);
  for (int64_t i = 0; i < Bm * D; ++i) cot[i] = -a[i]; /* grad_outputs=-adj_y */
  orc_mlp_vjp_batch(m, y, cot, Bm, dY + 1, dY + 1 + Bm * D, dY + 1 + 2 * Bm * D);
  free(cot);
}

/* default_adjoint_norm / adjoint_seminorm (functional/odeint_adjoint.py:284-309) with
 * state_norm = _rms_norm and _mixed_norm (utils/ode_utils.py:16-19) */
static float adj_norm(void *ctx, const float *V) {
  adj_ctx_t *c = (adj_ctx_t *)ctx;
  const orc_mlp_t *m = c->m;
  const int D = m->d, H = m->h;
  const int64_t Bm = c->Bm;
  float best = fabsf(V[0]);
  float ny = rms_f64(V + 1, Bm * D);
  if (ny > best) best = ny;
  float na = rms_f64(V + 1 + Bm * D, Bm * D);
  if (na > best) best = na;
  if (c->adj_norm == ORC_ADJ_NORM_MIXED) {
    const float *p = V + 1 + 2 * Bm * D;
    const int64_t sz[4] = {(int64_t)D * H, H, (int64_t)H * D, D};
    float pm = 0.0f;
    for (int q = 0; q < 4; ++q) {
      float r = rms_f64(p, sz[q]);
      if (q == 0 || r > pm) pm = r;
      p += sz[q];
    }
    if (pm > best) best = pm;
  }
  return best;
}

static int adj_solve(const orc_mlp_t *m, const float *t_span, int T, const float *y_ans,
                     const float *grad_y, int64_t B, int64_t b0, int64_t Bm, const orc_opts_t *opts,
                     int adj_norm_kind, double *gsum, float *gout, float *adj_y0, double *gt_sum,
                     orc_stats_t *st, orc_attempt_t *log, int64_t log_cap, int64_t *log_len) {
  const int D = m->d;
  const int64_t P = adj_nparams(m);
  const int64_t n = 1 + 2 * Bm * D + P;
  drv_t d;
  if (drv_alloc(&d, n)) return ORC_BAD_ARG;
  adj_ctx_t c = {m, Bm, adj_norm_kind};
  float *aug = (float *)calloc((size_t)n * 3, sizeof(float));
  float *sol = aug + n;
  int rc = ORC_OK;
  /* aug_state = [0, y_ans[-1], grad_y[-1], zeros...] (:75-82) */
  for (int64_t b = 0; b < Bm; ++b)
    for (int e = 0; e < D; ++e) {
      aug[1 + b * D + e] = y_ans[((size_t)(T - 1) * B + b0 + b) * D + e];
      aug[1 + (Bm + b) * D + e] = grad_y[((size_t)(T - 1) * B + b0 + b) * D + e];
    }
  for (int i = T - 1; i >= 1; --i) {
    if (gt_sum) {
      /* t_requires_grad (:135-141): dLd_cur_t = func(t_i, y_i) . grad_y[i]; aug_state[0] -= dLd_cur_t;
       * grad_t_span[i] = dLd_cur_t.  Dot product: products rounded, summed left to right over (b, e). */
      float dl = 0.0f;
      for (int64_t b = 0; b < Bm; ++b) {
        float fe[ORC_MAX_D];
        orc_mlp_eval(m, aug + 1 + b * D, fe, NULL);
        for (int e = 0; e < D; ++e) {
          const float pr = fe[e] * grad_y[((size_t)i * B + b0 + b) * D + e];
          dl = (b == 0 && e == 0) ? pr : dl + pr;
        }
      }
      aug[0] = aug[0] - dl;
      gt_sum[i] += (double)dl;
    }
    float seg[2] = {t_span[i], t_span[i - 1]}; /* t_span[i-1:i+1].flip(0) (:147) */
    d.rhs = adj_rhs;
    d.norm = adj_norm;
    d.ctx = &c;
    d.o = opts;
    d.st = st;
    d.log = log;
    d.log_cap = log_cap;
    d.log_len = log_len;
    rc = drv_integrate(&d, aug, seg, 2, sol);
    if (rc) break;
    memcpy(aug, sol + n, sizeof(float) * n); /* a[1] for a in aug_state (:153) */
    for (int64_t b = 0; b < Bm; ++b)
      for (int e = 0; e < D; ++e) {
        size_t src = ((size_t)(i - 1) * B + b0 + b) * D + e;
        aug[1 + b * D + e] = y_ans[src];              /* :154-156 */
        aug[1 + (Bm + b) * D + e] += grad_y[src];     /* :157-159 */
      }
  }
  if (rc == ORC_OK) {
    const float *g = aug + 1 + 2 * Bm * D;
    if (gsum)
      for (int64_t p = 0; p < P; ++p) gsum[p] += (double)g[p];
    if (gout) memcpy(gout, g, sizeof(float) * P);
    if (adj_y0)
      for (int64_t b = 0; b < Bm; ++b)
        memcpy(adj_y0 + (b0 + b) * D, aug + 1 + (Bm + b) * D, sizeof(float) * D);
    if (gt_sum) gt_sum[0] += (double)aug[0]; /* grad_t_span[0] = aug_state[0] (:161-162) */
  }
  free(aug);
  drv_free(&d);
  return rc;
}

int orc_dopri5_mlp_adjoint(const orc_mlp_t *m, const float *t_span, int32_t T, const float *y_ans,
                           const float *grad_y, int64_t B, const orc_opts_t *opts,
                           int32_t controller, int32_t adj_norm_kind, float *out_gparams,
                           float *out_adj_y0, float *out_grad_t, orc_stats_t *stats, orc_attempt_t *log,
                           int64_t log_cap, int64_t log_traj, int64_t *log_len, int32_t nthreads) {
  const int64_t P = adj_nparams(m);
  if (m->d > ORC_MAX_D || m->h > ORC_MAX_H || T < 2 || B < 1) return ORC_BAD_ARG;
  if (log_len) *log_len = 0;
  if (controller == ORC_CTRL_BATCH) {
    orc_stats_t st;
    stats_reset(&st);
    double *gtb = out_grad_t ? (double *)calloc((size_t)T, sizeof(double)) : NULL;
    int rc = adj_solve(m, t_span, T, y_ans, grad_y, B, 0, B, opts, adj_norm_kind, NULL, out_gparams,
                       out_adj_y0, gtb, &st, log, log_cap, log_len);
    if (gtb) {
      for (int i = 0; i < T; ++i) out_grad_t[i] = (float)gtb[i];
      free(gtb);
    }
    st.status = rc;
    if (stats) stats[0] = st;
    return rc;
  }
  /* trajectory mode: reference run with B = 1 per trajectory; per-trajectory parameter-gradient
   * integrals are summed at the end (valid: g_theta never feeds back). The sum is
   * taken in fp64 so it does not depend on the trajectory order. */
  double *gtot = (double *)calloc((size_t)P, sizeof(double));
  double *gttot = out_grad_t ? (double *)calloc((size_t)T, sizeof(double)) : NULL;
  int worst = ORC_OK;
#ifdef _OPENMP
  if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
  {
    double *gloc = (double *)calloc((size_t)P, sizeof(double));
    double *gtloc = out_grad_t ? (double *)calloc((size_t)T, sizeof(double)) : NULL;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
    for (int64_t b = 0; b < B; ++b) {
      orc_stats_t st;
      stats_reset(&st);
      int want_log = (log && b == log_traj);
      int rc = adj_solve(m, t_span, T, y_ans, grad_y, B, b, 1, opts, adj_norm_kind, gloc, NULL,
                         out_adj_y0, gtloc, &st, want_log ? log : NULL, log_cap, want_log ? log_len : NULL);
      st.status = rc;
      if (stats) stats[b] = st;
      if (rc) {
#ifdef _OPENMP
#pragma omp critical
#endif
        { if (rc > worst) worst = rc; }
      }
    }
#ifdef _OPENMP
#pragma omp critical
#endif
    {
      for (int64_t p = 0; p < P; ++p) gtot[p] += gloc[p];
      if (gttot)
        for (int i = 0; i < T; ++i) gttot[i] += gtloc[i];
    }
    free(gloc);
    free(gtloc);
  }
  for (int64_t p = 0; p < P; ++p) out_gparams[p] = (float)gtot[p];
  free(gtot);
  if (gttot) {
    for (int i = 0; i < T; ++i) out_grad_t[i] = (float)gttot[i];
    free(gttot);
  }
  return worst;
}

/* ------------------------------------------------------------------------------------------ */
/* SDE: Euler-Maruyama (intended BaseSDE semantics, repairs R2/R3) and Milstein (extension)     */
/* ------------------------------------------------------------------------------------------ */
static void mlp_eval_diag_jac(const orc_mlp_t *m, const float *y, float *g, float *gp) {
  /* g = field(y); gp[d] = d g_d / d y_d  (diagonal of the Jacobian, analytic) */
  const int D = m->d, H = m->h;
  float h[ORC_MAX_H];
  orc_mlp_eval(m, y, g, h);
  float sw[ORC_MAX_H];
  for (int d = 0; d < D; ++d) {
    for (int j = 0; j < H; ++j) {
      float s = fmaf(-h[j], h[j], 1.0f);
      sw[j] = s * m->w1[(size_t)d * H + j];
    }
    gp[d] = chain2_dot(sw, m->w2 + d, D, H) * pre_act_grad(m->pre, y[d]);
  }
}

int orc_sde_mlp(int32_t scheme, const orc_mlp_t *drift, const orc_mlp_t *diffusion, const float *y0,
                int64_t B, const float *t_span, int32_t T, const float *dW, float *out,
                int32_t nthreads) {
  const int D = drift->d;
  if (diffusion->d != D || D > ORC_MAX_D || drift->h > ORC_MAX_H || diffusion->h > ORC_MAX_H)
    return ORC_BAD_ARG;
#ifdef _OPENMP
  if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
  for (int64_t b = 0; b < B; ++b) {
    float y[ORC_MAX_D], f[ORC_MAX_D], g[ORC_MAX_D], gp[ORC_MAX_D];
    float *o = out + (size_t)b * T * D;
    memcpy(y, y0 + b * D, sizeof(float) * D);
    memcpy(o, y, sizeof(float) * D);
    for (int i = 1; i < T; ++i) {
      const float dt = t_span[i] - t_span[i - 1];
      const float *w = dW + ((size_t)(i - 1) * B + b) * D;
      orc_mlp_eval(drift, y, f, NULL);
      if (scheme == ORC_SDE_MILSTEIN)
        mlp_eval_diag_jac(diffusion, y, g, gp);
      else
        orc_mlp_eval(diffusion, y, g, NULL);
      for (int e = 0; e < D; ++e) {
        /* y1 = y0 + f*dt + g*dW  (xde/base_sde.py:56-61 intent) */
        float v = (y[e] + f[e] * dt) + g[e] * w[e];
        if (scheme == ORC_SDE_MILSTEIN) v = v + ((0.5f * g[e]) * gp[e]) * (w[e] * w[e] - dt);
        y[e] = v;
      }
      memcpy(o + (size_t)i * D, y, sizeof(float) * D);
    }
  }
  return ORC_OK;
}

/* sdeint_adjoint backward.  The reference's SdeintAdjointMethod.backward (functional/sdeint_adjoint.py:57-230)
 * copies the ODE adjoint and its `augmented_diffusion` is a verbatim copy of the drift dynamics (:136-171), on
 * top of the uninstantiable BaseSDE: there is no behaviour to restate.  What is defined
 * here is what the code is reaching for on the solver's fixed grid: the EXACT adjoint of the Euler-Maruyama
 * recursion y[n+1] = (y[n] + f(y[n]) dt_n) + g(y[n]) * dW_n  (discretise-then-differentiate):
 *     lam[n] = lam[n+1] + J_f(y[n])^T (lam[n+1] dt_n) + J_g(y[n])^T (lam[n+1] * dW_n) + grad_y[n]
 *     g_theta_f += (df/dtheta)(y[n])^T (lam[n+1] dt_n),  g_theta_g += (dg/dtheta)(y[n])^T (lam[n+1] * dW_n)
 * with y[n] read from the stored forward solution.  PARITY UNPINNED; checked against fp64 autograd in tests.
 * y_all, grad_y: [B,T,D] (the fixed solver's layout); dW [T-1,B,D]; out_gf / out_gg: parameter gradients of
 * drift / diffusion (gW1,gb1,gW2,gb2), summed over trajectories in fp64; out_adj_y0 [B,D] optional. */
int orc_sde_mlp_adjoint(const orc_mlp_t *drift, const orc_mlp_t *diffusion, const float *t_span, int32_t T,
                        const float *y_all, const float *grad_y, int64_t B, const float *dW, float *out_gf,
                        float *out_gg, float *out_adj_y0, int32_t nthreads) {
  const int D = drift->d;
  if (diffusion->d != D || D > ORC_MAX_D || drift->h > ORC_MAX_H || diffusion->h > ORC_MAX_H || T < 1)
    return ORC_BAD_ARG;
  const int64_t Pf = adj_nparams(drift), Pg = adj_nparams(diffusion);
  double *accf = (double *)calloc((size_t)(Pf + Pg), sizeof(double));
  if (!accf) return ORC_BAD_ARG;
  double *accg = accf + Pf;
#ifdef _OPENMP
  if (nthreads < 1) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
  {
    float *gf = (float *)calloc((size_t)(Pf + Pg), sizeof(float));
    float *gg = gf + Pf;
    double *lf = (double *)calloc((size_t)(Pf + Pg), sizeof(double));
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t b = 0; b < B; ++b) {
      float lam[ORC_MAX_D], cf[ORC_MAX_D], cg[ORC_MAX_D], fo[ORC_MAX_D], dyf[ORC_MAX_D], dyg[ORC_MAX_D];
      const float *yb = y_all + (size_t)b * T * D, *gb = grad_y + (size_t)b * T * D;
      memset(gf, 0, sizeof(float) * (size_t)(Pf + Pg)); /* per-trajectory partials in fp32, batch sum in fp64 */
      for (int e = 0; e < D; ++e) lam[e] = gb[(size_t)(T - 1) * D + e];
      for (int n = T - 2; n >= 0; --n) {
        const float dt = t_span[n + 1] - t_span[n];
        const float *w = dW + ((size_t)n * B + b) * D;
        const float *y = yb + (size_t)n * D;
        for (int e = 0; e < D; ++e) {
          cf[e] = lam[e] * dt;
          cg[e] = lam[e] * w[e];
        }
        orc_mlp_vjp(drift, y, cf, fo, dyf, gf, gf + (size_t)D * drift->h, gf + (size_t)D * drift->h + drift->h,
                    gf + (size_t)2 * D * drift->h + drift->h);
        orc_mlp_vjp(diffusion, y, cg, fo, dyg, gg, gg + (size_t)D * diffusion->h,
                    gg + (size_t)D * diffusion->h + diffusion->h, gg + (size_t)2 * D * diffusion->h + diffusion->h);
        for (int e = 0; e < D; ++e) lam[e] = ((lam[e] + dyf[e]) + dyg[e]) + gb[(size_t)n * D + e];
      }
      if (out_adj_y0)
        for (int e = 0; e < D; ++e) out_adj_y0[b * D + e] = lam[e];
      for (int64_t q = 0; q < Pf + Pg; ++q) lf[q] += (double)gf[q];
    }
#ifdef _OPENMP
#pragma omp critical
#endif
    {
      for (int64_t q = 0; q < Pf + Pg; ++q) accf[q] += lf[q];
    }
    free(gf);
    free(lf);
  }
  for (int64_t q = 0; q < Pf; ++q) out_gf[q] = (float)accf[q];
  for (int64_t q = 0; q < Pg; ++q) out_gg[q] = (float)accg[q];
  free(accf);
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* History gather                                                                               */
/* ------------------------------------------------------------------------------------------ */
int orc_history_gather(int32_t kind, const float *his, int64_t R, int32_t Th, int32_t D,
                       const float *span, const float *lags, int32_t L, float *out_val,
                       float *out_der) {
  if (Th < 2 || (kind == ORC_INTERP_BEZIER && Th < 4)) return ORC_BAD_ARG;
  for (int l = 0; l < L; ++l) {
    const float t = lags[l];
    /* paddle.bucketize(t, _t) (right=False) = #{_t < t}; index = clip(. - 1, 0, maxlen)
     * (interpolation/interpolate_base.py:62-66) */
    int cnt = 0;
    while (cnt < Th && span[cnt] < t) ++cnt;
    int idx = cnt - 1;
    if (idx < 0) idx = 0;
    if (idx > Th - 1) idx = Th - 1;
    /* _make_series scale1/scale2 (interpolation/interpolate.py:52-54,147-149) */
#define SCALE1(i) (((i) < Th - 1) ? (span[(i) + 1] - span[(i)]) : (span[Th - 1] - span[Th - 2]))
#define SCALE2(i) (((i) == 0) ? (span[1] - span[0]) : SCALE1((i) - 1))
    if (kind == ORC_INTERP_BEZIER) {
      /* BezierSpline (interpolation/interpolate.py:207-298): control points p_i .. p_{i+3} (clamped to the
       * last sample, :258-261), each divided by its own shifted 3-interval span scale_m[i] =
       * scale1[max(i-(m-1),0)], scale1[i] = t[min(i,Th-4)+3] - t[min(i,Th-4)] (:252-256); Bernstein matrix
       * :240-245; value rescaled by scale1[i], derivative not (interpolate_base.py:89-114). */
#define BSC1(i) (span[((i) < Th - 4 ? (i) : Th - 4) + 3] - span[((i) < Th - 4 ? (i) : Th - 4)])
#define BSCM(i, m) BSC1(((i) - (m)) > 0 ? ((i) - (m)) : 0)
      const float b1 = BSC1(idx);
      float sb = t - span[idx];
      sb = sb / b1;
      const float s2 = sb * sb, s3 = s2 * sb;
      const float tv[4] = {s3, s2, sb, 1.0f};
      const float td[4] = {3.0f * s2, 2.0f * sb, 1.0f, 0.0f};
      static const float Bm[4][4] = {
          {-1.0f, 3.0f, -3.0f, 1.0f}, {3.0f, -6.0f, 3.0f, 0}, {-3.0f, 3.0f, 0, 0}, {1.0f, 0, 0, 0}};
      float cv[4], cd[4], scm[4];
      int im[4];
      for (int c = 0; c < 4; ++c) {
        cv[c] = ((tv[0] * Bm[0][c] + tv[1] * Bm[1][c]) + tv[2] * Bm[2][c]) + tv[3] * Bm[3][c];
        cd[c] = ((td[0] * Bm[0][c] + td[1] * Bm[1][c]) + td[2] * Bm[2][c]) + td[3] * Bm[3][c];
        scm[c] = BSCM(idx, c);
        im[c] = (idx + c < Th) ? idx + c : Th - 1;
      }
      for (int64_t r = 0; r < R; ++r) {
        const float *base = his + (size_t)r * Th * D;
        for (int e = 0; e < D; ++e) {
          float a[4];
          for (int c = 0; c < 4; ++c) a[c] = base[(size_t)im[c] * D + e] / scm[c];
          size_t o = ((size_t)r * L + l) * D + e;
          out_val[o] = (((cv[0] * a[0] + cv[1] * a[1]) + cv[2] * a[2]) + cv[3] * a[3]) * b1;
          out_der[o] = ((cd[0] * a[0] + cd[1] * a[1]) + cd[2] * a[2]) + cd[3] * a[3];
        }
      }
      continue;
    }
    const float sc1 = SCALE1(idx), sc2 = SCALE2(idx);
    float s = t - span[idx];
    s = s / sc1;
    const int i1 = (idx + 1 < Th) ? idx + 1 : Th - 1;
    if (kind == ORC_INTERP_LINEAR) {
      /* ts=[s,1] / [1,0]; H=[[-1,1],[1,0]] (interpolate.py:35-38,71-78) */
      const float cv0 = s * -1.0f + 1.0f * 1.0f, cv1 = s * 1.0f + 1.0f * 0.0f;
      const float cd0 = 1.0f * -1.0f + 0.0f * 1.0f, cd1 = 1.0f * 1.0f + 0.0f * 0.0f;
      for (int64_t r = 0; r < R; ++r) {
        const float *p0 = his + ((size_t)r * Th + idx) * D;
        const float *p1 = his + ((size_t)r * Th + i1) * D;
        for (int e = 0; e < D; ++e) {
          float a0 = p0[e] / sc1, a1 = p1[e] / sc2;
          size_t o = ((size_t)r * L + l) * D + e;
          out_val[o] = (cv0 * a0 + cv1 * a1) * sc1;
          out_der[o] = cd0 * a0 + cd1 * a1;
        }
      }
    } else {
      /* ts=[s^3,s^2,s,1] / [3s^2,2s,1,0]; H rows (interpolate.py:127-130,184-191) */
      const float s2 = s * s, s3 = s2 * s;
      const float tv[4] = {s3, s2, s, 1.0f};
      const float td[4] = {3.0f * s2, 2.0f * s, 1.0f, 0.0f};
      static const float Hm[4][4] = {
          {2.0f, -2.0f, 1.0f, 1.0f}, {-3.0f, 3.0f, -2.0f, -1.0f}, {0, 0, 1.0f, 0}, {1.0f, 0, 0, 0}};
      float cv[4], cd[4];
      for (int c = 0; c < 4; ++c) {
        cv[c] = ((tv[0] * Hm[0][c] + tv[1] * Hm[1][c]) + tv[2] * Hm[2][c]) + tv[3] * Hm[3][c];
        cd[c] = ((td[0] * Hm[0][c] + td[1] * Hm[1][c]) + td[2] * Hm[2][c]) + td[3] * Hm[3][c];
      }
      /* _make_derivative (interpolate.py:160-182): derivs[i], i in [0,Th] */
#define DIFFT(i) SCALE1(i)
      const int ia = (idx < Th - 1) ? idx : Th - 2;         /* derivs[idx]   */
      const int ib_raw = idx + 1;                           /* derivs[idx+1] */
      const int ib = (ib_raw < Th - 1) ? ib_raw : Th - 2;
      const float dta = DIFFT(idx < Th ? idx : Th - 1);
      const float dtb = DIFFT(ib_raw < Th ? ib_raw : Th - 1);
      for (int64_t r = 0; r < R; ++r) {
        const float *base = his + (size_t)r * Th * D;
        for (int e = 0; e < D; ++e) {
          float a0 = base[(size_t)idx * D + e] / sc1, a1 = base[(size_t)i1 * D + e] / sc2;
          float m0 = (base[(size_t)(ia + 1) * D + e] - base[(size_t)ia * D + e]) / dta;
          float m1 = (base[(size_t)(ib + 1) * D + e] - base[(size_t)ib * D + e]) / dtb;
          size_t o = ((size_t)r * L + l) * D + e;
          out_val[o] = (((cv[0] * a0 + cv[1] * a1) + cv[2] * m0) + cv[3] * m1) * sc1;
          out_der[o] = ((cd[0] * a0 + cd[1] * a1) + cd[2] * m0) + cd[3] * m1;
        }
      }
    }
  }
  return ORC_OK;
}

void orc_history_gather_bwd(const float *grad_y, const float *deriv, int64_t R, int32_t L, int32_t D,
                            float *g_lags) {
  /* paddle.sum(grad_y * deriv, axis=[0,1,3]) (xde/base_dde.py:125-126); fp64 accumulation */
  for (int l = 0; l < L; ++l) {
    double acc = 0.0;
    for (int64_t r = 0; r < R; ++r)
      for (int e = 0; e < D; ++e) {
        size_t o = ((size_t)r * L + l) * D + e;
        float p = grad_y[o] * deriv[o];
        acc += (double)p;
      }
    g_lags[l] = (float)acc;
  }
}

void orc_dde_fuse(const float *dy, float dt, const float *y0, int64_t n, float *y1) {
  for (int64_t e = 0; e < n; ++e) {
    float y = dy[e] * dt + y0[e];
    y1[e] = (dy[e] - 0.001f * y) * dt + y0[e];
  }
}
