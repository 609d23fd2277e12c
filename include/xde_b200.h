/*
 * xde_b200.h -- C ABI of libxde_b200.so: the B200 (sm_100a) implementation of PaddleXDE's batched
 * differential-equation integration hot path.
 *
 * The reference (DrownFish19/PaddleXDE) is pure Python on Paddle eager ops and has NO FFI / custom-op
 * boundary of its own; its de-facto extension point is the solver-class protocol
 *     s = solver(xde=xde, y0=xde.y0, rtol=rtol, atol=atol, **options); s.integrate(t_span)
 * (paddlexde/functional/odeint.py:30-31) plus the four functional entry points.  Each entry point
 * below replaces the inner loop that one of those Python call sites runs; the file:line it
 * replaces is cited on each declaration (paths relative to the reference root).  The Python shim
 * (paddlexde_b200/) and the reference-side stub shown in INTEGRATION.md bind these with ctypes.
 *
 * Conventions
 *   - plain C, no torch / DLPack types: raw device pointers + sizes.  All tensor arguments are
 *     DEVICE pointers to contiguous fp32 unless a name ends in _host.  Inputs are borrowed and
 *     never written; outputs are caller-allocated.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *     asynchronous with respect to the host; no call synchronises the device.
 *   - return value: XDE_OK or a negative XDE_E_* launch/argument error.  Solver conditions that the
 *     reference raises as Python asserts (dt underflow, non-finite state, max_num_steps,
 *     base_adaptive_solver_rk.py:120-122,200-203) are reported through xde_stats_t.status in device
 *     memory, to be read by the caller after it synchronises the stream.
 *   - re-entrant: no global mutable state except a monotonically increasing launch counter.
 */
#ifndef XDE_B200_H
#define XDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XDE_ABI_VERSION 3 /* 2: + tensor-core, table-driven adaptive RK, forced grid points, Philox, Bezier
                             3: + out_grad_t of the adjoint entry */

/* return codes */
enum {
  XDE_OK = 0,
  XDE_E_BAD_ARG = -1,           /* ValueError in the shim */
  XDE_E_UNSUPPORTED_FIELD = -2, /* field family / shape not covered by a fused kernel: hard error, no fallback */
  XDE_E_CUDA = -3               /* a CUDA runtime call failed; see xde_last_error() */
};

/* device-side solver status (xde_stats_t.status): mirrors the reference's asserts */
enum {
  XDE_ST_OK = 0,
  XDE_ST_DT_UNDERFLOW = 1,    /* assert t0 + dt > t0          base_adaptive_solver_rk.py:200 */
  XDE_ST_NONFINITE_STATE = 2, /* assert isfinite(y0).all()    base_adaptive_solver_rk.py:201-203 */
  XDE_ST_MAX_STEPS = 3,       /* max_num_steps exceeded       base_adaptive_solver_rk.py:120-122 */
  XDE_ST_INTERP_RANGE = 5,    /* invalid interpolation        utils/ode_utils.py:65-67 */
  XDE_ST_TC_RANGE = 6         /* tensor-core entries only: a stage input pre(y) was non-finite or >= 65504 in
                                 magnitude (not representable in the fp16 operand split): discard the result and
                                 use the FP32 entry (no reference counterpart) */
};

/* y ** p applied before the first Linear (example/ode_demo.py:33, example/sde_demo.py:183) */
enum { XDE_PRE_ID = 0, XDE_PRE_SQUARE = 1, XDE_PRE_CUBE = 2 };
/* error-controller granularity: TRAJECTORY = one controller per trajectory (north star; equals the
 * reference run with B = 1 per trajectory); BATCH = the reference's single global RMS norm and dt
 * (utils/ode_utils.py:8-9,80-82). */
enum { XDE_CTRL_TRAJECTORY = 0, XDE_CTRL_BATCH = 1 };
/* adjoint error norm (functional/odeint_adjoint.py:284-309) */
enum { XDE_ADJ_NORM_MIXED = 0, XDE_ADJ_NORM_SEMI = 1 };
enum { XDE_FIXED_EULER = 0, XDE_FIXED_RK4_38 = 1, XDE_FIXED_MIDPOINT = 2 };
/* embedded Runge-Kutta pairs exported by solver/__init__.py:1-6.  XDE_RK_DOPRI5_TABLE (diagnostics) runs
 * the table-driven kernel with the Dormand-Prince tableau: it must equal XDE_RK_DOPRI5 bit for bit. */
enum {
  XDE_RK_DOPRI5 = 0,
  XDE_RK_BOSH3 = 1,
  XDE_RK_FEHLBERG2 = 2,
  XDE_RK_ADAPTIVE_HEUN = 3,
  XDE_RK_DOPRI8 = 4,
  XDE_RK_DOPRI5_TABLE = 100
};
enum { XDE_SDE_EM = 0, XDE_SDE_MILSTEIN = 1 };
enum { XDE_MATH_FP32 = 0, XDE_MATH_TENSOR = 1 };
enum { XDE_INTERP_LINEAR = 0, XDE_INTERP_HERMITE = 1, XDE_INTERP_BEZIER = 2 };

/* The fused vector-field family: f(t, y) = tanh(pre(y) @ w1 + b1) @ w2 + b2
 * (example/ode_demo.py:17-33).  Weights in Paddle nn.Linear layout [in, out]. */
typedef struct {
  int32_t d;       /* state dim D */
  int32_t h;       /* hidden width H */
  int32_t pre;     /* XDE_PRE_* */
  int32_t _pad;
  const float *w1; /* [d, h] */
  const float *b1; /* [h]    */
  const float *w2; /* [h, d] */
  const float *b2; /* [d]    */
} xde_mlp_field_t;

/* keyword arguments of AdaptiveRKSolver.__init__ (solver/base_adaptive_solver_rk.py:32-49) */
typedef struct {
  float rtol, atol;
  float min_step, max_step;
  float first_step; /* NaN => select_initial_step (solver/base_adaptive_solver.py:33-72) */
  float safety, ifactor, dfactor;
  int32_t max_num_steps;
  int32_t _pad;
} xde_ctrl_opts_t;

/* written by the kernels (device memory, zeroed by the entry point before the launch) */
typedef struct {
  unsigned long long n_attempts; /* step attempts summed over trajectories (= trajectory-steps) */
  unsigned long long n_accepted;
  unsigned long long nfe;        /* vector-field evaluations, counted as the reference would issue them */
  int32_t status;                /* worst XDE_ST_* over trajectories */
  int32_t _pad;
} xde_stats_t;

/* optional per-attempt log, one row of `cap` records per trajectory (tests / diagnostics) */
typedef struct {
  float t0, dt, ratio;
  int32_t accepted;
} xde_attempt_t;

typedef struct {
  xde_attempt_t *records; /* [B, cap] device, or NULL */
  int32_t *counts;        /* [B] device attempts recorded per trajectory (counts beyond cap are counted, not stored) */
  int32_t cap;
  int32_t _pad;
} xde_attempt_log_t;

int xde_abi_version(void);
/* thread-local text of the last XDE_E_CUDA / argument failure */
const char *xde_last_error(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
unsigned long long xde_launch_count(void);
void xde_default_ctrl_opts(xde_ctrl_opts_t *o);
/* measurement aid (bench.py): saturates the FP32 FMA pipe for `iters` iterations of 8 independent
 * chains per thread on every SM; *n_flops_host receives the FLOPs the launch performs.  sink: one
 * device float (never written in practice). */
int xde_probe_ffma_f32(int32_t iters, float *sink, int64_t *n_flops_host, void *stream);
/* the same with the packed FFMA2 instruction (fma.rn.f32x2, sm_100+): 2 FMAs per lane and issue slot */
int xde_probe_ffma2_f32(int32_t iters, float *sink, int64_t *n_flops_host, void *stream);

/* odeint(func, y0, t_span, solver=Dopri5)            functional/odeint.py:28-35
 *   -> AdaptiveSolver.integrate                       solver/base_adaptive_solver.py:24-31
 *   -> AdaptiveRKSolver._adaptive_step/_runge_kutta_step  solver/base_adaptive_solver_rk.py:116-292
 * One launch integrates every trajectory over the whole t_span (device-resident controller).
 * y0 [B,D]; t_span [T] strictly increasing or strictly decreasing; out [T,B,D] time-major
 * (base_adaptive_solver.py:25).  stats / log may be NULL.
 * Fused shapes: D in 1..8 with any H that fits shared memory (both controllers); D = 64 with H in
 * {128,256}, D = 32 with H in {64,128,256}, D = 16 with H = 64 (XDE_CTRL_TRAJECTORY); anything else returns
 * XDE_E_UNSUPPORTED_FIELD. */
int xde_dopri5_mlp_f32(const xde_mlp_field_t *field, const float *y0, int64_t B, const float *t_span,
                       int32_t T, const xde_ctrl_opts_t *opts, int32_t controller, float *out,
                       xde_stats_t *stats, const xde_attempt_log_t *log, void *stream);

/* odeint(func, y0, t_span, solver=Bosh3|Fehlberg2|AdaptiveHeun|Dopri8|Dopri5): the same driver with the
 * tableau as data (stage count, FSAL shortcut base_adaptive_solver_rk.py:172-176, controller order):
 *   adaptive_solver/bosh3.py:5-27, fehlberg2.py:5-22, adaptive_heun.py:5-27, dopri8.py:5-252.
 * method = XDE_RK_*; XDE_RK_DOPRI5 forwards to xde_dopri5_mlp_f32.  The other pairs run the per-trajectory
 * controller (controller = XDE_CTRL_BATCH returns XDE_E_UNSUPPORTED_FIELD).  Same layouts as above.
 * Fused shapes: D in 1..8 (every pair); the large-state shapes of xde_dopri5_mlp_f32 for the pairs with at most
 * 6 stages (Dopri5, Bosh3, Fehlberg2, AdaptiveHeun; Dopri8 returns XDE_E_UNSUPPORTED_FIELD there). */
int xde_adaptive_rk_mlp_f32(int32_t method, const xde_mlp_field_t *field, const float *y0, int64_t B,
                            const float *t_span, int32_t T, const xde_ctrl_opts_t *opts, int32_t controller,
                            float *out, xde_stats_t *stats, const xde_attempt_log_t *log, void *stream);

/* ... with the solver's step_t / jump_t keyword arguments (solver/base_adaptive_solver_rk.py:38-46 ctor,
 * :94-114 setup, :209-224 clamping of the attempt to the next forced point, :263-273 index advance and the
 * re-evaluation of f after a jump).  step_t / jump_t: device arrays in t_span's time, sorted in integration
 * order and filtered to lie at or after t_span[0] (sort_tvals, utils/ode_utils.py:22-25, is the caller's);
 * null / 0 = none.  Any method, per-trajectory controller; Dopri5 with forced points runs on the
 * table-driven kernel (bit-identical arithmetic).  Same fused shapes as xde_adaptive_rk_mlp_f32. */
int xde_adaptive_rk_mlp_grid_f32(int32_t method, const xde_mlp_field_t *field, const float *y0, int64_t B,
                                 const float *t_span, int32_t T, const xde_ctrl_opts_t *opts, int32_t controller,
                                 const float *step_t, int32_t n_step, const float *jump_t, int32_t n_jump,
                                 float *out, xde_stats_t *stats, const xde_attempt_log_t *log, void *stream);

/* OdeintAdjointMethod.backward                        functional/odeint_adjoint.py:47-167
 * (augmented_dynamics :89-124 integrated backwards segment by segment :134-159).
 * y_ans, grad_y [T,B,D]; out_gparams [d*h + h + h*d + d] = (gW1, gb1, gW2, gb2) summed over the
 * B trajectories handed to this call (all-reduce across GPUs is the caller's, 8(e));
 * out_adj_y0 [B,D] optional (dL/dy0: computed and discarded by the reference, :167);
 * out_grad_t [T] optional (NULL = t_span does not require a gradient): grad_t_span of :129-141,161-162,
 * i.e. out_grad_t[i] = sum_b func(t_i, y_i) . grad_y[i] for i >= 1 and out_grad_t[0] = the final aug_state[0]
 * (XDE_CTRL_TRAJECTORY only; the g_t slot then takes part in the controller norm as the reference's does).
 * Fused shapes: D in 1..8 with H <= 64 (H <= 128 for D <= 2). */
int xde_dopri5_mlp_adjoint_f32(const xde_mlp_field_t *field, const float *t_span, int32_t T,
                               const float *y_ans, const float *grad_y, int64_t B,
                               const xde_ctrl_opts_t *opts, int32_t controller, int32_t adj_norm,
                               float *out_gparams, float *out_adj_y0, float *out_grad_t, xde_stats_t *stats,
                               const xde_attempt_log_t *log, void *stream);

/* odeint(func, y0, t_span, solver=Euler|RK4|Midpoint) functional/odeint.py:28-35
 *   -> FixedSolver.integrate                          solver/base_fixed_solver.py:103-144
 *   -> Euler.step fixed_solver/euler.py:7-11 | RK4.step fixed_solver/rk4.py:7-10 (3/8 rule,
 *      base_fixed_solver.py:166-197) | Midpoint.step fixed_solver/midpoint.py:7-18.  grid == t_span.  out [B,T,D] (base_fixed_solver.py:143).
 * out_stride_t: write only every out_stride_t-th grid point (1 = all; the last point is always
 * written); out then has ceil((T-1)/stride)+1 rows per trajectory. */
int xde_rk_fixed_mlp_f32(int32_t method, const xde_mlp_field_t *field, const float *y0, int64_t B,
                         const float *t_span, int32_t T, int32_t out_stride_t, float *out,
                         void *stream);

/* FixedSolver(step_size= | grid_constructor=)         solver/base_fixed_solver.py:49-89,119-139
 * The reference's loop takes its first len(t_span) steps on the constructed grid and reports, for output i, the linear
 * interpolant of step i evaluated at t_span[i] (interpolation/functional/interp_fn.py:4-10; an extrapolation whenever
 * the grid is finer than t_span -- reproduced as is).  y_grid [B,T,D] = the solution on grid[0..T) from
 * xde_rk_fixed_mlp_f32; out [B,T,D]. */
int xde_fixed_interp_linear_f32(const float *y_grid, const float *grid, const float *t_out, int64_t B, int32_t T,
                                int32_t D, float *out, void *stream);

/* sdeint(drift, diffusion, y0, t, solver=Euler)       functional/sdeint.py:30-37
 *   -> BaseSDE.move/fuse xde/base_sde.py:44-61 (intended Euler-Maruyama, diagonal noise) with the
 *      Brownian increments supplied by the caller: dW [T-1,B,D].  out [B,T,D].
 * XDE_SDE_MILSTEIN is an extension (no reference counterpart). */
int xde_sde_mlp_f32(int32_t scheme, const xde_mlp_field_t *drift, const xde_mlp_field_t *diffusion,
                    const float *y0, int64_t B, const float *t_span, int32_t T, const float *dW,
                    int32_t out_stride_t, float *out, void *stream);

/* Tensor-core variants of the two fixed-grid entry points above (same arguments, same reference lines):
 * the dense layers of the field run on tcgen05 (fp16-split three-product GEMMs, fp32 accumulation in
 * TMEM), for D in {16,32,64} x H in {64,128,256} (sde: H <= 128).  Results agree with the FP32 entry
 * points to ~1e-6 relative (north star: rtol 1e-5), not bit for bit; anything else returns
 * XDE_E_UNSUPPORTED_FIELD.  status: optional device word, zeroed by the entry point, set to
 * XDE_ST_TC_RANGE when some stage input left the fp16 operand range (read it after synchronising). */
int xde_rk_fixed_mlp_tc_f32(int32_t method, const xde_mlp_field_t *field, const float *y0, int64_t B,
                            const float *t_span, int32_t T, int32_t out_stride_t, float *out, int32_t *status,
                            void *stream);
int xde_sde_mlp_tc_f32(int32_t scheme, const xde_mlp_field_t *drift, const xde_mlp_field_t *diffusion,
                       const float *y0, int64_t B, const float *t_span, int32_t T, const float *dW,
                       int32_t out_stride_t, float *out, int32_t *status, void *stream);

/* SdeintAdjointMethod.backward                        functional/sdeint_adjoint.py:57-230
 * The reference's reverse solve is a placeholder (augmented_diffusion repeats the drift dynamics, :136-171; BaseSDE
 * cannot be instantiated): there is no behaviour to reproduce.  Implemented: the exact adjoint of the
 * Euler-Maruyama recursion the forward entry computes (discretise, then differentiate), small states (D <= 8,
 * H <= 96).  y_all, grad_y [B,T,D] = the forward solution at EVERY grid point (out_stride_t = 1) and its
 * cotangent; increments from dW [T-1,B,D], or dW = NULL and (seed, traj_offset) for the generator -- the same
 * source as the forward pass; out_gdrift / out_gdiffusion = (gW1,gb1,gW2,gb2) summed over the B trajectories
 * (all-reduce across GPUs is the caller's); out_adj_y0 [B,D] optional (dL/dy0). */
int xde_sde_mlp_adjoint_f32(const xde_mlp_field_t *drift, const xde_mlp_field_t *diffusion, const float *t_span,
                            int32_t T, const float *y_all, const float *grad_y, int64_t B, const float *dW,
                            uint64_t seed, int64_t traj_offset, float *out_gdrift, float *out_gdiffusion,
                            float *out_adj_y0, void *stream);

/* Brownian increments without a table.  The reference draws them on the host from
 * BrownianInterval (utils/brownian/brownian_interval.py:178-240; xde/base_sde.py:35-37); on a fixed grid the
 * solver only ever asks for W(t[n+1]) - W(t[n]), so a counter-based generator addressed by
 * (step n, GLOBAL trajectory index b + traj_offset, component) gives the same thing without state:
 * Philox4x32-10 keyed by `seed`, Box-Muller, dW[n,b,d] = sqrt(|t[n+1]-t[n]|) * N(0,1).  A batch shard that
 * passes its first global row as traj_offset sees exactly the increments of the unsharded run.
 *   xde_brownian_increments_f32 writes the table dW [T-1,B,D] (the parity bridge: feed it to the `dW`
 *   entries above, or to the oracle);  xde_sde_mlp_philox_f32 = xde_sde_mlp_f32 / _tc_f32 (math =
 *   XDE_MATH_FP32 / XDE_MATH_TENSOR) generating the same increments on the fly -- bit-identical results. */
int xde_brownian_increments_f32(uint64_t seed, int64_t traj_offset, const float *t_span, int32_t T, int64_t B,
                                int32_t D, float *dW, void *stream);
int xde_sde_mlp_philox_f32(int32_t scheme, int32_t math, const xde_mlp_field_t *drift,
                           const xde_mlp_field_t *diffusion, const float *y0, int64_t B, const float *t_span,
                           int32_t T, uint64_t seed, int64_t traj_offset, int32_t out_stride_t, float *out,
                           int32_t *status, void *stream);

/* HistoryIndex.forward                                xde/base_dde.py:84-118
 *   -> InterpolationBase.evaluate / derivative        interpolation/interpolate_base.py:49-114
 *      (LinearInterpolation interpolate.py:6-97, CubicHermiteSpline :100-204, BezierSpline :207-298;
 *      interp_method "linear" | "cubic" | "bez", xde/base_dde.py:104-111).
 * his [R,Th,D] (R = product of leading dims), his_span [Th], lags [L]; out_val, out_der [R,L,D]. */
int xde_history_gather_f32(int32_t kind, const float *his, int64_t R, int32_t Th, int32_t D,
                           const float *his_span, const float *lags, int32_t L, float *out_val,
                           float *out_der, void *stream);
/* HistoryIndex.backward                               xde/base_dde.py:121-127
 * g_lags[l] = sum_{r,d} grad_y[r,l,d] * deriv[r,l,d].  g_lags [L]. */
int xde_history_gather_bwd_f32(const float *grad_y, const float *deriv, int64_t R, int32_t L, int32_t D,
                               float *g_lags, void *stream);

/* BaseDDE.fuse (damped Euler update)                  xde/base_dde.py:55-58
 * y1 = (dy - 0.001*(dy*dt + y0))*dt + y0, elementwise over n values. */
int xde_dde_fuse_f32(const float *dy, float dt, const float *y0, int64_t n, float *y1, void *stream);
/* its cotangents (the solution of ddeint is differentiated through by the D3STN trainer,
 * example/D3STN/train_dde.py:424-454): grad_dy = grad_y1 * dt*(1 - 0.001*dt), grad_y0 = grad_y1 * (1 - 0.001*dt);
 * either output may be NULL. */
int xde_dde_fuse_bwd_f32(const float *grad_y1, float dt, int64_t n, float *grad_dy, float *grad_y0, void *stream);

/* ---- torch-free host surface (csrc/xde_hostapi.cu) ---------------------------------------------------------
 * The reference creates its tensors with Paddle eager ops (solver/base_adaptive_solver.py:25-31,
 * base_fixed_solver.py:119-143).  A host layer that wants no tensor library binds these instead: device memory from
 * the stream-ordered pool, copies, streams, and a DLPack producer through which Paddle / PyTorch / CuPy import the
 * results without a copy (paddlexde_b200/_native.py is such a host layer: ctypes + numpy only). */
int xde_device_count(int32_t *n);
int xde_set_device(int32_t dev);
int xde_get_device(int32_t *dev);
int xde_malloc(void **ptr, uint64_t bytes, void *stream);
int xde_free(void *ptr, void *stream);
/* kind: 0 host->device, 1 device->host, 2 device->device */
int xde_memcpy_async(void *dst, const void *src, uint64_t bytes, int32_t kind, void *stream);
int xde_memset_async(void *dst, int32_t value, uint64_t bytes, void *stream);
int xde_host_alloc(void **ptr, uint64_t bytes); /* pinned */
int xde_host_free(void *ptr);
int xde_stream_create(void **stream);
int xde_stream_destroy(void *stream);
int xde_stream_synchronize(void *stream);
/* -> DLManagedTensor* (DLPack v0.8) for a contiguous device buffer; dtype_code 2 = float, 0 = int; owns != 0: the
 * deleter frees the buffer.  Wrap it in a PyCapsule named "dltensor".  NULL on a bad argument. */
void *xde_dlpack_wrap(void *dev_ptr, int32_t ndim, const int64_t *shape, int32_t dtype_code, int32_t bits,
                      int32_t device_id, int32_t owns);
void xde_dlpack_release(void *managed);
/* in-place sum of the adjoint parameter gradients over the batch shards, the only collective of the
 * path (example/D3STN/train_dde.py:201-202,454-456).  comm: the caller's ncclComm_t. */
int xde_allreduce_grads(void *comm, float *buf, int64_t n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* XDE_B200_H */
