// xde_dopri5_adj_batch.cu -- OdeintAdjointMethod.backward with the REFERENCE-FAITHFUL controller:
// one dt and one error norm for the whole augmented state (functional/odeint_adjoint.py:47-167 with
// the solver of solver/base_adaptive_solver_rk.py), controller = XDE_CTRL_BATCH, and either adjoint norm:
//   XDE_ADJ_NORM_MIXED  max(|g_t|, rms(y), rms(a), max_p rms(g_theta,p))   (default, :284-291)
//   XDE_ADJ_NORM_SEMI   max(|g_t|, rms(y), rms(a))                          ("seminorm", :301-309)
//
// One cooperative launch for the whole backward pass; the controller is replicated in every thread
// (as in xde_dopri5_batch.cu).  Every thread owns a fixed grid-stride set of trajectories; their
// (y, a) and FSAL derivatives live in L2-resident double buffers that are flipped on accept.
//
// Parameter gradients (round 2: sequence-exact).  g_theta is a state of the solve like y and a, replicated in the
// shared memory of every CTA, and advanced with the oracle's flat-state arithmetic from the seven stage derivatives
// k_0..k_6 of the parameter-gradient dynamics.  Each k_i^theta is a sum over the batch, and at the reference's
// default tolerances the error estimate of the g_theta part is the rounding noise of that sum -- the accept/reject
// sequence of the mixed norm depends on HOW it is summed.  The sum is therefore specified so that it has one value on
// every machine (xde_fixed128.cuh, oracle/xde_oracle.c adj_rhs): per 32 consecutive trajectories a sequential fp32
// fma chain (this is what a warp's fold of its 32 tile columns computes), the chain values added exactly in a
// 128-bit fixed-point accumulator (carry-free integer limbs: shared-memory REDs per CTA, then global REDs),
// the total rounded once to fp32.  Per attempt the six new stage sums travel through one grid-wide reduction
// (the same two grid.sync() that the error norm of (y, a) needs); the FSAL property carries over: k_6^theta of an
// accepted attempt is k_0^theta of the next one.  Both norms run this path; with it the whole (dt, ratio, accept)
// sequence of the reference's DEFAULT configuration (mixed norm, rtol 1e-7) equals the oracle's bit for bit.
#include <cooperative_groups.h>

#include <type_traits>

#include "xde_common.cuh"
#include "xde_fixed128.cuh"

namespace cg = cooperative_groups;

namespace xde {

constexpr int kABThreads = 128;
constexpr int kABWarps = kABThreads / 32;
constexpr int kABTileStride = 33;
constexpr int kABScalars = 8;  // scalar slots in front of the vectors in a reduction row

struct AdjBatchParams {
  xde_mlp_field_t field;
  const float *t_span, *y_ans, *grad_y;
  float *out_g;    // [P] fp32
  unsigned long long *gfx;  // [3][6][P][kFxGLimbs] grid-wide fixed-point accumulators, triple buffered (zeroed)
  int *gbad;                // [3] "an addend was not representable" flags (zeroed)
  float *adj_y0;   // [B,D] or null
  long long B;
  int T;
  xde_ctrl_opts_t o;
  int mixed;
  xde_stats_t *stats;
  xde_attempt_t *log_records;
  int *log_counts;
  int log_cap;
  float *ws;        // [4][B*2D]: S[2], F[2]
  double *partial;  // [gridDim.x][row]
  double *reduced;  // [row]
  int row;          // kABScalars
};

template <int D, int HPL, int PRE>
__global__ void __launch_bounds__(kABThreads, 1) dopri5_adj_batch_kernel(const AdjBatchParams p) {
  constexpr int C = 2 * D;
  constexpr int NV = 2 * D + 1;        // accumulator pairs per hidden-unit pair
  constexpr int NTP = NV * HPL;
  constexpr int REC = SmallRec<D>::REC;
  constexpr int CST = ((2 * D + 1 + 3) / 4) * 4;
  cg::grid_group grid = cg::this_grid();

  extern __shared__ __align__(16) float smem[];
  __shared__ double s_warp[kABWarps];
  __shared__ int s_bad;

  const int H = p.field.h, NP = SmallRec<D>::pairs(H);
  const int P = 2 * D * H + H + D;
  const int P4 = ((P + 3) / 4) * 4;
  // accumulator slots of one stage vector: kind i (gW1 rows, gb1, gW2 columns) x parity e x hidden-unit pair, then the
  // D slots of gb2.  Consecutive lanes (pairs) hit consecutive 32-bit words: the shared-memory REDs are conflict free
  // (the parameter order [i*H + j] made them 4-way conflicts).  PS = slots per vector; limb k of slot s at [k*PS + s].
  const int HP2 = 32 * HPL;  // pairs per parity row, padded to the lanes
  const int PS = NV * 2 * HP2 + D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // shared-memory carve-up
  float *sw = smem;
  float *st = sw + SmallRec<D>::floats(H);
  float *g0 = st + ((p.T + 3) / 4) * 4;            // [P]     g_theta: a state of the solve, replicated in every CTA
  float *kth = g0 + P4;                            // [8][P4] k_0..k_6 of the g_theta dynamics, [7] = the Euler probe
  int *fxacc = reinterpret_cast<int *>(kth + 8 * P4);  // [6][kFxSLimbs][PS] CTA partial sums (carry-free limbs)
  double *sred = reinterpret_cast<double *>(fxacc + (((size_t)6 * PS * kFxSLimbs + 3) / 4) * 4);  // [kABScalars]
  float *wbase = reinterpret_cast<float *>(sred + p.row) + (size_t)warp * (4 * NP * kABTileStride + 32 * CST);
  float4 *tile = reinterpret_cast<float4 *>(wbase);
  float *coef = wbase + 4 * NP * kABTileStride;

  load_small_field<D>(sw, p.field);
  const float tsign = (p.t_span[1] > p.t_span[0]) ? -1.0f : 1.0f;  // backward sweep as s = tsign * t
  for (int i = threadIdx.x; i < p.T; i += blockDim.x) st[i] = tsign * p.t_span[i];
  for (int i = threadIdx.x; i < P; i += blockDim.x) g0[i] = 0.0f;
  for (int i = threadIdx.x; i < 8 * P4; i += blockDim.x) kth[i] = 0.0f;
  for (int i = threadIdx.x; i < 6 * PS * kFxSLimbs; i += blockDim.x) fxacc[i] = 0;
  for (int i = lane; i < NP * kABTileStride; i += 32) tile[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = lane; i < 32 * CST; i += 32) coef[i] = 0.0f;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();

  const xde_ctrl_opts_t o = p.o;
  const bool mixed = p.mixed != 0;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  const long long wbeg = gtid - lane;  // first trajectory of this warp in iteration 0 (a multiple of 32)
  const long long nel = p.B * C;
  float *Sb[2] = {p.ws, p.ws + nel};
  float *Fb[2] = {p.ws + 2 * nel, p.ws + 3 * nel};
  const double n_half = (double)(p.B * D);  // elements of the y part (= of the a part)
  const bool leader = (gtid == 0);
  const float ths = -tsign;  // d g_theta / ds = -tsign * vjp_theta(a)

  // field + VJP of this lane's trajectory (xde_dopri5_adj.cu, Appendix B); writes the tile column
  auto eval = [&](const float (&yin)[C], float (&fo)[C]) {
    float u[D];
    f32x2 accf[D], pdu[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      u[k] = pre_act<PRE>(yin[k]);
      accf[k] = pk1(0.0f);
      pdu[k] = pk1(0.0f);
    }
    // U hidden-unit pairs per trip, weight records read first and tile columns stored last (see xde_dopri5_adj.cu:
    // ptxas cannot prove that the tile and the records do not alias and would serialise the pairs)
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
#pragma unroll
      for (int i = 0; i < U; ++i) {
        float h0, h1, z0, z1;
        upk(h[i], h0, h1);
        upk(dz[i], z0, z1);
        tile[(jp0 + i) * kABTileStride + lane] = make_float4(h0, h1, z0, z1);
      }
    };
    constexpr int UT = (D <= 2) ? 5 : 2;
    int jp = 0;
#pragma unroll 1
    for (; jp + UT <= NP; jp += UT) eval_pairs(jp, std::integral_constant<int, UT>());
#pragma unroll 1
    for (; jp < NP; ++jp) eval_pairs(jp, std::integral_constant<int, 1>());
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float fe, fod, ue, uo;
      upk(accf[d], fe, fod);
      upk(pdu[d], ue, uo);
      fo[d] = tsign * ((fe + fod) + sw[NP * REC + d]);
      fo[D + d] = (-tsign) * ((ue + uo) * pre_act_grad<PRE>(yin[d]));
    }
  };
  // fold the warp's 32 tile columns into X = this lane's slice of sum_b k^theta(b); valid = 0/1 per lane
  auto fold = [&](const float (&yin)[C], float valid, f32x2 (&X)[NTP], float (&xb)[D]) {
    float *c = coef + lane * CST;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      c[d] = valid * pre_act<PRE>(yin[d]);
      c[D + 1 + d] = valid * yin[D + d];
      xb[d] = valid * yin[D + d];
    }
    c[D] = valid;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NTP; ++i) X[i] = pk1(0.0f);
#pragma unroll 8
    for (int b = 0; b < 32; ++b) {
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
        // branch-free (one basic block for the 32 columns): a lane without a pair folds a copy of the last row
        // and its sums are cleared below
        const int jp = lane + 32 * q;
        const float4 hv = tile[(jp < NP ? jp : NP - 1) * kABTileStride + b];
        const f32x2 hp = pk(hv.x, hv.y), dzp = pk(hv.z, hv.w);
#pragma unroll
        for (int k = 0; k < D; ++k) X[q * NV + k] = fma2(pk1(cb[k]), dzp, X[q * NV + k]);
        X[q * NV + D] = fma2(pk1(cb[D]), dzp, X[q * NV + D]);
#pragma unroll
        for (int d = 0; d < D; ++d) X[q * NV + D + 1 + d] = fma2(pk1(cb[D + 1 + d]), hp, X[q * NV + D + 1 + d]);
      }
    }
#pragma unroll
    for (int q = 0; q < HPL; ++q)
      if (lane + 32 * q >= NP) {
#pragma unroll
        for (int i = 0; i < NV; ++i) X[q * NV + i] = pk1(0.0f);
      }
    __syncwarp();
  };

  // ---------------- exact batch sums of a stage's parameter-gradient terms (xde_fixed128.cuh) ----------------
  // X: the fp32 chain values of this warp's 32 trajectories (lane = hidden-unit pair); xb: a_d of this lane's
  // trajectory (0 for a lane past the batch).  Added into the CTA's accumulators of stage vector `v`.
  auto accumulate = [&](const f32x2 (&X)[NTP], const float (&xb)[D], int v) {
    int *dst = fxacc + (size_t)v * PS * kFxSLimbs;
    bool ok = true;
#pragma unroll
    for (int q = 0; q < HPL; ++q)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * (lane + 32 * q) + e;
        if (j < H) {
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            float lo, hi;
            upk(X[q * NV + i], lo, hi);
            ok = fx_limbs_add_float(dst + (i * 2 + e) * HP2 + lane + 32 * q, PS, e ? hi : lo) && ok;
          }
        }
      }
    // gb2 slots: one addend per trajectory; summed over the warp in integer arithmetic first (one atomic per warp)
#pragma unroll
    for (int d = 0; d < D; ++d) {
      Fx128 f;
      if (!fx_from_float(xb[d], f)) {
        ok = false;
        f.lo = f.hi = 0ull;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        Fx128 g;
        g.lo = __shfl_xor_sync(XDE_FULL_MASK, f.lo, off);
        g.hi = __shfl_xor_sync(XDE_FULL_MASK, f.hi, off);
        f = fx_add(f, g);
      }
      if (lane == 0) {  // the warp's exact integer sum, split over the limbs (carry-free: signed limbs with headroom)
        const bool neg = (f.hi >> 63) != 0ull;
        unsigned long long lo = f.lo, hi = f.hi;
        if (neg) {
          lo = ~lo + 1ull;
          hi = ~hi + (lo == 0ull ? 1ull : 0ull);
        }
        int *a = dst + NV * 2 * HP2 + d;
#pragma unroll
        for (int k = 0; k < kFxSLimbs; ++k) {
          const int sh = kFxSBits * k;
          const unsigned long long w = (sh < 64) ? ((lo >> sh) | (sh ? (hi << (64 - sh)) : 0ull)) : (hi >> (sh - 64));
          const int limb = (int)(w & ((1u << kFxSBits) - 1u));
          if (limb) atomicAdd(a + k * PS, neg ? -limb : limb);
        }
      }
    }
    if (!ok) s_bad = 1;
  };

  // ---------------- grid reduction of the scalars: deterministic (thread -> warp -> CTA -> fixed-order re-sum) ----------------
  auto scalar_to_row = [&](double v, int slot) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(XDE_FULL_MASK, v, off);
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kABWarps; ++w) s += s_warp[w];
      sred[slot] = s;
    }
    __syncthreads();
  };
  auto clear_row = [&](int ncols) {
    for (int i = threadIdx.x; i < ncols; i += blockDim.x) sred[i] = 0.0;
    __syncthreads();
  };
  // publish the CTA row, reduce columns (column c by CTA c mod grid), broadcast the reduced row back to sred
  auto grid_reduce = [&](int ncols) {
    double *my = p.partial + (size_t)blockIdx.x * p.row;
    for (int i = threadIdx.x; i < ncols; i += blockDim.x) my[i] = sred[i];
    grid.sync();
    for (int col = blockIdx.x; col < ncols; col += gridDim.x) {
      double v = 0.0;
      for (unsigned r = threadIdx.x; r < gridDim.x; r += blockDim.x) v += p.partial[(size_t)r * p.row + col];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(XDE_FULL_MASK, v, off);
      if (lane == 0) s_warp[warp] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kABWarps; ++w) s += s_warp[w];
        p.reduced[col] = s;
      }
      __syncthreads();
    }
    grid.sync();
    for (int i = threadIdx.x; i < ncols; i += blockDim.x) sred[i] = p.reduced[i];
    __syncthreads();
  };
  // One reduction phase: the scalars already sit in sred[0..kABScalars); the CTA's 128-bit partial sums of `nv` stage
  // vectors go to the grid-wide accumulators of this phase's buffer, and after the two grid.sync() of the scalar
  // reduction every CTA converts the totals to fp32 into kth[slot0 + v].  Buffers rotate over three phases: the one
  // the NEXT phase will use was last read two phases ago, i.e. before the previous phase's grid.sync().
  int phase = 0;
  auto reduce_phase = [&](int nv, int slot0) {
    __syncthreads();  // every warp has added its last chain values
    unsigned long long *gb = p.gfx + (size_t)(phase % 3) * 6 * P * kFxGLimbs;
    for (int i = threadIdx.x; i < nv * P; i += blockDim.x) {
      const int v = i / P, q = i - v * P;  // parameter q of stage vector v -> its accumulator slot
      int slot;
      if (q < D * H) {
        const int k = q / H, j = q - k * H;
        slot = (k * 2 + (j & 1)) * HP2 + (j >> 1);
      } else if (q < D * H + H) {
        const int j = q - D * H;
        slot = (D * 2 + (j & 1)) * HP2 + (j >> 1);
      } else if (q < D * H + H + H * D) {
        const int r = q - D * H - H, j = r / D, d = r - j * D;
        slot = ((D + 1 + d) * 2 + (j & 1)) * HP2 + (j >> 1);
      } else {
        slot = NV * 2 * HP2 + (q - D * H - H - H * D);
      }
      int *a = fxacc + (size_t)v * PS * kFxSLimbs + slot;
      fx_glimbs_add(gb + (size_t)i * kFxGLimbs, fx_limbs_total(a, PS));
#pragma unroll
      for (int k = 0; k < kFxSLimbs; ++k) a[k * PS] = 0;
    }
    if (threadIdx.x == 0 && s_bad) atomicOr(p.gbad + phase % 3, 1);
    if (blockIdx.x == 0) {
      unsigned long long *gn = p.gfx + (size_t)((phase + 1) % 3) * 6 * P * kFxGLimbs;
      for (int i = threadIdx.x; i < 6 * P * kFxGLimbs; i += blockDim.x) gn[i] = 0ull;
      if (threadIdx.x == 0) p.gbad[(phase + 1) % 3] = 0;
    }
    __threadfence();
    grid_reduce(kABScalars);
    const bool bad = __ldcg(p.gbad + phase % 3) != 0;
    for (int i = threadIdx.x; i < nv * P; i += blockDim.x) {
      const int v = i / P, q = i - v * P;
      kth[(slot0 + v) * P4 + q] = bad ? NAN : ths * fx_to_float(fx_glimbs_total(gb + (size_t)i * kFxGLimbs));
    }
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    phase++;
  };
  // max_p rms over the four parameter tensors of a P-vector given element-wise by fn(idx) (computed by all
  // threads of the CTA redundantly per tensor; fixed order)
  auto param_norm = [&](auto fn) -> float {
    const int off[5] = {0, D * H, D * H + H, D * H + H + H * D, P};
    float pm = 0.0f;
    for (int q = 0; q < 4; ++q) {
      double s = 0.0;
      for (int i = off[q] + threadIdx.x; i < off[q + 1]; i += blockDim.x) {
        const float v = fn(i);
        s += (double)(v * v);
      }
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(XDE_FULL_MASK, s, o2);
      if (lane == 0) s_warp[warp] = s;
      __syncthreads();
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < kABWarps; ++w) tot += s_warp[w];
      __syncthreads();
      const float r = (float)sqrt(tot / (double)(off[q + 1] - off[q]));
      if (q == 0 || r > pm) pm = r;
    }
    return pm;
  };
  auto mixed_of = [&](double sy, double sa, float pn) -> float {  // Python max semantics, |g_t| = 0 first
    const float ny = (float)sqrt(sy / n_half), na = (float)sqrt(sa / n_half);
    float best = 0.0f;
    if (ny > best) best = ny;
    if (na > best) best = na;
    if (mixed && pn > best) best = pn;
    return best;
  };

  int cur = 0, status = 0, n_logged = 0;
  unsigned long long n_att = 0, n_acc = 0, n_fe = 0;
  const bool has_first = (o.first_step == o.first_step);

  for (int seg = p.T - 1; seg >= 1 && status == 0; --seg) {
    const float t_start = st[seg], te = st[seg - 1];
    // ======== segment start: y <- y_ans[seg], a <- a + grad_y[seg] (functional/odeint_adjoint.py:75-82,153-159) ========
    // ======== INIT 0: f0 = rhs(start), k_0^theta, d0, d1 (base_adaptive_solver.py:44-57) ========
    double sy0 = 0.0, sa0 = 0.0, sy1 = 0.0, sa1 = 0.0;
    for (long long base = wbeg; base < p.B; base += gstride) {
      const long long b = base + lane;
      const bool ok = b < p.B;
      float s0[C], fo[C];
#pragma unroll
      for (int e = 0; e < D; ++e) {
        const long long src = ((long long)seg * p.B + (ok ? b : 0)) * D + e;
        s0[e] = ok ? p.y_ans[src] : 0.0f;
        const float prev = (seg == p.T - 1 || !ok) ? 0.0f : Sb[cur][b * C + D + e];
        s0[D + e] = ok ? ((seg == p.T - 1) ? p.grad_y[src] : (prev + p.grad_y[src])) : 0.0f;
      }
      eval(s0, fo);
      f32x2 X[NTP];
      float xb[D];
      fold(s0, ok ? 1.0f : 0.0f, X, xb);
      accumulate(X, xb, 0);
      if (ok) {
#pragma unroll
        for (int e = 0; e < C; ++e) {
          Sb[cur][b * C + e] = s0[e];
          Fb[cur][b * C + e] = fo[e];
          const float sc = o.atol + fabsf(s0[e]) * o.rtol;
          const float v0 = __fdiv_rn(s0[e], sc), v1 = __fdiv_rn(fo[e], sc);
          if (e < D) {
            sy0 += (double)(v0 * v0);
            sy1 += (double)(v1 * v1);
          } else {
            sa0 += (double)(v0 * v0);
            sa1 += (double)(v1 * v1);
          }
        }
      }
    }
    clear_row(kABScalars);
    scalar_to_row(sy0, 0);
    scalar_to_row(sa0, 1);
    scalar_to_row(sy1, 2);
    scalar_to_row(sa1, 3);
    reduce_phase(1, 0);  // -> kth[0] = k_0^theta
    float t0 = t_start, dt;
    if (has_first) {
      dt = o.first_step;
      n_fe += 1;
    } else {
      float pn0 = 0.f, pn1 = 0.f;
      if (mixed) {
        pn0 = param_norm([&](int i) { return __fdiv_rn(g0[i], o.atol + fabsf(g0[i]) * o.rtol); });
        pn1 = param_norm([&](int i) { return __fdiv_rn(kth[i], o.atol + fabsf(g0[i]) * o.rtol); });
      }
      const float d0 = fabsf(mixed_of(sred[0], sred[1], pn0));
      const float d1 = fabsf(mixed_of(sred[2], sred[3], pn1));
      __syncthreads();
      float h0;
      if (d0 < 1e-5f || d1 < 1e-5f) h0 = 1e-6f; else h0 = __fdiv_rn(0.01f * d0, d1);
      h0 = fabsf(h0);
      // ======== INIT 1: Euler probe rhs(start + h0 f0) (base_adaptive_solver.py:60-64) ========
      double sy2 = 0.0, sa2 = 0.0;
      for (long long base = wbeg; base < p.B; base += gstride) {
        const long long b = base + lane;
        const bool ok = b < p.B;
        float s0[C], f0[C], yi[C], f1[C];
#pragma unroll
        for (int e = 0; e < C; ++e) {
          s0[e] = ok ? Sb[cur][b * C + e] : 0.0f;
          f0[e] = ok ? Fb[cur][b * C + e] : 0.0f;
          yi[e] = f0[e] * h0 + s0[e];
        }
        eval(yi, f1);
        if (mixed) {
          f32x2 X[NTP];
          float xb[D];
          fold(yi, ok ? 1.0f : 0.0f, X, xb);
          accumulate(X, xb, 0);
        } else {
          __syncwarp();  // the tile column is rewritten by the next evaluation
        }
        if (ok) {
#pragma unroll
          for (int e = 0; e < C; ++e) {
            const float sc = o.atol + fabsf(s0[e]) * o.rtol;
            const float v = __fdiv_rn(f1[e] - f0[e], sc);
            if (e < D) sy2 += (double)(v * v); else sa2 += (double)(v * v);
          }
        }
      }
      clear_row(kABScalars);
      scalar_to_row(sy2, 0);
      scalar_to_row(sa2, 1);
      reduce_phase(mixed ? 1 : 0, 7);  // -> kth[7] = the probe's k^theta
      float pn2 = 0.f;
      if (mixed)
        pn2 = param_norm([&](int i) {
          return __fdiv_rn(kth[7 * P4 + i] - kth[i], o.atol + fabsf(g0[i]) * o.rtol);
        });
      const float d2 = fabsf(__fdiv_rn(mixed_of(sred[0], sred[1], pn2), h0));
      __syncthreads();
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
      n_fe += 3;
    }

    // ======== attempts until the segment end is reached (AdaptiveRKSolver.step, :116-127) ========
    int n_steps = 0;
    bool seg_done = false;
    while (!seg_done) {
      if (!(n_steps < o.max_num_steps)) {
        status = XDE_ST_MAX_STEPS;
        break;
      }
      if (!(t0 + dt > t0)) {
        status = XDE_ST_DT_UNDERFLOW;
        break;
      }
      const float t1 = t0 + dt;
      const bool fin = !(te > t1);
      const float x = fin ? __fdiv_rn(te - t0, t1 - t0) : 0.f;
      double sqy = 0.0, sqa = 0.0, bad = 0.0;
      for (long long base = wbeg; base < p.B; base += gstride) {
        const long long b = base + lane;
        const bool ok = b < p.B;
        float s0[C], kk[7][C], yin[C], fo[C];
        bool finite = true;
#pragma unroll
        for (int e = 0; e < C; ++e) {
          s0[e] = ok ? Sb[cur][b * C + e] : 0.0f;
          kk[0][e] = ok ? Fb[cur][b * C + e] : 0.0f;
          finite = finite && (fabsf(s0[e]) < INFINITY);
        }
        if (!finite) bad += 1.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
          for (int e = 0; e < C; ++e) {
            float s = kk[0][e] * (DP::beta(i, 0) * dt);
#pragma unroll
            for (int j = 1; j <= i; ++j) s = s + kk[j][e] * (DP::beta(i, j) * dt);
            yin[e] = s0[e] + s;
          }
          eval(yin, fo);
#pragma unroll
          for (int e = 0; e < C; ++e) kk[i + 1][e] = fo[e];
          f32x2 X[NTP];
          float xb[D];
          fold(yin, ok ? 1.0f : 0.0f, X, xb);
          accumulate(X, xb, i);
        }
        if (ok) {
          const float two_dt = 2.0f * dt;
#pragma unroll
          for (int e = 0; e < C; ++e) {
            float er = kk[0][e] * (dt * DP::cerr(0));
#pragma unroll
            for (int j = 1; j < 7; ++j) er = er + kk[j][e] * (dt * DP::cerr(j));
            const float tol = o.atol + o.rtol * fmaxf(fabsf(s0[e]), fabsf(yin[e]));
            const float v = __fdiv_rn(er, tol);
            if (e < D) sqy += (double)(v * v); else sqa += (double)(v * v);
            float ynew = yin[e];
            if (fin) {  // dense output at the segment end (interp_fit + interp_evaluate, ode_utils.py:28-77)
              float sm = kk[0][e] * (dt * DP::cmid(0));
#pragma unroll
              for (int j = 1; j < 7; ++j) sm = sm + kk[j][e] * (dt * DP::cmid(j));
              const float ym = s0[e] + sm;
              const float F0 = kk[0][e], F1 = kk[6][e], Y0 = s0[e], Y1 = yin[e];
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
              ynew = total;
            }
            Sb[cur ^ 1][b * C + e] = ynew;
            Fb[cur ^ 1][b * C + e] = kk[6][e];
          }
        }
      }
      // ---- reduce: the error sums of (y, a) and the six new stage sums k_1..k_6 of the g_theta dynamics ----
      clear_row(kABScalars);
      scalar_to_row(sqy, 0);
      scalar_to_row(sqa, 1);
      scalar_to_row(bad, 2);
      reduce_phase(6, 1);
      if (sred[2] > 0.0) {
        status = XDE_ST_NONFINITE_STATE;
        break;
      }
      // the g_theta part of the step in the oracle's flat-state arithmetic (xde_oracle.c drv_adaptive_step):
      // y1 = g + sum_j k_j (beta_5j dt) [FSAL: the stage-6 input], err = sum_j k_j (dt c_err,j)
      auto g_y1 = [&](int i) -> float {
        float s = kth[i] * (DP::beta(5, 0) * dt);
#pragma unroll
        for (int j = 1; j < 6; ++j) s = s + kth[j * P4 + i] * (DP::beta(5, j) * dt);
        return g0[i] + s;
      };
      float pn = 0.f;
      if (mixed)
        pn = param_norm([&](int i) {
          const float g1 = g_y1(i);
          float er = kth[i] * (dt * DP::cerr(0));
#pragma unroll
          for (int j = 1; j < 7; ++j) er = er + kth[j * P4 + i] * (dt * DP::cerr(j));
          const float tol = o.atol + o.rtol * fmaxf(fabsf(g0[i]), fabsf(g1));
          return __fdiv_rn(er, tol);
        });
      const float ratio = fabsf(mixed_of(sred[0], sred[1], pn));
      bool accept = (ratio <= 1.0f);
      if (dt > o.max_step) accept = false;
      if (dt <= o.min_step) accept = true;
      const float dt_next = next_step_size(dt, ratio, o);
      n_att++;
      n_fe += 6;
      n_steps++;
      if (leader && p.log_records && n_logged < p.log_cap) {
        xde_attempt_t r;
        r.t0 = tsign * t0;
        r.dt = tsign * dt;
        r.ratio = ratio;
        r.accepted = accept ? 1 : 0;
        p.log_records[n_logged] = r;
      }
      n_logged++;
      __syncthreads();
      if (accept) {
        n_acc++;
        const float two_dt = 2.0f * dt;
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
          const float Y0 = g0[i], Y1 = g_y1(i), F0 = kth[i], F1 = kth[6 * P4 + i];
          float gnew = Y1;
          if (fin) {
            float sm = kth[i] * (dt * DP::cmid(0));
#pragma unroll
            for (int j = 1; j < 7; ++j) sm = sm + kth[j * P4 + i] * (dt * DP::cmid(j));
            const float ym = Y0 + sm;
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
            gnew = total;
          }
          g0[i] = gnew;
          kth[i] = F1;  // FSAL for the parameter-gradient dynamics (unused after a segment end)
        }
        cur ^= 1;
        t0 = t1;
        if (fin) seg_done = true;
      }
      __syncthreads();  // g0 / kth are rewritten above and read by the next phase
      dt = dt_next;
    }
  }

  // ---------------- epilogue ----------------
  if (status == 0) {
    // the state after the last segment still misses `a += grad_y[0]` (functional/odeint_adjoint.py:157-159)
    if (p.adj_y0)
      for (long long b = gtid; b < p.B; b += gstride)
#pragma unroll
        for (int d = 0; d < D; ++d) p.adj_y0[b * D + d] = Sb[cur][b * C + D + d] + p.grad_y[b * D + d];
  } else if (p.adj_y0) {
    for (long long b = gtid; b < p.B; b += gstride)
#pragma unroll
      for (int d = 0; d < D; ++d) p.adj_y0[b * D + d] = NAN;
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < P; i += blockDim.x) p.out_g[i] = g0[i];
  if (leader) {
    if (p.stats) {
      p.stats->n_attempts = n_att * (unsigned long long)p.B;
      p.stats->n_accepted = n_acc * (unsigned long long)p.B;
      p.stats->nfe = n_fe * (unsigned long long)p.B;
      p.stats->status = status;
    }
    if (p.log_counts) p.log_counts[0] = n_logged;
  }
}

template <int D, int HPL, int PRE>
static int launch_adj_batch(AdjBatchParams &p, cudaStream_t stream) {
  const int H = p.field.h, NP = SmallRec<D>::pairs(H);
  const int P = 2 * D * H + H + D;
  const int P4 = ((P + 3) / 4) * 4;
  constexpr int CST = ((2 * D + 1 + 3) / 4) * 4;
  p.row = kABScalars;
  const size_t smem = sizeof(float) * (SmallRec<D>::floats(H) + ((p.T + 3) / 4) * 4 + (size_t)9 * P4) +
                      sizeof(int) * ((((size_t)6 * ((2 * D + 1) * 2 * 32 * HPL + D) * kFxSLimbs + 3) / 4) * 4) +
                      sizeof(double) * p.row +
                      sizeof(float) * kABWarps * (4 * (size_t)NP * kABTileStride + 32 * CST);
  XDE_REQUIRE(smem <= 200 * 1024, XDE_E_UNSUPPORTED_FIELD, "adjoint (batch controller): field + t_span exceed shared memory");
  auto kern = dopri5_adj_batch_kernel<D, HPL, PRE>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  XDE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kABThreads, smem));
  XDE_REQUIRE(per_sm >= 1, XDE_E_CUDA, "adjoint (batch controller): kernel does not fit an SM");
  long long want = (p.B + kABThreads - 1) / kABThreads;
  long long grid = (long long)sm_count() * per_sm;
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  const long long nel = p.B * 2 * D;
  float *ws = nullptr;
  double *partial = nullptr, *reduced = nullptr;
  unsigned long long *gfx = nullptr;
  const size_t gfx_bytes = sizeof(unsigned long long) * (size_t)3 * 6 * P * kFxGLimbs + 4 * sizeof(int);
  // headroom of the shared-memory limbs: one addend per accumulator, warp and 32-trajectory block
  XDE_REQUIRE((p.B + 31) / 32 / grid < 4000, XDE_E_UNSUPPORTED_FIELD,
              "adjoint (batch controller): more than 4000 trajectory blocks per CTA (B = %lld on %lld CTAs)", p.B, grid);
  XDE_CUDA_CHECK(scratch_alloc((void **)&ws, sizeof(float) * 4 * nel, stream));
  XDE_CUDA_CHECK(scratch_alloc((void **)&partial, sizeof(double) * (size_t)grid * p.row, stream));
  XDE_CUDA_CHECK(scratch_alloc((void **)&reduced, sizeof(double) * p.row, stream));
  XDE_CUDA_CHECK(scratch_alloc((void **)&gfx, gfx_bytes, stream));
  XDE_CUDA_CHECK(cudaMemsetAsync(gfx, 0, gfx_bytes, stream));
  p.ws = ws;
  p.partial = partial;
  p.reduced = reduced;
  p.gfx = gfx;
  p.gbad = reinterpret_cast<int *>(gfx + (size_t)3 * 6 * P * kFxGLimbs);
  void *args[] = {(void *)&p};
  cudaError_t e = cudaLaunchCooperativeKernel((void *)kern, dim3((unsigned)grid), dim3(kABThreads), args, smem, stream);
  count_launch();
  cudaFreeAsync(ws, stream);
  cudaFreeAsync(partial, stream);
  cudaFreeAsync(reduced, stream);
  cudaFreeAsync(gfx, stream);
  if (e != cudaSuccess) {
    set_last_error("cooperative launch of dopri5_adj_batch_kernel failed: %s", cudaGetErrorString(e));
    return XDE_E_CUDA;
  }
  return XDE_OK;
}

template <int D, int HPL>
static int adj_batch_pre(AdjBatchParams &p, cudaStream_t s) {
  switch (p.field.pre) {
    case XDE_PRE_ID: return launch_adj_batch<D, HPL, XDE_PRE_ID>(p, s);
    case XDE_PRE_SQUARE: return launch_adj_batch<D, HPL, XDE_PRE_SQUARE>(p, s);
    case XDE_PRE_CUBE: return launch_adj_batch<D, HPL, XDE_PRE_CUBE>(p, s);
  }
  set_last_error("unknown pre-activation %d", p.field.pre);
  return XDE_E_BAD_ARG;
}

int dopri5_adj_batch(const xde_mlp_field_t *field, const float *t_span, int T, const float *y_ans,
                     const float *grad_y, long long B, const xde_ctrl_opts_t *opts, int adj_norm,
                     float *out_gparams, float *out_adj_y0, xde_stats_t *stats, const xde_attempt_log_t *log,
                     cudaStream_t s) {
  AdjBatchParams p{};
  p.field = *field;
  p.t_span = t_span;
  p.y_ans = y_ans;
  p.grad_y = grad_y;
  p.out_g = out_gparams;
  p.adj_y0 = out_adj_y0;
  p.B = B;
  p.T = T;
  p.o = *opts;
  p.mixed = (adj_norm == XDE_ADJ_NORM_MIXED) ? 1 : 0;
  p.stats = stats;
  p.log_records = log ? log->records : nullptr;
  p.log_counts = log ? log->counts : nullptr;
  p.log_cap = log ? log->cap : 0;
  const int D = field->d, H = field->h;
  if (H <= 64) {
    if (D == 1) return adj_batch_pre<1, 1>(p, s);
    if (D == 2) return adj_batch_pre<2, 1>(p, s);
    if (D == 4) return adj_batch_pre<4, 1>(p, s);
  }
  set_last_error("adjoint (batch controller): field D=%d H=%d has no fused kernel (D in {1,2,4}, H <= 64)", D, H);
  return XDE_E_UNSUPPORTED_FIELD;
}

}  // namespace xde
