// xde_adj_tile.cu -- OdeintAdjointMethod.backward for LARGE states (D = 16 / 32 / 64), per-trajectory controller,
// adjoint seminorm: the same reference loop as xde_dopri5_adj.cu (functional/odeint_adjoint.py:47-167, one fresh
// reverse-time dopri5 solve of (y, a, g_theta) per output segment) when one trajectory no longer fits a thread.
//
// Layout.  A CTA of 256 threads keeps W1, W2 resident in shared memory (xde_tile.cuh) and integrates a tile of TM
// trajectories; all rows of the tile walk through a segment together (select_initial_step, then attempts until the
// slowest row has reached the segment end), every row with its own t, dt and accept/reject decisions.
//   * state: the augmented state of a row is (y, a), 2 D values.  A thread owns R2 = 2 rows x C2 columns like in the
//     forward kernels, and the two half-warps that hold identical copies there split the two PARTS here: parity 0
//     keeps y and its seven Dormand-Prince stages, parity 1 keeps a and its stages -- the register budget of the
//     forward kernel carries the doubled state.
//   * one evaluation of the augmented dynamics (Appendix B) is four register-tiled FP32 GEMMs on operands that pass
//     through shared memory feature-major ([feature][row], so every fetch is a conflict-free LDS.128/LDS.64):
//       Z = U W1 (+b1, tanh -> h)          dH = A W2^T (dZ = dH (1 - h^2))       [TM x H], K = D, layer-1 mapping
//       F = h W2 (+b2)                     dU = dZ W1^T                          [TM x D], K = H, two-chain mapping
//     in exactly the oracle's summation orders (sequential-k fma chains; even / odd hidden-unit chains joined at
//     the end): (y, a), every dt, error ratio and accept/reject flag are bit-identical to the oracle's.
//   * parameter gradients: g_theta = sum over evaluations of W (u^T dZ | dZ | h^T a | a) with the scalar stage
//     weights W known when an attempt starts (xde_dopri5_adj.cu: theta_w).  Per evaluation two more GEMMs with
//     K = the TM rows of the tile, gW1 += (W u)^T dZ and gW2 += h^T (W a): 2 D H = 32 768 running sums per CTA for
//     64-256-64 -- too many for registers or for the shared memory the weights leave over.  They live in TENSOR
//     MEMORY: the SM's 256 KB of TMEM are otherwise idle in an FP32 kernel, every thread owns one TMEM lane
//     (columns [0,256) for warps 0-3, [256,512) for warps 4-7) and round-trips its 64 + 64 partial sums with
//     tcgen05.ld / tcgen05.st around the fused multiply-adds; they are flushed to an fp64 accumulator in global
//     memory every kFlushEvery folds (fp32 partial sums stay short, as in the small-state kernel).
//   * the stage-6 evaluation is folded AFTER the controller (its operands are still in shared memory): with its own
//     weight if the attempt is accepted, plus the stage-0 weight of the next attempt (FSAL);
//   * a rejected attempt has already been folded in: its rows replay stages 0..5 with the negated weights in a
//     theta-only pass (the two [TM x H] GEMMs + the two gradient GEMMs), the other rows carry weight 0.  The f0
//     evaluation of a segment is folded by such a pass too, once select_initial_step has produced dt.
// Bit-exact against the oracle: dL/dy0, step sequences, counters; parameter gradients at rtol 1e-5.
#include "xde_tile.cuh"

namespace xde {

#ifndef XDE_ADJ_TILE_FLUSH
#define XDE_ADJ_TILE_FLUSH 24
#endif
constexpr int kAdjTileFlushEvery = XDE_ADJ_TILE_FLUSH;

struct AdjTileParams {
  xde_mlp_field_t f;
  const float *t_span, *y_ans, *grad_y;
  double *gacc;               // [P] fp64 accumulator (zeroed)
  unsigned long long *queue;  // next unclaimed tile (zeroed)
  float *adj_y0;              // [B, D] or null
  long long B;
  int T;
  xde_ctrl_opts_t o;
  xde_stats_t *stats;
  xde_attempt_t *log_records;
  int *log_counts;
  int log_cap;
};

// ---- TMEM as per-thread scratch: lane = 32 * (warp % 4) + lane id, 256 columns per thread ----
__device__ __forceinline__ uint32_t at_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t a, float (&r)[N]);
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t a, const float (&r)[N]);
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t a, float (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "r"(a)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<8>(uint32_t a, const float (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "f"(r[0]), "f"(r[1]),
               "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t a, float (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3])
               : "r"(a)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<4>(uint32_t a, const float (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(r[0]), "f"(r[1]), "f"(r[2]),
               "f"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int D, int H, int TM, int R1, int C1, int R2, int C2>
struct AdjTileGeom {
  using G = TileGeom<D, H, TM, R1, C1, R2, C2>;
  // gradient blocks: thread t owns gW1[k0 .. k0+KB)[j0 .. j0+JB) and gW2[j0 .. j0+JB)[k0 .. k0+KB)
  static constexpr int PER = D * H / kTileThreads;  // running sums per matrix and thread
  static constexpr int JB = (PER <= 32) ? 4 : 8;
  static constexpr int KB = PER / JB;
  static constexpr int NKB = D / KB;
  static constexpr int NJB = kTileThreads / NKB;
  static_assert(KB >= 1 && KB * JB == PER && KB * NKB == D && JB * NJB == H, "gradient blocks must tile [D x H]");
  static constexpr int KBH = (KB > 4) ? 4 : KB;  // k rows of a block folded per pass (register budget)
  static constexpr int NACC = KB * JB;
  static constexpr int NBIAS = (H + D + kTileThreads - 1) / kTileThreads;  // bias-gradient outputs per thread
  static_assert(2 * NACC + 8 <= 256, "TMEM columns per thread");
  static constexpr int NS = D + 1;  // row stride of the norm tile
  static constexpr size_t floats(int T) {
    return (size_t)G::net_floats + 2 * D * TM + 2 * H * TM + 2 * TM * NS + TM + ((T + 3) / 4) * 4;
  }
};

template <int D, int H, int TM, int R1, int C1, int R2, int C2, int PRE>
__global__ void __launch_bounds__(kTileThreads, 1) adjoint_tile_kernel(const AdjTileParams p) {
  using G = TileGeom<D, H, TM, R1, C1, R2, C2>;
  using AG = AdjTileGeom<D, H, TM, R1, C1, R2, C2>;
  constexpr int NS = AG::NS;
  constexpr int KB = AG::KB, JB = AG::JB, NKB = AG::NKB;
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned long long s_cnt[3];
  __shared__ int s_status;
  __shared__ long long s_tile;
  __shared__ uint32_t s_tmem;
  float *net = smem;
  float *sU = smem + G::net_floats;  // [D][TM]  pre(y) of the evaluation point
  float *sA = sU + D * TM;           // [D][TM]  a of the evaluation point
  float *sH = sA + D * TM;           // [H][TM]  tanh(Z + b1)
  float *sDZ = sH + H * TM;          // [H][TM]  dH (1 - h^2)
  float *sN = sDZ + H * TM;          // [TM][2][NS] squares for the row norms (y part, a part)
  float *sWt = sN + 2 * TM * NS;     // [TM] theta weight of each row for the current fold
  float *st = sWt + TM;              // [T] t_span in solver time
  const float *sW1 = net, *sW2 = net + D * H, *sb1 = sW2 + H * D, *sb2 = sb1 + H;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  load_net<D, H>(net, p.f);
  // direction of the backward sweep: t_span increasing (usual) -> integrate s = -t
  const float tsign = (p.t_span[1] > p.t_span[0]) ? -1.0f : 1.0f;
  for (int i = tid; i < p.T; i += blockDim.x) st[i] = tsign * p.t_span[i];
  if (tid == 0) {
    s_cnt[0] = s_cnt[1] = s_cnt[2] = 0ull;
    s_status = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(at_smem_u32(&s_tmem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // this thread's TMEM scratch: lane 32 (warp % 4) + lane, columns (warp / 4) * 256 ...
  const uint32_t tm = s_tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 256);
  {
    float z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) z[i] = 0.0f;
    for (int c = 0; c < 2 * AG::NACC + 8; c += 8) tmem_st<8>(tm + c, z);
    tmem_wait_st();
  }

  // ---- roles ----
  const int e = lane >> 4;  // 0: this thread keeps the y part of its rows, 1: the a part
  const int pidx = warp * 16 + (lane & 15);
  const int rg2 = pidx % G::NRG2, cg2 = pidx / G::NRG2;
  const int c0 = cg2 * C2;
  const int rg1 = tid % G::NRG1, cg1 = tid / G::NRG1;
  const int kb0 = (tid % NKB) * KB, jb0 = (tid / NKB) * JB;
  const xde_ctrl_opts_t o = p.o;
  const long long n_tiles = (p.B + TM - 1) / TM;
  const bool writer = (cg2 == 0);  // speaks for its rows (stats, log, status): the y-part thread of column group 0
  const bool speaker = writer && e == 0;
  unsigned long long n_att = 0, n_acc = 0, n_fe = 0;
  int status = 0, since_flush = 0;
  const bool has_first = (o.first_step == o.first_step);

  // dense-output weight polynomials of the stage values (interp_fit / interp_evaluate expanded, xde_dopri5_adj.cu)
  auto theta_w = [&](int stg, float dtv, bool finv, float xv) -> float {
    const float csv[7] = {DP::csol(0), DP::csol(1), DP::csol(2), DP::csol(3), DP::csol(4), DP::csol(5), DP::csol(6)};
    const float cmv[7] = {DP::cmid(0), DP::cmid(1), DP::cmid(2), DP::cmid(3), DP::cmid(4), DP::cmid(5), DP::cmid(6)};
    float w;
    if (finv) {
      const float cs = csv[stg], cm = cmv[stg];
      const float d0f = (stg == 0) ? 1.f : 0.f, d6f = (stg == 6) ? 1.f : 0.f;
      const float wa = fmaf(16.0f, cm, fmaf(-5.0f, cs, d6f - 4.0f * d0f));
      const float wb = fmaf(-32.0f, cm, fmaf(14.0f, cs, 5.0f * d0f - 3.0f * d6f));
      const float wc = fmaf(16.0f, cm, fmaf(-8.0f, cs, 2.0f * d6f - 2.0f * d0f));
      const float x2 = xv * xv, x3 = x2 * xv, x4 = x3 * xv;
      const float lin = (stg == 0) ? xv : 0.0f;
      w = dtv * (((lin + x2 * wa) + x3 * wb) + x4 * wc);
    } else {
      w = dtv * csv[stg];
    }
    return -tsign * w;  // d g_theta / ds = -tsign * vjp_theta(a)
  };

  // ---- one evaluation of the augmented dynamics for the tile ----
  // in:  v = this thread's part of the evaluation point (y columns for e = 0, a columns for e = 1)
  // out: fo = the same part of the dynamics (dy/ds = tsign f ; da/ds = -tsign vjp_y(a)); theta_only skips the two
  //      [TM x D] GEMMs (fo is not written).  sU, sA, sH, sDZ hold the operands of the gradient fold afterwards.
  auto eval = [&](const float (&v)[R2][C2], float (&fo)[R2][C2], bool theta_only) {
    __syncthreads();  // the previous readers of sU / sA / sH / sDZ are done
#pragma unroll
    for (int r = 0; r < R2; ++r)
#pragma unroll
      for (int c = 0; c < C2; ++c) {
        if (e == 0) sU[(c0 + c) * TM + rg2 * R2 + r] = pre_act<PRE>(v[r][c]);
        else sA[(c0 + c) * TM + rg2 * R2 + r] = v[r][c];
      }
    __syncthreads();
    // ---- Z = U W1: sequential-k fma chain; h = tanh(Z + b1) -> sH ----
    {
      f32x2 acc[R1][C1 / 2];
#pragma unroll
      for (int r = 0; r < R1; ++r)
#pragma unroll
        for (int c = 0; c < C1 / 2; ++c) acc[r][c] = pk1(0.0f);
      const float4 *u4 = reinterpret_cast<const float4 *>(sU + rg1 * R1);
      const float4 *w4 = reinterpret_cast<const float4 *>(sW1 + cg1 * C1);
#pragma unroll 4
      for (int k = 0; k < D; ++k) {
        const float4 uv = u4[k * (TM / 4)];
        const float ur[4] = {uv.x, uv.y, uv.z, uv.w};
        f32x2 w[C1 / 2];
#pragma unroll
        for (int q = 0; q < C1 / 4; ++q) {
          const float4 wv = w4[k * (H / 4) + q];
          w[2 * q] = pk(wv.x, wv.y);
          w[2 * q + 1] = pk(wv.z, wv.w);
        }
#pragma unroll
        for (int r = 0; r < R1; ++r)
#pragma unroll
          for (int c = 0; c < C1 / 2; ++c) acc[r][c] = fma2(pk1(ur[r]), w[c], acc[r][c]);
      }
#pragma unroll
      for (int c = 0; c < C1 / 2; ++c) {
        const int j = cg1 * C1 + 2 * c;
        const f32x2 b = pk(sb1[j], sb1[j + 1]);
        float h0[R1], h1[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) upk(tanh_rat2(add2(acc[r][c], b)), h0[r], h1[r]);
        *reinterpret_cast<float4 *>(sH + j * TM + rg1 * R1) = make_float4(h0[0], h0[1], h0[2], h0[3]);
        *reinterpret_cast<float4 *>(sH + (j + 1) * TM + rg1 * R1) = make_float4(h1[0], h1[1], h1[2], h1[3]);
      }
    }
    // ---- dH = A W2^T: sequential-d fma chain per hidden unit (orc_mlp_vjp); dZ = dH (1 - h^2) -> sDZ ----
    {
      float acc[R1][C1];
#pragma unroll
      for (int r = 0; r < R1; ++r)
#pragma unroll
        for (int c = 0; c < C1; ++c) acc[r][c] = 0.0f;
      const float4 *a4 = reinterpret_cast<const float4 *>(sA + rg1 * R1);
#pragma unroll 2
      for (int d = 0; d < D; d += 4) {
        float ar[4][4];
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          const float4 av = a4[(d + dd) * (TM / 4)];
          ar[dd][0] = av.x;
          ar[dd][1] = av.y;
          ar[dd][2] = av.z;
          ar[dd][3] = av.w;
        }
#pragma unroll
        for (int c = 0; c < C1; ++c) {
          const float4 wv = *reinterpret_cast<const float4 *>(sW2 + (cg1 * C1 + c) * D + d);
          const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
          for (int dd = 0; dd < 4; ++dd)
#pragma unroll
            for (int r = 0; r < R1; ++r) acc[r][c] = fmaf(ar[dd][r], wr[dd], acc[r][c]);
        }
      }
#pragma unroll
      for (int c = 0; c < C1; ++c) {
        const int j = cg1 * C1 + c;
        const float4 hv = *reinterpret_cast<const float4 *>(sH + j * TM + rg1 * R1);
        const float hr[4] = {hv.x, hv.y, hv.z, hv.w};
        float dz[4];
#pragma unroll
        for (int r = 0; r < R1; ++r) dz[r] = acc[r][c] * fmaf(-hr[r], hr[r], 1.0f);
        *reinterpret_cast<float4 *>(sDZ + j * TM + rg1 * R1) = make_float4(dz[0], dz[1], dz[2], dz[3]);
      }
    }
    if (theta_only) return;
    __syncthreads();
    // ---- F = h W2 (+b2): even / odd hidden-unit chains on the two half-warps, joined by one shuffle ----
    float F[R2][C2];
    {
      f32x2 acc[R2][C2 / 2];
#pragma unroll
      for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int c = 0; c < C2 / 2; ++c) acc[r][c] = pk1(0.0f);
      const float2 *h2 = reinterpret_cast<const float2 *>(sH + rg2 * R2);
      const float4 *w4 = reinterpret_cast<const float4 *>(sW2 + c0);
#pragma unroll 4
      for (int j = e; j < H; j += 2) {
        const float2 hv = h2[j * (TM / 2)];
        f32x2 w[C2 / 2];
#pragma unroll
        for (int q = 0; q < C2 / 4; ++q) {
          const float4 wv = w4[j * (D / 4) + q];
          w[2 * q] = pk(wv.x, wv.y);
          w[2 * q + 1] = pk(wv.z, wv.w);
        }
#pragma unroll
        for (int c = 0; c < C2 / 2; ++c) {
          acc[0][c] = fma2(pk1(hv.x), w[c], acc[0][c]);
          acc[1][c] = fma2(pk1(hv.y), w[c], acc[1][c]);
        }
      }
#pragma unroll
      for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int c = 0; c < C2 / 2; ++c) {
          const f32x2 other = __shfl_xor_sync(XDE_FULL_MASK, acc[r][c], 16);
          float s0, s1;
          upk(add2(acc[r][c], other), s0, s1);
          F[r][2 * c] = s0 + sb2[c0 + 2 * c];
          F[r][2 * c + 1] = s1 + sb2[c0 + 2 * c + 1];
        }
    }
    // ---- dU = dZ W1^T: chain2_dot over the hidden axis (even chain | odd chain packed in one FFMA2); parity e takes
    // half of the thread pair's C2 columns with both chains ----
    float dU[R2][C2];
    {
      constexpr int CH = C2 / 2;
      f32x2 acc[R2][CH];  // (even chain, odd chain)
#pragma unroll
      for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[r][c] = pk1(0.0f);
      const float2 *z2 = reinterpret_cast<const float2 *>(sDZ + rg2 * R2);
      const int cc0 = c0 + e * CH;
#pragma unroll 2
      for (int j = 0; j < H; j += 4) {
        float2 zv[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) zv[jj] = z2[(j + jj) * (TM / 2)];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float4 wv = *reinterpret_cast<const float4 *>(sW1 + (cc0 + c) * H + j);
          const f32x2 w01 = pk(wv.x, wv.y), w23 = pk(wv.z, wv.w);
          acc[0][c] = fma2(pk(zv[0].x, zv[1].x), w01, acc[0][c]);
          acc[1][c] = fma2(pk(zv[0].y, zv[1].y), w01, acc[1][c]);
          acc[0][c] = fma2(pk(zv[2].x, zv[3].x), w23, acc[0][c]);
          acc[1][c] = fma2(pk(zv[2].y, zv[3].y), w23, acc[1][c]);
        }
      }
      float mine[R2][CH];
#pragma unroll
      for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          float ev, od;
          upk(acc[r][c], ev, od);
          mine[r][c] = ev + od;
        }
#pragma unroll
      for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float theirs = __shfl_xor_sync(XDE_FULL_MASK, mine[r][c], 16);
          dU[r][e * CH + c] = mine[r][c];
          dU[r][(1 - e) * CH + c] = theirs;
        }
    }
    // solver-time dynamics of this thread's part; the a part needs pre'(y) of the evaluation point (partner lane)
#pragma unroll
    for (int r = 0; r < R2; ++r)
#pragma unroll
      for (int c = 0; c < C2; ++c) {
        const float yv = __shfl_sync(XDE_FULL_MASK, v[r][c], lane & 15);  // the y-part thread of this pair
        fo[r][c] = (e == 0) ? tsign * F[r][c] : (-tsign) * (dU[r][c] * pre_act_grad<PRE>(yv));
      }
  };

  // ---- gradient fold: gW1 += (w u)^T dZ, gb1 += w dZ, gW2 += h^T (w a), gb2 += w a over the rows of the tile ----
  // sWt holds the row weights (0 for rows that do not take part); the running sums live in this thread's TMEM lane.
  // bias gradients: output q = tid + 256 z is gb1[q] (q < H) or gb2[q - H] (q < H + D)
  double gb_acc[AG::NBIAS];
  float gb_run[AG::NBIAS];
#pragma unroll
  for (int z = 0; z < AG::NBIAS; ++z) {
    gb_acc[z] = 0.0;
    gb_run[z] = 0.0f;
  }
  auto flush_theta = [&]() {
    const int P1 = D * H;
#pragma unroll
    for (int kq = 0; kq < KB; ++kq) {
      for (int jq = 0; jq < JB; jq += 4) {
        float a1[4], a2[4];
        tmem_ld<4>(tm + kq * JB + jq, a1);
        tmem_ld<4>(tm + AG::NACC + kq * JB + jq, a2);
        tmem_wait_ld();
#pragma unroll
        for (int z = 0; z < 4; ++z) {
          atomicAdd(&p.gacc[(kb0 + kq) * H + jb0 + jq + z], (double)a1[z]);                 // gW1[k][j]
          atomicAdd(&p.gacc[P1 + H + (jb0 + jq + z) * D + kb0 + kq], (double)a2[z]);        // gW2[j][d]
          a1[z] = 0.0f;
          a2[z] = 0.0f;
        }
        tmem_st<4>(tm + kq * JB + jq, a1);
        tmem_st<4>(tm + AG::NACC + kq * JB + jq, a2);
      }
    }
    tmem_wait_st();
#pragma unroll
    for (int z = 0; z < AG::NBIAS; ++z) {
      gb_acc[z] += (double)gb_run[z];
      gb_run[z] = 0.0f;
    }
  };
  auto fold = [&]() {
    // caller: sWt written and __syncthreads() passed; skipped by the caller when every weight is zero
    constexpr int KBH = AG::KBH;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      // half 0: gW1[k][j] += sum_b (w_b U[k][b]) dZ[j][b];  half 1: gW2[j][k] += sum_b h[j][b] (w_b A[k][b])
      const float *sK = half ? sA : sU;   // [D][TM], scaled by the row weight
      const float *sJ = half ? sH : sDZ;  // [H][TM]
#pragma unroll 1
      for (int kh = 0; kh < KB; kh += KBH) {
        float acc[KBH][JB];
#pragma unroll
        for (int kq = 0; kq < KBH; ++kq)
#pragma unroll
          for (int jq = 0; jq < JB; ++jq) acc[kq][jq] = 0.0f;
#pragma unroll 1
        for (int b = 0; b < TM; b += 4) {
          const float4 wv = *reinterpret_cast<const float4 *>(sWt + b);
          float kv[KBH][4];
#pragma unroll
          for (int kq = 0; kq < KBH; ++kq) {
            const float4 x = *reinterpret_cast<const float4 *>(sK + (kb0 + kh + kq) * TM + b);
            kv[kq][0] = x.x * wv.x;
            kv[kq][1] = x.y * wv.y;
            kv[kq][2] = x.z * wv.z;
            kv[kq][3] = x.w * wv.w;
          }
#pragma unroll
          for (int jq = 0; jq < JB; ++jq) {
            const float4 x = *reinterpret_cast<const float4 *>(sJ + (jb0 + jq) * TM + b);
            const float jv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int z = 0; z < 4; ++z)
#pragma unroll
              for (int kq = 0; kq < KBH; ++kq) acc[kq][jq] = fmaf(kv[kq][z], jv[z], acc[kq][jq]);
          }
        }
        // add to the running sums in TMEM
#pragma unroll
        for (int kq = 0; kq < KBH; ++kq)
#pragma unroll
          for (int jq = 0; jq < JB; jq += 4) {
            float t4[4];
            const uint32_t col = tm + half * AG::NACC + (kh + kq) * JB + jq;
            tmem_ld<4>(col, t4);
            tmem_wait_ld();
#pragma unroll
            for (int z = 0; z < 4; ++z) t4[z] += acc[kq][jq + z];
            tmem_st<4>(col, t4);
          }
      }
    }
    tmem_wait_st();
    // bias gradients
#pragma unroll
    for (int z = 0; z < AG::NBIAS; ++z) {
      const int q = tid + kTileThreads * z;
      if (q < H + D) {
        const float *src = (q < H) ? (sDZ + q * TM) : (sA + (q - H) * TM);
        float sacc = 0.0f;
        for (int b = 0; b < TM; ++b) sacc = fmaf(sWt[b], src[b], sacc);
        gb_run[z] += sacc;
      }
    }
    if (++since_flush >= kAdjTileFlushEvery) {
      since_flush = 0;
      flush_theta();
    }
  };

  // rms over each part of this thread's rows: every thread contributes the squares of its part's columns, then sums
  // BOTH parts of its rows sequentially in fp64 (the oracle's order) and returns the adjoint seminorm
  // max(|g_t| = 0, rms(y), rms(a)) with Python max semantics (functional/odeint_adjoint.py:304-307)
  auto row_seminorm = [&](const float (&v)[R2][C2], float (&res)[R2]) {
    __syncthreads();  // previous readers of sN are done
#pragma unroll
    for (int r = 0; r < R2; ++r)
#pragma unroll
      for (int c = 0; c < C2; ++c) sN[((rg2 * R2 + r) * 2 + e) * NS + c0 + c] = v[r][c] * v[r][c];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R2; ++r) {
      const float *ry = sN + ((rg2 * R2 + r) * 2) * NS, *ra = ry + NS;
      double sy = 0.0, sa = 0.0;
#pragma unroll 8
      for (int k = 0; k < D; ++k) {
        sy += (double)ry[k];
        sa += (double)ra[k];
      }
      const float ny = rms_from_sumsq(sy, (double)D), na = rms_from_sumsq(sa, (double)D);
      float best = 0.0f;
      if (ny > best) best = ny;
      if (na > best) best = na;
      res[r] = best;
    }
  };
  // publish the row weights of the next fold; returns whether any row takes part
  auto put_weights = [&](const float (&w)[R2]) -> bool {
    __syncthreads();  // the previous fold has read sWt
    if (speaker) {
#pragma unroll
      for (int r = 0; r < R2; ++r) sWt[rg2 * R2 + r] = w[r];
    }
    bool any = false;
#pragma unroll
    for (int r = 0; r < R2; ++r) any = any || (w[r] != 0.0f);
    return __syncthreads_or(any);
  };

  while (true) {
    __syncthreads();
    if (tid == 0) s_tile = (long long)atomicAdd(p.queue, 1ull);
    __syncthreads();
    const long long tile = s_tile;
    if (tile >= n_tiles) break;
    const long long b0 = tile * TM + rg2 * R2;  // first of this thread's R2 rows
    float s0[R2][C2], kk[7][R2][C2], yin[R2][C2];
    float t0[R2], dt[R2], xfin[R2], dt_old[R2], xfin_old[R2];
    bool fin[R2], fin_old[R2], dead[R2], seg_done[R2];
    int n_steps[R2], n_logged[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) {
      dead[r] = !(b0 + r < p.B);
      n_logged[r] = 0;
      dt[r] = dt_old[r] = xfin[r] = xfin_old[r] = 0.0f;
      fin[r] = fin_old[r] = false;
    }
    // aug_state = [y_ans[-1], grad_y[-1]] (functional/odeint_adjoint.py:75-82)
    {
      const float *src = (e == 0) ? p.y_ans : p.grad_y;
#pragma unroll
      for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int c = 0; c < C2; ++c)
          s0[r][c] = dead[r] ? 0.0f : src[((long long)(p.T - 1) * p.B + b0 + r) * D + c0 + c];
    }

    for (int seg = p.T - 1; seg >= 1; --seg) {
      const float t_start = st[seg], te = st[seg - 1];
      float wrow[R2];
#pragma unroll
      for (int r = 0; r < R2; ++r) {
        t0[r] = t_start;
        n_steps[r] = 0;
        seg_done[r] = dead[r];
      }
      auto plan_attempt = [&](int r) {  // (t0, dt, te) -> does the attempt reach the segment end, and where
        const float t1n = t0[r] + dt[r];
        fin[r] = !(te > t1n);
        xfin[r] = fin[r] ? __fdiv_rn(te - t0[r], t1n - t0[r]) : 0.f;
      };
      // ================= select_initial_step (solver/base_adaptive_solver.py:33-72) =================
      {
        float scale[R2][C2], v[R2][C2], d0[R2], d1[R2], d2[R2], h0[R2], f1[R2][C2];
        eval(s0, kk[0], false);
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) {
            scale[r][c] = o.atol + fabsf(s0[r][c]) * o.rtol;
            v[r][c] = __fdiv_rn(s0[r][c], scale[r][c]);
          }
        row_seminorm(v, d0);
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) v[r][c] = __fdiv_rn(kk[0][r][c], scale[r][c]);
        row_seminorm(v, d1);
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          d0[r] = fabsf(d0[r]);
          d1[r] = fabsf(d1[r]);
          if (d0[r] < 1e-5f || d1[r] < 1e-5f) h0[r] = 1e-6f; else h0[r] = __fdiv_rn(0.01f * d0[r], d1[r]);
          h0[r] = fabsf(h0[r]);
#pragma unroll
          for (int c = 0; c < C2; ++c) yin[r][c] = kk[0][r][c] * h0[r] + s0[r][c];  // Euler probe (:60)
        }
        eval(yin, f1, false);
#pragma unroll
        for (int r = 0; r < R2; ++r)
#pragma unroll
          for (int c = 0; c < C2; ++c) v[r][c] = __fdiv_rn(f1[r][c] - kk[0][r][c], scale[r][c]);
        row_seminorm(v, d2);
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          const float dd2 = fabsf(__fdiv_rn(d2[r], h0[r]));
          float h1;
          if (d1[r] <= 1e-15f && dd2 <= 1e-15f) {
            h1 = fmaxf(1e-6f, h0[r] * 1e-3f);
          } else {
            const float mx = (dd2 > d1[r]) ? dd2 : d1[r];
            const float arg = __fdiv_rn(0.01f, mx);
            h1 = (arg > 0.0f && arg < INFINITY) ? root5(arg) : arg;
          }
          h1 = fabsf(h1);
          dt[r] = has_first ? o.first_step : fminf(100.0f * h0[r], h1);
          if (speaker && !dead[r]) n_fe += has_first ? 1u : 3u;
          plan_attempt(r);
          wrow[r] = dead[r] ? 0.0f : theta_w(0, dt[r], fin[r], xfin[r]);
        }
        // the f0 evaluation enters g_theta with the stage-0 weight of the first attempt: theta-only pass at s0
        if (put_weights(wrow)) {
          eval(s0, f1, true);
          __syncthreads();
          fold();
        }
      }

      // ================= attempts until every row of the tile has reached the segment end =================
      while (true) {
        // assertions of _adaptive_step (base_adaptive_solver_rk.py:200-203) + max_num_steps (:120-122)
        {
          float fz[R2];
          row_seminorm(s0, fz);  // isfinite(state).all(): a non-finite component makes a sum of squares non-finite
#pragma unroll
          for (int r = 0; r < R2; ++r) {
            if (seg_done[r]) continue;
            int bad = 0;
            if (!(n_steps[r] < o.max_num_steps)) bad = XDE_ST_MAX_STEPS;
            else if (!(t0[r] + dt[r] > t0[r])) bad = XDE_ST_DT_UNDERFLOW;
            else if (!(fabsf(fz[r]) < INFINITY)) bad = XDE_ST_NONFINITE_STATE;
            if (bad) {
              status = max(status, bad);
              seg_done[r] = dead[r] = true;
              if (e == 1 && p.adj_y0) {
#pragma unroll
                for (int c = 0; c < C2; ++c) p.adj_y0[(b0 + r) * D + c0 + c] = NAN;
              }
              if (speaker && p.log_counts) p.log_counts[b0 + r] = n_logged[r];
            }
          }
        }
        bool all_done = true;
#pragma unroll
        for (int r = 0; r < R2; ++r) all_done = all_done && seg_done[r];
        if (__syncthreads_and(all_done)) break;
#pragma unroll
        for (int r = 0; r < R2; ++r)
          if (seg_done[r]) dt[r] = 0.0f;  // idle rows re-evaluate their state (finite, weight 0)

        // ---- the six stages (base_adaptive_solver_rk.py:129-181): s0 + sum_j k_j (beta_ij dt), products first ----
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
          for (int r = 0; r < R2; ++r)
#pragma unroll
            for (int c = 0; c < C2; ++c) {
              float s = kk[0][r][c] * (DP::beta(i, 0) * dt[r]);
#pragma unroll
              for (int j = 1; j <= i; ++j) s = s + kk[j][r][c] * (DP::beta(i, j) * dt[r]);
              yin[r][c] = s0[r][c] + s;
            }
          eval(yin, kk[i + 1], false);
          if (i < 5) {  // stage 6 is folded after the controller: its weight exists only if the attempt is accepted
#pragma unroll
            for (int r = 0; r < R2; ++r) wrow[r] = seg_done[r] ? 0.0f : theta_w(i + 1, dt[r], fin[r], xfin[r]);
            if (put_weights(wrow)) fold();
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
              float er = kk[0][r][c] * (dt[r] * DP::cerr(0));
#pragma unroll
              for (int j = 1; j < 7; ++j) er = er + kk[j][r][c] * (dt[r] * DP::cerr(j));
              const float tol = o.atol + o.rtol * fmaxf(fabsf(s0[r][c]), fabsf(yin[r][c]));
              v[r][c] = __fdiv_rn(er, tol);
            }
          row_seminorm(v, ratio);
        }
        bool rejected[R2];
        float w_next[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) {
          rejected[r] = false;
          w_next[r] = 0.0f;
          if (seg_done[r]) continue;
          const float t1 = t0[r] + dt[r];
          const float rt = fabsf(ratio[r]);
          bool accept = (rt <= 1.0f);
          if (dt[r] > o.max_step) accept = false;
          if (dt[r] <= o.min_step) accept = true;
          const float dt_next = next_step_size(dt[r], rt, o);
          if (speaker) {
            n_att++;
            n_fe += 6;
            if (p.log_records && n_logged[r] < p.log_cap) {
              xde_attempt_t rec;
              rec.t0 = tsign * t0[r];
              rec.dt = tsign * dt[r];
              rec.ratio = rt;
              rec.accepted = accept ? 1 : 0;
              p.log_records[(b0 + r) * p.log_cap + n_logged[r]] = rec;
            }
          }
          n_logged[r]++;
          n_steps[r]++;
          if (accept) {
            if (speaker) n_acc++;
            const float w_eval = theta_w(6, dt[r], fin[r], xfin[r]);  // weight of the stage-6 evaluation in this attempt
            w_next[r] = w_eval;
            if (fin[r]) {
              // dense output at the segment end (interp_fit + interp_evaluate), then
              // y <- y_ans[i-1], a += grad_y[i-1] (functional/odeint_adjoint.py:153-159)
              const float two_dt = 2.0f * dt[r];
              const float x = xfin[r];
#pragma unroll
              for (int c = 0; c < C2; ++c) {
                const long long src = ((long long)(seg - 1) * p.B + b0 + r) * D + c0 + c;
                if (e == 0) {
                  s0[r][c] = p.y_ans[src];
                } else {
                  float sm = kk[0][r][c] * (dt[r] * DP::cmid(0));
#pragma unroll
                  for (int j = 1; j < 7; ++j) sm = sm + kk[j][r][c] * (dt[r] * DP::cmid(j));
                  const float ym = s0[r][c] + sm;
                  const float F0 = kk[0][r][c], F1 = kk[6][r][c], Y0 = s0[r][c], Y1 = yin[r][c];
                  const float ca = (two_dt * (F1 - F0) - 8.0f * (Y1 + Y0)) + 16.0f * ym;
                  const float cb = ((dt[r] * (5.0f * F0 - 3.0f * F1) + 18.0f * Y0) + 14.0f * Y1) - 32.0f * ym;
                  const float cc = ((dt[r] * (F1 - 4.0f * F0) - 11.0f * Y0) - 5.0f * Y1) + 16.0f * ym;
                  const float cd = dt[r] * F0;
                  float total = Y0 + x * cd;
                  float xp = x * x;
                  total = total + xp * cc;
                  xp = xp * x;
                  total = total + xp * cb;
                  xp = xp * x;
                  total = total + xp * ca;
                  s0[r][c] = total + p.grad_y[src];
                }
              }
              seg_done[r] = true;
            } else {
#pragma unroll
              for (int c = 0; c < C2; ++c) {
                s0[r][c] = yin[r][c];
                kk[0][r][c] = kk[6][r][c];
              }
              t0[r] = t1;
              dt[r] = dt_next;
              plan_attempt(r);
              // FSAL: the stage-6 evaluation (still in shared memory) is also stage 0 of the next attempt
              w_next[r] = w_eval + theta_w(0, dt[r], fin[r], xfin[r]);
            }
          } else {
            // rejected: everything folded for this attempt is taken back by a replay pass below
            dt_old[r] = dt[r];
            fin_old[r] = fin[r];
            xfin_old[r] = xfin[r];
            dt[r] = dt_next;
            plan_attempt(r);
            rejected[r] = true;
          }
        }
        if (put_weights(w_next)) fold();
        // ---- replay: stages 0..5 of the rejected attempts with the weights negated (stage 0: new minus old) ----
        {
          bool any_rej = false;
#pragma unroll
          for (int r = 0; r < R2; ++r) any_rej = any_rej || rejected[r];
          if (__syncthreads_or(any_rej)) {
#pragma unroll 1
            for (int i = 0; i < 6; ++i) {
#pragma unroll
              for (int r = 0; r < R2; ++r)
#pragma unroll
                for (int c = 0; c < C2; ++c) {
                  float s = 0.0f;
                  if (i >= 1) {
                    s = kk[0][r][c] * (DP::beta(i > 0 ? i - 1 : 0, 0) * dt_old[r]);
#pragma unroll
                    for (int j = 1; j < 6; ++j)
                      if (j < i) s = s + kk[j][r][c] * (DP::beta(i > 0 ? i - 1 : 0, j) * dt_old[r]);
                  }
                  yin[r][c] = (i >= 1 && rejected[r]) ? (s0[r][c] + s) : s0[r][c];
                }
#pragma unroll
              for (int r = 0; r < R2; ++r) {
                float w = 0.0f;
                if (rejected[r])
                  w = (i == 0) ? theta_w(0, dt[r], fin[r], xfin[r]) - theta_w(0, dt_old[r], fin_old[r], xfin_old[r])
                               : -theta_w(i, dt_old[r], fin_old[r], xfin_old[r]);
                wrow[r] = w;
              }
              float dummy[R2][C2];
              eval(yin, dummy, true);
              if (put_weights(wrow)) fold();
            }
          }
        }
      }
    }
    // dL/dy0 of the tile's rows
    if (e == 1 && p.adj_y0) {
#pragma unroll
      for (int r = 0; r < R2; ++r)
        if (!dead[r]) {
#pragma unroll
          for (int c = 0; c < C2; ++c) p.adj_y0[(b0 + r) * D + c0 + c] = s0[r][c];
        }
    }
    if (speaker && p.log_counts) {
#pragma unroll
      for (int r = 0; r < R2; ++r)
        if (!dead[r]) p.log_counts[b0 + r] = n_logged[r];
    }
  }

  // ---- epilogue: gradients, stats, TMEM ----
  flush_theta();
#pragma unroll
  for (int z = 0; z < AG::NBIAS; ++z) {
    const int q = tid + kTileThreads * z;
    if (q < H) atomicAdd(&p.gacc[D * H + q], gb_acc[z]);                            // gb1
    else if (q < H + D) atomicAdd(&p.gacc[D * H + H + H * D + (q - H)], gb_acc[z]);  // gb2
  }
  if (n_att | n_acc | n_fe) {
    atomicAdd(&s_cnt[0], n_att);
    atomicAdd(&s_cnt[1], n_acc);
    atomicAdd(&s_cnt[2], n_fe);
  }
  if (status) atomicMax(&s_status, status);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
  if (tid == 0 && p.stats) {
    atomicAdd(&p.stats->n_attempts, s_cnt[0]);
    atomicAdd(&p.stats->n_accepted, s_cnt[1]);
    atomicAdd(&p.stats->nfe, s_cnt[2]);
    atomicMax(&p.stats->status, s_status);
  }
}

template <int D, int H, int TM, int R1, int C1, int R2, int C2, int PRE>
static int launch_adj_tile(const AdjTileParams &p, cudaStream_t s) {
  using AG = AdjTileGeom<D, H, TM, R1, C1, R2, C2>;
  const size_t smem = sizeof(float) * AG::floats(p.T);
  XDE_REQUIRE(smem <= 227 * 1024 - 64, XDE_E_UNSUPPORTED_FIELD,
              "tiled adjoint: weights + tiles + t_span need %zu bytes of shared memory (> 227 KB)", smem);
  auto kern = adjoint_tile_kernel<D, H, TM, R1, C1, R2, C2, PRE>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_tiles = (p.B + TM - 1) / TM;
  long long grid = sm_count();  // one CTA per SM: each allocates the SM's whole tensor memory
  if (grid > n_tiles) grid = n_tiles;
  kern<<<(unsigned)grid, kTileThreads, smem, s>>>(p);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  return XDE_OK;
}

template <int D, int H, int TM, int R1, int C1, int R2, int C2>
static int adj_tile_pre(const AdjTileParams &p, cudaStream_t s) {
  switch (p.f.pre) {
    case XDE_PRE_ID: return launch_adj_tile<D, H, TM, R1, C1, R2, C2, XDE_PRE_ID>(p, s);
    case XDE_PRE_SQUARE: return launch_adj_tile<D, H, TM, R1, C1, R2, C2, XDE_PRE_SQUARE>(p, s);
    case XDE_PRE_CUBE: return launch_adj_tile<D, H, TM, R1, C1, R2, C2, XDE_PRE_CUBE>(p, s);
  }
  set_last_error("unknown pre-activation %d", p.f.pre);
  return XDE_E_BAD_ARG;
}

// called by xde_dopri5_mlp_adjoint_f32 (xde_dopri5_adj.cu) for D >= 16; gacc / queue are the caller's scratch
int dopri5_adj_tile(const xde_mlp_field_t *field, const float *t_span, int T, const float *y_ans, const float *grad_y,
                    long long B, const xde_ctrl_opts_t *opts, double *gacc, unsigned long long *queue, float *out_adj_y0,
                    xde_stats_t *stats, const xde_attempt_log_t *log, cudaStream_t s) {
  AdjTileParams p{};
  p.f = *field;
  p.t_span = t_span;
  p.y_ans = y_ans;
  p.grad_y = grad_y;
  p.gacc = gacc;
  p.queue = queue;
  p.adj_y0 = out_adj_y0;
  p.B = B;
  p.T = T;
  p.o = *opts;
  p.stats = stats;
  p.log_records = log ? log->records : nullptr;
  p.log_counts = log ? log->counts : nullptr;
  p.log_cap = log ? log->cap : 0;
  const int D = field->d, H = field->h;
  // the forward geometries of xde_tile_adaptive.cuh (R2 x C2 <= 16 state values per thread and part)
  if (D == 64 && H == 256) return adj_tile_pre<64, 256, 32, 4, 8, 2, 8>(p, s);
  if (D == 64 && H == 128) return adj_tile_pre<64, 128, 32, 4, 4, 2, 8>(p, s);
  if (D == 32 && H == 256) return adj_tile_pre<32, 256, 32, 4, 8, 2, 4>(p, s);
  if (D == 32 && H == 128) return adj_tile_pre<32, 128, 64, 4, 8, 2, 8>(p, s);
  if (D == 32 && H == 64) return adj_tile_pre<32, 64, 64, 4, 4, 2, 8>(p, s);
  if (D == 16 && H == 64) return adj_tile_pre<16, 64, 64, 4, 4, 2, 4>(p, s);
  set_last_error("adjoint: no fused kernel for D=%d H=%d (small states: D in 1..8; tiles: D=64: H in {128,256}; D=32: H in "
                 "{64,128,256}; D=16: H=64)", D, H);
  return XDE_E_UNSUPPORTED_FIELD;
}

}  // namespace xde
