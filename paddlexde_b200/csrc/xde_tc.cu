// xde_tc.cu -- fixed-grid steppers for LARGE states on the 5th-generation tensor cores (tcgen05 + TMEM).
//   * odeint(..., solver=Euler|Midpoint|RK4): FixedSolver.integrate (solver/base_fixed_solver.py:103-144),
//     Euler.step (fixed_solver/euler.py:7-11), Midpoint.step (midpoint.py:7-18), RK4.step = 3/8 rule
//     (base_fixed_solver.py:166-197);
//   * sdeint(..., solver=Euler): y1 = y0 + f*dt + g*dW (xde/base_sde.py:44-61), dW from the caller's table or
//     from the counter-based generator (xde_common.cuh: BmSource).
// Same boundary as xde_tile.cu (the FP32 FFMA2 path, bit-exact against the oracle); this file is the
// tensor-core path: the two dense layers of the field
//     Z[128 x H] = U[128 x D] W1[D x H],   F[128 x D] = tanh(Z + b1)[128 x H] W2[H x D]
// run as tcgen05.mma (kind::f16, M = 128, fp32 accumulators in TMEM); everything else (tanh, RK stage
// combines, the SDE update) stays in fp32 registers.
//
// Precision.  fp32 operands are split into two fp16 pieces, x = hi + lo (hi = rn16(x), lo = rn16(x - hi):
// 22 significant bits), and each GEMM is three MMAs hi*hi + hi*lo + lo*hi accumulated in fp32 -- the
// dropped lo*lo term and the split error are ~2^-22 relative, the level of fp32 rounding itself.
// Weights are pre-scaled by a power of two (exact) so that their lo pieces stay normal fp16 numbers; the
// tensor core truncates its fp32 accumulator once per MMA, so correction products are accumulated first or
// apart (Geom::NFM / SPLIT_CORR).  Measured against fp64 the error equals the FP32 kernels' own (0.6-1.9 x);
// results are not bit-identical to them (tests: rtol 1e-5).  Stage inputs must fit fp16 (|pre(y)| < 65504):
// checked in the kernel, reported through TcParams::status (XDE_ST_TC_RANGE).
//
// Data flow (one CTA per SM, persistent over tiles of 128 trajectories; row r of the tile = TMEM lane r):
//   * all operands of the activations live in TMEM, never in shared memory:
//       U  (layer-1 A operand): the stage input, written as packed fp16 by the threads that own the
//          state (tcgen05.st), hi and lo side by side;
//       Z  (layer-1 accumulator, fp32) is read back 16 columns per thread (tcgen05.ld), bias + tanh in
//          registers, and the SAME 16 columns are overwritten IN PLACE with the fp16 hi (8 columns) and
//          lo (8 columns) of tanh -- which is exactly the A operand of layer 2 for one K = 16 step;
//       F  (layer-2 accumulator) is read back by the state owners.
//   * shared memory holds only the weights (fp16 hi/lo, K-major, no swizzle: 8 x 16 B core matrices),
//     64-256-64: 128 KB, resident for the whole solve.
//   * 16 compute warps (warp w: lane quarter w%4, column group w/4) + 1 warp whose elected thread issues
//     every MMA.  Hand-offs are mbarriers: u_ready (compute -> MMA), z_ready[c] (tcgen05.commit ->
//     compute, per 64-wide hidden chunk), h_ready[c] (compute -> MMA), f_ready (commit -> compute).
//     Layer-1 chunks are issued back to back, so the tanh epilogue of chunk c overlaps the MMAs of
//     chunk c+1 and the layer-2 MMAs of chunk c-1.  No barrier is ever needed in the other direction:
//     every buffer's next writer is ordered behind its last reader by the chain itself.
//   * the epilogue binds (1 tanh per 4 D algorithmic FLOP): tanh is a mix of the 13/6 rational (FMA pipe)
//     and an ex2/rcp form (XU pipe), 3 : 5 per thread block of 8 pairs (tanh_mixed2).
// Two kernels: fixed_tc_kernel (one tile per CTA, any supported shape) and fixed_tc2_kernel (two tiles in
// flight per CTA, software-pipelined, for fields that need <= 256 TMEM columns: H = 64 networks, cfg4).
#include <cuda_fp16.h>

#include <algorithm>

#include "xde_common.cuh"

namespace xde {
namespace tc {

// compute warps = 4 lane quarters x NJ column groups (NJ = 4 is what is dispatched; see launch_tc_auto)
__host__ __device__ constexpr int compute_warps(int NJ) { return 4 * NJ; }
__host__ __device__ constexpr int cta_threads(int NJ) { return (4 * NJ + 1) * 32; }
constexpr int kTM = 128;
constexpr int kMaxChunks = 4;

template <int D, int H, int NETS, int NJ>
struct Geom {
  static_assert(NJ == 4 || NJ == 2, "column groups");
  static_assert(D % 16 == 0 && D >= 16 && D <= 64, "state dim: 16, 32, 48, 64");
  static_assert(H % 64 == 0 && H >= 64 && H <= 256, "hidden width: 64, 128, 192, 256");
  static constexpr int NC = D / NJ;   // state columns per compute thread
  static_assert(NC % 4 == 0 && NC <= 16, "state columns per thread: 4, 8 or 16");
  static constexpr int SB = 4 / NJ;   // 16-column tanh blocks per thread and hidden chunk
  static constexpr int CH = H / 64;   // 64-wide hidden chunks per network
  static constexpr int NCHUNK = NETS * CH;
  static_assert(NCHUNK <= kMaxChunks, "too many hidden chunks");
  // layer-2 accumulators per network: NFM "main" ones (hi*hi products; chunk c goes to c % NFM) and one
  // for the hi*lo + lo*hi corrections.  The tensor core truncates the fp32 accumulator once per MMA, an
  // error of ~0.5 ulp(|accumulator|) each time and always towards zero: keeping the 2^-11-times-smaller
  // correction products away from the full-size sums, and halving the chain length, keeps that bias at
  // a few fp32 ulps.  The epilogue adds the partial sums in fp32 (round to nearest).
  // A single-chunk network (H = 64) has all of its K steps at hand at once: corrections first, then the
  // hi*hi products, in ONE accumulator (the ordering does what the separate accumulator does).
  static constexpr int NFM = CH >= 2 ? 2 : 1;
  static constexpr bool SPLIT_CORR = CH >= 2;
  static constexpr int FW = (NFM + (SPLIT_CORR ? 1 : 0)) * D;
  static constexpr int Z0 = 0, F0 = NETS * H, U0 = NETS * (H + FW);
  static constexpr int COLS = NETS * (H + FW + D);
  static_assert(COLS <= 512, "TMEM has 512 columns");
  static constexpr int ALLOC = COLS <= 32 ? 32 : COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
  static constexpr int MAT_BYTES = D * H * 2;          // one fp16 copy of one weight matrix
  static constexpr int NET_W_BYTES = 4 * MAT_BYTES;    // W1hi | W1lo | W2hi | W2lo
  static constexpr int W_BYTES = NETS * NET_W_BYTES;
  // shared memory: [barriers + tmem slot: 256 B][weights][b1 NETS*H][b2 NETS*D][sinv NETS*2 (+pad)][t grid]
  static constexpr int OFF_W = 256;
  static constexpr int OFF_B1 = OFF_W + W_BYTES;
  static constexpr int OFF_B2 = OFF_B1 + NETS * H * 4;
  static constexpr int OFF_SINV = OFF_B2 + NETS * D * 4;
  static constexpr int OFF_T = OFF_SINV + 16;
  static size_t bytes(int T) { return (size_t)OFF_T + 4 * (size_t)((T + 3) / 4) * 4; }
};

struct TcParams {
  xde_mlp_field_t f, g;
  const unsigned char *wbuf;  // prepared weights: Geom::W_BYTES, then float sinv[NETS][2]
  const float *y0, *t_span;
  BmSource bm;
  float *out;
  long long B;
  int T, stride, n_out;
  int *status;  // device word or nullptr: XDE_ST_TC_RANGE when a stage input left the fp16 range
};

// ---- optional phase trace (debug builds only: -DXDE_TC_TRACE, tools/tc_trace.py) ------------------------
#ifdef XDE_TC_TRACE
__device__ long long *g_trace = nullptr;  // [2][kTraceCap]: row 0 = compute warp 0, row 1 = MMA thread (CTA 0)
constexpr int kTraceCap = 4096;
#define XDE_TRACE(row, tag)                                                              \
  do {                                                                                   \
    if (g_trace && blockIdx.x == 0 && trace_n < kTraceCap / 2) {                          \
      g_trace[(row) * kTraceCap + 2 * trace_n] = (tag);                                  \
      g_trace[(row) * kTraceCap + 2 * trace_n + 1] = clock64();                          \
      ++trace_n;                                                                         \
    }                                                                                    \
  } while (0)
#else
#define XDE_TRACE(row, tag) \
  do {                      \
  } while (0)
#endif

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// completion of every MMA this thread has issued so far -> one arrival on the mbarrier
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// one lane of a CONVERGED warp (the MMA warp runs its loop with all 32 lanes; only the elected lane issues:
// inside a divergent `if (lane == 0)` ptxas wraps every UTCHMMA in an ELECT/BRA.U.ANY waterfall loop,
// ~65 cycles per MMA instead of the ~34 the tensor pipe needs)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[tmem, fp16 packed] * B[smem descriptor], M = 128, K = 16
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d),
      "r"(a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; LBO = byte step between core matrices
// along K, SBO = byte step between 8-row groups along N (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46);
}
// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t instr_desc(int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
}

template <int N>
struct Tmem;
template <>
struct Tmem<2> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&r)[2]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(a), "r"(r[0]), "r"(r[1]) : "memory");
  }
};
template <>
struct Tmem<4> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a)
                 : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3])
                 : "memory");
  }
};
template <>
struct Tmem<8> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(a)
                 : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  }
};
template <>
struct Tmem<16> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(a)
        : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
  }
};

// x0, x1 -> packed fp16 hi pieces and packed fp16 lo pieces (element 0 in the low half).  The residual is one
// packed subtraction: the split is a quarter of the epilogue's instructions.
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  float l0, l1;
  upk(sub2(pk(x0, x1), pk(hf.x, hf.y)), l0, l1);
  const __half2 l = __floats2half2_rn(l0, l1);
  hi = *reinterpret_cast<const uint32_t *>(&h);
  lo = *reinterpret_cast<const uint32_t *>(&l);
}

// tanh on a packed pair for the tolerance path: the same 13/6 rational as tanh_rat2 (xde_common.cuh), but
// the quotient is p * rcp.approx(q) (~2 ulp instead of correctly rounded) and the |a| < 4e-4 -> a select
// is dropped (there the rational is a * (1 - 1.3e-7)).  10 packed FMA-pipe ops + 2 MUFU + 4 FMNMX per
// pair instead of 17 + 2 + 4 + 4 (FSETP/FSEL): the epilogue is issue-bound, so this is ~20 % of the step.
__device__ __forceinline__ f32x2 tanh_fast2(f32x2 a) {
  const float c = 7.90531110763549805f;
  float a0, a1;
  upk(a, a0, a1);
  const f32x2 x = pk(fminf(fmaxf(a0, -c), c), fminf(fmaxf(a1, -c), c));
  const f32x2 x2 = mul2(x, x);
  f32x2 p = fma2(x2, pk1(-2.76076847742355e-16f), pk1(2.00018790482477e-13f));
  p = fma2(x2, p, pk1(-8.60467152213735e-11f));
  p = fma2(x2, p, pk1(5.12229709037114e-08f));
  p = fma2(x2, p, pk1(1.48572235717979e-05f));
  p = fma2(x2, p, pk1(6.37261928875436e-04f));
  p = fma2(x2, p, pk1(4.89352455891786e-03f));
  p = mul2(x, p);
  f32x2 q = fma2(x2, pk1(1.19825839466702e-06f), pk1(1.18534705686654e-04f));
  q = fma2(x2, q, pk1(2.26843463243900e-03f));
  q = fma2(x2, q, pk1(4.89352518554385e-03f));
  float q0, q1;
  upk(q, q0, q1);
  return mul2(p, pk(rcp_approx(q0), rcp_approx(q1)));
}

// tanh on a packed pair through the SFU: 1 - 2 / (2^(2 a log2 e) + 1) -- 3 packed FMA-pipe ops + 4 MUFU
// (ex2, rcp per element) instead of 13 + 2.  Absolute error ~2e-7 (relative accuracy is lost for |a| << 1,
// where the rational keeps it).  The epilogue mixes the two forms so that the FMA and XU pipes are both busy:
// XDE_TC_EXP_MASK selects, per thread, which of its 8 pairs per 16-column block take this form.
__device__ __forceinline__ f32x2 tanh_sfu2(f32x2 a) {
  float t0, t1;
  upk(mul2(a, pk1(2.885390081777927f)), t0, t1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
  float d0, d1;
  upk(add2(pk(e0, e1), pk1(1.0f)), d0, d1);
  return fma2(pk(rcp_approx(d0), rcp_approx(d1)), pk1(-2.0f), pk1(1.0f));
}
// Measured on B200 (cfg3 / cfg4, ms): none 12.6 / 4.17, 3 of 8 11.3 / 3.75, 4 of 8 10.9 / 3.61, 5 of 8
// 10.9 / 3.55, 6 of 8 10.8 / 3.57, all 11.5 / 3.69; error against fp64 unchanged (0.6-1.9 x the FP32 kernels').
#ifndef XDE_TC_EXP_MASK
#define XDE_TC_EXP_MASK 0xB5
#endif
// m = index of the pair inside the thread's 16-column block (a constant after unrolling)
__device__ __forceinline__ f32x2 tanh_mixed2(f32x2 a, int m) {
  return ((XDE_TC_EXP_MASK >> m) & 1) ? tanh_sfu2(a) : tanh_fast2(a);
}

__device__ __forceinline__ f32x2 pre_rt2(int pre, f32x2 y) {  // the same on a packed pair
  if (pre == XDE_PRE_CUBE) return mul2(mul2(y, y), y);
  if (pre == XDE_PRE_SQUARE) return mul2(y, y);
  return y;
}
__device__ __forceinline__ float pre_rt(int pre, float y) {
  if (pre == XDE_PRE_CUBE) return (y * y) * y;
  if (pre == XDE_PRE_SQUARE) return y * y;
  return y;
}

// ---- weight preparation ---------------------------------------------------------------------------
// One CTA per (network, layer): power-of-two scale from max|W|, then the scaled matrix as fp16 hi / lo
// in the K-major core-matrix order the MMA's B descriptor walks:
//   layer 1 (B rows = hidden unit j, K = input i):  half index (i/8)*(H*8) + j*8 + i%8,  value W1[i*H + j]
//   layer 2 (B rows = output d,      K = hidden j): half index (j/8)*(D*8) + d*8 + j%8,  value W2[j*D + d]
__global__ void __launch_bounds__(1024) tc_prep_kernel(xde_mlp_field_t f, xde_mlp_field_t g, unsigned char *wbuf,
                                                       int nets) {
  const int net = blockIdx.x >> 1, layer = blockIdx.x & 1;
  const xde_mlp_field_t &fld = net ? g : f;
  const int D = fld.d, H = fld.h, n = D * H;
  const float *W = layer ? fld.w2 : fld.w1;
  __shared__ float red[32];
  float m = 0.0f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = fabsf(W[i]);
    if (a > m && a < INFINITY) m = a;
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(XDE_FULL_MASK, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = 0.0f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  int e = 0;
  if (m > 0.0f) e = 13 - ilogbf(m);  // max|W| * 2^e in [2^13, 2^14): hi never overflows, lo stays normal
  e = max(-100, min(100, e));
  const float scale = ldexpf(1.0f, e);
  const size_t mat = (size_t)n * 2;
  __half *hi = reinterpret_cast<__half *>(wbuf + (size_t)net * 4 * mat + (size_t)layer * 2 * mat);
  __half *lo = reinterpret_cast<__half *>(wbuf + (size_t)net * 4 * mat + (size_t)layer * 2 * mat + mat);
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    int k, r, R;  // K index, B-row index, number of B rows
    if (layer == 0) {
      k = idx / H;
      r = idx % H;
      R = H;
    } else {
      k = idx / D;
      r = idx % D;
      R = D;
    }
    const float v = W[idx] * scale;
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const int dst = (k >> 3) * (R * 8) + r * 8 + (k & 7);
    hi[dst] = h;
    lo[dst] = l;
  }
  if (threadIdx.x == 0) {
    float *sinv = reinterpret_cast<float *>(wbuf + (size_t)nets * 4 * mat);
    sinv[net * 2 + layer] = ldexpf(1.0f, -e);
  }
}

// ---- the solver -----------------------------------------------------------------------------------
// KIND 0: ODE Euler, 1: ODE RK4 (3/8 rule), 2: SDE Euler-Maruyama (two networks; increments from the caller's
// table), 3: ODE Midpoint, 4: SDE Euler-Maruyama with the increments generated in the kernel (BmSource)
template <int D, int H, int KIND, int NJ>
__global__ void __launch_bounds__(cta_threads(NJ), NJ == 4 ? 1 : 2) fixed_tc_kernel(const TcParams p) {
  constexpr int NETS = (KIND == 2 || KIND == 4) ? 2 : 1;
  constexpr int kComputeWarps = compute_warps(NJ), kThreads = cta_threads(NJ);
  using G = Geom<D, H, NETS, NJ>;
  constexpr int NC = G::NC, CH = G::CH, NCHUNK = G::NCHUNK;
  constexpr int EVALS = (KIND == 1) ? 4 : (KIND == 3) ? 2 : 1;  // field evaluations per step

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *u_ready = reinterpret_cast<uint64_t *>(smem);
  uint64_t *f_ready = u_ready + 1;
  uint64_t *z_ready = u_ready + 2;
  uint64_t *h_ready = z_ready + kMaxChunks;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(h_ready + kMaxChunks);
  unsigned char *sW = smem + G::OFF_W;
  const float *sb1 = reinterpret_cast<const float *>(smem + G::OFF_B1);
  const float *sb2 = reinterpret_cast<const float *>(smem + G::OFF_B2);
  const float *ssinv = reinterpret_cast<const float *>(smem + G::OFF_SINV);
  const float *st = reinterpret_cast<const float *>(smem + G::OFF_T);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup ----
  if (warp == kComputeWarps) {
    if (lane == 0) {
      mbar_init(u_ready, kComputeWarps);
      mbar_init(f_ready, 1);
      for (int c = 0; c < kMaxChunks; ++c) {
        mbar_init(z_ready + c, 1);
        mbar_init(h_ready + c, kComputeWarps);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)G::ALLOC)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.wbuf);
    uint4 *dst = reinterpret_cast<uint4 *>(sW);
    for (int i = tid; i < G::W_BYTES / 16; i += kThreads) dst[i] = src[i];
    float *b1w = reinterpret_cast<float *>(smem + G::OFF_B1);
    float *b2w = reinterpret_cast<float *>(smem + G::OFF_B2);
    float *siw = reinterpret_cast<float *>(smem + G::OFF_SINV);
    float *tw = reinterpret_cast<float *>(smem + G::OFF_T);
    for (int i = tid; i < NETS * H; i += kThreads) b1w[i] = (i < H ? p.f.b1[i] : p.g.b1[i - H]);
    for (int i = tid; i < NETS * D; i += kThreads) b2w[i] = (i < D ? p.f.b2[i] : p.g.b2[i - D]);
    if (tid < NETS * 2) siw[tid] = reinterpret_cast<const float *>(p.wbuf + G::W_BYTES)[tid];
    for (int i = tid; i < p.T; i += kThreads) tw[i] = p.t_span[i];
  }
  // the MMA unit reads shared memory through the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const long long n_tiles = (p.B + kTM - 1) / kTM;
  long long my_tiles = 0;
  if ((long long)blockIdx.x < n_tiles) my_tiles = (n_tiles - 1 - blockIdx.x) / gridDim.x + 1;
  const long long total_evals = my_tiles * (long long)(p.T - 1) * EVALS;

  if (warp == kComputeWarps) {
    // =============================== MMA issuer (one elected lane of a converged warp) ===============
    constexpr uint32_t idesc1 = instr_desc(64), idesc2 = instr_desc(D);
    const uint32_t w_addr = smem_u32(sW);
    [[maybe_unused]] int trace_n = (lane == 0) ? 0 : (1 << 30);
    for (long long ev = 0; ev < total_evals; ++ev) {
      const uint32_t par = (uint32_t)(ev & 1);
      mbar_wait(u_ready, par);
      tc_fence_after();
      XDE_TRACE(1, 100);
      // layer 1, chunk by chunk: Z[:, 64 c ..] = U W1[:, 64 c ..]
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const int net = c / CH, cc = c % CH;
        const uint32_t d = tmem + G::Z0 + net * H + cc * 64;
        const uint32_t w1hi = w_addr + net * G::NET_W_BYTES, w1lo = w1hi + G::MAT_BYTES;
        if (elect_one()) {
          // corrections first, while the accumulator is still small (see Geom::NFM), then the hi*hi products
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t a_hi = tmem + G::U0 + net * D + 8 * ks, a_lo = a_hi + D / 2;
            const uint32_t off = (2 * ks) * (H * 16) + (cc * 64) * 16;
            mma_ts(d, a_hi, smem_desc(w1lo + off, H * 16, 128), idesc1, ks > 0);
            mma_ts(d, a_lo, smem_desc(w1hi + off, H * 16, 128), idesc1, 1);
          }
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t a_hi = tmem + G::U0 + net * D + 8 * ks;
            const uint32_t off = (2 * ks) * (H * 16) + (cc * 64) * 16;
            mma_ts(d, a_hi, smem_desc(w1hi + off, H * 16, 128), idesc1, 1);
          }
          tc_commit(z_ready + c);
        }
        __syncwarp();
        XDE_TRACE(1, 110 + c);
      }
      // layer 2, as each chunk of tanh lands: F += Hh[:, 64 c ..] W2[64 c .., :]
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const int net = c / CH, cc = c % CH;
        const uint32_t d_main = tmem + G::F0 + net * G::FW + (cc % G::NFM) * D;
        const uint32_t d_corr = G::SPLIT_CORR ? tmem + G::F0 + net * G::FW + G::NFM * D : d_main;
        const uint32_t w2hi = w_addr + net * G::NET_W_BYTES + 2 * G::MAT_BYTES, w2lo = w2hi + G::MAT_BYTES;
        mbar_wait(h_ready + c, par);
        tc_fence_after();
        XDE_TRACE(1, 120 + c);
        if (elect_one()) {
          if (G::SPLIT_CORR) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const int s = cc * 4 + ks;  // K step = hidden units 16 s .. 16 s + 15
              const uint32_t a_hi = tmem + G::Z0 + net * H + 16 * s, a_lo = a_hi + 8;
              const uint32_t off = (2 * s) * (D * 16);
              const uint64_t b_hi = smem_desc(w2hi + off, D * 16, 128), b_lo = smem_desc(w2lo + off, D * 16, 128);
              mma_ts(d_corr, a_hi, b_lo, idesc2, (cc | ks) != 0);
              mma_ts(d_corr, a_lo, b_hi, idesc2, 1);
              mma_ts(d_main, a_hi, b_hi, idesc2, !(cc < G::NFM && ks == 0));
            }
          } else {  // one chunk, one accumulator: corrections first, then the hi*hi products
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t a_hi = tmem + G::Z0 + net * H + 16 * ks, a_lo = a_hi + 8;
              const uint32_t off = (2 * ks) * (D * 16);
              mma_ts(d_main, a_hi, smem_desc(w2lo + off, D * 16, 128), idesc2, ks != 0);
              mma_ts(d_main, a_lo, smem_desc(w2hi + off, D * 16, 128), idesc2, 1);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ts(d_main, tmem + G::Z0 + net * H + 16 * ks, smem_desc(w2hi + (2 * ks) * (D * 16), D * 16, 128),
                     idesc2, 1);
          }
          if (c == NCHUNK - 1) tc_commit(f_ready);
        }
        __syncwarp();
      }
      XDE_TRACE(1, 130);
    }
  } else {
    // =============================== compute warps ===============================
    const int q = warp & 3, j = warp >> 2;  // TMEM lane quarter, column group
    const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
    const int c0 = j * NC;  // first state column of this thread
    const int pref = p.f.pre, preg = p.g.pre;
    const float one_third = (float)(1.0 / 3.0);
    uint32_t par = 0;
    bool in_range = true;  // every stage input representable in fp16 (|pre(y)| < 65504, finite)
    [[maybe_unused]] int trace_n = (tid == 0) ? 0 : (1 << 30);

    // one evaluation of the field(s) at yi: kf (and kg) <- f(yi) (, g(yi)); state columns travel as packed pairs
    constexpr int NP = NC / 2;
    auto eval = [&](const f32x2 (&yi)[NP], f32x2 (&kf)[NP], f32x2 (&kg)[NP]) {
      XDE_TRACE(0, 0);
      // 1. stage input -> U (fp16 hi | lo)
#pragma unroll
      for (int net = 0; net < NETS; ++net) {
        uint32_t uh[NC / 2], ul[NC / 2];
        const int pre = net ? preg : pref;
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          float v0, v1;
          upk(pre_rt2(pre, yi[c]), v0, v1);
          in_range = in_range && (fabsf(v0) < 65504.0f) && (fabsf(v1) < 65504.0f);  // also false for NaN
          split2(v0, v1, uh[c], ul[c]);
        }
        Tmem<NC / 2>::st(tl + G::U0 + net * D + j * (NC / 2), uh);
        Tmem<NC / 2>::st(tl + G::U0 + net * D + D / 2 + j * (NC / 2), ul);
      }
      XDE_TRACE(0, 3);
      tc_wait_st();
      XDE_TRACE(0, 4);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(u_ready);
      XDE_TRACE(0, 1);
      // 2. hidden chunks: Z -> tanh -> fp16 hi | lo, in place.  The TMEM load of chunk c+1 is issued half
      // way through the tanh of chunk c (its MMAs were queued right behind chunk c's), so its latency and
      // the mbarrier round trip hide behind arithmetic.  (Deferring the store-completion wait / hand-off
      // of chunk c into chunk c+1 was measured: no gain -- the phase is issue-bound, the other warps of
      // the scheduler already cover those latencies.)
      constexpr int SB = G::SB, NB = NCHUNK * SB;  // 16-column blocks of this thread: SB per chunk
      auto zcol = [&](int blk) {
        const int c = blk / SB, sb = blk % SB;
        return (c / CH) * H + (c % CH) * 64 + (j * SB + sb) * 16;  // first hidden unit (= Z column) of the block
      };
      uint32_t z[16];
      mbar_wait(z_ready, par);
      tc_fence_after();
      Tmem<16>::ld(tl + G::Z0 + zcol(0), z);
      XDE_TRACE(0, 2);
#pragma unroll
      for (int blk = 0; blk < NB; ++blk) {
        const int c = blk / SB, net = c / CH;
        const int h0 = zcol(blk);
        const f32x2 s1 = pk1(ssinv[net * 2]);
        uint32_t o[16], zn[16];
        tc_wait_ld();
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          if (m == 4 && blk + 1 < NB) {
            if ((blk + 1) % SB == 0) {
              mbar_wait(z_ready + (blk + 1) / SB, par);
              tc_fence_after();
            }
            Tmem<16>::ld(tl + G::Z0 + zcol(blk + 1), zn);
          }
          const float2 b = *reinterpret_cast<const float2 *>(sb1 + h0 + 2 * m);
          const f32x2 a = fma2(pk(__uint_as_float(z[2 * m]), __uint_as_float(z[2 * m + 1])), s1, pk(b.x, b.y));
          float t0, t1;
          upk(tanh_mixed2(a, m), t0, t1);
          split2(t0, t1, o[m], o[8 + m]);
        }
        Tmem<16>::st(tl + G::Z0 + h0, o);
        XDE_TRACE(0, 30 + c);
        if (blk % SB == SB - 1) {
          tc_wait_st();
          XDE_TRACE(0, 40 + c);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(h_ready + c);
          XDE_TRACE(0, 10 + c);
        }
        if (blk + 1 < NB) {
#pragma unroll
          for (int m = 0; m < 16; ++m) z[m] = zn[m];
        }
      }
      // 3. F -> registers
      mbar_wait(f_ready, par);
      tc_fence_after();
      XDE_TRACE(0, 20);
#pragma unroll
      for (int net = 0; net < NETS; ++net) {
        f32x2(&kk)[NP] = net ? kg : kf;
        const uint32_t fa = tl + G::F0 + net * G::FW + c0;
        uint32_t r0[NC], r1[NC];
        Tmem<NC>::ld(fa, r0);
        if (G::SPLIT_CORR) Tmem<NC>::ld(fa + D, r1);  // second main accumulator
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          kk[c] = pk(__uint_as_float(r0[2 * c]), __uint_as_float(r0[2 * c + 1]));
          if (G::SPLIT_CORR) kk[c] = add2(kk[c], pk(__uint_as_float(r1[2 * c]), __uint_as_float(r1[2 * c + 1])));
        }
        if (G::SPLIT_CORR) {
          Tmem<NC>::ld(fa + 2 * D, r0);  // corrections
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < NP; ++c)
            kk[c] = add2(kk[c], pk(__uint_as_float(r0[2 * c]), __uint_as_float(r0[2 * c + 1])));
        }
        const f32x2 s2 = pk1(ssinv[net * 2 + 1]);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          const float2 bb = *reinterpret_cast<const float2 *>(sb2 + net * D + c0 + 2 * c);
          kk[c] = fma2(kk[c], s2, pk(bb.x, bb.y));
        }
      }
      par ^= 1u;
      XDE_TRACE(0, 21);
    };

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long b = tile * kTM + 32 * q + lane;
      const bool ok = b < p.B;
      f32x2 y[NP];
#pragma unroll
      for (int v = 0; v < NC / 4; ++v) {
        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
          t4 = *reinterpret_cast<const float4 *>(p.y0 + b * D + c0 + 4 * v);
          *reinterpret_cast<float4 *>(p.out + b * (long long)p.n_out * D + c0 + 4 * v) = t4;
        }
        y[2 * v] = pk(t4.x, t4.y);
        y[2 * v + 1] = pk(t4.z, t4.w);
      }
      for (int i = 1; i < p.T; ++i) {
        const float dt = st[i] - st[i - 1];
        const f32x2 dt2 = pk1(dt);
        f32x2 k[NP], kg[NP];
        if (KIND == 0) {
          eval(y, k, kg);
#pragma unroll
          for (int c = 0; c < NP; ++c) y[c] = fma2(k[c], dt2, y[c]);
        } else if (KIND == 3) {
          // Midpoint.step (fixed_solver/midpoint.py:7-18)
          f32x2 yi[NP];
          eval(y, k, kg);
          const f32x2 half_dt = pk1(0.5f * dt);
#pragma unroll
          for (int c = 0; c < NP; ++c) yi[c] = fma2(k[c], half_dt, y[c]);
          eval(yi, k, kg);
#pragma unroll
          for (int c = 0; c < NP; ++c) y[c] = fma2(k[c], dt2, y[c]);
        } else if (KIND == 1) {
          // RK4.step = rk4_alt_step_func (base_fixed_solver.py:166-197), as written there:
          //   k2 = f(y + dt k1/3), k3 = f(y + dt (k1 - k2/3)), k4 = f(y + dt (k1 - k2 + k3)),
          //   y1 = y + dt (k1 + 3 k2 + 3 k3 + k4) / 8
          f32x2 yi[NP], A[NP], S[NP];
          eval(y, k, kg);
          const f32x2 dt13 = pk1(dt * one_third);
#pragma unroll
          for (int c = 0; c < NP; ++c) {
            A[c] = k[c];
            yi[c] = fma2(k[c], dt13, y[c]);
          }
          eval(yi, k, kg);
#pragma unroll
          for (int c = 0; c < NP; ++c) {
            yi[c] = fma2(fma2(pk1(-one_third), k[c], A[c]), dt2, y[c]);
            S[c] = fma2(pk1(3.0f), k[c], A[c]);
            A[c] = sub2(A[c], k[c]);
          }
          eval(yi, k, kg);
#pragma unroll
          for (int c = 0; c < NP; ++c) {
            yi[c] = fma2(add2(A[c], k[c]), dt2, y[c]);
            S[c] = fma2(pk1(3.0f), k[c], S[c]);
          }
          eval(yi, k, kg);
          const f32x2 dt8 = pk1(dt * 0.125f);
#pragma unroll
          for (int c = 0; c < NP; ++c) y[c] = fma2(add2(S[c], k[c]), dt8, y[c]);
        } else {
          // the increments are requested before the evaluation and first touched after it (packing them
          // here would stall on the load before the evaluation starts)
          float4 w4[NC / 4];
#pragma unroll
          for (int v = 0; v < NC / 4; ++v) {
            w4[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) w4[v] = bm_increment4<KIND == 4>(p.bm, i - 1, b, p.B, D, c0 / 4 + v, dt);
          }
          eval(y, k, kg);
#pragma unroll
          for (int v = 0; v < NC / 4; ++v) {
            y[2 * v] = fma2(kg[2 * v], pk(w4[v].x, w4[v].y), fma2(k[2 * v], dt2, y[2 * v]));
            y[2 * v + 1] = fma2(kg[2 * v + 1], pk(w4[v].z, w4[v].w), fma2(k[2 * v + 1], dt2, y[2 * v + 1]));
          }
        }
        // linear_interp at t == t1 is the identity (interpolation/functional/interp_fn.py:4-10)
        if (ok && (i % p.stride == 0 || i == p.T - 1)) {
          const int row = (i == p.T - 1) ? p.n_out - 1 : i / p.stride;
#pragma unroll
          for (int v = 0; v < NC / 4; ++v) {
            float4 t4;
            upk(y[2 * v], t4.x, t4.y);
            upk(y[2 * v + 1], t4.z, t4.w);
            *reinterpret_cast<float4 *>(p.out + (b * (long long)p.n_out + row) * D + c0 + 4 * v) = t4;
          }
        }
      }
    }
    // a stage input outside the fp16 range (or non-finite) makes the MMA operands inf / NaN: report, the caller
    // discards the result (the shim reruns on the FP32 kernels when math="auto")
    if (p.status && !__all_sync(XDE_FULL_MASK, in_range) && lane == 0) atomicMax(p.status, XDE_ST_TC_RANGE);
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == kComputeWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)G::ALLOC)
                 : "memory");
  }
}

// ---- two tiles in flight (small fields) ----------------------------------------------------------------
// For H = 64 networks one tile needs <= 256 TMEM columns and the tensor-core work per evaluation is tiny:
// in fixed_tc_kernel almost half of an evaluation is then exposed latency (first layer-1 chunk, last layer-2
// chunk, mbarrier / commit / TMEM round trips; phase trace of cfg4: ~3 000 of ~6 300 cycles).  Two 9-warp CTAs
// per SM do not co-reside (register file granularity), so this kernel keeps TWO tiles (slots A, B; 256 TMEM
// columns and one set of mbarriers each) in flight inside one CTA and software-pipelines them:
//     compute warps:  H(A) H(B) | F(A) update(A) U(A) | F(B) update(B) U(B) | H(A) H(B) | ...
//     MMA warp:       L1(A) L1(B) | L2(A) | L2(B) | L1(A) L1(B) | ...
// U(x) = stage input to TMEM, L1/L2 = the layers' MMAs, H(x) = tanh epilogue, F(x) = read the field value.
// Whenever the compute warps wait on a barrier of one slot, the MMAs it stands for were issued a whole phase
// of the other slot earlier.  Same arithmetic, same TMEM layouts per slot as fixed_tc_kernel.
template <int D, int H, int KIND>
__global__ void __launch_bounds__(cta_threads(4), 1) fixed_tc2_kernel(const TcParams p) {
  constexpr int NETS = (KIND == 2 || KIND == 4) ? 2 : 1;
  constexpr int kComputeWarps = compute_warps(4), kThreads = cta_threads(4);
  using G = Geom<D, H, NETS, 4>;
  static_assert(G::COLS <= 256 && G::NC <= 8, "two tiles in flight: <= 256 TMEM columns and <= 8 state columns per thread");
  constexpr int NC = G::NC, NP = NC / 2, CH = G::CH, NCHUNK = G::NCHUNK;
  constexpr int EVALS = (KIND == 1) ? 4 : (KIND == 3) ? 2 : 1;
  constexpr uint32_t SLOT = 256;  // TMEM columns per tile slot

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *u_ready = reinterpret_cast<uint64_t *>(smem);  // [2]
  uint64_t *f_ready = u_ready + 2;                           // [2]
  uint64_t *z_ready = u_ready + 4;                           // [2][kMaxChunks]
  uint64_t *h_ready = z_ready + 2 * kMaxChunks;              // [2][kMaxChunks]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(h_ready + 2 * kMaxChunks);
  static_assert((4 + 4 * kMaxChunks) * 8 + 4 <= G::OFF_W, "barrier block");
  unsigned char *sW = smem + G::OFF_W;
  const float *sb1 = reinterpret_cast<const float *>(smem + G::OFF_B1);
  const float *sb2 = reinterpret_cast<const float *>(smem + G::OFF_B2);
  const float *ssinv = reinterpret_cast<const float *>(smem + G::OFF_SINV);
  const float *st = reinterpret_cast<const float *>(smem + G::OFF_T);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == kComputeWarps) {
    if (lane == 0) {
      for (int s = 0; s < 2; ++s) {
        mbar_init(u_ready + s, kComputeWarps);
        mbar_init(f_ready + s, 1);
        for (int c = 0; c < kMaxChunks; ++c) {
          mbar_init(z_ready + s * kMaxChunks + c, 1);
          mbar_init(h_ready + s * kMaxChunks + c, kComputeWarps);
        }
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(2 * SLOT)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.wbuf);
    uint4 *dst = reinterpret_cast<uint4 *>(sW);
    for (int i = tid; i < G::W_BYTES / 16; i += kThreads) dst[i] = src[i];
    float *b1w = reinterpret_cast<float *>(smem + G::OFF_B1);
    float *b2w = reinterpret_cast<float *>(smem + G::OFF_B2);
    float *siw = reinterpret_cast<float *>(smem + G::OFF_SINV);
    float *tw = reinterpret_cast<float *>(smem + G::OFF_T);
    for (int i = tid; i < NETS * H; i += kThreads) b1w[i] = (i < H ? p.f.b1[i] : p.g.b1[i - H]);
    for (int i = tid; i < NETS * D; i += kThreads) b2w[i] = (i < D ? p.f.b2[i] : p.g.b2[i - D]);
    if (tid < NETS * 2) siw[tid] = reinterpret_cast<const float *>(p.wbuf + G::W_BYTES)[tid];
    for (int i = tid; i < p.T; i += kThreads) tw[i] = p.t_span[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const long long n_tiles = (p.B + kTM - 1) / kTM;
  const long long stride = 2LL * gridDim.x;  // a CTA takes tiles (t, t + gridDim.x) together
  long long my_pairs = 0;
  if ((long long)blockIdx.x < n_tiles) my_pairs = (n_tiles - 1 - blockIdx.x) / stride + 1;
  const long long rounds = my_pairs * (long long)(p.T - 1) * EVALS;

  if (warp == kComputeWarps) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc1 = instr_desc(64), idesc2 = instr_desc(D);
    const uint32_t w_addr = smem_u32(sW);
    for (long long r = 0; r < rounds; ++r) {
      const uint32_t par = (uint32_t)(r & 1);
#pragma unroll
      for (int s = 0; s < 2; ++s) {  // layer 1 of both slots
        const uint32_t tm = tmem + s * SLOT;
        mbar_wait(u_ready + s, par);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          const int net = c / CH, cc = c % CH;
          const uint32_t d = tm + G::Z0 + net * H + cc * 64;
          const uint32_t w1hi = w_addr + net * G::NET_W_BYTES, w1lo = w1hi + G::MAT_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
              const uint32_t a_hi = tm + G::U0 + net * D + 8 * ks, a_lo = a_hi + D / 2;
              const uint32_t off = (2 * ks) * (H * 16) + (cc * 64) * 16;
              mma_ts(d, a_hi, smem_desc(w1lo + off, H * 16, 128), idesc1, ks > 0);
              mma_ts(d, a_lo, smem_desc(w1hi + off, H * 16, 128), idesc1, 1);
            }
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
              const uint32_t a_hi = tm + G::U0 + net * D + 8 * ks;
              const uint32_t off = (2 * ks) * (H * 16) + (cc * 64) * 16;
              mma_ts(d, a_hi, smem_desc(w1hi + off, H * 16, 128), idesc1, 1);
            }
            tc_commit(z_ready + s * kMaxChunks + c);
          }
          __syncwarp();
        }
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {  // layer 2 of both slots, as their tanh chunks land
        const uint32_t tm = tmem + s * SLOT;
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          const int net = c / CH, cc = c % CH;
          const uint32_t d_main = tm + G::F0 + net * G::FW + (cc % G::NFM) * D;
          const uint32_t d_corr = G::SPLIT_CORR ? tm + G::F0 + net * G::FW + G::NFM * D : d_main;
          const uint32_t w2hi = w_addr + net * G::NET_W_BYTES + 2 * G::MAT_BYTES, w2lo = w2hi + G::MAT_BYTES;
          mbar_wait(h_ready + s * kMaxChunks + c, par);
          tc_fence_after();
          if (elect_one()) {
            if (G::SPLIT_CORR) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const int kk = cc * 4 + ks;
                const uint32_t a_hi = tm + G::Z0 + net * H + 16 * kk, a_lo = a_hi + 8;
                const uint32_t off = (2 * kk) * (D * 16);
                mma_ts(d_corr, a_hi, smem_desc(w2lo + off, D * 16, 128), idesc2, (cc | ks) != 0);
                mma_ts(d_corr, a_lo, smem_desc(w2hi + off, D * 16, 128), idesc2, 1);
                mma_ts(d_main, a_hi, smem_desc(w2hi + off, D * 16, 128), idesc2, !(cc < G::NFM && ks == 0));
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t a_hi = tm + G::Z0 + net * H + 16 * ks, a_lo = a_hi + 8;
                const uint32_t off = (2 * ks) * (D * 16);
                mma_ts(d_main, a_hi, smem_desc(w2lo + off, D * 16, 128), idesc2, ks != 0);
                mma_ts(d_main, a_lo, smem_desc(w2hi + off, D * 16, 128), idesc2, 1);
              }
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                mma_ts(d_main, tm + G::Z0 + net * H + 16 * ks, smem_desc(w2hi + (2 * ks) * (D * 16), D * 16, 128),
                       idesc2, 1);
            }
            if (c == NCHUNK - 1) tc_commit(f_ready + s);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =============================== compute warps ===============================
    const int q = warp & 3, j = warp >> 2;
    const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
    const int c0 = j * NC;
    const int pref = p.f.pre, preg = p.g.pre;
    const float one_third = (float)(1.0 / 3.0);
    uint32_t par = 0;
    bool in_range = true;  // every stage input representable in fp16 (|pre(y)| < 65504, finite)
    [[maybe_unused]] int trace_n = (tid == 0) ? 0 : (1 << 30);

    auto phaseU_issue = [&](int s, const f32x2(&yi)[NP]) {  // stage input -> U of slot s (fp16 hi | lo)
      const uint32_t ts = tl + s * SLOT;
#pragma unroll
      for (int net = 0; net < NETS; ++net) {
        uint32_t uh[NP], ul[NP];
        const int pre = net ? preg : pref;
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          float v0, v1;
          upk(pre_rt2(pre, yi[c]), v0, v1);
          in_range = in_range && (fabsf(v0) < 65504.0f) && (fabsf(v1) < 65504.0f);  // also false for NaN
          split2(v0, v1, uh[c], ul[c]);
        }
        Tmem<NP>::st(ts + G::U0 + net * D + j * NP, uh);
        Tmem<NP>::st(ts + G::U0 + net * D + D / 2 + j * NP, ul);
      }
    };
    auto phaseU_finish = [&](int s) {  // stores complete -> hand the slot to the MMA warp
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(u_ready + s);
    };
    auto phaseH = [&](int s) {  // Z -> tanh -> fp16 hi | lo, in place, chunk by chunk
      const uint32_t ts = tl + s * SLOT;
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const int net = c / CH, cc = c % CH;
        const int h0 = net * H + cc * 64 + j * 16;
        mbar_wait(z_ready + s * kMaxChunks + c, par);
        tc_fence_after();
        uint32_t z[16], o[16];
        Tmem<16>::ld(ts + G::Z0 + h0, z);
        tc_wait_ld();
        const f32x2 s1 = pk1(ssinv[net * 2]);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const float2 b = *reinterpret_cast<const float2 *>(sb1 + h0 + 2 * m);
          const f32x2 a = fma2(pk(__uint_as_float(z[2 * m]), __uint_as_float(z[2 * m + 1])), s1, pk(b.x, b.y));
          float t0, t1;
          upk(tanh_mixed2(a, m), t0, t1);
          split2(t0, t1, o[m], o[8 + m]);
        }
        Tmem<16>::st(ts + G::Z0 + h0, o);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(h_ready + s * kMaxChunks + c);
      }
    };
    // F of slot s -> registers, in two halves so that the TMEM load of slot B can be in flight while slot A's
    // stage input is being stored (and the store's completion wait overlaps the load's)
    auto phaseF_issue = [&](int s, uint32_t(&r)[NETS][NC]) {
      const uint32_t ts = tl + s * SLOT;
      mbar_wait(f_ready + s, par);
      tc_fence_after();
#pragma unroll
      for (int net = 0; net < NETS; ++net) Tmem<NC>::ld(ts + G::F0 + net * G::FW + c0, r[net]);
    };
    auto phaseF_finish = [&](int s, uint32_t(&r)[NETS][NC], f32x2(&kf)[NP], f32x2(&kg)[NP]) {
      const uint32_t ts = tl + s * SLOT;
      tc_wait_ld();
#pragma unroll
      for (int net = 0; net < NETS; ++net) {
        f32x2(&kk)[NP] = net ? kg : kf;
#pragma unroll
        for (int c = 0; c < NP; ++c) kk[c] = pk(__uint_as_float(r[net][2 * c]), __uint_as_float(r[net][2 * c + 1]));
        if (G::SPLIT_CORR) {  // second main accumulator and the corrections
          const uint32_t fa = ts + G::F0 + net * G::FW + c0;
          uint32_t r1[NC], r2[NC];
          Tmem<NC>::ld(fa + D, r1);
          Tmem<NC>::ld(fa + 2 * D, r2);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < NP; ++c)
            kk[c] = add2(add2(kk[c], pk(__uint_as_float(r1[2 * c]), __uint_as_float(r1[2 * c + 1]))),
                         pk(__uint_as_float(r2[2 * c]), __uint_as_float(r2[2 * c + 1])));
        }
        const f32x2 s2 = pk1(ssinv[net * 2 + 1]);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          const float2 bb = *reinterpret_cast<const float2 *>(sb2 + net * D + c0 + 2 * c);
          kk[c] = fma2(kk[c], s2, pk(bb.x, bb.y));
        }
      }
    };

    for (long long t0 = blockIdx.x; t0 < n_tiles; t0 += stride) {
      long long b[2];
      bool ok[2];
      f32x2 y[2][NP], A[2][NP], S[2][NP];
      float4 w4[2][NC / 4];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const long long tile = t0 + (long long)s * gridDim.x;
        b[s] = tile * kTM + 32 * q + lane;
        ok[s] = tile < n_tiles && b[s] < p.B;  // slot B of the last pair may be empty: it runs on zeros
#pragma unroll
        for (int v = 0; v < NC / 4; ++v) {
          float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok[s]) {
            t4 = *reinterpret_cast<const float4 *>(p.y0 + b[s] * D + c0 + 4 * v);
            *reinterpret_cast<float4 *>(p.out + b[s] * (long long)p.n_out * D + c0 + 4 * v) = t4;
          }
          y[s][2 * v] = pk(t4.x, t4.y);
          y[s][2 * v + 1] = pk(t4.z, t4.w);
          w4[s][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((KIND == 2 || KIND == 4) && ok[s] && p.T > 1)  // increments of step 1
            w4[s][v] = bm_increment4<KIND == 4>(p.bm, 0, b[s], p.B, D, c0 / 4 + v, st[1] - st[0]);
        }
#pragma unroll
        for (int c = 0; c < NP; ++c) A[s][c] = S[s][c] = pk1(0.0f);
        phaseU_issue(s, y[s]);
        phaseU_finish(s);
      }
      for (int i = 1; i < p.T; ++i) {
        const float dt = st[i] - st[i - 1];
        const f32x2 dt2 = pk1(dt);
#pragma unroll
        for (int e = 0; e < EVALS; ++e) {
          XDE_TRACE(0, 200);
          phaseH(0);
          XDE_TRACE(0, 201);
          phaseH(1);
          XDE_TRACE(0, 202);
          uint32_t fr[NETS][NC];
          phaseF_issue(0, fr);
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            f32x2 k[NP], kg[NP], yi[NP];
            phaseF_finish(s, fr, k, kg);
            XDE_TRACE(0, 210 + s);
            bool stored = false;
            if (KIND == 0) {
#pragma unroll
              for (int c = 0; c < NP; ++c) y[s][c] = fma2(k[c], dt2, y[s][c]);
            } else if (KIND == 3) {  // Midpoint.step (fixed_solver/midpoint.py:7-18)
#pragma unroll
              for (int c = 0; c < NP; ++c) {
                if (e == 0) yi[c] = fma2(k[c], pk1(0.5f * dt), y[s][c]);
                else y[s][c] = fma2(k[c], dt2, y[s][c]);
              }
            } else if (KIND == 1) {  // rk4_alt_step_func (base_fixed_solver.py:166-197), see fixed_tc_kernel
#pragma unroll
              for (int c = 0; c < NP; ++c) {
                if (e == 0) {
                  A[s][c] = k[c];
                  yi[c] = fma2(k[c], pk1(dt * one_third), y[s][c]);
                } else if (e == 1) {
                  yi[c] = fma2(fma2(pk1(-one_third), k[c], A[s][c]), dt2, y[s][c]);
                  S[s][c] = fma2(pk1(3.0f), k[c], A[s][c]);
                  A[s][c] = sub2(A[s][c], k[c]);
                } else if (e == 2) {
                  yi[c] = fma2(add2(A[s][c], k[c]), dt2, y[s][c]);
                  S[s][c] = fma2(pk1(3.0f), k[c], S[s][c]);
                } else {
                  y[s][c] = fma2(add2(S[s][c], k[c]), pk1(dt * 0.125f), y[s][c]);
                }
              }
            } else {  // Euler-Maruyama
#pragma unroll
              for (int v = 0; v < NC / 4; ++v) {
                y[s][2 * v] = fma2(kg[2 * v], pk(w4[s][v].x, w4[s][v].y), fma2(k[2 * v], dt2, y[s][2 * v]));
                y[s][2 * v + 1] = fma2(kg[2 * v + 1], pk(w4[s][v].z, w4[s][v].w), fma2(k[2 * v + 1], dt2, y[s][2 * v + 1]));
              }
            }
            if (e == EVALS - 1) {
              if (ok[s] && (i % p.stride == 0 || i == p.T - 1)) {
                const int row = (i == p.T - 1) ? p.n_out - 1 : i / p.stride;
#pragma unroll
                for (int v = 0; v < NC / 4; ++v) {
                  float4 t4;
                  upk(y[s][2 * v], t4.x, t4.y);
                  upk(y[s][2 * v + 1], t4.z, t4.w);
                  *reinterpret_cast<float4 *>(p.out + (b[s] * (long long)p.n_out + row) * D + c0 + 4 * v) = t4;
                }
              }
              if (i < p.T - 1) {  // next step's increments (first touched after its evaluation), then its input
                if (KIND == 2 || KIND == 4) {
#pragma unroll
                  for (int v = 0; v < NC / 4; ++v)
                    if (ok[s])
                      w4[s][v] = bm_increment4<KIND == 4>(p.bm, i, b[s], p.B, D, c0 / 4 + v, st[i + 1] - st[i]);
                }
                phaseU_issue(s, y[s]);
                stored = true;
              }
            } else {
              phaseU_issue(s, yi);
              stored = true;
            }
            if (s == 0) phaseF_issue(1, fr);  // slot B's field value: its load flies while slot A's store completes
            if (stored) phaseU_finish(s);
            XDE_TRACE(0, 220 + s);
          }
          par ^= 1u;
        }
      }
    }
    // a stage input outside the fp16 range (or non-finite) makes the MMA operands inf / NaN: report, the caller
    // discards the result (the shim reruns on the FP32 kernels when math="auto")
    if (p.status && !__all_sync(XDE_FULL_MASK, in_range) && lane == 0) atomicMax(p.status, XDE_ST_TC_RANGE);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kComputeWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * SLOT) : "memory");
  }
}

template <int D, int H, int KIND>
static int launch_tc2(TcParams p, cudaStream_t s) {
  constexpr int NETS = (KIND == 2 || KIND == 4) ? 2 : 1;
  using G = Geom<D, H, NETS, 4>;
  const size_t smem = G::bytes(p.T);
  XDE_REQUIRE(smem <= 227 * 1024, XDE_E_UNSUPPORTED_FIELD,
              "tensor-core solver: weights + time grid need %zu bytes of shared memory (> 227 KB)", smem);
  void *wbuf = nullptr;
  XDE_CUDA_CHECK(scratch_alloc(&wbuf, (size_t)G::W_BYTES + 16, s));
  tc_prep_kernel<<<NETS * 2, 1024, 0, s>>>(p.f, p.g, (unsigned char *)wbuf, NETS);
  count_launch();
  p.wbuf = (const unsigned char *)wbuf;
  auto kern = fixed_tc2_kernel<D, H, KIND>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long n_tiles = (p.B + kTM - 1) / kTM;
  long long grid = sm_count();  // persistent: one CTA, two tile slots, all 512 TMEM columns per SM
  if (grid > (n_tiles + 1) / 2) grid = (n_tiles + 1) / 2;
  kern<<<(unsigned)grid, cta_threads(4), smem, s>>>(p);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  XDE_CUDA_CHECK(cudaFreeAsync(wbuf, s));
  return XDE_OK;
}

template <int D, int H, int KIND, int NJ>
static int launch_tc(TcParams p, cudaStream_t s) {
  constexpr int NETS = (KIND == 2 || KIND == 4) ? 2 : 1;
  using G = Geom<D, H, NETS, NJ>;
  const size_t smem = G::bytes(p.T);
  XDE_REQUIRE(smem <= 227 * 1024, XDE_E_UNSUPPORTED_FIELD,
              "tensor-core solver: weights + time grid need %zu bytes of shared memory (> 227 KB)", smem);
  void *wbuf = nullptr;
  const size_t wbytes = (size_t)G::W_BYTES + 16;
  XDE_CUDA_CHECK(scratch_alloc(&wbuf, wbytes, s));
  tc_prep_kernel<<<NETS * 2, 1024, 0, s>>>(p.f, p.g, (unsigned char *)wbuf, NETS);
  count_launch();
  p.wbuf = (const unsigned char *)wbuf;
  auto kern = fixed_tc_kernel<D, H, KIND, NJ>;
  XDE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  XDE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, cta_threads(NJ), smem));
  per_sm = std::max(1, std::min(per_sm, 512 / G::ALLOC));  // each resident CTA owns G::ALLOC TMEM columns
  const long long n_tiles = (p.B + kTM - 1) / kTM;
  long long grid = (long long)sm_count() * per_sm;  // persistent
  if (grid > n_tiles) grid = n_tiles;
  kern<<<(unsigned)grid, cta_threads(NJ), smem, s>>>(p);
  count_launch();
  XDE_CUDA_CHECK(cudaGetLastError());
  XDE_CUDA_CHECK(cudaFreeAsync(wbuf, s));
  return XDE_OK;
}

// NJ = 2 (8 compute warps, meant for two CTAs per SM on small fields) is kept compilable but not dispatched:
// measured on B200, the register file is allocated per 4 warps, so two 9-warp CTAs at > 80 registers per
// thread do not co-reside (launch__waves_per_multiprocessor stayed 0.5) and the variant is 4 % slower
// than one 17-warp CTA.  Small fields use fixed_tc2_kernel (two tiles in flight inside one CTA) instead,
// whenever every SM gets at least one PAIR of tiles.
template <int D, int H, int KIND>
static int launch_tc_auto(const TcParams &p, cudaStream_t s) {
  constexpr int NETS = (KIND == 2 || KIND == 4) ? 2 : 1;
  using G = Geom<D, H, NETS, 4>;
  if constexpr (G::COLS <= 256 && G::NC <= 8) {
    const long long n_tiles = (p.B + kTM - 1) / kTM;
    if (n_tiles >= 2LL * sm_count()) return launch_tc2<D, H, KIND>(p, s);
  }
  return launch_tc<D, H, KIND, 4>(p, s);
}

template <int KIND>
static int tc_dispatch(const TcParams &p, cudaStream_t s) {
  const int D = p.f.d, H = p.f.h;
#define XDE_TC_CASE(DD, HH) \
  if (D == DD && H == HH) return launch_tc_auto<DD, HH, KIND>(p, s);
  XDE_TC_CASE(64, 64) XDE_TC_CASE(32, 128) XDE_TC_CASE(32, 64) XDE_TC_CASE(16, 128) XDE_TC_CASE(16, 64)
  if constexpr (KIND != 2 && KIND != 4) {  // two networks: 2 (H + FW + D) <= 512 TMEM columns
    XDE_TC_CASE(64, 256) XDE_TC_CASE(64, 128) XDE_TC_CASE(32, 256)
  }
#undef XDE_TC_CASE
  set_last_error("tensor-core solver: no kernel for D=%d H=%d (D in {16,32,64} x H in {64,128,256}; sde: TMEM limits D=64 to H=64, others to H<=128)", D, H);
  return XDE_E_UNSUPPORTED_FIELD;
}

}  // namespace tc

#ifdef XDE_TC_TRACE
extern "C" XDE_EXPORT int xde_tc_trace_set(long long *buf) {
  return cudaMemcpyToSymbol(tc::g_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -3;
}
#endif

int rk_fixed_tc(int method, const xde_mlp_field_t *f, const float *y0, long long B, const float *t_span, int T,
                int stride, float *out, int *status, cudaStream_t s) {
  tc::TcParams p{};
  p.f = *f;
  p.g = *f;
  p.y0 = y0;
  p.t_span = t_span;
  p.out = out;
  p.B = B;
  p.T = T;
  p.stride = stride;
  p.n_out = (T - 1 + stride - 1) / stride + 1;
  p.status = status;
  if (status) XDE_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int), s));
  if (method == XDE_FIXED_EULER) return tc::tc_dispatch<0>(p, s);
  if (method == XDE_FIXED_MIDPOINT) return tc::tc_dispatch<3>(p, s);
  return tc::tc_dispatch<1>(p, s);
}

int sde_tc(int scheme, const xde_mlp_field_t *f, const xde_mlp_field_t *g, const float *y0, long long B,
           const float *t_span, int T, const BmSource &bm, int stride, float *out, int *status, cudaStream_t s) {
  XDE_REQUIRE(scheme == XDE_SDE_EM, XDE_E_UNSUPPORTED_FIELD,
              "Milstein (an extension without a reference counterpart) is fused for small states (D <= 8) only");
  XDE_REQUIRE(f->h == g->h, XDE_E_UNSUPPORTED_FIELD, "tensor-core sde: drift and diffusion must share the hidden width");
  tc::TcParams p{};
  p.f = *f;
  p.g = *g;
  p.y0 = y0;
  p.t_span = t_span;
  p.bm = bm;
  p.out = out;
  p.B = B;
  p.T = T;
  p.stride = stride;
  p.n_out = (T - 1 + stride - 1) / stride + 1;
  p.status = status;
  if (status) XDE_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(int), s));
  return bm.table ? tc::tc_dispatch<2>(p, s) : tc::tc_dispatch<4>(p, s);
}

}  // namespace xde
