// xde_fixed128.cuh -- exact, order-independent accumulation of fp32 addends (arithmetic specification, round 2).
//
// The batch sums of the parameter-gradient dynamics (functional/odeint_adjoint.py:108-114: autograd sums the
// per-trajectory outer products over the batch) decide the accept/reject sequence of the reference's default
// mixed adjoint norm at tolerances below fp32 epsilon, so their value must not depend on the summation tree.
// Specification (shared with oracle/xde_oracle.c: fx_add_float / fx_to_float): addends are truncated toward zero
// to a grid of 2^-59 and added in a 128-bit two's-complement integer; the total is rounded once to fp32 (RN-even).
// Integer addition is associative: any grid, any order, atomics included, give the same bits.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace xde {

struct Fx128 {
  unsigned long long lo, hi;
};

// false: the addend is not representable (non-finite or |x| >= 2^40) -> the sum is NaN.  24-bit mantissa << 75
// < 2^99: 2^28 addends fit below 2^127.
__device__ __forceinline__ bool fx_from_float(float x, Fx128 &r) {
  const unsigned bits = __float_as_uint(x);
  const int E = (int)((bits >> 23) & 0xffu);
  const unsigned M = bits & 0x7fffffu;
  if (E == 255) return false;
  const unsigned long long mant = E ? (unsigned long long)(M | 0x800000u) : (unsigned long long)M;
  const int shift = (E ? E - 150 : -149) + 59;
  unsigned long long lo, hi;
  if (shift >= 0) {
    if (shift > 75) return false;
    if (shift >= 64) {
      lo = 0ull;
      hi = mant << (shift - 64);
    } else {
      lo = mant << shift;
      hi = shift ? (mant >> (64 - shift)) : 0ull;
    }
  } else {
    const int s = -shift;
    lo = (s >= 32) ? 0ull : (mant >> s);
    hi = 0ull;
  }
  if (bits >> 31) {  // two's complement negation
    lo = ~lo + 1ull;
    hi = ~hi + (lo == 0ull ? 1ull : 0ull);
  }
  r.lo = lo;
  r.hi = hi;
  return true;
}

__device__ __forceinline__ Fx128 fx_add(Fx128 a, Fx128 b) {
  Fx128 r;
  r.lo = a.lo + b.lo;
  r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
  return r;
}

// acc[0] = low word, acc[1] = high word; shared or global memory.  The carry out of the low word is decided by the
// atomic that performed the addition, so the two words end up exact whatever the interleaving.
__device__ __forceinline__ void fx_atomic_add(unsigned long long *acc, Fx128 v) {
  if ((v.lo | v.hi) == 0ull) return;
  unsigned long long hi = v.hi;
  if (v.lo) {
    const unsigned long long old = atomicAdd(acc, v.lo);
    if (old + v.lo < old) hi += 1ull;
  }
  if (hi) atomicAdd(acc + 1, hi);
}

// the total, rounded once to fp32 (round to nearest even)
__device__ __forceinline__ float fx_to_float(Fx128 a) {
  const bool neg = (a.hi >> 63) != 0ull;
  unsigned long long lo = a.lo, hi = a.hi;
  if (neg) {
    lo = ~lo + 1ull;
    hi = ~hi + (lo == 0ull ? 1ull : 0ull);
  }
  if ((lo | hi) == 0ull) return 0.0f;
  const int p = hi ? 127 - __clzll((long long)hi) : 63 - __clzll((long long)lo);
  float r;
  if (p <= 23) {
    r = (float)(unsigned)lo * 1.7347234759768071e-18f;  // 2^-59: exact scaling
  } else {
    const int sh = p - 23;  // 1 .. 104
    unsigned long long top, rem_lo, rem_hi, half_lo, half_hi;
    if (sh >= 64) {
      top = hi >> (sh - 64);
      rem_lo = lo;
      rem_hi = (sh == 64) ? 0ull : (hi & ((1ull << (sh - 64)) - 1ull));
      half_lo = (sh == 64) ? (1ull << 63) : 0ull;
      half_hi = (sh == 64) ? 0ull : (1ull << (sh - 65));
    } else {
      top = (lo >> sh) | (hi << (64 - sh));
      rem_lo = lo & ((1ull << sh) - 1ull);
      rem_hi = 0ull;
      half_lo = 1ull << (sh - 1);
      half_hi = 0ull;
    }
    unsigned t = (unsigned)top;  // 24 bits
    const bool gt = (rem_hi > half_hi) || (rem_hi == half_hi && rem_lo > half_lo);
    const bool eq = (rem_hi == half_hi) && (rem_lo == half_lo);
    if (gt || (eq && (t & 1u))) t++;
    r = (float)t * __uint_as_float((unsigned)(sh - 59 + 127) << 23);
  }
  return neg ? -r : r;
}

// ---- carry-free limb accumulators (same value, cheaper atomics) --------------------------------------------------
// A 128-bit atomic add needs the low word's return value for the carry: ~150 cycles of shared-memory round trip per
// addend, serialised at the occupancy of the cooperative kernel (r2d: 3x slower).  Integer limbs with HEADROOM need
// no carries while accumulating: the value is sum_i limb_i * 2^(W i) with signed limbs, every add is a
// fire-and-forget RED, and the carries are propagated once when the total is read.
//   shared memory: 6 x int32 limbs of 19 bits (an addend touches <= 3 of them; 2^12 adds of headroom per limb)
//   global memory: 3 x int64 limbs of 43 bits (2^20 adds of headroom)
constexpr int kFxSLimbs = 6, kFxSBits = 19;
constexpr int kFxGLimbs = 3, kFxGBits = 43;

__device__ __forceinline__ Fx128 fx_shl_signed(long long v, int sh) {  // sign-extended v * 2^sh, 0 <= sh < 128
  Fx128 r;
  const unsigned long long u = (unsigned long long)v, ext = (v < 0) ? ~0ull : 0ull;
  if (sh == 0) {
    r.lo = u;
    r.hi = ext;
  } else if (sh < 64) {
    r.lo = u << sh;
    r.hi = (ext << sh) | (u >> (64 - sh));
  } else {
    r.lo = 0ull;
    r.hi = u << (sh - 64);
  }
  return r;
}

// acc: kFxSLimbs int32 words in shared memory, `stride` words apart
__device__ __forceinline__ bool fx_limbs_add_float(int *acc, int stride, float x) {
  const unsigned bits = __float_as_uint(x);
  const int E = (int)((bits >> 23) & 0xffu);
  const unsigned M = bits & 0x7fffffu;
  if (E == 255) return false;
  const unsigned long long mant = E ? (unsigned long long)(M | 0x800000u) : (unsigned long long)M;
  const int shift = (E ? E - 150 : -149) + 59;
  if (shift > 75) return false;
  int i0 = 0;
  unsigned long long m;
  if (shift >= 0) {
    i0 = shift / kFxSBits;
    m = mant << (shift - kFxSBits * i0);  // < 2^(24 + 18)
  } else {
    m = (-shift >= 32) ? 0ull : (mant >> (-shift));
  }
  const int neg = (int)(bits >> 31);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int limb = (int)((m >> (kFxSBits * k)) & ((1u << kFxSBits) - 1u));
    if (limb) atomicAdd(acc + (i0 + k) * stride, neg ? -limb : limb);
  }
  return true;
}

__device__ __forceinline__ Fx128 fx_limbs_total(const int *acc, int stride) {
  Fx128 t;
  t.lo = t.hi = 0ull;
#pragma unroll
  for (int i = 0; i < kFxSLimbs; ++i) t = fx_add(t, fx_shl_signed((long long)acc[i * stride], kFxSBits * i));
  return t;
}

// add a 128-bit value to kFxGLimbs int64 words in global memory (no return values needed)
__device__ __forceinline__ void fx_glimbs_add(unsigned long long *acc, Fx128 v) {
  const bool neg = (v.hi >> 63) != 0ull;
  unsigned long long lo = v.lo, hi = v.hi;
  if (neg) {
    lo = ~lo + 1ull;
    hi = ~hi + (lo == 0ull ? 1ull : 0ull);
  }
  const unsigned long long mask = (1ull << kFxGBits) - 1ull;
  const unsigned long long l0 = lo & mask;
  const unsigned long long l1 = ((lo >> kFxGBits) | (hi << (64 - kFxGBits))) & mask;
  const unsigned long long l2 = hi >> (2 * kFxGBits - 64);
  if (l0) atomicAdd(acc, neg ? (0ull - l0) : l0);
  if (l1) atomicAdd(acc + 1, neg ? (0ull - l1) : l1);
  if (l2) atomicAdd(acc + 2, neg ? (0ull - l2) : l2);
}

__device__ __forceinline__ Fx128 fx_glimbs_total(const unsigned long long *acc) {
  Fx128 t;
  t.lo = t.hi = 0ull;
#pragma unroll
  for (int i = 0; i < kFxGLimbs; ++i) t = fx_add(t, fx_shl_signed((long long)__ldcg(acc + i), kFxGBits * i));
  return t;
}

}  // namespace xde
