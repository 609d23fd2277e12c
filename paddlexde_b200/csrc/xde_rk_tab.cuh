// xde_rk_tab.cuh -- an embedded Runge-Kutta pair as data (stage count, FSAL test, controller order, tables rounded
// once to fp32: solver/base_adaptive_solver_rk.py:73-79, 172-176) and the root the controller needs; shared by the
// table-driven kernels for small states (xde_adaptive_rk.cu) and for large states (xde_tile_adaptive.cuh).
#pragma once

#include "xde_common.cuh"

namespace xde {

constexpr int kMaxStages = 13;  // len(alpha) of Dopri8; k has one more column

struct RkTab {
  int S, order, fsal, _pad;
  float alpha[kMaxStages];
  float beta[kMaxStages][kMaxStages];
  float csol[kMaxStages + 1], cerr[kMaxStages + 1], cmid[kMaxStages + 1];
};

// float64 tableau of `method` (XDE_RK_*), rounded once to fp32; false for an unknown method (xde_adaptive_rk.cu)
bool make_tab(int method, RkTab &t);

// r ** (1/p), r finite and > 0: p = 2 sqrt (IEEE), p = 8 three sqrts, p = 3 integer seed + 4 Newton steps
// x <- (2x + r/x^2)/3, p = 5 root5 -- the oracle's orc_rootpf, operation for operation.
__device__ __forceinline__ float rootp(float r, int p) {
  if (p == 5) return root5(r);
  if (p == 2) return __fsqrt_rn(r);
  if (p == 8) return __fsqrt_rn(__fsqrt_rn(__fsqrt_rn(r)));
  if (p == 3) {
    float x = __uint_as_float(__float_as_uint(r) / 3u + 0x2A555555u);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const float q = __fdiv_rn(r, x * x);
      x = fmaf(2.0f, x, q) * (float)(1.0 / 3.0);
    }
    return x;
  }
  return powf(r, __fdiv_rn(1.0f, (float)p));
}

}  // namespace xde
