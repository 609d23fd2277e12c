// xde_runtime.cu -- host-side plumbing of libxde_b200: error text, launch counter, device queries.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "xde_common.cuh"

namespace xde {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_last_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace xde

extern "C" XDE_EXPORT int xde_abi_version(void) { return XDE_ABI_VERSION; }
extern "C" XDE_EXPORT const char *xde_last_error(void) { return xde::g_err; }
extern "C" XDE_EXPORT unsigned long long xde_launch_count(void) { return xde::g_launches.load(); }
extern "C" XDE_EXPORT void xde_default_ctrl_opts(xde_ctrl_opts_t *o) {
  // solver/base_adaptive_solver_rk.py:32-49, functional/odeint.py:14-15
  o->rtol = 1e-7f;
  o->atol = 1e-9f;
  o->min_step = 0.0f;
  o->max_step = INFINITY;
  o->first_step = NAN;
  o->safety = 0.9f;
  o->ifactor = 10.0f;
  o->dfactor = 0.2f;
  o->max_num_steps = 2147483647;
  o->_pad = 0;
}
