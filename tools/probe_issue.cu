// probe_issue.cu -- how many issue cycles does a packed FFMA2 cost next to ALU / XU instructions on sm_100a?
//
// The cfg2 kernels (dopri5 forward / adjoint, D = 2, H = 50) are made of packed fp32 instructions
// (FFMA2 / FMUL2 / FADD2) plus a few FMNMX / FSETP / FSEL / MUFU per hidden-unit pair.  ncu reports
// "issue active" ~53 % and "FMA pipe active" ~49 % for the adjoint, yet neither more warps nor more
// independent chains per warp make it faster.  This probe measures the cycles one scheduler needs for
// fixed instruction mixes with plenty of independent work (8 warps per scheduler, 8 independent chains
// per warp), to tell whether an FFMA2 leaves its second FMA-pipe cycle free for another instruction.
//
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_probe_issue tools/probe_issue.cu
// run:   tools/_probe_issue            (prints cycles per loop iteration and scheduler for every mix)
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;

template <int NF2, int NF1, int NALU, int NXU>
__global__ void __launch_bounds__(256) mix_kernel(float *sink, int iters, float x, float y, long long *cycles) {
  u64 a[8];
  float b[8], c[8], m[4];
  u64 xx, yy;
  asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
  asm("mov.b64 %0, {%1, %1};" : "=l"(yy) : "f"(y));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float v = (float)(threadIdx.x + i);
    asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(v));
    b[i] = v;
    c[i] = v * 0.5f;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) m[i] = 1.0f + (float)(threadIdx.x + i);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < NF2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(xx), "l"(yy));
      if (i < NF1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(x), "f"(y));
      if (i < NALU) asm volatile("max.f32 %0, %0, %1;" : "+f"(c[i]) : "f"(y));
      if (i < NXU) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(m[i & 3]) : "f"(x));  // no dependent chain
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
    s += lo + hi + b[i] + c[i];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) s += m[i];
  if (s == 123456.789f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int NF2, int NF1, int NALU, int NXU>
static void run(const char *what, int sms, double clock_hz) {
  float *sink;
  long long *cyc, h = 0;
  cudaMalloc(&sink, 4);
  cudaMalloc(&cyc, 8);
  const int iters = 1 << 15;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mix_kernel<NF2, NF1, NALU, NXU>, 256, 0);
  if (per_sm > 4) per_sm = 4;  // 4 CTAs of 256 threads per SM = 32 warps per SM = 8 per scheduler
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  mix_kernel<NF2, NF1, NALU, NXU><<<sms * per_sm, 256>>>(sink, iters, 0.999f, 0.001f, cyc);
  cudaEventRecord(e0);
  mix_kernel<NF2, NF1, NALU, NXU><<<sms * per_sm, 256>>>(sink, iters, 0.999f, 0.001f, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double warps_per_sched = per_sm * 8 / 4.0;
  // cycles one scheduler spends per loop iteration of ONE warp: from the event time at the maximum SM clock, and
  // from clock64() of one thread (the two agree when the kernel runs at the maximum clock)
  const double per_ev = ms * 1e-3 * clock_hz / iters / warps_per_sched;
  const double per_ck = (double)h / iters / warps_per_sched;
  printf("%-44s %6.2f (events) %6.2f (clock64) cycles per (%d FFMA2 + %d FFMA + %d FMNMX + %d MUFU), %d CTAs/SM\n", what,
         per_ev, per_ck, NF2, NF1, NALU, NXU, per_sm);
  cudaFree(sink);
  cudaFree(cyc);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  const int sms = p.multiProcessorCount;
  const double hz = 1e3 * (double)p.clockRate;
  printf("clock %.0f MHz\n", hz * 1e-6);
  run<0, 8, 0, 0>("scalar FFMA only", sms, hz);
  run<8, 0, 0, 0>("packed FFMA2 only", sms, hz);
  run<0, 8, 8, 0>("FFMA + FMNMX 1:1", sms, hz);
  run<8, 0, 4, 0>("FFMA2 + FMNMX 2:1", sms, hz);
  run<8, 0, 8, 0>("FFMA2 + FMNMX 1:1", sms, hz);
  run<4, 0, 8, 0>("FFMA2 + FMNMX 1:2", sms, hz);
  run<0, 0, 8, 0>("FMNMX only", sms, hz);
  run<8, 0, 0, 1>("FFMA2 + MUFU 8:1", sms, hz);
  run<8, 0, 0, 2>("FFMA2 + MUFU 4:1", sms, hz);
  run<8, 0, 4, 1>("FFMA2 + FMNMX + MUFU 8:4:1 (tanh-like mix)", sms, hz);
  run<4, 4, 0, 0>("FFMA2 + FFMA 1:1", sms, hz);
  return 0;
}
