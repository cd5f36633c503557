// MUFU.EX2 issue-rate microbenchmark: exps per clock per SM as a function of warps per SM sub-partition and of the
// independent chains per thread (ILP), with and without FFMA work interleaved. Answers: can TWO warps per scheduler (the
// ping-pong attention kernel's softmax occupancy) saturate the SFU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/mufu_rate tools/microbench/mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int FMA_PER_EX2>
__global__ void k(float* out, int iters, float seed) {
  float x[ILP], y[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { x[i] = seed * (threadIdx.x + i + 1) * 1e-3f; y[i] = x[i]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[i]) : "f"(x[i]));
#pragma unroll
      for (int f = 0; f < FMA_PER_EX2; ++f) y[i] = fmaf(y[i], 0.999f, x[i] * 1e-9f + 0.5f);
      x[i] = x[i] * 0.25f - 1.0f;   // keep the argument bounded (1 FMUL/FFMA per ex2 in every variant)
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
  if (s == 123.456f) out[0] = s;
}

template <int ILP, int FMA>
void run(int warps_per_sm, int sms, float mhz) {
  const int iters = 4096;
  float* d;
  cudaMalloc(&d, 4);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<ILP, FMA><<<sms, warps_per_sm * 32>>>(d, 64, 1.f);
  cudaEventRecord(a);
  k<ILP, FMA><<<sms, warps_per_sm * 32>>>(d, iters, 1.f);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double exps = (double)sms * warps_per_sm * 32 * ILP * iters;
  const double clk = ms * 1e-3 * mhz * 1e6;
  printf("MUFU_RATE ilp=%d fma_per_ex2=%d warps_per_smsp=%.1f  %.2f ex2/clk/SM  (%.3f ms)\n", ILP, FMA, warps_per_sm / 4.0,
         exps / clk / sms, ms);
  cudaFree(d);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float mhz = khz / 1000.f;   // nominal boost clock: the rates below are relative to it
  printf("device %s, %d SMs, clock %.0f MHz (nominal)\n", p.name, p.multiProcessorCount, mhz);
  for (int w : {4, 8, 16, 32}) {
    run<4, 0>(w, p.multiProcessorCount, mhz);
    run<8, 0>(w, p.multiProcessorCount, mhz);
    run<16, 0>(w, p.multiProcessorCount, mhz);
    run<8, 2>(w, p.multiProcessorCount, mhz);
    run<8, 4>(w, p.multiProcessorCount, mhz);
  }
  return 0;
}
