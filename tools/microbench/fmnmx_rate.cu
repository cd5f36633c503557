// Issue rate of the 3-input FMNMX3 (max.f32 a, b, c) against the 2-input FMNMX and FADD on one SM sub-partition: the row max of the
// ping-pong attention kernel is 43 FMNMX3 in 8 chains and takes ~350 cycles in the phase trace.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/fmnmx_rate tools/microbench/fmnmx_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int kOp>
__global__ void k(float* out, const float* in, int iters, long long* clocks) {
  float a[8], x = in[threadIdx.x], y = in[threadIdx.x + 32];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + 64 + i];
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {          // 8 independent chains
      if (kOp == 0) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(x), "f"(y));
      if (kOp == 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x));
      if (kOp == 2) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(x));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

int main() {
  float *in, *out;
  long long* dc;
  cudaMalloc(&in, 4096 * 4);
  cudaMemset(in, 0, 4096 * 4);
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&dc, 148 * 8);
  const int iters = 4096;
  const char* names[3] = {"FMNMX3 (max.f32 a,b,c)", "FMNMX  (max.f32 a,b)", "FADD"};
  for (int warps : {1, 2, 4}) {            // warps per SM sub-partition
    for (int op = 0; op < 3; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) k<0><<<148, warps * 128>>>(out, in, iters, dc);
        if (op == 1) k<1><<<148, warps * 128>>>(out, in, iters, dc);
        if (op == 2) k<2><<<148, warps * 128>>>(out, in, iters, dc);
      }
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
      printf("FMNMX_RATE %-24s warps/SMSP=%d  %.2f cycles per warp-instruction per SMSP\n", names[op], warps,
             (double)c / ((double)iters * 8 * warps));
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
