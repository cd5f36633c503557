// tcgen05.ld (TMEM -> registers) throughput per SM: the level-0 attention reads one fp32 score tile (128 x 128 x 4 B =
// 64 KB) out of tensor memory per 128 x 128 scores, so this rate is a roofline of the softmax kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/ldtm_rate tools/microbench/ldtm_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__global__ void k(uint32_t* out, int iters, long long* clocks) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {                // 4 x 32 columns = one 128-column score tile row set per warp
      uint32_t v[32];
      tmem_ld32(base + ((it & 3) * 128 + c * 32) % 512, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) out[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  uint32_t* d;
  long long* dc;
  cudaMalloc(&d, 4);
  cudaMalloc(&dc, sizeof(long long) * p.multiProcessorCount);
  const int iters = 2048;
  for (int warps : {4, 8, 16}) {
    k<<<p.multiProcessorCount, warps * 32>>>(d, 16, dc);
    k<<<p.multiProcessorCount, warps * 32>>>(d, iters, dc);
    cudaDeviceSynchronize();
    long long c0;
    cudaMemcpy(&c0, dc, sizeof(long long), cudaMemcpyDeviceToHost);
    const double bytes = (double)warps * iters * 4 * 32 * 32 * 4;      // per SM
    printf("LDTM_RATE warps=%d  %.1f bytes/clk/SM  (%lld clk; one 64 KB score tile = %.0f clk)\n", warps, bytes / c0, c0,
           65536.0 * c0 / bytes);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
