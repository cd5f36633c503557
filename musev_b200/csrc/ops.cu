// HBM-bound kernels (see ops.cuh). Design rules: 128-bit loads/stores on the contiguous channel axis, fp32
// statistics, grids sized to cover >= 2 waves of the 148 SMs where the tensor is large enough.
#include "ops.cuh"

#include <math.h>

#include "stats.cuh"

namespace mvb {

// x * sigmoid(x) with the two SFU approximations (ex2.approx.ftz, rcp.approx.ftz): ~1e-6 relative, five instructions
// (__expf without -ftz adds a denormal-range fix-up of three more)
__device__ __forceinline__ float silu_f(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}

struct alignas(16) Half8 {
  __half2 h[4];
};
// explicit 128-bit accesses: a plain struct copy of Half8 is compiled to four 32-bit loads / stores, which quadruples
// the L1 sector traffic and pins every streaming kernel at ~3.3 TB/s (ncu: l1tex throughput 93 %)
__device__ __forceinline__ Half8 ld_half8(const __half* p) {
  Half8 r;
  *reinterpret_cast<uint4*>(&r) = __ldg(reinterpret_cast<const uint4*>(p));
  return r;
}
__device__ __forceinline__ void st_half8(__half* p, const Half8& v) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}

// ------------------------------------------------------------------------------------------------ GroupNorm
// grid (chunks, NF); block 256. Thread owns a fixed 8-channel vector column and strides over pixels, so its
// accumulators stay in registers; groups are even-sized, so a half2 never straddles two groups. The block reduction is
// deterministic (no atomics): per-(row slot, channel pair) partials go to smem and thread g sums group g in fixed order.
// Statistics are carried as (count, mean, M2 = sum of squared deviations) triples and merged with Chan's parallel update, never
// as raw sum / sum-of-squares: E[x^2] - E[x]^2 in fp32 loses the variance once |mean| >> std (activations of real checkpoints
// have such channels). Each thread accumulates deviations from a pilot value (its first sample), which
// keeps the running sums at the scale of the spread, not of the mean.
__device__ __forceinline__ void chan_merge(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
  if (nb <= 0.f) return;
  const float nt = n + nb;
  const float delta = mb - mean;
  const float w = __fdividef(nb, nt);   // counts are small integers: the approximate reciprocal is exact enough (1 ulp)
  mean = fmaf(delta, w, mean);
  m2 = m2 + m2b + delta * delta * n * w;
  n = nt;
}

__device__ __forceinline__ void gn_stats_unit(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                              int G, float* __restrict__ part, int f, int chunk, int chunks, float2* spair) {
  const int C = C0 + C1;
  const int vecs = C / 8;
  const int pairs = C / 2;
  const int cpg2 = (C / G) / 2;      // channel pairs per group
  const int cols_per_pass = vecs < (int)blockDim.x ? vecs : (int)blockDim.x;
  const int rows_per_iter = blockDim.x / cols_per_pass;
  const int r0 = threadIdx.x / cols_per_pass;
  const int p_begin = (int)(((long long)HW * chunk) / chunks);
  const int p_end = (int)(((long long)HW * (chunk + 1)) / chunks);
  for (int v0 = 0; v0 < vecs; v0 += cols_per_pass) {
    const int v = v0 + threadIdx.x % cols_per_pass;
    if (r0 < rows_per_iter && v < vecs) {
      const int c = v * 8;
      const __half* src = (c < C0) ? x0 + (size_t)f * HW * C0 + c : x1 + (size_t)f * HW * C1 + (c - C0);
      const int ld = (c < C0) ? C0 : C1;
      float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
      __half2 piv2[4];
      int p = p_begin + r0;
      {                                // pilot: the first sample of every channel pair (independent of the loads below)
        const Half8 h0 = ld_half8(src + (size_t)(p < p_end ? p : p_begin) * ld);
#pragma unroll
        for (int j = 0; j < 4; ++j) piv2[j] = __low2half2(h0.h[j]);
      }
      // four independent 16-byte loads in flight per thread; the deviation from the pilot is taken in fp16 (one HSUB2 per
      // pair: exact for neighbours of the pilot, 2^-11 relative otherwise), everything after it in fp32
      for (; p + 3 * rows_per_iter < p_end; p += 4 * rows_per_iter) {
        Half8 hv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) hv[u] = ld_half8(src + (size_t)(p + u * rows_per_iter) * ld);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 t = __half22float2(__hsub2(hv[u].h[j], piv2[j]));
            s[j] += t.x + t.y;
            q[j] = fmaf(t.x, t.x, fmaf(t.y, t.y, q[j]));
          }
      }
      for (; p < p_end; p += rows_per_iter) {
        const Half8 hv = ld_half8(src + (size_t)p * ld);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(__hsub2(hv.h[j], piv2[j]));
          s[j] += t.x + t.y;
          q[j] = fmaf(t.x, t.x, fmaf(t.y, t.y, q[j]));
        }
      }
      float piv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) piv[j] = __low2float(piv2[j]);
      // (mean, M2) of this thread's 2 x count samples of each pair
      const int span = p_end - p_begin - r0;
      const float cnt = span > 0 ? 2.f * (float)((span + rows_per_iter - 1) / rows_per_iter) : 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float ms = cnt > 0.f ? s[j] / cnt : 0.f;
        spair[(size_t)r0 * pairs + v * 4 + j] = make_float2(piv[j] + ms, fmaxf(q[j] - s[j] * ms, 0.f));
      }
    }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float n = 0.f, mean = 0.f, m2 = 0.f;
    for (int r = 0; r < rows_per_iter; ++r) {
      const int span = p_end - p_begin - r;
      const float cnt = span > 0 ? 2.f * (float)((span + rows_per_iter - 1) / rows_per_iter) : 0.f;
      for (int pc = g * cpg2; pc < (g + 1) * cpg2; ++pc) {
        const float2 t = spair[(size_t)r * pairs + pc];
        chan_merge(n, mean, m2, cnt, t.x, t.y);
      }
    }
    float* dst = part + ((size_t)f * chunks + chunk) * 2 * G + 2 * g;
    dst[0] = mean;                     // count is implied: (p_end - p_begin) * channels per group
    dst[1] = m2;
  }
}

__global__ void __launch_bounds__(256, 8)     // 8 resident blocks per SM = 32 registers: the kernel is latency-bound
gn_stats_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW, int G,
                float* __restrict__ part) {
  extern __shared__ float2 spair[];  // [rows_per_iter][C/2] (sum, sumsq) per channel pair
  // frames in reverse order: the tail of the tensor is what the producing GEMM wrote last and is still in L2
  gn_stats_unit(x0, C0, x1, C1, HW, G, part, gridDim.y - 1 - blockIdx.y, blockIdx.x, gridDim.x, spair);
}

cudaError_t gn_stats(cudaStream_t s, const __half* x0, int C0, const __half* x1, int C1, int NF, int HW, int G,
                     float* part, int* chunks_out) {
  ProfScope prof(s, KC_GROUPNORM);
  const int C = C0 + (x1 ? C1 : 0);
  if (!x1) C1 = 0;
  if ((C0 % 8) || (C1 % 8) || (C % G) || ((C / G) % 2)) return cudaErrorInvalidValue;
  const int threads = 256;
  // ~8 resident blocks per SM (the kernel is latency-bound below that), at least ~32 pixels per block
  int chunks = (8 * 148) / NF;          // rounded down: one full wave
  const int maxc = HW / 32 > 0 ? HW / 32 : 1;
  if (chunks > maxc) chunks = maxc;
  if (chunks > kGnMaxChunks) chunks = kGnMaxChunks;
  if (chunks < 1) chunks = 1;
  *chunks_out = chunks;
  const int cols = (C / 8) < threads ? (C / 8) : threads;
  const size_t smem = (size_t)(threads / cols) * (C / 2) * sizeof(float2);
  gn_stats_kernel<<<dim3(chunks, NF), threads, smem, s>>>(x0, C0, x1, C1, HW, G, part);
  return cudaGetLastError();
}

// Reduces the per-frame partials of one statistics group (fps consecutive frames) to mean / rstd per group:
// stats[NF/fps][G][2]. grid (NF/fps, G), one warp per (statistics group, channel group): lanes walk the partials in a
// fixed interleaved order, then a fixed shuffle tree -- deterministic.
__device__ __forceinline__ void gn_finalize_unit(const float* __restrict__ part, int chunks, int fps, int G, int HW, int cpg, float eps,
                                                 float* __restrict__ stats, int sg, int g, int lane) {
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int i = lane; i < fps * chunks; i += 32) {
    const int chunk = i % chunks;
    const int pix = (int)(((long long)HW * (chunk + 1)) / chunks) - (int)(((long long)HW * chunk) / chunks);
    const float2 pp = *reinterpret_cast<const float2*>(part + ((size_t)sg * fps * chunks + i) * 2 * G + 2 * g);
    chan_merge(n, mean, m2, (float)pix * (float)cpg, pp.x, pp.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {   // fixed butterfly: deterministic
    const float nb = __shfl_xor_sync(0xffffffffu, n, o);
    const float mb = __shfl_xor_sync(0xffffffffu, mean, o);
    const float qb = __shfl_xor_sync(0xffffffffu, m2, o);
    chan_merge(n, mean, m2, nb, mb, qb);
  }
  if (lane == 0) {
    const float var = n > 0.f ? fmaxf(m2 / n, 0.f) : 0.f;
    stats[((size_t)sg * G + g) * 2] = mean;
    stats[((size_t)sg * G + g) * 2 + 1] = rsqrtf(var + eps);
  }
}
__global__ void gn_finalize_kernel(const float* __restrict__ part, int chunks, int fps, int G, int HW, int cpg, float eps,
                                   float* __restrict__ stats) {
  gn_finalize_unit(part, chunks, fps, G, HW, cpg, eps, stats, blockIdx.x, blockIdx.y, threadIdx.x);
}

// 5-D statistics (fps = T frames per group of statistics) have fps x chunks partials per channel group -- 578 at level 0, which
// one warp walks in 18.6 us (ncu, profiles/r02_ncu_kernels_full_summary.txt), and 104 of the 166 GroupNorms of a forward are
// of this kind. Same reduction on kGnFinalizeWarps warps: every thread merges its strided share, a fixed butterfly per warp,
// then warp 0 merges the warp results in index order -- deterministic.
static constexpr int kGnFinalizeWarps = 8;
__global__ void __launch_bounds__(kGnFinalizeWarps * 32)
gn_finalize_wide_kernel(const float* __restrict__ part, int chunks, int fps, int G, int HW, int cpg, float eps,
                        float* __restrict__ stats) {
  __shared__ float sred[kGnFinalizeWarps][3];
  const int sg = blockIdx.x, g = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float n = 0.f, mean = 0.f, m2 = 0.f;
  for (int i = threadIdx.x; i < fps * chunks; i += kGnFinalizeWarps * 32) {
    const int chunk = i % chunks;
    const int pix = (int)(((long long)HW * (chunk + 1)) / chunks) - (int)(((long long)HW * chunk) / chunks);
    const float2 pp = *reinterpret_cast<const float2*>(part + ((size_t)sg * fps * chunks + i) * 2 * G + 2 * g);
    chan_merge(n, mean, m2, (float)pix * (float)cpg, pp.x, pp.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float nb = __shfl_xor_sync(0xffffffffu, n, o);
    const float mb = __shfl_xor_sync(0xffffffffu, mean, o);
    const float qb = __shfl_xor_sync(0xffffffffu, m2, o);
    chan_merge(n, mean, m2, nb, mb, qb);
  }
  if (lane == 0) { sred[warp][0] = n; sred[warp][1] = mean; sred[warp][2] = m2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kGnFinalizeWarps; ++w) chan_merge(n, mean, m2, sred[w][0], sred[w][1], sred[w][2]);
    const float var = n > 0.f ? fmaxf(m2 / n, 0.f) : 0.f;
    stats[((size_t)sg * G + g) * 2] = mean;
    stats[((size_t)sg * G + g) * 2 + 1] = rsqrtf(var + eps);
  }
}

// grid (pixel blocks, NF); block 256; each block streams ~64 KB.
__device__ __forceinline__ void gn_apply_unit(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW,
                                              int G, const float* __restrict__ stats, int fps, const float* __restrict__ gamma,
                                              const float* __restrict__ beta, int silu, __half* __restrict__ y, int pix_per_block,
                                              int f, int pblock, float* sm_gn) {
  const int C = C0 + C1;
  const int cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    sm_gn[g] = stats[((size_t)(f / fps) * G + g) * 2];
    sm_gn[G + g] = stats[((size_t)(f / fps) * G + g) * 2 + 1];
  }
  __syncthreads();
  const int vecs = C / 8;
  const int p0 = pblock * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  const int total = (p1 - p0) * vecs;
  // a thread keeps the same channel vector when blockDim is a multiple of vecs; the affine terms are then loaded once
  const bool fixed_col = (blockDim.x % vecs) == 0 && cpg >= 8;
  float ga[8], be[8];
  if (fixed_col) {
    // 8 consecutive channels touch at most two groups (cpg >= 8 whenever this path is taken): one division per thread
    const int c = (threadIdx.x % vecs) * 8;
    const int g0 = c / cpg;
    const int edge = (g0 + 1) * cpg;
    const float4 gm0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), gm1 = __ldg(reinterpret_cast<const float4*>(gamma + c) + 1);
    const float4 bt0 = __ldg(reinterpret_cast<const float4*>(beta + c)), bt1 = __ldg(reinterpret_cast<const float4*>(beta + c) + 1);
    const float gmv[8] = {gm0.x, gm0.y, gm0.z, gm0.w, gm1.x, gm1.y, gm1.z, gm1.w};
    const float btv[8] = {bt0.x, bt0.y, bt0.z, bt0.w, bt1.x, bt1.y, bt1.z, bt1.w};
    const int g1 = min(g0 + 1, G - 1);
    const float mean0 = sm_gn[g0], rstd0 = sm_gn[G + g0], mean1 = sm_gn[g1], rstd1 = sm_gn[G + g1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool hi = c + j >= edge;
      ga[j] = (hi ? rstd1 : rstd0) * gmv[j];
      be[j] = btv[j] - (hi ? mean1 : mean0) * ga[j];
    }
  }
  if (fixed_col) {
    // the thread walks down its channel vector: no per-element index arithmetic, affine terms in registers
    const int c = (threadIdx.x % vecs) * 8;
    const int rstep = blockDim.x / vecs;
    const __half* src = (c < C0) ? x0 + (size_t)f * HW * C0 + c : x1 + (size_t)f * HW * C1 + (c - C0);
    const int ld = (c < C0) ? C0 : C1;
    __half* dst = y + (size_t)f * HW * C + c;
#pragma unroll 2
    for (int p = p0 + threadIdx.x / vecs; p < p1; p += rstep) {
      const Half8 hv = ld_half8(src + (size_t)p * ld);
      Half8 ov;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 t = __half22float2(hv.h[j]);
        t.x = fmaf(t.x, ga[2 * j], be[2 * j]);
        t.y = fmaf(t.y, ga[2 * j + 1], be[2 * j + 1]);
        if (silu) { t.x = silu_f(t.x); t.y = silu_f(t.y); }
        ov.h[j] = __floats2half2_rn(t.x, t.y);
      }
      st_half8(dst + (size_t)p * C, ov);
    }
    return;
  }
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int p = p0 + i / vecs;
    const int c = (i % vecs) * 8;
    const __half* src = (c < C0) ? x0 + ((size_t)f * HW + p) * C0 + c : x1 + ((size_t)f * HW + p) * C1 + (c - C0);
    const Half8 hv = ld_half8(src);
    Half8 ov;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 t = __half22float2(hv.h[j]);
      if (fixed_col) {
        t.x = fmaf(t.x, ga[2 * j], be[2 * j]);
        t.y = fmaf(t.y, ga[2 * j + 1], be[2 * j + 1]);
      } else {
        const int cc = c + 2 * j;
        const int g = cc / cpg;
        const float mean = sm_gn[g], rstd = sm_gn[G + g];
        t.x = (t.x - mean) * rstd * __ldg(gamma + cc) + __ldg(beta + cc);
        t.y = (t.y - mean) * rstd * __ldg(gamma + cc + 1) + __ldg(beta + cc + 1);
      }
      if (silu) { t.x = silu_f(t.x); t.y = silu_f(t.y); }
      ov.h[j] = __floats2half2_rn(t.x, t.y);
    }
    st_half8(y + ((size_t)f * HW + p) * C + c, ov);
  }
}

__global__ void __launch_bounds__(256)
gn_apply_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int HW, int G,
                const float* __restrict__ stats, int fps, const float* __restrict__ gamma,
                const float* __restrict__ beta, int silu, __half* __restrict__ y, int pix_per_block) {
  extern __shared__ float sm_gn[];  // mean[G], rstd[G]
  gn_apply_unit(x0, C0, x1, C1, HW, G, stats, fps, gamma, beta, silu, y, pix_per_block, blockIdx.y, blockIdx.x, sm_gn);
}

// ---- the three passes in ONE persistent launch: statistics -> grid barrier -> finalize -> grid barrier -> apply.
// All blocks are co-resident (grid = occupancy x SMs), so a software barrier on a global counter is safe. The apply phase
// walks the frames in the opposite order of the statistics phase, i.e. it starts with the frames the statistics phase read
// last, which are still in L2 (126 MB): up to ~2/3 of the second read of a level-0 tensor (89 MB) no longer goes to HBM, and
// two launches per GroupNorm (332 per forward) disappear. Bit-identical to the three-kernel path for 4-D statistics (same partial layout, same
// fixed-order reductions).
__device__ __forceinline__ void gn_grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int spins = 0;
    for (;;) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if ((int)(v - target) >= 0) break;          // wrap-safe
      if (++spins > (1u << 27)) __trap();         // a protocol bug must not hang the GPU box
      __nanosleep(40);
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
gn_fused_kernel(const __half* __restrict__ x0, int C0, const __half* __restrict__ x1, int C1, int NF, int HW, int G, int chunks,
                float* __restrict__ part, float* __restrict__ stats, int fps, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, int silu, __half* __restrict__ y,
                int pix_per_block, int pblocks, unsigned int* counter, unsigned int base) {
  extern __shared__ float2 sm_fused[];
  // phase 1: per-(frame, chunk) partial sums, frames in reverse order (the producer's tail is still in L2)
  for (int u = blockIdx.x; u < NF * chunks; u += gridDim.x) {
    gn_stats_unit(x0, C0, x1, C1, HW, G, part, NF - 1 - u / chunks, u % chunks, chunks, sm_fused);
    __syncthreads();                              // the smem tile is reused by the next unit
  }
  gn_grid_barrier(counter, base + gridDim.x);
  // phase 2: one warp per (statistics group, channel group)
  {
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = (NF / fps) * G;
    if (warp < wpb)
      for (int u = blockIdx.x * wpb + warp; u < units; u += gridDim.x * wpb)
        gn_finalize_unit(part, chunks, fps, G, HW, (C0 + C1) / G, eps, stats, u / G, u % G, lane);
  }
  gn_grid_barrier(counter, base + 2u * gridDim.x);
  // phase 3: apply, frames ascending (phase 1 finished on frame 0)
  for (int u = blockIdx.x; u < NF * pblocks; u += gridDim.x) {
    gn_apply_unit(x0, C0, x1, C1, HW, G, stats, fps, gamma, beta, silu, y, pix_per_block, u / pblocks, u % pblocks,
                  reinterpret_cast<float*>(sm_fused));
    __syncthreads();
  }
}

cudaError_t gn_apply(cudaStream_t s, const __half* x0, int C0, const __half* x1, int C1, int NF, int HW, int G,
                     const float* part, int chunks, int fps, float eps, const float* gamma, const float* beta, int silu,
                     __half* y) {
  ProfScope prof(s, KC_GROUPNORM, 2);
  if (!x1) C1 = 0;
  const int C = C0 + C1;
  if (fps < 1 || (NF % fps) || G > 64) return cudaErrorInvalidValue;
  // mean / rstd live right behind the partial sums in the caller's scratch: NF*(kGnMaxChunks+1)*G*2 floats in total
  float* stats = const_cast<float*>(part) + (size_t)NF * kGnMaxChunks * G * 2;
  if (fps * chunks > 64) gn_finalize_wide_kernel<<<dim3(NF / fps, G), kGnFinalizeWarps * 32, 0, s>>>(part, chunks, fps, G, HW, C / G, eps, stats);
  else gn_finalize_kernel<<<dim3(NF / fps, G), 32, 0, s>>>(part, chunks, fps, G, HW, C / G, eps, stats);
  // ~64 KB of fp16 per block, block size a multiple of the number of channel vectors when possible
  const int vecs = C / 8;
  int threads = 256;
  if (vecs <= 256 && (256 % vecs) != 0) threads = (256 / vecs) * vecs;   // e.g. C=320: 240 threads, C=960: 240
  if (threads < 64) threads = 256;
  // ~4 waves of 6 resident blocks per SM (a 1.5-wave grid leaves half the machine idle for the second half), at least
  // 4 rows per thread so that the per-block prologue stays small
  const int rows_per_pass = threads / (vecs < threads ? vecs : threads) > 0 ? threads / (vecs < threads ? vecs : threads) : 1;
  int ppb = (int)(((long long)HW * NF + 148 * 24 - 1) / (148 * 24));
  if (ppb < 4 * rows_per_pass) ppb = 4 * rows_per_pass;
  if (ppb > HW) ppb = HW;
  int blocks = (HW + ppb - 1) / ppb;
  gn_apply_kernel<<<dim3(blocks, NF), threads, 2 * G * sizeof(float), s>>>(x0, C0, x1, C1, HW, G, stats, fps, gamma, beta,
                                                                           silu, y, ppb);
  return cudaGetLastError();
}

// One-launch GroupNorm (statistics + finalize + apply behind two grid barriers). `counter` is a zero-initialised device word
// owned by the caller, `*base` the number of arrivals it has seen so far (host bookkeeping; launches must be stream-ordered).
cudaError_t gn_fused(cudaStream_t s, const __half* x0, int C0, const __half* x1, int C1, int NF, int HW, int G, float* part,
                     int fps, float eps, const float* gamma, const float* beta, int silu, __half* y, int num_sms,
                     unsigned int* counter, unsigned int* base) {
  ProfScope prof(s, KC_GROUPNORM);
  if (!x1) C1 = 0;
  const int C = C0 + C1;
  if ((C0 % 8) || (C1 % 8) || (C % G) || ((C / G) % 2) || fps < 1 || (NF % fps) || G > 64) return cudaErrorInvalidValue;
  // same work decomposition as gn_stats / gn_apply
  int chunks = (8 * 148) / NF;
  const int maxc = HW / 32 > 0 ? HW / 32 : 1;
  if (chunks > maxc) chunks = maxc;
  if (chunks > kGnMaxChunks) chunks = kGnMaxChunks;
  if (chunks < 1) chunks = 1;
  const int vecs = C / 8;
  int threads = 256;
  if (vecs <= 256 && (256 % vecs) != 0) threads = (256 / vecs) * vecs;
  if (threads < 64) threads = 256;
  const int cols = vecs < threads ? vecs : threads;
  const int rows_per_pass = threads / cols > 0 ? threads / cols : 1;
  int ppb = (int)(((long long)HW * NF + 148 * 24 - 1) / (148 * 24));
  if (ppb < 4 * rows_per_pass) ppb = 4 * rows_per_pass;
  if (ppb > HW) ppb = HW;
  const int pblocks = (HW + ppb - 1) / ppb;
  size_t smem = (size_t)(threads / cols) * (C / 2) * sizeof(float2);
  if (smem < 2 * G * sizeof(float)) smem = 2 * G * sizeof(float);
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_fused_kernel, threads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorInvalidValue;
  if (per_sm > 8) per_sm = 8;
  long long units = (long long)NF * chunks;
  if ((long long)NF * pblocks > units) units = (long long)NF * pblocks;
  long long grid = (long long)per_sm * num_sms;      // every block must be resident: the kernel spins on a grid barrier
  if (grid > units) grid = units;
  float* stats = part + (size_t)NF * kGnMaxChunks * G * 2;
  gn_fused_kernel<<<(unsigned)grid, threads, smem, s>>>(x0, C0, x1, C1, NF, HW, G, chunks, part, stats, fps, eps, gamma, beta,
                                                       silu, y, ppb, pblocks, counter, *base);
  *base += 2u * (unsigned)grid;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// one warp handles R rows at a time (R independent row streams in flight per lane); C <= 2560
template <int VPL, int R>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, long long M, int C, float eps, const float* __restrict__ gamma,
                 const float* __restrict__ beta, __half* __restrict__ y) {
  const long long row0 = ((long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * R;
  if (row0 >= M) return;
  const int lane = threadIdx.x & 31;
  const int vecs = C / 8;
  Half8 hv[R][VPL];
#pragma unroll
  for (int r = 0; r < R; ++r) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < vecs && row0 + r < M) hv[r][i] = ld_half8(x + (row0 + r) * C + vi * 8);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (row0 + r >= M) break;
    float v[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < vecs) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(hv[r][i].h[j]);
          v[i][2 * j] = t.x; v[i][2 * j + 1] = t.y;
          s += t.x + t.y;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < vecs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q += d * d; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < vecs) {
        Half8 ov;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8) + 1);
        ov.h[0] = __floats2half2_rn((v[i][0] - mean) * rstd * g0.x + b0.x, (v[i][1] - mean) * rstd * g0.y + b0.y);
        ov.h[1] = __floats2half2_rn((v[i][2] - mean) * rstd * g0.z + b0.z, (v[i][3] - mean) * rstd * g0.w + b0.w);
        ov.h[2] = __floats2half2_rn((v[i][4] - mean) * rstd * g1.x + b1.x, (v[i][5] - mean) * rstd * g1.y + b1.y);
        ov.h[3] = __floats2half2_rn((v[i][6] - mean) * rstd * g1.z + b1.z, (v[i][7] - mean) * rstd * g1.w + b1.w);
        st_half8(y + (row0 + r) * C + vi * 8, ov);
      }
    }
  }
}

// C = 40 L (320 / 640 / 1280, the three transformer widths): L lanes share a row, each lane owns 5 vectors of 8
// channels (vector index = lane_in_row + L i, so a row is read as 5 fully coalesced segments) and a warp streams 32/L
// rows at a time; gamma / beta sit in smem. Two-pass (centered) variance on the register copy, ~8 instructions per
// element, which keeps the kernel on the HBM roofline instead of the issue port.
template <int L>
__global__ void __launch_bounds__(256)
layernorm40_kernel(const __half* __restrict__ x, long long M, float eps, const float* __restrict__ gamma,
                   const float* __restrict__ beta, __half* __restrict__ y) {
  constexpr int C = 40 * L;
  constexpr int RPW = 32 / L;   // rows per warp pass
  __shared__ __align__(16) float sg[C], sb[C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sg[i] = gamma[i]; sb[i] = beta[i]; }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int lr = lane % L;          // lane within the row
  const int rw = lane / L;          // row within the warp pass
  const long long groups = (M + RPW - 1) / RPW;
  const long long gstride = (long long)gridDim.x * (blockDim.x / 32);
  for (long long g = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); g < groups; g += gstride) {
    const long long row = g * RPW + rw;
    const bool ok = row < M;
    Half8 hv[5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
      hv[i] = ok ? ld_half8(x + row * C + (lr + L * i) * 8) : Half8{};
    float v[5][8];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(hv[i].h[j]);
        v[i][2 * j] = t.x; v[i][2 * j + 1] = t.y;
        s0 += t.x; s1 += t.y;
      }
    float sum = s0 + s1;
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.f / C);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float d0 = v[i][j] - mean, d1 = v[i][j + 1] - mean;
        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
      }
    float q = q0 + q1;
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / C) + eps);
    const float nmr = -mean * rstd;
    if (ok) {
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const int c = (lr + L * i) * 8;
        const float4 g0 = *reinterpret_cast<const float4*>(sg + c), g1 = *reinterpret_cast<const float4*>(sg + c + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(sb + c), b1 = *reinterpret_cast<const float4*>(sb + c + 4);
        Half8 ov;
        ov.h[0] = __floats2half2_rn(fmaf(fmaf(v[i][0], rstd, nmr), g0.x, b0.x), fmaf(fmaf(v[i][1], rstd, nmr), g0.y, b0.y));
        ov.h[1] = __floats2half2_rn(fmaf(fmaf(v[i][2], rstd, nmr), g0.z, b0.z), fmaf(fmaf(v[i][3], rstd, nmr), g0.w, b0.w));
        ov.h[2] = __floats2half2_rn(fmaf(fmaf(v[i][4], rstd, nmr), g1.x, b1.x), fmaf(fmaf(v[i][5], rstd, nmr), g1.y, b1.y));
        ov.h[3] = __floats2half2_rn(fmaf(fmaf(v[i][6], rstd, nmr), g1.z, b1.z), fmaf(fmaf(v[i][7], rstd, nmr), g1.w, b1.w));
        st_half8(y + row * C + c, ov);
      }
    }
  }
}

cudaError_t layernorm(cudaStream_t s, const __half* x, long long M, int C, float eps, const float* gamma,
                      const float* beta, __half* y) {
  ProfScope prof(s, KC_LAYERNORM);
  if (C % 8) return cudaErrorInvalidValue;
  if (C == 320 || C == 640 || C == 1280) {
    const int L = C / 40;
    const long long groups = (M + 32 / L - 1) / (32 / L);
    long long blocks = (groups + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;   // 8 resident blocks per SM, grid-stride over row groups
    if (L == 8) layernorm40_kernel<8><<<(unsigned)blocks, 256, 0, s>>>(x, M, eps, gamma, beta, y);
    else if (L == 16) layernorm40_kernel<16><<<(unsigned)blocks, 256, 0, s>>>(x, M, eps, gamma, beta, y);
    else layernorm40_kernel<32><<<(unsigned)blocks, 256, 0, s>>>(x, M, eps, gamma, beta, y);
    return cudaGetLastError();
  }
  const int vecs = C / 8;
  auto blocks_for = [&](int R) { return (unsigned)((M + 8LL * R - 1) / (8LL * R)); };
  if (vecs <= 32) layernorm_kernel<1, 4><<<blocks_for(4), 256, 0, s>>>(x, M, C, eps, gamma, beta, y);
  else if (vecs <= 64) layernorm_kernel<2, 4><<<blocks_for(4), 256, 0, s>>>(x, M, C, eps, gamma, beta, y);
  else if (vecs <= 160) layernorm_kernel<5, 2><<<blocks_for(2), 256, 0, s>>>(x, M, C, eps, gamma, beta, y);
  else if (vecs <= 320) layernorm_kernel<10, 1><<<blocks_for(1), 256, 0, s>>>(x, M, C, eps, gamma, beta, y);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ layout / elementwise
__global__ void upsample2x_kernel(const __half* __restrict__ x, int H, int W, int C, __half* __restrict__ y,
                                  long long total_vecs) {
  const int vecs = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vecs;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;
    const int ox = (int)(p % (2 * W)); p /= (2 * W);
    const int oy = (int)(p % (2 * H));
    const long long f = p / (2 * H);
    const Half8 hv = ld_half8(x + (((size_t)f * H + oy / 2) * W + ox / 2) * C + v * 8);
    st_half8(y + i * 8, hv);
  }
}
cudaError_t upsample2x(cudaStream_t s, const __half* x, int NF, int H, int W, int C, __half* y) {
  ProfScope prof(s, KC_OTHER);
  if (C % 8) return cudaErrorInvalidValue;
  const long long total = (long long)NF * 4 * H * W * (C / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  upsample2x_kernel<<<blocks, 256, 0, s>>>(x, H, W, C, y, total);
  return cudaGetLastError();
}

__global__ void add_kernel(const __half* __restrict__ a, const __half* __restrict__ b, long long nvec,
                           __half* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const Half8 x = ld_half8(a + i * 8);
    const Half8 z = ld_half8(b + i * 8);
    Half8 o;
#pragma unroll
    for (int j = 0; j < 4; ++j) o.h[j] = __hadd2(x.h[j], z.h[j]);
    st_half8(y + i * 8, o);
  }
}
cudaError_t add_tensors(cudaStream_t s, const __half* a, const __half* b, long long n, __half* y) {
  ProfScope prof(s, KC_OTHER);
  if (n % 8) return cudaErrorInvalidValue;
  const long long nv = n / 8;
  const int blocks = (int)((nv + 255) / 256 < 148 * 16 ? (nv + 255) / 256 : 148 * 16);
  add_kernel<<<blocks, 256, 0, s>>>(a, b, nv, y);
  return cudaGetLastError();
}

__global__ void silu_kernel(const __half* __restrict__ x, long long n, __half* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(silu_f(__half2float(x[i])));
}
cudaError_t silu_copy(cudaStream_t s, const __half* x, long long n, __half* y) {
  ProfScope prof(s, KC_OTHER);
  const int blocks = (int)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  silu_kernel<<<blocks, 256, 0, s>>>(x, n, y);
  return cudaGetLastError();
}

// NCTHW -> tokens. Tile transpose through smem: block handles 32 pixels x up to 32 channels.
template <typename TIn>
__global__ void ncthw_to_tokens_kernel(const TIn* __restrict__ x, int B, int C, int T, int HW, __half* __restrict__ y,
                                       int ldy, float scale) {
  __shared__ float tile[32][33];
  const int bt = blockIdx.z;
  const int b = bt / T, t = bt % T;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int cy = threadIdx.y; cy < 32; cy += blockDim.y) {
    const int c = c0 + cy, p = p0 + threadIdx.x;
    if (c < C && p < HW) tile[cy][threadIdx.x] = (float)x[(((size_t)b * C + c) * T + t) * HW + p];
  }
  __syncthreads();
  for (int py = threadIdx.y; py < 32; py += blockDim.y) {
    const int p = p0 + py, c = c0 + threadIdx.x;
    if (c < C && p < HW) y[((size_t)bt * HW + p) * ldy + c] = __float2half_rn(tile[threadIdx.x][py] * scale);
  }
}
cudaError_t ncthw_to_tokens(cudaStream_t s, const void* x, int is_f32, int B, int C, int T, int HW, __half* y, int ldy,
                            float scale) {
  ProfScope prof(s, KC_OTHER);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B * T), block(32, 8);
  if (is_f32) ncthw_to_tokens_kernel<float><<<grid, block, 0, s>>>((const float*)x, B, C, T, HW, y, ldy, scale);
  else ncthw_to_tokens_kernel<__half><<<grid, block, 0, s>>>((const __half*)x, B, C, T, HW, y, ldy, scale);
  return cudaGetLastError();
}

template <typename TOut>
__global__ void tokens_to_ncthw_kernel(const __half* __restrict__ x, int ldx, int B, int C, int T, int HW,
                                       TOut* __restrict__ y) {
  __shared__ float tile[32][33];
  const int bt = blockIdx.z;
  const int b = bt / T, t = bt % T;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int py = threadIdx.y; py < 32; py += blockDim.y) {
    const int p = p0 + py, c = c0 + threadIdx.x;
    if (c < C && p < HW) tile[py][threadIdx.x] = __half2float(x[((size_t)bt * HW + p) * ldx + c]);
  }
  __syncthreads();
  for (int cy = threadIdx.y; cy < 32; cy += blockDim.y) {
    const int c = c0 + cy, p = p0 + threadIdx.x;
    if (c < C && p < HW) y[(((size_t)b * C + c) * T + t) * HW + p] = (TOut)tile[threadIdx.x][cy];
  }
}
cudaError_t tokens_to_ncthw(cudaStream_t s, const __half* x, int ldx, int B, int C, int T, int HW, void* y, int is_f32) {
  ProfScope prof(s, KC_OTHER);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B * T), block(32, 8);
  if (is_f32) tokens_to_ncthw_kernel<float><<<grid, block, 0, s>>>(x, ldx, B, C, T, HW, (float*)y);
  else tokens_to_ncthw_kernel<__half><<<grid, block, 0, s>>>(x, ldx, B, C, T, HW, (__half*)y);
  return cudaGetLastError();
}

template <typename TIn>
__global__ void add_nchw_residual_kernel(__half* __restrict__ x, int C, int HW, const TIn* __restrict__ r) {
  __shared__ float tile[32][33];
  const int f = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int cy = threadIdx.y; cy < 32; cy += blockDim.y) {
    const int c = c0 + cy, p = p0 + threadIdx.x;
    if (c < C && p < HW) tile[cy][threadIdx.x] = (float)r[((size_t)f * C + c) * HW + p];
  }
  __syncthreads();
  for (int py = threadIdx.y; py < 32; py += blockDim.y) {
    const int p = p0 + py, c = c0 + threadIdx.x;
    if (c < C && p < HW) {
      __half* d = x + ((size_t)f * HW + p) * C + c;
      *d = __float2half_rn(__half2float(*d) + tile[threadIdx.x][py]);
    }
  }
}
cudaError_t add_nchw_residual(cudaStream_t s, __half* x, int NF, int C, int HW, const void* r, int is_f32) {
  ProfScope prof(s, KC_OTHER);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, NF), block(32, 8);
  if (is_f32) add_nchw_residual_kernel<float><<<grid, block, 0, s>>>(x, C, HW, (const float*)r);
  else add_nchw_residual_kernel<__half><<<grid, block, 0, s>>>(x, C, HW, (const __half*)r);
  return cudaGetLastError();
}

// conv_in im2col: one thread per (output pixel, tap); writes Cin values; column = tap*Cin + c; cols >= 9*Cin are zero
template <typename TIn>
__global__ void im2col_latent_kernel(const TIn* __restrict__ x, int B, int Cin, int T, int H, int W,
                                     __half* __restrict__ A) {
  const long long total = (long long)B * T * H * W;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < total; m += (long long)gridDim.x * blockDim.x) {
    long long r = m;
    const int xw = (int)(r % W); r /= W;
    const int yh = (int)(r % H); r /= H;
    const int t = (int)(r % T);
    const int b = (int)(r / T);
    __align__(16) __half row[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) row[i] = __float2half_rn(0.f);
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = yh + tap / 3 - 1, xx = xw + tap % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      for (int c = 0; c < Cin; ++c)
        row[tap * Cin + c] = __float2half_rn((float)x[((((size_t)b * Cin + c) * T + t) * H + yy) * W + xx]);
    }
    uint4* dst = reinterpret_cast<uint4*>(A + m * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = reinterpret_cast<const uint4*>(row)[i];
  }
}
cudaError_t im2col_latent(cudaStream_t s, const void* x, int is_f32, int B, int Cin, int T, int H, int W, __half* A) {
  ProfScope prof(s, KC_OTHER);
  if (9 * Cin > 64) return cudaErrorInvalidValue;
  const long long total = (long long)B * T * H * W;
  const int blocks = (int)((total + 127) / 128);
  if (is_f32) im2col_latent_kernel<float><<<blocks, 128, 0, s>>>((const float*)x, B, Cin, T, H, W, A);
  else im2col_latent_kernel<__half><<<blocks, 128, 0, s>>>((const __half*)x, B, Cin, T, H, W, A);
  return cudaGetLastError();
}

__global__ void sinusoid_kernel(const float* __restrict__ values, int n, int dim, __half* __restrict__ out, int ld) {
  const int half_dim = dim / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * half_dim; i += gridDim.x * blockDim.x) {
    const int r = i / half_dim, k = i % half_dim;
    const float freq = expf(-logf(10000.f) * (float)k / (float)half_dim);
    const float a = values[r] * freq;
    out[(size_t)r * ld + k] = __float2half_rn(cosf(a));             // flip_sin_to_cos: cos first
    out[(size_t)r * ld + half_dim + k] = __float2half_rn(sinf(a));
  }
}
cudaError_t sinusoid(cudaStream_t s, const float* values, int n, int dim, __half* out, int ld) {
  ProfScope prof(s, KC_OTHER);
  sinusoid_kernel<<<(n * dim / 2 + 255) / 256, 256, 0, s>>>(values, n, dim, out, ld);
  return cudaGetLastError();
}

__global__ void expand_rows_kernel(const __half* __restrict__ src, int B, int T, int D, const int* __restrict__ zero_t,
                                   int nzero, int act, __half* __restrict__ out) {
  const int total = B * T * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int d = i % D;
    const int bt = i / D;
    const int b = bt / T, t = bt % T;
    bool z = false;
    for (int k = 0; k < nzero; ++k) z |= (zero_t[k] == t);
    float v = __half2float(src[(size_t)b * D + d]);
    if (act == 1) v = silu_f(v);
    out[i] = __float2half_rn(z ? 0.f : v);
  }
}
cudaError_t expand_rows(cudaStream_t s, const __half* src, int B, int T, int D, const int* zero_t, int nzero, int act,
                        __half* out) {
  ProfScope prof(s, KC_OTHER);
  expand_rows_kernel<<<(B * T * D + 255) / 256, 256, 0, s>>>(src, B, T, D, zero_t, nzero, act, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ VAE decoder helpers
// post_quant_conv (AutoencoderKL.decode, diffusers models/autoencoder_kl.py:275-287): a 1x1 convolution over the 4 latent
// channels, applied before the decoder's 3x3 conv_in (whose zero padding acts on ITS input, so the two cannot be folded).
// y = in_scale * W x + b on [N, C, HW] (NCHW), fp32 out.
template <typename TIn>
__global__ void latent_pointwise_kernel(const TIn* __restrict__ x, int N, int C, int HW, const float* __restrict__ w,
                                        const float* __restrict__ b, float in_scale, float* __restrict__ y) {
  const long long total = (long long)N * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i % HW;
    float v[8];
    for (int k = 0; k < C; ++k) v[k] = (float)x[(n * C + k) * HW + p] * in_scale;
    for (int c = 0; c < C; ++c) {
      float acc = b[c];
      for (int k = 0; k < C; ++k) acc = fmaf(w[c * C + k], v[k], acc);
      y[(n * C + c) * HW + p] = acc;
    }
  }
}
cudaError_t latent_pointwise(cudaStream_t s, const void* x, int is_f32, int N, int C, int HW, const float* w, const float* b,
                             float in_scale, float* y) {
  ProfScope prof(s, KC_OTHER);
  if (C > 8) return cudaErrorInvalidValue;
  const long long total = (long long)N * HW;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  if (is_f32) latent_pointwise_kernel<float><<<blocks, 256, 0, s>>>((const float*)x, N, C, HW, w, b, in_scale, y);
  else latent_pointwise_kernel<__half><<<blocks, 256, 0, s>>>((const __half*)x, N, C, HW, w, b, in_scale, y);
  return cudaGetLastError();
}

// In-place row softmax of an fp16 score matrix [M, N] (N % 8 == 0), fp32 statistics: p = exp(scale (s - max)) / sum. One
// warp per row, the row held in registers for N <= 8192 (three passes over registers, one read + one write of HBM).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(__half* __restrict__ x, long long M, int N, long long ld, float scale_log2) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  __half* r = x + row * ld;
  const int vecs = N / 8;
  constexpr int kMaxV = 32;                      // 32 lanes x 32 vectors x 8 = 8192 columns
  Half8 hv[kMaxV];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    const int vi = lane + i * 32;
    if (vi < vecs) {
      hv[i] = ld_half8(r + vi * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(hv[i].h[j]);
        mx = fmaxf(mx, fmaxf(t.x, t.y));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  const float nm = -mx * scale_log2;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    const int vi = lane + i * 32;
    if (vi < vecs) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 t = __half22float2(hv[i].h[j]);
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(fmaf(t.x, scale_log2, nm)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(fmaf(t.y, scale_log2, nm)));
        const __half2 h2 = __floats2half2_rn(t.x, t.y);
        hv[i].h[j] = h2;
        const float2 back = __half22float2(h2);    // normalise by the sum of what will actually be multiplied
        sum += back.x + back.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    const int vi = lane + i * 32;
    if (vi < vecs) {
      Half8 ov;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(hv[i].h[j]);
        ov.h[j] = __floats2half2_rn(t.x * inv, t.y * inv);
      }
      st_half8(r + vi * 8, ov);
    }
  }
}
cudaError_t softmax_rows(cudaStream_t s, __half* x, long long M, int N, long long ld, float scale) {
  ProfScope prof(s, KC_OTHER);
  if ((N % 8) || N > 8192 || (ld % 8)) return cudaErrorInvalidValue;
  const unsigned blocks = (unsigned)((M + 7) / 8);
  softmax_rows_kernel<<<blocks, 256, 0, s>>>(x, M, N, ld, scale * 1.4426950408889634f);
  return cudaGetLastError();
}

// tokens [B*T*HW, ldx] -> NCTHW with y = clamp(alpha x + beta, lo, hi): the image post-processing of `decode_latents`
// (diffusers pipeline_stable_diffusion_img2img.py:490-492: image / 2 + 0.5, clamp(0, 1)) folded into the layout change.
template <typename TOut>
__global__ void tokens_to_ncthw_affine_kernel(const __half* __restrict__ x, int ldx, int B, int C, int T, int HW,
                                              TOut* __restrict__ y, float alpha, float beta, float lo, float hi) {
  __shared__ float tile[32][33];
  const int bt = blockIdx.z;
  const int b = bt / T, t = bt % T;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int py = threadIdx.y; py < 32; py += blockDim.y) {
    const int p = p0 + py, c = c0 + threadIdx.x;
    if (c < C && p < HW) tile[py][threadIdx.x] = __half2float(x[((size_t)bt * HW + p) * ldx + c]);
  }
  __syncthreads();
  for (int cy = threadIdx.y; cy < 32; cy += blockDim.y) {
    const int c = c0 + cy, p = p0 + threadIdx.x;
    if (c < C && p < HW) y[(((size_t)b * C + c) * T + t) * HW + p] = (TOut)fminf(fmaxf(fmaf(tile[threadIdx.x][cy], alpha, beta), lo), hi);
  }
}
cudaError_t tokens_to_ncthw_affine(cudaStream_t s, const __half* x, int ldx, int B, int C, int T, int HW, void* y, int is_f32,
                                   float alpha, float beta, float lo, float hi) {
  ProfScope prof(s, KC_OTHER);
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B * T), block(32, 8);
  if (grid.x > 65535u * 32u) return cudaErrorInvalidValue;
  if (is_f32) tokens_to_ncthw_affine_kernel<float><<<grid, block, 0, s>>>(x, ldx, B, C, T, HW, (float*)y, alpha, beta, lo, hi);
  else tokens_to_ncthw_affine_kernel<__half><<<grid, block, 0, s>>>(x, ldx, B, C, T, HW, (__half*)y, alpha, beta, lo, hi);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ temporal attention
// ---- tensor-core version: one warp per (batch, pixel, head); T <= 32 frames padded to a 32x32 score tile.
// S = Q K^T and O = P V run on mma.sync m16n8k16 (the problem is 32 x 32 x dp per warp: far too small for tcgen05's
// 128-row tiles), so the kernel is left with its HBM traffic: q,k,v read once, o written once.
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int DP>
__global__ void __launch_bounds__(128)
temporal_attention_mma_kernel(const __half* __restrict__ qkv, int ld, int B, int T, int HW, int heads, int d,
                              float scale_log2, __half* __restrict__ out, int ldo) {
  constexpr int DS = DP + 8;            // padded smem row (keeps ldmatrix rows on distinct banks, 16-byte aligned)
  constexpr int D8 = DP / 8;
  extern __shared__ __align__(16) __half sm_tm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long prob = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long nprob = (long long)B * HW * heads;
  if (prob >= nprob) return;
  const int h = (int)(prob % heads);
  const long long bp = prob / heads;
  const int pix = (int)(bp % HW);
  const int b = (int)(bp / HW);
  // only T rows (+ one row of zeros that every padded row index is clamped to) are staged per operand: the smem
  // footprint, not registers, bounds the resident warps, and the kernel is latency-bound (one problem per warp)
  const int R = T + 1;
  __half* sq = sm_tm + (size_t)warp * 3 * R * DS;
  __half* sk = sq + R * DS;
  __half* sv = sk + R * DS;
  const int hd = heads * DP;
  for (int i = lane; i < R * D8; i += 32) {
    const int t = i / D8, v8 = i % D8;
    uint4 q4 = make_uint4(0, 0, 0, 0), k4 = q4, v4 = q4;
    if (t < T) {
      const __half* row = qkv + (((size_t)b * T + t) * HW + pix) * ld + h * DP + v8 * 8;
      q4 = __ldg(reinterpret_cast<const uint4*>(row));
      k4 = __ldg(reinterpret_cast<const uint4*>(row + hd));
      v4 = __ldg(reinterpret_cast<const uint4*>(row + 2 * hd));
    }
    *reinterpret_cast<uint4*>(sq + t * DS + v8 * 8) = q4;
    *reinterpret_cast<uint4*>(sk + t * DS + v8 * 8) = k4;
    *reinterpret_cast<uint4*>(sv + t * DS + v8 * 8) = v4;
  }
  __syncwarp();
  auto rowc = [&](int r) { return r < T ? r : T; };   // padded rows read the zero row
  // ---- S = Q K^T (32 x 32)
  float sacc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[mt][nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < DP / 16; ++ks) {
    uint32_t a[2][4], bk[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) ldsm_x4(a[mt], sq + rowc(mt * 16 + (lane & 15)) * DS + ks * 16 + (lane >> 4) * 8);
#pragma unroll
    for (int np = 0; np < 2; ++np)   // two key tiles (16 keys) per ldmatrix.x4
      ldsm_x4(bk[np], sk + rowc(np * 16 + (lane & 7) + ((lane >> 4) << 3)) * DS + ks * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma16816(sacc[mt][nt], a[mt], bk[nt >> 1][(nt & 1) * 2], bk[nt >> 1][(nt & 1) * 2 + 1]);
  }
  // ---- softmax over the key axis (rows: mt*16 + lane/4 and +8; columns nt*8 + 2*(lane%4) + {0,1})
  float inv[2][2];
  uint32_t pa[2][2][4];   // P as A fragments: [m tile][key step of 16]
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = nt * 8 + 2 * (lane & 3) + (e & 1);
        if (col >= T) sacc[mt][nt][e] = -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], sacc[mt][nt][e]);
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float pv;   // SFU approximation, flush-to-zero: exp2(-inf) = 0 for the masked keys
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv) : "f"((sacc[mt][nt][e] - mx[e >> 1]) * scale_log2));
        sacc[mt][nt][e] = pv;
        sum[e >> 1] += pv;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
      sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv[mt][r]) : "f"(sum[r]));
    }
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
      pa[mt][k2][0] = pack_h2(sacc[mt][2 * k2][0], sacc[mt][2 * k2][1]);
      pa[mt][k2][1] = pack_h2(sacc[mt][2 * k2][2], sacc[mt][2 * k2][3]);
      pa[mt][k2][2] = pack_h2(sacc[mt][2 * k2 + 1][0], sacc[mt][2 * k2 + 1][1]);
      pa[mt][k2][3] = pack_h2(sacc[mt][2 * k2 + 1][2], sacc[mt][2 * k2 + 1][3]);
    }
  }
  __syncwarp();   // all lanes are done reading Q: its buffer becomes the output staging area
  // ---- O = P V, 16 output columns per step
#pragma unroll
  for (int n0 = 0; n0 < DP; n0 += 16) {
    float oacc[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) oacc[mt][nt][e] = 0.f;
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
      uint32_t bv[4];
      ldsm_x4_trans(bv, sv + rowc(k2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * DS + n0 + (lane >> 4) * 8);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma16816(oacc[mt][0], pa[mt][k2], bv[0], bv[1]);
        mma16816(oacc[mt][1], pa[mt][k2], bv[2], bv[3]);
      }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = n0 + nt * 8 + 2 * (lane & 3);
        const int r_lo = mt * 16 + (lane >> 2);
        if (r_lo < T)
          *reinterpret_cast<uint32_t*>(sq + r_lo * DS + col) = pack_h2(oacc[mt][nt][0] * inv[mt][0], oacc[mt][nt][1] * inv[mt][0]);
        if (r_lo + 8 < T)
          *reinterpret_cast<uint32_t*>(sq + (r_lo + 8) * DS + col) = pack_h2(oacc[mt][nt][2] * inv[mt][1], oacc[mt][nt][3] * inv[mt][1]);
      }
  }
  __syncwarp();
  const int dv = d / 8;
  for (int i = lane; i < T * dv; i += 32) {
    const int t = i / dv, v8 = i % dv;
    *reinterpret_cast<uint4*>(out + (((size_t)b * T + t) * HW + pix) * ldo + h * d + v8 * 8) =
        *reinterpret_cast<const uint4*>(sq + t * DS + v8 * 8);
  }
}

cudaError_t temporal_attention(cudaStream_t s, const __half* qkv, int ld, int B, int T, int HW, int heads, int d, int dp,
                               float scale, __half* out, int ldo) {
  ProfScope prof(s, KC_TEMPORAL_ATTN);
  if (T > 32 || T < 1 || (d % 8) || (dp % 16) || dp < d) return cudaErrorInvalidValue;
  const long long nprob = (long long)B * HW * heads;
  const int wpb = 4;
  const unsigned blocks = (unsigned)((nprob + wpb - 1) / wpb);
  const float sl2 = scale * 1.4426950408889634f;
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
#define MVB_TAM(N)                                                                                                  \
  case N: {                                                                                                          \
    const size_t smem = (size_t)wpb * 3 * (T + 1) * (N + 8) * sizeof(__half);                                        \
    static size_t set##N[64] = {};   /* per device */                                                                \
    if (smem > set##N[cur_dev & 63]) {                                                                               \
      cudaFuncSetAttribute(temporal_attention_mma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      set##N[cur_dev & 63] = smem;                                                                                   \
    }                                                                                                                \
    temporal_attention_mma_kernel<N><<<blocks, wpb * 32, smem, s>>>(qkv, ld, B, T, HW, heads, d, sl2, out, ldo);    \
    break;                                                                                                           \
  }
  switch (dp) {
    MVB_TAM(16) MVB_TAM(32) MVB_TAM(48) MVB_TAM(64) MVB_TAM(80) MVB_TAM(96) MVB_TAM(160)
    default: return cudaErrorInvalidValue;
  }
#undef MVB_TAM
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ step epilogue
template <typename TLat>
__global__ void fuse_cfg_ddim_kernel(const float* __restrict__ eps_sum, const float* __restrict__ counter,
                                     const TLat* __restrict__ lat_in, TLat* __restrict__ lat_out, int B, int C, int T,
                                     int HW, int cfg, float g, float a_t, float a_prev, int pred, float clip,
                                     int use_clipped, float std_dev, const float* __restrict__ noise,
                                     float* __restrict__ eps_out, float* __restrict__ x0_out) {
  const long long n = (long long)B * C * T * HW;
  const float sa = sqrtf(a_t), sb = sqrtf(1.f - a_t), sap = sqrtf(a_prev);
  const float sdir = sqrtf(fmaxf(1.f - a_prev - std_dev * std_dev, 0.f));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)((i / HW) % T);
    const float cnt = counter ? counter[t] : 1.f;
    float e = eps_sum[i] / cnt;                       // uncond half comes first (prompt_embeds = [neg, pos])
    if (cfg) {
      const float tx = eps_sum[n + i] / cnt;
      e = e + g * (tx - e);
    }
    const float x = (float)lat_in[i];
    float x0, eps;
    if (pred == 0) { x0 = (x - sb * e) / sa; eps = e; }
    else if (pred == 1) { x0 = sa * x - sb * e; eps = sa * e + sb * x; }      // v_prediction
    else { x0 = e; eps = (x - sa * x0) / sb; }                                // sample
    if (clip > 0.f) x0 = fminf(fmaxf(x0, -clip), clip);
    if (use_clipped) eps = (x - sa * x0) / sb;
    float prev = sap * x0 + sdir * eps;
    if (noise) prev += std_dev * noise[i];
    lat_out[i] = (TLat)prev;
    if (eps_out) eps_out[i] = e;
    if (x0_out) x0_out[i] = x0;
  }
}
cudaError_t fuse_cfg_ddim(cudaStream_t s, const float* eps_sum, const float* counter, const void* latents_in,
                          void* latents_out, int is_f32, int B, int C, int T, int HW, int cfg, float guidance,
                          float alpha_t, float alpha_prev, int prediction_type, float clip_range, int use_clipped,
                          float std_dev, const float* noise, float* eps_out, float* x0_out) {
  ProfScope prof(s, KC_OTHER);
  const long long n = (long long)B * C * T * HW;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  if (is_f32)
    fuse_cfg_ddim_kernel<float><<<blocks, 256, 0, s>>>(eps_sum, counter, (const float*)latents_in, (float*)latents_out,
                                                       B, C, T, HW, cfg, guidance, alpha_t, alpha_prev, prediction_type,
                                                       clip_range, use_clipped, std_dev, noise, eps_out, x0_out);
  else
    fuse_cfg_ddim_kernel<__half><<<blocks, 256, 0, s>>>(eps_sum, counter, (const __half*)latents_in,
                                                        (__half*)latents_out, B, C, T, HW, cfg, guidance, alpha_t,
                                                        alpha_prev, prediction_type, clip_range, use_clipped, std_dev,
                                                        noise, eps_out, x0_out);
  return cudaGetLastError();
}

// Overlap mean + CFG + an AFFINE sampler update in one pass: every eps-linear sampler step without clipping is
//   x_prev = c_x x + c_e eps + c_n noise     (DDIM eta >= 0, Euler discrete incl. churn, LCM; coefficients from the host)
// and an optional second affine output aux = a_x x + a_e eps (pred_original_sample / LCM's `denoised`).
template <typename TLat>
__global__ void fuse_cfg_affine_kernel(const float* __restrict__ eps_sum, const float* __restrict__ counter,
                                       const TLat* __restrict__ lat_in, TLat* __restrict__ lat_out, long long n, int T, int HW,
                                       int cfg, float g, float c_x, float c_e, float c_n, const float* __restrict__ noise,
                                       float a_x, float a_e, float* __restrict__ aux_out, float* __restrict__ eps_out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)((i / HW) % T);
    const float cnt = counter ? counter[t] : 1.f;
    float e = eps_sum[i] / cnt;
    if (cfg) {
      const float tx = eps_sum[n + i] / cnt;
      e = e + g * (tx - e);
    }
    const float x = (float)lat_in[i];
    float prev = fmaf(c_x, x, c_e * e);
    if (noise) prev = fmaf(c_n, noise[i], prev);
    lat_out[i] = (TLat)prev;
    if (aux_out) aux_out[i] = fmaf(a_x, x, a_e * e);
    if (eps_out) eps_out[i] = e;
  }
}
cudaError_t fuse_cfg_affine(cudaStream_t s, const float* eps_sum, const float* counter, const void* latents_in,
                            void* latents_out, int is_f32, int B, int C, int T, int HW, int cfg, float guidance, float c_x,
                            float c_e, float c_n, const float* noise, float a_x, float a_e, float* aux_out, float* eps_out) {
  ProfScope prof(s, KC_OTHER);
  const long long n = (long long)B * C * T * HW;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  if (is_f32)
    fuse_cfg_affine_kernel<float><<<blocks, 256, 0, s>>>(eps_sum, counter, (const float*)latents_in, (float*)latents_out, n, T,
                                                         HW, cfg, guidance, c_x, c_e, c_n, noise, a_x, a_e, aux_out, eps_out);
  else
    fuse_cfg_affine_kernel<__half><<<blocks, 256, 0, s>>>(eps_sum, counter, (const __half*)latents_in, (__half*)latents_out, n,
                                                          T, HW, cfg, guidance, c_x, c_e, c_n, noise, a_x, a_e, aux_out, eps_out);
  return cudaGetLastError();
}

template <typename TIn>
__global__ void accumulate_window_kernel(float* __restrict__ eps_sum, int B2, int C, int T, int HW,
                                         const TIn* __restrict__ win, int Tw, int src_t0, const int* __restrict__ frames,
                                         int nframes) {
  const long long n = (long long)B2 * C * nframes * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    long long r = i / HW;
    const int k = (int)(r % nframes); r /= nframes;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    const float v = (float)win[(((size_t)b * C + c) * Tw + src_t0 + k) * HW + p];
    eps_sum[(((size_t)b * C + c) * T + frames[k]) * HW + p] += v;
  }
}
cudaError_t accumulate_window(cudaStream_t s, float* eps_sum, int B2, int C, int T, int HW, const void* eps_win,
                              int is_f32, int Tw, int src_t0, const int* frames_dev, int nframes) {
  ProfScope prof(s, KC_OTHER);
  const long long n = (long long)B2 * C * nframes * HW;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  if (is_f32)
    accumulate_window_kernel<float><<<blocks, 256, 0, s>>>(eps_sum, B2, C, T, HW, (const float*)eps_win, Tw, src_t0,
                                                           frames_dev, nframes);
  else
    accumulate_window_kernel<__half><<<blocks, 256, 0, s>>>(eps_sum, B2, C, T, HW, (const __half*)eps_win, Tw, src_t0,
                                                            frames_dev, nframes);
  return cudaGetLastError();
}

}  // namespace mvb
