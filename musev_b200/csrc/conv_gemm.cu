// tcgen05 / TMA implicit-GEMM kernel. See conv_gemm.cuh for the math and the reference call sites.
//
// CTA = 384 threads, persistent over output tiles (128 rows x block_n columns; 256 rows on a CTA pair):
//   warp 0   : TMA producer  -- per K block (64 channels of one tap) loads the A box {64, bw, bh, bn}
//              (out-of-image taps are zero-filled by TMA = conv padding) and the weight tile {64, block_n};
//              warp-uniform loop, elect.sync picks the issuing lane (a single-thread loop was the mainloop bottleneck)
//   warp 1   : MMA issuer    -- 4 x tcgen05.mma (K=16 each) per K block into TMEM, same warp-uniform structure
//   warp 2   : TMEM allocator (512 columns = 2 accumulator stages x up to 256 fp32 columns)
//   warps 4-11: epilogue     -- two warps per TMEM lane quarter, alternating 32-column chunks: tcgen05.ld accumulator
//              rows -> bias / time-embedding / residual / GEGLU (four template variants) -> fp16 -> swizzled smem
//              staging -> TMA store (4-deep staging ring per column half, one named barrier per chunk)
// Pipelines: smem operand ring (full/empty mbarriers, 160 KB cut into 3..8 stages per launch) and the TMEM double
// buffer (tmem_full/tmem_empty). kTwoCta: cluster of 2 with cta_group::2 MMAs (M = 256), used for K >= 1280.
#include "conv_gemm.cuh"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "ptx.cuh"
#include "stats.cuh"

namespace mvb {

static constexpr int kMaxStages = 8;   // ring depth is chosen per launch: as many (16 KB A + block_n x 128 B) stages as fit
static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;
static constexpr int kMaxBlockN = 256;
static constexpr int kABytes = kBlockM * kBlockK * 2;        // 16 KB
static constexpr int kStagingBufBytes = 128 * 64;               // 128 rows x 32 fp16 output columns
static constexpr int kStagingDepth = 4;                         // staging buffers per column half (3 TMA stores in flight)
static constexpr int kStagingBytes = 2 * kStagingDepth * kStagingBufBytes;
static constexpr int kRingBytes = 156 * 1024;                 // operand ring, split into nstages stages
static constexpr int kBiasBytes = 8 * 128 * 4;                // per epilogue warp: the bias of its (up to) 128 accumulator columns of a tile
static constexpr int kSmemBytes = kRingBytes + kStagingBytes + kBiasBytes + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// exact-erf GELU (diffusers activations.py:89-102 uses F.gelu default) with erf from Abramowitz-Stegun 7.1.26
// (|error| < 1.5e-7, far below the fp16 output rounding): one reciprocal, one exp2, a degree-5 polynomial.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-poly, e, 1.f);
  const float erf_v = copysignf(erf_abs, x);
  return 0.5f * x * (1.f + erf_v);
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// value * gelu_erf(gate) for two columns at once. A&S 7.1.26: erfc(z) = t P(t) exp(-z^2), t = 1 / (1 + p z), z >= 0.
// With u = |x| and w = 0.5 u erfc(u / sqrt 2) >= 0, gelu(x) = max(x, 0) - w for either sign of x. The reciprocal is
// the SFU approximation (its argument lies in [1, inf), no special cases), -0.5 is folded into the coefficients.
__device__ __forceinline__ F2 geglu2(F2 val, F2 gate) {
  float g0, g1;
  f2_get(gate, g0, g1);
  const F2 u = f2_make(fabsf(g0), fabsf(g1));
  float d0, d1;
  f2_get(f2_fma(u, f2_make(0.231641888f, 0.231641888f), f2_make(1.f, 1.f)), d0, d1);   // 1 + p u / sqrt 2
  const F2 t = f2_make(rcp_approx(d0), rcp_approx(d1));
  F2 poly = f2_fma(t, f2_make(-0.5307027145f, -0.5307027145f), f2_make(0.7265760135f, 0.7265760135f));
  poly = f2_fma(poly, t, f2_make(-0.7107068705f, -0.7107068705f));
  poly = f2_fma(poly, t, f2_make(0.142248368f, 0.142248368f));
  poly = f2_fma(poly, t, f2_make(-0.127414796f, -0.127414796f));
  poly = f2_mul(poly, t);                                                                 // -0.5 t P(t)
  float a0, a1;
  f2_get(f2_mul(f2_mul(u, f2_make(-0.72134752044f, -0.72134752044f)), u), a0, a1);        // -u^2 / 2 * log2 e
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const F2 gelu = f2_fma(u, f2_mul(poly, f2_make(e0, e1)), f2_make(fmaxf(g0, 0.f), fmaxf(g1, 0.f)));
  return f2_mul(val, gelu);
}

// Epilogue variants (template parameter kEpi). The generic one takes every option at run time; the three fast ones
// cover the shapes that are epilogue-bound in the UNet (K <= 640) with packed f32x2 arithmetic and no per-element
// branches: ~2-3 instructions per output element instead of ~14.
enum : int {
  kEpiGeneric = 0,
  kEpiPlain = 1,      // fp16 out, bias / row-add optional, alpha == 1, no residual, no activation, N % 32 == 0
  kEpiResidual = 2,   // fp16 out, alpha * (acc + bias) + residual (beta == 1), N % 32 == 0
  kEpiGeglu = 3       // fp16 out, value * gelu(gate) on the packed [16 value | 16 gate] column layout
};

// kTwoCta: the kernel runs as CTA pairs (cluster of 2, tcgen05 cta_group::2): one 256 x block_n tile per pair, each CTA
// stages its own 128 A rows and HALF of the weight tile, the leader issues M=256 MMAs that read both halves. This cuts
// the L2->SM operand traffic per FLOP (the measured bound of the 1-CTA kernel, profiles/r01_ncu_full_summary.txt) and
// halves the MMA issue count.
struct AMaps {
  CUtensorMap m[4];   // activation views: [0] source 0, [1] skip-concat source / stride-2 phases 1..3
};

template <bool kTwoCta, int kEpi>
__global__ void __launch_bounds__(384, 1)
conv_gemm_kernel(const __grid_constant__ AMaps tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC,
                 const __grid_constant__ ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + kRingBytes;
  float* sbias = reinterpret_cast<float*>(staging + kStagingBytes);      // [8 epilogue warps][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes + kBiasBytes);
  uint64_t* full = bars;                       // [kMaxStages]
  uint64_t* empty = bars + kMaxStages;         // [kMaxStages]
  uint64_t* tfull = bars + 2 * kMaxStages;     // [2]
  uint64_t* tempty = bars + 2 * kMaxStages + 2;// [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  const int nstages = p.nstages;
  const int stage_bytes = p.stage_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA.m[0]);
    tma_prefetch_desc(&tmA.m[1]);
    tma_prefetch_desc(&tmA.m[2]);
    tma_prefetch_desc(&tmA.m[3]);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kTwoCta ? 16 : 8);   // every epilogue warp of the pair reports to the leader
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kTwoCta) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (kTwoCta) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kb_per_tap = p.kb0 + p.kb1;
  const int num_kb = p.ntaps * kb_per_tap;
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
  // work units: a unit is one tile (1-CTA) or a pair of vertically adjacent tiles (2-CTA), n fastest
  const uint32_t crank = kTwoCta ? cluster_ctarank() : 0u;
  const int unit0 = kTwoCta ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = kTwoCta ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_units = (kTwoCta ? (tiles_m + 1) / 2 : tiles_m) * p.tiles_nn;
  const int b_rows = kTwoCta ? p.block_n / 2 : p.block_n;      // weight rows staged by this CTA
  const uint32_t stage_tx = kABytes + (uint32_t)b_rows * kBlockK * 2;

  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues the copies of a stage
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < num_units; unit += unit_stride) {
        const int nt = unit % p.tiles_nn;
        const int mt = kTwoCta ? 2 * (unit / p.tiles_nn) + (int)crank : unit / p.tiles_nn;
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int tn = mt / (p.tiles_w * p.tiles_h);   // may run past the last frame in the odd tail of a pair: TMA zero-fills
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        int kcol = 0;                                   // K coordinate of the weight tile
        const int ncoord = nt * p.block_n + (kTwoCta ? (int)crank * b_rows : 0);
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int xw = w0 + p.dx[tap], yh = h0 + p.dy[tap];
          const int src0 = p.tap_src[tap];
          for (int cb = 0; cb < kb_per_tap; ++cb, kcol += kBlockK) {
            const bool first = cb < p.kb0;
            const int src = src0 + (first ? 0 : 1);
            const int c0 = (first ? cb : cb - p.kb0) * kBlockK;
            const CUtensorMap* tm = &tmA.m[src];
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              uint8_t* sa = smem + stage * stage_bytes;
              uint8_t* sb = sa + kABytes;
              if (kTwoCta) {
                // both CTAs' copies complete on the LEADER's barrier, which expects the bytes of the whole pair
                if (crank == 0) mbar_expect_tx(&full[stage], 2 * stage_tx);
                tma_load_4d_2sm(sa, tm, &full[stage], c0, xw, yh, n0);
                tma_load_2d_2sm(sb, &tmB, &full[stage], kcol, ncoord);
              } else {
                mbar_expect_tx(&full[stage], stage_tx);
                tma_load_4d(sa, tm, &full[stage], c0, xw, yh, n0);
                tma_load_2d(sb, &tmB, &full[stage], kcol, ncoord);
              }
            }
            __syncwarp();
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issue: the whole warp walks the loop (warp-uniform control flow keeps the operands in uniform registers and
    // avoids the per-instruction lane-election loops a divergent single-thread region needs); one elected lane issues.
    if (crank == 0) {
      const uint32_t idesc = make_idesc_f16(kTwoCta ? 2 * kBlockM : kBlockM, p.block_n, 0, 0);
      const uint64_t desc0_a = make_desc_k_sw128(smem_u32(smem));
      const uint64_t desc0_b = make_desc_k_sw128(smem_u32(smem) + kABytes);
      const uint32_t stage_step = (uint32_t)stage_bytes >> 4;     // descriptor address field counts 16-byte units
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int unit = unit0; unit < num_units; unit += unit_stride) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * kMaxBlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = desc0_a + (uint64_t)((uint32_t)stage * stage_step);
          const uint64_t db = desc0_b + (uint64_t)((uint32_t)stage * stage_step);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // advance 16 K-elements = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
              if (kTwoCta) umma_f16_ss_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
              else umma_f16_ss(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            if (kTwoCta) umma_commit_2sm(&empty[stage], 3); else umma_commit(&empty[stage]);
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
          if (kTwoCta) umma_commit_2sm(&tfull[as], 3); else umma_commit(&tfull[as]);
        }
        __syncwarp();
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // 8 epilogue warps: warp (q, half) owns TMEM lanes [32q, 32q+32) and every other group of output columns.
    // fp16 results leave through smem staging (64-byte swizzle) + TMA store: full-line coalesced writes, clipped by the
    // tensor map at the image / matrix edges, so the store path needs no masks.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r = q * 32 + lane;     // accumulator row
    int as = 0;
    uint32_t aphase = 0;
    const bool use_res = (p.res != nullptr) && !p.geglu && !p.out_f32;
    const bool issuer = (q == 0) && (lane == 0);
    const int acc_step = p.geglu ? 64 : 32;             // accumulator columns consumed per 32 output columns
    uint32_t chunk_iter = 0;
    for (int unit = unit0; unit < num_units; unit += unit_stride) {
      const int nt = unit % p.tiles_nn;
      const int mt = kTwoCta ? 2 * (unit / p.tiles_nn) + (int)crank : unit / p.tiles_nn;
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int rw = r % p.bw;
      const int rh = (r / p.bw) % p.bh;
      const int rn = r / (p.bw * p.bh);
      const int w = tw * p.bw + rw, h = th * p.bh + rh, n = tn * p.bn + rn;
      const bool row_ok = (w < p.W) && (h < p.H) && (n < p.NF);
      const long long m = ((long long)n * p.H + h) * p.W + w;
      const float* radd = (p.rowadd && row_ok) ? p.rowadd + (long long)(m / p.rows_per_group) * p.ld_rowadd : nullptr;
      const int ncol0 = nt * p.block_n;
      const __half* res_row = use_res ? p.res + m * p.ld_res : nullptr;

      uint4 rcur[4] = {}, rnxt[4] = {};
      auto load_res = [&](int c0, uint4 (&dst)[4]) {
        if (!use_res || !row_ok) return;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int nn = ncol0 + c0 + g * 8;
          if (c0 + g * 8 < p.block_n && nn < p.N) dst[g] = __ldg(reinterpret_cast<const uint4*>(res_row + nn));
        }
      };
      load_res(half * acc_step, rcur);
      // Bias of this warp's accumulator columns of the tile -> its private smem slice, issued BEFORE the accumulator wait so
      // that the global-load latency hides behind the mainloop. (ncu on the level-0 N = K = 320 + residual GEMM: 20 % of
      // all samples were long-scoreboard stalls on the per-chunk bias LDGs, which every thread issued and consumed at once.)
      float* wbias = sbias + (warp - 4) * 128;
      if constexpr (kEpi != kEpiGeneric) {
        __syncwarp();                                   // the previous tile's reads of this slice are done
#pragma unroll
        for (int k = 0; k < 128 / 32; ++k) {
          const int cw = k * 32 + lane;                 // position inside this warp's column list
          const int cc = half * acc_step + (cw / acc_step) * 2 * acc_step + (cw % acc_step);   // accumulator column in the tile
          float bv = 0.f;
          if (p.bias && cc < p.block_n && ncol0 + cc < p.N) bv = __ldg(p.bias + ncol0 + cc);
          wbias[cw] = bv;
        }
        __syncwarp();
      }

      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * kMaxBlockN;
      for (int c0 = half * acc_step; c0 < p.block_n; c0 += 2 * acc_step) {
        if constexpr (kEpi != kEpiGeneric) {
          const int nbase = ncol0 + c0;
          uint32_t v[32], vg[32];
          tmem_ld32(t_row + c0, v);
          if constexpr (kEpi == kEpiGeglu) tmem_ld32(t_row + c0 + 32, vg);
          if constexpr (kEpi == kEpiResidual) load_res(c0 + 2 * acc_step, rnxt);
          tmem_ld_wait();
          if (nbase < p.N) {     // uniform per column half: the skipped chunks skip their barrier as a group
            uint8_t* buf = staging + (half * kStagingDepth + (chunk_iter % kStagingDepth)) * kStagingBufBytes;
            const F2 alpha2 = f2_make(p.alpha, p.alpha);
            const float* cbias = wbias + ((c0 - half * acc_step) / (2 * acc_step)) * acc_step;   // this chunk's staged bias
#pragma unroll
            for (int g = 0; g < 4; ++g) {      // 8 output columns = one 16-byte staging store
              uint32_t o[4];
              if constexpr (kEpi == kEpiGeglu) {
                // output columns 8g..8g+7 of this chunk: accumulator block g/2 (va | vb), value j, gate 16 + j
                const uint32_t* vv = (g < 2) ? v : vg;
                const int j0 = (g & 1) * 8;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int j = j0 + 2 * e;
                  // staged bias: this chunk's 64 accumulator columns start at cbias; block g/2 holds [16 value | 16 gate]
                  const float2 bv = *reinterpret_cast<const float2*>(cbias + (g >> 1) * 32 + j);
                  const float2 bg = *reinterpret_cast<const float2*>(cbias + (g >> 1) * 32 + 16 + j);
                  const F2 val = f2_add(f2_make(__uint_as_float(vv[j]), __uint_as_float(vv[j + 1])), f2_make(bv.x, bv.y));
                  const F2 gat = f2_add(f2_make(__uint_as_float(vv[16 + j]), __uint_as_float(vv[16 + j + 1])),
                                        f2_make(bg.x, bg.y));
                  float x0, x1;
                  f2_get(geglu2(val, gat), x0, x1);
                  const __half2 h2 = __floats2half2_rn(x0, x1);
                  o[e] = *reinterpret_cast<const uint32_t*>(&h2);
                }
              } else {
                float4 b0 = *reinterpret_cast<const float4*>(cbias + 8 * g);
                float4 b1 = *reinterpret_cast<const float4*>(cbias + 8 * g + 4);
                if (radd) {
                  const float4 a0 = __ldg(reinterpret_cast<const float4*>(radd + nbase) + 2 * g);
                  const float4 a1 = __ldg(reinterpret_cast<const float4*>(radd + nbase) + 2 * g + 1);
                  b0.x += a0.x; b0.y += a0.y; b0.z += a0.z; b0.w += a0.w;
                  b1.x += a1.x; b1.y += a1.y; b1.z += a1.z; b1.w += a1.w;
                }
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const __half2* rh2 = reinterpret_cast<const __half2*>(&rcur[g]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int j = g * 8 + 2 * e;
                  F2 x = f2_add(f2_make(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), f2_make(bb[2 * e], bb[2 * e + 1]));
                  if constexpr (kEpi == kEpiResidual) {
                    const float2 rr = __half22float2(rh2[e]);
                    x = f2_fma(x, alpha2, f2_make(rr.x, rr.y));
                  }
                  float x0, x1;
                  f2_get(x, x0, x1);
                  const __half2 h2 = __floats2half2_rn(x0, x1);
                  o[e] = *reinterpret_cast<const uint32_t*>(&h2);
                }
              }
              *reinterpret_cast<uint4*>(buf + r * 64 + ((g ^ ((r >> 1) & 3)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
            fence_proxy_async();
            if (issuer) tma_store_wait_read<kStagingDepth - 2>();
            named_bar_sync(1 + half, 128);
            if (issuer) {
              const int oc = (kEpi == kEpiGeglu) ? nbase / 2 : nbase;
              tma_store_4d(&tmC, buf, oc, tw * p.bw, th * p.bh, tn * p.bn);
              tma_store_commit();
            }
            ++chunk_iter;
          }
        } else {
          // accumulator -> f[32] = the 32 output columns of this chunk, bias / row-add applied
          float f[32];
          const int nbase = ncol0 + c0;
          if (p.geglu) {
            uint32_t va[32], vb[32];
            tmem_ld32(t_row + c0, va);
            tmem_ld32(t_row + c0 + 32, vb);
            tmem_ld_wait();
            // packed columns: [16 value | 16 gate] per 32 accumulator columns
  #pragma unroll
            for (int hsel = 0; hsel < 2; ++hsel) {
              const uint32_t* v = hsel ? vb : va;
              const int nb = nbase + hsel * 32;
  #pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                // bias of 4 value columns and their 4 gate columns (nb is a multiple of 32: 16-byte aligned float4)
                const float4 bv = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + nb) + j4) : make_float4(0, 0, 0, 0);
                const float4 bg = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + nb + 16) + j4) : make_float4(0, 0, 0, 0);
                const int j = j4 * 4;
                const F2 v01 = f2_add(f2_make(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), f2_make(bv.x, bv.y));
                const F2 v23 = f2_add(f2_make(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), f2_make(bv.z, bv.w));
                const F2 g01 = f2_add(f2_make(__uint_as_float(v[16 + j]), __uint_as_float(v[16 + j + 1])), f2_make(bg.x, bg.y));
                const F2 g23 = f2_add(f2_make(__uint_as_float(v[16 + j + 2]), __uint_as_float(v[16 + j + 3])), f2_make(bg.z, bg.w));
                f2_get(geglu2(v01, g01), f[hsel * 16 + j], f[hsel * 16 + j + 1]);
                f2_get(geglu2(v23, g23), f[hsel * 16 + j + 2], f[hsel * 16 + j + 3]);
              }
            }
          } else {
            uint32_t v[32];
            tmem_ld32(t_row + c0, v);      // block_n is a multiple of 32
            load_res(c0 + 2 * acc_step, rnxt);
            tmem_ld_wait();
            if (nbase + 32 <= p.N) {
  #pragma unroll
              for (int g = 0; g < 8; ++g) {
                float4 b = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + nbase) + g) : make_float4(0, 0, 0, 0);
                if (radd) {
                  const float4 a4 = __ldg(reinterpret_cast<const float4*>(radd + nbase) + g);
                  b.x += a4.x; b.y += a4.y; b.z += a4.z; b.w += a4.w;
                }
                f[g * 4 + 0] = __uint_as_float(v[g * 4 + 0]) + b.x;
                f[g * 4 + 1] = __uint_as_float(v[g * 4 + 1]) + b.y;
                f[g * 4 + 2] = __uint_as_float(v[g * 4 + 2]) + b.z;
                f[g * 4 + 3] = __uint_as_float(v[g * 4 + 3]) + b.w;
              }
            } else {
  #pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int nn = nbase + j;
                float x = __uint_as_float(v[j]);
                if (nn < p.N) {
                  if (p.bias) x += __ldg(p.bias + nn);
                  if (radd) x += __ldg(radd + nn);
                }
                f[j] = x;
              }
            }
          }
          if (p.out_f32) {
            if (row_ok) {
  #pragma unroll
              for (int g = 0; g < 8; ++g) {
                const int nn = nbase + g * 4;
                if (nn < p.N) {
                  float4 o4;
                  o4.x = f[g * 4 + 0] * p.alpha; o4.y = f[g * 4 + 1] * p.alpha;
                  o4.z = f[g * 4 + 2] * p.alpha; o4.w = f[g * 4 + 3] * p.alpha;
                  if (p.act == 1) { o4.x = silu(o4.x); o4.y = silu(o4.y); o4.z = silu(o4.z); o4.w = silu(o4.w); }
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + m * p.ldc + nn) = o4;
                }
              }
            }
          } else {
            // finish in fp32, round to fp16, stage, store
            // staging ring of kStagingDepth buffers, ONE barrier per chunk: chunk i is written while stores i-1..i-3 drain
            // (a TMA store takes ~1 us to release its smem source); before barrier i the issuer makes sure store
            // i-(depth-1) is done, so after the barrier everybody knows the buffer of chunk i+1 is free
            uint8_t* buf = staging + (half * kStagingDepth + (chunk_iter % kStagingDepth)) * kStagingBufBytes;
  #pragma unroll
            for (int g = 0; g < 4; ++g) {
              __align__(16) __half o[8];
              const __half* rh8 = reinterpret_cast<const __half*>(&rcur[g]);
  #pragma unroll
              for (int j = 0; j < 8; ++j) {
                float x = f[g * 8 + j];
                if (!p.geglu) {
                  x *= p.alpha;
                  if (use_res) x = fmaf(p.beta, __half2float(rh8[j]), x);
                  if (p.act == 1) x = silu(x);
                }
                o[j] = __float2half_rn(x);
              }
              *reinterpret_cast<uint4*>(buf + r * 64 + ((g ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(o);
            }
            fence_proxy_async();
            if (issuer) tma_store_wait_read<kStagingDepth - 2>();
            named_bar_sync(1 + half, 128);
            if (issuer) {
              const int oc = p.geglu ? nbase / 2 : nbase;
              tma_store_4d(&tmC, buf, oc, tw * p.bw, th * p.bh, tn * p.bn);
              tma_store_commit();
            }
            ++chunk_iter;
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) rcur[g] = rnxt[g];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kTwoCta) mbar_arrive_remote(&tempty[as], 0); else mbar_arrive(&tempty[as]);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (issuer) tma_store_wait_all();   // smem must stay valid until the last bulk store has read it
  }

  tc_fence_before();
  if (kTwoCta) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    if (kTwoCta) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) fn = reinterpret_cast<EncodeTiledFn>(dlsym(h, "cuTensorMapEncodeTiled"));
  }
  return fn;
}

// ---- encoded tensor maps are memoized. A CUtensorMap is a pure function of (address, extents, strides, box, swizzle); the engine's
// bump arena hands out the same addresses for the same shapes on every forward, so after the first window-step every one of the
// ~2 500 cuTensorMapEncodeTiled calls of a forward (2-6 per GEMM, 5 per attention; VERDICT r01 "host encodes tensor maps per launch")
// becomes a hash lookup. Open addressing on 2^14 slots, cleared wholesale when 3/4 full (callers that stream fresh torch
// buffers through the op-level ABI only ever cost themselves re-encodes). MVB_TMAP_CACHE=0 switches it off (A/B runs).
namespace {
struct MapKey {
  uint64_t ptr, d[4], s[3];
  uint32_t box[4], rank, swz;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapSlot {
  MapKey key;
  CUtensorMap map;
  bool used;
};
static_assert(sizeof(MapKey) == 88, "MapKey is compared and hashed bytewise: no padding allowed");
static_assert(sizeof(MapSlot) % 64 == 0, "slots keep the 64-byte alignment of CUtensorMap");
constexpr uint32_t kMapSlots = 1u << 14;
std::mutex g_map_mutex;
MapSlot* g_map_slots = nullptr;
uint32_t g_map_count = 0;
unsigned long long g_map_hits = 0, g_map_misses = 0;

uint64_t hash_key(const MapKey& k) {
  const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
  uint64_t h = 0x9e3779b97f4a7c15ull;
  for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) {
    h ^= w[i] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
  }
  return h;
}
bool cache_enabled() {
  static const bool on = !(getenv("MVB_TMAP_CACHE") && atoi(getenv("MVB_TMAP_CACHE")) == 0);
  return on;
}

bool encode_raw(CUtensorMap* m, const MapKey& k) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gd[4] = {k.d[0], k.d[1], k.d[2], k.d[3]};
  cuuint64_t gs[3] = {k.s[0] * 2, k.s[1] * 2, k.s[2] * 2};
  cuuint32_t bx[4] = {k.box[0], k.box[1], k.box[2], k.box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, k.rank, reinterpret_cast<void*>(k.ptr), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)k.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool encode_cached(CUtensorMap* m, const MapKey& k) {
  if (!cache_enabled()) return encode_raw(m, k);
  std::lock_guard<std::mutex> lock(g_map_mutex);
  if (!g_map_slots) {                                // CUtensorMap is alignas(64): aligned_alloc, not calloc
    g_map_slots = static_cast<MapSlot*>(aligned_alloc(64, sizeof(MapSlot) * kMapSlots));
    if (g_map_slots) memset(g_map_slots, 0, sizeof(MapSlot) * kMapSlots);
  }
  if (!g_map_slots) return encode_raw(m, k);
  uint32_t i = (uint32_t)hash_key(k) & (kMapSlots - 1);
  while (g_map_slots[i].used) {
    if (g_map_slots[i].key == k) {
      *m = g_map_slots[i].map;
      ++g_map_hits;
      return true;
    }
    i = (i + 1) & (kMapSlots - 1);
  }
  if (!encode_raw(m, k)) return false;
  ++g_map_misses;
  if (g_map_count >= kMapSlots / 4 * 3) {            // full: start over (the live working set re-enters within one forward)
    memset(g_map_slots, 0, sizeof(MapSlot) * kMapSlots);
    g_map_count = 0;
    i = (uint32_t)hash_key(k) & (kMapSlots - 1);
  }
  g_map_slots[i].key = k;
  g_map_slots[i].map = *m;
  g_map_slots[i].used = true;
  ++g_map_count;
  return true;
}
}  // namespace

void tensor_map_cache_stats(unsigned long long* hits, unsigned long long* misses) {
  std::lock_guard<std::mutex> lock(g_map_mutex);
  *hits = g_map_hits;
  *misses = g_map_misses;
}

// rank-4 fp16 map, inner box 64 elements, swizzle as given, zero fill out of bounds
bool encode_map_4d_sw(CUtensorMap* m, const void* ptr, const uint64_t dims[4], const uint64_t strides_elems[3],
                      const uint32_t box[4], CUtensorMapSwizzle swz) {
  MapKey k;
  memset(&k, 0, sizeof(k));                          // the key is compared bytewise: no uninitialised padding
  k.ptr = reinterpret_cast<uint64_t>(ptr);
  for (int i = 0; i < 4; ++i) { k.d[i] = dims[i]; k.box[i] = box[i]; }
  for (int i = 0; i < 3; ++i) k.s[i] = strides_elems[i];
  k.rank = 4;
  k.swz = (uint32_t)swz;
  return encode_cached(m, k);
}
bool encode_map_4d(CUtensorMap* m, const void* ptr, const uint64_t dims[4], const uint64_t strides_elems[3],
                   const uint32_t box[4]) {
  return encode_map_4d_sw(m, ptr, dims, strides_elems, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

bool encode_map_2d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t stride1_elems, uint32_t b0,
                   uint32_t b1) {
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.ptr = reinterpret_cast<uint64_t>(ptr);
  k.d[0] = d0; k.d[1] = d1;
  k.s[0] = stride1_elems;
  k.box[0] = b0; k.box[1] = b1;
  k.rank = 2;
  k.swz = (uint32_t)CU_TENSOR_MAP_SWIZZLE_128B;
  return encode_cached(m, k);
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Pick the 128-row pixel box {bw, bh, bn} (powers of two) that wastes the fewest padded rows.
static void pick_box(int W, int H, int NF, int* bw, int* bh, int* bn) {
  long long best = -1;
  for (int w = 128; w >= 1; w >>= 1) {
    for (int h = 128 / w; h >= 1; h >>= 1) {
      const int n = 128 / (w * h);
      const long long padded = (long long)ceil_div(W, w) * ceil_div(H, h) * ceil_div(NF, n);
      if (best < 0 || padded < best) {
        best = padded;
        *bw = w; *bh = h; *bn = n;
      }
    }
  }
}

static int pick_block_n(int N, int geglu, long long tiles_m, int num_sms) {
  // candidates are multiples of 32 (GEGLU chunks) or 16; prefer few padded columns, then fewer waves
  int best = 0;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 32; bn -= 32) {
    if (geglu && (bn % 64)) continue;
    const int tn = ceil_div(N, bn);
    const long long tiles = tiles_m * tn;
    const long long waves = (tiles + num_sms - 1) / num_sms;
    // time ~ waves * (bn + epilogue/fixed overhead per tile)
    const double cost = (double)waves * (bn + 24.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

static cudaError_t launch_common(cudaStream_t stream, const CUtensorMap* maps, int nmaps, ConvGemmParams& p,
                                 const __half* wt, long long ktot, const Epilogue& ep, int num_sms, const char** err) {
  // kernel variants indexed [cta pair][epilogue]
  typedef void (*KernelFn)(const AMaps, const CUtensorMap, const CUtensorMap, const ConvGemmParams);
  static const KernelFn kernels[2][4] = {
      {conv_gemm_kernel<false, kEpiGeneric>, conv_gemm_kernel<false, kEpiPlain>, conv_gemm_kernel<false, kEpiResidual>,
       conv_gemm_kernel<false, kEpiGeglu>},
      {conv_gemm_kernel<true, kEpiGeneric>, conv_gemm_kernel<true, kEpiPlain>, conv_gemm_kernel<true, kEpiResidual>,
       conv_gemm_kernel<true, kEpiGeglu>}};
  // the opt-in is per device: key the "already set" state by the current device ordinal
  static bool attr_set_dev[64] = {};
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& attr_set = attr_set_dev[cur_dev & 63];
  if (!attr_set) {
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 4; ++b) {
        cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(kernels[a][b]),
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) { *err = "cudaFuncSetAttribute(conv_gemm_kernel)"; return e; }
      }
    attr_set = true;
  }
  const long long tiles_m = (long long)p.tiles_w * p.tiles_h * p.tiles_n;
  p.block_n = pick_block_n(p.N, ep.geglu, tiles_m, num_sms);
  p.tiles_nn = ceil_div(p.N, p.block_n);
  p.stage_bytes = kABytes + ((p.block_n * kBlockK * 2 + 1023) / 1024) * 1024;   // B tile rounded up to the swizzle period
  p.nstages = kRingBytes / p.stage_bytes;
  if (p.nstages > kMaxStages) p.nstages = kMaxStages;
  static const int stage_cap = getenv("MVB_STAGES") ? atoi(getenv("MVB_STAGES")) : 0;   // experiment knob
  if (stage_cap > 1 && p.nstages > stage_cap) p.nstages = stage_cap;
  p.out = ep.out; p.ldc = ep.ldc; p.bias = ep.bias; p.rowadd = ep.rowadd;
  p.rows_per_group = ep.rows_per_group > 0 ? ep.rows_per_group : 1;
  p.ld_rowadd = ep.ld_rowadd; p.res = ep.res; p.ld_res = ep.ld_res; p.alpha = ep.alpha; p.beta = ep.beta;
  p.geglu = ep.geglu; p.act = ep.act; p.out_f32 = ep.out_f32;
  if (ep.out_f32 && (ep.geglu || ep.res)) { *err = "conv_gemm: fp32 output excludes geglu/residual"; return cudaErrorInvalidValue; }
  // CTA pairs when there is enough work for all 74 pairs (env MVB_TWOCTA=0/1 forces the choice for experiments)
  static const int twocta_env = getenv("MVB_TWOCTA") ? atoi(getenv("MVB_TWOCTA")) : -1;
  const long long pair_units = ((tiles_m + 1) / 2) * p.tiles_nn;
  // pairs pay off when the mainloop dominates (K >= 1280: +5..9 % measured); short-K GEMMs are epilogue-bound and
  // lose ~15 % to the pair-wide accumulator hand-off
  bool two_cta = tiles_m >= 2 && pair_units >= (num_sms / 2) && (num_sms % 2 == 0) && ktot >= 1280;
  if (twocta_env == 0) two_cta = false;
  if (twocta_env == 1 && tiles_m >= 2 && (num_sms % 2 == 0)) two_cta = true;
  CUtensorMap tmB;
  if (!encode_map_2d(&tmB, wt, (uint64_t)ktot, (uint64_t)p.N, (uint64_t)ktot, 64u,
                     (uint32_t)(two_cta ? p.block_n / 2 : p.block_n))) {
    *err = "cuTensorMapEncodeTiled(B) failed";
    return cudaErrorInvalidValue;
  }
  // output tensor map {cols, W, H, NF}: mirrors the A box, 32 columns x 128 pixels, 64-byte swizzle
  CUtensorMap tmC = tmB;
  p.tma_store = ep.out_f32 ? 0 : 1;
  if (p.tma_store) {
    const int ncols = ep.geglu ? p.N / 2 : p.N;
    const uint64_t dims[4] = {(uint64_t)ncols, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.NF};
    const uint64_t st[3] = {(uint64_t)ep.ldc, (uint64_t)ep.ldc * p.W, (uint64_t)ep.ldc * p.W * p.H};
    const uint32_t box[4] = {32u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
    if ((ep.ldc % 8) || !encode_map_4d_sw(&tmC, ep.out, dims, st, box, CU_TENSOR_MAP_SWIZZLE_64B)) {
      *err = "cuTensorMapEncodeTiled(C) failed (output row stride must be a multiple of 8 elements)";
      return cudaErrorInvalidValue;
    }
  }
  const long long num_tiles = tiles_m * p.tiles_nn;
  int grid = (int)(num_tiles < num_sms ? num_tiles : num_sms);
  if (two_cta) grid = (int)(2 * (pair_units < num_sms / 2 ? pair_units : num_sms / 2));
  // epilogue variant: the fast ones need whole 32-column chunks and the common alpha / beta
  static const int epi_env = getenv("MVB_EPI") ? atoi(getenv("MVB_EPI")) : -1;   // 0 forces the generic epilogue
  int epi = kEpiGeneric;
  if (!ep.out_f32 && ep.act == 0 && epi_env != 0) {
    if (ep.geglu) { if (p.N % 64 == 0 && !ep.rowadd) epi = kEpiGeglu; }
    else if (p.N % 32 == 0) {
      if (ep.res && ep.beta == 1.f) epi = kEpiResidual;
      else if (!ep.res && ep.alpha == 1.f) epi = kEpiPlain;
    }
  }
  static const bool trace = getenv("MVB_TRACE") != nullptr;
  if (trace)
    fprintf(stderr, "MVB_TRACE gemm M=%lld N=%d K=%lld taps=%d block_n=%d tiles=%lld geglu=%d res=%d f32=%d cta2=%d epi=%d\n",
            (long long)p.W * p.H * p.NF, p.N, ktot, p.ntaps, p.block_n, num_tiles, p.geglu, p.res != nullptr, p.out_f32,
            (int)two_cta, epi);
  AMaps am;
  for (int i = 0; i < 4; ++i) am.m[i] = maps[i < nmaps ? i : 0];
  ProfScope prof(stream, KC_GEMM);
  cudaError_t e;
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = two_cta ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, kernels[two_cta ? 1 : 0][epi], am, tmB, tmC, p);
  }
  if (e != cudaSuccess) *err = "conv_gemm_kernel launch";
  return e;
}

cudaError_t launch_conv_gemm(cudaStream_t stream, const ASource& a0, const ASource* a1, int W, int H, int NF,
                             int ntaps, const int8_t* dy, const int8_t* dx, const __half* wt, int N,
                             const Epilogue& ep, int num_sms, const char** err) {
  if ((a0.C % 64) || (a1 && (a1->C % 64)) || (N % 8) || ntaps < 1 || ntaps > 9) {
    *err = "conv_gemm: channels must be multiples of 64, N a multiple of 8, 1..9 taps";
    return cudaErrorInvalidValue;
  }
  ConvGemmParams p{};
  p.W = W; p.H = H; p.NF = NF;
  pick_box(W, H, NF, &p.bw, &p.bh, &p.bn);
  p.tiles_w = ceil_div(W, p.bw); p.tiles_h = ceil_div(H, p.bh); p.tiles_n = ceil_div(NF, p.bn);
  p.ntaps = ntaps;
  for (int i = 0; i < ntaps; ++i) { p.dy[i] = dy[i]; p.dx[i] = dx[i]; p.tap_src[i] = 0; }
  p.kb0 = a0.C / 64;
  p.kb1 = a1 ? a1->C / 64 : 0;
  p.N = N;
  CUtensorMap maps[2];
  const uint32_t box[4] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
  {
    const uint64_t dims[4] = {(uint64_t)a0.C, (uint64_t)W, (uint64_t)H, (uint64_t)NF};
    const uint64_t st[3] = {(uint64_t)a0.sW, (uint64_t)a0.sH, (uint64_t)a0.sN};
    if (!encode_map_4d(&maps[0], a0.ptr, dims, st, box)) { *err = "cuTensorMapEncodeTiled(A0) failed"; return cudaErrorInvalidValue; }
  }
  if (a1) {
    const uint64_t dims[4] = {(uint64_t)a1->C, (uint64_t)W, (uint64_t)H, (uint64_t)NF};
    const uint64_t st[3] = {(uint64_t)a1->sW, (uint64_t)a1->sH, (uint64_t)a1->sN};
    if (!encode_map_4d(&maps[1], a1->ptr, dims, st, box)) { *err = "cuTensorMapEncodeTiled(A1) failed"; return cudaErrorInvalidValue; }
  }
  const long long ktot = (long long)ntaps * (a0.C + (a1 ? a1->C : 0));
  return launch_common(stream, maps, a1 ? 2 : 1, p, wt, ktot, ep, num_sms, err);
}

cudaError_t launch_conv_s2(cudaStream_t stream, const __half* x, int C, int W, int H, int NF, const __half* wt, int N,
                           const Epilogue& ep, int num_sms, const char** err) {
  if ((C % 64) || (N % 8) || (W % 2) || (H % 2)) {
    *err = "conv_s2: channels multiple of 64, even H and W";
    return cudaErrorInvalidValue;
  }
  const int Wo = W / 2, Ho = H / 2;
  ConvGemmParams p{};
  p.W = Wo; p.H = Ho; p.NF = NF;
  pick_box(Wo, Ho, NF, &p.bw, &p.bh, &p.bn);
  p.tiles_w = ceil_div(Wo, p.bw); p.tiles_h = ceil_div(Ho, p.bh); p.tiles_n = ceil_div(NF, p.bn);
  p.ntaps = 9;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      const int i = ky * 3 + kx;
      // input row 2y + ky - 1: ky=0 -> odd phase at y-1; ky=1 -> even phase at y; ky=2 -> odd phase at y
      const int py = (ky == 1) ? 0 : 1, px = (kx == 1) ? 0 : 1;
      p.dy[i] = (ky == 0) ? -1 : 0;
      p.dx[i] = (kx == 0) ? -1 : 0;
      p.tap_src[i] = (int8_t)(py * 2 + px);
    }
  p.kb0 = C / 64; p.kb1 = 0; p.N = N;
  CUtensorMap maps[4];
  const uint32_t box[4] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bn};
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      const uint64_t dims[4] = {(uint64_t)C, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)NF};
      const uint64_t st[3] = {(uint64_t)2 * C, (uint64_t)2 * W * C, (uint64_t)H * W * C};
      if (!encode_map_4d(&maps[py * 2 + px], x + ((long long)py * W + px) * C, dims, st, box)) {
        *err = "cuTensorMapEncodeTiled(phase) failed";
        return cudaErrorInvalidValue;
      }
    }
  return launch_common(stream, maps, 4, p, wt, 9LL * C, ep, num_sms, err);
}

}  // namespace mvb
