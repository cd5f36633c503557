// C-ABI entry points for the individual device ops (used by the per-op parity tests and by embedders that
// only want one kernel). The whole-forward entry points live in capi_engine.cu.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/musev_b200.h"
#include "attention.cuh"
#include "conv_gemm.cuh"
#include "ops.cuh"

using namespace mvb;

static thread_local char g_err[512] = "";

static int fail(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what ? what : "error", e == cudaSuccess ? "invalid argument" : cudaGetErrorString(e));
  return e == cudaSuccess ? MVB_ERR_INVALID : MVB_ERR_CUDA;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

extern "C" {

const char* mvb_last_error(void) { return g_err; }

int mvb_version(void) { return 1; }

int mvb_op_conv_gemm(const mvb_conv_gemm_desc* d, void* stream) {
  if (!d || !d->a0 || !d->weight || !d->out) return fail("mvb_op_conv_gemm: null pointer", cudaSuccess);
  ASource a0{(const __half*)d->a0, d->c0, d->a0_stride_w, d->a0_stride_h, d->a0_stride_n};
  ASource a1{(const __half*)d->a1, d->c1, d->a1_stride_w, d->a1_stride_h, d->a1_stride_n};
  Epilogue ep;
  ep.out = (__half*)d->out; ep.ldc = d->ldc; ep.bias = d->bias; ep.rowadd = d->rowadd;
  ep.rows_per_group = d->rows_per_group; ep.ld_rowadd = d->ld_rowadd; ep.res = (const __half*)d->residual;
  ep.ld_res = d->ld_res; ep.alpha = d->alpha; ep.beta = d->beta; ep.geglu = d->geglu; ep.act = d->act; ep.out_f32 = d->out_f32;
  const char* err = nullptr;
  if (d->stride2) {
    cudaError_t e2 = launch_conv_s2((cudaStream_t)stream, (const __half*)d->a0, d->c0, d->W, d->H, d->NF,
                                    (const __half*)d->weight, d->N, ep, sm_count(), &err);
    if (e2 != cudaSuccess) return fail(err, e2);
    return MVB_OK;
  }
  cudaError_t e = launch_conv_gemm((cudaStream_t)stream, a0, d->a1 ? &a1 : nullptr, d->W, d->H, d->NF, d->ntaps,
                                   d->dy, d->dx, (const __half*)d->weight, d->N, ep, sm_count(), &err);
  if (e != cudaSuccess) return fail(err, e);
  return MVB_OK;
}


int mvb_op_attention(const mvb_attention_desc* d, void* stream) {
  if (!d || !d->q || !d->out || !d->k[0] || !d->v[0]) return fail("mvb_op_attention: null pointer", cudaSuccess);
  AttnArgs a{};
  a.q = (const __half*)d->q; a.ldq = d->ldq; a.NF = d->NF; a.Nq = d->Nq; a.heads = d->heads; a.d = d->d; a.dp = d->dp;
  a.scale = d->scale; a.nseg = d->nseg;
  for (int s = 0; s < 2 && s < d->nseg; ++s) {
    a.seg[s].k = (const __half*)d->k[s]; a.seg[s].v = (const __half*)d->v[s]; a.seg[s].ld = d->ldkv[s];
    a.seg[s].rows = d->kv_rows[s]; a.seg[s].nk = d->nk[s]; a.seg[s].fdiv = d->fdiv[s]; a.seg[s].fmul = d->fmul[s];
    a.seg[s].fadd = d->fadd[s];
  }
  a.out = (__half*)d->out; a.ldo = d->ldo; a.out_scale = d->out_scale; a.accumulate = d->accumulate; a.v_ones_col = d->v_ones_col; a.variant = d->variant;
  const char* err = nullptr;
  cudaError_t e = launch_attention((cudaStream_t)stream, a, &err);
  if (e != cudaSuccess) return fail(err, e);
  return MVB_OK;
}

int mvb_tensor_map_cache_stats(unsigned long long* hits, unsigned long long* misses) {
  if (!hits || !misses) return fail("mvb_tensor_map_cache_stats: null pointer", cudaSuccess);
  tensor_map_cache_stats(hits, misses);
  return MVB_OK;
}

int mvb_debug_attention_trace(long long* device_buffer) {
  set_attention_trace(device_buffer);
  return MVB_OK;
}

int mvb_op_temporal_attention(const void* qkv, int ld, int B, int T, int HW, int heads, int d, int dp, float scale,
                              void* out, int ldo, void* stream) {
  cudaError_t e = temporal_attention((cudaStream_t)stream, (const __half*)qkv, ld, B, T, HW, heads, d, dp, scale,
                                     (__half*)out, ldo);
  if (e != cudaSuccess) return fail("mvb_op_temporal_attention", e == cudaErrorInvalidValue ? cudaSuccess : e);
  return MVB_OK;
}

int mvb_op_groupnorm(const void* x0, int c0, const void* x1, int c1, int NF, int HW, int groups, int frames_per_stat,
                     float eps, const float* gamma, const float* beta, int silu, void* y, float* scratch, void* stream) {
  int chunks = 0;
  cudaError_t e = gn_stats((cudaStream_t)stream, (const __half*)x0, c0, (const __half*)x1, c1, NF, HW, groups, scratch,
                           &chunks);
  if (e == cudaSuccess)
    e = gn_apply((cudaStream_t)stream, (const __half*)x0, c0, (const __half*)x1, c1, NF, HW, groups, scratch, chunks,
                 frames_per_stat, eps, gamma, beta, silu, (__half*)y);
  if (e != cudaSuccess) return fail("mvb_op_groupnorm", e == cudaErrorInvalidValue ? cudaSuccess : e);
  return MVB_OK;
}

int mvb_op_groupnorm_fused(const void* x0, int c0, const void* x1, int c1, int NF, int HW, int groups, int frames_per_stat,
                           float eps, const float* gamma, const float* beta, int silu, void* y, float* scratch,
                           unsigned int* barrier_word, unsigned int* arrivals, void* stream) {
  if (!barrier_word || !arrivals) return fail("mvb_op_groupnorm_fused: null barrier word", cudaSuccess);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = gn_fused((cudaStream_t)stream, (const __half*)x0, c0, (const __half*)x1, c1, NF, HW, groups, scratch,
                           frames_per_stat, eps, gamma, beta, silu, (__half*)y, sms, barrier_word, arrivals);
  if (e != cudaSuccess) return fail("mvb_op_groupnorm_fused", e == cudaErrorInvalidValue ? cudaSuccess : e);
  return MVB_OK;
}

int mvb_op_layernorm(const void* x, long long M, int C, float eps, const float* gamma, const float* beta, void* y,
                     void* stream) {
  cudaError_t e = layernorm((cudaStream_t)stream, (const __half*)x, M, C, eps, gamma, beta, (__half*)y);
  if (e != cudaSuccess) return fail("mvb_op_layernorm", e == cudaErrorInvalidValue ? cudaSuccess : e);
  return MVB_OK;
}

int mvb_fuse_cfg_ddim(const float* eps_sum, const float* counter, const void* latents_in, void* latents_out,
                      int is_f32, int B, int C, int T, int HW, int cfg, float guidance_scale, float alpha_prod_t,
                      float alpha_prod_t_prev, int prediction_type, float clip_range, int use_clipped_model_output,
                      float std_dev_t, const float* variance_noise, float* eps_out, float* x0_out, void* stream) {
  if (!eps_sum || !latents_in || !latents_out) return fail("mvb_fuse_cfg_ddim: null pointer", cudaSuccess);
  if (alpha_prod_t <= 0.f || alpha_prod_t > 1.f || prediction_type < 0 || prediction_type > 2)
    return fail("mvb_fuse_cfg_ddim: bad alpha / prediction_type", cudaSuccess);
  cudaError_t e = fuse_cfg_ddim((cudaStream_t)stream, eps_sum, counter, latents_in, latents_out, is_f32, B, C, T, HW, cfg,
                                guidance_scale, alpha_prod_t, alpha_prod_t_prev, prediction_type, clip_range,
                                use_clipped_model_output, std_dev_t, variance_noise, eps_out, x0_out);
  if (e != cudaSuccess) return fail("mvb_fuse_cfg_ddim", e);
  return MVB_OK;
}

int mvb_fuse_cfg_affine(const float* eps_sum, const float* counter, const void* latents_in, void* latents_out, int is_f32,
                        int B, int C, int T, int HW, int cfg, float guidance_scale, float c_x, float c_e, float c_n,
                        const float* noise, float a_x, float a_e, float* aux_out, float* eps_out, void* stream) {
  if (!eps_sum || !latents_in || !latents_out) return fail("mvb_fuse_cfg_affine: null pointer", cudaSuccess);
  cudaError_t e = fuse_cfg_affine((cudaStream_t)stream, eps_sum, counter, latents_in, latents_out, is_f32, B, C, T, HW, cfg,
                                  guidance_scale, c_x, c_e, c_n, noise, a_x, a_e, aux_out, eps_out);
  if (e != cudaSuccess) return fail("mvb_fuse_cfg_affine", e);
  return MVB_OK;
}

int mvb_accumulate_window(float* eps_sum, int B2, int C, int T, int HW, const void* eps_window, int is_f32, int Tw,
                          int src_t0, const int* frames_dev, int nframes, void* stream) {
  if (!eps_sum || !eps_window || !frames_dev) return fail("mvb_accumulate_window: null pointer", cudaSuccess);
  cudaError_t e = accumulate_window((cudaStream_t)stream, eps_sum, B2, C, T, HW, eps_window, is_f32, Tw, src_t0,
                                    frames_dev, nframes);
  if (e != cudaSuccess) return fail("mvb_accumulate_window", e);
  return MVB_OK;
}

}  // extern "C"
