// C-ABI entry points for the individual device ops (used by the per-op parity tests and by embedders that
// only want one kernel). The whole-forward entry points live in capi_engine.cu.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "../../include/musev_b200.h"
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
  ep.ld_res = d->ld_res; ep.alpha = d->alpha; ep.beta = d->beta; ep.geglu = d->geglu; ep.act = d->act;
  const char* err = nullptr;
  cudaError_t e = launch_conv_gemm((cudaStream_t)stream, a0, d->a1 ? &a1 : nullptr, d->W, d->H, d->NF, d->ntaps,
                                   d->dy, d->dx, (const __half*)d->weight, d->N, ep, sm_count(), &err);
  if (e != cudaSuccess) return fail(err, e);
  return MVB_OK;
}

}  // extern "C"
