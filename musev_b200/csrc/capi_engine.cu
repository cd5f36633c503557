// C-ABI of the whole-model engine (declarations and reference call sites: include/musev_b200.h).
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "../../include/musev_b200.h"
#include "engine.cuh"

struct mvb_handle {
  mvb::Engine* e;
};

extern "C" {

int mvb_create(const mvb_config* cfg, int device, mvb_handle** out) {
  if (!cfg || !out) return MVB_ERR_INVALID;
  if (cfg->num_blocks < 1 || cfg->num_blocks > 4 || cfg->heads < 1 || cfg->norm_num_groups < 1) return MVB_ERR_INVALID;
  for (int i = 0; i < cfg->num_blocks; ++i) {
    const int c = cfg->block_out_channels[i];
    if (c % 64 || c % cfg->heads || (c / cfg->heads) % 8 || c % cfg->norm_num_groups) return MVB_ERR_INVALID;
  }
  if (cfg->cross_attention_dim % 64 || cfg->in_channels * 9 > 64 || cfg->out_channels > 16) return MVB_ERR_INVALID;
  mvb::Engine* e = new (std::nothrow) mvb::Engine(*cfg, device);
  if (!e) return MVB_ERR_STATE;
  if (e->error()[0]) { delete e; return MVB_ERR_CUDA; }
  mvb_handle* h = new (std::nothrow) mvb_handle{e};
  if (!h) { delete e; return MVB_ERR_STATE; }
  *out = h;
  return MVB_OK;
}

static int check_config(const mvb_config* cfg) {
  if (cfg->num_blocks < 1 || cfg->num_blocks > 4 || cfg->heads < 1 || cfg->norm_num_groups < 1) return MVB_ERR_INVALID;
  for (int i = 0; i < cfg->num_blocks; ++i) {
    const int c = cfg->block_out_channels[i];
    if (c % 64 || c % cfg->heads || (c / cfg->heads) % 8 || c % cfg->norm_num_groups) return MVB_ERR_INVALID;
  }
  if (cfg->cross_attention_dim % 64 || cfg->in_channels * 9 > 64) return MVB_ERR_INVALID;
  return MVB_OK;
}

int mvb_create_controlnet(const mvb_config* cfg, int device, mvb_handle** out) {
  if (!cfg || !out) return MVB_ERR_INVALID;
  if (check_config(cfg) != MVB_OK) return MVB_ERR_INVALID;
  int n_out = 2;
  for (int i = 0; i < cfg->num_blocks; ++i) n_out += cfg->layers_per_block + (i == cfg->num_blocks - 1 ? 0 : 1);
  if (n_out > MVB_CONTROLNET_MAX_OUT) return MVB_ERR_INVALID;
  mvb::Engine* e = new (std::nothrow) mvb::Engine(*cfg, device, 1);
  if (!e) return MVB_ERR_STATE;
  if (e->error()[0]) { delete e; return MVB_ERR_CUDA; }
  mvb_handle* h = new (std::nothrow) mvb_handle{e};
  if (!h) { delete e; return MVB_ERR_STATE; }
  *out = h;
  return MVB_OK;
}

int mvb_create_referencenet(const mvb_config* cfg, int device, mvb_handle** out) {
  if (!cfg || !out) return MVB_ERR_INVALID;
  if (check_config(cfg) != MVB_OK) return MVB_ERR_INVALID;
  int n_out = 2;
  for (int i = 0; i < cfg->num_blocks; ++i) n_out += cfg->layers_per_block + (i == cfg->num_blocks - 1 ? 0 : 1);
  if (n_out > MVB_CONTROLNET_MAX_OUT) return MVB_ERR_INVALID;
  mvb::Engine* e = new (std::nothrow) mvb::Engine(*cfg, device, 2);
  if (!e) return MVB_ERR_STATE;
  if (e->error()[0]) { delete e; return MVB_ERR_CUDA; }
  mvb_handle* h = new (std::nothrow) mvb_handle{e};
  if (!h) { delete e; return MVB_ERR_STATE; }
  *out = h;
  return MVB_OK;
}

long long mvb_referencenet_workspace_bytes(mvb_handle* h, const mvb_controlnet_args* args) {
  if (!h || !args || h->e->kind() != 2) return -1;
  return h->e->controlnet_workspace_bytes(*args);
}

int mvb_referencenet_forward(mvb_handle* h, const mvb_controlnet_args* args, void* workspace, long long workspace_bytes,
                             void* stream) {
  if (!h || !args) return MVB_ERR_INVALID;
  if (h->e->kind() != 2) return MVB_ERR_STATE;
  return h->e->controlnet_forward(*args, workspace, workspace_bytes, (cudaStream_t)stream);
}

long long mvb_controlnet_workspace_bytes(mvb_handle* h, const mvb_controlnet_args* args) {
  if (!h || !args) return -1;
  return h->e->controlnet_workspace_bytes(*args);
}

int mvb_controlnet_forward(mvb_handle* h, const mvb_controlnet_args* args, void* workspace, long long workspace_bytes,
                           void* stream) {
  if (!h || !args) return MVB_ERR_INVALID;
  return h->e->controlnet_forward(*args, workspace, workspace_bytes, (cudaStream_t)stream);
}

int mvb_create_vae_decoder(const mvb_config* cfg, int device, mvb_handle** out) {
  if (!cfg || !out) return MVB_ERR_INVALID;
  if (cfg->num_blocks < 1 || cfg->num_blocks > 4 || cfg->norm_num_groups < 1 || cfg->layers_per_block < 1) return MVB_ERR_INVALID;
  for (int i = 0; i < cfg->num_blocks; ++i) {
    const int c = cfg->block_out_channels[i];
    if (c % 64 || c % cfg->norm_num_groups || (c / cfg->norm_num_groups) % 2) return MVB_ERR_INVALID;
  }
  if (cfg->in_channels < 1 || cfg->in_channels > 7 || cfg->out_channels < 1 || cfg->out_channels > 16) return MVB_ERR_INVALID;
  mvb::Engine* e = new (std::nothrow) mvb::Engine(*cfg, device, 3);
  if (!e) return MVB_ERR_STATE;
  if (e->error()[0]) { delete e; return MVB_ERR_CUDA; }
  mvb_handle* h = new (std::nothrow) mvb_handle{e};
  if (!h) { delete e; return MVB_ERR_STATE; }
  *out = h;
  return MVB_OK;
}

long long mvb_vae_decode_workspace_bytes(mvb_handle* h, const mvb_vae_decode_args* args) {
  if (!h || !args) return -1;
  return h->e->vae_workspace_bytes(*args);
}

int mvb_vae_decode(mvb_handle* h, const mvb_vae_decode_args* args, void* workspace, long long workspace_bytes, void* stream) {
  if (!h || !args) return MVB_ERR_INVALID;
  return h->e->vae_decode(*args, workspace, workspace_bytes, (cudaStream_t)stream);
}

void mvb_destroy(mvb_handle* h) {
  if (!h) return;
  delete h->e;
  delete h;
}

int mvb_load_weight(mvb_handle* h, const char* name, const void* device_ptr, int is_f32, const long long* shape, int ndim) {
  if (!h || !name || !device_ptr || !shape) return MVB_ERR_INVALID;
  return h->e->load_weight(name, device_ptr, is_f32, shape, ndim);
}

int mvb_load_weights(mvb_handle* h, const mvb_named_tensor* tensors, int n) {
  if (!h || (!tensors && n > 0) || n < 0) return MVB_ERR_INVALID;
  return h->e->load_weights(tensors, n);
}

int mvb_finalize(mvb_handle* h) { return h ? h->e->finalize() : MVB_ERR_INVALID; }
int mvb_num_params(mvb_handle* h) { return h ? h->e->num_params() : 0; }

long long mvb_workspace_bytes(mvb_handle* h, const mvb_unet_args* args) {
  if (!h || !args) return -1;
  return h->e->workspace_bytes(*args);
}

int mvb_unet_forward(mvb_handle* h, const mvb_unet_args* args, void* workspace, long long workspace_bytes, void* stream) {
  if (!h || !args) return MVB_ERR_INVALID;
  return h->e->forward(*args, workspace, workspace_bytes, (cudaStream_t)stream);
}

const char* mvb_handle_error(mvb_handle* h) { return h ? h->e->error() : "null handle"; }

/* Debug: layer outputs of the last forward (pointers into the caller's workspace; fp16 [rows, C] channels-last). */
int mvb_debug_num_taps(mvb_handle* h) { return h ? (int)h->e->taps().size() : 0; }
int mvb_debug_tap(mvb_handle* h, int i, char* name, int name_cap, const void** ptr, long long* rows, int* C) {
  if (!h || i < 0 || i >= (int)h->e->taps().size()) return MVB_ERR_INVALID;
  const auto& t = h->e->taps()[i];
  snprintf(name, name_cap, "%s", t.name.c_str());
  *ptr = t.p; *rows = t.rows; *C = t.C;
  return MVB_OK;
}

}  // extern "C"
