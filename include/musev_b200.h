/* musev_b200 C ABI -- B200 (sm_100a) kernels for MuseV's denoising hot path.
 *
 * Conventions (mirrors how the reference drives its model):
 *   - every data pointer is a DEVICE pointer into caller-owned memory (e.g. torch `tensor.data_ptr()`);
 *   - calls are asynchronous on the `stream` argument (a cudaStream_t passed as void*; NULL = default stream);
 *   - return value: MVB_OK (0) or a negative error code; the message is available from mvb_last_error()
 *     (thread local). Nothing throws or aborts;
 *   - activations are channels-last fp16: a video batch is [B, T, H, W, C] which is at the same time the token
 *     matrix [(b t h w), C] of every 1x1 conv / nn.Linear of the reference.
 */
#ifndef MUSEV_B200_H_
#define MUSEV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVB_OK 0
#define MVB_ERR_INVALID (-1)
#define MVB_ERR_CUDA (-2)
#define MVB_ERR_STATE (-3)

const char* mvb_last_error(void);
int mvb_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * Op level: implicit-GEMM convolution / linear on tcgen05 tensor cores.
 *
 * Replaces, for channels-last fp16 activations:
 *   F.conv2d 3x3 pad 1      diffusers/src/diffusers/models/resnet.py:643,666 (ResnetBlock2D.conv1/conv2),
 *                           :159,201 (Upsample2D.conv), :247,272 (Downsample2D.conv, via mvb_op_space_to_depth)
 *   F.conv3d (3,1,1) pad 1  musev/models/resnet.py:56-78 (TemporalConvLayer.conv1..4)
 *   F.conv2d 1x1, F.linear  diffusers models/transformer_2d.py:150,212; attention_processor.py:181-196;
 *                           attention.py:342-395 (GEGLU feed-forward); musev/models/temporal_transformer.py:121-167
 *
 * The input is viewed as an image {C, W, H, NF} with arbitrary element strides; `ntaps` offsets (dy[i], dx[i])
 * are accumulated, out-of-image taps read zeros. K index of the packed weight [N, ntaps*(c0+c1)] is
 * tap-major, then source-0 channels, then source-1 channels (torch.cat order of a skip connection).
 *   out[m, n] = act((acc + bias[n] + rowadd[m / rows_per_group, n]) * alpha + beta * residual[m, n])
 * with m = (frame*H + h)*W + w. geglu=1: packed columns come in [16 value | 16 gate] chunks and
 * out[m, j] = value * gelu_erf(gate) has N/2 columns.
 */
typedef struct mvb_conv_gemm_desc {
  const void* a0; int c0; long long a0_stride_w, a0_stride_h, a0_stride_n;
  const void* a1; int c1; long long a1_stride_w, a1_stride_h, a1_stride_n; /* a1 may be NULL */
  int W, H, NF;
  int ntaps; int8_t dy[9]; int8_t dx[9];
  const void* weight; int N;
  void* out; long long ldc;
  const float* bias;
  const float* rowadd; int rows_per_group; int ld_rowadd;
  const void* residual; long long ld_res;
  float alpha, beta;
  int geglu;
  int act; /* 0 none, 1 SiLU */
  int out_f32; /* store fp32 (no residual / geglu) */
  int stride2; /* 1: 3x3 stride-2 pad-1 conv of a contiguous [NF,H,W,c0] input (taps ignored) */
} mvb_conv_gemm_desc;

int mvb_op_conv_gemm(const mvb_conv_gemm_desc* desc, void* stream);


/* ---------------------------------------------------------------------------------------------------------
 * Op level: flash attention on tcgen05 (spatial self / reference / cross attention of the transformer blocks).
 *
 * Replaces xformers.ops.memory_efficient_attention at musev/models/attention_processor.py:258,292,519,724 and
 * F.scaled_dot_product_attention at diffusers/src/diffusers/models/attention_processor.py:1166-1250.
 * Q/K/V are head-padded token matrices (head h occupies columns [h*dp, h*dp+d), dp a multiple of 16, padding = 0);
 * the keys of query frame f are the concatenation of up to two segments, segment s starting at row
 * (f / fdiv[s]) * fmul[s] + fadd[s] of its K/V matrices and holding nk[s] rows.
 *   out[f*Nq + q, h*d + :] (+)= out_scale * softmax_k(scale * q.k) v
 */
typedef struct mvb_attention_desc {
  const void* q; long long ldq;
  int NF, Nq, heads, d, dp;
  float scale;
  int nseg;
  const void* k[2]; const void* v[2]; long long ldkv[2]; long long kv_rows[2];
  int nk[2]; int fdiv[2]; long long fmul[2]; long long fadd[2];
  void* out; long long ldo;
  float out_scale;
  int accumulate;
  int v_ones_col;  /* every V row holds 1.0 at column h*dp + d (dp > d): the P.V MMA also yields the softmax row sum */
  int variant;     /* 0: default kernel; 2: split-KV kernel (dp <= 64 only; two independent softmax groups, kept for A/B runs) */
} mvb_attention_desc;

int mvb_op_attention(const mvb_attention_desc* desc, void* stream);
/* Measurement aid (no reference equivalent): while `device_buffer` (>= 9*32*8 int64 on the device) is set, CTA (0,0,0) of every
 * ping-pong attention launch (head dim <= 64) writes the SM clock at each phase of its first 32 key/value tiles:
 * [role][tile][slot], role 4t+q = softmax warp of query tile t, lane quarter q (slots: 0 wait S, 1 S ready, 2 scores in registers, 3 row max,
 * 4 exponentials done, 5 previous P.V complete, 6 P published), role 8 = the MMA-issuing warp (per query tile t, slots 4t..4t+3:
 * S_t(j) freed, S_t(j+1) issued, P_t(j) ready, P_t(j)V(j) issued). NULL switches it off. tools/gpu_attention_trace.py. */
int mvb_debug_attention_trace(long long* device_buffer);
/* Encoded TMA descriptors (cuTensorMapEncodeTiled: 2-6 per GEMM / convolution, 5 per attention call) are memoized process-wide by
 * (address, extents, strides, box, swizzle); the engine's workspace arena reproduces its addresses on every forward of the same
 * shapes, so from the second window-step on they are hash lookups. Hits / misses since the library was loaded (no reference
 * equivalent; MVB_TMAP_CACHE=0 disables the cache). */
int mvb_tensor_map_cache_stats(unsigned long long* hits, unsigned long long* misses);

/* Temporal self-attention over the frame axis (musev/models/temporal_transformer.py:241-273 ->
 * musev/models/attention.py:293-365 -> AttnProcessor2_0). qkv: [B, T, HW, 3*heads*dp] (q | k | v). */
int mvb_op_temporal_attention(const void* qkv, int ld, int B, int T, int HW, int heads, int d, int dp, float scale,
                              void* out, int ldo, void* stream);

/* GroupNorm (+SiLU) on channels-last fp16 [NF, HW, C0 (+C1)] (F.group_norm at diffusers models/resnet.py:641,662;
 * musev/models/resnet.py:57-74; temporal_transformer.py:117). frames_per_stat = 1: per-frame statistics;
 * = T: the reference's 5-D GroupNorm over (c/g, t, h, w). `scratch` >= NF*65*groups*2 floats. */
int mvb_op_groupnorm(const void* x0, int c0, const void* x1, int c1, int NF, int HW, int groups, int frames_per_stat,
                     float eps, const float* gamma, const float* beta, int silu, void* y, float* scratch, void* stream);

/* The same GroupNorm as ONE persistent launch (statistics, finalize and apply separated by grid barriers; the engine's
 * default path). `barrier_word`: a zero-initialised device uint32 owned by the caller; `*arrivals`: host-side count of the
 * arrivals that word has seen, updated by the call (calls sharing a word must be issued on one stream). */
int mvb_op_groupnorm_fused(const void* x0, int c0, const void* x1, int c1, int NF, int HW, int groups, int frames_per_stat,
                           float eps, const float* gamma, const float* beta, int silu, void* y, float* scratch,
                           unsigned int* barrier_word, unsigned int* arrivals, void* stream);

/* LayerNorm over the channel axis of [M, C] fp16 (F.layer_norm at musev/models/attention.py:193,346-350,399). */
int mvb_op_layernorm(const void* x, long long M, int C, float eps, const float* gamma, const float* beta, void* y,
                     void* stream);

/* Fused overlap mean + classifier-free guidance + DDIM step
 * (musev/pipelines/pipeline_controlnet.py:2079,2101-2117; musev/schedulers/scheduling_ddim.py:198-295).
 *   eps = eps_sum / counter[t];  cfg: eps = uncond + g * (text - uncond)   (eps_sum fp32 [2B,C,T,HW], uncond first)
 *   x0 from eps / v / sample prediction (prediction_type 0 / 1 / 2), optional clip to +-clip_range (<= 0: off),
 *   x_prev = sqrt(a_prev) x0 + sqrt(1 - a_prev - std^2) eps (+ std * variance_noise when eta > 0).
 * cfg = 0: eps_sum is a single [B,C,T,HW] prediction -- this is plain `DDIMScheduler.step`. counter, variance_noise,
 * eps_out, x0_out may be NULL. latents fp32 (is_f32) or fp16 [B,C,T,HW]. */
int mvb_fuse_cfg_ddim(const float* eps_sum, const float* counter, const void* latents_in, void* latents_out,
                      int is_f32, int B, int C, int T, int HW, int cfg, float guidance_scale, float alpha_prod_t,
                      float alpha_prod_t_prev, int prediction_type, float clip_range, int use_clipped_model_output,
                      float std_dev_t, const float* variance_noise, float* eps_out, float* x0_out, void* stream);

/* Fused overlap mean + classifier-free guidance + an AFFINE sampler step ("other samplers ... pure
 * elementwise epilogues like DDIM"): x_prev = c_x x + c_e eps + c_n noise, aux = a_x x + a_e eps. Covers, with
 * host-computed scalars and no clipping / thresholding,
 *   EulerDiscreteScheduler.step  musev/schedulers/scheduling_euler_discrete.py:47-170 (default sampler of the predictor,
 *                                pipeline_controlnet_predictor.py:258-261): c_x 1, c_e sigma_next - sigma_hat, aux = x0;
 *   LCMScheduler.step            musev/schedulers/scheduling_lcm.py:196-312: prev = sqrt(a_prev) denoised + sqrt(1-a_prev) z,
 *                                aux = denoised = c_out x0 + c_skip x;
 *   DDIMScheduler.step           eps prediction without clipping, any eta.
 * Same tensor conventions as mvb_fuse_cfg_ddim; noise, aux_out, eps_out may be NULL. */
int mvb_fuse_cfg_affine(const float* eps_sum, const float* counter, const void* latents_in, void* latents_out, int is_f32,
                        int B, int C, int T, int HW, int cfg, float guidance_scale, float c_x, float c_e, float c_n,
                        const float* noise, float a_x, float a_e, float* aux_out, float* eps_out, void* stream);

/* eps_sum[:, :, frames[i]] += eps_window[:, :, src_t0 + i] (musev/pipelines/pipeline_controlnet.py:2068-2078).
 * eps_window [2B, C, Tw, HW] fp32/fp16; frames_dev: device int32[nframes]. */
int mvb_accumulate_window(float* eps_sum, int B2, int C, int T, int HW, const void* eps_window, int is_f32, int Tw,
                          int src_t0, const int* frames_dev, int nframes, void* stream);


/* ---------------------------------------------------------------------------------------------------------
 * Whole-model level: the denoiser `UNet3DConditionModel` (musev/models/unet_3d_condition.py:179-1280).
 *
 * mvb_config mirrors the constructor arguments that change the computation (unet_3d_condition.py:213-258) for the
 * released presets of musev/models/unet_loader.py:232-268. One handle per device, not re-entrant (the reference is
 * driven by a single Python thread on the default stream).
 */
typedef struct mvb_config {
  int in_channels, out_channels;
  int num_blocks;                 /* len(block_out_channels), <= 4 */
  int block_out_channels[4];
  int layers_per_block;
  int heads;                      /* `attention_head_dim` of the reference config (it is the head COUNT) */
  int cross_attention_dim;
  int norm_num_groups;
  float norm_eps;
  int need_transformer_in;
  int use_anivv1_cfg;
  int resnet_2d_skip_time_act;
  int keep_vision_condtion;
  int need_refer_emb;
  int ip_adapter_cross_attn;
  int need_t2i_ip_adapter;        /* reference-only self attention toward the vision-condition frame(s) */
} mvb_config;

typedef struct mvb_handle mvb_handle;

#define MVB_MAX_REFER 16
/* Arguments of one `UNet3DConditionModel.forward` call (unet_3d_condition.py:773-803). Video tensors are the
 * reference's NCTHW layout, fp16 or fp32 (flag per tensor group), contiguous. */
typedef struct mvb_unet_args {
  const void* sample; int sample_is_f32;       /* [B, in_channels, T, H, W], vision-condition frames included */
  int B, T, H, W;
  float timestep;
  const void* encoder_hidden_states; int ehs_is_f32; int n_text;   /* [B, n_text, cross_attention_dim] */
  int has_sample_index;                        /* sample_index is not None */
  int n_vis_cond, vis_cond_first;              /* vision_conditon_frames_sample_index = [first, first + n) ; n = 0: None */
  float sample_frame_rate;
  const void* vision_clip_emb; int clip_is_f32; int n_clip; float ip_adapter_scale;  /* [B, n_clip, cross_dim] or NULL */
  int n_refer;                                 /* 0 or the number of down_block_refer_embs */
  const void* refer_embs[MVB_MAX_REFER]; int refer_t[MVB_MAX_REFER], refer_h[MVB_MAX_REFER], refer_w[MVB_MAX_REFER];
  const void* mid_refer_emb; int mid_refer_t, mid_refer_h, mid_refer_w;
  int refer_is_f32;                            /* refer maps are [B, C, t, h, w] */
  int n_down_residuals;                        /* ControlNet: 0 or 1 + num_blocks*(layers_per_block+1) - 1 tensors */
  const void* down_residuals[MVB_MAX_REFER];   /* [(B T), C, h, w] */
  const void* mid_residual; int residual_is_f32;
  int skip_temporal_layers;
  void* out; int out_is_f32;                   /* [B, out_channels, T, H, W] */
} mvb_unet_args;

/* Reference: UNet3DConditionModel.__init__ (unet_3d_condition.py:213-610). */
int mvb_create(const mvb_config* cfg, int device, mvb_handle** out);
void mvb_destroy(mvb_handle* h);
/* Reference: from_pretrained_2d / load_state_dict (unet_3d_condition.py:1284-1637): feed every tensor of the
 * reference state_dict by its reference name; the library packs it into its kernel layout on the device. */
int mvb_load_weight(mvb_handle* h, const char* name, const void* device_ptr, int is_f32, const long long* shape, int ndim);
/* Batched form (the `mvb_load_weights(h, const mvb_named_tensor*, n)` of SURVEY.md 8b): every entry is validated, then the
 * whole batch is packed by one kernel launch; synchronous (the sources may be freed on return). Entries not in a batch can
 * still be fed through mvb_load_weight; mvb_finalize checks that the union covers the schema. */
typedef struct mvb_named_tensor {
  const char* name;          /* reference state_dict key */
  const void* device_ptr;    /* contiguous tensor on the handle's device */
  int is_f32;                /* 1: float32, 0: float16 */
  int ndim;                  /* 0..5 */
  long long shape[5];
} mvb_named_tensor;
int mvb_load_weights(mvb_handle* h, const mvb_named_tensor* tensors, int n);
/* Checks that every tensor of the schema has been loaded. */
int mvb_finalize(mvb_handle* h);
int mvb_num_params(mvb_handle* h);
/* Bytes of scratch `mvb_unet_forward` needs for these shapes (activation arena; the caller owns it). */
long long mvb_workspace_bytes(mvb_handle* h, const mvb_unet_args* args);
/* Reference: UNet3DConditionModel.forward (unet_3d_condition.py:773-1280). Asynchronous on `stream`. */
int mvb_unet_forward(mvb_handle* h, const mvb_unet_args* args, void* workspace, long long workspace_bytes, void* stream);
const char* mvb_handle_error(mvb_handle* h);
/* Debug aid for bisecting parity: layer outputs of the last forward, fp16 [rows, C] inside the caller's workspace. */
int mvb_debug_num_taps(mvb_handle* h);
int mvb_debug_tap(mvb_handle* h, int i, char* name, int name_cap, const void** ptr, long long* rows, int* C);


/* ---------------------------------------------------------------------------------------------------------
 * ControlNet encoder per window-step. Reference: diffusers `ControlNetModel.forward`
 * (diffusers/src/diffusers/models/controlnet.py:645-852) as `get_controlnet_emb` calls it
 * (musev/pipelines/pipeline_controlnet.py:1238-1262): frames on the batch axis, the prompt embedding repeated per
 * frame, the condition embedding pre-computed once per call (`controlnet_cond_latents`, :1258) -- the embedding conv
 * stack itself (controlnet.py:101-112) is a one-shot on the 512x512 condition image and stays with the caller.
 * The handle is created with `mvb_create_controlnet` from the same `mvb_config` (UNet-only switches ignored) and fed
 * with `mvb_load_weight` by the reference names (`controlnet_cond_embedding.*` is not part of it). */
#define MVB_CONTROLNET_MAX_OUT 13
typedef struct mvb_controlnet_args {
  const void* sample; int sample_is_f32;            /* [NF, in_channels, H, W] */
  int NF, H, W;
  float timestep;
  const void* encoder_hidden_states; int ehs_is_f32; int n_text;   /* [NF, n_text, cross_attention_dim] */
  const void* cond_latents; int cond_is_f32;        /* [NF, block_out_channels[0], H, W] */
  int n_out;                                        /* 1 + sum over blocks (layers_per_block + has_downsampler) + 1 (mid) */
  float scales[MVB_CONTROLNET_MAX_OUT];             /* conditioning_scale, times logspace(-1, 0) in guess mode (:826-833) */
  void* outs[MVB_CONTROLNET_MAX_OUT];               /* down residuals in order, then the mid residual: [NF, C_k, h_k, w_k] */
  int out_is_f32;
  int out_frames;                                   /* ReferenceNet only (`num_frames`, referencenet.py:1041-1049): outputs are
                                                       [NF / out_frames, C_k, out_frames, h_k, w_k]; 0 or 1 = (b t) c h w */
} mvb_controlnet_args;
int mvb_create_controlnet(const mvb_config* cfg, int device, mvb_handle** out);
long long mvb_controlnet_workspace_bytes(mvb_handle* h, const mvb_controlnet_args* args);
int mvb_controlnet_forward(mvb_handle* h, const mvb_controlnet_args* args, void* workspace, long long workspace_bytes,
                           void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * ReferenceNet one-shot.
 * Reference: `ReferenceNet2D.forward` (musev/models/referencenet.py:640-1127) as `get_referencenet_emb` calls it once per
 * pipeline call at step 0 (musev/pipelines/pipeline_controlnet.py:867-964,1883-1899): the reference-image VAE latents
 * flattened to (b t) c h w, timestep 0, `encoder_hidden_states` = the IP-Adapter image tokens (or the prompt), returning the
 * 12 down-block feature maps + the mid-block map as [b, C, t, h, w] (`need_block_embs=True`, `return_ndim=5`; the up blocks
 * are dropped, referencenet.py:624-636). Same SD-1.5 encoder as the ControlNet above but built from the musev blocks
 * (LayerNorm eps 0 / 1e-5 / 0), no condition embedding, no zero convolutions: the maps are the taps.
 * Uses `mvb_controlnet_args` (`cond_latents` NULL, `scales` ignored, `out_frames` = num_frames). */
int mvb_create_referencenet(const mvb_config* cfg, int device, mvb_handle** out);
long long mvb_referencenet_workspace_bytes(mvb_handle* h, const mvb_controlnet_args* args);
int mvb_referencenet_forward(mvb_handle* h, const mvb_controlnet_args* args, void* workspace, long long workspace_bytes,
                             void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Accounting (used by bench.py). category: 0 conv/linear GEMM, 1 spatial attention, 2 temporal attention,
 * 3 GroupNorm, 4 LayerNorm, 5 other; -1 = all. */
long long mvb_launch_count(int category);
/* When enabled every launcher brackets its kernel with two CUDA events on the launching stream. */
void mvb_profile_enable(int on);
/* Synchronises the device, sums the recorded event pairs per category (6 entries each) and clears them. */
int mvb_profile_collect(double* ms_per_category, long long* scopes_per_category);

/* ------------------------------------------------------------------------------------------------------------------
 * VAE decode, the step after the path.
 * Reference: `MusevControlNetPipeline.decode_latents` (musev/pipelines/pipeline_controlnet.py:233-238, called in T-segments
 * at :2157-2171) -> diffusers `decode_latents` (pipelines/stable_diffusion/pipeline_stable_diffusion_img2img.py:486-495:
 * latents / scaling_factor, `vae.decode`, image / 2 + 0.5, clamp(0, 1)) -> `AutoencoderKL.decode`
 * (models/autoencoder_kl.py:275-302: post_quant_conv + `Decoder.forward`, models/vae.py:265-316: conv_in, UNetMidBlock2D
 * with one single-head attention, 4 UpDecoderBlock2D, GroupNorm + SiLU + conv_out).
 * The handle is created from an `mvb_config` whose block_out_channels are the VAE's (128, 256, 512, 512 for SD-1.5),
 * in_channels = latent channels, out_channels = image channels, norm_eps 1e-6; weights by the `AutoencoderKL.state_dict()`
 * names `post_quant_conv.*` and `decoder.*`. */
typedef struct mvb_vae_decode_args {
  const void* latents; int latents_is_f32;   /* [N, latent_channels, h, w] (frames on the batch axis) */
  int N, h, w;
  float latent_scale;                        /* multiplies the latents first: 1 / scaling_factor (or 1 for plain vae.decode) */
  void* out; int out_is_f32;                 /* [N, out_channels, 8h, 8w] */
  int postprocess;                           /* 1: out = clamp(image / 2 + 0.5, 0, 1) (decode_latents), 0: raw decoder output */
} mvb_vae_decode_args;
int mvb_create_vae_decoder(const mvb_config* cfg, int device, mvb_handle** out);
long long mvb_vae_decode_workspace_bytes(mvb_handle* h, const mvb_vae_decode_args* args);
int mvb_vae_decode(mvb_handle* h, const mvb_vae_decode_args* args, void* workspace, long long workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MUSEV_B200_H_ */
