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
} mvb_conv_gemm_desc;

int mvb_op_conv_gemm(const mvb_conv_gemm_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MUSEV_B200_H_ */
