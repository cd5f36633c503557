// Whole-forward engine: owns the packed weights of one UNet3DConditionModel and launches the fixed kernel sequence
// of musev/models/unet_3d_condition.py:773-1280 on a caller-provided stream and workspace.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/musev_b200.h"

namespace mvb {

struct Mat {
  __half* w = nullptr;   // packed [N, K] fp16 (K-major)
  float* bias = nullptr; // packed [N] fp32 or null
  int N = 0, K = 0;
};
struct Norm {
  float* g = nullptr;
  float* b = nullptr;
  int C = 0;
};
struct TBlock {
  Norm n1, n2, n3;
  Mat qkv1, out1;          // self attention (spatial reference-only / temporal attn1)
  Mat qkv2;                // temporal attn2 (self)
  Mat q2, kv2, kv2_ip;     // spatial attn2 (text cross attention, optional IP-Adapter k/v)
  Mat out2;
  Mat ff1, ff2;            // GEGLU feed-forward
  bool cross = false, has_ip = false;
};
struct Resnet {
  Norm n1, n2;
  Mat conv1, conv2, shortcut;
  int cin = 0, C = 0, temb_off = 0;
  bool has_shortcut = false;
  bool has_temb = true;    // false in the VAE decoder (temb_channels=None)
};
struct TempConv {
  Norm n[4];
  Mat conv[4];
  float tw = 0.f;
  int C = 0;
};
struct SpatialT {
  Norm norm;
  Mat proj_in, proj_out;
  TBlock blk;
  int C = 0;
};
struct TemporalT {
  Norm norm;
  Mat proj_in, proj_out;
  TBlock blk;
  float tw = 0.f;
  int femb_off = 0, C = 0;
};
struct ReferAttn {
  Mat qkv, out;   // K/V of the reference tokens use rows [H*dp, 3*H*dp) of qkv
  int C = 0;
  bool present = false;
};
struct Layer {
  Resnet res;
  TempConv tc;
  SpatialT st;
  TemporalT tt;
  ReferAttn ref;
  bool has_attn = false;
};
struct Block {
  std::vector<Layer> layers;
  Mat sampler;        // downsample (stride 2) or upsample conv
  bool has_sampler = false;
  ReferAttn ref_down; // ReferEmbFuseAttention applied after the downsampler
};

enum LoadKind { LK_MAT, LK_VEC, LK_ABS_SCALAR };
struct Loader {
  LoadKind kind;
  // LK_MAT: dst [.., ld] rows [row0, row0+rows_dst): source [Nsrc, Ksrc]
  __half* dst = nullptr;
  long long ld = 0;
  int rows_dst = 0, kdst = 0;
  int rowmode = 0;  // 0 copy, 1 pad heads (p0 = d, p1 = dp), 2 geglu interleave
  int p0 = 0, p1 = 0;
  int colmode = 0;  // 0 identity (zero fill beyond Ksrc), 1 conv [N, Cin, taps] -> (tap, c)
  int cin = 0, taps = 1;
  int nsrc = 0, ksrc = 0;
  // LK_VEC
  float* vdst = nullptr;
  int vn = 0;       // destination length
  int vmode = 0;    // 0 copy (zero fill beyond source), 2 geglu interleave
  // LK_ABS_SCALAR
  float* host_scalar = nullptr;
  bool loaded = false;
};

struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = true;
  void* alloc(size_t bytes) {
    const size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    if (off > peak) peak = off;
    if (dry) return reinterpret_cast<void*>(size_t(4096) + a);  // fake, never dereferenced
    return (off <= cap) ? base + a : nullptr;
  }
};

class Engine {
 public:
  // kind 0: UNet3DConditionModel; kind 1: ControlNet encoder (diffusers models/controlnet.py);
  // kind 2: ReferenceNet2D encoder + mid block (musev/models/referencenet.py);
  // kind 3: AutoencoderKL decoder (diffusers models/autoencoder_kl.py, vae.py)
  explicit Engine(const mvb_config& cfg, int device, int kind = 0);
  ~Engine();
  int load_weight(const char* name, const void* dev_ptr, int is_f32, const long long* shape, int ndim);
  int load_weights(const mvb_named_tensor* tensors, int n);
  int finalize();
  long long workspace_bytes(const mvb_unet_args& a);
  int forward(const mvb_unet_args& a, void* workspace, long long workspace_bytes, cudaStream_t stream);
  long long vae_workspace_bytes(const mvb_vae_decode_args& a);
  int vae_decode(const mvb_vae_decode_args& a, void* workspace, long long workspace_bytes, cudaStream_t stream);
  long long controlnet_workspace_bytes(const mvb_controlnet_args& a);
  int controlnet_forward(const mvb_controlnet_args& a, void* workspace, long long workspace_bytes, cudaStream_t stream);
  int kind() const { return kind_; }
  const char* error() const { return err_.c_str(); }
  struct Tap { std::string name; const __half* p; long long rows; int C; };
  const std::vector<Tap>& taps() const { return taps_; }
  int num_params() const { return (int)loaders_.size(); }

 private:
  // construction
  void build();
  void build_unet();
  void build_controlnet();
  void build_vae();
  template <typename T> T* slab(size_t n);
  Mat make_mat(int N, int K, bool bias);
  Norm make_norm(const std::string& p, int C);
  void reg_mat(const std::string& name, Mat& m, int row0, int rows_dst, int rowmode, int p0, int p1, int nsrc, int ksrc,
               int colmode = 0, int cin = 0, int taps = 1);
  void reg_vec(const std::string& name, float* dst, int n, int nsrc_expected, int vmode = 0);
  void reg_linear(const std::string& p, Mat& m, int N, int K, bool bias);
  void reg_conv(const std::string& p, Mat& m, int N, int Cin, int taps);
  void build_tblock(const std::string& p, TBlock& b, int C, bool cross);
  void build_resnet(const std::string& p, Resnet& r, int cin, int C, bool has_temb = true);
  void build_tempconv(const std::string& p, TempConv& t, int C);
  void build_spatial(const std::string& p, SpatialT& s, int C);
  void build_temporal(const std::string& p, TemporalT& t, int C);
  void build_refer(const std::string& p, ReferAttn& r, int C);

  // forward helpers (all return false on error, message in err_)
  struct Fwd;
  bool run(const mvb_unet_args& a, Arena& ar, cudaStream_t s);
  bool run_controlnet(const mvb_controlnet_args& a, Arena& ar, cudaStream_t s);
  bool run_vae(const mvb_vae_decode_args& a, Arena& ar, cudaStream_t s);

  mvb_config cfg_;
  int device_ = 0, num_sms_ = 148;
  int kind_ = 0;
  float ln_eps13_ = 0.f;       // LayerNorm eps of norm1 / norm3: 0 in the musev blocks (Q1), 1e-5 in the vanilla diffusers blocks
  int heads_ = 8;
  bool finalized_ = false;
  std::string err_;
  std::vector<Tap> taps_;   // layer outputs of the last forward (pointers into the caller's workspace)
  std::unordered_map<std::string, Loader> loaders_;
  struct OnesInit { float* v_bias; int heads, d, dp; };
  std::vector<OnesInit> ones_init_;   // V-part biases that carry the ones column used for MMA row sums
  float* v_ones_bias(int rows_before_v, int total_rows, int d, int dp);
  char* slab_ = nullptr;
  size_t slab_bytes_ = 0, slab_off_ = 0;
  bool slab_counting_ = true;

  // model
  Mat conv_in_, conv_out_;
  Norm norm_out_;
  Mat time_l1_, time_l2_, frame_l1_, frame_l2_;
  Mat temb_all_, femb_all_;     // concatenated time_emb_proj / frame_emb_proj of every layer
  int temb_total_ = 0, femb_total_ = 0;
  bool has_tin_ = false;
  TemporalT tin_;
  ReferAttn first_ref_, mid_ref_;
  std::vector<Block> down_, up_;
  Resnet mid_res_[2];
  TempConv mid_tc_[2];
  SpatialT mid_st_;
  TemporalT mid_tt_;
  unsigned int* gn_counter_dev_ = nullptr;   // grid-barrier word of the one-launch GroupNorm
  unsigned int gn_base_ = 0;                 // arrivals it has seen (host bookkeeping)
  bool gn_fused_ = false;                    // env MVB_GN_FUSED=1 selects the one-launch GroupNorm
  int* zero_idx_dev_ = nullptr;  // device int[32] scratch for vis-cond frame indices
  float* fidx_dev_ = nullptr;    // device float[64] scratch for timestep / frame index values
  Mat zero_convs_[MVB_CONTROLNET_MAX_OUT];   // ControlNet: controlnet_down_blocks.* then controlnet_mid_block
  int n_zero_convs_ = 0;
  // VAE decoder: post_quant_conv (fp32 [C, C] + bias), single-head mid-block attention (q/k/v/out with bias, GroupNorm)
  float* vae_pq_w_ = nullptr;
  float* vae_pq_b_ = nullptr;
  Norm vae_attn_norm_;
  Mat vae_q_, vae_k_, vae_v_, vae_o_;
};

}  // namespace mvb
