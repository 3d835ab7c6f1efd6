// Model runner: the UNet2DCondition / AutoencoderKL op graphs expressed as sequences of the sm_100a
// kernels in ops.h.  A model owns its packed weights (device memory) and nothing else; activations live
// in a caller-provided workspace carved by a two-ended bump allocator (persistent block outputs grow
// from the bottom, per-block temporaries from the top).  The same code path runs "dry" to size the
// workspace, so the plan and the execution cannot drift.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/gyre_b200.h"
#include "ops.h"

namespace gyre {

struct Exec {
  cudaStream_t st = nullptr;
  bool dry = false;
  uint8_t* base = nullptr;
  size_t cap = 0;
  size_t bottom = 0;     // persistent bytes used
  size_t top = 0;        // scratch bytes used (from the end)
  size_t peak = 0;
  bool overflow = false;
  void* alloc_p(size_t bytes);
  void* alloc_s(size_t bytes);
  void reset_scratch() { top = 0; }
  __half* p16(size_t elems) { return static_cast<__half*>(alloc_p(elems * 2)); }
  __half* s16(size_t elems) { return static_cast<__half*>(alloc_s(elems * 2)); }
  float* s32(size_t elems) { return static_cast<float*>(alloc_s(elems * 4)); }
};

struct NormW { float* g = nullptr; float* b = nullptr; int C = 0; };
struct LinW { __half* w = nullptr; float* bias = nullptr; int N = 0, K = 0; };
struct Conv3W {
  __half* wp = nullptr;
  __half* wp_up = nullptr;   // upsampler convs only: the four 2x2 phase kernels of the folded nearest-2x upsample
  float* bias = nullptr;
  int Cin = 0, Cout = 0;
};
struct SmallConvW { float* w = nullptr; float* bias = nullptr; int Cin = 0, Cout = 0; };   // fp32 1x1 on <= 8 channels

struct ResnetW {
  NormW n1, n2;
  Conv3W c1, c2;
  LinW sc;
  bool has_sc = false;
  int cin = 0, cout = 0;
  int temb_off = -1;    // column offset into the fused time_emb_proj output (-1: no temb)
};

// A Linear that directly follows a LayerNorm, prepared for the folded form (ops.h, Epilogue::ln_rowstat): gamma-scaled
// copy of the packed weight, its column sums and bias + beta @ W^T.  Derived data: rebuilt after any (re)load.
struct LnFoldW { __half* w = nullptr; float* colsum = nullptr; float* bias = nullptr; };

struct TBlockW {          // one BasicTransformerBlock
  NormW ln1, ln2, ln3;
  LinW qkv, o1, q2, kv2, o2, geglu, ff2;
  LnFoldW qkv_ln, q2_ln, geglu_ln;
  int flat = 0;           // index among all BasicTransformerBlocks of the model (bound-context cache slot)
};

struct TransformerW {     // Transformer2DModel: GroupNorm, proj_in, `depth` blocks, proj_out
  NormW gn;
  LinW proj_in, proj_out;
  std::vector<TBlockW> blocks;
  int C = 0, heads = 0;
};

struct VaeAttnW {
  NormW gn;
  LinW qk, v, proj;      // qk fused [2C, C]; v [C, C] (its bias is applied after P.V: softmax rows sum to 1)
  float* v_bias = nullptr;
  int C = 0;
};

enum ParamKind { P_F32 = 0, P_LINEAR, P_CONV3, P_GEGLU_W, P_GEGLU_B, P_F32MAT, P_CONV3_UP };

struct ParamSlot {
  int kind = P_F32;
  void* dst = nullptr;
  int64_t shape[4] = {0, 0, 0, 0};
  int ndim = 0;
  int ld = 0;            // P_LINEAR: destination row pitch (elements)
  void* dst2 = nullptr;  // P_CONV3_UP: phase-packed copy
  bool loaded = false;
};

class Model {
 public:
  virtual ~Model();
  int load(const char* key, const void* data, int dtype, const int64_t* shape, int ndim, cudaStream_t st);
  int finalize();
  int device() const { return device_; }
  virtual bool is_unet() const = 0;
  virtual int kind() const { return is_unet() ? 0 : 1; }   // 0 unet, 1 vae, 2 clip text encoder

 protected:
  Model();
  void* dalloc(size_t bytes);
  void reg(const std::string& key, int kind, void* dst, std::initializer_list<int64_t> shape, int ld = 0);
  void reg_norm(const std::string& p, int C, NormW* n);
  void reg_linear(const std::string& p, int N, int K, bool bias, LinW* l);
  void reg_conv3(const std::string& p, int Cin, int Cout, Conv3W* c, bool upsampler = false);
  // conv3x3 applied to the nearest-2x upsample of x [B, H, W, c.Cin] -> out [B, 2H, 2W, c.Cout]
  int upsample_conv(Exec& ex, const Conv3W& c, const __half* x, int B, int H, int W, __half* out);
  void reg_resnet(const std::string& p, int cin, int cout, bool temb, ResnetW* r);
  int resnet(Exec& ex, const ResnetW& r, const __half* x1, int C1, const __half* x2, int C2, int B, int H, int W,
             float eps, const __half* temb_all, int temb_ld, __half* out);
  int ensure_device();
  virtual void on_load(const std::string& key) { (void)key; }   // a parameter was (re)loaded
  // epilogue descriptor carrying this model's stream-K scratch
  Epilogue ep_out(__half* out, int ldo, const float* bias = nullptr, const __half* residual = nullptr, int ldr = 0,
                  int act = 0) const;

  // GroupNorm statistics from the producing convolution (Epilogue::gn_out).  gn_offer() asks the conv that is about to
  // write e.out for them when the shape allows it and remembers the tensor; gnorm() - every GroupNorm of the models goes
  // through it - takes the one-pass path when its input is exactly that tensor, the two-pass kernels otherwise, and
  // forgets the offer either way.  Anything that rewrites a tensor in place between the two must call gn_forget().
  void gn_offer(Epilogue& e, int B, int H, int W, int Cout, int stride, int pad);
  int gnorm(Exec& ex, const __half* x1, int C1, const __half* x2, int C2, int B, int HW, float eps, const NormW& n, bool silu,
            __half* out);
  void gn_forget() { gn_pre_.x = nullptr; }
  struct GnPre { const __half* x = nullptr; int B = 0, HW = 0, C = 0, nparts = 0; };
  GnPre gn_pre_;
  float* gn_pre_buf_ = nullptr;     // partials [B][nparts][G][2], then kGnPreStats floats of folded (mean, rstd)
  size_t gn_pre_floats_ = 0;        // capacity of the partials part

  std::unordered_map<std::string, ParamSlot> slots_;
  std::vector<void*> allocs_;
  int device_ = 0;
  int groups_ = 32;
  bool alloc_failed_ = false;
  // fused time_emb_proj of every resnet: [temb_total, temb_dim] + bias
  LinW temb_proj_;
  int temb_total_ = 0;
  int temb_dim_ = 0;
  // stream-K scratch shared by all GEMM / conv launches of this model (they are serialised on one stream)
  void* sk_ws_ = nullptr;
  size_t sk_ws_bytes_ = 0;
  int* sk_flags_ = nullptr;
};

// ControlNet mode of UNetModel::forward: the conditioning image in, one residual per skip + the mid residual out (NCHW)
struct ControlNetIO {
  const __half* cond = nullptr;        // [B, conditioning_channels, 8H, 8W] NCHW fp16
  __half* const* down_out = nullptr;   // host array of num_skips() device pointers
  int n_down = 0;
  __half* mid_out = nullptr;
};

class UNetModel : public Model {
 public:
  explicit UNetModel(const gyre_b200_unet_config& cfg);
  bool is_unet() const override { return true; }
  int num_transformer_blocks() const { return static_cast<int>(tblocks_.size()); }
  // dry == true sizes the workspace (ex.peak) without launching anything
  int forward(Exec& ex, const __half* sample, const int64_t* t, const __half* ctx, const __half* add_cond, int B, int H,
              int W, int L, const int32_t* tome_r, __half* out, const ControlNetIO* cn = nullptr);
  bool is_controlnet() const { return cfg_.controlnet != 0; }
  // The caller promises that the next forwards' batches are [x ; x] with equal timesteps in both halves (what
  // CFGUNet_Parallel feeds the UNet, gyre/pipeline/unet/cfg.py:47-57): everything before the first cross-attention is
  // then computed once for both halves.
  void set_cfg_duplicate(bool on) { cfg_dup_ = on; }
  // Binds a text context for the following forwards (the reference binds the embeddings once per request:
  // UNetWithEmbeddings, gyre/pipeline/unet/core.py:253-259): the cross-attention K/V projections of every
  // transformer block depend only on ctx, so they are computed here ONCE instead of once per step.
  // ctx == nullptr drops the binding.
  int set_context(const __half* ctx, int B, int L, cudaStream_t st);
  // ControlNet residuals for the NEXT forward only (diffusers' down_block_additional_residuals /
  // mid_block_additional_residual, passed through by gyre/pipeline/unet/core.py:213-239): NCHW fp16 tensors, one per
  // skip connection in push order (conv_in, every resnet/attention pair, every downsampler) plus one for the mid block
  // output.  They are added to the skip tensors AFTER the down path has run (they do not change the down path itself).
  int set_control_residuals(const __half* const* down, int n_down, const __half* mid);
  int num_skips() const;
  // T2I-adapter states for the NEXT forward only (`adapter_states=`, gyre/pipeline/t2i_adapter/unet_patcher.py:21-60,
  // 95-110): one NCHW fp16 tensor per down block, added IN PLACE to the block's last hidden state before its
  // downsampler - i.e. to the block's last skip tensor and to everything computed from it.
  int set_adapter_states(const __half* const* states, int n);
  ~UNetModel() override;

 private:
  int transformer(Exec& ex, const TransformerW& t, const __half* x, int B, int HW, const __half* ctx, int L, int r,
                  __half* out, bool share = false);
  void on_load(const std::string& key) override;
  int refold_layernorms(cudaStream_t st);
  bool cfg_dup_ = false;    // set_cfg_duplicate
  bool ln_fuse_ = false;    // LayerNorm folded into the GEMMs around it (tunable LN_FUSE at creation)
  bool ln_dirty_ = true;    // a transformer-block parameter changed since the folded weights were derived
  gyre_b200_unet_config cfg_;
  Conv3W conv_in_;      // Cin = in_channels (4/5/9): the input is staged NHWC with the channel pitch padded to 8
  LinW time1_, time2_;
  LinW add1_, add2_;    // add_embedding (text_time conditioning), only when cfg_.addition_embed_dim > 0
  int n_tblocks_flat_ = 0;
  std::vector<const __half*> ctrl_down_;   // pending ControlNet residuals (consumed by the next forward)
  const __half* ctrl_mid_ = nullptr;
  std::vector<const __half*> adapter_;     // pending T2I-adapter states (consumed by the next forward)
  std::vector<ResnetW> resnets_;        // in module execution order
  std::vector<TransformerW> tblocks_;   // in module execution order (== ToMe r-list order)
  std::vector<Conv3W> downs_, ups_;
  NormW norm_out_;
  Conv3W conv_out_;
  // ControlNet only: conditioning embedding (conv_in, 6 blocks, conv_out) and the zero convolutions
  std::vector<Conv3W> cn_embed_;
  std::vector<LinW> cn_down_;
  LinW cn_mid_;
  // bound-context cache: per transformer block [ctx_B_ * ctx_L_, 2C] fp16 (k | v)
  std::vector<__half*> kv_cache_;
  std::vector<size_t> kv_cache_elems_;
  int ctx_B_ = 0, ctx_L_ = 0;
};

class VAEModel : public Model {
 public:
  explicit VAEModel(const gyre_b200_vae_config& cfg);
  bool is_unet() const override { return false; }
  int decode(Exec& ex, const __half* z, int B, int h, int w, bool postprocess, __half* img, uint8_t* img_u8);
  int encode(Exec& ex, const __half* img, int B, int H, int W, __half* moments);

 private:
  int attn(Exec& ex, const VaeAttnW& a, const __half* x, int B, int HW, __half* out);
  gyre_b200_vae_config cfg_;
  // decoder
  SmallConvW post_quant_;
  Conv3W dec_conv_in_;
  ResnetW dec_mid_[2];
  VaeAttnW dec_attn_;
  std::vector<ResnetW> dec_res_;
  std::vector<Conv3W> dec_ups_;
  NormW dec_norm_out_;
  Conv3W dec_conv_out_;
  // encoder
  Conv3W enc_conv_in_;
  SmallConvW quant_;
  std::vector<ResnetW> enc_res_;
  std::vector<Conv3W> enc_downs_;
  ResnetW enc_mid_[2];
  VaeAttnW enc_attn_;
  NormW enc_norm_out_;
  Conv3W enc_conv_out_;
};

struct ClipLayerW {
  NormW ln1, ln2;
  LinW qkv, out, fc1, fc2;
};

class ClipTextModel : public Model {
 public:
  explicit ClipTextModel(const gyre_b200_clip_config& cfg);
  bool is_unet() const override { return false; }
  int kind() const override { return 2; }
  int forward(Exec& ex, const int64_t* ids, int B, int L, int skip_last, bool final_ln, __half* out);

 private:
  gyre_b200_clip_config cfg_;
  __half* tok_emb_ = nullptr;   // [vocab, C]
  __half* pos_emb_ = nullptr;   // [max_positions, C]
  std::vector<ClipLayerW> layers_;
  NormW final_ln_;
};

// CLIP vision tower + concept embeddings of the safety checker (gyre/pipeline/safety_checkers.py:13-66)
class ClipVisionModel : public Model {
 public:
  explicit ClipVisionModel(const gyre_b200_clip_vision_config& cfg);
  bool is_unet() const override { return false; }
  int kind() const override { return 4; }
  // hidden != nullptr: stop after (num_layers - skip_last) layers and hand out the hidden states [B, tokens, C]
  int forward(Exec& ex, const __half* pixel_values, int B, __half* image_embeds, float* scores, __half* hidden = nullptr,
              int skip_last = 0, bool hidden_only = false);
  int tokens() const { return (cfg_.image_size / cfg_.patch_size) * (cfg_.image_size / cfg_.patch_size) + 1; }

 private:
  gyre_b200_clip_vision_config cfg_;
  int Kp_ = 0;                    // patch-embedding K (3 * P * P) padded to a multiple of 8
  LinW patch_;                    // [C, Kp]
  float* class_emb_ = nullptr;    // [C]
  __half* pos_emb_ = nullptr;     // [tokens, C]
  NormW pre_ln_, post_ln_;
  std::vector<ClipLayerW> layers_;
  LinW proj_;                     // visual_projection [projection_dim, C]
  float* embeds_ = nullptr;       // [num_special + num_concepts, projection_dim]: special-care rows first
  float* thresholds_ = nullptr;   // [num_special + num_concepts] (kept for completeness; the host applies them)
};

// Style T2I-adapter (gyre/pipeline/t2i_adapter/adapter.py:173-199)
class StyleAdapterModel : public Model {
 public:
  explicit StyleAdapterModel(const gyre_b200_style_adapter_config& cfg);
  bool is_unet() const override { return false; }
  int kind() const override { return 5; }
  int forward(Exec& ex, const __half* x, int B, int L, __half* out);

 private:
  gyre_b200_style_adapter_config cfg_;
  std::vector<ClipLayerW> layers_;
  __half* style_emb_ = nullptr;   // [num_token, width]
  NormW ln_pre_, ln_post_;
  LinW proj_;                     // [context_dim, width] (the module's `proj` transposed)
};

// T2I-adapter encoder (gyre/pipeline/t2i_adapter/adapter.py:65-132)
struct AdapterLightW {     // one `extractor` of Adapter_light
  LinW in1, out1;
  std::vector<Conv3W> b1, b2;
  int in_c = 0, inter_c = 0, out_c = 0;
};

struct AdapterBlockW {
  Conv3W down3;            // Downsample with conv (use_conv)
  Conv3W in3, b2_3, sk3;   // ksize 3 variants
  LinW in1, b2_1, sk1;     // ksize 1 variants
  Conv3W b1;               // block1 is always 3x3
  bool down = false, has_in = false, has_skep = false;
  int in_c = 0, out_c = 0;
};

class AdapterModel : public Model {
 public:
  explicit AdapterModel(const gyre_b200_adapter_config& cfg);
  bool is_unet() const override { return false; }
  int kind() const override { return 3; }
  int num_levels() const { return cfg_.num_levels; }
  // image [B, cin / 64, H, W] NCHW fp16 (H, W multiples of 8) -> one NCHW fp16 feature map per level (host pointer array)
  int forward(Exec& ex, const __half* image, int B, int H, int W, __half* const* features);

 private:
  gyre_b200_adapter_config cfg_;
  Conv3W conv_in_;
  std::vector<AdapterBlockW> body_;
  std::vector<AdapterLightW> light_;
  int forward_light(Exec& ex, const __half* image, int B, int H, int W, __half* const* features);
};

}  // namespace gyre
