/* gyre_b200 -- C ABI of the Blackwell-native (sm_100a) diffusion sampling path.
 *
 * This is the drop-in boundary for the hot path of stablecabal/gyre.  The
 * reference has no native code: its hot path is Python calling third-party libraries.  Each entry
 * point below cites the reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - every function returns 0 on success and a negative status on failure; the message is
 *     available (per calling thread) from gyre_b200_last_error();
 *   - all pointers are DEVICE pointers unless the name ends in _host; the library never allocates
 *     or frees caller memory; it owns only the packed weights behind a handle;
 *   - all work is enqueued on the caller-supplied CUDA stream (a cudaStream_t passed as void*);
 *     nothing synchronises the device;
 *   - activations handed across this boundary use the reference's layouts (NCHW latents / images,
 *     [B, L, C] token tensors); the token-major NHWC fp16 layout the kernels use is internal;
 *   - handles are bound to the device that was current at creation; functions are re-entrant
 *     across handles; one handle must not be entered by two threads at once (same rule as the
 *     reference: gyre/manager.py:2047-2103 hands one pipeline clone to one thread).
 */
#ifndef GYRE_B200_H
#define GYRE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GYRE_B200_ABI_VERSION 2

typedef void* gyre_b200_stream;  /* cudaStream_t */
typedef struct gyre_b200_model* gyre_b200_handle;

enum { GYRE_B200_F16 = 0, GYRE_B200_F32 = 1, GYRE_B200_BF16 = 2 };

int gyre_b200_abi_version(void);
/* Copies the calling thread's last error message; returns its length. */
int gyre_b200_last_error(char* buf, size_t n);

/* ------------------------------------------------------------------------------------------
 * Measurement hooks (no reference counterpart: the reference only prints wall time and peak VRAM,
 * tests/test_harness.py:155-166).  launch_count: kernels this library has launched in this process.
 * prof_*: when enabled, every op brackets its kernels with CUDA events on the launching stream;
 * prof_read synchronises the device and returns, per kernel family (0 gemm, 1 conv3x3, 2 attention,
 * 3 groupnorm, 4 layernorm, 5 softmax, 6 elementwise, 7 tome), the launch-group count, summed
 * device milliseconds, algorithmic FLOPs and algorithmic bytes.  Not for use under graph capture.
 * ------------------------------------------------------------------------------------------ */
unsigned long long gyre_b200_launch_count(void);
int gyre_b200_prof_enable(int on);
int gyre_b200_prof_reset(void);
int gyre_b200_prof_read(int family, unsigned long long* count, double* ms, double* flops, double* bytes);
/* Sum over the family's recorded launches of max(flops / peak_tflops, bytes / peak_gbs) in ms: what they would take if
 * every launch ran on whichever roof binds it (a family mixes tensor-bound and HBM-bound shapes). */
int gyre_b200_prof_read_roofline(int family, double peak_tflops, double peak_gbs, double* ideal_ms);
/* Measurement hook: while a device buffer is registered, CTA 0 of the lean-softmax attention kernel writes clock64
 * time stamps of its pipeline events into it (layout: scripts/attn_trace.py); NULL switches it off.  Not part of
 * the product path. */
int gyre_b200_debug_attention_trace(long long* dev_buf, int capacity);

/* Kernel-selection knobs for A/B measurement (no reference counterpart).  Names: "ATT_VARIANT" (softmax
 * variant bit flags of the d<=64 flash kernel), "PDL" (programmatic dependent launch), "GELU_FAST",
 * "GN_CHUNKS", "GN_PHASE", "UPCONV_FOLD" (nearest-2x upsample folded into four 2x2 phase convolutions),
 * "CTX_KV_CACHE" (cross-attention K/V projections of an unchanged text context are reused across
 * steps), "XATTN" (short-key attention kernel), "ATT_D128", "MCAST" (GEMM / conv CTA pairs: 0 off,
 * 1 TMA-multicast pairs, 2 tcgen05 cta_group::2 pairs where they win [default], 3 cta_group::2 pairs
 * everywhere), "STREAMK", "FORCE_BN", "GEMM_STAGES" (cap of the operand ring depth), "DEBUG" (GEMM
 * measurement bits: 1 no output stores, 2 no epilogue - both give WRONG results -, 4 all-variants image),
 * "LN_SUB" (LayerNorm: several rows per warp for C <= 320 [1, default] / <= 640 [2]), "GN_THREADS", "LN_FUSE"
 * (LayerNorm folded into the GEMMs around it; read when a UNet is created [derived weights] and at every forward),
 * "CFG_SHARE" (0: ignore gyre_b200_unet_set_cfg_duplicate), "GN_FUSE" (GroupNorm statistics produced by the epilogue of
 * the 3x3 convolution that writes the tensor [1, default] instead of a statistics pass over it), "SK_MIN" (stream-K floor:
 * k-iterations x tile width below which a launch keeps whole tiles [14000]).
 * Every knob also reads GYRE_B200_<NAME> from the environment at first use.  Results stay within the
 * documented tolerances for every setting except the DEBUG store / epilogue bits. */
int gyre_b200_set_tunable(const char* name, int value);
/* Measurement only: clocks one thread needs to issue and retire `reps` tcgen05.mma (M=128, K=16, N=n) round-robin
 * over `naccs` TMEM accumulators, A from shared memory (a_tmem = 0) or TMEM; out_dev[blocks] (int64, device). */
int gyre_b200_debug_mma_bench(int n, int naccs, int a_tmem, int reps, int blocks, long long* out_dev,
                              gyre_b200_stream stream);
int gyre_b200_get_tunable(const char* name, int* value);

/* ------------------------------------------------------------------------------------------
 * UNet  (replaces diffusers UNet2DConditionModel.forward as called from
 *        gyre/pipeline/unet/core.py:274 through the DiffusersUNet protocol,
 *        gyre/pipeline/unet/types.py:30-39; ToMe variant nonfree/tome_unet.py:229-247)
 * ------------------------------------------------------------------------------------------ */
typedef struct gyre_b200_unet_config {
  int32_t in_channels;             /* 4, 9 (inpaint), 5 (depth)  unified_pipeline.py:186          */
  int32_t out_channels;            /* 4                                                            */
  int32_t num_levels;              /* <= 4                                                         */
  int32_t block_out_channels[4];   /* (320,640,1280,1280)        gyre/ldm_config/v1-inference.yaml */
  int32_t num_heads[4];            /* diffusers-0.16 `attention_head_dim` == number of heads       */
  int32_t attn_levels[4];          /* 1: CrossAttn block at this level                             */
  int32_t layers_per_block;        /* 2                                                            */
  int32_t cross_attention_dim;     /* 768 / 1024                                                   */
  int32_t norm_num_groups;         /* 32                                                           */
  float norm_eps;                  /* 1e-5                                                         */
  int32_t use_linear_projection;   /* SD2.x                                                        */
  int32_t upcast_attention;        /* SD2.1-768 (softmax/QK^T are fp32 in this library regardless) */
  /* SDXL-style topologies (BASELINE config 5; no counterpart in the reference, public diffusers config names): */
  int32_t transformer_depth[4];    /* `transformer_layers_per_block` per level; 0 is read as 1; the mid block
                                      uses the last level's value                                            */
  int32_t addition_embed_dim;      /* 0: none.  > 0: input width of `add_embedding.linear_1` (text_time
                                      conditioning: pooled text embedding ++ sinusoids of the 6 time ids)    */
  /* ControlNet encoder (gyre/pipeline/controlnet/models.py:97-544): the UNet's conv_in / down blocks / mid block plus
   * `controlnet_cond_embedding` (conv stack 3 -> 16/32/96/256 -> C0 on the 8x larger conditioning image), one 1x1
   * `controlnet_down_blocks.k` per skip tensor and `controlnet_mid_block`; no up path.  0: plain UNet. */
  int32_t controlnet;
  int32_t conditioning_channels;   /* 3 */
} gyre_b200_unet_config;

int gyre_b200_unet_create(const gyre_b200_unet_config* cfg, gyre_b200_handle* out);
/* ControlNetModel.forward (controlnet/models.py:420-544) of a handle created with cfg.controlnet = 1:
 * sample [B, Cin, H, W], cond [B, conditioning_channels, 8H, 8W], ctx [B, L, Cc] fp16 -> the gyre_b200_unet_num_skips(h)
 * residual tensors (NCHW fp16, shapes of the UNet's skips, device pointers in the HOST array down_out) and the mid
 * residual [B, C_last, H/8, W/8] - ready for gyre_b200_unet_set_control_residuals of the UNet they condition.
 * Workspace: gyre_b200_unet_workspace_bytes on the same handle. */
int gyre_b200_controlnet_forward(gyre_b200_handle h, const void* sample, const int64_t* timestep, const void* ctx,
                                 const void* cond, int batch, int height, int width, int ctx_len, void* const* down_out,
                                 int n_down, void* mid_out, void* workspace, size_t workspace_bytes,
                                 gyre_b200_stream stream);
/* Number of transformer blocks == length of the ToMe r-list (nonfree/tome_unet.py:243). */
int gyre_b200_unet_num_transformer_blocks(gyre_b200_handle h);

/* Hands one parameter to the library under its diffusers state-dict key,
 * e.g. "down_blocks.0.resnets.0.conv1.weight".  `data` is a device pointer to a dense tensor of
 * `dtype` (GYRE_B200_F16/F32); the library packs its own copy (K-major fp16, tap-major convs) on
 * `stream`.  Replaces the parameter ownership of the nn.Module the reference clones to the GPU
 * (gyre/pipeline/model_utils.py:172-259); call again with the same key to re-pack after a
 * LoRA-style weight edit. */
int gyre_b200_load_weight(gyre_b200_handle h, const char* key, const void* data, int dtype,
                          const int64_t* shape, int ndim, gyre_b200_stream stream);
/* Verifies that every parameter the architecture needs has been loaded. */
int gyre_b200_finalize(gyre_b200_handle h);

/* Scratch bytes one forward needs for a CFG-doubled batch `batch`, latent height/width, context
 * length.  The caller allocates (torch.empty) and passes it to every forward. */
int gyre_b200_unet_workspace_bytes(gyre_b200_handle h, int batch, int height, int width, int ctx_len,
                                   size_t* bytes);

/* eps = unet(sample, timestep, encoder_hidden_states)
 *   sample  [batch, in_channels, height, width]  fp16 NCHW
 *   timestep[batch]                               int64 (the reference broadcasts scalars first)
 *   ctx     [batch, ctx_len, cross_attention_dim] fp16
 *   tome_r  NULL, or one int32 per transformer block in module execution order: the number of
 *           K/V tokens to merge (nonfree/tome_memory_efficient_cross_attention.py:28-50; the list
 *           parse_r builds, nonfree/ToMe/tome/utils.py:80-105) -- HOST pointer
 *   out     [batch, out_channels, height, width]  fp16 NCHW
 * ctx may be NULL after gyre_b200_unet_set_context bound a context of the same [batch, ctx_len]. */
int gyre_b200_unet_forward(gyre_b200_handle h, const void* sample, const int64_t* timestep, const void* ctx,
                           int batch, int height, int width, int ctx_len, const int32_t* tome_r_host,
                           void* out, void* workspace, size_t workspace_bytes, gyre_b200_stream stream);
/* Same with the additional conditioning vector of `addition_embed_type = "text_time"` models:
 * add_cond [batch, addition_embed_dim] fp16 = cat([text_embeds, sinusoid_256(time_ids).flatten(1)]). */
int gyre_b200_unet_forward_cond(gyre_b200_handle h, const void* sample, const int64_t* timestep, const void* ctx,
                                const void* add_cond, int batch, int height, int width, int ctx_len,
                                const int32_t* tome_r_host, void* out, void* workspace, size_t workspace_bytes,
                                gyre_b200_stream stream);

/* Binds the text context for the following forwards (the reference binds the embeddings once per
 * request: UNetWithEmbeddings, gyre/pipeline/unet/core.py:253-259).  The cross-attention K/V
 * projections depend only on ctx, so they are computed here once instead of once per denoising step;
 * later forwards pass ctx = NULL.  The library keeps the projected K/V (batch*ctx_len*2C fp16 per
 * transformer block) until the next call; ctx = NULL drops the binding.  The caller must re-bind
 * after changing the context tensor's contents or any attn2.to_k / to_v weight. */
int gyre_b200_unet_set_context(gyre_b200_handle h, const void* ctx, int batch, int ctx_len, gyre_b200_stream stream);
/* Classifier-free guidance as the reference runs it (CFGUNet_Parallel, gyre/pipeline/unet/cfg.py:47-57) feeds the UNet
 * `torch.cat([latents] * 2)` with the [uncond ; cond] text contexts: the two halves of the batch are the SAME samples
 * and timesteps until the first cross-attention reads the context.  on = 1 is the caller's promise that the following
 * forwards have that form (sample[b] == sample[b + batch / 2], timestep likewise): conv_in, the first resnet and the first
 * transformer's self-attention half are then computed once for both halves - the outputs are what the full computation
 * gives (bit for bit when the kernels involved are batch-invariant, i.e. with stream-K off; to the last fp16 bit
 * otherwise).  Sticky until changed; 0 (default) makes no assumption.  Tunable CFG_SHARE = 0 disables it globally. */
int gyre_b200_unet_set_cfg_duplicate(gyre_b200_handle h, int on);
/* ControlNet residual injection (replaces the `down_block_additional_residuals=` /
 * `mid_block_additional_residual=` keyword arguments that gyre/pipeline/unet/core.py:213-239 passes to
 * UNet2DConditionModel.forward).  Binds, for the NEXT gyre_b200_unet_forward* call only, one NCHW fp16 device tensor
 * per skip connection ([batch, C_i, h_i, w_i], in the order diffusers lists them: conv_in output, each down block's
 * layer outputs, each downsampler output - gyre_b200_unet_num_skips() of them) and/or one for the mid block output.
 * n_down = 0 and mid_residual = NULL clear the binding.  The tensors must stay valid until that forward has run. */
int gyre_b200_unet_set_control_residuals(gyre_b200_handle h, const void* const* down_residuals, int n_down,
                                         const void* mid_residual);
int gyre_b200_unet_num_skips(gyre_b200_handle h);
/* T2I-adapter states (the `adapter_states=` keyword the reference's patched UNet takes,
 * gyre/pipeline/t2i_adapter/unet_patcher.py:21-60,95-110; chosen per CFG half by unet/core.py:213-219): one NCHW fp16
 * device tensor per down block ([batch, C_i, h_i, w_i] at the block's resolution), added in place to the block's last
 * hidden state before its downsampler (so the block's last skip connection sees it too, as in the reference).
 * Valid for the NEXT forward only; n_states = 0 clears. */
int gyre_b200_unet_set_adapter_states(gyre_b200_handle h, const void* const* states, int n_states);

/* ------------------------------------------------------------------------------------------
 * AutoencoderKL  (replaces vae.decode(x).sample, gyre/pipeline/unified_pipeline.py:1523-1536,
 *                 and vae.encode(img).latent_dist, :305-318)
 * ------------------------------------------------------------------------------------------ */
typedef struct gyre_b200_vae_config {
  int32_t in_channels;             /* 3 */
  int32_t out_channels;            /* 3 */
  int32_t latent_channels;         /* 4 */
  int32_t num_levels;              /* 4 */
  int32_t block_out_channels[4];   /* (128,256,512,512)  gyre/ldm_config/v1-inference.yaml:46-67 */
  int32_t layers_per_block;        /* 2 */
  int32_t norm_num_groups;         /* 32 */
} gyre_b200_vae_config;

int gyre_b200_vae_create(const gyre_b200_vae_config* cfg, gyre_b200_handle* out);
int gyre_b200_vae_workspace_bytes(gyre_b200_handle h, int batch, int latent_h, int latent_w, size_t* bytes);
/* z [batch, 4, h, w] fp16 NCHW (already divided by the scaling factor, as the reference passes it)
 * -> img [batch, 3, 8h, 8w] fp16 NCHW.  postprocess != 0 additionally applies the pipeline tail
 * (img/2+0.5).clamp(0,1) (unified_pipeline.py:2491); img_u8 (optional, may be NULL) receives the
 * same image as [batch, 8h, 8w, 3] uint8 for the multi-GPU gather. */
int gyre_b200_vae_decode(gyre_b200_handle h, const void* z, int batch, int latent_h, int latent_w, int postprocess,
                         void* img, uint8_t* img_u8, void* workspace, size_t workspace_bytes,
                         gyre_b200_stream stream);
/* img [batch, 3, H, W] fp16 NCHW -> moments [batch, 8, H/8, W/8] fp16 NCHW (mean | logvar), the
 * parameters of the reference's DiagonalGaussianDistribution. */
int gyre_b200_vae_encode(gyre_b200_handle h, const void* img, int batch, int height, int width, void* moments,
                         void* workspace, size_t workspace_bytes, gyre_b200_stream stream);

/* ------------------------------------------------------------------------------------------
 * CLIP text encoder  (the step right before the sampling path: replaces
 *   transformers CLIPTextModel.forward as called through gyre's TextEncoderAltLayer,
 *   gyre/pipeline/text_embedding/text_encoder_alt_layer.py:6-36, from the LPW embedding code,
 *   gyre/pipeline/text_embedding/lpw_text_embedding.py:195-386)
 * Parameter keys are the transformers state-dict names ("text_model.embeddings.token_embedding.weight",
 * "text_model.encoder.layers.3.self_attn.q_proj.weight", ...), handed over with gyre_b200_load_weight.
 * ------------------------------------------------------------------------------------------ */
typedef struct gyre_b200_clip_config {
  int32_t vocab_size;              /* 49408                          */
  int32_t hidden_size;             /* 768 (CLIP-L) / 1024 (ViT-H)    */
  int32_t intermediate_size;       /* 3072 / 4096                    */
  int32_t num_layers;              /* 12 / 23-24                     */
  int32_t num_heads;               /* 12 / 16                        */
  int32_t max_positions;           /* 77                             */
  int32_t hidden_act;              /* 0 quick_gelu, 1 gelu (erf)     */
  float layer_norm_eps;            /* 1e-5                           */
} gyre_b200_clip_config;

int gyre_b200_clip_create(const gyre_b200_clip_config* cfg, gyre_b200_handle* out);
int gyre_b200_clip_workspace_bytes(gyre_b200_handle h, int batch, int seq_len, size_t* bytes);
/* input_ids [batch, seq_len] int64 -> out [batch, seq_len, hidden] fp16.
 * skip_last = 0: last_hidden_state ("final").  skip_last = k > 0: final_layer_norm(hidden_states[-(k+1)]),
 * i.e. TextEncoderAltLayer's "penultimate" (k = 1) / integer layer (k = layer - 1).
 * apply_final_ln = 0 returns the selected hidden state without the final LayerNorm. */
int gyre_b200_clip_forward(gyre_b200_handle h, const int64_t* input_ids, int batch, int seq_len, int skip_last,
                           int apply_final_ln, void* out, void* workspace, size_t workspace_bytes,
                           gyre_b200_stream stream);

/* ------------------------------------------------------------------------------------------
 * Safety checker  (replaces gyre/pipeline/safety_checkers.py:13-66 FlagOnlySafetyChecker.forward - and the vision half
 *   of diffusers' StableDiffusionSafetyChecker it is derived from - plus the CLIPFeatureExtractor call in front of it,
 *   gyre/pipeline/unified_pipeline.py:2512-2522)
 * Parameter keys are the checker's state-dict names: "vision_model.embeddings.class_embedding",
 * "vision_model.embeddings.patch_embedding.weight", "vision_model.embeddings.position_embedding.weight",
 * "vision_model.pre_layrnorm.*" (sic), "vision_model.encoder.layers.N.*", "vision_model.post_layernorm.*",
 * "visual_projection.weight", "concept_embeds", "special_care_embeds", "concept_embeds_weights",
 * "special_care_embeds_weights".
 * ------------------------------------------------------------------------------------------ */
typedef struct gyre_b200_clip_vision_config {
  int32_t image_size;              /* 224                            */
  int32_t patch_size;              /* 14                             */
  int32_t hidden_size;             /* 1024 (ViT-L/14)                */
  int32_t intermediate_size;       /* 4096                           */
  int32_t num_layers;              /* 24                             */
  int32_t num_heads;               /* 16                             */
  int32_t hidden_act;              /* 0 quick_gelu, 1 gelu (erf)     */
  float layer_norm_eps;            /* 1e-5                           */
  int32_t projection_dim;          /* 768; 0: no visual projection (tower only)  */
  int32_t num_concepts;            /* 17; 0: no concept embeddings (tower only)  */
  int32_t num_special;             /* 3                              */
} gyre_b200_clip_vision_config;
int gyre_b200_clip_vision_create(const gyre_b200_clip_vision_config* cfg, gyre_b200_handle* out);
int gyre_b200_clip_vision_workspace_bytes(gyre_b200_handle h, int batch, size_t* bytes);
/* pixel_values [batch, 3, S, S] fp16 -> image_embeds [batch, projection_dim] fp16 (nullable) and
 * scores [batch, num_special + num_concepts] fp32 = cosine similarity against the special-care embeddings, then the
 * concept embeddings (the thresholding is a handful of host scalars per image: gyre_b200.safety_checker). */
int gyre_b200_safety_scores(gyre_b200_handle h, const void* pixel_values, int batch, void* image_embeds, float* scores,
                            void* workspace, size_t workspace_bytes, gyre_b200_stream stream);
/* CLIPFeatureExtractor on the device.  One pass of PIL's 8-bit resample along the middle dimension of
 * src [n_outer, in_size, inner] -> dst [n_outer, out_size, inner] (bounds [out_size, 2] = first source index and tap
 * count, coeffs [out_size, ksize] 22-bit fixed point: built on the host with PIL's expressions), and the centre crop +
 * x 1/255 + (x - mean) / std of u8 NHWC [batch, H, W, 3] -> fp16 NCHW [batch, 3, S, S]. */
/* The vision tower alone (CLIPModel.vision_model as the style T2I-adapter uses it, gyre/pipeline/unified_pipeline.py:941-958):
 * hidden states [batch, tokens, hidden_size] fp16 after (num_layers - skip_last) encoder layers, before post_layernorm -
 * skip_last = 0 is `last_hidden_state`, 1 is `hidden_states[-2]` ...  A handle made with num_concepts = 0 and
 * projection_dim = 0 registers the tower's parameters only. */
int gyre_b200_clip_vision_hidden(gyre_b200_handle h, const void* pixel_values, int batch, int skip_last, void* hidden,
                                 void* workspace, size_t workspace_bytes, gyre_b200_stream stream);

/* Style T2I-adapter (replaces StyleAdapter.forward, gyre/pipeline/t2i_adapter/adapter.py:173-199; T2iAdapter_style,
 * models.py:146-160): x [batch, tokens, width] fp16 (CLIP vision hidden states) -> out [batch, num_token, context_dim] fp16,
 * the extra context tokens of the guided side.  Parameter keys are the module's, with two reshapes done by the caller:
 * "style_embedding" as [num_token, width] and "proj" TRANSPOSED to [context_dim, width]. */
typedef struct gyre_b200_style_adapter_config {
  int32_t width;        /* 1024 */
  int32_t context_dim;  /* 768  */
  int32_t num_head;     /* 8    */
  int32_t n_layers;     /* 3    */
  int32_t num_token;    /* 8    */
} gyre_b200_style_adapter_config;
int gyre_b200_style_adapter_create(const gyre_b200_style_adapter_config* cfg, gyre_b200_handle* out);
int gyre_b200_style_adapter_workspace_bytes(gyre_b200_handle h, int batch, int tokens, size_t* bytes);
int gyre_b200_style_adapter_forward(gyre_b200_handle h, const void* x, int batch, int tokens, void* out, void* workspace,
                                    size_t workspace_bytes, gyre_b200_stream stream);

int gyre_b200_resample_u8(const void* src, int64_t n_outer, int in_size, int inner, const int32_t* bounds,
                          const int32_t* coeffs, int ksize, int out_size, void* dst, gyre_b200_stream stream);
int gyre_b200_clip_normalize(const void* src_u8_nhwc, int batch, int height, int width, int crop, const float* mean3,
                             const float* std3, void* out, gyre_b200_stream stream);

/* ------------------------------------------------------------------------------------------
 * PNG encoding  (replaces gyre/images.py:93-111 toPngBytes -> torchvision.io.encode_png per image on the host, called by
 *   gyre/services/generate.py:79 image_to_artifact)
 * images u8 NHWC [batch, height, width, channels] on the device (channels 1 grey, 2 grey + alpha, 3 RGB, 4 RGBA; 8 bits) ->
 * out [batch][out_stride] bytes, one complete PNG file per image, its length in out_lengths[i].  gyre_b200_png_sizes gives
 * the workspace size and the smallest out_stride (worst case: incompressible input).  Lossless: any PNG decoder returns
 * the input pixels; the byte stream is this library's own (chunk-parallel filters + Huffman coding), not libpng's.
 * Scanlines longer than 32768 bytes are refused.
 * ------------------------------------------------------------------------------------------ */
int gyre_b200_png_sizes(int batch, int height, int width, int channels, size_t* workspace_bytes, size_t* out_stride);
int gyre_b200_png_encode(const void* images_u8_nhwc, int batch, int height, int width, int channels, void* out,
                         size_t out_stride, int64_t* out_lengths, void* workspace, size_t workspace_bytes,
                         gyre_b200_stream stream);

/* Lossless WebP  (replaces gyre/images.py:125-135 toWebpBytes -> cv.imencode(".webp", .., [IMWRITE_WEBP_QUALITY, 500]) on the
 *   host, chosen by gyre/services/generate.py:73-76 for clients that accept image/webp)
 * images u8 NHWC [batch, height, width, channels] (3 RGB | 4 RGBA) -> out [batch][out_stride] bytes, one simple-format VP8L
 * file per image ("RIFF" .. "WEBP" "VP8L"), its length in out_lengths[i]; out must be 4-byte aligned, out_stride a multiple
 * of 4 and at least what gyre_b200_webp_sizes returns.  Lossless: any WebP decoder returns the input pixels; the stream is
 * this library's own (gradient predictor + per-image Huffman codes, no LZ77). */
int gyre_b200_webp_sizes(int batch, int height, int width, int channels, size_t* workspace_bytes, size_t* out_stride);
int gyre_b200_webp_encode(const void* images_u8_nhwc, int batch, int height, int width, int channels, void* out,
                          size_t out_stride, int64_t* out_lengths, void* workspace, size_t workspace_bytes,
                          gyre_b200_stream stream);

int gyre_b200_destroy(gyre_b200_handle h);

/* ------------------------------------------------------------------------------------------
 * Scheduler inner loop  (replaces the per-step elementwise chain of
 *   k_diffusion/external.py:96-113,149-167 (denoiser scalings), gyre/pipeline/unet/cfg.py:47-57 (CFG),
 *   k_diffusion/sampling.py:46-58,139-155 (to_d / Euler / Euler-ancestral update) and
 *   gyre/pipeline/schedulers/scheduling_ddim.py:259-316 (DDIM step) -- one launch per step)
 * ------------------------------------------------------------------------------------------ */
typedef struct gyre_b200_step {
  int32_t kind;        /* 0: k-diffusion Euler family   1: DDIM                                    */
  int32_t v_pred;      /* model predicts v (SD2.1-768)                                              */
  int32_t cfg;         /* model_out holds [uncond ; cond] halves (CFGUNet_Parallel)                 */
  float guidance;      /* CFG scale                                                                 */
  float sigma;         /* k: sigma_i (after fp16 quantisation, common_scheduler.py:560)             */
  float c_in_next;     /* k: 1/sqrt(sigma_{i+1}^2+1) -- scale of the NEXT unet input; 0: skip       */
  float dt;            /* k: sigma_down - sigma_i                                                   */
  float sigma_up;      /* k: ancestral noise scale (0: none)                                        */
  float sqrt_a_t, sqrt_1m_a_t, sqrt_a_prev, dir_coef, noise_coef; /* DDIM coefficients             */
} gyre_b200_step;

/* x, x_out, denoised_out: fp32 [B, C, h, w]; model_out: fp16 [2B or B, C, h, w]; noise fp32 or NULL;
 * x_in_next: fp16 [2B or B, ...] = x_out * c_in_next (duplicated for CFG) or NULL. */
int gyre_b200_sched_step(const gyre_b200_step* s, const float* x, const void* model_out, const float* noise,
                         float* x_out, float* denoised_out, void* x_in_next, int batch, int64_t per_sample,
                         gyre_b200_stream stream);
/* CFG combine on its own, for callers that keep the reference's un-fused wrapper stack
 * (gyre/pipeline/unet/cfg.py:54-57): out = u + guidance * (g - u), model_out fp16 [2B, ...] =
 * [uncond ; cond]; either output may be NULL. */
int gyre_b200_cfg_combine(const void* model_out, float guidance, int batch, int64_t per_sample, void* out_f16,
                          float* out_f32, gyre_b200_stream stream);
/* Building blocks of the multi-evaluation samplers (k_diffusion/sampling.py:159-278,509-581 Heun, DPM-2,
 * DPM-2 ancestral, LMS, DPM++ 2S ancestral, DPM++ SDE; gyre/pipeline/schedulers/sample_dpmpp_2m.py:6-50):
 * every update there is a denoiser call followed by a linear combination of latent-sized tensors whose
 * scalar coefficients the host computes exactly as the reference does.
 *   denoise: denoised = x * c_skip + cfg(model_out) * c_out   (external.py:96-113: eps c_skip=1, c_out=-sigma;
 *            :149-167: v c_skip=1/(sigma^2+1), c_out=-sigma/sqrt(sigma^2+1)); cfg != 0: model_out = [uncond ; cond]
 *   lincomb: out = sum_k coefs_host[k] * inputs_host[k]  (n_terms <= 6 device pointers in a HOST array, fp32);
 *            x_in_next (optional) = fp16(out * c_in), duplicated when dup != 0; out may alias an input or be NULL */
int gyre_b200_denoise(const float* x, const void* model_out, int cfg, float guidance, float c_skip, float c_out,
                      int batch, int64_t per_sample, float* denoised, gyre_b200_stream stream);
int gyre_b200_lincomb(int n_terms, const float* const* inputs_host, const float* coefs_host, int batch,
                      int64_t per_sample, float* out, void* x_in_next, float c_in, int dup, gyre_b200_stream stream);
/* Error estimate of the adaptive DPM-Solver (`sample_dpm_adaptive`, k_diffusion/sampling.py:461-462):
 *   delta = max(atol, rtol * max(|x_low|, |x_prev|));  error = ||(x_low - x_high) / delta||_2 / sqrt(n).
 * Writes gyre_b200_dpm_error_num_partials() fp64 block sums of ((x_low - x_high) / delta)^2 to `partials` (device);
 * the caller adds them in index order (reproducible accept / reject decisions) and takes sqrt(sum / n). */
int gyre_b200_dpm_error_partials(const float* x_low, const float* x_high, const float* x_prev, float atol, float rtol,
                                 int64_t n, double* partials, gyre_b200_stream stream);
int gyre_b200_dpm_error_num_partials(void);
/* Legacy (4-channel UNet) inpainting: EnhancedInpaintMode.wrap_k_unet / _blend
 * (gyre/pipeline/unified_pipeline.py:620-636) replaces the predicted x0 by the original image latents
 * wherever blend_mask > u (u = progress in [0, 1), common_scheduler.py:358-389).  Same as the functions
 * above with that substitution applied to `denoised` before the update. */
int gyre_b200_denoise_blend(const float* x, const void* model_out, int cfg, float guidance, float c_skip, float c_out,
                            int batch, int64_t per_sample, float* denoised, const float* blend_orig,
                            const float* blend_mask, float blend_u, gyre_b200_stream stream);
int gyre_b200_sched_step_blend(const gyre_b200_step* s, const float* x, const void* model_out, const float* noise,
                               float* x_out, float* denoised_out, void* x_in_next, int batch, int64_t per_sample,
                               const float* blend_orig, const float* blend_mask, float blend_u,
                               gyre_b200_stream stream);
/* UNet input of the inpaint (9-channel) / depth (5-channel) models: out[b] = cat([x[b], extra[b % extra_batch]], dim=1),
 * NCHW fp16 (EnhancedRunwayInpaintMode.wrap_unet, unified_pipeline.py:668-690; UnetWithExtraChannels,
 * gyre/pipeline/unet/core.py:21-37).  The extra channels are not scaled by c_in. */
int gyre_b200_cat_channels(const void* x, int channels, const void* extra, int extra_channels, int extra_batch, int batch,
                           int64_t hw, void* out, gyre_b200_stream stream);
/* out_f16[(dup?2:1) * B, ...] = x * c_in  (first unet input of a run) */
int gyre_b200_scale_latents(const float* x, float c_in, int dup, int batch, int64_t per_sample, void* out,
                            gyre_b200_stream stream);

/* ------------------------------------------------------------------------------------------
 * T2I-adapter encoder  (replaces `Adapter.forward`, gyre/pipeline/t2i_adapter/adapter.py:102-132; defaults of
 * T2iAdapter_main, t2i_adapter/models.py:80-88).  Parameter keys are the nn.Module state-dict names
 * ("conv_in.weight", "body.3.block1.bias", "body.2.down_opt.op.weight", ...).
 * image [B, cin / 64, H, W] NCHW fp16 (H, W multiples of 8) -> num_levels feature maps, NCHW fp16
 * [B, channels[i], H / 8 / 2^i, W / 8 / 2^i] (device pointers in the HOST array `features`) - the `adapter_states` of
 * gyre_b200_unet_set_adapter_states.
 * ------------------------------------------------------------------------------------------ */
typedef struct gyre_b200_adapter_config {
  int32_t cin;            /* channels after PixelUnshuffle(8): 64 x image channels (192 for RGB hints) */
  int32_t num_levels;     /* <= 4 */
  int32_t channels[4];    /* (320, 640, 1280, 1280) */
  int32_t nums_rb;        /* ResnetBlocks per level (2) */
  int32_t ksize;          /* 1 | 3: kernel of in_conv / block2 / skep (1) */
  int32_t sk;             /* 1: identity skip, in_conv only where the width changes (1) */
  int32_t use_conv;       /* 0: AvgPool2d(2) downsample, 1: stride-2 conv3x3 (0) */
  int32_t light;          /* 1: Adapter_light (adapter.py:240-263): per level [AvgPool2d(2)] -> 1x1 in_conv to channels / 4 ->
                             nums_rb x (conv3x3, ReLU, conv3x3, + x) -> 1x1 out_conv; no conv_in; ksize / sk / use_conv unused */
} gyre_b200_adapter_config;
int gyre_b200_adapter_create(const gyre_b200_adapter_config* cfg, gyre_b200_handle* out);
int gyre_b200_adapter_workspace_bytes(gyre_b200_handle h, int batch, int height, int width, size_t* bytes);
int gyre_b200_adapter_forward(gyre_b200_handle h, const void* image, int batch, int height, int width, void* const* features,
                              int n_features, void* workspace, size_t workspace_bytes, gyre_b200_stream stream);

/* Image-space tail of an outpaint request (gyre/pipeline/unified_pipeline.py:2493-2510 with gyre/images.py:667-672 and
 * gyre/match_histograms.py:12-37 - done on the host with numpy in the reference): result / source / outmask / out are
 * [batch, 3, hw] fp16 images in [0, 1]; out = source * (1 - outmask) + match_histograms(result, source * (1 - outmask) +
 * result * outmask) * outmask, the match being the per-channel uint8 CDF match over the whole batch.  Bit-identical to
 * the reference's fp16 evaluation.  scratch: gyre_b200_outpaint_scratch_bytes() bytes. */
size_t gyre_b200_outpaint_scratch_bytes(void);
int gyre_b200_outpaint_match_histograms(const void* result, const void* source, const void* outmask, int batch, int64_t hw,
                                        void* out, void* scratch, gyre_b200_stream stream);

/* Prompt weighting of the LPW text embedding (gyre/pipeline/text_embedding/lpw_text_embedding.py:352-371):
 * emb [batch, tokens, channels] fp16, weights [batch, tokens] fp32 ->
 * out = emb * w * (mean(emb) / mean(emb * w)) per prompt (the weighted embedding keeps its mean). */
int gyre_b200_lpw_weight(const void* emb, const float* weights, int batch, int tokens, int channels, void* out,
                         gyre_b200_stream stream);

/* Hires-fix / graft blending of the scheduler-UNet wrappers (replaces the ~20 torch ops per step of
 * gyre/pipeline/unet/hires_fix.py:142-205 HiresUnetWrapper.__call__ and gyre/pipeline/unet/graft.py:31-48):
 *   resample_select: `scale_into` (hires_fix.py:45-92) - a separable 4-tap lanczos2 resample of src [planes, src_h, src_w]
 *   fp32 (tap tables [resized, 4] built on the host with ResizeRight's expressions, resize_right.py:70-118,203-213; NULL
 *   tables = that dimension is not resampled), placed at (off_y, off_x) in a [target_h, target_w] frame (negative
 *   offsets crop the centre), outside it replicate padding (mode 0, strategy "pad") or `background` (mode 1, strategy
 *   "clone") - then, when `other` is given, where(rand_map >= p, A, B) with (A, B) = (resampled, other) if
 *   resampled_if_ge else (other, resampled) - written at (frame_y, frame_x) into out [planes, frame_h, frame_w], zero
 *   elsewhere (`lo_expanded`, hires_fix.py:196-198).
 *   rand_select: out = where(rand_map >= p, a, b)  (graft.py:44-46). */
int gyre_b200_resample_select(const float* src, int planes, int src_h, int src_w, const int32_t* taps_y_idx,
                              const float* taps_y_w, int resized_h, const int32_t* taps_x_idx, const float* taps_x_w,
                              int resized_w, int target_h, int target_w, int off_y, int off_x, int mode,
                              const float* background, const float* other, const float* rand_map, float p,
                              int resampled_if_ge, float* out, int frame_h, int frame_w, int frame_y, int frame_x,
                              gyre_b200_stream stream);
/* One pass of `images.resize` (gyre/images.py:324-340: ResizeRight, lanczos3, reflect padding, antialiasing for sharpness 1;
 * used for hint masks, unified_pipeline.py:790-808, and the depth hint, :2007-2008): src [n_outer, in_size, inner] fp32 ->
 * dst [n_outer, out_size, inner] along the middle dimension with host-built tables idx / weights [out_size, ksize] (the
 * reflect padding is folded into idx; gyre_b200.images.resize builds them with ResizeRight's expressions). */
int gyre_b200_resample_f32(const float* src, int64_t n_outer, int in_size, int inner, const int32_t* idx, const float* weights,
                           int ksize, int out_size, int clamp01, float* dst, gyre_b200_stream stream);
int gyre_b200_rand_select(const float* a, const float* b, const float* rand_map, float p, int64_t n, float* out,
                          gyre_b200_stream stream);

/* ------------------------------------------------------------------------------------------
 * Attention patcher  (replaces ToMeMemoryEfficientCrossAttention.forward's merge,
 *   nonfree/tome_memory_efficient_cross_attention.py:28-50, i.e. tome/merge.py:18-97,210-224)
 * k, v [B, N, C] fp16 -> k_out, v_out [B, N - r, C] fp16 with the reference's token order
 * ([unmerged even tokens in descending-score order ..., odd tokens ...]).
 * ------------------------------------------------------------------------------------------ */
int gyre_b200_tome_workspace_bytes(int batch, int tokens, int channels, size_t* bytes);
/* The merge plan (merge.py:41-64) stays in the workspace after gyre_b200_tome_merge_kv; byte offsets of its int32 arrays,
 * Na = ceil(tokens / 2) entries per sample each:
 *   node_idx [batch, Na]  every even ("a") token's best odd ("b") partner  (merge.py: node_idx)
 *   unm_idx  [batch, Na]  first Na - r entries: the a-tokens that stay, in descending-score order (merge.py: unm_idx)
 *   src_idx  [batch, Na]  first r entries: the merged a-tokens, ordered by (destination, token) - the summation order;
 *                         as a set they are merge.py's src_idx, their destinations are node_idx[src]
 * Tie rule: equal best scores are ordered by ascending token index; equal scores of one a-token pick the lowest b. */
int gyre_b200_tome_plan_offsets(int batch, int tokens, int channels, size_t* node_idx, size_t* unm_idx, size_t* src_idx);
int gyre_b200_tome_merge_kv(const void* k, const void* v, int batch, int tokens, int channels, int r, void* k_out,
                            void* v_out, void* workspace, size_t workspace_bytes, gyre_b200_stream stream);

/* ------------------------------------------------------------------------------------------
 * Building-block kernels, exported for the per-kernel parity tests and for callers that patch a
 * single module (the reference's attention patcher swaps one module's forward:
 * nonfree/tome_patcher.py:14-52).  Token-major fp16 tensors ([rows, C], C contiguous).
 * ------------------------------------------------------------------------------------------ */
typedef struct gyre_b200_epilogue {
  const float* bias;          /* [N] fp32 or NULL                                                    */
  const void* rowgroup_bias;  /* fp16 [rows / rows_per_group, rgb_ld] added per row group or NULL    */
  int32_t rows_per_group;
  int32_t rgb_ld;
  const void* residual;       /* fp16 [rows, ldr] or NULL                                            */
  int32_t ldr;
  int32_t act;                /* 0 none, 1 GEGLU (value*gelu(gate); W holds [value rows ; gate rows]), 2 SiLU */
  int32_t out_f32;            /* 1: out is fp32                                                      */
  void* out;                  /* [rows, ldo]                                                         */
  int32_t ldo;
  /* optional stream-K scratch (NULL: whole tiles only).  When the last wave of 128 x BN tiles would leave most
   * SMs idle, its tiles are cut along K across ALL SMs; partial fp32 accumulators travel through sk_ws and are
   * summed in a fixed order (deterministic).  sk_flags: >= 148 int32, zero before the first use (the kernel
   * leaves them zero); both may be shared by every launch on one stream. */
  void* sk_ws;
  size_t sk_ws_bytes;
  int32_t* sk_flags;
  int32_t sk_flags_count;
  /* LayerNorm folded into the GEMMs around it (gyre_b200_gemm only; all NULL: off).  Replaces the F.layer_norm between
   * two Linears of a BasicTransformerBlock (diffusers attention.py; structure restated in nonfree/tome_unet.py:114-136):
   *   rowstat_out [gyre_b200_gemm_rowstat_parts(M, N)][M] float2 - the producing GEMM also leaves, per output row, the
   *     (sum, sum of squares) of every row segment it writes; gyre_b200_ln_finalize_rows folds them to (mean, rstd);
   *   ln_rowstat [M] float2 (mean, rstd) + ln_colsum [N] - the consuming GEMM takes the RAW rows as A and weights
   *     prepared by gyre_b200_ln_fold_linear (W' = gamma (.) W, colsum, bias' = bias + beta @ W^T) and stores
   *     rstd * (acc - mean * colsum[n]) + bias'[n]  ==  LayerNorm(x) @ W^T + bias  (act: none or GEGLU). */
  void* rowstat_out;
  const void* ln_rowstat;
  const float* ln_colsum;
  /* ln_parts in 1..8: ln_rowstat holds the producer's raw partials [ln_parts][M] and the consumer folds them itself
   * with ln_inv_c = 1 / C and ln_eps (no finalize launch); 0: ln_rowstat holds finished (mean, rstd) pairs. */
  int32_t ln_parts;
  float ln_inv_c;
  float ln_eps;
  /* GroupNorm statistics of the output, produced by the convolution that writes it (gyre_b200_conv3x3 only; NULL: off).
   * Replaces the statistics half of the torch.nn.GroupNorm that follows a conv in diffusers' ResnetBlock2D / AttentionBlock
   * (structure restated in SURVEY.md A.2): gn_out [B][gn_nparts][gn_groups][2] float = per output tile of a sample the
   * (sum, sum of squares) of every group's fp16-rounded outputs, summed in a fixed order; gyre_b200_groupnorm_pre
   * consumes them.  gn_nparts must be gyre_b200_conv3x3_gn_parts() of the launch (0: this shape cannot produce them). */
  float* gn_out;
  int32_t gn_groups;
  int32_t gn_nparts;
} gyre_b200_epilogue;

/* out = epilogue([A | A2] @ W^T): A [M, K1] pitch lda, A2 [M, K2] pitch lda2 (may be NULL/0),
 * W [N, K1+K2] pitch ldw; fp16, fp32 accumulate (replaces F.linear / 1x1 conv: cuBLAS HGEMM). */
int gyre_b200_gemm(const void* A, int lda, int K1, const void* A2, int lda2, int K2, const void* W, int ldw, int M,
                   int N, const gyre_b200_epilogue* ep, gyre_b200_stream stream);
int gyre_b200_gemm_rowstat_parts(int M, int N);
int gyre_b200_ln_finalize_rows(const void* parts, int nparts, int M, int C, float eps, void* mean_rstd,
                               gyre_b200_stream stream);
/* W [N, K] fp16 (any row order, e.g. GEGLU-packed) -> W' [N, K] fp16, colsum [N], lnbias [N] (bias may be NULL) */
int gyre_b200_ln_fold_linear(const void* W, int N, int K, const float* gamma, const float* beta, const float* bias,
                             void* W_out, float* colsum, float* lnbias, gyre_b200_stream stream);
/* GEGLU weights must be re-ordered so that each 256-row tile holds 128 value rows then their 128
 * gate rows: packs W [2*F, K] (diffusers ff.net.0.proj layout) into Wp [2*F, K]. F % 128 == 0. */
int gyre_b200_pack_geglu(const void* W, int dtype, int F, int K, const void* bias, int bias_dtype, void* Wp,
                         float* bias_p, gyre_b200_stream stream);
/* 3x3 convolution, implicit GEMM: X [B, H, W, Cin] NHWC fp16 (pitch ldx), Wp packed by
 * gyre_b200_pack_conv3x3 -> out rows = output pixels (replaces F.conv2d: cuDNN). */
int gyre_b200_conv3x3(const void* X, int ldx, int B, int H, int W, int Cin, const void* Wp, int Cout, int stride,
                      int pad, const gyre_b200_epilogue* ep, gyre_b200_stream stream);
/* W [Cout, Cin, 3, 3] (dtype f16/f32) -> Wp fp16 [Cout, 9, round_up(Cin, 64)], tap = kh*3+kw. */
size_t gyre_b200_conv3x3_packed_elems(int Cin, int Cout);
int gyre_b200_pack_conv3x3(const void* W, int dtype, int Cin, int Cout, void* Wp, gyre_b200_stream stream);
/* conv3x3(nearest-2x upsample(X)) without writing the upsampled tensor (replaces diffusers Upsample2D:
 * F.interpolate(scale_factor=2, mode="nearest") + conv): X [B, H, W, Cin] NHWC fp16 ->
 * out [B, 2H, 2W, Cout].  Wp4 = the four 2x2 phase kernels packed by gyre_b200_pack_upconv3x3
 * (fp16 [4, Cout, 4, round_up(Cin, 64)]); epilogue: bias only. */
size_t gyre_b200_upconv3x3_packed_elems(int Cin, int Cout);
int gyre_b200_pack_upconv3x3(const void* W, int dtype, int Cin, int Cout, void* Wp4, gyre_b200_stream stream);
int gyre_b200_upconv2x(const void* X, int ldx, int B, int H, int W, int Cin, const void* Wp4, int Cout,
                       const gyre_b200_epilogue* ep, gyre_b200_stream stream);
/* GroupNorm (+SiLU) over NHWC fp16; the input may be the channel concatenation x1 ++ x2
 * (replaces at::native_group_norm + silu).  scratch: fp32, gyre_b200_groupnorm_scratch_floats. */
size_t gyre_b200_groupnorm_scratch_floats(int B, int HW, int G);
int gyre_b200_groupnorm(const void* x1, int C1, const void* x2, int C2, int B, int HW, int G, float eps,
                        const float* gamma, const float* beta, int silu, void* out, float* scratch,
                        gyre_b200_stream stream);
/* The same from statistics a producing convolution left (gyre_b200_epilogue::gn_out): one pass over x.  pre
 * [B][nparts][G][2] float; stats: fp32 scratch of >= 2 * B * G floats.  gyre_b200_groupnorm_pre_ok tells whether the
 * shape takes this path at all (small maps are normalised by a single-pass kernel that needs no statistics). */
int gyre_b200_conv3x3_gn_parts(int B, int H, int W, int Cout, int stride, int pad, int groups);
int gyre_b200_groupnorm_pre_ok(int C, int HW, int G);
int gyre_b200_groupnorm_pre(const void* x, int C, int B, int HW, int G, float eps, const float* gamma, const float* beta,
                            int silu, void* out, const float* pre, int nparts, float* stats, gyre_b200_stream stream);
int gyre_b200_layernorm(const void* x, int rows, int C, float eps, const float* gamma, const float* beta, void* out,
                        gyre_b200_stream stream);
/* softmax(Q K^T * scale) V per head, reading Q/K/V in place from token-major projections:
 * q [B, Nq, ldq] (head h at columns h*d), k / v [B, Nk, ldk / ldv]; out [B, Nq, heads*d]
 * (replaces xformers.ops.memory_efficient_attention,
 *  gyre/pipeline/models/memory_efficient_cross_attention.py:39-60). */
int gyre_b200_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int B, int heads,
                        int Nq, int Nk, int d, float scale, void* out, int ldo, gyre_b200_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* GYRE_B200_H */
