// Internal op layer of libgyre_b200: every function enqueues sm_100a kernels on the given stream and
// returns 0 or a negative status (message via gyre::set_last_error).  Activations are NHWC fp16
// ("token-major": [B, H*W, C]); weights are pre-packed K-major fp16.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gyre {

enum OutMode { OUT_F16 = 0, OUT_F32 = 1 };
enum Act { ACT_NONE = 0, ACT_GEGLU = 1, ACT_SILU = 2, ACT_ROWMAX = 3, ACT_QUICKGELU = 4, ACT_GELU = 5, ACT_RELU = 6 };

// Epilogue description shared by the GEMM and the implicit-GEMM conv.
struct Epilogue {
  const float* bias = nullptr;            // [N] fp32, indexed by accumulator column
  const __half* rowgroup_bias = nullptr;  // [groups, rgb_ld] added per (row / rows_per_group), e.g. temb projection
  int rows_per_group = 1;
  int rgb_ld = 0;
  const __half* residual = nullptr;       // [rows, ldr] fp16 added after bias/activation
  int ldr = 0;
  void* out = nullptr;                    // OUT_F16 / OUT_F32 destination
  int ldo = 0;
  int out_mode = OUT_F16;
  int act = ACT_NONE;
  // ACT_ROWMAX (ToMe scoring): nothing is stored but, per row, the max and argmax over each (n-tile, column-half)
  // of the product: rowmax_val / rowmax_idx [rows, rowmax_ld], entry n_tile * 2 + half
  float* rowmax_val = nullptr;
  int* rowmax_idx = nullptr;
  int rowmax_ld = 0;
  // LayerNorm folded into the GEMMs around it (GEMM only, TMA epilogue):
  //   producer side - rowstat_out [2 * n_tiles][M] float2: per output row the (sum, sum of squares) of the row segment
  //   each (n-tile, column-half) of this launch produced (after bias / residual); ln_finalize_rows folds them into
  //   (mean, rstd) per row.
  //   consumer side - A holds the RAW rows x, W the gamma-scaled weights W' = gamma (.) W; with ln_rowstat [M] float2
  //   (mean, rstd) and ln_colsum [N] = sum_k W'[n, k] the epilogue turns the accumulator into LayerNorm(x) @ W^T:
  //   rstd * (acc - mean * colsum[n]) + bias[n], bias already holding beta @ W^T.
  //   With ln_parts in 1..8 the consumer folds the producer's partials itself (ln_rowstat = [ln_parts][M] raw partials,
  //   ln_inv_c = 1 / C, ln_eps) and no finalize launch is needed.
  float2* rowstat_out = nullptr;
  int rowstat_ld = 0;                     // rows between two partials of rowstat_out (0: M) - producers that write a row range
  const float2* ln_rowstat = nullptr;
  const float* ln_colsum = nullptr;
  int ln_parts = 0;
  float ln_inv_c = 0.f;
  float ln_eps = 1e-5f;
  // GroupNorm statistics of the OUTPUT, produced by the 3x3 convolution that writes it (conv only, TMA epilogue):
  // gn_out [B][gn_nparts][gn_groups][2] float = per output tile of a sample the (sum, sum of squares) of every group's
  // fp16-rounded outputs, summed in a fixed order (batch-size independent); groupnorm_nhwc_pre consumes them in place of
  // its own statistics pass.  gn_nparts must be what conv3x3_gn_parts() returns for the launch (0 = cannot be fused).
  float* gn_out = nullptr;
  int gn_groups = 0;
  int gn_nparts = 0;
  // optional stream-K scratch (owned by the caller, reused by every launch on one stream): fp32 partial
  // accumulators + one int flag per tile of the partial wave (flags must be zero; the kernel leaves them zero)
  void* sk_ws = nullptr;
  size_t sk_ws_bytes = 0;
  int* sk_flags = nullptr;
  int sk_flags_count = 0;
};

// out[M, N] = epilogue(A[M, K] @ W[N, K]^T).  A: fp16 row pitch lda; W: packed fp16 [N, Kp] row pitch ldw.
// For ACT_GEGLU the packed W interleaves value/gate rows per 256-column tile and N counts both.
int gemm_f16(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const Epilogue& ep,
             cudaStream_t st);
// Same with the K dimension split over two sources: out = [A | A2] @ W^T  (UNet skip-concat feeding a 1x1
// shortcut without materialising the concatenation).  K1 must be a multiple of 64.
int gemm2_f16(const __half* A, int lda, int K1, const __half* A2, int lda2, int K2, const __half* W, int ldw, int M,
              int N, const Epilogue& ep, cudaStream_t st);
// Partials per sample a stride-1/2 conv3x3 of this shape leaves in Epilogue::gn_out, or 0 when the statistics cannot be
// produced there (tiles that span samples or overhang the image, group boundaries that do not fall on tile boundaries).
int conv3x3_gn_parts(int B, int H, int W, int Cout, int stride, int pad, int groups);
// output extent of conv3x3_f16 along one axis (stride 1: n; stride 2: pad 1 -> (n - 1) / 2 + 1, pad 0 = F.pad(0, 1) -> (n - 2) / 2 + 1)
int conv3x3_out_extent(int n, int stride, int pad);
// number of (max, argmax) partials per row an ACT_ROWMAX GEMM of this shape writes
int gemm_rowmax_partials(int M, int N);
// number of float2 partials per row a GEMM of this shape writes to Epilogue::rowstat_out
int gemm_rowstat_parts(int M, int N);
// (mean, rstd) per row from the row-segment partials of a producing GEMM: parts [nparts][M] -> out [M]
int ln_finalize_rows(const float2* parts, int nparts, int M, int C, float eps, float2* out, cudaStream_t st);
// LayerNorm affine folded into the following Linear: w_out[n, k] = fp16(w[n, k] * gamma[k]), colsum[n] = sum_k w_out[n, k],
// lnbias[n] = bias[n] (0 if null) + sum_k beta[k] * w[n, k].  w / w_out: [N, K] fp16 row pitch K (any row order).
int ln_fold_linear(const __half* w, int N, int K, const float* gamma, const float* beta, const float* bias, __half* w_out,
                   float* colsum, float* lnbias, cudaStream_t st);

// 3x3 convolution as implicit GEMM.  X: [B, H, W, Cin] fp16 NHWC with channel pitch ldx; Wp: packed
// [Cout, 9*cin_pad] (tap-major, cin_pad = round_up(Cin, 64)); output rows are pixels of [B, Ho, Wo].
// stride 1|2; pad 1 (symmetric) or 0 (diffusers Downsample2D padding=0: zero pad right/bottom only).
int conv3x3_f16(const __half* X, int ldx, int B, int H, int W, int Cin, const __half* Wp, int Cout, int stride,
                int pad, const Epilogue& ep, cudaStream_t st);

// conv3x3(nearest_upsample_2x(X)) without materialising the upsampled tensor: four 2x2 phase convolutions on
// the low-res X [B, H, W, Cin] with weights packed by pack_upconv3x3; output [B, 2H, 2W, Cout] (ep.out, pitch
// ep.ldo, fp16, 16-byte aligned rows).  Epilogue: bias only.
int upconv2x_f16(const __half* X, int ldx, int B, int H, int W, int Cin, const __half* Wp4, int Cout,
                 const Epilogue& ep, cudaStream_t st);

// GroupNorm (+optional SiLU) over NHWC fp16.  Input may be the channel-concatenation of two tensors
// (x1 [.., C1] ++ x2 [.., C2]); output is one dense [B, HW, C1+C2] tensor.  `partials` is fp32 scratch
// of at least gn_partials_floats(B, HW, G) floats.
size_t gn_partials_floats(int B, int HW, int G);
int groupnorm_nhwc(const __half* x1, int C1, const __half* x2, int C2, int B, int HW, int G, float eps,
                   const float* gamma, const float* beta, bool silu, __half* out, float* partials, cudaStream_t st);

// The same with the statistics already there: pre [B][nparts][G][2] float (sum, sum of squares) partials left by the
// producing convolution (Epilogue::gn_out).  `stats` is fp32 scratch of >= 2 * B * G floats (used when nparts > 64: a
// tiny launch folds the partials into (mean, rstd) first).  Whether a given x can take this path: groupnorm_pre_ok().
int groupnorm_nhwc_pre(const __half* x, int C, int B, int HW, int G, float eps, const float* gamma, const float* beta,
                       bool silu, __half* out, const float* pre, int nparts, float* stats, cudaStream_t st);
// true when groupnorm_nhwc would run its two-pass kernels for this shape (the single-pass small-map kernel reads the
// tensor once anyway and does not take precomputed statistics)
bool groupnorm_pre_ok(int C, int HW, int G);

// LayerNorm over the last dim of [rows, C] fp16 (fp32 statistics, affine).
int layernorm_rows(const __half* x, int rows, int C, float eps, const float* gamma, const float* beta, __half* out,
                   cudaStream_t st);

// Flash attention reading Q/K/V in place from token-major projections: q [B, Nq, ldq] (head h at columns
// h*d), k/v [B, Nk, ldk/ldv]; writes out [B, Nq, ldo] (head h at columns h*d).  d % 8 == 0, d <= 192.
void attention_set_trace(long long* dev_buf, int capacity);
int attention_f16(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv, int B, int heads,
                  int Nq, int Nk, int d, float scale, __half* out, int ldo, cudaStream_t st);

// Text-encoder kernels (CLIP): token + position embedding gather, and causal self-attention over short sequences
// (L <= 128) read in place from the fused q|k|v projection [B, L, 3C].
int embed_tokens(const int64_t* ids, const __half* tok_emb, const __half* pos_emb, int B, int L, int C, int vocab,
                 __half* out, cudaStream_t st);
int causal_attention_short(const __half* qkv, int B, int L, int heads, int d, float scale, __half* out, cudaStream_t st);

// Row softmax on fp32 scores [rows, n] -> fp16 probs [rows, ldp] (VAE single-head attention).
int softmax_rows_f32(const float* s, int rows, int n, float scale, __half* p, int ldp, cudaStream_t st);

// Layout / elementwise helpers
int nchw_to_nhwc_f16(const __half* x, int B, int C, int H, int W, __half* out, int ldo, cudaStream_t st);
int add_nchw_to_nhwc_f16(__half* dst_nhwc, const __half* src_nchw, int B, int C, int HW, cudaStream_t st);
int nhwc_to_nchw_f16(const __half* x, int ldx, int B, int C, int H, int W, __half* out, cudaStream_t st);
int upsample2x_nhwc(const __half* x, int B, int H, int W, int C, __half* out, cudaStream_t st);
int concat_channels(const __half* a, int Ca, const __half* b, int Cb, int64_t rows, __half* out, cudaStream_t st);
int timestep_embed(const int64_t* t, int B, int dim, __half* out, cudaStream_t st);
int silu_f16(const __half* x, int64_t n, __half* out, cudaStream_t st);
// 1x1 conv on tiny channel counts (post_quant_conv 4->4, quant_conv 8->8): NHWC fp16.
int conv1x1_small(const __half* X, int64_t rows, int Cin, const float* Wt, const float* bias, int Cout, __half* out,
                  int ldo, cudaStream_t st);

// Scheduler-step fusions (fp32 latents NCHW [B,4,h,w]; eps fp16 NCHW [2B or B, ...]).
struct StepScalars {
  int kind;            // 0 euler/euler-a style (k-diffusion eps or v denoiser), 1 ddim
  int v_pred;          // 1: model output is v
  int cfg;             // 1: model_out holds [uncond; cond] halves
  float guidance;      // CFG scale
  float sigma;         // current sigma (k) ; unused for ddim
  float c_in_next;     // scaling of x for the NEXT unet call (written to x_in_next), 0 to skip
  float dt;            // sigma_down - sigma
  float sigma_up;      // ancestral noise scale (0: no noise)
  // ddim
  float sqrt_a_t, sqrt_1m_a_t, sqrt_a_prev, dir_coef, noise_coef;
};
int sched_step(const StepScalars& s, const float* x, const __half* model_out, const float* noise, float* x_out,
               float* denoised_out, __half* x_in_next, int B, int64_t per_sample, cudaStream_t st,
               const float* blend_orig = nullptr, const float* blend_mask = nullptr, float blend_u = 0.f);
int cat_channels_nchw(const __half* x, int Cx, const __half* extra, int Ce, int extra_batch, int B, int64_t hw,
                      __half* out, cudaStream_t st);
// generic sampler blocks: den = x * c_skip + cfg(model_out) * c_out ; out = sum coef[k] * in[k] (+ next UNet input)
int denoise_combine(const float* x, const __half* model_out, int cfg, float guidance, float c_skip, float c_out, int B,
                    int64_t per_sample, float* den, cudaStream_t st, const float* blend_orig = nullptr,
                    const float* blend_mask = nullptr, float blend_u = 0.f);
int dpm_error_partials(const float* x_low, const float* x_high, const float* x_prev, float atol, float rtol, int64_t n,
                       double* partials, cudaStream_t st);
int dpm_error_num_partials();
int lincomb(int n_terms, const float* const* in, const float* coef, int B, int64_t per_sample, float* out,
            __half* x_in, float c_in, int dup, cudaStream_t st);
int cfg_combine(const __half* model_out, float guidance, int B, int64_t per_sample, __half* out16, float* out32,
                cudaStream_t st);
// hires-fix / graft blending (blend.cu): separable lanczos resample + placement + where(rand >= p, ..) in one launch
int resample_select(const float* src, int BC, int SH, int SW, const int* ty_idx, const float* ty_w, int RH,
                    const int* tx_idx, const float* tx_w, int RW, int TH, int TW, int offy, int offx, int mode,
                    const float* bg, const float* other, const float* rnd, float p, int resampled_if_ge, float* out,
                    int FH, int FW, int oy, int ox, cudaStream_t st);
int rand_select(const float* a, const float* b, const float* rnd, float p, int64_t n, float* out, cudaStream_t st);
// Outpaint tail: histogram-match `result` to (source outside the mask + result inside it) over the whole batch, then mix the
// source back over it (postprocess.cu).  result / source / mask / out: [B, 3, HW] fp16 in [0, 1].
int resample_f32(const float* src, int64_t n_outer, int in_sz, int inner, const int* idx, const float* w, int ksize, int out_sz,
                 int clamp01, float* dst, cudaStream_t st);
// PNG encoder (png.cu): u8 NHWC [B, H, W, C] (C = 1 grey, 2 grey + alpha, 3 RGB, 4 RGBA) -> one PNG file per image
int png_sizes(int B, int H, int W, int C, size_t* workspace_bytes, size_t* out_stride);
int png_encode(const uint8_t* images, int B, int H, int W, int C, uint8_t* out, size_t out_stride, int64_t* out_len,
               void* workspace, size_t workspace_bytes, cudaStream_t st);
// Lossless WebP (VP8L) encoder (webp.cu): u8 NHWC [B, H, W, C] (C = 3 RGB, 4 RGBA) -> one WebP file per image
int webp_sizes(int B, int H, int W, int C, size_t* workspace_bytes, size_t* out_stride);
int webp_encode(const uint8_t* images, int B, int H, int W, int C, uint8_t* out, size_t out_stride, int64_t* out_len,
                void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t outpaint_scratch_bytes();
int outpaint_match_histograms(const __half* result, const __half* source, const __half* mask, int B, int64_t hw, __half* out,
                              void* scratch, cudaStream_t st);
// T2I-adapter front end: PixelUnshuffle(8) of an NCHW image into token-major NHWC [B, H/8 * W/8, C * 64] (channel
// c * 64 + dy * 8 + dx), and 2x2 average pooling of NHWC [B, H, W, C] (floor sizes, fp32 accumulation)
int pixel_unshuffle8_nchw_to_nhwc(const __half* x, int B, int C, int H, int W, __half* out, cudaStream_t st);
int avg_pool2x2_nhwc(const __half* x, int B, int H, int W, int C, __half* out, cudaStream_t st);
// Safety-checker front end and CLIP vision tower helpers (postprocess.cu)
int resample_u8(const uint8_t* src, int64_t n_outer, int in_sz, int inner, const int* bounds, const int* kk, int ksize,
                int out_sz, uint8_t* dst, cudaStream_t st);
int clip_normalize(const uint8_t* src, int B, int H, int W, int S, const float* mean3, const float* std3, __half* out,
                   cudaStream_t st);
int patchify(const __half* x, int B, int S, int P, int Kp, __half* out, cudaStream_t st);
int vision_embed(const __half* patches, const float* cls, const __half* pos, int B, int Ntok, int C, __half* out, cudaStream_t st);
int cosine_scores(const __half* img, int B, int D, const float* emb, int n_emb, float* scores, cudaStream_t st);
// LPW prompt weighting: out = emb * w[b, l] * (mean(emb[b]) / mean(emb[b] * w[b]))
int lpw_weight(const __half* emb, const float* weights, int B, int L, int C, __half* out, cudaStream_t st);
// unet input prep: out_f16[2B or B] = x * c_in (duplicated for CFG)
int scale_dup_latents(const float* x, float c_in, int dup, int B, int64_t per_sample, __half* out, cudaStream_t st);
// VAE tail: img = clamp(x/2+0.5, 0, 1): NHWC fp16 [B,H,W,ldx>=3] -> NCHW fp16 [B,3,H,W] (+ optional uint8 copy)
// postprocess=false: plain NHWC->NCHW copy of the 3 channels (the u8 copy, if requested, is always post-processed)
int vae_tail(const __half* x, int ldx, int B, int H, int W, bool postprocess, __half* out_nchw, uint8_t* out_u8_nhwc,
             cudaStream_t st);

}  // namespace gyre

namespace gyre {
// ---- weight packing (pack.cu); dtype: 0 fp16, 1 fp32
size_t conv3x3_packed_elems(int Cin, int Cout);
int pack_conv3x3(const void* w, int dtype, int Cin, int Cout, __half* out, cudaStream_t st);
size_t upconv3x3_packed_elems(int Cin, int Cout);
int pack_upconv3x3(const void* w, int dtype, int Cin, int Cout, __half* out, cudaStream_t st);
int cast_to_f16(const void* src, int dtype, int64_t rows, int cols, __half* dst, int ldd, cudaStream_t st);
int cast_to_f32(const void* src, int dtype, int64_t n, float* dst, cudaStream_t st);
int pack_geglu(const void* w, int dtype, int F, int K, const void* bias, int bias_dtype, __half* wp, float* bias_p,
               cudaStream_t st);
const char* last_error();
// ---- ToMe K/V merge (tome.cu): k, v [B, N, C] with row pitch ld -> k_out, v_out [B, N-r, C] dense
int tome_workspace_bytes(int B, int N, int C, size_t* bytes);
int tome_plan_offsets(int B, int N, int C, size_t* node_idx, size_t* unm_idx, size_t* src_idx);
int tome_merge_kv(const __half* k, const __half* v, int ld, int B, int N, int C, int r, __half* k_out, __half* v_out,
                  void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace gyre

namespace gyre {
// measurement-only (microbench.cu)
int mma_bench(int n, int naccs, int a_tmem, int reps, int blocks, long long* out_dev, cudaStream_t st);
// ---- tunables: small integer knobs read on the host at launch time.  Each starts from the environment
// variable GYRE_B200_<NAME> (if set) and can be changed through gyre_b200_set_tunable (A/B measurements).
enum Tunable { TUNE_ATT_VARIANT = 0, TUNE_PDL = 1, TUNE_GELU_FAST = 2, TUNE_GN_CHUNKS = 3, TUNE_UPCONV_FOLD = 4,
               TUNE_CTX_KV_CACHE = 5, TUNE_XATTN = 6, TUNE_GN_PHASE = 7, TUNE_MCAST = 8, TUNE_ATT_D128 = 9, TUNE_STREAMK = 10, TUNE_FORCE_BN = 11, TUNE_GEMM_STAGES = 12, TUNE_DEBUG = 13, TUNE_LN_SUB = 14, TUNE_GN_THREADS = 15, TUNE_LN_FUSE = 16, TUNE_CFG_SHARE = 17, TUNE_GN_FUSE = 18, TUNE_SK_MIN = 19, TUNE_COUNT = 20 };
int tunable(int id);
int set_tunable_by_name(const char* name, int value);
int get_tunable_by_name(const char* name, int* value);

namespace prof {
// Kernel families for the launch counter and the optional CUDA-event profiler (bench.py's roofline leg).
enum Family { F_GEMM = 0, F_CONV = 1, F_ATTN = 2, F_GROUPNORM = 3, F_LAYERNORM = 4, F_SOFTMAX = 5, F_ELEMENTWISE = 6,
              F_TOME = 7, F_COUNT = 8 };
unsigned long long launch_count();
void enable(int on);
bool enabled();
void reset();
int read(int family, unsigned long long* count, double* ms, double* flops, double* bytes);
int read_roofline_ms(int family, double peak_tflops, double peak_gbs, double* ideal_ms);
// RAII: counts `kernels` launches; when profiling is enabled brackets them with events on `st`.
class Scope {
 public:
  Scope(int family, double flops, double bytes, cudaStream_t st, int kernels = 1);
  ~Scope();
 private:
  int family_;
  double flops_, bytes_;
  cudaStream_t st_;
  void* e0_;
};
}  // namespace prof
}  // namespace gyre
