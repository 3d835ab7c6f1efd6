// extern "C" surface of libgyre_b200.so (include/gyre_b200.h): argument checking + forwarding to the op
// layer.  Nothing here throws; every failure becomes a negative status + thread-local message.
#include <cstring>
#include <new>

#include "../../include/gyre_b200.h"
#include "common.cuh"
#include "model.h"
#include "ops.h"

using namespace gyre;

namespace {
inline cudaStream_t S(gyre_b200_stream s) { return reinterpret_cast<cudaStream_t>(s); }

int to_epilogue(const gyre_b200_epilogue* e, Epilogue* out) {
  GYRE_REQUIRE(e != nullptr && e->out != nullptr, "epilogue: null output");
  out->bias = e->bias;
  out->rowgroup_bias = static_cast<const __half*>(e->rowgroup_bias);
  out->rows_per_group = e->rows_per_group > 0 ? e->rows_per_group : 1;
  out->rgb_ld = e->rgb_ld;
  out->residual = static_cast<const __half*>(e->residual);
  out->ldr = e->ldr;
  out->act = e->act;
  out->out = e->out;
  out->ldo = e->ldo;
  out->out_mode = e->out_f32 ? OUT_F32 : OUT_F16;
  out->sk_ws = e->sk_ws;
  out->sk_ws_bytes = e->sk_ws_bytes;
  out->sk_flags = e->sk_flags;
  out->sk_flags_count = e->sk_flags_count;
  out->rowstat_out = static_cast<float2*>(e->rowstat_out);
  out->ln_rowstat = static_cast<const float2*>(e->ln_rowstat);
  out->ln_colsum = e->ln_colsum;
  out->ln_parts = e->ln_parts;
  out->ln_inv_c = e->ln_inv_c;
  out->ln_eps = e->ln_eps;
  out->gn_out = e->gn_out;
  out->gn_groups = e->gn_groups;
  out->gn_nparts = e->gn_nparts;
  return 0;
}
}  // namespace

extern "C" {

int gyre_b200_abi_version(void) { return GYRE_B200_ABI_VERSION; }

int gyre_b200_last_error(char* buf, size_t n) {
  const char* e = last_error();
  const size_t len = strlen(e);
  if (buf && n > 0) {
    const size_t c = len < n - 1 ? len : n - 1;
    memcpy(buf, e, c);
    buf[c] = 0;
  }
  return static_cast<int>(len);
}

unsigned long long gyre_b200_launch_count(void) { return prof::launch_count(); }
int gyre_b200_prof_enable(int on) {
  prof::enable(on);
  return 0;
}
int gyre_b200_prof_reset(void) {
  prof::reset();
  return 0;
}
int gyre_b200_prof_read(int family, unsigned long long* count, double* ms, double* flops, double* bytes) {
  GYRE_REQUIRE(family >= 0 && family < prof::F_COUNT, "prof_read: family %d", family);
  return prof::read(family, count, ms, flops, bytes);
}

int gyre_b200_prof_read_roofline(int family, double peak_tflops, double peak_gbs, double* ideal_ms) {
  GYRE_REQUIRE(family >= 0 && family < prof::F_COUNT, "prof_read_roofline: family %d", family);
  return prof::read_roofline_ms(family, peak_tflops, peak_gbs, ideal_ms);
}

int gyre_b200_debug_mma_bench(int n, int naccs, int a_tmem, int reps, int blocks, long long* out_dev,
                              gyre_b200_stream stream) {
  GYRE_REQUIRE(out_dev, "mma_bench: null output");
  return mma_bench(n, naccs, a_tmem, reps, blocks, out_dev, S(stream));
}

int gyre_b200_debug_attention_trace(long long* dev_buf, int capacity) {
  attention_set_trace(dev_buf, capacity);
  return 0;
}

int gyre_b200_set_tunable(const char* name, int value) { return set_tunable_by_name(name, value); }
int gyre_b200_get_tunable(const char* name, int* value) { return get_tunable_by_name(name, value); }

int gyre_b200_gemm(const void* A, int lda, int K1, const void* A2, int lda2, int K2, const void* W, int ldw, int M,
                   int N, const gyre_b200_epilogue* ep, gyre_b200_stream stream) {
  Epilogue e;
  GYRE_TRY(to_epilogue(ep, &e));
  GYRE_REQUIRE(A && W, "gemm: null operand");
  return gemm2_f16(static_cast<const __half*>(A), lda, K1, static_cast<const __half*>(A2), lda2, K2,
                   static_cast<const __half*>(W), ldw, M, N, e, S(stream));
}

int gyre_b200_pack_geglu(const void* W, int dtype, int F, int K, const void* bias, int bias_dtype, void* Wp,
                         float* bias_p, gyre_b200_stream stream) {
  GYRE_REQUIRE((W && Wp) || (bias && bias_p), "pack_geglu: null operand");
  return pack_geglu(W, dtype, F, K, bias, bias_dtype, static_cast<__half*>(Wp), bias_p, S(stream));
}

int gyre_b200_conv3x3(const void* X, int ldx, int B, int H, int W, int Cin, const void* Wp, int Cout, int stride,
                      int pad, const gyre_b200_epilogue* ep, gyre_b200_stream stream) {
  Epilogue e;
  GYRE_TRY(to_epilogue(ep, &e));
  GYRE_REQUIRE(X && Wp, "conv3x3: null operand");
  return conv3x3_f16(static_cast<const __half*>(X), ldx, B, H, W, Cin, static_cast<const __half*>(Wp), Cout, stride,
                     pad, e, S(stream));
}

size_t gyre_b200_conv3x3_packed_elems(int Cin, int Cout) { return conv3x3_packed_elems(Cin, Cout); }

int gyre_b200_pack_conv3x3(const void* W, int dtype, int Cin, int Cout, void* Wp, gyre_b200_stream stream) {
  GYRE_REQUIRE(W && Wp, "pack_conv3x3: null operand");
  return pack_conv3x3(W, dtype, Cin, Cout, static_cast<__half*>(Wp), S(stream));
}

size_t gyre_b200_upconv3x3_packed_elems(int Cin, int Cout) { return upconv3x3_packed_elems(Cin, Cout); }

int gyre_b200_pack_upconv3x3(const void* W, int dtype, int Cin, int Cout, void* Wp4, gyre_b200_stream stream) {
  GYRE_REQUIRE(W && Wp4, "pack_upconv3x3: null operand");
  return pack_upconv3x3(W, dtype, Cin, Cout, static_cast<__half*>(Wp4), S(stream));
}

int gyre_b200_upconv2x(const void* X, int ldx, int B, int H, int W, int Cin, const void* Wp4, int Cout,
                       const gyre_b200_epilogue* ep, gyre_b200_stream stream) {
  Epilogue e;
  GYRE_TRY(to_epilogue(ep, &e));
  GYRE_REQUIRE(X && Wp4, "upconv2x: null operand");
  return upconv2x_f16(static_cast<const __half*>(X), ldx, B, H, W, Cin, static_cast<const __half*>(Wp4), Cout, e,
                      S(stream));
}

size_t gyre_b200_groupnorm_scratch_floats(int B, int HW, int G) { return gn_partials_floats(B, HW, G); }

int gyre_b200_groupnorm(const void* x1, int C1, const void* x2, int C2, int B, int HW, int G, float eps,
                        const float* gamma, const float* beta, int silu, void* out, float* scratch,
                        gyre_b200_stream stream) {
  GYRE_REQUIRE(x1 && gamma && beta && out && scratch, "groupnorm: null operand");
  return groupnorm_nhwc(static_cast<const __half*>(x1), C1, static_cast<const __half*>(x2), C2, B, HW, G, eps, gamma,
                        beta, silu != 0, static_cast<__half*>(out), scratch, S(stream));
}

int gyre_b200_conv3x3_gn_parts(int B, int H, int W, int Cout, int stride, int pad, int groups) {
  return conv3x3_gn_parts(B, H, W, Cout, stride, pad, groups);
}

int gyre_b200_groupnorm_pre_ok(int C, int HW, int G) { return groupnorm_pre_ok(C, HW, G) ? 1 : 0; }

int gyre_b200_groupnorm_pre(const void* x, int C, int B, int HW, int G, float eps, const float* gamma, const float* beta,
                            int silu, void* out, const float* pre, int nparts, float* stats, gyre_b200_stream stream) {
  GYRE_REQUIRE(x && gamma && beta && out && pre, "groupnorm_pre: null operand");
  return groupnorm_nhwc_pre(static_cast<const __half*>(x), C, B, HW, G, eps, gamma, beta, silu != 0,
                            static_cast<__half*>(out), pre, nparts, stats, S(stream));
}

int gyre_b200_layernorm(const void* x, int rows, int C, float eps, const float* gamma, const float* beta, void* out,
                        gyre_b200_stream stream) {
  GYRE_REQUIRE(x && gamma && beta && out, "layernorm: null operand");
  return layernorm_rows(static_cast<const __half*>(x), rows, C, eps, gamma, beta, static_cast<__half*>(out),
                        S(stream));
}

int gyre_b200_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int B, int heads,
                        int Nq, int Nk, int d, float scale, void* out, int ldo, gyre_b200_stream stream) {
  GYRE_REQUIRE(q && k && v && out, "attention: null operand");
  return attention_f16(static_cast<const __half*>(q), ldq, static_cast<const __half*>(k), ldk,
                       static_cast<const __half*>(v), ldv, B, heads, Nq, Nk, d, scale, static_cast<__half*>(out), ldo,
                       S(stream));
}

int gyre_b200_sched_step(const gyre_b200_step* s, const float* x, const void* model_out, const float* noise,
                         float* x_out, float* denoised_out, void* x_in_next, int batch, int64_t per_sample,
                         gyre_b200_stream stream) {
  GYRE_REQUIRE(s && x && model_out && x_out, "sched_step: null operand");
  StepScalars k;
  static_assert(sizeof(StepScalars) == sizeof(gyre_b200_step), "StepScalars must mirror gyre_b200_step");
  memcpy(&k, s, sizeof(k));
  return sched_step(k, x, static_cast<const __half*>(model_out), noise, x_out, denoised_out,
                    static_cast<__half*>(x_in_next), batch, per_sample, S(stream));
}

int gyre_b200_cfg_combine(const void* model_out, float guidance, int batch, int64_t per_sample, void* out_f16,
                          float* out_f32, gyre_b200_stream stream) {
  GYRE_REQUIRE(model_out, "cfg_combine: null operand");
  return cfg_combine(static_cast<const __half*>(model_out), guidance, batch, per_sample, static_cast<__half*>(out_f16),
                     out_f32, S(stream));
}

int gyre_b200_denoise(const float* x, const void* model_out, int cfg, float guidance, float c_skip, float c_out,
                      int batch, int64_t per_sample, float* denoised, gyre_b200_stream stream) {
  GYRE_REQUIRE(x && model_out && denoised, "denoise: null operand");
  return denoise_combine(x, static_cast<const __half*>(model_out), cfg, guidance, c_skip, c_out, batch, per_sample,
                         denoised, S(stream));
}

int gyre_b200_denoise_blend(const float* x, const void* model_out, int cfg, float guidance, float c_skip, float c_out,
                            int batch, int64_t per_sample, float* denoised, const float* blend_orig,
                            const float* blend_mask, float blend_u, gyre_b200_stream stream) {
  GYRE_REQUIRE(x && model_out && denoised && blend_orig && blend_mask, "denoise_blend: null operand");
  return denoise_combine(x, static_cast<const __half*>(model_out), cfg, guidance, c_skip, c_out, batch, per_sample,
                         denoised, S(stream), blend_orig, blend_mask, blend_u);
}

int gyre_b200_sched_step_blend(const gyre_b200_step* s, const float* x, const void* model_out, const float* noise,
                               float* x_out, float* denoised_out, void* x_in_next, int batch, int64_t per_sample,
                               const float* blend_orig, const float* blend_mask, float blend_u,
                               gyre_b200_stream stream) {
  GYRE_REQUIRE(s && x && model_out && x_out && blend_orig && blend_mask, "sched_step_blend: null operand");
  StepScalars k;
  static_assert(sizeof(StepScalars) == sizeof(gyre_b200_step), "step struct drift");
  memcpy(&k, s, sizeof(k));
  return sched_step(k, x, static_cast<const __half*>(model_out), noise, x_out, denoised_out,
                    static_cast<__half*>(x_in_next), batch, per_sample, S(stream), blend_orig, blend_mask, blend_u);
}

int gyre_b200_cat_channels(const void* x, int channels, const void* extra, int extra_channels, int extra_batch, int batch,
                           int64_t hw, void* out, gyre_b200_stream stream) {
  return cat_channels_nchw(static_cast<const __half*>(x), channels, static_cast<const __half*>(extra), extra_channels,
                           extra_batch, batch, hw, static_cast<__half*>(out), S(stream));
}

int gyre_b200_lincomb(int n_terms, const float* const* inputs_host, const float* coefs_host, int batch,
                      int64_t per_sample, float* out, void* x_in_next, float c_in, int dup, gyre_b200_stream stream) {
  GYRE_REQUIRE(inputs_host && coefs_host, "lincomb: null operand");
  return lincomb(n_terms, inputs_host, coefs_host, batch, per_sample, out, static_cast<__half*>(x_in_next), c_in, dup,
                 S(stream));
}

int gyre_b200_dpm_error_partials(const float* x_low, const float* x_high, const float* x_prev, float atol, float rtol,
                                 int64_t n, double* partials, gyre_b200_stream stream) {
  return dpm_error_partials(x_low, x_high, x_prev, atol, rtol, n, partials, S(stream));
}

int gyre_b200_dpm_error_num_partials(void) { return dpm_error_num_partials(); }

int gyre_b200_scale_latents(const float* x, float c_in, int dup, int batch, int64_t per_sample, void* out,
                            gyre_b200_stream stream) {
  GYRE_REQUIRE(x && out, "scale_latents: null operand");
  return scale_dup_latents(x, c_in, dup, batch, per_sample, static_cast<__half*>(out), S(stream));
}

int gyre_b200_resample_f32(const float* src, int64_t n_outer, int in_size, int inner, const int32_t* idx, const float* weights,
                           int ksize, int out_size, int clamp01, float* dst, gyre_b200_stream stream) {
  return resample_f32(src, n_outer, in_size, inner, idx, weights, ksize, out_size, clamp01, dst, S(stream));
}

int gyre_b200_png_sizes(int batch, int height, int width, int channels, size_t* workspace_bytes, size_t* out_stride) {
  return png_sizes(batch, height, width, channels, workspace_bytes, out_stride);
}

int gyre_b200_png_encode(const void* images_u8_nhwc, int batch, int height, int width, int channels, void* out,
                         size_t out_stride, int64_t* out_lengths, void* workspace, size_t workspace_bytes,
                         gyre_b200_stream stream) {
  return png_encode(static_cast<const uint8_t*>(images_u8_nhwc), batch, height, width, channels, static_cast<uint8_t*>(out),
                    out_stride, out_lengths, workspace, workspace_bytes, S(stream));
}

int gyre_b200_webp_sizes(int batch, int height, int width, int channels, size_t* workspace_bytes, size_t* out_stride) {
  return webp_sizes(batch, height, width, channels, workspace_bytes, out_stride);
}

int gyre_b200_webp_encode(const void* images_u8_nhwc, int batch, int height, int width, int channels, void* out,
                          size_t out_stride, int64_t* out_lengths, void* workspace, size_t workspace_bytes,
                          gyre_b200_stream stream) {
  return webp_encode(static_cast<const uint8_t*>(images_u8_nhwc), batch, height, width, channels, static_cast<uint8_t*>(out),
                     out_stride, out_lengths, workspace, workspace_bytes, S(stream));
}

size_t gyre_b200_outpaint_scratch_bytes(void) { return outpaint_scratch_bytes(); }

int gyre_b200_outpaint_match_histograms(const void* result, const void* source, const void* outmask, int batch, int64_t hw,
                                        void* out, void* scratch, gyre_b200_stream stream) {
  return outpaint_match_histograms(static_cast<const __half*>(result), static_cast<const __half*>(source),
                                   static_cast<const __half*>(outmask), batch, hw, static_cast<__half*>(out), scratch,
                                   S(stream));
}

int gyre_b200_lpw_weight(const void* emb, const float* weights, int batch, int tokens, int channels, void* out,
                         gyre_b200_stream stream) {
  return lpw_weight(static_cast<const __half*>(emb), weights, batch, tokens, channels, static_cast<__half*>(out), S(stream));
}

int gyre_b200_gemm_rowstat_parts(int M, int N) { return gemm_rowstat_parts(M, N); }

int gyre_b200_ln_finalize_rows(const void* parts, int nparts, int M, int C, float eps, void* mean_rstd,
                               gyre_b200_stream stream) {
  return ln_finalize_rows(static_cast<const float2*>(parts), nparts, M, C, eps, static_cast<float2*>(mean_rstd), S(stream));
}

int gyre_b200_ln_fold_linear(const void* W, int N, int K, const float* gamma, const float* beta, const float* bias,
                             void* W_out, float* colsum, float* lnbias, gyre_b200_stream stream) {
  return ln_fold_linear(static_cast<const __half*>(W), N, K, gamma, beta, bias, static_cast<__half*>(W_out), colsum,
                        lnbias, S(stream));
}

int gyre_b200_resample_select(const float* src, int planes, int src_h, int src_w, const int32_t* taps_y_idx,
                              const float* taps_y_w, int resized_h, const int32_t* taps_x_idx, const float* taps_x_w,
                              int resized_w, int target_h, int target_w, int off_y, int off_x, int mode,
                              const float* background, const float* other, const float* rand_map, float p,
                              int resampled_if_ge, float* out, int frame_h, int frame_w, int frame_y, int frame_x,
                              gyre_b200_stream stream) {
  return resample_select(src, planes, src_h, src_w, taps_y_idx, taps_y_w, resized_h, taps_x_idx, taps_x_w, resized_w,
                         target_h, target_w, off_y, off_x, mode, background, other, rand_map, p, resampled_if_ge, out,
                         frame_h, frame_w, frame_y, frame_x, S(stream));
}

int gyre_b200_rand_select(const float* a, const float* b, const float* rand_map, float p, int64_t n, float* out,
                          gyre_b200_stream stream) {
  return rand_select(a, b, rand_map, p, n, out, S(stream));
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ models
namespace {
inline Model* M(gyre_b200_handle h) { return reinterpret_cast<Model*>(h); }

int make_exec(Exec* ex, void* workspace, size_t bytes, gyre_b200_stream stream) {
  GYRE_REQUIRE(workspace != nullptr, "null workspace");
  GYRE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  ex->st = S(stream);
  ex->dry = false;
  ex->base = static_cast<uint8_t*>(workspace);
  ex->cap = bytes & ~static_cast<size_t>(1023);
  return 0;
}
}  // namespace

extern "C" {

int gyre_b200_unet_create(const gyre_b200_unet_config* cfg, gyre_b200_handle* out) {
  GYRE_REQUIRE(cfg && out, "unet_create: null argument");
  GYRE_REQUIRE(cfg->num_levels >= 1 && cfg->num_levels <= 4 && cfg->layers_per_block >= 1, "unet_create: bad topology");
  for (int i = 0; i < cfg->num_levels; ++i) {
    const int c = cfg->block_out_channels[i];
    GYRE_REQUIRE(c > 0 && c % 64 == 0, "unet_create: block_out_channels[%d]=%d must be a multiple of 64", i, c);
    GYRE_REQUIRE(c % cfg->norm_num_groups == 0, "unet_create: channels %d not divisible by %d groups", c,
                 cfg->norm_num_groups);
    if (cfg->attn_levels[i]) {
      const int hds = cfg->num_heads[i];
      GYRE_REQUIRE(hds > 0 && c % hds == 0 && (c / hds) % 8 == 0 && c / hds <= 192,
                   "unet_create: level %d head dim %d unsupported", i, hds > 0 ? c / hds : -1);
    }
  }
  GYRE_REQUIRE(cfg->in_channels >= 1 && cfg->in_channels <= 16, "unet_create: in_channels %d", cfg->in_channels);
  GYRE_REQUIRE(cfg->cross_attention_dim > 0 && cfg->cross_attention_dim % 8 == 0, "unet_create: cross_attention_dim");
  GYRE_REQUIRE(cfg->addition_embed_dim >= 0 && cfg->addition_embed_dim % 8 == 0, "unet_create: addition_embed_dim");
  for (int i = 0; i < cfg->num_levels; ++i)
    GYRE_REQUIRE(cfg->transformer_depth[i] >= 0 && cfg->transformer_depth[i] <= 16, "unet_create: transformer_depth");
  GYRE_REQUIRE(cfg->norm_num_groups > 0 && cfg->norm_num_groups <= 32, "unet_create: norm_num_groups");
  UNetModel* m = new (std::nothrow) UNetModel(*cfg);
  GYRE_REQUIRE(m != nullptr, "unet_create: out of host memory");
  *out = reinterpret_cast<gyre_b200_handle>(static_cast<Model*>(m));
  return 0;
}

int gyre_b200_load_weight(gyre_b200_handle h, const char* key, const void* data, int dtype, const int64_t* shape,
                          int ndim, gyre_b200_stream stream) {
  GYRE_REQUIRE(h, "load_weight: null handle");
  return M(h)->load(key, data, dtype, shape, ndim, S(stream));
}

int gyre_b200_finalize(gyre_b200_handle h) {
  GYRE_REQUIRE(h, "finalize: null handle");
  return M(h)->finalize();
}

int gyre_b200_unet_workspace_bytes(gyre_b200_handle h, int batch, int height, int width, int ctx_len, size_t* bytes) {
  GYRE_REQUIRE(h && bytes, "unet_workspace_bytes: null argument");
  GYRE_REQUIRE(M(h)->is_unet(), "unet_workspace_bytes: handle is not a UNet");
  Exec ex;
  ex.dry = true;
  ex.cap = static_cast<size_t>(1) << 60;
  // the dry run reserves the ToMe scratch as well (see UNetModel::transformer)
  GYRE_TRY(static_cast<UNetModel*>(M(h))->forward(ex, nullptr, nullptr, nullptr, nullptr, batch, height, width, ctx_len,
                                                  nullptr, nullptr));
  *bytes = ex.peak + (1u << 20);
  return 0;
}

int gyre_b200_unet_forward_cond(gyre_b200_handle h, const void* sample, const int64_t* timestep, const void* ctx,
                                const void* add_cond, int batch, int height, int width, int ctx_len,
                                const int32_t* tome_r_host, void* out, void* workspace, size_t workspace_bytes,
                                gyre_b200_stream stream) {
  GYRE_REQUIRE(h && sample && timestep && out, "unet_forward: null argument");
  GYRE_REQUIRE(M(h)->is_unet(), "unet_forward: handle is not a UNet");
  Exec ex;
  GYRE_TRY(make_exec(&ex, workspace, workspace_bytes, stream));
  return static_cast<UNetModel*>(M(h))->forward(ex, static_cast<const __half*>(sample), timestep,
                                                static_cast<const __half*>(ctx), static_cast<const __half*>(add_cond),
                                                batch, height, width, ctx_len, tome_r_host, static_cast<__half*>(out));
}

int gyre_b200_controlnet_forward(gyre_b200_handle h, const void* sample, const int64_t* timestep, const void* ctx,
                                 const void* cond, int batch, int height, int width, int ctx_len, void* const* down_out,
                                 int n_down, void* mid_out, void* workspace, size_t workspace_bytes,
                                 gyre_b200_stream stream) {
  GYRE_REQUIRE(h && sample && timestep && ctx && cond && down_out && mid_out, "controlnet_forward: null argument");
  GYRE_REQUIRE(M(h)->is_unet() && static_cast<UNetModel*>(M(h))->is_controlnet(),
               "controlnet_forward: handle is not a ControlNet (create it with cfg.controlnet = 1)");
  Exec ex;
  GYRE_TRY(make_exec(&ex, workspace, workspace_bytes, stream));
  ControlNetIO io;
  io.cond = static_cast<const __half*>(cond);
  io.down_out = reinterpret_cast<__half* const*>(down_out);
  io.n_down = n_down;
  io.mid_out = static_cast<__half*>(mid_out);
  return static_cast<UNetModel*>(M(h))->forward(ex, static_cast<const __half*>(sample), timestep,
                                                static_cast<const __half*>(ctx), nullptr, batch, height, width, ctx_len,
                                                nullptr, nullptr, &io);
}

int gyre_b200_unet_forward(gyre_b200_handle h, const void* sample, const int64_t* timestep, const void* ctx, int batch,
                           int height, int width, int ctx_len, const int32_t* tome_r_host, void* out, void* workspace,
                           size_t workspace_bytes, gyre_b200_stream stream) {
  return gyre_b200_unet_forward_cond(h, sample, timestep, ctx, nullptr, batch, height, width, ctx_len, tome_r_host, out,
                                     workspace, workspace_bytes, stream);
}

int gyre_b200_unet_set_cfg_duplicate(gyre_b200_handle h, int on) {
  GYRE_REQUIRE(h, "unet_set_cfg_duplicate: null handle");
  GYRE_REQUIRE(M(h)->is_unet(), "unet_set_cfg_duplicate: handle is not a UNet");
  static_cast<UNetModel*>(M(h))->set_cfg_duplicate(on != 0);
  return 0;
}

int gyre_b200_unet_set_context(gyre_b200_handle h, const void* ctx, int batch, int ctx_len, gyre_b200_stream stream) {
  GYRE_REQUIRE(h, "unet_set_context: null handle");
  GYRE_REQUIRE(M(h)->is_unet(), "unet_set_context: handle is not a UNet");
  return static_cast<UNetModel*>(M(h))->set_context(static_cast<const __half*>(ctx), batch, ctx_len, S(stream));
}

int gyre_b200_unet_set_control_residuals(gyre_b200_handle h, const void* const* down_residuals, int n_down,
                                         const void* mid_residual) {
  GYRE_REQUIRE(h, "unet_set_control_residuals: null handle");
  GYRE_REQUIRE(M(h)->is_unet(), "unet_set_control_residuals: handle is not a UNet");
  return static_cast<UNetModel*>(M(h))->set_control_residuals(reinterpret_cast<const __half* const*>(down_residuals), n_down,
                                                              static_cast<const __half*>(mid_residual));
}

int gyre_b200_unet_set_adapter_states(gyre_b200_handle h, const void* const* states, int n_states) {
  GYRE_REQUIRE(h, "unet_set_adapter_states: null handle");
  GYRE_REQUIRE(M(h)->is_unet(), "unet_set_adapter_states: handle is not a UNet");
  return static_cast<UNetModel*>(M(h))->set_adapter_states(reinterpret_cast<const __half* const*>(states), n_states);
}

int gyre_b200_unet_num_skips(gyre_b200_handle h) {
  if (!h || !M(h)->is_unet()) return -1;
  return static_cast<UNetModel*>(M(h))->num_skips();
}

int gyre_b200_unet_num_transformer_blocks(gyre_b200_handle h) {
  if (!h || !M(h)->is_unet()) return -1;
  return static_cast<UNetModel*>(M(h))->num_transformer_blocks();
}

int gyre_b200_vae_create(const gyre_b200_vae_config* cfg, gyre_b200_handle* out) {
  GYRE_REQUIRE(cfg && out, "vae_create: null argument");
  GYRE_REQUIRE(cfg->num_levels >= 1 && cfg->num_levels <= 4 && cfg->layers_per_block >= 1, "vae_create: bad topology");
  for (int i = 0; i < cfg->num_levels; ++i) {
    const int c = cfg->block_out_channels[i];
    GYRE_REQUIRE(c > 0 && c % 64 == 0 && c % cfg->norm_num_groups == 0,
                 "vae_create: block_out_channels[%d]=%d must be a multiple of 64 and of the group count", i, c);
  }
  GYRE_REQUIRE(cfg->in_channels == 3 && cfg->out_channels == 3, "vae_create: RGB only");
  GYRE_REQUIRE(cfg->latent_channels >= 1 && cfg->latent_channels <= 8, "vae_create: latent_channels");
  VAEModel* m = new (std::nothrow) VAEModel(*cfg);
  GYRE_REQUIRE(m != nullptr, "vae_create: out of host memory");
  *out = reinterpret_cast<gyre_b200_handle>(static_cast<Model*>(m));
  return 0;
}

int gyre_b200_vae_workspace_bytes(gyre_b200_handle h, int batch, int latent_h, int latent_w, size_t* bytes) {
  GYRE_REQUIRE(h && bytes, "vae_workspace_bytes: null argument");
  GYRE_REQUIRE(M(h)->kind() == 1, "vae_workspace_bytes: handle is not a VAE");
  VAEModel* v = static_cast<VAEModel*>(M(h));
  Exec ex;
  ex.dry = true;
  ex.cap = static_cast<size_t>(1) << 60;
  GYRE_TRY(v->decode(ex, nullptr, batch, latent_h, latent_w, true, reinterpret_cast<__half*>(16), nullptr));
  const size_t dec = ex.peak;
  Exec ex2;
  ex2.dry = true;
  ex2.cap = static_cast<size_t>(1) << 60;
  GYRE_TRY(v->encode(ex2, nullptr, batch, latent_h * 8, latent_w * 8, nullptr));
  *bytes = (dec > ex2.peak ? dec : ex2.peak) + 4096;
  return 0;
}

int gyre_b200_vae_decode(gyre_b200_handle h, const void* z, int batch, int latent_h, int latent_w, int postprocess,
                         void* img, uint8_t* img_u8, void* workspace, size_t workspace_bytes, gyre_b200_stream stream) {
  GYRE_REQUIRE(h && z, "vae_decode: null argument");
  GYRE_REQUIRE(M(h)->kind() == 1, "vae_decode: handle is not a VAE");
  Exec ex;
  GYRE_TRY(make_exec(&ex, workspace, workspace_bytes, stream));
  return static_cast<VAEModel*>(M(h))->decode(ex, static_cast<const __half*>(z), batch, latent_h, latent_w,
                                              postprocess != 0, static_cast<__half*>(img), img_u8);
}

int gyre_b200_vae_encode(gyre_b200_handle h, const void* img, int batch, int height, int width, void* moments,
                         void* workspace, size_t workspace_bytes, gyre_b200_stream stream) {
  GYRE_REQUIRE(h && img && moments, "vae_encode: null argument");
  GYRE_REQUIRE(M(h)->kind() == 1, "vae_encode: handle is not a VAE");
  Exec ex;
  GYRE_TRY(make_exec(&ex, workspace, workspace_bytes, stream));
  return static_cast<VAEModel*>(M(h))->encode(ex, static_cast<const __half*>(img), batch, height, width,
                                              static_cast<__half*>(moments));
}

int gyre_b200_clip_vision_create(const gyre_b200_clip_vision_config* cfg, gyre_b200_handle* out) {
  GYRE_REQUIRE(cfg && out, "clip_vision_create: null argument");
  GYRE_REQUIRE(cfg->image_size > 0 && cfg->patch_size > 0 && cfg->image_size % cfg->patch_size == 0 && cfg->num_layers > 0,
               "clip_vision_create: bad sizes");
  GYRE_REQUIRE(cfg->hidden_size % 8 == 0 && cfg->intermediate_size % 8 == 0 && cfg->projection_dim % 8 == 0 &&
                   cfg->num_heads > 0 && cfg->hidden_size % cfg->num_heads == 0 && (cfg->hidden_size / cfg->num_heads) % 8 == 0 &&
                   (cfg->hidden_size / cfg->num_heads) <= 192,
               "clip_vision_create: hidden %d / heads %d / projection %d unsupported", cfg->hidden_size, cfg->num_heads,
               cfg->projection_dim);
  GYRE_REQUIRE(cfg->hidden_act == 0 || cfg->hidden_act == 1, "clip_vision_create: hidden_act must be quick_gelu (0) or gelu (1)");
  GYRE_REQUIRE(cfg->num_concepts >= 0 && cfg->num_special >= 0 && cfg->projection_dim >= 0, "clip_vision_create: bad concept counts");
  ClipVisionModel* m = new (std::nothrow) ClipVisionModel(*cfg);
  GYRE_REQUIRE(m != nullptr, "clip_vision_create: out of host memory");
  *out = reinterpret_cast<gyre_b200_handle>(static_cast<Model*>(m));
  return 0;
}

int gyre_b200_clip_vision_workspace_bytes(gyre_b200_handle h, int batch, size_t* bytes) {
  GYRE_REQUIRE(h && bytes, "clip_vision_workspace_bytes: null argument");
  GYRE_REQUIRE(M(h)->kind() == 4, "clip_vision_workspace_bytes: handle is not a CLIP vision model");
  Exec ex;
  ex.dry = true;
  ex.cap = static_cast<size_t>(1) << 60;
  GYRE_TRY(static_cast<ClipVisionModel*>(M(h))->forward(ex, nullptr, batch, nullptr, nullptr, nullptr, 0, true));
  *bytes = ex.peak + 65536;     // (the scores path adds a few class-token rows on top of the tower's buffers)
  return 0;
}

int gyre_b200_clip_vision_hidden(gyre_b200_handle h, const void* pixel_values, int batch, int skip_last, void* hidden,
                                 void* workspace, size_t workspace_bytes, gyre_b200_stream stream) {
  GYRE_REQUIRE(h && pixel_values && hidden, "clip_vision_hidden: null argument");
  GYRE_REQUIRE(M(h)->kind() == 4, "clip_vision_hidden: handle is not a CLIP vision model");
  Exec ex;
  GYRE_TRY(make_exec(&ex, workspace, workspace_bytes, stream));
  return static_cast<ClipVisionModel*>(M(h))->forward(ex, static_cast<const __half*>(pixel_values), batch, nullptr, nullptr,
                                                      static_cast<__half*>(hidden), skip_last, true);
}

int gyre_b200_style_adapter_create(const gyre_b200_style_adapter_config* cfg, gyre_b200_handle* out) {
  GYRE_REQUIRE(cfg && out, "style_adapter_create: null argument");
  GYRE_REQUIRE(cfg->width > 0 && cfg->width % 8 == 0 && cfg->context_dim > 0 && cfg->context_dim % 8 == 0 && cfg->num_head > 0 &&
                   cfg->width % cfg->num_head == 0 && (cfg->width / cfg->num_head) % 8 == 0 && cfg->width / cfg->num_head <= 192,
               "style_adapter_create: width %d / heads %d / context %d unsupported", cfg->width, cfg->num_head, cfg->context_dim);
  GYRE_REQUIRE(cfg->n_layers > 0 && cfg->num_token > 0, "style_adapter_create: bad sizes");
  StyleAdapterModel* m = new (std::nothrow) StyleAdapterModel(*cfg);
  GYRE_REQUIRE(m != nullptr, "style_adapter_create: out of host memory");
  *out = reinterpret_cast<gyre_b200_handle>(static_cast<Model*>(m));
  return 0;
}

int gyre_b200_style_adapter_workspace_bytes(gyre_b200_handle h, int batch, int tokens, size_t* bytes) {
  GYRE_REQUIRE(h && bytes, "style_adapter_workspace_bytes: null argument");
  GYRE_REQUIRE(M(h)->kind() == 5, "style_adapter_workspace_bytes: handle is not a style adapter");
  Exec ex;
  ex.dry = true;
  ex.cap = static_cast<size_t>(1) << 60;
  GYRE_TRY(static_cast<StyleAdapterModel*>(M(h))->forward(ex, nullptr, batch, tokens, nullptr));
  *bytes = ex.peak + 4096;
  return 0;
}

int gyre_b200_style_adapter_forward(gyre_b200_handle h, const void* x, int batch, int tokens, void* out, void* workspace,
                                    size_t workspace_bytes, gyre_b200_stream stream) {
  GYRE_REQUIRE(h && x && out, "style_adapter_forward: null argument");
  GYRE_REQUIRE(M(h)->kind() == 5, "style_adapter_forward: handle is not a style adapter");
  Exec ex;
  GYRE_TRY(make_exec(&ex, workspace, workspace_bytes, stream));
  return static_cast<StyleAdapterModel*>(M(h))->forward(ex, static_cast<const __half*>(x), batch, tokens, static_cast<__half*>(out));
}

int gyre_b200_safety_scores(gyre_b200_handle h, const void* pixel_values, int batch, void* image_embeds, float* scores,
                            void* workspace, size_t workspace_bytes, gyre_b200_stream stream) {
  GYRE_REQUIRE(h && pixel_values && scores, "safety_scores: null argument");
  GYRE_REQUIRE(M(h)->kind() == 4, "safety_scores: handle is not a CLIP vision model");
  Exec ex;
  GYRE_TRY(make_exec(&ex, workspace, workspace_bytes, stream));
  return static_cast<ClipVisionModel*>(M(h))->forward(ex, static_cast<const __half*>(pixel_values), batch,
                                                      static_cast<__half*>(image_embeds), scores);
}

int gyre_b200_resample_u8(const void* src, int64_t n_outer, int in_size, int inner, const int32_t* bounds,
                          const int32_t* coeffs, int ksize, int out_size, void* dst, gyre_b200_stream stream) {
  return resample_u8(static_cast<const uint8_t*>(src), n_outer, in_size, inner, bounds, coeffs, ksize, out_size,
                     static_cast<uint8_t*>(dst), S(stream));
}

int gyre_b200_clip_normalize(const void* src_u8_nhwc, int batch, int height, int width, int crop, const float* mean3,
                             const float* std3, void* out, gyre_b200_stream stream) {
  return clip_normalize(static_cast<const uint8_t*>(src_u8_nhwc), batch, height, width, crop, mean3, std3,
                        static_cast<__half*>(out), S(stream));
}

int gyre_b200_adapter_create(const gyre_b200_adapter_config* cfg, gyre_b200_handle* out) {
  GYRE_REQUIRE(cfg && out, "adapter_create: null argument");
  GYRE_REQUIRE(cfg->num_levels >= 1 && cfg->num_levels <= 4 && cfg->nums_rb >= 1 && cfg->nums_rb <= 8, "adapter_create: bad sizes");
  GYRE_REQUIRE(cfg->cin > 0 && cfg->cin % 64 == 0, "adapter_create: cin = %d must be 64 x image channels", cfg->cin);
  GYRE_REQUIRE(cfg->light || cfg->ksize == 1 || cfg->ksize == 3, "adapter_create: ksize must be 1 or 3");
  if (cfg->light)
    for (int i = 0; i < cfg->num_levels; ++i)
      GYRE_REQUIRE(cfg->channels[i] % 32 == 0, "adapter_create: light adapters work at channels / 4 (channels[%d] = %d)", i,
                   cfg->channels[i]);
  for (int i = 0; i < cfg->num_levels; ++i)
    GYRE_REQUIRE(cfg->channels[i] > 0 && cfg->channels[i] % 8 == 0, "adapter_create: channels[%d] = %d must be a multiple of 8",
                 i, cfg->channels[i]);
  AdapterModel* m = new (std::nothrow) AdapterModel(*cfg);
  GYRE_REQUIRE(m != nullptr, "adapter_create: out of host memory");
  *out = reinterpret_cast<gyre_b200_handle>(static_cast<Model*>(m));
  return 0;
}

int gyre_b200_adapter_workspace_bytes(gyre_b200_handle h, int batch, int height, int width, size_t* bytes) {
  GYRE_REQUIRE(h && bytes, "adapter_workspace_bytes: null argument");
  GYRE_REQUIRE(M(h)->kind() == 3, "adapter_workspace_bytes: handle is not a T2I adapter");
  Exec ex;
  ex.dry = true;
  ex.cap = static_cast<size_t>(1) << 60;
  GYRE_TRY(static_cast<AdapterModel*>(M(h))->forward(ex, nullptr, batch, height, width, nullptr));
  *bytes = ex.peak + 4096;
  return 0;
}

int gyre_b200_adapter_forward(gyre_b200_handle h, const void* image, int batch, int height, int width, void* const* features,
                              int n_features, void* workspace, size_t workspace_bytes, gyre_b200_stream stream) {
  GYRE_REQUIRE(h && image && features, "adapter_forward: null argument");
  GYRE_REQUIRE(M(h)->kind() == 3, "adapter_forward: handle is not a T2I adapter");
  AdapterModel* m = static_cast<AdapterModel*>(M(h));
  GYRE_REQUIRE(n_features == m->num_levels(), "adapter_forward: %d feature buffers given, the model has %d levels", n_features,
               m->num_levels());
  for (int i = 0; i < n_features; ++i) GYRE_REQUIRE(features[i] != nullptr, "adapter_forward: null feature buffer %d", i);
  Exec ex;
  GYRE_TRY(make_exec(&ex, workspace, workspace_bytes, stream));
  return m->forward(ex, static_cast<const __half*>(image), batch, height, width, reinterpret_cast<__half* const*>(features));
}

int gyre_b200_clip_create(const gyre_b200_clip_config* cfg, gyre_b200_handle* out) {
  GYRE_REQUIRE(cfg && out, "clip_create: null argument");
  GYRE_REQUIRE(cfg->vocab_size > 0 && cfg->num_layers > 0 && cfg->num_heads > 0 && cfg->max_positions > 0 &&
                   cfg->max_positions <= 128,
               "clip_create: bad sizes");
  GYRE_REQUIRE(cfg->hidden_size % 8 == 0 && cfg->intermediate_size % 8 == 0 && cfg->hidden_size % cfg->num_heads == 0 &&
                   (cfg->hidden_size / cfg->num_heads) <= 128,
               "clip_create: hidden %d / heads %d unsupported", cfg->hidden_size, cfg->num_heads);
  GYRE_REQUIRE(cfg->hidden_act == 0 || cfg->hidden_act == 1, "clip_create: hidden_act must be quick_gelu (0) or gelu (1)");
  ClipTextModel* m = new (std::nothrow) ClipTextModel(*cfg);
  GYRE_REQUIRE(m != nullptr, "clip_create: out of host memory");
  *out = reinterpret_cast<gyre_b200_handle>(static_cast<Model*>(m));
  return 0;
}

int gyre_b200_clip_workspace_bytes(gyre_b200_handle h, int batch, int seq_len, size_t* bytes) {
  GYRE_REQUIRE(h && bytes, "clip_workspace_bytes: null argument");
  GYRE_REQUIRE(M(h)->kind() == 2, "clip_workspace_bytes: handle is not a text encoder");
  Exec ex;
  ex.dry = true;
  ex.cap = static_cast<size_t>(1) << 60;
  GYRE_TRY(static_cast<ClipTextModel*>(M(h))->forward(ex, nullptr, batch, seq_len, 0, true, nullptr));
  *bytes = ex.peak + 4096;
  return 0;
}

int gyre_b200_clip_forward(gyre_b200_handle h, const int64_t* input_ids, int batch, int seq_len, int skip_last,
                           int apply_final_ln, void* out, void* workspace, size_t workspace_bytes,
                           gyre_b200_stream stream) {
  GYRE_REQUIRE(h && input_ids && out, "clip_forward: null argument");
  GYRE_REQUIRE(M(h)->kind() == 2, "clip_forward: handle is not a text encoder");
  Exec ex;
  GYRE_TRY(make_exec(&ex, workspace, workspace_bytes, stream));
  return static_cast<ClipTextModel*>(M(h))->forward(ex, input_ids, batch, seq_len, skip_last, apply_final_ln != 0,
                                                    static_cast<__half*>(out));
}

int gyre_b200_destroy(gyre_b200_handle h) {
  if (h) delete M(h);
  return 0;
}

int gyre_b200_tome_plan_offsets(int batch, int tokens, int channels, size_t* node_idx, size_t* unm_idx, size_t* src_idx) {
  GYRE_REQUIRE(node_idx && unm_idx && src_idx, "tome_plan_offsets: null argument");
  return tome_plan_offsets(batch, tokens, channels, node_idx, unm_idx, src_idx);
}
int gyre_b200_tome_workspace_bytes(int batch, int tokens, int channels, size_t* bytes) {
  GYRE_REQUIRE(bytes, "tome_workspace_bytes: null argument");
  return tome_workspace_bytes(batch, tokens, channels, bytes);
}

int gyre_b200_tome_merge_kv(const void* k, const void* v, int batch, int tokens, int channels, int r, void* k_out,
                            void* v_out, void* workspace, size_t workspace_bytes, gyre_b200_stream stream) {
  GYRE_REQUIRE(k && v && k_out && v_out, "tome_merge_kv: null operand");
  return tome_merge_kv(static_cast<const __half*>(k), static_cast<const __half*>(v), channels, batch, tokens, channels,
                       r, static_cast<__half*>(k_out), static_cast<__half*>(v_out), workspace, workspace_bytes,
                       S(stream));
}

}  // extern "C"
