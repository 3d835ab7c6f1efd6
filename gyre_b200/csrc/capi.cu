// extern "C" surface of libgyre_b200.so (include/gyre_b200.h): argument checking + forwarding to the op
// layer.  Nothing here throws; every failure becomes a negative status + thread-local message.
#include <cstring>

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

size_t gyre_b200_groupnorm_scratch_floats(int B, int HW, int G) { return gn_partials_floats(B, HW, G); }

int gyre_b200_groupnorm(const void* x1, int C1, const void* x2, int C2, int B, int HW, int G, float eps,
                        const float* gamma, const float* beta, int silu, void* out, float* scratch,
                        gyre_b200_stream stream) {
  GYRE_REQUIRE(x1 && gamma && beta && out && scratch, "groupnorm: null operand");
  return groupnorm_nhwc(static_cast<const __half*>(x1), C1, static_cast<const __half*>(x2), C2, B, HW, G, eps, gamma,
                        beta, silu != 0, static_cast<__half*>(out), scratch, S(stream));
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

int gyre_b200_scale_latents(const float* x, float c_in, int dup, int batch, int64_t per_sample, void* out,
                            gyre_b200_stream stream) {
  GYRE_REQUIRE(x && out, "scale_latents: null operand");
  return scale_dup_latents(x, c_in, dup, batch, per_sample, static_cast<__half*>(out), S(stream));
}

}  // extern "C"

// ---- TEMPORARY stubs (replaced by model.cu / tome.cu)
extern "C" {
#define GYRE_NOT_YET(name) do { set_last_error(name ": not implemented yet"); return -100; } while (0)
int gyre_b200_unet_create(const gyre_b200_unet_config*, gyre_b200_handle*) { GYRE_NOT_YET("unet_create"); }
int gyre_b200_load_weight(gyre_b200_handle, const char*, const void*, int, const int64_t*, int, gyre_b200_stream) { GYRE_NOT_YET("load_weight"); }
int gyre_b200_finalize(gyre_b200_handle) { GYRE_NOT_YET("finalize"); }
int gyre_b200_unet_workspace_bytes(gyre_b200_handle, int, int, int, int, size_t*) { GYRE_NOT_YET("unet_workspace_bytes"); }
int gyre_b200_unet_forward(gyre_b200_handle, const void*, const int64_t*, const void*, int, int, int, int, const int32_t*, void*, void*, size_t, gyre_b200_stream) { GYRE_NOT_YET("unet_forward"); }
int gyre_b200_vae_create(const gyre_b200_vae_config*, gyre_b200_handle*) { GYRE_NOT_YET("vae_create"); }
int gyre_b200_vae_workspace_bytes(gyre_b200_handle, int, int, int, size_t*) { GYRE_NOT_YET("vae_workspace_bytes"); }
int gyre_b200_vae_decode(gyre_b200_handle, const void*, int, int, int, int, void*, uint8_t*, void*, size_t, gyre_b200_stream) { GYRE_NOT_YET("vae_decode"); }
int gyre_b200_vae_encode(gyre_b200_handle, const void*, int, int, int, void*, void*, size_t, gyre_b200_stream) { GYRE_NOT_YET("vae_encode"); }
int gyre_b200_destroy(gyre_b200_handle) { GYRE_NOT_YET("destroy"); }
int gyre_b200_tome_workspace_bytes(int, int, int, size_t*) { GYRE_NOT_YET("tome_workspace_bytes"); }
int gyre_b200_tome_merge_kv(const void*, const void*, int, int, int, int, void*, void*, void*, size_t, gyre_b200_stream) { GYRE_NOT_YET("tome_merge_kv"); }
}
