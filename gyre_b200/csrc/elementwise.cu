// Layout shuffles, tiny-channel convolutions, timestep embedding, the fused scheduler step (K7) and the
// VAE image tail (K8).  All HBM/latency-bound: coalesced, vectorised where the layout allows.
#include "common.cuh"
#include "ops.h"

namespace gyre {

static inline unsigned blocks_for(int64_t n, int threads) {
  return static_cast<unsigned>((n + threads - 1) / threads);
}

// ------------------------------------------------------------------ NCHW <-> NHWC (small C: latents / images)
__global__ void nchw_to_nhwc_kernel(const __half* __restrict__ x, int C, int HW, __half* __restrict__ out, int ldo,
                                    int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // over b*HW*ldo (padded NHWC)
  if (i >= total) return;
  const int c = static_cast<int>(i % ldo);
  const int64_t bp = i / ldo;
  const int p = static_cast<int>(bp % HW);
  const int64_t b = bp / HW;
  out[i] = c < C ? x[(b * C + c) * HW + p] : __float2half_rn(0.f);   // pad channels are written as zeros
}
int nchw_to_nhwc_f16(const __half* x, int B, int C, int H, int W, __half* out, int ldo, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(B) * ldo * H * W;
  GYRE_REQUIRE(total > 0 && ldo >= C, "nchw_to_nhwc: bad shape");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  nchw_to_nhwc_kernel<<<blocks_for(total, 256), 256, 0, st>>>(x, C, H * W, out, ldo, total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
__global__ void nhwc_to_nchw_kernel(const __half* __restrict__ x, int ldx, int C, int HW, __half* __restrict__ out,
                                    int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // over b*C*HW (NCHW order)
  if (i >= total) return;
  const int p = static_cast<int>(i % HW);
  const int64_t bc = i / HW;
  const int c = static_cast<int>(bc % C);
  const int64_t b = bc / C;
  out[i] = x[(b * HW + p) * ldx + c];
}
int nhwc_to_nchw_f16(const __half* x, int ldx, int B, int C, int H, int W, __half* out, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(B) * C * H * W;
  GYRE_REQUIRE(total > 0, "nhwc_to_nchw: empty");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  nhwc_to_nchw_kernel<<<blocks_for(total, 256), 256, 0, st>>>(x, ldx, C, H * W, out, total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dst (NHWC, fp16) += src (NCHW, fp16): ControlNet residuals arrive in the reference's NCHW layout (core.py:213-239).
// 32 x 32 (pixel, channel) tiles through shared memory: reads coalesced along pixels, writes along channels.
__global__ void add_nchw_to_nhwc_kernel(__half* __restrict__ dst, const __half* __restrict__ src, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;           // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, p = p0 + tx;
    tile[k][tx] = (c < C && p < HW) ? __half2float(src[(static_cast<int64_t>(b) * C + c) * HW + p]) : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int p = p0 + k, c = c0 + tx;
    if (p < HW && c < C) {
      __half* d = dst + (static_cast<int64_t>(b) * HW + p) * C + c;
      *d = __float2half_rn(__half2float(*d) + tile[tx][k]);
    }
  }
}
int add_nchw_to_nhwc_f16(__half* dst_nhwc, const __half* src_nchw, int B, int C, int HW, cudaStream_t st) {
  GYRE_REQUIRE(B > 0 && C > 0 && HW > 0 && B <= 65535, "add_nchw_to_nhwc: bad shape");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  add_nchw_to_nhwc_kernel<<<dim3((HW + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, st>>>(dst_nhwc, src_nchw, C, HW);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ nearest 2x upsample, NHWC, 16B vectors
__global__ void upsample2x_kernel(const uint4* __restrict__ x, int H, int W, int nvec, uint4* __restrict__ out,
                                  int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // over output vectors
  if (i >= total) return;
  const int v = static_cast<int>(i % nvec);
  int64_t p = i / nvec;
  const int xo = static_cast<int>(p % (2 * W));
  p /= (2 * W);
  const int yo = static_cast<int>(p % (2 * H));
  const int64_t b = p / (2 * H);
  out[i] = x[((b * H + (yo >> 1)) * W + (xo >> 1)) * nvec + v];
}
int upsample2x_nhwc(const __half* x, int B, int H, int W, int C, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(C % 8 == 0, "upsample2x: C must be a multiple of 8");
  const int64_t total = static_cast<int64_t>(B) * 4 * H * W * (C / 8);
  GYRE_REQUIRE(total > 0, "upsample2x: empty");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  upsample2x_kernel<<<blocks_for(total, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(x), H, W, C / 8,
                                                            reinterpret_cast<uint4*>(out), total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ channel concat (rows x (Ca+Cb))
__global__ void concat_kernel(const uint4* __restrict__ a, int na, const uint4* __restrict__ b, int nb,
                              uint4* __restrict__ out, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int n = na + nb;
  const int v = static_cast<int>(i % n);
  const int64_t r = i / n;
  out[i] = v < na ? a[r * na + v] : b[r * nb + (v - na)];
}
int concat_channels(const __half* a, int Ca, const __half* b, int Cb, int64_t rows, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0, "concat: channel counts must be multiples of 8");
  const int64_t total = rows * ((Ca + Cb) / 8);
  GYRE_REQUIRE(total > 0, "concat: empty");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  concat_kernel<<<blocks_for(total, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(a), Ca / 8,
                                                        reinterpret_cast<const uint4*>(b), Cb / 8,
                                                        reinterpret_cast<uint4*>(out), total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ timestep embedding [cos | sin], fp16
__global__ void timestep_embed_kernel(const int64_t* __restrict__ t, int dim, __half* __restrict__ out) {
  const int b = blockIdx.x;
  const int half = dim / 2;
  const float tv = static_cast<float>(t[b]);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float freq = expf(-logf(10000.0f) * static_cast<float>(i) / static_cast<float>(half));
    const float e = tv * freq;
    out[static_cast<int64_t>(b) * dim + i] = __float2half_rn(cosf(e));
    out[static_cast<int64_t>(b) * dim + half + i] = __float2half_rn(sinf(e));
  }
}
int timestep_embed(const int64_t* t, int B, int dim, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(B > 0 && dim > 0 && dim % 2 == 0, "timestep_embed: bad shape");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  timestep_embed_kernel<<<B, 128, 0, st>>>(t, dim, out);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void silu_kernel(const __half* __restrict__ x, int64_t n, __half* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float v = __half2float(x[i]);
  out[i] = __float2half_rn(v / (1.0f + __expf(-v)));
}
int silu_f16(const __half* x, int64_t n, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(n > 0, "silu: empty");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  silu_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, n, out);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void conv1x1_small_kernel(const __half* __restrict__ X, int64_t rows, int Cin, const float* __restrict__ Wt,
                                     const float* __restrict__ bias, int Cout, __half* __restrict__ out, int ldo) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // over rows * ldo
  if (i >= rows * ldo) return;
  const int o = static_cast<int>(i % ldo);
  const int64_t r = i / ldo;
  float acc = 0.f;
  if (o < Cout) {
    acc = bias ? bias[o] : 0.f;
    for (int c = 0; c < Cin; ++c) acc = fmaf(__half2float(X[r * Cin + c]), Wt[o * Cin + c], acc);
  }
  out[i] = __float2half_rn(acc);   // columns [Cout, ldo) are zero padding for a following TMA-fed conv
}
int conv1x1_small(const __half* X, int64_t rows, int Cin, const float* Wt, const float* bias, int Cout, __half* out,
                  int ldo, cudaStream_t st) {
  GYRE_REQUIRE(rows > 0 && Cin > 0 && Cout > 0 && ldo >= Cout, "conv1x1_small: bad shape");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  conv1x1_small_kernel<<<blocks_for(rows * ldo, 256), 256, 0, st>>>(X, rows, Cin, Wt, bias, Cout, out, ldo);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ K7: fused scheduler step
// One launch does: CFG combine (u + s*(g-u)), denoiser scalings (eps: x - sigma*eps ; v: c_out*v + c_skip*x),
// to_d + Euler update (x + (x-den)/sigma * dt), ancestral noise add (+ noise*sigma_up) - or the DDIM update -
// and emits the NEXT step's CFG-duplicated, c_in-scaled fp16 UNet input.  Latents stay fp32 across steps.
__global__ void sched_step_kernel(StepScalars s, const float* __restrict__ x, const __half* __restrict__ mo,
                                  const float* __restrict__ noise, float* __restrict__ x_out,
                                  float* __restrict__ den_out, __half* __restrict__ x_in_next, int64_t n_total,
                                  const float* __restrict__ blend_orig, const float* __restrict__ blend_mask,
                                  float blend_u) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_total) return;
  float m;
  if (s.cfg) {
    const float u = __half2float(mo[i]);
    const float g = __half2float(mo[n_total + i]);
    m = u + s.guidance * (g - u);
  } else {
    m = __half2float(mo[i]);
  }
  const float xv = x[i];
  float xn, den;
  if (s.kind == 0) {
    if (s.v_pred) {
      const float q = s.sigma * s.sigma + 1.0f;
      den = m * (-s.sigma * rsqrtf(q)) + xv * (1.0f / q);
    } else {
      den = xv - s.sigma * m;
    }
    // EnhancedInpaintMode.wrap_k_unet / _blend (unified_pipeline.py:620-636): where the mask still protects the
    // original at progress u, the predicted x0 is replaced by the original latents
    if (blend_mask != nullptr && blend_mask[i] > blend_u) den = blend_orig[i];
    const float d = (xv - den) / s.sigma;
    xn = xv + d * s.dt;
    if (s.sigma_up != 0.f) xn += noise[i] * s.sigma_up;
  } else {
    float eps = m;
    if (s.v_pred) {
      den = s.sqrt_a_t * xv - s.sqrt_1m_a_t * m;
      eps = s.sqrt_a_t * m + s.sqrt_1m_a_t * xv;
    } else {
      den = (xv - s.sqrt_1m_a_t * m) / s.sqrt_a_t;
    }
    xn = s.sqrt_a_prev * den + s.dir_coef * eps;
    if (s.noise_coef != 0.f) xn += s.noise_coef * noise[i];
  }
  x_out[i] = xn;
  if (den_out) den_out[i] = den;
  if (x_in_next) {
    const __half h = __float2half_rn(xn * s.c_in_next);
    x_in_next[i] = h;
    if (s.cfg) x_in_next[n_total + i] = h;
  }
}
int sched_step(const StepScalars& s, const float* x, const __half* model_out, const float* noise, float* x_out,
               float* denoised_out, __half* x_in_next, int B, int64_t per_sample, cudaStream_t st,
               const float* blend_orig, const float* blend_mask, float blend_u) {
  const int64_t n = static_cast<int64_t>(B) * per_sample;
  GYRE_REQUIRE(n > 0, "sched_step: empty");
  GYRE_REQUIRE((blend_orig == nullptr) == (blend_mask == nullptr), "sched_step: blend needs both original and mask");
  GYRE_REQUIRE(blend_mask == nullptr || s.kind == 0, "sched_step: x0 blending is defined for the k-diffusion step only");
  GYRE_REQUIRE(!((s.sigma_up != 0.f || (s.kind == 1 && s.noise_coef != 0.f)) && noise == nullptr),
               "sched_step: noise required");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  sched_step_kernel<<<blocks_for(n, 256), 256, 0, st>>>(s, x, model_out, noise, x_out, denoised_out,
                                                        s.c_in_next != 0.f ? x_in_next : nullptr, n, blend_orig,
                                                        blend_mask, blend_u);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ generic sampler building blocks
// The multi-evaluation samplers (Heun, DPM-2, LMS, DPM++ ...) are sequences of "denoise" and of linear
// combinations of latent-sized tensors with host-computed scalar coefficients; two kernels cover all of them.
//
// denoise: CFG combine + k-diffusion denoiser scalings (external.py:96-113 eps, :149-167 v):
//   den = x * c_skip + model * c_out      (eps-prediction: c_skip = 1, c_out = -sigma)
__global__ void denoise_kernel(const float* __restrict__ x, const __half* __restrict__ mo, int cfg, float guidance,
                               float c_skip, float c_out, int64_t n, float* __restrict__ den,
                               const float* __restrict__ blend_orig, const float* __restrict__ blend_mask,
                               float blend_u) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float m;
  if (cfg) {
    const float u = __half2float(mo[i]);
    const float g = __half2float(mo[n + i]);
    m = u + guidance * (g - u);
  } else {
    m = __half2float(mo[i]);
  }
  float d = x[i] * c_skip + m * c_out;
  if (blend_mask != nullptr && blend_mask[i] > blend_u) d = blend_orig[i];
  den[i] = d;
}
int denoise_combine(const float* x, const __half* model_out, int cfg, float guidance, float c_skip, float c_out, int B,
                    int64_t per_sample, float* den, cudaStream_t st, const float* blend_orig, const float* blend_mask,
                    float blend_u) {
  const int64_t n = static_cast<int64_t>(B) * per_sample;
  GYRE_REQUIRE(n > 0 && x && model_out && den, "denoise: bad arguments");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  GYRE_REQUIRE((blend_orig == nullptr) == (blend_mask == nullptr), "denoise: blend needs both original and mask");
  denoise_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, model_out, cfg, guidance, c_skip, c_out, n, den, blend_orig,
                                                     blend_mask, blend_u);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// out = sum_k coef[k] * in[k]  (k < n_terms <= 6, fp32); optionally also the next UNet input
// x_in = fp16(out * c_in), duplicated for CFG.  `out` may alias any input (pure elementwise).
struct LinArgs {
  const float* in[6];
  float coef[6];
  int n_terms;
};
__global__ void lincomb_kernel(LinArgs a, int64_t n, float* __restrict__ out, __half* __restrict__ x_in, float c_in,
                               int dup) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (k < a.n_terms) acc = fmaf(a.coef[k], a.in[k][i], acc);
  if (out) out[i] = acc;
  if (x_in) {
    const __half h = __float2half_rn(acc * c_in);
    x_in[i] = h;
    if (dup) x_in[n + i] = h;
  }
}
int lincomb(int n_terms, const float* const* in, const float* coef, int B, int64_t per_sample, float* out,
            __half* x_in, float c_in, int dup, cudaStream_t st) {
  const int64_t n = static_cast<int64_t>(B) * per_sample;
  GYRE_REQUIRE(n > 0 && n_terms >= 1 && n_terms <= 6 && in && coef && (out || x_in), "lincomb: bad arguments");
  LinArgs a;
  for (int k = 0; k < 6; ++k) {
    a.in[k] = k < n_terms ? in[k] : nullptr;
    a.coef[k] = k < n_terms ? coef[k] : 0.f;
    GYRE_REQUIRE(k >= n_terms || in[k] != nullptr, "lincomb: null input %d", k);
  }
  a.n_terms = n_terms;
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  lincomb_kernel<<<blocks_for(n, 256), 256, 0, st>>>(a, n, out, x_in, c_in, dup);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Error norm of the adaptive DPM-Solver (k_diffusion/sampling.py:461-462):
//   delta = max(atol, rtol * max(|x_low|, |x_prev|));  partial[b] = sum over the block's elements of ((x_low - x_high) / delta)^2
// One fp64 partial per block, a fixed number of blocks and a fixed in-block tree: the host adds the partials in order,
// so the accept / reject decision is reproducible.
constexpr int kErrBlocks = 256;
__global__ void __launch_bounds__(256) dpm_error_kernel(const float* __restrict__ x_low, const float* __restrict__ x_high,
                                                        const float* __restrict__ x_prev, float atol, float rtol,
                                                        int64_t n, double* __restrict__ partials) {
  double acc = 0.0;
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * 256) {
    const float lo = x_low[i];
    const float delta = fmaxf(atol, rtol * fmaxf(fabsf(lo), fabsf(x_prev[i])));
    const float q = (lo - x_high[i]) / delta;
    acc += static_cast<double>(q) * static_cast<double>(q);
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}
int dpm_error_partials(const float* x_low, const float* x_high, const float* x_prev, float atol, float rtol, int64_t n,
                       double* partials, cudaStream_t st) {
  GYRE_REQUIRE(x_low && x_high && x_prev && partials && n > 0, "dpm_error: bad arguments");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  dpm_error_kernel<<<kErrBlocks, 256, 0, st>>>(x_low, x_high, x_prev, atol, rtol, n, partials);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int dpm_error_num_partials() { return kErrBlocks; }

// T2I-adapter front end (gyre/pipeline/t2i_adapter/adapter.py:104, 119-122): nn.PixelUnshuffle(8) fused with the
// NCHW -> NHWC transpose: out[b, y, x, c * 64 + dy * 8 + dx] = in[b, c, 8y + dy, 8x + dx].
__global__ void pixel_unshuffle8_kernel(const __half* __restrict__ x, int C, int H, int W, __half* __restrict__ out,
                                        int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // over the output
  if (i >= total) return;
  const int Co = C * 64, Ho = H / 8, Wo = W / 8;
  const int co = static_cast<int>(i % Co);
  const int64_t p = i / Co;
  const int xo = static_cast<int>(p % Wo);
  const int yo = static_cast<int>((p / Wo) % Ho);
  const int64_t b = p / (static_cast<int64_t>(Wo) * Ho);
  const int c = co >> 6, dy = (co >> 3) & 7, dx = co & 7;
  out[i] = x[((b * C + c) * H + (8 * yo + dy)) * W + (8 * xo + dx)];
}
int pixel_unshuffle8_nchw_to_nhwc(const __half* x, int B, int C, int H, int W, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(x && out && B > 0 && C > 0 && H % 8 == 0 && W % 8 == 0 && H > 0 && W > 0, "pixel_unshuffle: bad shape");
  const int64_t total = static_cast<int64_t>(B) * C * H * W;
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  pixel_unshuffle8_kernel<<<blocks_for(total, 256), 256, 0, st>>>(x, C, H, W, out, total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// nn.AvgPool2d(kernel_size=2, stride=2) (adapter.py Downsample without conv, :56-58) on NHWC, 8 channels per thread
__global__ void avg_pool2x2_kernel(const __half* __restrict__ x, int H, int W, int C, __half* __restrict__ out, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // over output vectors of 8 channels
  if (i >= total) return;
  const int nv = C >> 3, Ho = H / 2, Wo = W / 2;
  const int v = static_cast<int>(i % nv);
  const int64_t p = i / nv;
  const int xo = static_cast<int>(p % Wo);
  const int yo = static_cast<int>((p / Wo) % Ho);
  const int64_t b = p / (static_cast<int64_t>(Wo) * Ho);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + (((b * H + 2 * yo + dy) * W + 2 * xo + dx) * C + v * 8));
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        acc[2 * k] += f.x;
        acc[2 * k + 1] += f.y;
      }
    }
  __align__(16) __half2 o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = __floats2half2_rn(acc[2 * k] * 0.25f, acc[2 * k + 1] * 0.25f);
  *reinterpret_cast<uint4*>(out + (((b * Ho + yo) * Wo + xo) * C + v * 8)) = *reinterpret_cast<uint4*>(o);
}
int avg_pool2x2_nhwc(const __half* x, int B, int H, int W, int C, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(x && out && B > 0 && H >= 2 && W >= 2 && C > 0 && C % 8 == 0, "avg_pool2x2: bad shape");
  const int64_t total = static_cast<int64_t>(B) * (H / 2) * (W / 2) * (C / 8);
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  avg_pool2x2_kernel<<<blocks_for(total, 256), 256, 0, st>>>(x, H, W, C, out, total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Prompt weighting of the LPW text embedding (gyre/pipeline/text_embedding/lpw_text_embedding.py:352-371):
//   previous_mean = emb.mean([-2, -1]); emb *= weights[..., None]; emb *= previous_mean / emb.mean([-2, -1])
// One CTA per prompt, two passes over its [L, C] block (the second one from L2): fixed-order fp32 sums, one rounding.
__global__ void __launch_bounds__(256) lpw_weight_kernel(const __half* __restrict__ emb, const float* __restrict__ w, int L,
                                                         int C, __half* __restrict__ out) {
  const int b = blockIdx.x;
  const int64_t n = static_cast<int64_t>(L) * C;
  const __half* e = emb + b * n;
  const float* wb = w + static_cast<int64_t>(b) * L;
  float s0 = 0.f, s1 = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    const float v = __half2float(e[i]);
    s0 += v;
    s1 = fmaf(v, wb[i / C], s1);
  }
  __shared__ float r0[256], r1[256];
  r0[threadIdx.x] = s0;
  r1[threadIdx.x] = s1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) {
      r0[threadIdx.x] += r0[threadIdx.x + o];
      r1[threadIdx.x] += r1[threadIdx.x + o];
    }
    __syncthreads();
  }
  const float scale = r0[0] / r1[0];     // ratio of the means == ratio of the sums
  __half* o_ = out + b * n;
  for (int64_t i = threadIdx.x; i < n; i += 256) o_[i] = __float2half_rn(__half2float(e[i]) * wb[i / C] * scale);
}
int lpw_weight(const __half* emb, const float* weights, int B, int L, int C, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(emb && weights && out && B > 0 && L > 0 && C > 0, "lpw_weight: bad arguments");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  lpw_weight_kernel<<<B, 256, 0, st>>>(emb, weights, L, C, out);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// UNet input of the inpaint / depth models (EnhancedRunwayInpaintMode.wrap_unet, unified_pipeline.py:668-690;
// UnetWithExtraChannels, unet/core.py:21-37): out[b] = cat([x[b], extra[b % extra_batch]], dim=channel), NCHW fp16.
// The extra channels (mask + masked-image latents) are NOT scaled by c_in (see the reference's comment there).
__global__ void cat_channels_nchw_kernel(const __half* __restrict__ x, int Cx, const __half* __restrict__ extra, int Ce,
                                         int extra_batch, int64_t hw, int64_t total, __half* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int C = Cx + Ce;
  const int64_t p = i % hw;
  const int64_t bc = i / hw;
  const int c = static_cast<int>(bc % C);
  const int64_t b = bc / C;
  out[i] = c < Cx ? x[(b * Cx + c) * hw + p] : extra[((b % extra_batch) * Ce + (c - Cx)) * hw + p];
}
int cat_channels_nchw(const __half* x, int Cx, const __half* extra, int Ce, int extra_batch, int B, int64_t hw,
                      __half* out, cudaStream_t st) {
  GYRE_REQUIRE(x && extra && out && B > 0 && Cx > 0 && Ce > 0 && extra_batch > 0 && hw > 0, "cat_channels: bad arguments");
  const int64_t total = static_cast<int64_t>(B) * (Cx + Ce) * hw;
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  cat_channels_nchw_kernel<<<blocks_for(total, 256), 256, 0, st>>>(x, Cx, extra, Ce, extra_batch, hw, total, out);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void scale_dup_kernel(const float* __restrict__ x, float c_in, int dup, int64_t n, __half* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const __half h = __float2half_rn(x[i] * c_in);
  out[i] = h;
  if (dup) out[n + i] = h;
}
int scale_dup_latents(const float* x, float c_in, int dup, int B, int64_t per_sample, __half* out, cudaStream_t st) {
  const int64_t n = static_cast<int64_t>(B) * per_sample;
  GYRE_REQUIRE(n > 0, "scale_dup: empty");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  scale_dup_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, c_in, dup, n, out);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// CFG combine for callers that drive the UNet through the NoisePredictionUNet protocol instead of the fused
// step: out = u + s * (g - u) over [uncond ; cond] halves (gyre/pipeline/unet/cfg.py:54-57)
__global__ void cfg_combine_kernel(const __half* __restrict__ mo, float s, int64_t n, __half* __restrict__ o16,
                                   float* __restrict__ o32) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float u = __half2float(mo[i]);
  const float g = __half2float(mo[n + i]);
  const float m = u + s * (g - u);
  if (o16) o16[i] = __float2half_rn(m);
  if (o32) o32[i] = m;
}
int cfg_combine(const __half* model_out, float guidance, int B, int64_t per_sample, __half* out16, float* out32,
                cudaStream_t st) {
  const int64_t n = static_cast<int64_t>(B) * per_sample;
  GYRE_REQUIRE(n > 0 && (out16 || out32), "cfg_combine: empty");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  cfg_combine_kernel<<<blocks_for(n, 256), 256, 0, st>>>(model_out, guidance, n, out16, out32);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ text encoder: embeddings + causal attention
// hidden = token_embedding[ids] + position_embedding[pos]   (transformers CLIPTextEmbeddings.forward)
__global__ void embed_tokens_kernel(const int64_t* __restrict__ ids, const uint4* __restrict__ tok,
                                    const uint4* __restrict__ pos, int L, int nvec, int vocab, uint4* __restrict__ out,
                                    int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % nvec);
  const int64_t row = i / nvec;
  const int p = static_cast<int>(row % L);
  int64_t id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const uint4 a = tok[id * nvec + v], b = pos[static_cast<int64_t>(p) * nvec + v];
  const __half2* ah = reinterpret_cast<const __half2*>(&a);
  const __half2* bh = reinterpret_cast<const __half2*>(&b);
  uint4 o;
  __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 x = __half22float2(ah[k]), y = __half22float2(bh[k]);
    oh[k] = __floats2half2_rn(x.x + y.x, x.y + y.y);
  }
  out[i] = o;
}
int embed_tokens(const int64_t* ids, const __half* tok_emb, const __half* pos_emb, int B, int L, int C, int vocab,
                 __half* out, cudaStream_t st) {
  GYRE_REQUIRE(ids && tok_emb && pos_emb && out && B > 0 && L > 0 && C % 8 == 0, "embed_tokens: bad arguments");
  const int64_t total = static_cast<int64_t>(B) * L * (C / 8);
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  embed_tokens_kernel<<<blocks_for(total, 256), 256, 0, st>>>(ids, reinterpret_cast<const uint4*>(tok_emb),
                                                              reinterpret_cast<const uint4*>(pos_emb), L, C / 8, vocab,
                                                              reinterpret_cast<uint4*>(out), total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Causal self-attention of one (batch, head) per CTA for L <= 128, d <= 128 (CLIP: L = 77, d = 64): 0.1 % of the
// sampling path's work, so plain FMA code - K and V of the head sit in shared memory as fp32, one warp per query
// row, lanes over keys for the scores and over channels for P.V.  (The tcgen05 flash kernels have no mask input.)
__global__ void __launch_bounds__(128) causal_attn_kernel(const __half* __restrict__ qkv, int L, int heads, int d,
                                                          float scale, __half* __restrict__ out) {
  extern __shared__ float sm[];
  const int C = heads * d;
  const int h = blockIdx.x, b = blockIdx.y;
  const int dp = d + 1;
  float* sK = sm;                     // [L][d + 1]
  float* sV = sK + L * dp;            // [L][d + 1]
  float* sQ = sV + L * dp;            // [4 warps][d]
  float* sP = sQ + 4 * d;             // [4 warps][L]
  const __half* base = qkv + static_cast<int64_t>(b) * L * 3 * C + h * d;
  for (int i = threadIdx.x; i < L * d; i += 128) {
    const int j = i / d, c = i - j * d;
    sK[j * dp + c] = __half2float(base[static_cast<int64_t>(j) * 3 * C + C + c]);
    sV[j * dp + c] = __half2float(base[static_cast<int64_t>(j) * 3 * C + 2 * C + c]);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = sQ + warp * d;
  float* pr = sP + warp * L;
  for (int i = warp; i < L; i += 4) {
    for (int c = lane; c < d; c += 32) q[c] = __half2float(base[static_cast<int64_t>(i) * 3 * C + c]) * scale;
    __syncwarp();
    float sc[4];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane + 32 * t;
      float a = -INFINITY;
      if (j <= i) {
        a = 0.f;
        for (int c = 0; c < d; ++c) a = fmaf(q[c], sK[j * dp + c], a);
      }
      sc[t] = a;
      m = fmaxf(m, a);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane + 32 * t;
      const float e = j <= i ? __expf(sc[t] - m) : 0.f;
      if (j < L) pr[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int c = lane; c < d; c += 32) {
      float acc = 0.f;
      for (int j = 0; j <= i; ++j) acc = fmaf(pr[j], sV[j * dp + c], acc);
      out[(static_cast<int64_t>(b) * L + i) * C + h * d + c] = __float2half_rn(acc * inv);
    }
    __syncwarp();
  }
}
int causal_attention_short(const __half* qkv, int B, int L, int heads, int d, float scale, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(qkv && out && B > 0 && heads > 0, "causal_attention: bad arguments");
  GYRE_REQUIRE(L >= 1 && L <= 128 && d >= 8 && d <= 128, "causal_attention: L=%d (<= 128), d=%d (<= 128)", L, d);
  const size_t smem = (static_cast<size_t>(2) * L * (d + 1) + 4 * d + 4 * L) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    GYRE_CHECK_CUDA(cudaFuncSetAttribute(causal_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  prof::Scope ps(prof::F_ATTN, 2.0 * B * heads * static_cast<double>(L) * L * d, 2.0 * B * L * heads * d * 4.0, st);
  causal_attn_kernel<<<dim3(heads, B), 128, smem, st>>>(qkv, L, heads, d, scale, out);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ K8: VAE tail  (x/2+0.5).clamp(0,1), NHWC -> NCHW
__global__ void vae_tail_kernel(const __half* __restrict__ x, int ldx, int HW, int post, __half* __restrict__ out,
                                uint8_t* __restrict__ u8, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // over b*HW pixels
  if (i >= total) return;
  const int64_t b = i / HW;
  const int p = static_cast<int>(i % HW);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float raw = __half2float(x[i * ldx + c]);
    const float v = fminf(fmaxf(raw * 0.5f + 0.5f, 0.f), 1.f);
    if (out) out[(b * 3 + c) * HW + p] = post ? __float2half_rn(v) : x[i * ldx + c];
    if (u8) u8[i * 3 + c] = static_cast<uint8_t>(__float2int_rn(v * 255.0f));
  }
}
int vae_tail(const __half* x, int ldx, int B, int H, int W, bool postprocess, __half* out_nchw, uint8_t* out_u8_nhwc,
             cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(B) * H * W;
  GYRE_REQUIRE(total > 0, "vae_tail: empty");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  vae_tail_kernel<<<blocks_for(total, 256), 256, 0, st>>>(x, ldx, H * W, postprocess ? 1 : 0, out_nchw, out_u8_nhwc, total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gyre
