// Image-space tail of an outpaint request (reference: gyre/pipeline/unified_pipeline.py:2493-2510 - reference image =
// source outside the mask + result inside it, `images.match_histograms(result, reference)`, source mixed back over the
// result; gyre/images.py:667-672 quantises both to uint8 and calls gyre/match_histograms.py:12-37, a per-channel CDF
// match over the WHOLE batch whose float64 table is truncated back to uint8).
//
// The reference does this on the host (`.cpu().numpy()`, np.bincount, np.interp) - at tens of images per second per
// node that round trip is the serial tail SURVEY 8f2 points at.  Here: one histogram kernel (integer atomics: exact
// and order-independent), one single-CTA kernel that builds the three 256-entry tables with numpy's interpolation
// expression in fp64, one apply kernel fused with the final composite.  Every arithmetic step rounds where the
// reference's tensor statements round (fp16 images: per-operation fp16 rounding), so the result is bit-identical.
#include "common.cuh"
#include "ops.h"

namespace gyre {

static inline unsigned blocks_for(int64_t n, int threads) {
  return static_cast<unsigned>((n + threads - 1) / threads);
}

// (x.to(float32) * 255).round().to(uint8): round half to even like torch.round, values are in [0, 1]
__device__ __forceinline__ int quant8(float v) {
  const float r = rintf(v * 255.0f);
  return static_cast<int>(fminf(fmaxf(r, 0.f), 255.f));
}

// reference = source * (1 - outmask) + result * outmask, evaluated in the image dtype (fp16: one rounding per operation)
__device__ __forceinline__ __half composite(__half src, __half val, __half m) {
  const __half one = __float2half_rn(1.0f);
  return __hadd(__hmul(src, __hsub(one, m)), __hmul(val, m));
}

// hist [2][3][256]: [0] the result image, [1] the reference image; element i of [B, 3, HW]
__global__ void outpaint_hist_kernel(const __half* __restrict__ result, const __half* __restrict__ source,
                                     const __half* __restrict__ mask, int64_t hw, int64_t total, int* __restrict__ hist) {
  __shared__ int sh[2 * 3 * 256];
  for (int i = threadIdx.x; i < 2 * 3 * 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>((i / hw) % 3);
    const __half r = result[i];
    atomicAdd(&sh[c * 256 + quant8(__half2float(r))], 1);
    atomicAdd(&sh[768 + c * 256 + quant8(__half2float(composite(source[i], r, mask[i])))], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 3 * 256; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// _match_cumulative_cdf (match_histograms.py:12-37), unsigned branch, one thread per (channel, source value):
//   src_quantiles = cumsum(bincount(src)) / src.size ; tmpl values = bins with a non-zero count, their quantiles likewise
//   lut[v] = uint8(np.interp(src_quantiles[v], tmpl_quantiles, tmpl_values))       (float64 -> uint8 truncates)
__global__ void __launch_bounds__(768) outpaint_lut_kernel(const int* __restrict__ hist, double size, uint8_t* __restrict__ lut) {
  __shared__ double sq[3][256];       // source quantiles
  __shared__ double tq[3][256];       // template quantiles (compacted)
  __shared__ int tv[3][256];          // template values (compacted)
  __shared__ int tn[3];
  const int c = threadIdx.x / 256, v = threadIdx.x % 256;
  if (v == 0) {
    long long acc = 0;
    for (int k = 0; k < 256; ++k) {
      acc += hist[c * 256 + k];
      sq[c][k] = static_cast<double>(acc) / size;
    }
    acc = 0;
    int n = 0;
    for (int k = 0; k < 256; ++k) {
      const int cnt = hist[768 + c * 256 + k];
      if (cnt == 0) continue;
      acc += cnt;
      tv[c][n] = k;
      tq[c][n] = static_cast<double>(acc) / size;
      ++n;
    }
    tn[c] = n;
  }
  __syncthreads();
  const int n = tn[c];
  const double x = sq[c][v];
  double y;
  if (n == 0) {
    y = 0.0;
  } else if (x < tq[c][0]) {
    y = tv[c][0];                                      // np.interp: left = fp[0]
  } else if (x > tq[c][n - 1]) {
    y = tv[c][n - 1];                                  // right = fp[-1]
  } else {
    int j = 0;                                         // last j with xp[j] <= x
    for (int k = 0; k < n; ++k)
      if (tq[c][k] <= x) j = k;
    if (j == n - 1 || tq[c][j] == x) {
      y = tv[c][j];
    } else {
      const double slope = static_cast<double>(tv[c][j + 1] - tv[c][j]) / (tq[c][j + 1] - tq[c][j]);
      y = slope * (x - tq[c][j]) + tv[c][j];
    }
  }
  lut[c * 256 + v] = static_cast<uint8_t>(static_cast<int>(y));
}

// matched = lut[quant8(result)] / 255 (float32 -> image dtype); out = source * (1 - outmask) + matched * outmask
__global__ void outpaint_apply_kernel(const __half* __restrict__ result, const __half* __restrict__ source,
                                      const __half* __restrict__ mask, const uint8_t* __restrict__ lut, int64_t hw,
                                      int64_t total, __half* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>((i / hw) % 3);
  const float matched = static_cast<float>(lut[c * 256 + quant8(__half2float(result[i]))]) / 255.0f;
  out[i] = composite(source[i], __float2half_rn(matched), mask[i]);
}

size_t outpaint_scratch_bytes() { return 2 * 3 * 256 * sizeof(int) + 3 * 256; }

int outpaint_match_histograms(const __half* result, const __half* source, const __half* mask, int B, int64_t hw, __half* out,
                              void* scratch, cudaStream_t st) {
  GYRE_REQUIRE(result && source && mask && out && scratch && B > 0 && hw > 0, "outpaint_match_histograms: bad arguments");
  const int64_t total = static_cast<int64_t>(B) * 3 * hw;
  int* hist = static_cast<int*>(scratch);
  uint8_t* lut = reinterpret_cast<uint8_t*>(hist + 2 * 3 * 256);
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  GYRE_CHECK_CUDA(cudaMemsetAsync(hist, 0, 2 * 3 * 256 * sizeof(int), st));
  unsigned grid = blocks_for(total, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  outpaint_hist_kernel<<<grid, 256, 0, st>>>(result, source, mask, hw, total, hist);
  outpaint_lut_kernel<<<1, 768, 0, st>>>(hist, static_cast<double>(static_cast<int64_t>(B) * hw), lut);
  outpaint_apply_kernel<<<blocks_for(total, 256), 256, 0, st>>>(result, source, mask, lut, hw, total, out);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------ safety checker front end
// CLIPFeatureExtractor as the reference calls it on the decoded images (unified_pipeline.py:2512-2516; transformers
// ~= 4.28.1 CLIPImageProcessor: resize shortest edge to 224 with PIL bicubic, centre crop 224, x 1/255, (x - mean) / std).
// PIL's 8-bit resample (ImagingResample, Resample.c) is integer arithmetic: per pass the coefficients are rounded to
// 22-bit fixed point, each output is clip8((1 << 21) + sum pixel * k >> 22), horizontal pass first - reproduced here
// exactly (the coefficient tables come from the host, built with PIL's double-precision expressions).
__global__ void resample_u8_kernel(const uint8_t* __restrict__ src, int64_t n_outer, int in_sz, int inner, const int* __restrict__ bounds,
                                   const int* __restrict__ kk, int ksize, int out_sz, uint8_t* __restrict__ dst) {
  // src [n_outer, in_sz, inner] -> dst [n_outer, out_sz, inner]: one pass along the middle dimension
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t total = n_outer * out_sz * inner;
  if (i >= total) return;
  const int c = static_cast<int>(i % inner);
  const int o = static_cast<int>((i / inner) % out_sz);
  const int64_t b = i / (static_cast<int64_t>(inner) * out_sz);
  const int xmin = bounds[2 * o], xcnt = bounds[2 * o + 1];
  const int* k = kk + o * ksize;
  int ss = 1 << 21;
  const uint8_t* s0 = src + (b * in_sz + xmin) * inner + c;
  for (int x = 0; x < xcnt; ++x) ss += static_cast<int>(s0[static_cast<int64_t>(x) * inner]) * k[x];
  ss >>= 22;                                   // arithmetic shift, like C on a signed int
  dst[i] = static_cast<uint8_t>(ss < 0 ? 0 : (ss > 255 ? 255 : ss));
}

int resample_u8(const uint8_t* src, int64_t n_outer, int in_sz, int inner, const int* bounds, const int* kk, int ksize,
                int out_sz, uint8_t* dst, cudaStream_t st) {
  GYRE_REQUIRE(src && dst && bounds && kk && n_outer > 0 && in_sz > 0 && inner > 0 && out_sz > 0 && ksize > 0,
               "resample_u8: bad arguments");
  const int64_t total = n_outer * out_sz * inner;
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  resample_u8_kernel<<<blocks_for(total, 256), 256, 0, st>>>(src, n_outer, in_sz, inner, bounds, kk, ksize, out_sz, dst);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// centre crop + rescale + normalise: u8 NHWC [B, H, W, 3] -> fp16 NCHW [B, 3, S, S]
//   x = float32(double(u8) * (1 / 255));  y = (x - mean[c]) / std[c]  (float32, IEEE division);  out = fp16(y)
__global__ void clip_normalize_kernel(const uint8_t* __restrict__ src, int H, int W, int S, int oy, int ox, float m0, float m1,
                                      float m2, float s0, float s1, float s2, __half* __restrict__ out, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % S);
  const int y = static_cast<int>((i / S) % S);
  const int c = static_cast<int>((i / (static_cast<int64_t>(S) * S)) % 3);
  const int64_t b = i / (3ll * S * S);
  const uint8_t v = src[((b * H + (y + oy)) * W + (x + ox)) * 3 + c];
  const float r = static_cast<float>(static_cast<double>(v) * 0.00392156862745098);
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
  out[i] = __float2half_rn(__fdiv_rn(__fsub_rn(r, mean), sd));
}

int clip_normalize(const uint8_t* src, int B, int H, int W, int S, const float* mean3, const float* std3, __half* out,
                   cudaStream_t st) {
  GYRE_REQUIRE(src && out && mean3 && std3 && B > 0 && H >= S && W >= S && S > 0, "clip_normalize: bad arguments");
  const int64_t total = static_cast<int64_t>(B) * 3 * S * S;
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  clip_normalize_kernel<<<blocks_for(total, 256), 256, 0, st>>>(src, H, W, S, (H - S) / 2, (W - S) / 2, mean3[0], mean3[1],
                                                                mean3[2], std3[0], std3[1], std3[2], out, total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ViT patch embedding input: pixel_values NCHW fp16 [B, 3, S, S] -> A [B * np * np, Kp], k = c * P * P + ky * P + kx
// (the Conv2d(3, C, P, stride P) weight flattened the same way), zero-padded to the pitch Kp
__global__ void patchify_kernel(const __half* __restrict__ x, int S, int P, int Kp, __half* __restrict__ out, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int k = static_cast<int>(i % Kp);
  const int64_t row = i / Kp;
  const int np = S / P;
  const int px = static_cast<int>(row % np), py = static_cast<int>((row / np) % np);
  const int64_t b = row / (static_cast<int64_t>(np) * np);
  if (k >= 3 * P * P) {
    out[i] = __float2half_rn(0.f);
    return;
  }
  const int c = k / (P * P), ky = (k / P) % P, kx = k % P;
  out[i] = x[((b * 3 + c) * S + (py * P + ky)) * S + (px * P + kx)];
}

int patchify(const __half* x, int B, int S, int P, int Kp, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(x && out && B > 0 && S > 0 && P > 0 && S % P == 0 && Kp >= 3 * P * P, "patchify: bad arguments");
  const int64_t total = static_cast<int64_t>(B) * (S / P) * (S / P) * Kp;
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  patchify_kernel<<<blocks_for(total, 256), 256, 0, st>>>(x, S, P, Kp, out, total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// tokens[b, 0] = class_embedding + pos[0]; tokens[b, 1 + i] = patches[b, i] + pos[1 + i]   (CLIPVisionEmbeddings)
__global__ void vision_embed_kernel(const __half* __restrict__ patches, const float* __restrict__ cls, const __half* __restrict__ pos,
                                    int Ntok, int C, __half* __restrict__ out, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const int tkn = static_cast<int>((i / C) % Ntok);
  const int64_t b = i / (static_cast<int64_t>(C) * Ntok);
  const float base = tkn == 0 ? __half2float(__float2half_rn(cls[c])) : __half2float(patches[(b * (Ntok - 1) + (tkn - 1)) * C + c]);
  out[i] = __float2half_rn(base + __half2float(pos[static_cast<int64_t>(tkn) * C + c]));
}

int vision_embed(const __half* patches, const float* cls, const __half* pos, int B, int Ntok, int C, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(patches && cls && pos && out && B > 0 && Ntok > 1 && C > 0, "vision_embed: bad arguments");
  const int64_t total = static_cast<int64_t>(B) * Ntok * C;
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  vision_embed_kernel<<<blocks_for(total, 256), 256, 0, st>>>(patches, cls, pos, Ntok, C, out, total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// cosine_distance(image_embeds, embeds) (gyre/pipeline/safety_checkers.py:8-11): rows normalised with
// F.normalize's x / max(||x||, 1e-12), one CTA per image, scores [B, n_embeds] fp32
__global__ void __launch_bounds__(256) cosine_scores_kernel(const __half* __restrict__ img, int D, const float* __restrict__ emb,
                                                            int n_emb, float* __restrict__ scores) {
  const int b = blockIdx.x;
  __shared__ float red[256];
  __shared__ float inv_norm;
  const __half* x = img + static_cast<int64_t>(b) * D;
  float s = 0.f;
  for (int k = threadIdx.x; k < D; k += 256) {
    const float v = __half2float(x[k]);
    s = fmaf(v, v, s);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) inv_norm = 1.0f / fmaxf(sqrtf(red[0]), 1e-12f);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = warp; e < n_emb; e += 8) {
    const float* t = emb + static_cast<int64_t>(e) * D;
    float dot = 0.f, tn = 0.f;
    for (int k = lane; k < D; k += 32) {
      const float tv = t[k];
      dot = fmaf(__half2float(x[k]), tv, dot);
      tn = fmaf(tv, tv, tn);
    }
    dot = warp_sum(dot);
    tn = warp_sum(tn);
    if (lane == 0) scores[static_cast<int64_t>(b) * n_emb + e] = dot * inv_norm / fmaxf(sqrtf(tn), 1e-12f);
  }
}

int cosine_scores(const __half* img, int B, int D, const float* emb, int n_emb, float* scores, cudaStream_t st) {
  GYRE_REQUIRE(img && emb && scores && B > 0 && D > 0 && n_emb > 0, "cosine_scores: bad arguments");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  cosine_scores_kernel<<<B, 256, 0, st>>>(img, D, emb, n_emb, scores);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gyre
