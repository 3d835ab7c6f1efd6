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

}  // namespace gyre
