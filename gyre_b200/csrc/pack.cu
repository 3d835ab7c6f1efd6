// Weight packing (runs once per parameter at load time): diffusers layouts -> the K-major fp16 layouts
// the tcgen05 kernels consume.
#include "common.cuh"
#include "ops.h"

namespace gyre {

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// [Cout, Cin, 3, 3] -> [Cout, 9, cin_pad] (tap = kh*3+kw), zero padded channels
template <typename T>
__global__ void pack_conv3x3_kernel(const T* __restrict__ w, int Cin, int cin_pad, int64_t total,
                                    __half* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % cin_pad);
  const int64_t r = i / cin_pad;
  const int tap = static_cast<int>(r % 9);
  const int64_t o = r / 9;
  float v = 0.f;
  if (c < Cin) v = to_f32(w[(o * Cin + c) * 9 + tap]);
  out[i] = __float2half_rn(v);
}

size_t conv3x3_packed_elems(int Cin, int Cout) {
  const int cin_pad = (Cin + 63) / 64 * 64;
  return static_cast<size_t>(Cout) * 9 * cin_pad;
}

int pack_conv3x3(const void* w, int dtype, int Cin, int Cout, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(Cin > 0 && Cout > 0, "pack_conv3x3: empty");
  const int cin_pad = (Cin + 63) / 64 * 64;
  const int64_t total = static_cast<int64_t>(Cout) * 9 * cin_pad;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (dtype == 0)
    pack_conv3x3_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(w), Cin, cin_pad, total, out);
  else if (dtype == 1)
    pack_conv3x3_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(w), Cin, cin_pad, total, out);
  else
    GYRE_REQUIRE(false, "pack_conv3x3: dtype %d", dtype);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// nearest-2x upsample followed by a 3x3 conv == four 2x2 convs on the LOW-RES input, one per output phase
// (dy, dx) = (Y & 1, X & 1): the 3 taps of a row collapse onto 2 source pixels, so their weights add
//   dy = 0: source row y-1 <- ky{0},   source row y   <- ky{1,2}
//   dy = 1: source row y   <- ky{0,1}, source row y+1 <- ky{2}            (same for columns)
// 2.25x fewer MACs than convolving the upsampled tensor, which is never written.  Zero padding of the
// upsampled image coincides with zero padding of the source (TMA out-of-bounds fill), so borders are exact.
// out: [4 phases][Cout][4 taps = th*2+tw][cin_pad]; sums are formed in fp32 and rounded to fp16 once.
template <typename T>
__global__ void pack_upconv3x3_kernel(const T* __restrict__ w, int Cin, int Cout, int cin_pad, int64_t total,
                                      __half* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % cin_pad);
  int64_t r = i / cin_pad;
  const int tap = static_cast<int>(r % 4);
  r /= 4;
  const int o = static_cast<int>(r % Cout);
  const int phase = static_cast<int>(r / Cout);
  const int dy = phase >> 1, dx = phase & 1;
  const int th = tap >> 1, tw = tap & 1;
  // [k_lo, k_hi] of the 3x3 taps that land on source offset t for phase d
  const int ky0 = dy == 0 ? (th == 0 ? 0 : 1) : (th == 0 ? 0 : 2);
  const int ky1 = dy == 0 ? (th == 0 ? 0 : 2) : (th == 0 ? 1 : 2);
  const int kx0 = dx == 0 ? (tw == 0 ? 0 : 1) : (tw == 0 ? 0 : 2);
  const int kx1 = dx == 0 ? (tw == 0 ? 0 : 2) : (tw == 0 ? 1 : 2);
  float v = 0.f;
  if (c < Cin) {
    const T* base = w + (static_cast<int64_t>(o) * Cin + c) * 9;
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) v += to_f32(base[ky * 3 + kx]);
  }
  out[i] = __float2half_rn(v);
}

size_t upconv3x3_packed_elems(int Cin, int Cout) {
  const int cin_pad = (Cin + 63) / 64 * 64;
  return static_cast<size_t>(4) * Cout * 4 * cin_pad;
}

int pack_upconv3x3(const void* w, int dtype, int Cin, int Cout, __half* out, cudaStream_t st) {
  GYRE_REQUIRE(Cin > 0 && Cout > 0, "pack_upconv3x3: empty");
  const int cin_pad = (Cin + 63) / 64 * 64;
  const int64_t total = static_cast<int64_t>(upconv3x3_packed_elems(Cin, Cout));
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (dtype == 0)
    pack_upconv3x3_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(w), Cin, Cout, cin_pad, total, out);
  else if (dtype == 1)
    pack_upconv3x3_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(w), Cin, Cout, cin_pad, total, out);
  else
    GYRE_REQUIRE(false, "pack_upconv3x3: dtype %d", dtype);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// generic cast / copy with row pitch: src [rows, cols] dtype -> dst fp16 or fp32 [rows, ldd]
template <typename T, typename U>
__global__ void cast_rows_kernel(const T* __restrict__ src, int64_t rows, int cols, U* __restrict__ dst, int ldd) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= rows * cols) return;
  const int c = static_cast<int>(i % cols);
  const int64_t r = i / cols;
  const float v = to_f32(src[i]);
  if constexpr (sizeof(U) == 2) dst[r * ldd + c] = __float2half_rn(v);
  else dst[r * ldd + c] = v;
}

int cast_to_f16(const void* src, int dtype, int64_t rows, int cols, __half* dst, int ldd, cudaStream_t st) {
  GYRE_REQUIRE(rows > 0 && cols > 0, "cast: empty");
  const unsigned blocks = static_cast<unsigned>((rows * cols + 255) / 256);
  if (dtype == 0)
    cast_rows_kernel<__half, __half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(src), rows, cols, dst, ldd);
  else if (dtype == 1)
    cast_rows_kernel<float, __half><<<blocks, 256, 0, st>>>(static_cast<const float*>(src), rows, cols, dst, ldd);
  else
    GYRE_REQUIRE(false, "cast: dtype %d", dtype);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int cast_to_f32(const void* src, int dtype, int64_t n, float* dst, cudaStream_t st) {
  GYRE_REQUIRE(n > 0, "cast: empty");
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  if (dtype == 0)
    cast_rows_kernel<__half, float><<<blocks, 256, 0, st>>>(static_cast<const __half*>(src), n, 1, dst, 1);
  else if (dtype == 1)
    cast_rows_kernel<float, float><<<blocks, 256, 0, st>>>(static_cast<const float*>(src), n, 1, dst, 1);
  else
    GYRE_REQUIRE(false, "cast: dtype %d", dtype);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// GEGLU: W [2F, K] = [value rows ; gate rows]  ->  per 256-row tile: 128 value rows then their 128 gate rows
template <typename T>
__global__ void pack_geglu_kernel(const T* __restrict__ w, int F, int K, int64_t total, __half* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int k = static_cast<int>(i % K);
  const int64_t r = i / K;           // packed row
  const int tile = static_cast<int>(r / 256);
  const int j = static_cast<int>(r % 256);
  const int64_t src_row = (j < 128) ? (static_cast<int64_t>(tile) * 128 + j) : (F + static_cast<int64_t>(tile) * 128 + (j - 128));
  out[i] = __float2half_rn(to_f32(w[src_row * K + k]));
}
template <typename T>
__global__ void pack_geglu_bias_kernel(const T* __restrict__ b, int F, float* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= 2 * F) return;
  const int tile = r / 256, j = r % 256;
  const int src = (j < 128) ? tile * 128 + j : F + tile * 128 + (j - 128);
  out[r] = to_f32(b[src]);
}

int pack_geglu(const void* w, int dtype, int F, int K, const void* bias, int bias_dtype, __half* wp, float* bias_p,
               cudaStream_t st) {
  GYRE_REQUIRE(F > 0 && F % 128 == 0 && K > 0, "pack_geglu: F=%d must be a multiple of 128", F);
  if (w) {
    const int64_t total = static_cast<int64_t>(2) * F * K;
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
    if (dtype == 0)
      pack_geglu_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(w), F, K, total, wp);
    else if (dtype == 1)
      pack_geglu_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(w), F, K, total, wp);
    else
      GYRE_REQUIRE(false, "pack_geglu: dtype %d", dtype);
  }
  if (bias) {
    const unsigned blocks = (2 * F + 255) / 256;
    if (bias_dtype == 0)
      pack_geglu_bias_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(bias), F, bias_p);
    else if (bias_dtype == 1)
      pack_geglu_bias_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(bias), F, bias_p);
    else
      GYRE_REQUIRE(false, "pack_geglu: bias dtype %d", bias_dtype);
  }
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gyre
