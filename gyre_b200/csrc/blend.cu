// Latent-space blending of the hires-fix and graft scheduler-UNet wrappers
// (reference: gyre/pipeline/unet/hires_fix.py:45-92 scale_into, :142-205 HiresUnetWrapper.__call__;
//  gyre/pipeline/unet/graft.py:31-48 GraftUnets.__call__; the resize is the vendored ResizeRight with a 4-tap lanczos2
//  window, gyre/src/ResizeRight/resize_right.py:70-118,216-250).
//
// One kernel does what the reference spends ~20 small torch launches per step on:
//   separable 4x4-tap resample of `src` (the per-dimension tap tables are built on the host with the reference's own
//   fp32 expressions, so the weights are bit-identical) -> placement into the target frame (centre offset / centre crop,
//   replicate padding or a cloned background) -> `where(rand >= p, A, B)` against the partner tensor -> insertion into
//   a zero frame.  Latents are a few hundred KB: the only cost that matters is the launch count.
#include "common.cuh"
#include "ops.h"

namespace gyre {

static inline unsigned blocks_for(int64_t n, int threads) {
  return static_cast<unsigned>((n + threads - 1) / threads);
}

struct ResampleArgs {
  const float* src;      // [BC, SH, SW]
  const int* ty_idx;     // [RH, 4] source rows (already clamped: replicate padding of the source)
  const float* ty_w;     // [RH, 4]
  const int* tx_idx;     // [RW, 4]
  const float* tx_w;     // [RW, 4]
  int SH, SW, RH, RW;    // source and resized sizes (RH == SH && ty_idx == nullptr: no resampling along H; same for W)
  int TH, TW;            // target frame of scale_into
  int offy, offx;        // position of the resized image in the target frame (negative: centre crop)
  int mode;              // 0: replicate padding outside the resized image ("pad"); 1: background `bg` ("clone")
  const float* bg;       // [BC, TH, TW] (mode 1)
  const float* other;    // [BC, TH, TW] select partner, or nullptr: no select
  const float* rnd;      // [BC, TH, TW] uniform map (select)
  float p;
  int resampled_if_ge;   // select: rand >= p ? resampled : other   (else rand >= p ? other : resampled)
  float* out;            // [BC, FH, FW]
  int FH, FW, oy, ox;    // output frame; the target frame sits at (oy, ox), zero elsewhere
};

__device__ __forceinline__ float resample_at(const ResampleArgs& a, const float* __restrict__ s, int ry, int rx) {
  // H pass first, then W (ResizeRight walks the dims in ascending scale order; equal scales keep H before W):
  // out = sum_j wx[j] * (sum_i wy[i] * src[iy[i], ix[j]]), products rounded before the sequential adds like
  // (neighbors * weights).sum(1)
  int iy[4], ix[4];
  float wy[4], wx[4];
  int ny = 4, nx = 4;
  if (a.ty_idx) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      iy[i] = a.ty_idx[ry * 4 + i];
      wy[i] = a.ty_w[ry * 4 + i];
    }
  } else {
    ny = 1;
    iy[0] = ry;
    wy[0] = 1.f;
  }
  if (a.tx_idx) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ix[j] = a.tx_idx[rx * 4 + j];
      wx[j] = a.tx_w[rx * 4 + j];
    }
  } else {
    nx = 1;
    ix[0] = rx;
    wx[0] = 1.f;
  }
  float acc = 0.f;
  for (int j = 0; j < nx; ++j) {
    float col;
    if (ny == 1) {
      col = s[static_cast<int64_t>(iy[0]) * a.SW + ix[j]];
    } else {
      col = __fmul_rn(wy[0], s[static_cast<int64_t>(iy[0]) * a.SW + ix[j]]);
#pragma unroll
      for (int i = 1; i < 4; ++i) col = __fadd_rn(col, __fmul_rn(wy[i], s[static_cast<int64_t>(iy[i]) * a.SW + ix[j]]));
    }
    if (nx == 1) return col;
    const float t = __fmul_rn(wx[j], col);
    acc = j == 0 ? t : __fadd_rn(acc, t);
  }
  return acc;
}

__global__ void resample_select_kernel(ResampleArgs a, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int fx = static_cast<int>(i % a.FW);
  const int fy = static_cast<int>((i / a.FW) % a.FH);
  const int64_t bc = i / (static_cast<int64_t>(a.FW) * a.FH);
  const int ty = fy - a.oy, tx = fx - a.ox;
  if (ty < 0 || ty >= a.TH || tx < 0 || tx >= a.TW) {
    a.out[i] = 0.f;
    return;
  }
  const int64_t t = (bc * a.TH + ty) * a.TW + tx;
  int ry = ty - a.offy, rx = tx - a.offx;
  const bool inside = ry >= 0 && ry < a.RH && rx >= 0 && rx < a.RW;
  float v;
  if (!inside && a.mode == 1) {
    v = a.bg[t];
  } else {
    ry = min(max(ry, 0), a.RH - 1);
    rx = min(max(rx, 0), a.RW - 1);
    v = resample_at(a, a.src + bc * static_cast<int64_t>(a.SH) * a.SW, ry, rx);
  }
  if (a.other) {
    const bool ge = a.rnd[t] >= a.p;
    v = (ge == (a.resampled_if_ge != 0)) ? v : a.other[t];
  }
  a.out[i] = v;
}

int resample_select(const float* src, int BC, int SH, int SW, const int* ty_idx, const float* ty_w, int RH,
                    const int* tx_idx, const float* tx_w, int RW, int TH, int TW, int offy, int offx, int mode,
                    const float* bg, const float* other, const float* rnd, float p, int resampled_if_ge, float* out,
                    int FH, int FW, int oy, int ox, cudaStream_t st) {
  GYRE_REQUIRE(src && out && BC > 0 && SH > 0 && SW > 0 && RH > 0 && RW > 0 && TH > 0 && TW > 0 && FH > 0 && FW > 0,
               "resample_select: bad arguments");
  GYRE_REQUIRE((ty_idx == nullptr) == (ty_w == nullptr) && (tx_idx == nullptr) == (tx_w == nullptr),
               "resample_select: a tap table needs both indices and weights");
  GYRE_REQUIRE(ty_idx != nullptr || RH == SH, "resample_select: no H taps given but the height changes");
  GYRE_REQUIRE(tx_idx != nullptr || RW == SW, "resample_select: no W taps given but the width changes");
  GYRE_REQUIRE(mode == 0 || (mode == 1 && bg != nullptr), "resample_select: mode 1 needs a background");
  GYRE_REQUIRE((other == nullptr) == (rnd == nullptr), "resample_select: the select needs both the partner and the random map");
  GYRE_REQUIRE(oy >= 0 && ox >= 0 && oy + TH <= FH && ox + TW <= FW, "resample_select: target frame outside the output frame");
  ResampleArgs a{src, ty_idx, ty_w, tx_idx, tx_w, SH, SW, RH, RW, TH, TW, offy, offx, mode, bg, other, rnd, p,
                 resampled_if_ge, out, FH, FW, oy, ox};
  const int64_t total = static_cast<int64_t>(BC) * FH * FW;
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  resample_select_kernel<<<blocks_for(total, 256), 256, 0, st>>>(a, total);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// GraftUnets: out = where(rand >= p, a, b)
__global__ void rand_select_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ rnd,
                                   float p, int64_t n, float* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  out[i] = rnd[i] >= p ? a[i] : b[i];
}

int rand_select(const float* a, const float* b, const float* rnd, float p, int64_t n, float* out, cudaStream_t st) {
  GYRE_REQUIRE(a && b && rnd && out && n > 0, "rand_select: bad arguments");
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  rand_select_kernel<<<blocks_for(n, 256), 256, 0, st>>>(a, b, rnd, p, n, out);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// One pass of a separable resample with host-built tap tables (gyre/images.py:324-340 `resize` = ResizeRight with lanczos3,
// reflect padding folded into the indices, antialiasing = a window stretched by 1 / scale: up to a few hundred taps):
// src [n_outer, in_size, inner] -> dst [n_outer, out_size, inner] along the middle dimension, products rounded before they
// are added like `(neighbors * weights).sum(1)`, optional clamp to [0, 1].
__global__ void resample_f32_kernel(const float* __restrict__ src, int64_t n_outer, int in_sz, int inner,
                                    const int* __restrict__ idx, const float* __restrict__ w, int ksize, int out_sz, int clamp01,
                                    float* __restrict__ dst) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t total = n_outer * out_sz * inner;
  if (i >= total) return;
  const int c = static_cast<int>(i % inner);
  const int o = static_cast<int>((i / inner) % out_sz);
  const int64_t b = i / (static_cast<int64_t>(inner) * out_sz);
  const float* s0 = src + b * in_sz * inner + c;
  const int* ii = idx + static_cast<int64_t>(o) * ksize;
  const float* ww = w + static_cast<int64_t>(o) * ksize;
  float acc = 0.f;
  for (int k = 0; k < ksize; ++k) acc = __fadd_rn(acc, __fmul_rn(s0[static_cast<int64_t>(ii[k]) * inner], ww[k]));
  if (clamp01) acc = fminf(fmaxf(acc, 0.f), 1.f);
  dst[i] = acc;
}

int resample_f32(const float* src, int64_t n_outer, int in_sz, int inner, const int* idx, const float* w, int ksize, int out_sz,
                 int clamp01, float* dst, cudaStream_t st) {
  GYRE_REQUIRE(src && dst && idx && w && n_outer > 0 && in_sz > 0 && inner > 0 && out_sz > 0 && ksize > 0,
               "resample_f32: bad arguments");
  const int64_t total = n_outer * out_sz * inner;
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  resample_f32_kernel<<<blocks_for(total, 256), 256, 0, st>>>(src, n_outer, in_sz, inner, idx, w, ksize, out_sz, clamp01, dst);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gyre
