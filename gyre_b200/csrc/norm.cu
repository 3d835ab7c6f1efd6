// K2 / K3: HBM-bound normalisation fusions.  GroupNorm(+SiLU) over NHWC fp16 (optionally reading the
// channel-concatenation of two tensors, so the UNet skip-concat is never materialised before a norm),
// LayerNorm over token rows, and the fp32 row softmax used by the VAE's single-head attention.
// All loads/stores are 128-bit; statistics are fp32 with warp-shuffle / shared-memory reductions and a
// Chan-style merge of per-chunk (n, mean, M2) partials (deterministic: no float atomics to global).
#include "common.cuh"
#include "ops.h"

namespace gyre {

constexpr int kGnRows = 64;      // pixels per CTA chunk
constexpr int kMaxGroups = 32;

size_t gn_partials_floats(int B, int HW, int G) {
  const int chunks = (HW + kGnRows - 1) / kGnRows;
  return static_cast<size_t>(B) * chunks * G * 2 + static_cast<size_t>(B) * G * 2;   // partials + (mean, rstd)
}

__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8h(__half* p, const float (&v)[8]) {
  __align__(16) __half2 h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<uint4*>(h);
}

// grid (chunks, B); block = nvec * rpar threads, thread -> (vector v of 8 channels, row lane ty)
__global__ void gn_stats_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2, int HW,
                                int G, int rpar, float* __restrict__ partials) {
  const int C = C1 + C2;
  const int nvec = C >> 3;
  const int cpg = C / G;
  const int v = threadIdx.x % nvec;
  const int ty = threadIdx.x / nvec;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * kGnRows;
  const int row1 = min(row0 + kGnRows, HW);
  const int c0 = v * 8;
  const __half* src;
  int ld, cc;
  if (c0 < C1) { src = x1; ld = C1; cc = c0; } else { src = x2; ld = C2; cc = c0 - C1; }
  src += (static_cast<int64_t>(b) * HW) * ld + cc;

  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
  // 4 independent 16-byte loads in flight per thread
  int r = row0 + ty;
  for (; r + 3 * rpar < row1; r += 4 * rpar) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(src + static_cast<int64_t>(r + k * rpar) * ld);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2* h = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        s[2 * i] += f.x;
        ss[2 * i] += f.x * f.x;
        s[2 * i + 1] += f.y;
        ss[2 * i + 1] += f.y * f.y;
      }
    }
  }
  for (; r < row1; r += rpar) {
    float xv[8];
    load8(src + static_cast<int64_t>(r) * ld, xv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i] += xv[i];
      ss[i] += xv[i] * xv[i];
    }
  }
  // Deterministic fold: per-thread channel sums go to shared memory, then one thread per group adds its
  // channels x row-lanes in a fixed order (no float atomics: results are bit-reproducible, which the
  // reference's batch-independence contract, tests/batch_independance.py, is checked against).
  extern __shared__ float sh[];            // [rpar][C][2]
  float* mine = sh + (static_cast<size_t>(ty) * C + c0) * 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mine[2 * i] = s[i];
    mine[2 * i + 1] = ss[i];
  }
  __syncthreads();
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    float a = 0.f, q = 0.f;
    for (int t = 0; t < rpar; ++t) {
      const float* row = sh + (static_cast<size_t>(t) * C + g * cpg) * 2;
      for (int c = 0; c < cpg; ++c) {
        a += row[2 * c];
        q += row[2 * c + 1];
      }
    }
    float* dst = partials + (static_cast<int64_t>(b) * gridDim.x + blockIdx.x) * (2 * G) + 2 * g;
    dst[0] = a;
    dst[1] = q;
  }
}

// one CTA per (sample, group): merges the per-chunk (sum, sumsq) partials in fp64 -> (mean, rstd)
__global__ void __launch_bounds__(128) gn_finalize_kernel(const float* __restrict__ partials, int chunks, int G, int HW,
                                                          int cpg, float eps, float* __restrict__ stats) {
  const int g = blockIdx.x, b = blockIdx.y;
  double s = 0.0, q = 0.0;
  for (int c = threadIdx.x; c < chunks; c += 128) {
    const float* pp = partials + (static_cast<int64_t>(b) * chunks + c) * (2 * G) + 2 * g;
    s += static_cast<double>(pp[0]);
    q += static_cast<double>(pp[1]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  __shared__ double sh[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[warp] = s; sh[4 + warp] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    s = sh[0] + sh[1] + sh[2] + sh[3];
    q = sh[4] + sh[5] + sh[6] + sh[7];
    const double n = static_cast<double>(HW) * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[(static_cast<int64_t>(b) * G + g) * 2] = static_cast<float>(mean);
    stats[(static_cast<int64_t>(b) * G + g) * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

__global__ void gn_apply_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2, int HW,
                                int G, int rpar, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int silu, const float* __restrict__ stats,
                                __half* __restrict__ out) {
  const int C = C1 + C2;
  const int nvec = C >> 3;
  const int cpg = C / G;
  const int b = blockIdx.y;
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  if (threadIdx.x < G) {
    s_mean[threadIdx.x] = stats[(static_cast<int64_t>(b) * G + threadIdx.x) * 2];
    s_rstd[threadIdx.x] = stats[(static_cast<int64_t>(b) * G + threadIdx.x) * 2 + 1];
  }
  __syncthreads();
  const int v = threadIdx.x % nvec;
  const int ty = threadIdx.x / nvec;
  const int row0 = blockIdx.x * kGnRows;
  const int row1 = min(row0 + kGnRows, HW);
  const int c0 = v * 8;
  const __half* src;
  int ld, cc;
  if (c0 < C1) { src = x1; ld = C1; cc = c0; } else { src = x2; ld = C2; cc = c0 - C1; }
  src += (static_cast<int64_t>(b) * HW) * ld + cc;
  __half* dst = out + (static_cast<int64_t>(b) * HW) * C + c0;
  float sc[8], sf[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (c0 + i) / cpg;
    const float ga = gamma[c0 + i] * s_rstd[g];
    sc[i] = ga;
    sf[i] = beta[c0 + i] - s_mean[g] * ga;
  }
  int r = row0 + ty;
  for (; r + 3 * rpar < row1; r += 4 * rpar) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(src + static_cast<int64_t>(r + k * rpar) * ld);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2* h = reinterpret_cast<const __half2*>(&u[k]);
      float xv[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        xv[2 * i] = f.x;
        xv[2 * i + 1] = f.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = fmaf(xv[i], sc[i], sf[i]);
        if (silu) y = y / (1.0f + __expf(-y));
        xv[i] = y;
      }
      store8h(dst + static_cast<int64_t>(r + k * rpar) * C, xv);
    }
  }
  for (; r < row1; r += rpar) {
    float xv[8];
    load8(src + static_cast<int64_t>(r) * ld, xv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = fmaf(xv[i], sc[i], sf[i]);
      if (silu) y = y / (1.0f + __expf(-y));
      xv[i] = y;
    }
    store8h(dst + static_cast<int64_t>(r) * C, xv);
  }
}

// Small feature maps (UNet levels 2-3): ONE kernel, one CTA per (sample, group).  The group's HW x cpg slab
// (<= 48 KB) is read once into shared memory, reduced in a fixed order (deterministic), normalised and
// written back - a single pass over HBM and a single launch instead of three.
__global__ void __launch_bounds__(256) gn_small_kernel(const __half* __restrict__ x1, int C1,
                                                       const __half* __restrict__ x2, int C2, int HW, int G, float eps,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int silu, __half* __restrict__ out) {
  extern __shared__ __half2 slab[];          // [HW][cpg / 2]
  const int C = C1 + C2;
  const int cpg = C / G;
  const int hp = cpg >> 1;                   // half2 pairs per row (cpg is even)
  const int g = blockIdx.x, b = blockIdx.y;
  const int cbase = g * cpg;
  const int total = HW * hp;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < total; i += 256) {
    const int row = i / hp;
    const int c = cbase + 2 * (i - row * hp);
    const __half* src = c < C1 ? x1 + (static_cast<int64_t>(b) * HW + row) * C1 + c
                               : x2 + (static_cast<int64_t>(b) * HW + row) * C2 + (c - C1);
    const __half2 v = *reinterpret_cast<const __half2*>(src);
    slab[i] = v;
    const float2 f = __half22float2(v);
    s += f.x + f.y;
    ss += f.x * f.x + f.y * f.y;
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  __shared__ float red[16];
  __shared__ float stat[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp] = s;
    red[8 + warp] = ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, q = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += static_cast<double>(red[w]);
      q += static_cast<double>(red[8 + w]);
    }
    const double n = static_cast<double>(HW) * cpg;
    const double mean = a / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = static_cast<float>(mean);
    stat[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1];
  for (int i = threadIdx.x; i < total; i += 256) {
    const int row = i / hp;
    const int cl = 2 * (i - row * hp);
    const int c = cbase + cl;
    const float2 f = __half22float2(slab[i]);
    const float ga0 = gamma[c] * rstd, ga1 = gamma[c + 1] * rstd;
    float y0 = fmaf(f.x - mean, ga0, beta[c]);
    float y1 = fmaf(f.y - mean, ga1, beta[c + 1]);
    if (silu) {
      y0 = y0 / (1.0f + __expf(-y0));
      y1 = y1 / (1.0f + __expf(-y1));
    }
    *reinterpret_cast<__half2*>(out + (static_cast<int64_t>(b) * HW + row) * C + c) = __floats2half2_rn(y0, y1);
  }
}

int groupnorm_nhwc(const __half* x1, int C1, const __half* x2, int C2, int B, int HW, int G, float eps,
                   const float* gamma, const float* beta, bool silu, __half* out, float* partials, cudaStream_t st) {
  const int C = C1 + C2;
  GYRE_REQUIRE(B > 0 && HW > 0 && C > 0, "groupnorm: empty input");
  GYRE_REQUIRE(G > 0 && G <= kMaxGroups && C % G == 0, "groupnorm: C=%d not divisible into %d groups", C, G);
  GYRE_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0, "groupnorm: channel counts must be multiples of 8");
  GYRE_REQUIRE(x2 != nullptr || C2 == 0, "groupnorm: missing second source");
  {
    const int cpg = C / G;
    const size_t slab_bytes = static_cast<size_t>(HW) * cpg * sizeof(__half);
    if ((cpg & 1) == 0 && HW <= 256 && slab_bytes <= 40 * 1024 && (C1 % 2 == 0)) {
      prof::Scope ps(prof::F_GROUPNORM, 0.0, 2.0 * 2.0 * B * HW * C, st, 1);
      gn_small_kernel<<<dim3(G, B), 256, slab_bytes, st>>>(x1, C1, x2, C2, HW, G, eps, gamma, beta, silu ? 1 : 0, out);
      GYRE_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
  }
  const int nvec = C / 8;
  GYRE_REQUIRE(nvec <= 1024, "groupnorm: C=%d too large", C);
  int rpar = 256 / nvec;
  if (rpar < 1) rpar = 1;
  const int threads = nvec * rpar;
  GYRE_REQUIRE(threads >= G, "groupnorm: too few threads for %d groups", G);
  const int chunks = (HW + kGnRows - 1) / kGnRows;
  dim3 grid(chunks, B);
  float* stats = partials + static_cast<size_t>(B) * chunks * G * 2;
  prof::Scope ps(prof::F_GROUPNORM, 0.0, 2.0 * 2.0 * B * HW * C, st, 3);
  gn_stats_kernel<<<grid, threads, static_cast<size_t>(rpar) * C * 2 * sizeof(float), st>>>(x1, C1, x2, C2, HW, G, rpar,
                                                                                          partials);
  gn_finalize_kernel<<<dim3(G, B), 128, 0, st>>>(partials, chunks, G, HW, C / G, eps, stats);
  gn_apply_kernel<<<grid, threads, 0, st>>>(x1, C1, x2, C2, HW, G, rpar, gamma, beta, silu ? 1 : 0, stats, out);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ LayerNorm
// one warp per row, the row lives in registers: NV 16-byte vectors per lane (C <= 256 * NV), sized per call so
// that small C keeps the register count - and therefore the bytes in flight per SM - where HBM needs them
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, int rows, int C, float eps,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __half* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int nvec = C >> 3;
  const __half* src = x + static_cast<int64_t>(row) * C;
  uint4 u[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = lane + 32 * k;
    if (vi < nvec) u[k] = *reinterpret_cast<const uint4*>(src + vi * 8);
  }
  float v[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = lane + 32 * k;
    if (vi < nvec) {
      const __half2* h = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        v[k][2 * i] = f.x;
        v[k][2 * i + 1] = f.y;
        sum += f.x + f.y;
      }
    }
  }
  const float mean = warp_sum(sum) / C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = lane + 32 * k;
    if (vi < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[k][i] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / C + eps);
  __half* dst = out + static_cast<int64_t>(row) * C;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = lane + 32 * k;
    if (vi < nvec) {
      float y[8];
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = (v[k][i] - mean) * rstd * gg[i] + bb[i];
      store8h(dst + vi * 8, y);
    }
  }
}

int layernorm_rows(const __half* x, int rows, int C, float eps, const float* gamma, const float* beta, __half* out,
                   cudaStream_t st) {
  GYRE_REQUIRE(rows > 0 && C > 0, "layernorm: empty input");
  GYRE_REQUIRE(C % 8 == 0 && C <= 2048, "layernorm: C=%d must be a multiple of 8 and <= 2048", C);
  prof::Scope ps(prof::F_LAYERNORM, 0.0, 2.0 * 2.0 * rows * C, st);
  const int nv = (C + 255) / 256;
  const unsigned grid = (rows + 7) / 8;
  switch (nv) {
    case 1: layernorm_kernel<1><<<grid, 256, 0, st>>>(x, rows, C, eps, gamma, beta, out); break;
    case 2: layernorm_kernel<2><<<grid, 256, 0, st>>>(x, rows, C, eps, gamma, beta, out); break;
    case 3: layernorm_kernel<3><<<grid, 256, 0, st>>>(x, rows, C, eps, gamma, beta, out); break;
    case 4: layernorm_kernel<4><<<grid, 256, 0, st>>>(x, rows, C, eps, gamma, beta, out); break;
    case 5: layernorm_kernel<5><<<grid, 256, 0, st>>>(x, rows, C, eps, gamma, beta, out); break;
    case 6: layernorm_kernel<6><<<grid, 256, 0, st>>>(x, rows, C, eps, gamma, beta, out); break;
    default: layernorm_kernel<8><<<grid, 256, 0, st>>>(x, rows, C, eps, gamma, beta, out); break;
  }
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ row softmax (fp32 -> fp16)
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, int n, float scale,
                                                           __half* __restrict__ p, int ldp) {
  const int64_t row = blockIdx.x;
  const float* src = s + row * n;
  __shared__ float red[8];
  __shared__ float bcast;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) m = fmaxf(m, src[i]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = red[0];
    for (int i = 1; i < 8; ++i) mm = fmaxf(mm, red[i]);
    bcast = mm;
  }
  __syncthreads();
  m = bcast;
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) sum += __expf((src[i] - m) * scale);
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bcast = 1.0f / t;
  }
  __syncthreads();
  const float inv = bcast;
  __half* dst = p + row * ldp;
  for (int i = threadIdx.x; i < n; i += 256) dst[i] = __float2half_rn(__expf((src[i] - m) * scale) * inv);
}

int softmax_rows_f32(const float* s, int rows, int n, float scale, __half* p, int ldp, cudaStream_t st) {
  GYRE_REQUIRE(rows > 0 && n > 0 && ldp >= n, "softmax: bad shape");
  prof::Scope ps(prof::F_SOFTMAX, 0.0, static_cast<double>(rows) * n * 6.0, st);
  softmax_rows_kernel<<<rows, 256, 0, st>>>(s, n, scale, p, ldp);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gyre
