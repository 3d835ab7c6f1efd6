// K2 / K3: HBM-bound normalisation fusions.  GroupNorm(+SiLU) over NHWC fp16 (optionally reading the
// channel-concatenation of two tensors, so the UNet skip-concat is never materialised before a norm),
// LayerNorm over token rows, and the fp32 row softmax used by the VAE's single-head attention.
// All loads/stores are 128-bit; statistics are fp32 with warp-shuffle / shared-memory reductions and a
// Chan-style merge of per-chunk (n, mean, M2) partials (deterministic: no float atomics to global).
#include <type_traits>

#include "common.cuh"
#include "ops.h"

namespace gyre {

constexpr int kMaxGroups = 32;
constexpr int kGnMaxChunks = 256;   // upper bound of per-sample partials (sizes the scratch)
constexpr int kGnUnroll = 8;        // independent 16-byte loads in flight per thread

// Rows (pixels) per CTA: a function of HW ONLY, so the summation order - and with it every bit of the result -
// does not depend on the batch size (tests/batch_independance.py contract of the reference).
static inline int gn_rows_per_cta(int HW) {
  int max_chunks = tunable(TUNE_GN_CHUNKS);   // per-sample partials the apply kernel folds in its prologue
  if (max_chunks < 8) max_chunks = 8;
  if (max_chunks > kGnMaxChunks) max_chunks = kGnMaxChunks;
  int rows = (HW + max_chunks - 1) / max_chunks;
  if (rows < 16) rows = 16;
  return (rows + 7) & ~7;
}

size_t gn_partials_floats(int B, int HW, int G) {
  (void)HW;
  return static_cast<size_t>(B) * kGnMaxChunks * G * 2 + static_cast<size_t>(B) * G * 2;
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

struct GnArgs {
  const __half* x1;
  const __half* x2;
  int C1, C2, HW, G, rpar, rows_per_cta;
};

// grid (chunks, B); block = nvec * rpar threads, thread -> (vector v of 8 channels, row lane ty).
// Pass 1: per-chunk (sum, sum of squares) of every group -> partials[b][chunk][G][2].
__global__ void gn_stats_kernel(const GnArgs a, float* __restrict__ partials) {
  pdl_wait();
  pdl_trigger();
  const int C = a.C1 + a.C2;
  const int nvec = C >> 3;
  const int cpg = C / a.G;
  const int v = threadIdx.x % nvec;
  const int ty = threadIdx.x / nvec;
  const int b = blockIdx.y;
  const int row0 = blockIdx.x * a.rows_per_cta;
  const int row1 = min(row0 + a.rows_per_cta, a.HW);
  const int c0 = v * 8;
  const __half* src;
  int ld, cc;
  if (c0 < a.C1) { src = a.x1; ld = a.C1; cc = c0; } else { src = a.x2; ld = a.C2; cc = c0 - a.C1; }
  src += (static_cast<int64_t>(b) * a.HW) * ld + cc;

  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
  const int rstep = a.rpar;
  for (int base = row0 + ty; base < row1; base += kGnUnroll * rstep) {
    uint4 u[kGnUnroll];
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const int r = base + k * rstep;
      u[k] = r < row1 ? *reinterpret_cast<const uint4*>(src + static_cast<int64_t>(r) * ld) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const __half2* h = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        s[2 * i] += f.x;
        ss[2 * i] = fmaf(f.x, f.x, ss[2 * i]);
        s[2 * i + 1] += f.y;
        ss[2 * i + 1] = fmaf(f.y, f.y, ss[2 * i + 1]);
      }
    }
  }
  // Deterministic fold: per-thread channel sums go to shared memory, then one thread per group adds its
  // channels x row-lanes in a fixed order (no float atomics: results are bit-reproducible).
  extern __shared__ float sh[];            // [rpar][C][2]
  float* mine = sh + (static_cast<size_t>(ty) * C + c0) * 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mine[2 * i] = s[i];
    mine[2 * i + 1] = ss[i];
  }
  __syncthreads();
  // four threads per group, each a fixed quarter of the (row-lane, channel) entries, then a fixed-order combine
  if (threadIdx.x < 4 * a.G) {
    const int g = threadIdx.x >> 2;
    const int k = threadIdx.x & 3;
    const int n = a.rpar * cpg;
    float acc = 0.f, q = 0.f;
    for (int e = k; e < n; e += 4) {
      const int t = e / cpg;
      const int c = e - t * cpg;
      const float2 v2 = *reinterpret_cast<const float2*>(sh + (static_cast<size_t>(t) * C + g * cpg + c) * 2);
      acc += v2.x;
      q += v2.y;
    }
    const unsigned m = __activemask();   // whole quads: 4 * G threads take this branch
    acc += __shfl_xor_sync(m, acc, 1);
    q += __shfl_xor_sync(m, q, 1);
    acc += __shfl_xor_sync(m, acc, 2);
    q += __shfl_xor_sync(m, q, 2);
    if (k == 0) {
      float* dst = partials + (static_cast<int64_t>(b) * gridDim.x + blockIdx.x) * (2 * a.G) + 2 * g;
      dst[0] = acc;
      dst[1] = q;
    }
  }
}

// Pass 2: every CTA folds its sample's per-chunk partials (fp64, fixed order) into (mean, rstd) - no separate
// finalize launch - then normalises (+SiLU) its chunk.  The first batch of rows is requested BEFORE the fold so
// the statistics latency hides under the loads.  Chunks and samples are walked in the REVERSE order of pass 1:
// what pass 1 touched last is still in L2.
// nparts: partials per sample behind `partials` (the statistics pass leaves one per chunk; a producing convolution one
// per output tile); ready != 0: `partials` holds finished (mean, rstd) pairs [B][G][2] instead (gn_fold_kernel).
__global__ void gn_apply_kernel(const GnArgs a, const float* __restrict__ gamma, const float* __restrict__ beta,
                                int silu, float eps, const float* __restrict__ partials, int nparts, int ready,
                                __half* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int C = a.C1 + a.C2;
  const int nvec = C >> 3;
  const int cpg = C / a.G;
  const int chunks = gridDim.x;
  const int chunk = chunks - 1 - blockIdx.x;
  const int b = gridDim.y - 1 - blockIdx.y;
  const int v = threadIdx.x % nvec;
  const int ty = threadIdx.x / nvec;
  const int row0 = chunk * a.rows_per_cta;
  const int row1 = min(row0 + a.rows_per_cta, a.HW);
  const int c0 = v * 8;
  const __half* src;
  int ld, cc;
  if (c0 < a.C1) { src = a.x1; ld = a.C1; cc = c0; } else { src = a.x2; ld = a.C2; cc = c0 - a.C1; }
  src += (static_cast<int64_t>(b) * a.HW) * ld + cc;
  __half* dst = out + (static_cast<int64_t>(b) * a.HW) * C + c0;
  const int rstep = a.rpar;

  uint4 u[kGnUnroll];
  int base = row0 + ty;
#pragma unroll
  for (int k = 0; k < kGnUnroll; ++k) {
    const int r = base + k * rstep;
    u[k] = r < row1 ? __ldcs(reinterpret_cast<const uint4*>(src + static_cast<int64_t>(r) * ld)) : make_uint4(0, 0, 0, 0);
  }

  __shared__ double sh_s[8][kMaxGroups], sh_q[8][kMaxGroups];
  __shared__ float s_mean[kMaxGroups], s_rstd[kMaxGroups];
  int parts = static_cast<int>(blockDim.x) / a.G;
  if (parts > 8) parts = 8;
  if (ready) {
    if (static_cast<int>(threadIdx.x) < a.G) {
      const float2 ms = *reinterpret_cast<const float2*>(partials + (static_cast<int64_t>(b) * a.G + threadIdx.x) * 2);
      s_mean[threadIdx.x] = ms.x;
      s_rstd[threadIdx.x] = ms.y;
    }
  } else if (static_cast<int>(threadIdx.x) < parts * a.G) {
    const int g = threadIdx.x % a.G;
    const int part = threadIdx.x / a.G;
    const float* pp = partials + static_cast<int64_t>(b) * nparts * (2 * a.G) + 2 * g;
    double s = 0.0, q = 0.0;
    for (int c = part; c < nparts; c += parts) {
      const float2 t = *reinterpret_cast<const float2*>(pp + static_cast<int64_t>(c) * (2 * a.G));
      s += static_cast<double>(t.x);
      q += static_cast<double>(t.y);
    }
    sh_s[part][g] = s;
    sh_q[part][g] = q;
  }
  __syncthreads();
  if (!ready && static_cast<int>(threadIdx.x) < a.G) {
    const int g = threadIdx.x;
    double s = 0.0, q = 0.0;
    for (int t = 0; t < parts; ++t) {
      s += sh_s[t][g];
      q += sh_q[t][g];
    }
    const double n = static_cast<double>(a.HW) * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = static_cast<float>(mean);
    s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();

  float sc[8], sf[8];
  {
    // 16-byte loads: a scalar load per channel would cost one 32-byte sector per lane per instruction
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (c0 + i) / cpg;
      const float ga = gg[i] * s_rstd[g];
      sc[i] = ga;
      sf[i] = bb[i] - s_mean[g] * ga;
    }
  }
  while (true) {
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const int r = base + k * rstep;
      if (r < row1) {
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
          if (silu) y = silu_fast(y);
          xv[i] = y;
        }
        store8h(dst + static_cast<int64_t>(r) * C, xv);
      }
    }
    base += kGnUnroll * rstep;
    if (base >= row1) break;
#pragma unroll
    for (int k = 0; k < kGnUnroll; ++k) {
      const int r = base + k * rstep;
      u[k] = r < row1 ? __ldcs(reinterpret_cast<const uint4*>(src + static_cast<int64_t>(r) * ld)) : make_uint4(0, 0, 0, 0);
    }
  }
}

// Small feature maps (UNet levels 2-3): ONE kernel, one CTA per (sample, group).  The group's HW x cpg slab
// (<= 40 KB) is read once into shared memory with the widest vector the group width allows (VEC halfs per
// load), reduced in a fixed order (deterministic), normalised and written back - a single pass over HBM and a
// single launch.
template <int VEC>
__global__ void __launch_bounds__(256) gn_small_kernel(const __half* __restrict__ x1, int C1,
                                                       const __half* __restrict__ x2, int C2, int HW, int G, float eps,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int silu, __half* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) uint8_t slab_raw[];
  typedef typename std::conditional<VEC == 8, uint4, typename std::conditional<VEC == 4, uint2, uint32_t>::type>::type vec_t;
  vec_t* slab = reinterpret_cast<vec_t*>(slab_raw);   // [HW][cpg / VEC]
  const int C = C1 + C2;
  const int cpg = C / G;
  const int vp = cpg / VEC;                  // vectors per row
  const int g = blockIdx.x, b = blockIdx.y;
  const int cbase = g * cpg;
  const int total = HW * vp;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < total; i += 256) {
    const int row = i / vp;
    const int c = cbase + VEC * (i - row * vp);
    const __half* src = c < C1 ? x1 + (static_cast<int64_t>(b) * HW + row) * C1 + c
                               : x2 + (static_cast<int64_t>(b) * HW + row) * C2 + (c - C1);
    const vec_t v = *reinterpret_cast<const vec_t*>(src);
    slab[i] = v;
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int k = 0; k < VEC / 2; ++k) {
      const float2 f = __half22float2(h[k]);
      s += f.x + f.y;
      ss += f.x * f.x + f.y * f.y;
    }
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
    double acc = 0.0, q = 0.0;
    for (int w = 0; w < 8; ++w) {
      acc += static_cast<double>(red[w]);
      q += static_cast<double>(red[8 + w]);
    }
    const double n = static_cast<double>(HW) * cpg;
    const double mean = acc / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[0] = static_cast<float>(mean);
    stat[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const float mean = stat[0], rstd = stat[1];
  for (int i = threadIdx.x; i < total; i += 256) {
    const int row = i / vp;
    const int c = cbase + VEC * (i - row * vp);
    const vec_t v = slab[i];
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    vec_t o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < VEC / 2; ++k) {
      const float2 f = __half22float2(h[k]);
      const float ga0 = gamma[c + 2 * k] * rstd, ga1 = gamma[c + 2 * k + 1] * rstd;
      float y0 = fmaf(f.x - mean, ga0, beta[c + 2 * k]);
      float y1 = fmaf(f.y - mean, ga1, beta[c + 2 * k + 1]);
      if (silu) {
        y0 = silu_fast(y0);
        y1 = silu_fast(y1);
      }
      oh[k] = __floats2half2_rn(y0, y1);
    }
    *reinterpret_cast<vec_t*>(out + (static_cast<int64_t>(b) * HW + row) * C + c) = o;
  }
}

// Many partials per sample (VAE-sized maps: one per 128-pixel conv tile): one CTA per (group, sample) folds them in a
// fixed order (per-thread strided sums, then a fixed tree in fp64) into (mean, rstd) -> stats [B][G][2].
__global__ void __launch_bounds__(256) gn_fold_kernel(const float* __restrict__ partials, int nparts, int G, double inv_n,
                                                      float eps, float* __restrict__ stats) {
  pdl_wait();
  pdl_trigger();
  const int g = blockIdx.x, b = blockIdx.y;
  const float* pp = partials + static_cast<int64_t>(b) * nparts * (2 * G) + 2 * g;
  double s = 0.0, q = 0.0;
  for (int c = threadIdx.x; c < nparts; c += 256) {
    const float2 t = *reinterpret_cast<const float2*>(pp + static_cast<int64_t>(c) * (2 * G));
    s += static_cast<double>(t.x);
    q += static_cast<double>(t.y);
  }
  __shared__ double sh_s[256], sh_q[256];
  sh_s[threadIdx.x] = s;
  sh_q[threadIdx.x] = q;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (static_cast<int>(threadIdx.x) < w) {
      sh_s[threadIdx.x] += sh_s[threadIdx.x + w];
      sh_q[threadIdx.x] += sh_q[threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = sh_s[0] * inv_n;
    double var = sh_q[0] * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    float* dst = stats + (static_cast<int64_t>(b) * G + g) * 2;
    dst[0] = static_cast<float>(mean);
    dst[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
}

static bool gn_small_path(int C, int HW, int G) {
  const int cpg = C / G;
  const size_t slab_bytes = static_cast<size_t>(HW) * cpg * sizeof(__half);
  return (cpg & 1) == 0 && HW <= 256 && slab_bytes <= 40 * 1024;
}

// geometry of the two-pass kernels for a shape (shared by groupnorm_nhwc and groupnorm_nhwc_pre)
static int gn_two_pass_args(const __half* x1, int C1, const __half* x2, int C2, int HW, int G, GnArgs* a, int* threads) {
  const int C = C1 + C2;
  const int nvec = C / 8;
  GYRE_REQUIRE(nvec <= 1024, "groupnorm: C=%d too large", C);
  int gn_threads = tunable(TUNE_GN_THREADS);          // CTA size target (a function of nothing but the tunable)
  if (gn_threads < 64 || gn_threads > 1024) gn_threads = 256;
  int rpar = gn_threads / nvec;
  if (rpar < 1) rpar = 1;
  *threads = nvec * rpar;
  GYRE_REQUIRE(*threads >= 4 * G, "groupnorm: too few threads for %d groups", G);
  a->x1 = x1;
  a->x2 = x2;
  a->C1 = C1;
  a->C2 = C2;
  a->HW = HW;
  a->G = G;
  a->rpar = rpar;
  a->rows_per_cta = gn_rows_per_cta(HW);
  return 0;
}

bool groupnorm_pre_ok(int C, int HW, int G) {
  return G > 0 && G <= kMaxGroups && C % G == 0 && C % 8 == 0 && !gn_small_path(C, HW, G);
}

int groupnorm_nhwc_pre(const __half* x, int C, int B, int HW, int G, float eps, const float* gamma, const float* beta,
                       bool silu, __half* out, const float* pre, int nparts, float* stats, cudaStream_t st) {
  GYRE_REQUIRE(B > 0 && HW > 0 && C > 0 && nparts > 0 && pre != nullptr, "groupnorm_pre: empty input");
  GYRE_REQUIRE(groupnorm_pre_ok(C, HW, G), "groupnorm_pre: C=%d HW=%d G=%d takes the single-pass kernel", C, HW, G);
  GnArgs a;
  int threads = 0;
  GYRE_TRY(gn_two_pass_args(x, C, nullptr, 0, HW, G, &a, &threads));
  const int chunks = (HW + a.rows_per_cta - 1) / a.rows_per_cta;
  // one pass over the tensor: read + write
  prof::Scope ps(prof::F_GROUPNORM, 0.0, 2.0 * 2.0 * B * HW * C, st, nparts > 64 ? 2 : 1);
  int ready = 0;
  if (nparts > 64) {
    GYRE_REQUIRE(stats != nullptr, "groupnorm_pre: %d partials per sample need the statistics scratch", nparts);
    const double inv_n = 1.0 / (static_cast<double>(HW) * (C / G));
    GYRE_TRY(launch_kernel(gn_fold_kernel, dim3(G, B), dim3(256), 0, st, pre, nparts, G, inv_n, eps, stats));
    pre = stats;
    ready = 1;
  }
  return launch_kernel(gn_apply_kernel, dim3(chunks, B), dim3(threads), 0, st, a, gamma, beta, silu ? 1 : 0, eps, pre,
                       nparts, ready, out);
}

int groupnorm_nhwc(const __half* x1, int C1, const __half* x2, int C2, int B, int HW, int G, float eps,
                   const float* gamma, const float* beta, bool silu, __half* out, float* partials, cudaStream_t st) {
  const int C = C1 + C2;
  GYRE_REQUIRE(B > 0 && HW > 0 && C > 0, "groupnorm: empty input");
  GYRE_REQUIRE(G > 0 && G <= kMaxGroups && C % G == 0, "groupnorm: C=%d not divisible into %d groups", C, G);
  GYRE_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0, "groupnorm: channel counts must be multiples of 8");
  GYRE_REQUIRE(x2 != nullptr || C2 == 0, "groupnorm: missing second source");
  if (gn_small_path(C, HW, G)) {
    const int cpg = C / G;
    const size_t slab_bytes = static_cast<size_t>(HW) * cpg * sizeof(__half);
    prof::Scope ps(prof::F_GROUPNORM, 0.0, 2.0 * 2.0 * B * HW * C, st, 1);
    // a vector must not straddle the x1 / x2 boundary or a group boundary: C1, cpg multiples of VEC
    const int vec = (cpg % 8 == 0) ? 8 : (cpg % 4 == 0 ? 4 : 2);
    const dim3 grid(G, B);
    if (vec == 8)
      return launch_kernel(gn_small_kernel<8>, grid, dim3(256), slab_bytes, st, x1, C1, x2, C2, HW, G, eps, gamma, beta,
                           silu ? 1 : 0, out);
    if (vec == 4)
      return launch_kernel(gn_small_kernel<4>, grid, dim3(256), slab_bytes, st, x1, C1, x2, C2, HW, G, eps, gamma, beta,
                           silu ? 1 : 0, out);
    return launch_kernel(gn_small_kernel<2>, grid, dim3(256), slab_bytes, st, x1, C1, x2, C2, HW, G, eps, gamma, beta,
                         silu ? 1 : 0, out);
  }
  GnArgs a;
  int threads = 0;
  GYRE_TRY(gn_two_pass_args(x1, C1, x2, C2, HW, G, &a, &threads));
  const int rpar = a.rpar;
  const int chunks = (HW + a.rows_per_cta - 1) / a.rows_per_cta;
  dim3 grid(chunks, B);
  prof::Scope ps(prof::F_GROUPNORM, 0.0, 2.0 * 2.0 * B * HW * C, st, 2);
  const int phase = tunable(TUNE_GN_PHASE);   // measurement only: 1 = statistics pass alone, 2 = apply pass alone
  if (phase != 2)
    GYRE_TRY(launch_kernel(gn_stats_kernel, grid, dim3(threads), static_cast<size_t>(rpar) * C * 2 * sizeof(float), st, a,
                           partials));
  if (phase != 1)
    GYRE_TRY(launch_kernel(gn_apply_kernel, grid, dim3(threads), 0, st, a, gamma, beta, silu ? 1 : 0, eps,
                           static_cast<const float*>(partials), chunks, 0, out));
  return 0;
}

// ------------------------------------------------------------------------------------ LayerNorm
// one warp per row, the row lives in registers: NV 16-byte vectors per lane (C <= 256 * NV), sized per call so
// that small C keeps the register count - and therefore the bytes in flight per SM - where HBM needs them
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, int rows, int C, float eps,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __half* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  constexpr int R = 1;   // rows per warp (2 measured slower on B200: 23.9 vs 21.9 us at 65536 x 320)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + warp) * R;
  if (row0 >= rows) return;
  const int nvec = C >> 3;
  uint4 u[R][NV];
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int row = min(row0 + rr, rows - 1);
    const __half* src = x + static_cast<int64_t>(row) * C;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int vi = lane + 32 * k;
      if (vi < nvec) u[rr][k] = *reinterpret_cast<const uint4*>(src + vi * 8);
    }
  }
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int row = row0 + rr;
    if (row >= rows) break;
    float v[NV][8];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int vi = lane + 32 * k;
      if (vi < nvec) {
        const __half2* h = reinterpret_cast<const __half2*>(&u[rr][k]);
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
}

// Sub-warp rows: LPR lanes per row, 32 / LPR rows per warp.  At C = 320 a full warp per row keeps 640 B in flight per
// warp and leaves three quarters of the second load idle; 8 lanes per row (5 vectors each) keep 2.5 KB in flight per
// warp with the same registers and need 3 shuffle levels instead of 5.  The reduction tree depends on C only, never on
// the batch, so results stay batch-invariant.
template <int LPR, int NV>
__global__ void __launch_bounds__(256) layernorm_sub_kernel(const __half* __restrict__ x, int rows, int C, float eps,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, __half* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  constexpr int RPW = 32 / LPR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int row = (blockIdx.x * 8 + warp) * RPW + lane / LPR;
  const bool live = row < rows;
  const int nvec = C >> 3;
  const __half* src = x + static_cast<int64_t>(live ? row : rows - 1) * C;
  uint4 u[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = sub + LPR * k;
    if (vi < nvec) u[k] = *reinterpret_cast<const uint4*>(src + vi * 8);
  }
  float v[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = sub + LPR * k;
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
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = sub + LPR * k;
    if (vi < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[k][i] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
  if (!live) return;
  __half* dst = out + static_cast<int64_t>(row) * C;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = sub + LPR * k;
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

template <int LPR>
static int launch_ln_sub(int nv, unsigned grid, cudaStream_t st, const __half* x, int rows, int C, float eps,
                         const float* gamma, const float* beta, __half* out) {
  switch (nv) {
    case 1: return launch_kernel(layernorm_sub_kernel<LPR, 1>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out);
    case 2: return launch_kernel(layernorm_sub_kernel<LPR, 2>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out);
    case 3: return launch_kernel(layernorm_sub_kernel<LPR, 3>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out);
    case 4: return launch_kernel(layernorm_sub_kernel<LPR, 4>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out);
    default: return launch_kernel(layernorm_sub_kernel<LPR, 5>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out);
  }
}

int layernorm_rows(const __half* x, int rows, int C, float eps, const float* gamma, const float* beta, __half* out,
                   cudaStream_t st) {
  GYRE_REQUIRE(rows > 0 && C > 0, "layernorm: empty input");
  GYRE_REQUIRE(C % 8 == 0 && C <= 2048, "layernorm: C=%d must be a multiple of 8 and <= 2048", C);
  prof::Scope ps(prof::F_LAYERNORM, 0.0, 2.0 * 2.0 * rows * C, st);
  // narrow rows: several rows per warp (tunable LN_SUB, default on)
  // (measured: 65536 x 320: 22.5 -> 18.6 us; at C = 640 two rows per warp is no faster than one: 11.3 vs 10.9 us)
  if ((tunable(TUNE_LN_SUB) == 1 && C <= 320) || (tunable(TUNE_LN_SUB) == 2 && C <= 640)) {
    const int nvec = C >> 3;
    const int lpr = nvec <= 40 ? 8 : 16;
    const int nvs = (nvec + lpr - 1) / lpr;               // <= 5
    const int rows_cta = 8 * (32 / lpr);
    const unsigned g = (rows + rows_cta - 1) / rows_cta;
    if (lpr == 8) GYRE_TRY(launch_ln_sub<8>(nvs, g, st, x, rows, C, eps, gamma, beta, out));
    else GYRE_TRY(launch_ln_sub<16>(nvs, g, st, x, rows, C, eps, gamma, beta, out));
    GYRE_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const int nv = (C + 255) / 256;
  const int rows_per_cta = 8;
  const unsigned grid = (rows + rows_per_cta - 1) / rows_per_cta;
  switch (nv) {
    case 1: GYRE_TRY(launch_kernel(layernorm_kernel<1>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out)); break;
    case 2: GYRE_TRY(launch_kernel(layernorm_kernel<2>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out)); break;
    case 3: GYRE_TRY(launch_kernel(layernorm_kernel<3>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out)); break;
    case 4: GYRE_TRY(launch_kernel(layernorm_kernel<4>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out)); break;
    case 5: GYRE_TRY(launch_kernel(layernorm_kernel<5>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out)); break;
    case 6: GYRE_TRY(launch_kernel(layernorm_kernel<6>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out)); break;
    default: GYRE_TRY(launch_kernel(layernorm_kernel<8>, dim3(grid), dim3(256), 0, st, x, rows, C, eps, gamma, beta, out)); break;
  }
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------ LayerNorm folded into GEMMs
// (mean, rstd) per row from the per-(n-tile, column-half) partial sums the producing GEMM's epilogue left
// (Epilogue::rowstat_out).  Fixed summation order; the final E[x^2] - mean^2 in fp64.
__global__ void __launch_bounds__(256) ln_finalize_kernel(const float2* __restrict__ parts, int nparts, int M, float inv_c,
                                                          float eps, float2* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * 256 + threadIdx.x;
  if (row >= M) return;
  double s = 0.0, q = 0.0;
  for (int p = 0; p < nparts; ++p) {
    const float2 v = parts[static_cast<int64_t>(p) * M + row];
    s += static_cast<double>(v.x);
    q += static_cast<double>(v.y);
  }
  const double mean = s * static_cast<double>(inv_c);
  double var = q * static_cast<double>(inv_c) - mean * mean;
  if (var < 0.0) var = 0.0;
  out[row] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
}

int ln_finalize_rows(const float2* parts, int nparts, int M, int C, float eps, float2* out, cudaStream_t st) {
  GYRE_REQUIRE(parts && out && nparts > 0 && M > 0 && C > 0, "ln_finalize: bad arguments");
  prof::Scope ps(prof::F_LAYERNORM, 0.0, 8.0 * M * (nparts + 1), st);
  GYRE_TRY(launch_kernel(ln_finalize_kernel, dim3((M + 255) / 256), dim3(256), 0, st, parts, nparts, M, 1.0f / C, eps, out));
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Weight preparation (once per load): one warp per output row n of W [N, K].
__global__ void __launch_bounds__(256) ln_fold_kernel(const __half* __restrict__ w, int N, int K,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ bias, __half* __restrict__ w_out,
                                                      float* __restrict__ colsum, float* __restrict__ lnbias) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float cs = 0.f, bw = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = __half2float(w[static_cast<int64_t>(n) * K + k]);
    const __half ws = __float2half_rn(wv * gamma[k]);
    w_out[static_cast<int64_t>(n) * K + k] = ws;
    cs += __half2float(ws);             // the column sum of what the tensor core will actually multiply by
    bw = fmaf(beta[k], wv, bw);
  }
  cs = warp_sum(cs);
  bw = warp_sum(bw);
  if (lane == 0) {
    colsum[n] = cs;
    lnbias[n] = (bias ? bias[n] : 0.f) + bw;
  }
}

int ln_fold_linear(const __half* w, int N, int K, const float* gamma, const float* beta, const float* bias, __half* w_out,
                   float* colsum, float* lnbias, cudaStream_t st) {
  GYRE_REQUIRE(w && gamma && beta && w_out && colsum && lnbias && N > 0 && K > 0, "ln_fold: bad arguments");
  ln_fold_kernel<<<(N + 7) / 8, 256, 0, st>>>(w, N, K, gamma, beta, bias, w_out, colsum, lnbias);
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
