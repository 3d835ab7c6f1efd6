// K6: ToMe K/V merge (reference: nonfree/tome_memory_efficient_cross_attention.py:28-50 calling
// nonfree/ToMe/tome/merge.py:18-97 `bipartite_soft_matching` and :210-224 `merge_wavg`).
//
//   1. tome_split_normalize   k -> unit-norm even tokens A [B, Na, C] and odd tokens B [B, Nb, C] (fp16)
//   2. scores = A . B^T on the tensor cores (gemm_tc with the ACT_ROWMAX epilogue): the Na x Nb score matrix
//      is never written - each row keeps only its running max / argmax per column block
//   3. tome_plan (one CTA per sample): fold the partial maxima, bitonic-sort the A tokens by best score
//      (descending, ties -> lower token first), take the first r as merge sources, sort the (dst, src) pairs so
//      that every B token gets a contiguous, ordered source list (deterministic summation order)
//   4. tome_apply: out = [A tokens that stay, in sorted order ..., B tokens averaged with their sources ...]
//      for K and V with the same plan; sums in fp32, one rounding to fp16.
#include "common.cuh"
#include "ops.h"

namespace gyre {

struct TomeLayout {
  int Na, Nb, P;
  size_t off_an, off_bn, off_pmax, off_pidx, off_nidx, off_unm, off_src, off_start, off_end, total;
};

static size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

static TomeLayout tome_layout(int B, int N, int C) {
  TomeLayout L;
  L.Na = (N + 1) / 2;
  L.Nb = N / 2;
  L.P = gemm_rowmax_partials(L.Na, L.Nb);
  size_t o = 0;
  L.off_an = o;    o += align256(static_cast<size_t>(B) * L.Na * C * 2);
  L.off_bn = o;    o += align256(static_cast<size_t>(B) * L.Nb * C * 2);
  L.off_pmax = o;  o += align256(static_cast<size_t>(B) * L.Na * L.P * 4);
  L.off_pidx = o;  o += align256(static_cast<size_t>(B) * L.Na * L.P * 4);
  L.off_nidx = o;  o += align256(static_cast<size_t>(B) * L.Na * 4);
  L.off_unm = o;   o += align256(static_cast<size_t>(B) * L.Na * 4);
  L.off_src = o;   o += align256(static_cast<size_t>(B) * L.Na * 4);
  L.off_start = o; o += align256(static_cast<size_t>(B) * (L.Nb + 1) * 4);
  L.off_end = o;   o += align256(static_cast<size_t>(B) * (L.Nb + 1) * 4);
  L.total = o;
  return L;
}

int tome_plan_offsets(int B, int N, int C, size_t* node_idx, size_t* unm_idx, size_t* src_idx) {
  GYRE_REQUIRE(B > 0 && N > 1 && C > 0, "tome: empty problem");
  const TomeLayout L = tome_layout(B, N, C);
  *node_idx = L.off_nidx;
  *unm_idx = L.off_unm;
  *src_idx = L.off_src;
  return 0;
}

int tome_workspace_bytes(int B, int N, int C, size_t* bytes) {
  GYRE_REQUIRE(B > 0 && N > 1 && C > 0, "tome: empty problem");
  *bytes = tome_layout(B, N, C).total;
  return 0;
}

// one warp per token: L2-normalise (fp32) and route to the even / odd set
__global__ void __launch_bounds__(256) tome_split_normalize_kernel(const __half* __restrict__ k, int ld, int N, int C,
                                                                   int Na, int Nb, __half* __restrict__ an,
                                                                   __half* __restrict__ bn) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tok = blockIdx.x * 8 + warp;
  const int b = blockIdx.y;
  if (tok >= N) return;
  const __half* src = k + (static_cast<int64_t>(b) * N + tok) * ld;
  float ss = 0.f;
  for (int c = lane * 8; c < C; c += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + c);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      ss += f.x * f.x + f.y * f.y;
    }
  }
  const float inv = rsqrtf(warp_sum(ss));
  __half* dst = (tok & 1) ? bn + (static_cast<int64_t>(b) * Nb + (tok >> 1)) * C
                          : an + (static_cast<int64_t>(b) * Na + (tok >> 1)) * C;
  for (int c = lane * 8; c < C; c += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(src + c);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    __align__(16) __half2 o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      o[i] = __floats2half2_rn(f.x * inv, f.y * inv);
    }
    *reinterpret_cast<uint4*>(dst + c) = *reinterpret_cast<uint4*>(o);
  }
}

// 64-bit keys, ascending bitonic sort in shared memory (n = power of two)
__device__ void bitonic_sort_u64(unsigned long long* a, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = a[i], y = a[ixj];
          const bool up = (i & k) == 0;
          if ((x > y) == up) {
            a[i] = y;
            a[ixj] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

// order-preserving map float -> uint32 (ascending)
__device__ __forceinline__ unsigned int float_key(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(1024) tome_plan_kernel(const float* __restrict__ pmax, const int* __restrict__ pidx,
                                                         int P, int Na, int Nb, int r, int npow2,
                                                         int* __restrict__ node_idx, int* __restrict__ unm_src,
                                                         int* __restrict__ src_sorted, int* __restrict__ start,
                                                         int* __restrict__ end) {
  extern __shared__ unsigned long long keys[];   // [npow2]
  const int b = blockIdx.x;
  pmax += static_cast<int64_t>(b) * Na * P;
  pidx += static_cast<int64_t>(b) * Na * P;
  node_idx += static_cast<int64_t>(b) * Na;
  unm_src += static_cast<int64_t>(b) * Na;
  src_sorted += static_cast<int64_t>(b) * Na;
  start += static_cast<int64_t>(b) * (Nb + 1);
  end += static_cast<int64_t>(b) * (Nb + 1);
  // 1. fold the per-block maxima: best score and its (lowest) column per A token
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    if (i < Na) {
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int q = 0; q < P; ++q) {
        const float v = pmax[static_cast<int64_t>(i) * P + q];
        const int c = pidx[static_cast<int64_t>(i) * P + q];
        if (v > best || (v == best && c < bi)) {
          best = v;
          bi = c;
        }
      }
      node_idx[i] = bi;
      // descending score, ascending token on ties  ==  ascending (~score_key, token)
      keys[i] = (static_cast<unsigned long long>(~float_key(best)) << 32) | static_cast<unsigned int>(i);
    } else {
      keys[i] = ~0ull;
    }
  }
  for (int j = threadIdx.x; j <= Nb; j += blockDim.x) {
    start[j] = 0;
    end[j] = 0;
  }
  __syncthreads();
  bitonic_sort_u64(keys, npow2);
  // 2. tokens that stay (argsort positions r..Na-1, in that order)
  for (int i = r + threadIdx.x; i < Na; i += blockDim.x) unm_src[i - r] = static_cast<int>(keys[i] & 0xffffffffu);
  __syncthreads();
  // 3. merge edges (dst, src), sorted so that each destination's sources are contiguous and ordered
  unsigned long long tmp[8];
  int cnt = 0;
  for (int t = threadIdx.x; t < npow2 && cnt < 8; t += blockDim.x, ++cnt) {
    if (t < r) {
      const int src = static_cast<int>(keys[t] & 0xffffffffu);
      tmp[cnt] = (static_cast<unsigned long long>(static_cast<unsigned int>(node_idx[src])) << 32) |
                 static_cast<unsigned int>(src);
    } else {
      tmp[cnt] = ~0ull;
    }
  }
  __syncthreads();
  cnt = 0;
  for (int t = threadIdx.x; t < npow2 && cnt < 8; t += blockDim.x, ++cnt) keys[t] = tmp[cnt];
  __syncthreads();
  bitonic_sort_u64(keys, npow2);
  for (int t = threadIdx.x; t < r; t += blockDim.x) {
    const int d = static_cast<int>(keys[t] >> 32);
    src_sorted[t] = static_cast<int>(keys[t] & 0xffffffffu);
    if (t == 0 || static_cast<int>(keys[t - 1] >> 32) != d) start[d] = t;
    if (t == r - 1 || static_cast<int>(keys[t + 1] >> 32) != d) end[d] = t + 1;
  }
}

// one warp per output token, K and V together
__global__ void __launch_bounds__(256) tome_apply_kernel(const __half* __restrict__ k, const __half* __restrict__ v,
                                                         int ld, int N, int C, int Na, int Nb, int r,
                                                         const int* __restrict__ unm_src,
                                                         const int* __restrict__ src_sorted,
                                                         const int* __restrict__ start, const int* __restrict__ end,
                                                         __half* __restrict__ k_out, __half* __restrict__ v_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_out = N - r;
  const int o = blockIdx.x * 8 + warp;
  const int b = blockIdx.y;
  if (o >= n_out) return;
  const int n_unm = Na - r;
  const __half* kb = k + static_cast<int64_t>(b) * N * ld;
  const __half* vb = v + static_cast<int64_t>(b) * N * ld;
  __half* ko = k_out + (static_cast<int64_t>(b) * n_out + o) * C;
  __half* vo = v_out + (static_cast<int64_t>(b) * n_out + o) * C;
  if (o < n_unm) {
    const int tok = 2 * unm_src[static_cast<int64_t>(b) * Na + o];
    for (int c = lane * 8; c < C; c += 256) {
      *reinterpret_cast<uint4*>(ko + c) = *reinterpret_cast<const uint4*>(kb + static_cast<int64_t>(tok) * ld + c);
      *reinterpret_cast<uint4*>(vo + c) = *reinterpret_cast<const uint4*>(vb + static_cast<int64_t>(tok) * ld + c);
    }
    return;
  }
  const int j = o - n_unm;                                  // B token index
  const int s0 = start[static_cast<int64_t>(b) * (Nb + 1) + j];
  const int s1 = end[static_cast<int64_t>(b) * (Nb + 1) + j];
  const int* srcs = src_sorted + static_cast<int64_t>(b) * Na;
  const float inv = 1.0f / static_cast<float>(1 + s1 - s0);  // merge_wavg with size == 1 everywhere: plain mean
  for (int c = lane * 8; c < C; c += 256) {
    float ak[8], av[8];
    {
      const int64_t off = static_cast<int64_t>(2 * j + 1) * ld + c;
      const uint4 uk = *reinterpret_cast<const uint4*>(kb + off);
      const uint4 uv = *reinterpret_cast<const uint4*>(vb + off);
      const __half2* hk = reinterpret_cast<const __half2*>(&uk);
      const __half2* hv = reinterpret_cast<const __half2*>(&uv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 fk = __half22float2(hk[i]), fv = __half22float2(hv[i]);
        ak[2 * i] = fk.x; ak[2 * i + 1] = fk.y;
        av[2 * i] = fv.x; av[2 * i + 1] = fv.y;
      }
    }
    for (int s = s0; s < s1; ++s) {
      const int64_t off = static_cast<int64_t>(2 * srcs[s]) * ld + c;
      const uint4 uk = *reinterpret_cast<const uint4*>(kb + off);
      const uint4 uv = *reinterpret_cast<const uint4*>(vb + off);
      const __half2* hk = reinterpret_cast<const __half2*>(&uk);
      const __half2* hv = reinterpret_cast<const __half2*>(&uv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 fk = __half22float2(hk[i]), fv = __half22float2(hv[i]);
        ak[2 * i] += fk.x; ak[2 * i + 1] += fk.y;
        av[2 * i] += fv.x; av[2 * i + 1] += fv.y;
      }
    }
    __align__(16) __half2 ok[4], ov[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ok[i] = __floats2half2_rn(ak[2 * i] * inv, ak[2 * i + 1] * inv);
      ov[i] = __floats2half2_rn(av[2 * i] * inv, av[2 * i + 1] * inv);
    }
    *reinterpret_cast<uint4*>(ko + c) = *reinterpret_cast<uint4*>(ok);
    *reinterpret_cast<uint4*>(vo + c) = *reinterpret_cast<uint4*>(ov);
  }
}

int tome_merge_kv(const __half* k, const __half* v, int ld, int B, int N, int C, int r, __half* k_out, __half* v_out,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
  GYRE_REQUIRE(B > 0 && N > 1 && C > 0, "tome: empty problem");
  GYRE_REQUIRE(C % 8 == 0 && ld % 8 == 0, "tome: channel count / pitch must be multiples of 8");
  GYRE_REQUIRE(r > 0 && r <= N / 2, "tome: r=%d must be in [1, N/2] (the reference clamps to 50%% of tokens)", r);
  const TomeLayout L = tome_layout(B, N, C);
  GYRE_REQUIRE(workspace != nullptr && workspace_bytes >= L.total, "tome: workspace too small (%zu < %zu)",
               workspace_bytes, L.total);
  GYRE_REQUIRE(L.Na <= 8192, "tome: more than 16384 tokens per sample is not supported");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __half* an = reinterpret_cast<__half*>(ws + L.off_an);
  __half* bn = reinterpret_cast<__half*>(ws + L.off_bn);
  float* pmax = reinterpret_cast<float*>(ws + L.off_pmax);
  int* pidx = reinterpret_cast<int*>(ws + L.off_pidx);
  int* nidx = reinterpret_cast<int*>(ws + L.off_nidx);
  int* unm = reinterpret_cast<int*>(ws + L.off_unm);
  int* src = reinterpret_cast<int*>(ws + L.off_src);
  int* start = reinterpret_cast<int*>(ws + L.off_start);
  int* end = reinterpret_cast<int*>(ws + L.off_end);
  {
    prof::Scope ps(prof::F_TOME, 0.0, 2.0 * 2.0 * B * N * C, st);
    tome_split_normalize_kernel<<<dim3((N + 7) / 8, B), 256, 0, st>>>(k, ld, N, C, L.Na, L.Nb, an, bn);
    GYRE_CHECK_CUDA(cudaGetLastError());
  }
  for (int b = 0; b < B; ++b) {
    Epilogue e;
    e.act = ACT_ROWMAX;
    e.rowmax_val = pmax + static_cast<size_t>(b) * L.Na * L.P;
    e.rowmax_idx = pidx + static_cast<size_t>(b) * L.Na * L.P;
    e.rowmax_ld = L.P;
    GYRE_TRY(gemm_f16(an + static_cast<size_t>(b) * L.Na * C, C, bn + static_cast<size_t>(b) * L.Nb * C, C, L.Na, L.Nb,
                      C, e, st));
  }
  int npow2 = 1;
  while (npow2 < L.Na) npow2 <<= 1;
  {
    prof::Scope ps(prof::F_TOME, 0.0, 0.0, st, 2);
    static bool attr_done = false;
    if (!attr_done) {
      GYRE_CHECK_CUDA(cudaFuncSetAttribute(tome_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
      attr_done = true;
    }
    GYRE_REQUIRE(npow2 <= 8 * 1024, "tome: plan kernel handles up to 16384 tokens per sample (got %d)", N);
    tome_plan_kernel<<<B, 1024, static_cast<size_t>(npow2) * 8, st>>>(pmax, pidx, L.P, L.Na, L.Nb, r, npow2, nidx, unm,
                                                                      src, start, end);
    tome_apply_kernel<<<dim3((N - r + 7) / 8, B), 256, 0, st>>>(k, v, ld, N, C, L.Na, L.Nb, r, unm, src, start, end,
                                                                k_out, v_out);
    GYRE_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace gyre
