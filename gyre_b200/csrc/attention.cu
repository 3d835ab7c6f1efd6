// K5: flash-style attention on tcgen05 for sm_100a.
//
// One CTA owns 128 query rows of one (batch, head).  Q, K and V are read IN PLACE from the token-major
// projection outputs ([B, N, ld], head h at columns h*d) through 4-D TMA tensor maps, so no head-split /
// transpose pass exists: K tiles land as K-major operands, V tiles as MN-major operands (keys are the
// MMA K dimension, d is contiguous), both with the 128B swizzle; the head dimension is zero-padded to the
// next 64 by the TMA unit's out-of-bounds fill (d = 40/80/160 in SD1.x).
//
//   warp 0      TMA producer (Q once, then K/V tiles through a STAGES-deep full/empty ring)
//   warp 1      TMEM owner + tcgen05.mma issuer:  S = Q K^T  (TMEM cols [0,128)),  O += P V (cols [128, ..))
//   warps 2..5  softmax: one query row per thread (tcgen05.ld 32x32b), online max/sum in fp32 with exp2,
//               P written as fp16 into a swizzled K-major smem tile, lazy rescale of O in TMEM
//               (only when the running max grows by more than 2^8), final O / l -> global.
#include "common.cuh"
#include "ops.h"

namespace gyre {

constexpr int kAttThreads = 192;
constexpr int kBQ = 128;
constexpr int kBKeys = 128;
constexpr int kChunkBytes = 128 * 128;   // one 64-wide (128 B) column chunk of a 128-row tile

struct AttnParams {
  int Nq, Nk, heads, d;
  int n_kv_tiles;
  int ksteps_qk;     // ceil(d / 16)
  int npv;           // round_up(d, 16): accumulator columns of O
  int q_tiles;
  int splits, tiles_per_cta;   // short-key kernel: CTAs per (batch, head) and 128-row query tiles per CTA
  float scale_log2;  // scale * log2(e)
  __half* out;
  int ldo;
  long long* trace;  // measurement only: clock64 stamps of CTA 0 (gyre_b200_debug_attention_trace)
  int trace_cap;
};

template <int DCH>
struct AttnCfg {
  static constexpr int STAGES = DCH >= 3 ? 1 : 2;   // d > 128: one K/V stage keeps the CTA under 227 KB
  static constexpr int Q_BYTES = DCH * kChunkBytes;
  static constexpr int KV_BYTES = DCH * kChunkBytes;          // per operand per stage
  static constexpr int P_BYTES = 2 * kChunkBytes;
  static constexpr int SMEM = Q_BYTES + P_BYTES + STAGES * 2 * KV_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = (128 + DCH * 64 <= 256) ? 256 : 512;
};

template <int DCH>
__global__ void __launch_bounds__(kAttThreads) attention_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                const __grid_constant__ CUtensorMap tmK,
                                                                const __grid_constant__ CUtensorMap tmV,
                                                                const AttnParams p) {
  using Cfg = AttnCfg<DCH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + Cfg::Q_BYTES;
  uint8_t* sK = sP + Cfg::P_BYTES;
  uint8_t* sV = sK + Cfg::STAGES * Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + Cfg::STAGES * Cfg::KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + Cfg::STAGES;
  uint64_t* s_full = kv_empty + Cfg::STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int q0 = q_tile * kBQ;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int c = 0; c < DCH; ++c) tma_load_4d(sQ + c * kChunkBytes, &tmQ, q_full, c * 64, q0, h, b);
      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int s = j % Cfg::STAGES;
        const uint32_t ph = (j / Cfg::STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * Cfg::KV_BYTES);
#pragma unroll
        for (int c = 0; c < DCH; ++c) {
          tma_load_4d(sK + s * Cfg::KV_BYTES + c * kChunkBytes, &tmK, &kv_full[s], c * 64, j * kBKeys, h, b);
          tma_load_4d(sV + s * Cfg::KV_BYTES + c * kChunkBytes, &tmV, &kv_full[s], c * 64, j * kBKeys, h, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      mbar_wait(q_full, 0);
      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int s = j % Cfg::STAGES;
        const uint32_t ph = (j / Cfg::STAGES) & 1;
        const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
        const int n_s = (nk_tile + 15) & ~15;
        mbar_wait(&kv_full[s], ph);
        tc_fence_after();
        {
          const uint32_t idesc = umma_idesc_f16(kBQ, n_s);
          const uint32_t qa = smem_u32(sQ);
          const uint32_t ka = smem_u32(sK + s * Cfg::KV_BYTES);
          for (int ks = 0; ks < p.ksteps_qk; ++ks) {
            const uint32_t off = (ks >> 2) * kChunkBytes + (ks & 3) * 32;
            umma_f16_ss(tmem_S, umma_desc_kmajor_sw128(qa + off), umma_desc_kmajor_sw128(ka + off), idesc,
                        ks != 0 ? 1u : 0u);
          }
        }
        umma_commit(s_full);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        {
          const uint32_t idesc = umma_idesc_f16(kBQ, p.npv, /*b_mn_major=*/1);
          const uint32_t pa = smem_u32(sP);
          const uint32_t va = smem_u32(sV + s * Cfg::KV_BYTES);
          const int ksteps = n_s >> 4;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da = umma_desc_kmajor_sw128(pa + (ks >> 2) * kChunkBytes + (ks & 3) * 32);
            const uint64_t db = umma_desc_mnmajor_sw128(va + ks * 2048, kChunkBytes, 1024);
            umma_f16_ss(tmem_O, da, db, idesc, (j | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&kv_empty[s]);
        umma_commit(pv_done);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float c = p.scale_log2;
    float m_ref = -INFINITY;
    float l = 0.f;
    uint8_t* prow = sP + r * 128;
    const int rx = r & 7;

    for (int j = 0; j < p.n_kv_tiles; ++j) {
      const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
      const int n_s = (nk_tile + 15) & ~15;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max over the valid keys of this tile
      float mt = -INFINITY;
      for (int c0 = 0; c0 < n_s; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_S + lane_off + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < nk_tile) mt = fmaxf(mt, __uint_as_float(v[i]));
      }
      if (j == 0) {
        m_ref = mt;
      } else {
        // P and O are free once PV(j-1) retired (it always has by now: the tensor pipe runs in order)
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
        const bool grow = (mt - m_ref) * c > 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = fmaxf(m_ref, mt);
          const float alpha = ex2_approx((m_ref - m_new) * c);
          m_ref = m_new;
          l *= alpha;
          int c0 = 0;
          for (; c0 + 32 <= p.npv; c0 += 32) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_O + lane_off + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x32(tmem_O + lane_off + c0, o);
          }
          if (c0 < p.npv) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(tmem_O + lane_off + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x16(tmem_O + lane_off + c0, o);
          }
          tmem_st_wait();
        }
      }
      // pass 2: p = exp2((s - m_ref) * c) -> fp16, swizzled K-major store
      const float mc = m_ref * c;
      for (int c0 = 0; c0 < n_s; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_S + lane_off + c0, v);
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (c0 + i < nk_tile) ? ex2_approx(fmaf(__uint_as_float(v[i]), c, -mc)) : 0.f;
          float p1 = (c0 + i + 1 < nk_tile) ? ex2_approx(fmaf(__uint_as_float(v[i + 1]), c, -mc)) : 0.f;
          const __half2 hh = __floats2half2_rn(p0, p1);
          // the sum uses the rounded values the tensor core will actually multiply
          const float2 back = __half22float2(hh);
          l += back.x + back.y;
          packed[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        uint8_t* pchunk = prow + (c0 >> 6) * kChunkBytes;
        const int u0 = (c0 & 63) >> 3;                   // first 16-byte unit of this 32-column group
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 w = make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
          *reinterpret_cast<uint4*>(pchunk + (((u0 + u) ^ rx) << 4)) = w;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> out[b, q0 + r, h*d + :]
    mbar_wait(pv_done, (p.n_kv_tiles - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    const bool valid = q0 + r < p.Nq;
    __half* dst = p.out + (static_cast<int64_t>(b) * p.Nq + q0 + r) * p.ldo + h * p.d;
    for (int c0 = 0; c0 < p.npv; c0 += 16) {
      uint32_t o[16];
      tmem_ld_32x32b_x16(tmem_O + lane_off + c0, o);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = c0 + g * 8;
          if (col < p.d) {    // d % 8 == 0
            __align__(16) __half2 hh[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              hh[i] = __floats2half2_rn(__uint_as_float(o[g * 8 + 2 * i]) * inv,
                                        __uint_as_float(o[g * 8 + 2 * i + 1]) * inv);
            *reinterpret_cast<uint4*>(dst + col) = *reinterpret_cast<uint4*>(hh);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}


// ------------------------------------------------------------------------------------------ d <= 64: ping-pong
// Two 128-row query tiles per CTA, each with its own softmax warpgroup, S buffer and O accumulator in TMEM.
// The single MMA thread interleaves them (S0, S1, then per tile: PV_g, S_g(next)), so one group's QK^T / PV
// runs on the tensor pipe while the other group is in its exp2 phase (the SFU is the bound at d = 40), and
// every K/V tile is loaded once for 256 queries.  Scores are read from TMEM once and kept in registers.
constexpr int kAtt2Threads = 352;   // TMA warp, S-MMA warp, 2 x 4 softmax warps, PV-MMA warp
constexpr int kXAttThreads = 320;   // TMA warp, MMA warp, 2 x 4 softmax warps

// DCH = 64-wide chunks of the head dimension (1: d <= 64, 2: d <= 128).  With DCH = 2 the O accumulators take
// 2 x 128 TMEM columns, so P cannot have columns of its own: it is written IN PLACE over the scores (packed fp16
// in the first 64 columns of the group's S region) and P V reads it from there - which rules out issuing the
// next S early (no AV_SPLIT), but also removes P from shared memory, making room for a second K/V stage.
template <int DCH, bool P_IN_SMEM>
struct Att2Cfg {
  static constexpr int STAGES = DCH == 1 ? 3 : 2;
  static constexpr int Q_BYTES = 2 * DCH * kChunkBytes;
  static constexpr int P_BYTES = P_IN_SMEM ? 2 * 2 * kChunkBytes : 0;
  static constexpr int KV_BYTES = DCH * kChunkBytes;           // per operand per stage
  static constexpr int SMEM = Q_BYTES + P_BYTES + STAGES * 2 * KV_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;
};

// Softmax variants (template parameter V of attention2_kernel; bit flags):
//   AV_STAGGER  the first S tile of group 1 is issued only after group 0 is half way through its first exp
//               phase, so the two groups' exp phases alternate on the MUFU instead of colliding
//   AV_PACKED   scale / sum with packed fp32x2 instructions (FFMA2 / FADD2: two elements per issue slot)
//   AV_POLY25 / AV_POLY50  every 4th / 2nd pair of scores takes exp2 on the FMA pipe (Cody-Waite split +
//               degree-3 minimax polynomial, rel err 7.5e-5 < fp16 rounding of P) instead of the MUFU
//   AV_F16EXP   ex2.approx.f16x2 (arguments rounded to fp16: measurement variant only)
//   AV_NOEXP    diagnostic: no exponential at all (timing floor of everything that is not the MUFU)
//   AV_SPLIT    S = Q K^T of tile j+1 is issued (by its own warp) as soon as the group has pulled tile j's scores
//               out of TMEM, i.e. it runs UNDER the exp phase of tile j instead of after P V of tile j; a second
//               issuing warp feeds P V.  Takes the MMA round trip out of the per-tile dependency chain.
//   AV_PTMEM    P stays in tensor memory: the softmax warps write fp16 P with tcgen05.st and P V takes its A operand
//               from TMEM - P never crosses shared memory (no 64 KB/tile-pair of stores + 64 KB of operand reads)
//   AV_LAZYMAX  optimistic softmax: a full tile is exponentiated against the RUNNING max while its own max is folded
//               on the side (no max pass in front of the exp phase, second half of the scores still in flight
//               from TMEM); only if the tile turns out to raise the max by more than 2^8 is it redone (rare)
enum : int { AV_STAGGER = 1, AV_PACKED = 2, AV_POLY25 = 4, AV_POLY50 = 8, AV_F16EXP = 16, AV_NOEXP = 32, AV_SPLIT = 64,
             AV_PTMEM = 128, AV_LAZYMAX = 256, AV_TURNS = 512 };
//   AV_TURNS    the two groups take turns in the exponential phase (an mbarrier hand-off each way): one group's exps
//               run at the full MUFU rate while the other loads / maxes / hands over, instead of both sharing the
//               MUFU and then leaving it idle together

__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  // 2^x for x <= ~8: n = round(x) via the 1.5*2^23 magic add, f = x - n in [-0.5, 0.5], 2^f by a degree-3
  // polynomial, exponent inserted with an integer add.  Arguments are clamped at -125 (result ~ 2^-125 ~ 0).
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 r = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(r, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __ffma2_rn(n, make_float2(-1.0f, -1.0f), x);
  float2 q = __ffma2_rn(make_float2(0.05517084f, 0.05517084f), f, make_float2(0.24260935f, 0.24260935f));
  q = __ffma2_rn(q, f, make_float2(0.69326096f, 0.69326096f));
  q = __ffma2_rn(q, f, make_float2(0.99992818f, 0.99992818f));
  float2 o;
  o.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(r.x) << 23));
  o.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(r.y) << 23));
  return o;
}

// p = exp2(s * c - mc) for 32 scores of one row -> fp16; row sum in lsum.  Destination: a swizzled K-major
// shared-memory tile (4 x 16 bytes) or, with AV_PTMEM, 16 packed columns of the group's P region in TMEM.
template <int V>
__device__ __forceinline__ void softmax_exp32(const uint32_t (&v)[32], float c, float mc, float2& lsum,
                                              uint32_t (&w)[16]) {
#pragma unroll
  for (int pair = 0; pair < 16; ++pair) {
    const int i = 2 * pair;
    float2 x;
    if constexpr ((V & AV_PACKED) != 0) {
      x = __ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), make_float2(c, c),
                     make_float2(-mc, -mc));
    } else {
      x.x = fmaf(__uint_as_float(v[i]), c, -mc);
      x.y = fmaf(__uint_as_float(v[i + 1]), c, -mc);
    }
    const bool poly = ((V & AV_POLY50) != 0 && (pair & 1) == 1) || ((V & AV_POLY25) != 0 && (pair & 3) == 3);
    float2 e;
    if constexpr ((V & AV_F16EXP) != 0) {
      const __half2 xh = __floats2half2_rn(x.x, x.y);
      uint32_t ph;
      asm("ex2.approx.f16x2 %0, %1;" : "=r"(ph) : "r"(*reinterpret_cast<const uint32_t*>(&xh)));
      w[pair] = ph;
      e = __half22float2(*reinterpret_cast<const __half2*>(&ph));
    } else {
      if constexpr ((V & AV_NOEXP) != 0) {
        e.x = fmaf(x.x, 1e-4f, 0.5f);
        e.y = fmaf(x.y, 1e-4f, 0.5f);
      } else if (poly) {
        e = exp2_poly2(x);
      } else {
        e.x = ex2_approx(x.x);
        e.y = ex2_approx(x.y);
      }
      const __half2 hh = __floats2half2_rn(e.x, e.y);
      w[pair] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    if constexpr ((V & AV_PACKED) != 0) {
      lsum = __fadd2_rn(lsum, e);
    } else {
      lsum.x += e.x;
      lsum.y += e.y;
    }
  }
}

template <int V>
__device__ __forceinline__ void softmax_store32(const uint32_t (&w)[16], uint8_t* pchunk, int u0, int rx, uint32_t tP) {
  if constexpr ((V & AV_PTMEM) != 0) {
    tmem_st_32x32b_x16(tP, w);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      *reinterpret_cast<uint4*>(pchunk + (((u0 + u) ^ rx) << 4)) = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
  }
}

template <int V>
__device__ __forceinline__ void softmax_chunk32(const uint32_t (&v)[32], float c, float mc, float2& lsum,
                                                uint8_t* pchunk, int u0, int rx, uint32_t tP) {
  uint32_t w[16];
  softmax_exp32<V>(v, c, mc, lsum, w);
  softmax_store32<V>(w, pchunk, u0, rx, tP);
}

__device__ __forceinline__ float max32(const uint32_t (&v)[32]) {
  float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
  for (int i = 4; i < 32; i += 4) {
    m0 = fmaxf(m0, __uint_as_float(v[i]));
    m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
    m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
    m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

template <int V, int DCH>
__global__ void __launch_bounds__(kAtt2Threads, 1) attention2_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                     const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV,
                                                                     const AttnParams p) {
  using Cfg = Att2Cfg<DCH, (V & AV_PTMEM) == 0>;
  static_assert(DCH == 1 || ((V & AV_PTMEM) != 0 && (V & AV_SPLIT) == 0), "d > 64 needs P aliased onto S, no early S");
  // TMEM map: S_g at g*128.  DCH 1 + AV_PTMEM: O_g at 256 + g*64, P_g at 384 + g*64.  Otherwise O_g at 256 + g*128
  // and (DCH 2) P_g aliased onto S_g.
  constexpr uint32_t kOBase = 256, kOStride = (DCH == 1 && (V & AV_PTMEM) != 0) ? 64 : 128;
  constexpr uint32_t kPBase = DCH == 1 ? 384 : 0, kPStride = DCH == 1 ? 64 : 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem;                          // [2 tiles][DCH chunks][128 x 128 B]
  uint8_t* sP = sQ + Cfg::Q_BYTES;             // [2 groups][2 chunks][128 x 128 B] (absent with AV_PTMEM)
  uint8_t* sK = sP + Cfg::P_BYTES;
  uint8_t* sV = sK + Cfg::STAGES * Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + Cfg::STAGES * Cfg::KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + Cfg::STAGES;
  uint64_t* s_full = kv_empty + Cfg::STAGES;   // [2]
  uint64_t* p_full = s_full + 2;               // [2]
  uint64_t* pv_done = p_full + 2;              // [2]
  uint64_t* half_bar = pv_done + 2;            // group 0 is half way through its first exp phase
  uint64_t* s_free = half_bar + 1;             // [2] the group holds tile j's scores in registers (AV_SPLIT)
  uint64_t* exp_turn = s_free + 2;             // [2] group g may enter its exp phase (AV_TURNS)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(exp_turn + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x % p.q_tiles;   // q_tiles counts PAIRS of 128-row tiles here
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int q0 = q_pair * 2 * kBQ;
  const bool two = q0 + kBQ < p.Nq;            // the second tile exists
  // staggering costs half an exp phase of latency once: only worth it on long key sequences
  const bool stagger = (V & AV_STAGGER) != 0 && two && p.n_kv_tiles >= 4;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 128);
      mbar_init(&pv_done[g], 1);
    }
    mbar_init(half_bar, 128);
    mbar_init(&s_free[0], 128);
    mbar_init(&s_free[1], 128);
    mbar_init(&exp_turn[0], 128);
    mbar_init(&exp_turn[1], 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, two ? Cfg::Q_BYTES : Cfg::Q_BYTES / 2);
#pragma unroll
      for (int cc = 0; cc < DCH; ++cc) {
        tma_load_4d(sQ + cc * kChunkBytes, &tmQ, q_full, cc * 64, q0, h, b);
        if (two) tma_load_4d(sQ + (DCH + cc) * kChunkBytes, &tmQ, q_full, cc * 64, q0 + kBQ, h, b);
      }
      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int s = j % Cfg::STAGES;
        const uint32_t ph = (j / Cfg::STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * Cfg::KV_BYTES);
#pragma unroll
        for (int cc = 0; cc < DCH; ++cc) {
          tma_load_4d(sK + s * Cfg::KV_BYTES + cc * kChunkBytes, &tmK, &kv_full[s], cc * 64, j * kBKeys, h, b);
          tma_load_4d(sV + s * Cfg::KV_BYTES + cc * kChunkBytes, &tmV, &kv_full[s], cc * 64, j * kBKeys, h, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const int ng = two ? 2 : 1;
      auto issue_s = [&](int g, int j) {
        const int s = j % Cfg::STAGES;
        const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
        const int n_s = (nk_tile + 15) & ~15;
        const uint32_t idesc = umma_idesc_f16(kBQ, n_s);
        const uint32_t qa = smem_u32(sQ + g * DCH * kChunkBytes);
        const uint32_t ka = smem_u32(sK + s * Cfg::KV_BYTES);
        for (int ks = 0; ks < p.ksteps_qk; ++ks) {
          const uint32_t off = (ks >> 2) * kChunkBytes + (ks & 3) * 32;
          umma_f16_ss(tmem_base + g * 128, umma_desc_kmajor_sw128(qa + off), umma_desc_kmajor_sw128(ka + off), idesc,
                      ks != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[g]);
      };
      mbar_wait(q_full, 0);
      if constexpr ((V & AV_SPLIT) != 0) {
        // ---- S issuer only: tile j+1's scores as soon as the group has read tile j's out of TMEM
        for (int j = 0; j < p.n_kv_tiles; ++j) {
          const int s = j % Cfg::STAGES;
          mbar_wait(&kv_full[s], (j / Cfg::STAGES) & 1);
          tc_fence_after();
          for (int g = 0; g < ng; ++g) {
            if (j > 0) {
              mbar_wait(&s_free[g], (j - 1) & 1);
              tc_fence_after();
            } else if (g == 1 && stagger) {
              mbar_wait(half_bar, 0);
            }
            issue_s(g, j);
          }
        }
      } else {
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        issue_s(0, 0);
        if (ng == 2) {
          if (stagger) mbar_wait(half_bar, 0);
          issue_s(1, 0);
        }
        for (int j = 0; j < p.n_kv_tiles; ++j) {
          const int s = j % Cfg::STAGES;
          const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
          const int ksteps = ((nk_tile + 15) & ~15) >> 4;
          const uint32_t idesc_pv = umma_idesc_f16(kBQ, p.npv, /*b_mn_major=*/1);
          const uint32_t va = smem_u32(sV + s * Cfg::KV_BYTES);
          for (int g = 0; g < ng; ++g) {
            mbar_wait(&p_full[g], j & 1);
            tc_fence_after();
            if constexpr ((V & AV_PTMEM) != 0) {
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t db = umma_desc_mnmajor_sw128(va + ks * 2048, kChunkBytes, 1024);
                umma_f16_ts(tmem_base + kOBase + g * kOStride, tmem_base + kPBase + g * kPStride + ks * 8, db, idesc_pv,
                            (j | ks) != 0 ? 1u : 0u);
              }
            } else {
              const uint32_t pa = smem_u32(sP + g * 2 * kChunkBytes);
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t da = umma_desc_kmajor_sw128(pa + (ks >> 2) * kChunkBytes + (ks & 3) * 32);
                const uint64_t db = umma_desc_mnmajor_sw128(va + ks * 2048, kChunkBytes, 1024);
                umma_f16_ss(tmem_base + kOBase + g * kOStride, da, db, idesc_pv, (j | ks) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&pv_done[g]);
            if (g == ng - 1) umma_commit(&kv_empty[s]);
            if (j + 1 < p.n_kv_tiles) {
              if (g == 0) {
                const int s1 = (j + 1) % Cfg::STAGES;
                mbar_wait(&kv_full[s1], ((j + 1) / Cfg::STAGES) & 1);
                tc_fence_after();
              }
              issue_s(g, j + 1);
            }
          }
        }
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------ P V issuer (AV_SPLIT)
    if constexpr ((V & AV_SPLIT) != 0) {
      if (elect_one()) {
        const int ng = two ? 2 : 1;
        const uint32_t idesc_pv = umma_idesc_f16(kBQ, p.npv, /*b_mn_major=*/1);
        for (int j = 0; j < p.n_kv_tiles; ++j) {
          const int s = j % Cfg::STAGES;
          const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
          const int ksteps = ((nk_tile + 15) & ~15) >> 4;
          const uint32_t va = smem_u32(sV + s * Cfg::KV_BYTES);
          mbar_wait(&kv_full[s], (j / Cfg::STAGES) & 1);   // V tile j landed (this thread's own observation)
          for (int g = 0; g < ng; ++g) {
            mbar_wait(&p_full[g], j & 1);
            tc_fence_after();
            if constexpr ((V & AV_PTMEM) != 0) {
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t db = umma_desc_mnmajor_sw128(va + ks * 2048, kChunkBytes, 1024);
                umma_f16_ts(tmem_base + kOBase + g * kOStride, tmem_base + kPBase + g * kPStride + ks * 8, db, idesc_pv,
                            (j | ks) != 0 ? 1u : 0u);
              }
            } else {
              const uint32_t pa = smem_u32(sP + g * 2 * kChunkBytes);
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t da = umma_desc_kmajor_sw128(pa + (ks >> 2) * kChunkBytes + (ks & 3) * 32);
                const uint64_t db = umma_desc_mnmajor_sw128(va + ks * 2048, kChunkBytes, 1024);
                umma_f16_ss(tmem_base + kOBase + g * kOStride, da, db, idesc_pv, (j | ks) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&pv_done[g]);
          }
          // S(j) of both groups retired long ago (their P tiles exist), so this commit also covers K tile j
          umma_commit(&kv_empty[s]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax groups
    const int g = (warp - 2) >> 2;
    if (g == 0 || two) {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t tS = tmem_base + g * 128 + lane_off;
      const uint32_t tO = tmem_base + kOBase + g * kOStride + lane_off;
      const uint32_t tP = tmem_base + kPBase + g * kPStride + lane_off;
      const float c = p.scale_log2;
      float m_ref = -INFINITY;
      float l = 0.f;
      uint8_t* prow = sP + g * 2 * kChunkBytes + r * 128;   // only dereferenced without AV_PTMEM
      const int rx = r & 7;

      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();
        bool done = false;
        if constexpr ((V & AV_LAZYMAX) != 0) {
          if (nk_tile == kBKeys && j > 0) {
            // ---- optimistic full tile: exponentiate against the running max, fold this tile's max on the side
            uint32_t v0[32], v1[32], v2[32], v3[32];
            tmem_ld_32x32b_x32(tS, v0);
            tmem_ld_32x32b_x32(tS + 32, v1);
            tmem_ld_wait_dep(v0);
            reg_fence(v1);
            tmem_ld_32x32b_x32(tS + 64, v2);      // second half stays in flight under the first exps
            tmem_ld_32x32b_x32(tS + 96, v3);
            const float l_before = l;
            float mc = m_ref * c;
            float2 ls = make_float2(0.f, 0.f);
            float mt = fmaxf(max32(v0), max32(v1));
            uint32_t w[16];
            softmax_exp32<V>(v0, c, mc, ls, w);
            mbar_wait(&pv_done[g], (j - 1) & 1);   // P (and O) are free once P V of the previous tile retired
            tc_fence_after();
            softmax_store32<V>(w, prow, 0, rx, tP);
            softmax_chunk32<V>(v1, c, mc, ls, prow, 4, rx, tP + 16);
            tmem_ld_wait_dep(v2);
            reg_fence(v3);
            if constexpr ((V & AV_SPLIT) != 0) {
              tc_fence_before();
              mbar_arrive(&s_free[g]);
            }
            mt = fmaxf(mt, fmaxf(max32(v2), max32(v3)));
            softmax_chunk32<V>(v2, c, mc, ls, prow + kChunkBytes, 0, rx, tP + 32);
            softmax_chunk32<V>(v3, c, mc, ls, prow + kChunkBytes, 4, rx, tP + 48);
            if (__any_sync(0xffffffffu, (mt - m_ref) * c > 8.0f)) {
              // the running max jumped: rescale O and redo the tile against the new max (overwrites P)
              const float m_new = fmaxf(m_ref, mt);
              const float alpha = ex2_approx((m_ref - m_new) * c);
              m_ref = m_new;
              l = l_before * alpha;
              if constexpr ((V & AV_PTMEM) != 0) tmem_st_wait();
              for (int c0 = 0; c0 < p.npv; c0 += 16) {
                uint32_t o[16];
                tmem_ld_32x32b_x16(tO + c0, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x32b_x16(tO + c0, o);
              }
              tmem_st_wait();
              mc = m_ref * c;
              ls = make_float2(0.f, 0.f);
              softmax_chunk32<V>(v0, c, mc, ls, prow, 0, rx, tP);
              softmax_chunk32<V>(v1, c, mc, ls, prow, 4, rx, tP + 16);
              softmax_chunk32<V>(v2, c, mc, ls, prow + kChunkBytes, 0, rx, tP + 32);
              softmax_chunk32<V>(v3, c, mc, ls, prow + kChunkBytes, 4, rx, tP + 48);
            }
            l += ls.x + ls.y;
            done = true;
          }
        }
        if (done) {
        } else if (nk_tile == kBKeys) {
          // ---- full tile: one TMEM read, scores stay in registers
          uint32_t v0[32], v1[32], v2[32], v3[32];
          tmem_ld_32x32b_x32(tS, v0);
          tmem_ld_32x32b_x32(tS + 32, v1);
          tmem_ld_32x32b_x32(tS + 64, v2);
          tmem_ld_32x32b_x32(tS + 96, v3);
          tmem_ld_wait();
          if constexpr ((V & AV_SPLIT) != 0) {
            tc_fence_before();
            mbar_arrive(&s_free[g]);       // the S buffer may be overwritten by tile j+1
          }
          const float mt = fmaxf(fmaxf(max32(v0), max32(v1)), fmaxf(max32(v2), max32(v3)));
          if (j == 0) {
            m_ref = mt;
          } else {
            mbar_wait(&pv_done[g], (j - 1) & 1);
            tc_fence_after();
            if (__any_sync(0xffffffffu, (mt - m_ref) * c > 8.0f)) {
              const float m_new = fmaxf(m_ref, mt);
              const float alpha = ex2_approx((m_ref - m_new) * c);
              m_ref = m_new;
              l *= alpha;
              for (int c0 = 0; c0 < p.npv; c0 += 16) {
                uint32_t o[16];
                tmem_ld_32x32b_x16(tO + c0, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x32b_x16(tO + c0, o);
              }
              tmem_st_wait();
            }
          }
          const float mc = m_ref * c;
          float2 ls = make_float2(0.f, 0.f);
          if constexpr ((V & AV_TURNS) != 0) {
            // group 0 enters tile j's exps after group 1 left tile j-1's; group 1 after group 0 left tile j's
            if (two && (g == 1 || j > 0)) mbar_wait(&exp_turn[g], (g == 1 ? j : j - 1) & 1);
          }
          softmax_chunk32<V>(v0, c, mc, ls, prow, 0, rx, tP);
          softmax_chunk32<V>(v1, c, mc, ls, prow, 4, rx, tP + 16);
          softmax_chunk32<V>(v2, c, mc, ls, prow + kChunkBytes, 0, rx, tP + 32);
          softmax_chunk32<V>(v3, c, mc, ls, prow + kChunkBytes, 4, rx, tP + 48);
          if constexpr ((V & AV_TURNS) != 0) {
            if (two) mbar_arrive(&exp_turn[g ^ 1]);
          }
          // group 1 starts its first tile when group 0 leaves its first exp phase: from then on one group's exp phase
          // falls into the other's load / max / hand-over phase instead of both sharing the MUFU and then idling it
          if (stagger && g == 0 && j == 0) mbar_arrive(half_bar);
          l += ls.x + ls.y;
        } else {
          // ---- ragged last tile (cross-attention Nk = 77, ToMe-merged Nk): masked two-pass path
          const int n_s = (nk_tile + 15) & ~15;
          float mt = -INFINITY;
          for (int c0 = 0; c0 < n_s; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tS + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i < nk_tile) mt = fmaxf(mt, __uint_as_float(v[i]));
          }
          if (j == 0) {
            m_ref = mt;
          } else {
            mbar_wait(&pv_done[g], (j - 1) & 1);
            tc_fence_after();
            if (__any_sync(0xffffffffu, (mt - m_ref) * c > 8.0f)) {
              const float m_new = fmaxf(m_ref, mt);
              const float alpha = ex2_approx((m_ref - m_new) * c);
              m_ref = m_new;
              l *= alpha;
              for (int c0 = 0; c0 < p.npv; c0 += 16) {
                uint32_t o[16];
                tmem_ld_32x32b_x16(tO + c0, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x32b_x16(tO + c0, o);
              }
              tmem_st_wait();
            }
          }
          const float mc = m_ref * c;
          if constexpr ((V & AV_TURNS) != 0) {
            if (two && (g == 1 || j > 0)) mbar_wait(&exp_turn[g], (g == 1 ? j : j - 1) & 1);
            if (two) mbar_arrive(&exp_turn[g ^ 1]);   // short tile: hand the turn on right away
          }
          for (int c0 = 0; c0 < n_s; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tS + c0, v);
            tmem_ld_wait();
            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = (c0 + i < nk_tile) ? ex2_approx(fmaf(__uint_as_float(v[i]), c, -mc)) : 0.f;
              const float p1 = (c0 + i + 1 < nk_tile) ? ex2_approx(fmaf(__uint_as_float(v[i + 1]), c, -mc)) : 0.f;
              l += p0 + p1;
              const __half2 hh = __floats2half2_rn(p0, p1);
              packed[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            if constexpr ((V & AV_PTMEM) != 0) {
              tmem_st_32x32b_x16(tP + (c0 >> 1), packed);
            } else {
              uint8_t* pchunk = prow + (c0 >> 6) * kChunkBytes;
              const int u0 = (c0 & 63) >> 3;
#pragma unroll
              for (int u = 0; u < 4; ++u)
                *reinterpret_cast<uint4*>(pchunk + (((u0 + u) ^ rx) << 4)) =
                    make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
            }
          }
        }
        if constexpr ((V & AV_PTMEM) != 0) {
          tmem_st_wait();
          tc_fence_before();
        } else {
          tc_fence_before();
          fence_proxy_async();
        }
        mbar_arrive(&p_full[g]);
      }
      // ---- epilogue
      mbar_wait(&pv_done[g], (p.n_kv_tiles - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / l;
      const int row = q0 + g * kBQ + r;
      const bool valid = row < p.Nq;
      __half* dst = p.out + (static_cast<int64_t>(b) * p.Nq + row) * p.ldo + h * p.d;
      for (int c0 = 0; c0 < p.npv; c0 += 16) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(tO + c0, o);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int col = c0 + gg * 8;
            if (col < p.d) {
              __align__(16) __half2 hh[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                hh[i] = __floats2half2_rn(__uint_as_float(o[gg * 8 + 2 * i]) * inv,
                                          __uint_as_float(o[gg * 8 + 2 * i + 1]) * inv);
              *reinterpret_cast<uint4*>(dst + col) = *reinterpret_cast<uint4*>(hh);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}


// ------------------------------------------------------------------------------------------ d <= 64, 16 softmax warps
// Same pipeline as attention2 (two 128-row query tiles per CTA, S of tile j+1 issued under tile j's exps, P kept in
// TMEM, P V issued by its own warp) but every score tile is shared by TWO warps per TMEM lane quadrant, each taking 64
// of its 128 columns: 16 softmax warps = 4 per SM sub-partition instead of 2.  Measured on B200 the two-warp version
// is bound by how fast ONE warp can issue its ~4 instructions per score (issue slots half empty, MUFU half idle,
// forcing the groups to alternate made it slower), so the cure is more warps, not less MUFU work.  The two warps of
// a row pair exchange their partial row max / row sum through shared memory (one 64-thread named barrier per tile).
// Scores are read from TMEM twice (max pass, exp pass) to keep the live registers under the 96 a 640-thread CTA allows.
constexpr int kAtt4Threads = 640;   // warp 0 TMA, 1 S issuer, 2 P V issuer, 3 idle, 4..19 softmax

struct Att4Cfg {
  static constexpr int STAGES = 4;
  static constexpr int Q_BYTES = 2 * kChunkBytes;
  static constexpr int KV_BYTES = kChunkBytes;
  static constexpr int XCH_BYTES = 2 /*parity*/ * 2 /*groups*/ * 2 /*halves*/ * 128 * 4;
  static constexpr int SMEM = Q_BYTES + STAGES * 2 * KV_BYTES + 2 * XCH_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;   // S0, S1 at 0 / 128; O0, O1 at 256 / 320; P0, P1 at 384 / 448
};

template <int V>
__global__ void __launch_bounds__(kAtt4Threads, 1) attention4_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                     const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV,
                                                                     const AttnParams p) {
  using Cfg = Att4Cfg;
  constexpr int VS = V | AV_PTMEM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + Cfg::STAGES * Cfg::KV_BYTES;
  float* xmax = reinterpret_cast<float*>(sV + Cfg::STAGES * Cfg::KV_BYTES);   // [parity][group][half][128]
  float* xsum = xmax + 2 * 2 * 2 * 128;                                       // [group][half][128] (+ unused)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xmax) + 2 * Cfg::XCH_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + Cfg::STAGES;
  uint64_t* s_full = kv_empty + Cfg::STAGES;   // [2]
  uint64_t* s_free = s_full + 2;               // [2] all 256 threads of the group hold / are done with the scores
  uint64_t* p_full = s_free + 2;               // [2]
  uint64_t* pv_done = p_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int q0 = q_pair * 2 * kBQ;
  const bool two = q0 + kBQ < p.Nq;
  const int ng = two ? 2 : 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 256);
      mbar_init(&p_full[g], 256);
      mbar_init(&pv_done[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, two ? Cfg::Q_BYTES : Cfg::Q_BYTES / 2);
      tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
      if (two) tma_load_4d(sQ + kChunkBytes, &tmQ, q_full, 0, q0 + kBQ, h, b);
      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int s = j % Cfg::STAGES;
        mbar_wait(&kv_empty[s], ((j / Cfg::STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * Cfg::KV_BYTES);
        tma_load_4d(sK + s * Cfg::KV_BYTES, &tmK, &kv_full[s], 0, j * kBKeys, h, b);
        tma_load_4d(sV + s * Cfg::KV_BYTES, &tmV, &kv_full[s], 0, j * kBKeys, h, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ S issuer
    if (elect_one()) {
      mbar_wait(q_full, 0);
      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int s = j % Cfg::STAGES;
        const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
        const uint32_t idesc = umma_idesc_f16(kBQ, (nk_tile + 15) & ~15);
        mbar_wait(&kv_full[s], (j / Cfg::STAGES) & 1);
        tc_fence_after();
        const uint32_t ka = smem_u32(sK + s * Cfg::KV_BYTES);
        for (int g = 0; g < ng; ++g) {
          if (j > 0) {
            mbar_wait(&s_free[g], (j - 1) & 1);
            tc_fence_after();
          }
          const uint32_t qa = smem_u32(sQ + g * kChunkBytes);
          for (int ks = 0; ks < p.ksteps_qk; ++ks)
            umma_f16_ss(tmem_base + g * 128, umma_desc_kmajor_sw128(qa + ks * 32), umma_desc_kmajor_sw128(ka + ks * 32),
                        idesc, ks != 0 ? 1u : 0u);
          umma_commit(&s_full[g]);
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ P V issuer
    if (elect_one()) {
      const uint32_t idesc_pv = umma_idesc_f16(kBQ, p.npv, /*b_mn_major=*/1);
      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int s = j % Cfg::STAGES;
        const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
        const int ksteps = ((nk_tile + 15) & ~15) >> 4;
        const uint32_t va = smem_u32(sV + s * Cfg::KV_BYTES);
        mbar_wait(&kv_full[s], (j / Cfg::STAGES) & 1);
        for (int g = 0; g < ng; ++g) {
          mbar_wait(&p_full[g], j & 1);
          tc_fence_after();
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t db = umma_desc_mnmajor_sw128(va + ks * 2048, kChunkBytes, 1024);
            umma_f16_ts(tmem_base + 256 + g * 64, tmem_base + 384 + g * 64 + ks * 8, db, idesc_pv,
                        (j | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&pv_done[g]);
        }
        umma_commit(&kv_empty[s]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax: 2 groups x 2 column halves x 4 quadrants
    const int sw = warp - 4;
    const int g = sw >> 3;
    if (g < ng) {
      const int hh = (sw >> 2) & 1;
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t tS = tmem_base + g * 128 + hh * 64 + lane_off;
      const uint32_t tP = tmem_base + 384 + g * 64 + hh * 32 + lane_off;
      const uint32_t tO = tmem_base + 256 + g * 64 + lane_off;
      const int pair_bar = 1 + g * 4 + q;           // named barrier of the two warps that share these 32 rows
      const float c = p.scale_log2;
      float m_ref = -INFINITY;
      float l = 0.f;
      const int col0 = hh * 64;

      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
        const bool ragged = nk_tile < kBKeys;
        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();
        // ---- pass 1: max of my 64 columns, then of the row
        float mt;
        {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tS, v);
          tmem_ld_wait();
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i >= nk_tile) v[i] = 0xff800000u;
          }
          mt = max32(v);
          tmem_ld_32x32b_x32(tS + 32, v);
          tmem_ld_wait();
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + 32 + i >= nk_tile) v[i] = 0xff800000u;
          }
          mt = fmaxf(mt, max32(v));
        }
        float* xm = xmax + ((j & 1) * 4 + g * 2) * 128;
        xm[hh * 128 + r] = mt;
        named_bar_sync(pair_bar, 64);
        mt = fmaxf(mt, xm[(hh ^ 1) * 128 + r]);
        if (j == 0) {
          m_ref = mt;
        } else {
          mbar_wait(&pv_done[g], (j - 1) & 1);   // P and O are free once P V of the previous tile retired
          tc_fence_after();
          // both warps of the pair see the same row maxima, so they take the same decision
          if (__any_sync(0xffffffffu, (mt - m_ref) * c > 8.0f)) {
            const float m_new = fmaxf(m_ref, mt);
            const float alpha = ex2_approx((m_ref - m_new) * c);
            m_ref = m_new;
            l *= alpha;
            for (int c0 = hh * 16; c0 < p.npv; c0 += 32) {   // the pair splits O's 16-column chunks
              uint32_t o[16];
              tmem_ld_32x32b_x16(tO + c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32b_x16(tO + c0, o);
            }
            tmem_st_wait();
          }
        }
        // ---- pass 2: exponentials of my 64 columns -> packed fp16 P in TMEM
        const float mc = m_ref * c;
        float2 ls = make_float2(0.f, 0.f);
        {
          uint32_t v[32], w[16];
          tmem_ld_32x32b_x32(tS, v);
          tmem_ld_wait();
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i >= nk_tile) v[i] = 0xff800000u;
          }
          softmax_exp32<VS>(v, c, mc, ls, w);
          softmax_store32<VS>(w, nullptr, 0, 0, tP);
          tmem_ld_32x32b_x32(tS + 32, v);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&s_free[g]);                 // the S buffer may be overwritten by tile j+1
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + 32 + i >= nk_tile) v[i] = 0xff800000u;
          }
          softmax_exp32<VS>(v, c, mc, ls, w);
          softmax_store32<VS>(w, nullptr, 0, 0, tP + 16);
        }
        l += ls.x + ls.y;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[g]);
      }
      // ---- epilogue: total row sum from both halves, then each warp normalises its share of O's columns
      float* xs = xsum + g * 2 * 128;
      xs[hh * 128 + r] = l;
      named_bar_sync(pair_bar, 64);
      const float inv = 1.0f / (l + xs[(hh ^ 1) * 128 + r]);
      mbar_wait(&pv_done[g], (p.n_kv_tiles - 1) & 1);
      tc_fence_after();
      const int row = q0 + g * kBQ + r;
      const bool valid = row < p.Nq;
      __half* dst = p.out + (static_cast<int64_t>(b) * p.Nq + row) * p.ldo + h * p.d;
      for (int c0 = hh * 16; c0 < p.npv; c0 += 32) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(tO + c0, o);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int col = c0 + gg * 8;
            if (col < p.d) {
              __align__(16) __half2 hv[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                hv[i] = __floats2half2_rn(__uint_as_float(o[gg * 8 + 2 * i]) * inv,
                                          __uint_as_float(o[gg * 8 + 2 * i + 1]) * inv);
              *reinterpret_cast<uint4*>(dst + col) = *reinterpret_cast<uint4*>(hv);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------ d < 64: lean softmax
// attention5: the pipeline of attention2<AV_PTMEM | AV_SPLIT> (two 128-row query tiles per CTA, S of tile j+1
// issued under tile j's exponentials, P kept in tensor memory, own warp feeding P V) with the per-score
// instruction count of the softmax cut from ~5.5 to ~2.3 - 3 issue slots.  What bounds these kernels at d = 40 was
// measured in isolation (scripts/ubench_pipes.cu, ubench_softmax.cu; profiles/r02_ubench_*.txt): a warp-wide
// MUFU.EX2 occupies the sub-partition's XU for 8 clocks (1024 clocks per 128 x 32 score tile), FFMA2 / F2FP / LOP3
// take 2 clocks each, and with two softmax warps per sub-partition the compiler's schedule reaches ~85 % of the MUFU
// rate at best; every instruction that is not the exponential itself makes that worse.
//   * no max pass.  A tile is exponentiated against the RUNNING reference max m_ref (minus a fixed head-room
//     of 2^kA5Shift), and the packed fp16 P words are OR-ed together on the side (one LOP3 per four scores).
//     Any p >= 2 - i.e. a score more than kA5Shift + 1 binades above the reference - shows up as bit 14 of a
//     half (the MUFU path gives >= 2 / inf / nan there, the polynomial path saturates at exactly 2.0); only then is
//     the tile redone the slow way (true max, rescale of O, exact recomputation from the scores still held in
//     registers).  The first tile and ragged tiles always take the slow path.  P lives in [2^-24, 2): 5 binades
//     of head-room above the reference, 20 below - the same bottom as fp16 P after an exact max within 2^-4.
//   * no row-sum arithmetic (A5_ONES, needs d < 64).  A patch warp writes 1.0 into column d of every V tile after
//     the TMA unit has zero-filled it, so the tensor core accumulates sum_k p[k] in column d of O - the sum of
//     exactly the fp16 values it multiplied, rescaled together with O for free.  P V's N stays <= 64 (the MMA
//     costs the same ~49 clocks for any N <= 96).
//   * polynomial share of the exponentials (A5_POLY*): the argument is mapped onto [0, 1] with a saturating FMA
//     (x' = sat((x + 125) / 126): the clamp at 2^-125 and the ceiling p = 2.0 - the value the overflow test looks
//     for - cost nothing), then Cody-Waite split + degree-3 minimax + exponent insert, all on the FMA / ALU pipes.
//   * softmax warps run with 232 registers (setmaxnreg), the four service warps with 40; mbarrier waits suspend in
//     hardware (try_wait with a time limit) instead of re-polling every ~40 clocks.
// Measured at B16 h8 N4096 d40 (profiles/r02_ab_attn.txt): 736 us (attention2, variant 198) -> 662 us.  What did NOT
// help, each tried on the same box: 16 softmax warps splitting the tile by columns (819 us: the two warps of a row
// pair must agree on the overflow bit every tile), handing P over in 64-key halves so that P V runs under the other
// half's exponentials (720 us), hand-staging scale / ex2 / pack groups with volatile asm (1043 clocks per row in
// isolation against 1200 compiler-ordered, but ptxas re-orders it inside the kernel: 675 us), more polynomial
// (POLY37 / POLY50: the FMA pipe's 2-clock packed ops make a polynomial pair cost 14 clocks against 16 on the MUFU),
// starting group 1 half a tile late (664 vs 662 us).
constexpr int kAttDefaultLean = 2002;   // ATT_VARIANT default: attention5<A5_POLY25> (+ A5_ONES) for d < 64
constexpr int kAtt5Threads = 384;   // warp 0 TMA, 1 S issuer, 2 P V issuer, 3 V patcher, 4..11 two softmax groups
constexpr float kA5Shift = 4.0f;
enum : int { A5_ONES = 1, A5_POLY25 = 2, A5_POLY50 = 4, A5_NOEXP = 8, A5_POLY12 = 16, A5_POLY37 = 64, A5_NOSTAGGER = 128, A5_TRACE = 256 };

struct Att5Cfg {
  static constexpr int STAGES = 4;
  static constexpr int Q_BYTES = 2 * kChunkBytes;
  static constexpr int KV_BYTES = kChunkBytes;
  static constexpr int SMEM = Q_BYTES + STAGES * 2 * KV_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;   // S0, S1 at 0 / 128; O0, O1 at 256 / 320; P0, P1 at 384 / 448
};

__device__ __forceinline__ void tmem_ld_32x32b_x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr) : "memory");
}
template <int N>
__device__ __forceinline__ void set_max_regs_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void set_max_regs_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// slow path of one 32-column chunk: exact, masked (columns >= nvalid give p = 0)
template <int V>
__device__ __forceinline__ void a5_exp32_masked(const uint32_t (&v)[32], int col0, int nvalid, float c, float mcs,
                                                uint32_t (&w)[16], float& lsum) {
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float p0 = (col0 + i < nvalid) ? ex2_approx(fmaf(__uint_as_float(v[i]), c, -mcs)) : 0.f;
    float p1 = (col0 + i + 1 < nvalid) ? ex2_approx(fmaf(__uint_as_float(v[i + 1]), c, -mcs)) : 0.f;
    if constexpr ((V & A5_NOEXP) != 0) {
      p0 = (col0 + i < nvalid) ? 0.03f : 0.f;
      p1 = (col0 + i + 1 < nvalid) ? 0.03f : 0.f;
    }
    const __half2 hh = __floats2half2_rn(p0, p1);
    w[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
    if constexpr ((V & A5_ONES) == 0) lsum += p0 + p1;
  }
}

__device__ __forceinline__ float max32_masked(const uint32_t (&v)[32], int col0, int nvalid) {
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (col0 + i < nvalid) m = fmaxf(m, __uint_as_float(v[i]));
  return m;
}

template <int V>
__device__ __forceinline__ constexpr bool a5_is_poly(int pair) {
  return ((V & A5_POLY50) != 0 && (pair & 1) == 1) || ((V & A5_POLY25) != 0 && (pair & 3) == 3) ||
         ((V & A5_POLY12) != 0 && (pair & 7) == 7) || ((V & A5_POLY37) != 0 && ((pair & 7) == 1 || (pair & 7) == 4 || (pair & 7) == 6));
}

// 2^y for y = 126 x' - 125, x' in [0, 1] (already saturated): n = round(y) through the 1.5 * 2^23 magic add,
// f = y - n in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial, exponent inserted with an integer add.
__device__ __forceinline__ float2 exp2_poly_sat(float2 xs) {
  const float kM = 12582912.0f - 125.0f;
  const float2 r = __ffma2_rn(xs, make_float2(126.0f, 126.0f), make_float2(kM, kM));          // magic + n
  const float2 m = __fadd2_rn(r, make_float2(-kM, -kM));                                         // n + 125
  const float2 f = __ffma2_rn(xs, make_float2(126.0f, 126.0f), make_float2(-m.x, -m.y));        // y - n  (exact: FMA)
  float2 q = __ffma2_rn(make_float2(0.05517084f, 0.05517084f), f, make_float2(0.24260935f, 0.24260935f));
  q = __ffma2_rn(q, f, make_float2(0.69326096f, 0.69326096f));
  q = __ffma2_rn(q, f, make_float2(0.99992818f, 0.99992818f));
  float2 o;
  o.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(r.x) << 23));
  o.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(r.y) << 23));
  return o;
}

__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float r;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// 32 scores of one row -> 16 packed fp16x2 words of p = 2^(s*c - mcs); (cs, os) = (c, 125 - mcs) / 126 feed the
// saturating map of the polynomial share.  ovf collects the OR of the packed words.
template <int V>
__device__ __forceinline__ void a5_exp32(const uint32_t (&v)[32], float c, float mcs, float cs, float os,
                                         uint32_t (&w)[16], uint32_t& ovf, float2& lsum) {
#pragma unroll
  for (int pair = 0; pair < 16; ++pair) {
    const int i = 2 * pair;
    float2 e;
    if constexpr ((V & A5_NOEXP) != 0) {
      e.x = fmaf(__uint_as_float(v[i]), 1e-4f, 0.03f);
      e.y = fmaf(__uint_as_float(v[i + 1]), 1e-4f, 0.03f);
    } else if (a5_is_poly<V>(pair)) {
      float2 xs;
      xs.x = fma_sat(__uint_as_float(v[i]), cs, os);
      xs.y = fma_sat(__uint_as_float(v[i + 1]), cs, os);
      e = exp2_poly_sat(xs);
    } else {
      const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), make_float2(c, c),
                                  make_float2(-mcs, -mcs));
      e.x = ex2_approx(x.x);
      e.y = ex2_approx(x.y);
    }
    const __half2 hh = __floats2half2_rn(e.x, e.y);
    w[pair] = *reinterpret_cast<const uint32_t*>(&hh);
    if constexpr ((V & A5_ONES) == 0) lsum = __fadd2_rn(lsum, e);
  }
  uint32_t a0 = ovf, a1 = 0;
#pragma unroll
  for (int k = 0; k < 16; k += 4) {
    a0 |= w[k] | w[k + 1];
    a1 |= w[k + 2] | w[k + 3];
  }
  ovf = a0 | a1;
}

template <int V>
__global__ void __launch_bounds__(kAtt5Threads, 1) attention5_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                     const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV,
                                                                     const AttnParams p) {
  using Cfg = Att5Cfg;
  constexpr bool kOnes = (V & A5_ONES) != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + Cfg::STAGES * Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + Cfg::STAGES * Cfg::KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + Cfg::STAGES;
  uint64_t* v_ready = kv_empty + Cfg::STAGES;   // the ones column of V stage s is in place
  uint64_t* s_full = v_ready + Cfg::STAGES;     // [2]
  uint64_t* s_free = s_full + 2;                // [2] the group holds tile j's scores in registers
  uint64_t* p_full = s_free + 2;                // [2]
  uint64_t* pv_done = p_full + 2;               // [2]
  uint64_t* stagger_bar = pv_done + 2;          // group 0 is half way through its first tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stagger_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int q0 = q_pair * 2 * kBQ;
  const bool two = q0 + kBQ < p.Nq;
  const int ng = two ? 2 : 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&v_ready[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&s_free[g], 128);
      mbar_init(&p_full[g], 128);
      mbar_init(&pv_done[g], 1);
    }
    mbar_init(stagger_bar, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();
  // trace slots (8 per tile per actor): actor 0 / 1 = softmax group (thread of row 0), 2 = S issuer, 3 = P V issuer
  // (compiled in only with A5_TRACE: the stamps cost ~100 clocks each and perturb the schedule of the softmax loop)
  const bool tr_on = (V & A5_TRACE) != 0 && p.trace != nullptr && blockIdx.x == 0;
  auto stamp = [&](int actor, int j, int k) {
    if constexpr ((V & A5_TRACE) != 0) {
      const int idx = (actor * p.n_kv_tiles + j) * 8 + k;
      if (tr_on && idx < p.trace_cap) p.trace[idx] = clock64();
    }
  };

  if (warp < 4) {
    set_max_regs_dec<40>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, two ? Cfg::Q_BYTES : Cfg::Q_BYTES / 2);
        tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
        if (two) tma_load_4d(sQ + kChunkBytes, &tmQ, q_full, 0, q0 + kBQ, h, b);
        for (int j = 0; j < p.n_kv_tiles; ++j) {
          const int s = j % Cfg::STAGES;
          mbar_wait(&kv_empty[s], ((j / Cfg::STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], 2 * Cfg::KV_BYTES);
          tma_load_4d(sK + s * Cfg::KV_BYTES, &tmK, &kv_full[s], 0, j * kBKeys, h, b);
          tma_load_4d(sV + s * Cfg::KV_BYTES, &tmV, &kv_full[s], 0, j * kBKeys, h, b);
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- S issuer
      if (elect_one()) {
        mbar_wait(q_full, 0);
        for (int j = 0; j < p.n_kv_tiles; ++j) {
          const int s = j % Cfg::STAGES;
          const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
          const uint32_t idesc = umma_idesc_f16(kBQ, (nk_tile + 15) & ~15);
          mbar_wait(&kv_full[s], (j / Cfg::STAGES) & 1);
          tc_fence_after();
          const uint32_t ka = smem_u32(sK + s * Cfg::KV_BYTES);
          for (int g = 0; g < ng; ++g) {
            if (j > 0) {
              mbar_wait(&s_free[g], (j - 1) & 1);
              tc_fence_after();
            }
            const uint32_t qa = smem_u32(sQ + g * kChunkBytes);
            stamp(2, j, g * 2);
            for (int ks = 0; ks < p.ksteps_qk; ++ks)
              umma_f16_ss(tmem_base + g * 128, umma_desc_kmajor_sw128(qa + ks * 32), umma_desc_kmajor_sw128(ka + ks * 32),
                          idesc, ks != 0 ? 1u : 0u);
            umma_commit(&s_full[g]);
            stamp(2, j, g * 2 + 1);
          }
        }
      }
    } else if (warp == 2) {
      // ---------------------------------------------------------------- P V issuer
      if (elect_one()) {
        const uint32_t idesc_pv = umma_idesc_f16(kBQ, p.npv, /*b_mn_major=*/1);
        for (int j = 0; j < p.n_kv_tiles; ++j) {
          const int s = j % Cfg::STAGES;
          const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
          const int ksteps = ((nk_tile + 15) & ~15) >> 4;
          const uint32_t va = smem_u32(sV + s * Cfg::KV_BYTES);
          mbar_wait(kOnes ? &v_ready[s] : &kv_full[s], (j / Cfg::STAGES) & 1);
          for (int g = 0; g < ng; ++g) {
            mbar_wait(&p_full[g], j & 1);
            tc_fence_after();
            stamp(3, j, g * 2);
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t db = umma_desc_mnmajor_sw128(va + ks * 2048, kChunkBytes, 1024);
              umma_f16_ts(tmem_base + 256 + g * 64, tmem_base + 384 + g * 64 + ks * 8, db, idesc_pv,
                          (j | ks) != 0 ? 1u : 0u);
            }
            umma_commit(&pv_done[g]);
            stamp(3, j, g * 2 + 1);
          }
          umma_commit(&kv_empty[s]);
        }
      }
    } else if (kOnes) {
      // ---------------------------------------------------------------- V patcher: column d of every key row := 1.0
      const int unit = p.d >> 3, sub = (p.d & 7) * 2;
      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int s = j % Cfg::STAGES;
        mbar_wait(&kv_full[s], (j / Cfg::STAGES) & 1);
        uint8_t* vt = sV + s * Cfg::KV_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = lane + 32 * i;
          *reinterpret_cast<uint16_t*>(vt + row * 128 + ((unit ^ (row & 7)) << 4) + sub) = 0x3C00;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_ready[s]);
      }
    }
  } else {
    set_max_regs_inc<232>();
    // ------------------------------------------------------------------ softmax groups
    const int g = (warp - 4) >> 2;
    if (g < ng) {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t tS = tmem_base + g * 128 + lane_off;
      const uint32_t tO = tmem_base + 256 + g * 64 + lane_off;
      const uint32_t tP = tmem_base + 384 + g * 64 + lane_off;
      const float c = p.scale_log2;
      float m_ref = -INFINITY;
      float mcs = 0.f;             // m_ref * c + kA5Shift
      float l = 0.f;               // row sum (only without the ones column)
      // Group 1 starts half a tile behind group 0 and stays there (the P V issuer serves g0, g1, g0, ... in turn):
      // the two warps of a sub-partition then reach their MUFU-free phases (barrier waits, TMEM loads, the vote and
      // the P hand-over: about a third of a tile) at different times instead of idling the MUFU together.
      const bool stagger = (V & A5_NOSTAGGER) == 0 && two && p.n_kv_tiles >= 4;
      if (stagger && g == 1) mbar_wait(stagger_bar, 0);

      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
        const int n_s = (nk_tile + 15) & ~15;
        const bool tr = (V & A5_TRACE) != 0 && r == 0;
        if (tr) stamp(g, j, 0);
        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();
        if (tr) stamp(g, j, 1);
        uint32_t v0[32], v1[32], v2[32], v3[32];
        tmem_ld_32x32b_x32(tS, v0);
        tmem_ld_32x32b_x32(tS + 32, v1);
        if (n_s > 64) {
          tmem_ld_32x32b_x32(tS + 64, v2);
          tmem_ld_32x32b_x32(tS + 96, v3);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&s_free[g]);             // the S buffer may be overwritten by tile j+1
        if (tr) stamp(g, j, 2);
        bool slow = (j == 0) || nk_tile < kBKeys;
        if (!slow) {
          // ---- optimistic tile against the running reference
          uint32_t ovf = 0;
          float2 ls = make_float2(0.f, 0.f);
          const float cs = c * (1.0f / 126.0f), os = (125.0f - mcs) * (1.0f / 126.0f);
          {
            // chunk-wise: three 32-score chunks are exponentiated before the wait on the previous tile's P V (P is
            // single-buffered), so that MMA (~400 clocks plus its queueing behind the other group's) stays hidden
            uint32_t w0[16], w1[16], w2[16];
            a5_exp32<V>(v0, c, mcs, cs, os, w0, ovf, ls);
            a5_exp32<V>(v1, c, mcs, cs, os, w1, ovf, ls);
            a5_exp32<V>(v2, c, mcs, cs, os, w2, ovf, ls);
            if (tr) stamp(g, j, 3);
            mbar_wait(&pv_done[g], (j - 1) & 1);
            tc_fence_after();
            if (tr) stamp(g, j, 4);
            tmem_st_32x32b_x16(tP, w0);
            tmem_st_32x32b_x16(tP + 16, w1);
            tmem_st_32x32b_x16(tP + 32, w2);
            a5_exp32<V>(v3, c, mcs, cs, os, w0, ovf, ls);
            tmem_st_32x32b_x16(tP + 48, w0);
          }
          // bit 14: p >= 2 (the polynomial saturates there; the MUFU path gives >= 2, inf or nan)
          const bool bad = (ovf & 0x40004000u) != 0u;
          if (tr) stamp(g, j, 5);
          slow = __any_sync(0xffffffffu, bad);
          if (!slow) l += ls.x + ls.y;
          else tmem_st_wait();
        }
        if (slow) {
          float mt = fmaxf(max32_masked(v0, 0, nk_tile), max32_masked(v1, 32, nk_tile));
          if (n_s > 64) mt = fmaxf(mt, fmaxf(max32_masked(v2, 64, nk_tile), max32_masked(v3, 96, nk_tile)));
          if (j == 0) {
            m_ref = mt;
          } else {
            if (nk_tile < kBKeys) {            // (the optimistic path has already waited)
              mbar_wait(&pv_done[g], (j - 1) & 1);
              tc_fence_after();
            }
            const float m_new = fmaxf(m_ref, mt);
            const float alpha = ex2_approx((m_ref - m_new) * c);
            m_ref = m_new;
            l *= alpha;
            if (__any_sync(0xffffffffu, alpha != 1.0f)) {
              for (int c0 = 0; c0 < p.npv; c0 += 16) {
                uint32_t o[16];
                tmem_ld_32x32b_x16(tO + c0, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x32b_x16(tO + c0, o);
              }
            }
          }
          mcs = fmaf(m_ref, c, kA5Shift);
          uint32_t w[16];
          a5_exp32_masked<V>(v0, 0, nk_tile, c, mcs, w, l);
          tmem_st_32x32b_x16(tP, w);
          if (n_s > 32) {
            a5_exp32_masked<V>(v1, 32, nk_tile, c, mcs, w, l);
            tmem_st_32x32b_x16(tP + 16, w);
          }
          if (stagger && g == 0 && j == 0) mbar_arrive(stagger_bar);
          if (n_s > 64) {
            a5_exp32_masked<V>(v2, 64, nk_tile, c, mcs, w, l);
            tmem_st_32x32b_x16(tP + 32, w);
          }
          if (n_s > 96) {
            a5_exp32_masked<V>(v3, 96, nk_tile, c, mcs, w, l);
            tmem_st_32x32b_x16(tP + 48, w);
          }
        }
        if (tr) stamp(g, j, 6);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[g]);
        if (tr) stamp(g, j, 7);
      }
      // ---- epilogue
      mbar_wait(&pv_done[g], (p.n_kv_tiles - 1) & 1);
      tc_fence_after();
      if constexpr (kOnes) {
        uint32_t lbits;
        tmem_ld_32x32b_x1(tO + p.d, lbits);
        tmem_ld_wait();
        l = __uint_as_float(lbits);
      }
      const float inv = 1.0f / l;
      const int row = q0 + g * kBQ + r;
      const bool valid = row < p.Nq;
      __half* dst = p.out + (static_cast<int64_t>(b) * p.Nq + row) * p.ldo + h * p.d;
      for (int c0 = 0; c0 < p.d; c0 += 16) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(tO + c0, o);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int col = c0 + gg * 8;
            if (col < p.d) {
              __align__(16) __half2 hh[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                hh[i] = __floats2half2_rn(__uint_as_float(o[gg * 8 + 2 * i]) * inv,
                                          __uint_as_float(o[gg * 8 + 2 * i + 1]) * inv);
              *reinterpret_cast<uint4*>(dst + col) = *reinterpret_cast<uint4*>(hh);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int V>
static int launch_attn5_v(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, AttnParams p,
                          unsigned blocks, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    GYRE_CHECK_CUDA(
        cudaFuncSetAttribute(attention5_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att5Cfg::SMEM));
    attr_done = true;
  }
  if ((V & A5_ONES) != 0) p.npv = (p.d + 1 + 15) & ~15;   // O carries the row sum in column d
  return launch_kernel(attention5_kernel<V>, dim3(blocks), dim3(kAtt5Threads), Att5Cfg::SMEM, st, tq, tk, tv, p);
}

template <int V>
static int launch_attn4_v(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                          unsigned blocks, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    GYRE_CHECK_CUDA(
        cudaFuncSetAttribute(attention4_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att4Cfg::SMEM));
    attr_done = true;
  }
  return launch_kernel(attention4_kernel<V>, dim3(blocks), dim3(kAtt4Threads), Att4Cfg::SMEM, st, tq, tk, tv, p);
}

// ------------------------------------------------------------------------------------------ short key sequences
// Cross-attention (77 text tokens) and any Nk <= 128: the whole K/V of one (batch, head) is ONE tile, so the
// roles of the flash kernel are turned around - K and V are loaded once per CTA and the CTA streams over a
// run of 128-row query tiles through a TMA ring.  Two softmax groups alternate tiles (own S / O accumulators
// in TMEM, own P buffer), the single MMA thread interleaves  S_g(i+2) right behind PV_g(i), and the
// normalised output leaves through the group's P buffer as a 128B-swizzled tile with one TMA bulk store
// per 64-column chunk.  The kernel is bound by the Q read + O write; nothing is re-read.
template <int DCH>
struct XAttCfg {
  static constexpr int STAGES = DCH == 1 ? 4 : 2;
  static constexpr int Q_BYTES = DCH * kChunkBytes;             // per stage
  static constexpr int KV_BYTES = DCH * kChunkBytes;            // per operand
  static constexpr int P_BYTES = 2 * kChunkBytes;               // per group (also the output staging tile)
  static constexpr int SMEM = STAGES * Q_BYTES + 2 * KV_BYTES + 2 * P_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;                         // S0, S1: [0,256); O0 at 256, O1 at 384
};

template <int DCH>
__global__ void __launch_bounds__(kXAttThreads, 1) xattention_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                     const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV,
                                                                     const __grid_constant__ CUtensorMap tmO,
                                                                     const AttnParams p) {
  using Cfg = XAttCfg<DCH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem;                                   // [STAGES][DCH chunks]
  uint8_t* sK = sQ + Cfg::STAGES * Cfg::Q_BYTES;
  uint8_t* sV = sK + Cfg::KV_BYTES;
  uint8_t* sP = sV + Cfg::KV_BYTES;                     // [2 groups][2 chunks]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * Cfg::P_BYTES);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = q_full + Cfg::STAGES;
  uint64_t* s_full = q_empty + Cfg::STAGES;             // [2]
  uint64_t* p_full = s_full + 2;                        // [2]
  uint64_t* o_full = p_full + 2;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x % p.splits;
  const int bh = blockIdx.x / p.splits;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int t0 = split * p.tiles_per_cta;
  const int n = min(p.tiles_per_cta, p.q_tiles - t0);   // > 0 by construction of the grid
  const int n_s = (p.Nk + 15) & ~15;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 128);
      mbar_init(&o_full[g], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * Cfg::KV_BYTES);
#pragma unroll
      for (int c = 0; c < DCH; ++c) {
        tma_load_4d(sK + c * kChunkBytes, &tmK, kv_full, c * 64, 0, h, b);
        tma_load_4d(sV + c * kChunkBytes, &tmV, kv_full, c * 64, 0, h, b);
      }
      for (int i = 0; i < n; ++i) {
        const int s = i % Cfg::STAGES;
        mbar_wait(&q_empty[s], ((i / Cfg::STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[s], Cfg::Q_BYTES);
#pragma unroll
        for (int c = 0; c < DCH; ++c)
          tma_load_4d(sQ + s * Cfg::Q_BYTES + c * kChunkBytes, &tmQ, &q_full[s], c * 64, (t0 + i) * kBQ, h, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_f16(kBQ, n_s);
      const uint32_t idesc_pv = umma_idesc_f16(kBQ, p.npv, /*b_mn_major=*/1);
      const uint32_t ka = smem_u32(sK);
      const uint32_t va = smem_u32(sV);
      auto issue_s = [&](int i) {
        const int g = i & 1;
        const int s = i % Cfg::STAGES;
        mbar_wait(&q_full[s], (i / Cfg::STAGES) & 1);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ + s * Cfg::Q_BYTES);
        for (int ks = 0; ks < p.ksteps_qk; ++ks) {
          const uint32_t off = (ks >> 2) * kChunkBytes + (ks & 3) * 32;
          umma_f16_ss(tmem_base + g * 128, umma_desc_kmajor_sw128(qa + off), umma_desc_kmajor_sw128(ka + off), idesc_s,
                      ks != 0 ? 1u : 0u);
        }
        umma_commit(&q_empty[s]);      // Q is only needed for S
        umma_commit(&s_full[g]);
      };
      mbar_wait(kv_full, 0);
      tc_fence_after();
      issue_s(0);
      if (n > 1) issue_s(1);
      const int ksteps = n_s >> 4;
      for (int i = 0; i < n; ++i) {
        const int g = i & 1;
        mbar_wait(&p_full[g], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP + g * Cfg::P_BYTES);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t da = umma_desc_kmajor_sw128(pa + (ks >> 2) * kChunkBytes + (ks & 3) * 32);
          const uint64_t db = umma_desc_mnmajor_sw128(va + ks * 2048, kChunkBytes, 1024);
          umma_f16_ss(tmem_base + 256 + g * 128, da, db, idesc_pv, ks != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[g]);
        if (i + 2 < n) issue_s(i + 2);   // S_g is free: the group read it before it arrived on p_full
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + output, two groups
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t tS = tmem_base + g * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + g * 128 + lane_off;
    const float c = p.scale_log2;
    uint8_t* pbase = sP + g * Cfg::P_BYTES;
    uint8_t* prow = pbase + r * 128;
    const int rx = r & 7;
    const bool leader = (threadIdx.x == 64 + 128 * g);
    const int nk = p.Nk;

    constexpr int VX = AV_PACKED | AV_POLY25;   // softmax arithmetic of the flash kernel's default variant, P in smem
    const int nchunks = (n_s + 31) >> 5;
    for (int i = g; i < n; i += 2) {
      const uint32_t ph = (i >> 1) & 1;
      mbar_wait(&s_full[g], ph);
      tc_fence_after();
      // one TMEM round trip for the whole score row: columns beyond n_s were never written by the MMA and are
      // masked to -inf together with the ragged tail, so they fall out of the max and exponentiate to zero
      uint32_t v0[32], v1[32], v2[32], v3[32];
      tmem_ld_32x32b_x32(tS, v0);
      tmem_ld_32x32b_x32(tS + 32, v1);
      tmem_ld_32x32b_x32(tS + 64, v2);
      if (nchunks > 3) tmem_ld_32x32b_x32(tS + 96, v3);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if (k >= nk) v0[k] = 0xff800000u;
        if (32 + k >= nk) v1[k] = 0xff800000u;
        if (64 + k >= nk) v2[k] = 0xff800000u;
        if (96 + k >= nk || nchunks <= 3) v3[k] = 0xff800000u;
      }
      const float mt = fmaxf(fmaxf(max32(v0), max32(v1)), fmaxf(max32(v2), max32(v3)));
      const float mc = mt * c;
      float2 ls = make_float2(0.f, 0.f);
      // the previous tile's output store must have finished READING this buffer before P overwrites it
      if (leader) tma_store_wait_read<0>();
      named_bar_sync(1 + g, 128);
      softmax_chunk32<VX>(v0, c, mc, ls, prow, 0, rx, 0);
      if (nchunks > 1) softmax_chunk32<VX>(v1, c, mc, ls, prow, 4, rx, 0);
      if (nchunks > 2) softmax_chunk32<VX>(v2, c, mc, ls, prow + kChunkBytes, 0, rx, 0);
      if (nchunks > 3) softmax_chunk32<VX>(v3, c, mc, ls, prow + kChunkBytes, 4, rx, 0);
      const float l = ls.x + ls.y;
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&p_full[g]);
      // ---- output: O / l -> swizzled staging tile (the P buffer, free once PV retired) -> TMA store
      mbar_wait(&o_full[g], ph);
      tc_fence_after();
      const float inv = 1.0f / l;
      for (int c0 = 0; c0 < p.npv; c0 += 16) {
        uint32_t o[16];
        tmem_ld_32x32b_x16(tO + c0, o);
        tmem_ld_wait();
        uint8_t* ochunk = prow + (c0 >> 6) * kChunkBytes;
        const int u0 = (c0 & 63) >> 3;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          __align__(16) __half2 hh[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            hh[k] = __floats2half2_rn(__uint_as_float(o[u * 8 + 2 * k]) * inv, __uint_as_float(o[u * 8 + 2 * k + 1]) * inv);
          *reinterpret_cast<uint4*>(ochunk + (((u0 + u) ^ rx) << 4)) = *reinterpret_cast<uint4*>(hh);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      named_bar_sync(1 + g, 128);
      if (leader) {
#pragma unroll
        for (int cc = 0; cc < DCH; ++cc)
          if (cc * 64 < p.d) tma_store_4d(&tmO, pbase + cc * kChunkBytes, cc * 64, (t0 + i) * kBQ, h, b);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int DCH>
static int launch_xattn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to,
                        const AttnParams& p, unsigned blocks, cudaStream_t st) {
  using Cfg = XAttCfg<DCH>;
  static bool attr_done = false;
  if (!attr_done) {
    GYRE_CHECK_CUDA(
        cudaFuncSetAttribute(xattention_kernel<DCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done = true;
  }
  GYRE_TRY(launch_kernel(xattention_kernel<DCH>, dim3(blocks), dim3(kXAttThreads), Cfg::SMEM, st, tq, tk, tv, to, p));
  return 0;
}

template <int V, int DCH = 1>
static int launch_attn2_v(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                          unsigned blocks, cudaStream_t st) {
  using Cfg = Att2Cfg<DCH, (V & AV_PTMEM) == 0>;
  static bool attr_done = false;
  if (!attr_done) {
    GYRE_CHECK_CUDA(
        cudaFuncSetAttribute(attention2_kernel<V, DCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done = true;
  }
  return launch_kernel(attention2_kernel<V, DCH>, dim3(blocks), dim3(kAtt2Threads), Cfg::SMEM, st, tq, tk, tv, p);
}

static int launch_attn2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, AttnParams p, int B,
                        cudaStream_t st) {
  p.q_tiles = (p.Nq + 2 * kBQ - 1) / (2 * kBQ);
  const long long blocks = static_cast<long long>(B) * p.heads * p.q_tiles;
  GYRE_REQUIRE(blocks < (1ll << 31), "attention: grid too large");
  const unsigned nb = static_cast<unsigned>(blocks);
  if (p.d > 64) return launch_attn2_v<AV_PTMEM | AV_PACKED | AV_POLY25, 2>(tq, tk, tv, p, nb, st);
  // Compiled-in variants = the ones the A/B record in profiles/r01_ab_attn.txt covers (each flag alone on top of its
  // predecessor, the default, the negative results and the no-exp timing floor).
  int variant = tunable(TUNE_ATT_VARIANT);
  // the lean-softmax kernel pays off where V has a spare column for the row sum (d < 64); at d = 64 it measures
  // within 2 % of attention2 (317 vs 311 us at B16 h10 N2304), so the default keeps the older kernel there
  if (variant == kAttDefaultLean && p.d >= 64) variant = AV_PTMEM | AV_SPLIT | AV_PACKED | AV_POLY25;
  switch (variant) {
    case 0: return launch_attn2_v<0>(tq, tk, tv, p, nb, st);
    case AV_STAGGER: return launch_attn2_v<AV_STAGGER>(tq, tk, tv, p, nb, st);
    case AV_STAGGER | AV_PACKED: return launch_attn2_v<AV_STAGGER | AV_PACKED>(tq, tk, tv, p, nb, st);
    case AV_STAGGER | AV_PACKED | AV_POLY25: return launch_attn2_v<AV_STAGGER | AV_PACKED | AV_POLY25>(tq, tk, tv, p, nb, st);
    case AV_SPLIT: return launch_attn2_v<AV_SPLIT>(tq, tk, tv, p, nb, st);
    case AV_SPLIT | AV_PACKED | AV_POLY25: return launch_attn2_v<AV_SPLIT | AV_PACKED | AV_POLY25>(tq, tk, tv, p, nb, st);
    case AV_PTMEM | AV_SPLIT: return launch_attn2_v<AV_PTMEM | AV_SPLIT>(tq, tk, tv, p, nb, st);
    case AV_PTMEM | AV_SPLIT | AV_PACKED | AV_POLY25:      // 198: default
      return launch_attn2_v<AV_PTMEM | AV_SPLIT | AV_PACKED | AV_POLY25>(tq, tk, tv, p, nb, st);
    case AV_PTMEM | AV_SPLIT | AV_PACKED | AV_POLY50:
      return launch_attn2_v<AV_PTMEM | AV_SPLIT | AV_PACKED | AV_POLY50>(tq, tk, tv, p, nb, st);
    case AV_PTMEM | AV_SPLIT | AV_NOEXP: return launch_attn2_v<AV_PTMEM | AV_SPLIT | AV_NOEXP>(tq, tk, tv, p, nb, st);
    case AV_LAZYMAX | AV_PTMEM | AV_SPLIT | AV_PACKED | AV_POLY25:
      return launch_attn2_v<AV_LAZYMAX | AV_PTMEM | AV_SPLIT | AV_PACKED | AV_POLY25>(tq, tk, tv, p, nb, st);
    case AV_TURNS | AV_PTMEM | AV_SPLIT | AV_PACKED | AV_POLY25:
      return launch_attn2_v<AV_TURNS | AV_PTMEM | AV_SPLIT | AV_PACKED | AV_POLY25>(tq, tk, tv, p, nb, st);
    // 1000+: the 16-softmax-warp kernel (attention4); low bits select the arithmetic of the exponentials
    case 1000 + (AV_PACKED | AV_POLY25): return launch_attn4_v<AV_PACKED | AV_POLY25>(tq, tk, tv, p, nb, st);
    case 1000 + (AV_PACKED | AV_NOEXP): return launch_attn4_v<AV_PACKED | AV_NOEXP>(tq, tk, tv, p, nb, st);
  }
  // 2000+: the lean-softmax kernel (attention5); low bits = A5_* flags.  The ones column needs a spare column in
  // the 64-wide V chunk, so d = 64 runs the same kernel with the packed row sum instead.
  const int v5 = variant - 2000;
  if (v5 >= 0 && v5 < 512) {
    const bool ones = p.d < 64;
#define GYRE_A5_CASE(F)                                                                              \
  case F:                                                                                            \
    return ones ? launch_attn5_v<(F) | A5_ONES>(tq, tk, tv, p, nb, st) : launch_attn5_v<(F)>(tq, tk, tv, p, nb, st);
    switch (v5 & ~A5_ONES) {
      GYRE_A5_CASE(0)
      GYRE_A5_CASE(A5_POLY12)
      GYRE_A5_CASE(A5_POLY25)
      GYRE_A5_CASE(A5_POLY37)
      GYRE_A5_CASE(A5_POLY50)
      GYRE_A5_CASE(A5_NOEXP)
      GYRE_A5_CASE(A5_NOSTAGGER)
      GYRE_A5_CASE(A5_NOSTAGGER | A5_POLY25)
      GYRE_A5_CASE(A5_TRACE | A5_POLY25)
      GYRE_A5_CASE(A5_TRACE | A5_NOSTAGGER | A5_POLY25)
    }
#undef GYRE_A5_CASE
  }
  set_last_error("attention: variant %d is not compiled in", variant);
  return -2;
}

template <int DCH>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                       long long blocks, cudaStream_t st) {
  using Cfg = AttnCfg<DCH>;
  static bool attr_done = false;
  if (!attr_done) {
    GYRE_CHECK_CUDA(
        cudaFuncSetAttribute(attention_kernel<DCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done = true;
  }
  return launch_kernel(attention_kernel<DCH>, dim3(static_cast<unsigned>(blocks)), dim3(kAttThreads), Cfg::SMEM, st, tq,
                       tk, tv, p);
}

static int make_head_map(CUtensorMap* m, const __half* base, int ld, int N, int heads, int d, int B, int box_rows) {
  uint64_t dims[4] = {static_cast<uint64_t>(d), static_cast<uint64_t>(N), static_cast<uint64_t>(heads),
                      static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(d) * 2,
                         static_cast<uint64_t>(N) * ld * 2};
  uint32_t box[4] = {64, static_cast<uint32_t>(box_rows), 1, 1};
  uint32_t es[4] = {1, 1, 1, 1};
  return encode_tmap_f16(m, base, 4, dims, strides, box, es, true);
}

static long long* g_attn_trace = nullptr;
static int g_attn_trace_cap = 0;
void attention_set_trace(long long* dev_buf, int capacity) {
  g_attn_trace = dev_buf;
  g_attn_trace_cap = capacity;
}

int attention_f16(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv, int B, int heads,
                  int Nq, int Nk, int d, float scale, __half* out, int ldo, cudaStream_t st) {
  GYRE_REQUIRE(B > 0 && heads > 0 && Nq > 0 && Nk > 0, "attention: empty problem");
  GYRE_REQUIRE(d >= 8 && d % 8 == 0 && d <= 192, "attention: head dim %d must be a multiple of 8 in [8, 192]", d);
  GYRE_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attention: pitches must be multiples of 8");
  GYRE_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                 reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "attention: operands must be 16B aligned");
  GYRE_REQUIRE(ldq >= heads * d && ldk >= heads * d && ldv >= heads * d && ldo >= heads * d,
               "attention: pitch smaller than heads*d");
  AttnParams p{};
  p.Nq = Nq;
  p.Nk = Nk;
  p.heads = heads;
  p.d = d;
  p.n_kv_tiles = (Nk + kBKeys - 1) / kBKeys;
  p.ksteps_qk = (d + 15) / 16;
  p.npv = (d + 15) & ~15;
  p.q_tiles = (Nq + kBQ - 1) / kBQ;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out;
  p.ldo = ldo;
  p.trace = g_attn_trace;
  p.trace_cap = g_attn_trace_cap;
  CUtensorMap tq, tk, tv;
  GYRE_TRY(make_head_map(&tq, q, ldq, Nq, heads, d, B, kBQ));
  GYRE_TRY(make_head_map(&tk, k, ldk, Nk, heads, d, B, kBKeys));
  GYRE_TRY(make_head_map(&tv, v, ldv, Nk, heads, d, B, kBKeys));
  const long long blocks = static_cast<long long>(B) * heads * p.q_tiles;
  GYRE_REQUIRE(blocks < (1ll << 31), "attention: grid too large");
  const int dch = (d + 63) / 64;
  prof::Scope ps(prof::F_ATTN, 4.0 * B * heads * static_cast<double>(Nq) * Nk * d,
                 2.0 * B * heads * d * (2.0 * Nq + 2.0 * Nk), st);
  if (tunable(TUNE_XATTN) && Nk <= kBKeys && dch <= 2 && Nq >= 4 * kBQ) {
    // short key sequence: K/V resident, CTA streams over query tiles; aim at ~2 CTAs per SM in total
    CUtensorMap to;
    GYRE_TRY(make_head_map(&to, out, ldo, Nq, heads, d, B, kBQ));
    const long long bh = static_cast<long long>(B) * heads;
    // CTAs per (batch, head): minimise  waves x (tiles per CTA + fixed cost of ~6 tiles for the K/V load and the
    // pipeline fill) over the split counts that leave every CTA at least 2 tiles
    int best_split = 1;
    long long best_cost = -1;
    for (int sp = 1; sp <= p.q_tiles / 2 || sp == 1; ++sp) {
      const int tpc = (p.q_tiles + sp - 1) / sp;
      const int real = (p.q_tiles + tpc - 1) / tpc;
      const long long waves = (bh * real + 147) / 148;
      const long long cost = waves * (tpc + 6);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_split = real;
      }
      if (sp >= 64) break;
    }
    p.tiles_per_cta = (p.q_tiles + best_split - 1) / best_split;
    p.splits = (p.q_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
    const long long nb = bh * p.splits;
    GYRE_REQUIRE(nb < (1ll << 31), "attention: grid too large");
    return dch == 1 ? launch_xattn<1>(tq, tk, tv, to, p, static_cast<unsigned>(nb), st)
                    : launch_xattn<2>(tq, tk, tv, to, p, static_cast<unsigned>(nb), st);
  }
  if (dch == 1 && Nq > kBQ) return launch_attn2(tq, tk, tv, p, B, st);
  if (dch == 2 && Nq > kBQ && tunable(TUNE_ATT_D128)) return launch_attn2(tq, tk, tv, p, B, st);
  switch (dch) {
    case 1: return launch_attn<1>(tq, tk, tv, p, blocks, st);
    case 2: return launch_attn<2>(tq, tk, tv, p, blocks, st);
    case 3: return launch_attn<3>(tq, tk, tv, p, blocks, st);
  }
  set_last_error("attention: unsupported head dim %d", d);
  return -2;
}

}  // namespace gyre
