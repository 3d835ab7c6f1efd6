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
  float scale_log2;  // scale * log2(e)
  __half* out;
  int ldo;
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
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
constexpr int kAtt2Threads = 320;   // TMA warp, MMA warp, 2 x 4 softmax warps

struct Att2Cfg {
  static constexpr int STAGES = 3;
  static constexpr int Q_BYTES = 2 * kChunkBytes;
  static constexpr int P_BYTES = 2 * 2 * kChunkBytes;
  static constexpr int KV_BYTES = kChunkBytes;                 // per operand per stage
  static constexpr int SMEM = Q_BYTES + P_BYTES + STAGES * 2 * KV_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 512;                        // S0, S1: [0,256); O0 at 256, O1 at 384
};

__global__ void __launch_bounds__(kAtt2Threads, 1) attention2_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                     const __grid_constant__ CUtensorMap tmK,
                                                                     const __grid_constant__ CUtensorMap tmV,
                                                                     const AttnParams p) {
  using Cfg = Att2Cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // [2 tiles][128 x 128 B]
  uint8_t* sP = sQ + Cfg::Q_BYTES;             // [2 groups][2 chunks][128 x 128 B]
  uint8_t* sK = sP + Cfg::P_BYTES;
  uint8_t* sV = sK + Cfg::STAGES * Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + Cfg::STAGES * Cfg::KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + Cfg::STAGES;
  uint64_t* s_full = kv_empty + Cfg::STAGES;   // [2]
  uint64_t* p_full = s_full + 2;               // [2]
  uint64_t* pv_done = p_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x % p.q_tiles;   // q_tiles counts PAIRS of 128-row tiles here
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int q0 = q_pair * 2 * kBQ;
  const bool two = q0 + kBQ < p.Nq;            // the second tile exists

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
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, two ? Cfg::Q_BYTES : Cfg::Q_BYTES / 2);
      tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
      if (two) tma_load_4d(sQ + kChunkBytes, &tmQ, q_full, 0, q0 + kBQ, h, b);
      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int s = j % Cfg::STAGES;
        const uint32_t ph = (j / Cfg::STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * Cfg::KV_BYTES);
        tma_load_4d(sK + s * Cfg::KV_BYTES, &tmK, &kv_full[s], 0, j * kBKeys, h, b);
        tma_load_4d(sV + s * Cfg::KV_BYTES, &tmV, &kv_full[s], 0, j * kBKeys, h, b);
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
        const uint32_t qa = smem_u32(sQ + g * kChunkBytes);
        const uint32_t ka = smem_u32(sK + s * Cfg::KV_BYTES);
        for (int ks = 0; ks < p.ksteps_qk; ++ks)
          umma_f16_ss(tmem_base + g * 128, umma_desc_kmajor_sw128(qa + ks * 32), umma_desc_kmajor_sw128(ka + ks * 32),
                      idesc, ks != 0 ? 1u : 0u);
        umma_commit(&s_full[g]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      for (int g = 0; g < ng; ++g) issue_s(g, 0);
      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int s = j % Cfg::STAGES;
        const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
        const int ksteps = ((nk_tile + 15) & ~15) >> 4;
        const uint32_t idesc_pv = umma_idesc_f16(kBQ, p.npv, /*b_mn_major=*/1);
        const uint32_t va = smem_u32(sV + s * Cfg::KV_BYTES);
        for (int g = 0; g < ng; ++g) {
          mbar_wait(&p_full[g], j & 1);
          tc_fence_after();
          const uint32_t pa = smem_u32(sP + g * 2 * kChunkBytes);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t da = umma_desc_kmajor_sw128(pa + (ks >> 2) * kChunkBytes + (ks & 3) * 32);
            const uint64_t db = umma_desc_mnmajor_sw128(va + ks * 2048, kChunkBytes, 1024);
            umma_f16_ss(tmem_base + 256 + g * 128, da, db, idesc_pv, (j | ks) != 0 ? 1u : 0u);
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
  } else {
    // ------------------------------------------------------------------ softmax groups
    const int g = (warp - 2) >> 2;
    if (g == 0 || two) {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      const uint32_t tS = tmem_base + g * 128 + lane_off;
      const uint32_t tO = tmem_base + 256 + g * 128 + lane_off;
      const float c = p.scale_log2;
      float m_ref = -INFINITY;
      float l = 0.f;
      uint8_t* prow = sP + g * 2 * kChunkBytes + r * 128;
      const int rx = r & 7;

      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int nk_tile = min(kBKeys, p.Nk - j * kBKeys);
        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();
        if (nk_tile == kBKeys) {
          // ---- full tile: one TMEM read, scores stay in registers
          uint32_t v0[32], v1[32], v2[32], v3[32];
          tmem_ld_32x32b_x32(tS, v0);
          tmem_ld_32x32b_x32(tS + 32, v1);
          tmem_ld_32x32b_x32(tS + 64, v2);
          tmem_ld_32x32b_x32(tS + 96, v3);
          tmem_ld_wait();
          float mt = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            mt = fmaxf(mt, fmaxf(fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])),
                                 fmaxf(__uint_as_float(v2[i]), __uint_as_float(v3[i]))));
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
          float l0 = 0.f, l1 = 0.f;
#define GYRE_SOFTMAX_CHUNK(V, CH, U0)                                                              \
          {                                                                                          \
            uint8_t* pchunk = prow + (CH) * kChunkBytes;                                             \
            _Pragma("unroll") for (int u = 0; u < 4; ++u) {                                          \
              uint32_t w[4];                                                                         \
              _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                        \
                const float p0 = ex2_approx(fmaf(__uint_as_float(V[u * 8 + 2 * k]), c, -mc));        \
                const float p1 = ex2_approx(fmaf(__uint_as_float(V[u * 8 + 2 * k + 1]), c, -mc));    \
                l0 += p0;                                                                            \
                l1 += p1;                                                                            \
                const __half2 hh = __floats2half2_rn(p0, p1);                                        \
                w[k] = *reinterpret_cast<const uint32_t*>(&hh);                                      \
              }                                                                                      \
              *reinterpret_cast<uint4*>(pchunk + ((((U0) + u) ^ rx) << 4)) = make_uint4(w[0], w[1], w[2], w[3]); \
            }                                                                                        \
          }
          GYRE_SOFTMAX_CHUNK(v0, 0, 0)
          GYRE_SOFTMAX_CHUNK(v1, 0, 4)
          GYRE_SOFTMAX_CHUNK(v2, 1, 0)
          GYRE_SOFTMAX_CHUNK(v3, 1, 4)
#undef GYRE_SOFTMAX_CHUNK
          l += l0 + l1;
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
            uint8_t* pchunk = prow + (c0 >> 6) * kChunkBytes;
            const int u0 = (c0 & 63) >> 3;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              *reinterpret_cast<uint4*>(pchunk + (((u0 + u) ^ rx) << 4)) =
                  make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
          }
        }
        tc_fence_before();
        fence_proxy_async();
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

static int launch_attn2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, AttnParams p, int B,
                        cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    GYRE_CHECK_CUDA(cudaFuncSetAttribute(attention2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Att2Cfg::SMEM));
    attr_done = true;
  }
  p.q_tiles = (p.Nq + 2 * kBQ - 1) / (2 * kBQ);
  const long long blocks = static_cast<long long>(B) * p.heads * p.q_tiles;
  GYRE_REQUIRE(blocks < (1ll << 31), "attention: grid too large");
  attention2_kernel<<<static_cast<unsigned>(blocks), kAtt2Threads, Att2Cfg::SMEM, st>>>(tq, tk, tv, p);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
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
  attention_kernel<DCH><<<static_cast<unsigned>(blocks), kAttThreads, Cfg::SMEM, st>>>(tq, tk, tv, p);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
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
  CUtensorMap tq, tk, tv;
  GYRE_TRY(make_head_map(&tq, q, ldq, Nq, heads, d, B, kBQ));
  GYRE_TRY(make_head_map(&tk, k, ldk, Nk, heads, d, B, kBKeys));
  GYRE_TRY(make_head_map(&tv, v, ldv, Nk, heads, d, B, kBKeys));
  const long long blocks = static_cast<long long>(B) * heads * p.q_tiles;
  GYRE_REQUIRE(blocks < (1ll << 31), "attention: grid too large");
  const int dch = (d + 63) / 64;
  prof::Scope ps(prof::F_ATTN, 4.0 * B * heads * static_cast<double>(Nq) * Nk * d,
                 2.0 * B * heads * d * (2.0 * Nq + 2.0 * Nk), st);
  if (dch == 1 && Nq > kBQ) return launch_attn2(tq, tk, tv, p, B, st);
  switch (dch) {
    case 1: return launch_attn<1>(tq, tk, tv, p, blocks, st);
    case 2: return launch_attn<2>(tq, tk, tv, p, blocks, st);
    case 3: return launch_attn<3>(tq, tk, tv, p, blocks, st);
  }
  set_last_error("attention: unsupported head dim %d", d);
  return -2;
}

}  // namespace gyre
