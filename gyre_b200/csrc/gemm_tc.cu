// K4 / K1: tcgen05 GEMM and 3x3 implicit-GEMM convolution for sm_100a.
//
// One CTA (persistent, one per SM) computes 128 x BN output tiles.  Warp 0 is the TMA producer, warp 1 owns TMEM and issues
// tcgen05.mma (one elected lane), warps 2..5 are the epilogue (one accumulator row per thread, read
// with tcgen05.ld 32x32b).  Operands are staged by TMA into 128B-swizzled K-major shared-memory tiles
// (64 halfs = one swizzle span per row), STAGES deep, handed over with full/empty mbarriers; the
// accumulator (128 lanes x BN fp32 columns) lives in TMEM.
//
// Convolution: the M tile is a tile_w x tile_h x tile_n patch of output pixels; for each of the 9
// taps and each 64-channel slice the producer issues ONE 4-D TMA box load at the shifted coordinate
// (x0*stride + kw - pad, y0*stride + kh - pad): out-of-bounds pixels are zero-filled by the TMA unit,
// which is exactly the conv zero padding, so no im2col buffer and no boundary code exist.  Stride-2
// convs use the tensor map's element strides.
#include "common.cuh"
#include "ops.h"

namespace gyre {

struct GemmParams {
  int M, N;
  int k_iters;
  int k1_iters;   // k-iterations served by the first A source (tmA); the rest come from tmA2 (skip-concat inputs)
  int n_tiles;
  int m_tiles;
  // conv
  int conv, cin_chunks, cin_pad;
  int tile_w, tile_h, tile_n;
  int Ho, Wo, Bn;
  int tiles_x, tiles_y;
  int stride, pad;
  Epilogue ep;
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kThreads = 320;        // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quadrant)
constexpr int kEpiThreads = 256;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = BN * kBK * 2;
  // as many stages as fit in ~200 KB (one persistent CTA per SM)
  static constexpr int STAGES_RAW = (200 * 1024) / (A_BYTES + B_BYTES);
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  // two accumulator buffers so that the epilogue of tile i overlaps the main loop of tile i+1
  static constexpr int BUF_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int TMEM_COLS = 2 * BUF_COLS;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/ +
                              2 * BN * 4 /*per-tile column bias, double buffered*/;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// Store 8 consecutive output columns (col .. col+7) of one row.
__device__ __forceinline__ void store8(const Epilogue& ep, int64_t row, int col, int n_valid_cols, const float* v) {
  if (ep.out_mode == OUT_F16) {
    __half* dst = reinterpret_cast<__half*>(ep.out) + row * ep.ldo + col;
    if (col + 8 <= n_valid_cols && (ep.ldo & 7) == 0) {
      __align__(16) __half2 h[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(h);
    } else {
      for (int i = 0; i < 8; ++i)
        if (col + i < n_valid_cols) dst[i] = __float2half_rn(v[i]);
    }
  } else if (ep.out_mode == OUT_F32) {
    float* dst = reinterpret_cast<float*>(ep.out) + row * ep.ldo + col;
    if (col + 8 <= n_valid_cols && (ep.ldo & 3) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      for (int i = 0; i < 8; ++i)
        if (col + i < n_valid_cols) dst[i] = v[i];
    }
  } else {  // OUT_SECTIONS; sec_width and hs_d are multiples of 8, so an 8-group never straddles
    if (col >= n_valid_cols) return;
    const int s = col / ep.sec_width;
    const int c = col - s * ep.sec_width;
    const OutSection sec = ep.sec[s];
    if (sec.mode == SEC_ROWMAJOR) {
      __half* dst = reinterpret_cast<__half*>(sec.ptr) + row * sec.ld + c;
      __align__(16) __half2 h[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(h);
    } else {
      const int b = static_cast<int>(row / ep.hs_tokens);
      const int t = static_cast<int>(row - static_cast<int64_t>(b) * ep.hs_tokens);
      const int hd = c / ep.hs_d;
      const int j = c - hd * ep.hs_d;
      const int64_t bh = static_cast<int64_t>(b) * ep.hs_heads + hd;
      if (sec.mode == SEC_HEADSPLIT) {
        __half* dst = reinterpret_cast<__half*>(sec.ptr) + (bh * ep.hs_tpad + t) * ep.hs_dpad + j;
        __align__(16) __half2 h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<uint4*>(h);
      } else {  // SEC_HEADSPLIT_T: [bh, dvpad(rows), tpad]; lanes hold consecutive tokens -> coalesced
        __half* dst = reinterpret_cast<__half*>(sec.ptr) + (bh * ep.hs_dpad + j) * ep.hs_tpad + t;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[static_cast<int64_t>(i) * ep.hs_tpad] = __float2half_rn(v[i]);
      }
    }
  }
}

// Persistent kernel: grid = min(#tiles, #SMs); each CTA walks tiles t = blockIdx.x, +gridDim.x, ...
// (n fastest, so CTAs sharing an A tile run concurrently and hit in L2).  The smem ring and the two TMEM
// accumulator buffers run continuously across tiles: while the epilogue warps drain buffer b, the MMA
// warp is already accumulating the next tile into buffer b^1.
template <int BN>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmA2,
                                                              const __grid_constant__ CUtensorMap tmB,
                                                              const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + Cfg::STAGES * Cfg::B_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;    // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmA2);
    prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t g = 0;   // k-iterations issued so far (ring position)
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n_tile = t % p.n_tiles;
        const int m_tile = t / p.n_tiles;
        int x0 = 0, y0 = 0, n0 = 0;
        if (p.conv) {
          x0 = (m_tile % p.tiles_x) * p.tile_w;
          y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.tile_h;
          n0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.tile_n;
        }
        for (int it = 0; it < p.k_iters; ++it, ++g) {
          const int s = g % Cfg::STAGES;
          const uint32_t ph = (g / Cfg::STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], Cfg::A_BYTES + Cfg::B_BYTES);
          if (p.conv) {
            const int tap = it / p.cin_chunks;
            const int cc = it - tap * p.cin_chunks;
            const int kh = tap / 3, kw = tap - kh * 3;
            tma_load_4d(sA + s * Cfg::A_BYTES, &tmA, &full_bar[s], cc * kBK, x0 * p.stride + kw - p.pad,
                        y0 * p.stride + kh - p.pad, n0);
            tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, &full_bar[s], tap * p.cin_pad + cc * kBK, n_tile * BN);
          } else {
            if (it < p.k1_iters)
              tma_load_2d(sA + s * Cfg::A_BYTES, &tmA, &full_bar[s], it * kBK, m_tile * kBM);
            else
              tma_load_2d(sA + s * Cfg::A_BYTES, &tmA2, &full_bar[s], (it - p.k1_iters) * kBK, m_tile * kBM);
            tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, &full_bar[s], it * kBK, n_tile * BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kBM, BN);
      uint32_t g = 0;
      uint32_t lt = 0;   // local tile counter
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
        const uint32_t buf = lt & 1;
        const uint32_t use = lt >> 1;
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);   // epilogue has drained this buffer's previous tile
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * Cfg::BUF_COLS;
        for (int it = 0; it < p.k_iters; ++it, ++g) {
          const int s = g % Cfg::STAGES;
          const uint32_t ph = (g / Cfg::STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA + s * Cfg::A_BYTES));
          const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sB + s * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_f16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);   // frees the smem slot when these MMAs retire
        }
        umma_commit(&tfull_bar[buf]);   // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: TMEM -> regs -> global
    // 8 warps: warp w reads TMEM lane quadrant (w & 3); the two warps of a quadrant alternate over the
    // 32-column chunks of the tile.  Per tile the column bias is staged once in shared memory; the row-wise
    // operands (residual, per-sample temb bias) are fetched as 16-byte vectors one chunk AHEAD of their use,
    // the first chunk before the accumulator is even ready, so their latency hides under the main loop.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;       // 0 | 1
    const int r = q * 32 + lane;            // accumulator row
    const int etid = threadIdx.x - 64;      // 0..255
    const Epilogue& ep = p.ep;
    float* sbias = reinterpret_cast<float*>(tmem_slot + 4);   // [2][BN]
    constexpr bool kGeglu = false;
    (void)kGeglu;
    uint32_t lt = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
      const int n_tile = t % p.n_tiles;
      const int m_tile = t / p.n_tiles;
      bool valid;
      int64_t out_row;
      if (p.conv) {
        const int x0 = (m_tile % p.tiles_x) * p.tile_w;
        const int y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.tile_h;
        const int n0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.tile_n;
        const int dx = r % p.tile_w;
        const int dy = (r / p.tile_w) % p.tile_h;
        const int dn = r / (p.tile_w * p.tile_h);
        const int x = x0 + dx, y = y0 + dy, n = n0 + dn;
        valid = (x < p.Wo) && (y < p.Ho) && (n < p.Bn);
        out_row = (static_cast<int64_t>(n) * p.Ho + y) * p.Wo + x;
      } else {
        out_row = static_cast<int64_t>(m_tile) * kBM + r;
        valid = out_row < p.M;
      }
      const uint32_t buf = lt & 1;
      const uint32_t use = lt >> 1;
      float* sb = sbias + buf * BN;
      for (int c = etid; c < BN; c += kEpiThreads) {
        const int col = n_tile * BN + c;
        sb[c] = (ep.bias != nullptr && col < p.N) ? __ldg(ep.bias + col) : 0.f;
      }
      const __half* rgb = (valid && ep.rowgroup_bias) ? ep.rowgroup_bias + (out_row / ep.rows_per_group) * ep.rgb_ld
                                                      : nullptr;
      const __half* res = (valid && ep.residual) ? ep.residual + out_row * ep.ldr : nullptr;
      const bool res_vec = res != nullptr && (ep.ldr & 7) == 0 && (p.N & 7) == 0;
      const bool rgb_vec = rgb != nullptr && (ep.rgb_ld & 7) == 0 && (p.N & 7) == 0 &&
                           ((reinterpret_cast<uintptr_t>(ep.rowgroup_bias) & 15) == 0);
      const int col_base = n_tile * BN;
      uint4 rv[4], gv[4];
      auto fetch_rows = [&](int c0, uint4 (&rr)[4], uint4 (&gg)[4]) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int col = col_base + c0 + g8 * 8;
          if (col + 8 <= p.N) {
            if (res_vec) rr[g8] = *reinterpret_cast<const uint4*>(res + col);
            if (rgb_vec) gg[g8] = __ldg(reinterpret_cast<const uint4*>(rgb + col));
          }
        }
      };
      if (ep.act != ACT_GEGLU) fetch_rows(half * 32, rv, gv);
      asm volatile("bar.sync 1, 256;" ::: "memory");   // column bias staged (epilogue warps only)
      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + buf * Cfg::BUF_COLS + (static_cast<uint32_t>(q * 32) << 16);

      if (ep.act == ACT_GEGLU) {
        // tile columns [0, BN/2) = value, [BN/2, BN) = gate (weights are packed that way)
        constexpr int HALF = BN / 2;
        const int n_out = p.N / 2;
#pragma unroll 1
        for (int c0 = half * 32; c0 < HALF; c0 += 64) {
          uint32_t ra[32], rg[32];
          tmem_ld_32x32b_x32(lane_addr + c0, ra);
          tmem_ld_32x32b_x32(lane_addr + HALF + c0, rg);
          tmem_ld_wait();
          if (valid) {
            const int col_out = n_tile * HALF + c0;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int j = g8 * 8 + i;
                const float a = __uint_as_float(ra[j]) + sb[c0 + j];
                const float g = __uint_as_float(rg[j]) + sb[HALF + c0 + j];
                v[i] = a * gelu_erf(g);
              }
              store8(ep, out_row, col_out + g8 * 8, n_out, v);
            }
          }
        }
      } else {
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          uint32_t ra[32];
          tmem_ld_32x32b_x32(lane_addr + c0, ra);
          uint4 rnext[4], gnext[4];
          if (c0 + 64 < BN) fetch_rows(c0 + 64, rnext, gnext);   // next chunk's row operands, before this chunk's stores
          tmem_ld_wait();
          const int col0 = col_base + c0;
          if (valid && col0 < p.N) {
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const int col = col0 + g8 * 8;
              if (col >= p.N) break;
              float v[8];
              const bool full8 = col + 8 <= p.N;
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(ra[g8 * 8 + i]) + sb[c0 + g8 * 8 + i];
              if (rgb) {
                if (full8 && rgb_vec) {
                  const __half2* gh = reinterpret_cast<const __half2*>(&gv[g8]);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(gh[i]);
                    v[2 * i] += f.x;
                    v[2 * i + 1] += f.y;
                  }
                } else {
                  for (int i = 0; i < 8; ++i)
                    if (col + i < p.N) v[i] += __half2float(__ldg(rgb + col + i));
                }
              }
              if (ep.act == ACT_SILU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = silu_f(v[i]);
              }
              if (res) {
                if (full8 && res_vec) {
                  const __half2* rh = reinterpret_cast<const __half2*>(&rv[g8]);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(rh[i]);
                    v[2 * i] += f.x;
                    v[2 * i + 1] += f.y;
                  }
                } else {
                  for (int i = 0; i < 8; ++i)
                    if (col + i < p.N) v[i] += __half2float(res[col + i]);
                }
              }
              store8(ep, out_row, col, p.N, v);
            }
          }
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            rv[g8] = rnext[g8];
            gv[g8] = gnext[g8];
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------ host
static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const GemmParams& p,
                  int m_tiles, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    GYRE_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_done = true;
  }
  const long long tiles = static_cast<long long>(m_tiles) * p.n_tiles;
  GYRE_REQUIRE(tiles > 0 && tiles < (1ll << 31), "gemm: bad tile count %lld", tiles);
  GemmParams pp = p;
  pp.m_tiles = m_tiles;
  const int sms = sm_count();
  const unsigned blocks = static_cast<unsigned>(tiles < sms ? tiles : sms);
  gemm_tc_kernel<BN><<<blocks, kThreads, Cfg::SMEM, st>>>(tmA, tmA2, tmB, pp);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Tile width: maximise (SM wave efficiency) x (1 - N padding) x (per-tile efficiency of the shape).
static int pick_bn(long long m_tiles, int N, int act) {
  if (act == ACT_GEGLU) return 256;
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  const int cand[4] = {256, 160, 128, 64};
  const double shape_eff[4] = {1.0, 0.96, 0.92, 0.78};
  const int sms = sm_count();
  int best = 128;
  double best_score = -1.0;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    const long long n_tiles = (N + bn - 1) / bn;
    const long long tiles = m_tiles * n_tiles;
    const long long waves = (tiles + sms - 1) / sms;
    const double wave_eff = static_cast<double>(tiles) / static_cast<double>(waves * sms);
    const double pad_eff = static_cast<double>(N) / static_cast<double>(n_tiles * bn);
    const double score = wave_eff * pad_eff * shape_eff[i];
    if (score > best_score + 1e-9) {
      best_score = score;
      best = bn;
    }
  }
  return best;
}

static int dispatch(int bn, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                    const GemmParams& p, int m_tiles, cudaStream_t st) {
  switch (bn) {
    case 32: return launch<32>(tmA, tmA2, tmB, p, m_tiles, st);
    case 64: return launch<64>(tmA, tmA2, tmB, p, m_tiles, st);
    case 128: return launch<128>(tmA, tmA2, tmB, p, m_tiles, st);
    case 160: return launch<160>(tmA, tmA2, tmB, p, m_tiles, st);
    case 256: return launch<256>(tmA, tmA2, tmB, p, m_tiles, st);
  }
  set_last_error("gemm: unsupported BN %d", bn);
  return -2;
}

static int check_epilogue(const Epilogue& ep, int N) {
  if (ep.out_mode == OUT_SECTIONS) {
    GYRE_REQUIRE(ep.sec_width > 0 && ep.sec_width % 8 == 0 && N % ep.sec_width == 0 && N / ep.sec_width <= 3,
                 "gemm: bad sections (N=%d width=%d)", N, ep.sec_width);
    for (int s = 0; s < N / ep.sec_width; ++s) {
      GYRE_REQUIRE(ep.sec[s].ptr != nullptr, "gemm: null section %d", s);
      if (ep.sec[s].mode != SEC_ROWMAJOR)
        GYRE_REQUIRE(ep.hs_d % 8 == 0 && ep.hs_dpad % 8 == 0 && ep.hs_heads * ep.hs_d == ep.sec_width &&
                         ep.hs_tokens > 0 && ep.hs_tpad >= ep.hs_tokens,
                     "gemm: bad head-split geometry");
      else
        GYRE_REQUIRE(ep.sec[s].ld % 8 == 0, "gemm: section ld must be a multiple of 8");
    }
  } else {
    GYRE_REQUIRE(ep.out != nullptr && ep.ldo > 0, "gemm: null output");
  }
  return 0;
}

int gemm_f16(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const Epilogue& ep,
             cudaStream_t st) {
  return gemm2_f16(A, lda, K, nullptr, 0, 0, W, ldw, M, N, ep, st);
}

int gemm2_f16(const __half* A, int lda, int K1, const __half* A2, int lda2, int K2, const __half* W, int ldw, int M,
              int N, const Epilogue& ep, cudaStream_t st) {
  const int K = K1 + K2;
  GYRE_REQUIRE(M > 0 && N > 0 && K1 > 0 && K2 >= 0, "gemm: empty problem %dx%dx%d", M, N, K);
  GYRE_REQUIRE(K2 == 0 || (A2 != nullptr && K1 % kBK == 0 && lda2 % 8 == 0 &&
                           (reinterpret_cast<uintptr_t>(A2) & 15) == 0),
               "gemm: two-source A needs K1 %% 64 == 0 and a 16B-aligned second source (K1=%d)", K1);
  GYRE_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "gemm: lda/ldw must be multiples of 8 halfs (TMA 16B pitch)");
  GYRE_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemm: operands must be 16B aligned");
  GYRE_TRY(check_epilogue(ep, ep.act == ACT_GEGLU ? N / 2 : N));
  const int bn = pick_bn((M + kBM - 1) / kBM, N, ep.act);
  if (ep.act == ACT_GEGLU) GYRE_REQUIRE(N % 256 == 0, "gemm: GEGLU needs N %% 256 == 0 (got %d)", N);
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.k1_iters = (K1 + kBK - 1) / kBK;
  p.k_iters = p.k1_iters + (K2 + kBK - 1) / kBK;
  p.n_tiles = (N + bn - 1) / bn;
  p.conv = 0;
  p.ep = ep;
  CUtensorMap tmA, tmA2, tmB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K1), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {kBK, kBM};
    uint32_t es[2] = {1, 1};
    GYRE_TRY(encode_tmap_f16(&tmA, A, 2, dims, strides, box, es, true));
    tmA2 = tmA;
  }
  if (K2 > 0) {
    uint64_t dims[2] = {static_cast<uint64_t>(K2), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda2) * 2};
    uint32_t box[2] = {kBK, kBM};
    uint32_t es[2] = {1, 1};
    GYRE_TRY(encode_tmap_f16(&tmA2, A2, 2, dims, strides, box, es, true));
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {kBK, static_cast<uint32_t>(bn)};
    uint32_t es[2] = {1, 1};
    GYRE_TRY(encode_tmap_f16(&tmB, W, 2, dims, strides, box, es, true));
  }
  prof::Scope ps(prof::F_GEMM, 2.0 * M * (ep.act == ACT_GEGLU ? N : N) * K,
                 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K + static_cast<double>(M) * (ep.act == ACT_GEGLU ? N / 2 : N)), st);
  return dispatch(bn, tmA, tmA2, tmB, p, (M + kBM - 1) / kBM, st);
}

static inline int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

int conv3x3_f16(const __half* X, int ldx, int B, int H, int W, int Cin, const __half* Wp, int Cout, int stride,
                int pad, const Epilogue& ep, cudaStream_t st) {
  GYRE_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3x3: empty problem");
  GYRE_REQUIRE(stride == 1 || stride == 2, "conv3x3: stride %d", stride);
  GYRE_REQUIRE(ldx % 8 == 0 && Cin % 8 == 0, "conv3x3: channel pitch must be a multiple of 8");
  GYRE_REQUIRE(ep.act != ACT_GEGLU, "conv3x3: GEGLU epilogue not supported");
  GYRE_TRY(check_epilogue(ep, Cout));
  const int Ho = (stride == 1) ? H : (pad == 1 ? (H - 1) / 2 + 1 : (H + 1 - 3) / 2 + 1);
  const int Wo = (stride == 1) ? W : (pad == 1 ? (W - 1) / 2 + 1 : (W + 1 - 3) / 2 + 1);
  // choose the 128-pixel patch shape with the least padded work
  int best_w = 0, best_h = 0, best_n = 0;
  long long best_cost = -1;
  const int max_edge = stride == 2 ? 128 : 256;
  for (int tw = 1; tw <= 128; tw <<= 1) {
    if (tw > pow2_ceil(Wo) || tw > max_edge) break;
    for (int th = 1; tw * th <= 128; th <<= 1) {
      if (th > pow2_ceil(Ho)) break;
      const int tn = 128 / (tw * th);
      if (tn > 1 && (tw < pow2_ceil(Wo) || th < pow2_ceil(Ho))) continue;   // batch-fold only whole images
      if (tn > 256) continue;
      const long long cost = 1ll * ((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * ((B + tn - 1) / tn);
      if (best_cost < 0 || cost < best_cost || (cost == best_cost && tw > best_w)) {
        best_cost = cost;
        best_w = tw;
        best_h = th;
        best_n = tn;
      }
    }
  }
  GYRE_REQUIRE(best_cost > 0, "conv3x3: no tile shape for %dx%d", Ho, Wo);
  const int bn = pick_bn(static_cast<long long>((Wo + best_w - 1) / best_w) * ((Ho + best_h - 1) / best_h) *
                             ((B + best_n - 1) / best_n),
                         Cout, ACT_NONE);
  GemmParams p{};
  p.M = 0;
  p.N = Cout;
  p.cin_chunks = (Cin + kBK - 1) / kBK;
  p.cin_pad = p.cin_chunks * kBK;
  p.k_iters = 9 * p.cin_chunks;
  p.k1_iters = p.k_iters;
  p.n_tiles = (Cout + bn - 1) / bn;
  p.conv = 1;
  p.tile_w = best_w;
  p.tile_h = best_h;
  p.tile_n = best_n;
  p.Ho = Ho;
  p.Wo = Wo;
  p.Bn = B;
  p.tiles_x = (Wo + best_w - 1) / best_w;
  p.tiles_y = (Ho + best_h - 1) / best_h;
  p.stride = stride;
  p.pad = pad;
  p.ep = ep;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(ldx) * 2 * W,
                           static_cast<uint64_t>(ldx) * 2 * W * H};
    uint32_t box[4] = {kBK, static_cast<uint32_t>(best_w * stride), static_cast<uint32_t>(best_h * stride),
                       static_cast<uint32_t>(best_n)};
    uint32_t es[4] = {1, static_cast<uint32_t>(stride), static_cast<uint32_t>(stride), 1};
    GYRE_TRY(encode_tmap_f16(&tmA, X, 4, dims, strides, box, es, true));
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(9) * p.cin_pad, static_cast<uint64_t>(Cout)};
    uint64_t strides[1] = {static_cast<uint64_t>(9) * p.cin_pad * 2};
    uint32_t box[2] = {kBK, static_cast<uint32_t>(bn)};
    uint32_t es[2] = {1, 1};
    GYRE_TRY(encode_tmap_f16(&tmB, Wp, 2, dims, strides, box, es, true));
  }
  const int m_tiles = p.tiles_x * p.tiles_y * ((B + best_n - 1) / best_n);
  prof::Scope ps(prof::F_CONV, 2.0 * 9 * Cin * Cout * static_cast<double>(B) * Ho * Wo,
                 2.0 * (static_cast<double>(B) * H * W * Cin + 9.0 * Cin * Cout + static_cast<double>(B) * Ho * Wo * Cout), st);
  return dispatch(bn, tmA, tmA, tmB, p, m_tiles, st);
}

}  // namespace gyre
