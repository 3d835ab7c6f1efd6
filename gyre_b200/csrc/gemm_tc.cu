// K4 / K1: tcgen05 GEMM and 3x3 implicit-GEMM convolution for sm_100a.
//
// One persistent worker per SM - a CTA computing 128 x BN output tiles, or (PAIR2) a 2-CTA cluster computing 256 x BN
// tiles with one tcgen05.mma.cta_group::2 per k step.  Warp 0 is the TMA producer, warp 1 owns TMEM and issues
// tcgen05.mma (one elected lane; in a pair only the leader CTA issues), warps 2..9 are the epilogue (one accumulator
// row per thread, read with tcgen05.ld 32x32b; two warps per TMEM lane quadrant alternate over 32-column slices).
// Operands are staged by TMA into 128B-swizzled K-major shared-memory tiles (64 halfs = one swizzle span per row),
// `stages` deep, handed over with full/empty mbarriers; the accumulators (128 lanes x BN fp32 columns, two buffers)
// live in TMEM.
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
  int stride;
  int taps_w;         // taps per kernel row (3; 2 for the phase convs of the folded upsample)
  int off_x, off_y;   // input offset of tap (0, 0): -pad, or (phase - 1) for the phase convs
  // epilogue routing
  int tma_epi;        // 1: output (and residual) tiles move through swizzled smem slices with TMA
  int rgb_rows;       // rows of a tile that share one rowgroup-bias vector (staged in smem); 0: per-thread loads
  int fast_gelu;      // GEGLU gate through gelu_fast instead of libdevice erff
  int debug;          // measurement only (tunable DEBUG): bit 0 skips the output TMA stores, bit 1 the whole epilogue body
                      // (main-loop time of a shape) - results are wrong with either; bit 2 forces the all-variants image
  int stages;         // smem ring depth in use (tunable GEMM_STAGES caps it; measurement only)
  int mcast;          // CTA pairs (cluster of 2 along M) on one N tile.  1: every B tile is fetched in halves and TMA-
                      // multicast into both CTAs; 2: ONE tcgen05.mma.cta_group::2 (M = 256) per k step, each CTA holds
                      // its A tile and half of the B tile only
  // stream-K over the tiles of the last, partial wave (sk_tiles == 0: off)
  int sk_first;       // first tile index handled by stream-K; tiles below it are dealt round-robin as whole tiles
  int sk_tiles;       // number of stream-K tiles (< grid size)
  int sk_maxp;        // partial-accumulator slots per stream-K tile
  float* sk_ws;       // [sk_tiles][sk_maxp][BN / 4][128] float4 partial accumulators (quad-major: coalesced per warp)
  int* sk_flags;      // [sk_tiles] partials delivered (zero between launches)
  // GroupNorm statistics of the output (EPI 6; Epilogue::gn_out): per 32-column slice of a tile, bit i of gn_mask is set
  // when a group ends behind column pair i of the slice, gn_g0 is the (tile-relative) group of the slice's first column
  uint32_t gn_mask[8];
  uint8_t gn_g0[8];
  int gn_cpg;
  Epilogue ep;
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kThreads = 320;        // TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quadrant)
constexpr int kEpiThreads = 256;
constexpr int kSliceBytes = kBM * 32 * 2;   // one 128-row x 32-column fp16 epilogue slice (64-byte rows)
constexpr int kMaxBiasGroups = 2;
constexpr int kGnTileGroups = 32;    // groups one tile may span when it produces GroupNorm statistics (EPI 6)

template <int BN, bool PAIR2 = false, int EPI = 1>
struct GemmCfg {
  static constexpr int A_BYTES = kBM * kBK * 2;
  // B rows held per CTA and stage: the whole BN-row tile, or half of it under cta_group::2 (-> a deeper ring)
  static constexpr int B_BYTES = (PAIR2 ? BN / 2 : BN) * kBK * 2;
  // epilogue staging: per column-half 2 output + 2 residual slices; per-tile column bias (double buffered)
  static constexpr int EPI_BYTES = 2 * 2 * 2 * kSliceBytes;
  static constexpr int BIAS_BYTES = 2 * kMaxBiasGroups * BN * 4;
  // EPI 6: [4 lane quadrants][kGnTileGroups][2 slice parities] float2 group totals of the tile in flight
  static constexpr int GN_BYTES = EPI == 6 ? 4 * 32 * 2 * 8 : 0;
  static constexpr int FIXED = EPI_BYTES + BIAS_BYTES + 1024 /*align slack*/ + 512 /*barriers*/ + GN_BYTES;
  static constexpr int STAGES_RAW = (227 * 1024 - FIXED) / (A_BYTES + B_BYTES);
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  // two accumulator buffers so that the epilogue of tile i overlaps the main loop of tile i+1
  static constexpr int BUF_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int TMEM_COLS = 2 * BUF_COLS;
  static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + FIXED;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu_f(float x) { return silu_fast(x); }
// Exact (erf) GELU without libdevice's branchy erff: x * Phi(x) with Phi(-|x|) = 0.5 * erfc(|x| / sqrt 2) from
// Abramowitz-Stegun 7.1.26 (|abs err of erf| <= 1.5e-7): branch-free, 2 MUFU (rcp, ex2) + ~14 FMA-pipe ops.
// Evaluating the tail Phi(-|x|) directly keeps the RELATIVE error of gelu(x) small for negative x as well.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float q = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  q = fmaf(q, t, 0.5f * 1.421413741f);
  q = fmaf(q, t, 0.5f * -0.284496736f);
  q = fmaf(q, t, 0.5f * 0.254829592f);
  q *= t;
  const float w = z * 1.2011224087864498f;          // sqrt(log2 e) * z : exp(-z^2) = 2^(-w^2)
  q *= ex2_approx(-w * w);                            // q = Phi(-|x|)
  const float phi = x < 0.f ? q : 1.0f - q;
  return x * phi;
}

// Direct store of 8 consecutive output columns of one row (fp32 outputs, unaligned pitches, tiny N).
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
  } else {
    float* dst = reinterpret_cast<float*>(ep.out) + row * ep.ldo + col;
    if (col + 8 <= n_valid_cols && (ep.ldo & 3) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
      for (int i = 0; i < 8; ++i)
        if (col + i < n_valid_cols) dst[i] = v[i];
    }
  }
}

// Persistent kernel: grid = min(#tiles, #SMs); each CTA walks tiles t = blockIdx.x, +gridDim.x, ...
// (n fastest, so CTAs sharing an A tile run concurrently and hit in L2).  The smem ring and the two TMEM
// accumulator buffers run continuously across tiles: while the epilogue warps drain buffer b, the MMA
// warp is already accumulating the next tile into buffer b^1.
//
// Epilogue (8 warps; warp w owns TMEM lane quadrant w&3, the two warps of a quadrant alternate over
// 32-column slices).  Fast path: a slice is converted in registers, written to a 64B-swizzled shared-memory
// slice and stored with ONE TMA bulk store per slice (full 64-byte row segments, clipped at the tensor
// edges by the TMA unit); the residual slice arrives the same way, prefetched one slice ahead.  A thread
// writing its own row 16 bytes at a time would be bound by L1/LSU request rate, not by HBM.
//
// CTA pairs (p.mcast): the operand stream of a 128 x BN tile needs (128 + BN) x 128 B per 64-wide k step, and one
// SM cannot pull more than ~100 GB/s out of L2 - that, not the tensor pipe, bounds the single-CTA kernel (measured:
// the same ~96 GB/s per SM at BN = 128, 160 and 256).  In pair mode two CTAs of a cluster work on the two M tiles
// of the same N tile; each loads its own A tile and HALF of the B tile, multicast into both CTAs' shared memory.
// A slot is refilled only when both CTAs' MMAs have drained it (their commits arrive on both empty barriers).
//
// cta_group::2 pairs (p.mcast == 2, the default for everything that is not stream-K): measured on B200, a CTA ingests
// at most ~45-50 B/clk from L2 whatever the tile shape, so a 128 x BN tile with its (128 + BN) operand rows per k step
// is ingest-bound at ~85 flop/B (BN = 256) = ~1.0 PFLOP/s - multicast does not help, both CTAs still receive all of B.
// With cta_group::2 the pair computes a 256 x BN tile from ONE instruction stream: each CTA loads its own 128 A rows
// and only BN/2 rows of B (the tensor core reads the other half from the peer's shared memory), i.e. 1.5x fewer bytes
// per flop at BN = 256.  The leader CTA (rank 0) issues every MMA and owns the "stage full" / "accumulator drained"
// barriers: the peer's TMA loads credit their bytes to the leader's full barrier, the peer's epilogue threads arrive
// remotely on the leader's drained barrier; commits are multicast to both CTAs' "slot free" / "accumulator ready"
// barriers.  Each CTA drains its own 128 accumulator rows exactly as in the single-CTA kernel.
// PAIR2 is a template parameter, not a run-time mode: a kernel image that contains cta_group::2 instructions can only be
// launched with an even cluster size ("cluster misconfiguration" otherwise), so the single-CTA kernel must not carry them.
//
// EPI selects how much epilogue code the image carries: 0 = bias / per-sample vector / residual through the TMA slices
// only (what almost every launch of a UNet step uses), 2 = the GEGLU gate (fast GELU) through the TMA slices, 1 = every
// variant behind run-time branches (activations, row-max, fp32 / unaligned direct stores).  With everything in one
// image the 32-way unrolled activation blocks sit between the instructions of the plain path: the slice loop no longer
// fits the instruction cache and the epilogue warps stall on fetches (`no_inst`), which on short-K tiles made the
// epilogue - not the main loop - the critical path.
template <int BN, bool CONV, bool PAIR2 = false, int EPI = 1>
__global__ void __launch_bounds__(kThreads, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmA2,
                                                              const __grid_constant__ CUtensorMap tmB,
                                                              const __grid_constant__ CUtensorMap tmOut,
                                                              const __grid_constant__ CUtensorMap tmRes,
                                                              const GemmParams p) {
  using Cfg = GemmCfg<BN, PAIR2, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::STAGES * Cfg::A_BYTES;
  uint8_t* sEpi = sB + Cfg::STAGES * Cfg::B_BYTES;            // [half][out0,out1,res0,res1][kSliceBytes]
  float* sbias = reinterpret_cast<float*>(sEpi + Cfg::EPI_BYTES);   // [2 bufs][kMaxBiasGroups][BN]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sbias) + Cfg::BIAS_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;    // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
  uint64_t* res_bar = tempty_bar + 2;               // [half][slot] residual slice landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 4);
  float2* gn_tot = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(full_bar) + 512);   // EPI 6 only (Cfg::GN_BYTES)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work units: tiles, or (pair mode) pairs of M tiles; worker = CTA or cluster
  const int rank = p.mcast ? static_cast<int>(cluster_ctarank()) : 0;
  constexpr bool pair2 = PAIR2;
  const int worker = p.mcast ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int n_workers = p.mcast ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int num_tiles = (p.mcast ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;

  // ---- work items.  Whole tiles are dealt round-robin; the tiles of the last, partial wave (p.sk_tiles of them) are
  // cut along K instead: the sk_tiles * k_iters k-iterations are divided evenly over ALL CTAs, every CTA gets one
  // contiguous range (at most two segments in adjacent tiles).  A segment that does not reach the end of its tile's K
  // range dumps its fp32 accumulator to a workspace slot and raises the tile's flag; the CTA holding the last segment
  // (the owner) waits for the flag, adds the slots in a fixed order (deterministic) and runs the normal epilogue.
  // Segments are processed FIRST and partial-producing ones before owned ones, so no CTA ever waits on work that
  // is queued behind another wait.
  struct Item { int t, k0, k1, kind, aux; };          // kind 0 whole tile, 1 partial producer (aux = slot), 2 owner (aux = #partials)
  Item sk_item[2];
  int n_sk = 0;
  const int n_whole = p.sk_tiles > 0 ? p.sk_first : num_tiles;
  if (p.sk_tiles > 0) {
    const long long W = static_cast<long long>(p.sk_tiles) * p.k_iters;
    const long long G = n_workers;
    const long long it0 = worker * W / G, it1 = (worker + 1) * W / G;
    auto make = [&](long long a, long long b) {          // iterations [a, b) inside one tile
      Item it;
      const int tt = static_cast<int>(a / p.k_iters);
      it.t = p.sk_first + tt;
      it.k0 = static_cast<int>(a - static_cast<long long>(tt) * p.k_iters);
      it.k1 = static_cast<int>(b - static_cast<long long>(tt) * p.k_iters);
      int cf = worker;                                    // first CTA whose range reaches into this tile
      while (cf > 0 && cf * W / G > static_cast<long long>(tt) * p.k_iters) --cf;
      it.aux = worker - cf;
      it.kind = it.k1 < p.k_iters ? 1 : (it.k0 > 0 ? 2 : 0);
      return it;
    };
    if (it1 > it0) {
      const long long endA = min(it1, (it0 / p.k_iters + 1) * static_cast<long long>(p.k_iters));
      const Item a = make(it0, endA);
      if (it1 > endA) {
        const Item b2 = make(endA, it1);
        // the segment that produces a partial goes first
        if (b2.kind == 1) { sk_item[0] = b2; sk_item[1] = a; } else { sk_item[0] = a; sk_item[1] = b2; }
        n_sk = 2;
      } else {
        sk_item[0] = a;
        n_sk = 1;
      }
    }
  }
  const int stages = p.stages;                          // ring depth in use (<= Cfg::STAGES)
  auto ring_next = [&](int& s, uint32_t& ph) {
    if (++s == stages) {
      s = 0;
      ph ^= 1;
    }
  };
  auto get_item = [&](int idx, Item& it) -> bool {
    if (idx < n_sk) {
      it = sk_item[idx];
      return true;
    }
    const int t = worker + (idx - n_sk) * n_workers;
    if (t >= n_whole) return false;
    it.t = t;
    it.k0 = 0;
    it.k1 = p.k_iters;
    it.kind = 0;
    it.aux = 0;
    return true;
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmA2);
    prefetch_tmap(&tmB);
    if (p.tma_epi) {
      prefetch_tmap(&tmOut);
      prefetch_tmap(&tmRes);
    }
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], p.mcast == 1 ? 2 : 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], (pair2 ? 2 : 1) * (kEpiThreads / 32));   // one arrival per epilogue warp
    }
    for (int b = 0; b < 4; ++b) mbar_init(&res_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR2) tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.mcast) cluster_sync_all();   // the peer's barriers exist before anything can arrive on them
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int s = 0;          // ring position
      uint32_t ph = 0;    // ring pass parity
      Item wi;
      for (int idx = 0; get_item(idx, wi); ++idx) {
        const int t = wi.t;
        const int n_tile = t % p.n_tiles;
        const int m_tile = p.mcast ? 2 * (t / p.n_tiles) + rank : t / p.n_tiles;
        int x0 = 0, y0 = 0, n0 = 0;
        if (CONV) {
          x0 = (m_tile % p.tiles_x) * p.tile_w;
          y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.tile_h;
          n0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.tile_n;
        }
        for (int it = wi.k0; it < wi.k1; ++it, ring_next(s, ph)) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if constexpr (PAIR2) {
            // both CTAs' loads of this stage are counted on the leader's barrier
            const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[s]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
            if (CONV) {
              const int tap = it / p.cin_chunks;
              const int cc = it - tap * p.cin_chunks;
              const int kh = tap / p.taps_w, kw = tap - kh * p.taps_w;
              tma_load_4d_2cta(sA + s * Cfg::A_BYTES, &tmA, lead_full, cc * kBK, x0 * p.stride + kw + p.off_x,
                               y0 * p.stride + kh + p.off_y, n0);
              tma_load_2d_2cta(sB + s * Cfg::B_BYTES, &tmB, lead_full, tap * p.cin_pad + cc * kBK,
                               n_tile * BN + rank * (BN / 2));
            } else {
              if (it < p.k1_iters)
                tma_load_2d_2cta(sA + s * Cfg::A_BYTES, &tmA, lead_full, it * kBK, m_tile * kBM);
              else
                tma_load_2d_2cta(sA + s * Cfg::A_BYTES, &tmA2, lead_full, (it - p.k1_iters) * kBK, m_tile * kBM);
              tma_load_2d_2cta(sB + s * Cfg::B_BYTES, &tmB, lead_full, it * kBK, n_tile * BN + rank * (BN / 2));
            }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[s], Cfg::A_BYTES + Cfg::B_BYTES);
          if (CONV) {
            const int tap = it / p.cin_chunks;
            const int cc = it - tap * p.cin_chunks;
            const int kh = tap / p.taps_w, kw = tap - kh * p.taps_w;
            tma_load_4d(sA + s * Cfg::A_BYTES, &tmA, &full_bar[s], cc * kBK, x0 * p.stride + kw + p.off_x,
                        y0 * p.stride + kh + p.off_y, n0);
            if (p.mcast)
              tma_load_2d_mcast(sB + s * Cfg::B_BYTES + rank * (Cfg::B_BYTES / 2), &tmB, &full_bar[s],
                                tap * p.cin_pad + cc * kBK, n_tile * BN + rank * (BN / 2), 3);
            else
              tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, &full_bar[s], tap * p.cin_pad + cc * kBK, n_tile * BN);
          } else {
            if (it < p.k1_iters)
              tma_load_2d(sA + s * Cfg::A_BYTES, &tmA, &full_bar[s], it * kBK, m_tile * kBM);
            else
              tma_load_2d(sA + s * Cfg::A_BYTES, &tmA2, &full_bar[s], (it - p.k1_iters) * kBK, m_tile * kBM);
            if (p.mcast)
              tma_load_2d_mcast(sB + s * Cfg::B_BYTES + rank * (Cfg::B_BYTES / 2), &tmB, &full_bar[s], it * kBK,
                                n_tile * BN + rank * (BN / 2), 3);
            else
              tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, &full_bar[s], it * kBK, n_tile * BN);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if constexpr (PAIR2) {
      // one instruction stream for the pair: the leader issues 256 x BN MMAs over both CTAs' operands
      if (rank == 0 && elect_one()) {
        constexpr uint32_t idesc2 = umma_idesc_f16(2 * kBM, BN);
        int s = 0;
        uint32_t ph = 0;
        uint32_t lt = 0;
        Item wi;
        for (int idx = 0; get_item(idx, wi); ++idx, ++lt) {
          const uint32_t buf = lt & 1;
          const uint32_t use = lt >> 1;
          mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);   // both CTAs' epilogues have drained this buffer
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * Cfg::BUF_COLS;
          for (int it = wi.k0; it < wi.k1; ++it, ring_next(s, ph)) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA + s * Cfg::A_BYTES));
            const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sB + s * Cfg::B_BYTES));
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_f16_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc2, (it > wi.k0 || k != 0) ? 1u : 0u);
            umma_commit_2cta_mcast(&empty_bar[s], 3);     // the slot is free in both CTAs
          }
          umma_commit_2cta_mcast(&tfull_bar[buf], 3);     // accumulator halves complete in both CTAs
        }
      }
    } else if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kBM, BN);
      int s = 0;
        uint32_t ph = 0;
      uint32_t lt = 0;   // local tile counter
      Item wi;
      for (int idx = 0; get_item(idx, wi); ++idx, ++lt) {
        const uint32_t buf = lt & 1;
        const uint32_t use = lt >> 1;
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);   // epilogue has drained this buffer's previous tile
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * Cfg::BUF_COLS;
        for (int it = wi.k0; it < wi.k1; ++it, ring_next(s, ph)) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA + s * Cfg::A_BYTES));
          const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sB + s * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_f16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (it > wi.k0 || k != 0) ? 1u : 0u);
          // frees the smem slot when these MMAs retire (pair mode: in both CTAs - the peer refills half of it)
          if (p.mcast) umma_commit_mcast(&empty_bar[s], 3);
          else umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[buf]);   // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;       // 0 | 1: which 32-column slices of the tile this warp handles
    const int r = q * 32 + lane;            // accumulator row == TMEM lane
    const int etid = threadIdx.x - 64;      // 0..255
    const bool leader = (threadIdx.x == 64 + 128 * half);
    const Epilogue& ep = p.ep;
    // EPI 3: the lean image + per-row (sum, sum of squares) of the produced row segments (LayerNorm statistics for the
    // next GEMM); EPI 4 / 5: the lean / GEGLU image consuming such statistics (LayerNorm folded into this GEMM)
    constexpr bool kGegluCode = (EPI == 1 || EPI == 2 || EPI == 5);
    constexpr bool kPlainCode = (EPI != 2 && EPI != 5);
    constexpr bool kRowStat = (EPI == 3);
    constexpr bool kGnStat = (EPI == 6);
    if constexpr (kGnStat) {
      for (int i = etid; i < 4 * kGnTileGroups * 2; i += kEpiThreads) gn_tot[i] = make_float2(0.f, 0.f);
    }
    constexpr bool kLnConsume = (EPI == 4 || EPI == 5);
    const bool geglu = EPI == 2 || EPI == 5 || (EPI == 1 && ep.act == ACT_GEGLU);
    const int bn_out = geglu ? BN / 2 : BN;           // output columns per tile
    const int n_out = geglu ? p.N / 2 : p.N;          // output columns of the problem
    uint8_t* sOut = sEpi + half * 4 * kSliceBytes;    // 2 slots
    uint8_t* sRes = sOut + 2 * kSliceBytes;           // 2 slots
    const int swz = (r >> 1) & 3;                     // 64B swizzle: 16-byte unit index ^= (row / 2) & 3
    uint8_t* my_out_row = sOut + r * 64;
    const uint8_t* my_res_row = sRes + r * 64;
    uint32_t slice_cnt = 0;                           // slices processed by this half (slot = cnt & 1)
    uint32_t lt = 0;
    // bias (+ per-sample vector) entry i of tile tt: i = group * BN + column
    const int nbias = (kLnConsume ? 2 : (p.rgb_rows > 0 ? (kBM / p.rgb_rows) : 1)) * BN;
    auto bias_of = [&](int tt, int i) -> float {
      if (i >= nbias) return 0.f;
      const int gi = i / BN;
      const int col = (tt % p.n_tiles) * BN + (i - gi * BN);        // accumulator column == bias index
      if (col >= p.N) return 0.f;
      if constexpr (kLnConsume) {
        if (gi == 1) return __ldg(ep.ln_colsum + col);              // second group: column sums of gamma (.) W
      }
      float v = ep.bias ? __ldg(ep.bias + col) : 0.f;
      if (p.rgb_rows > 0) {
        const int mt = p.mcast ? 2 * (tt / p.n_tiles) + rank : tt / p.n_tiles;
        const int nn = (mt / (p.tiles_x * p.tiles_y)) * p.tile_n + gi;
        if (nn < p.Bn) v += __half2float(__ldg(ep.rowgroup_bias + static_cast<int64_t>(nn) * ep.rgb_ld + col));
      }
      return v;
    };
    // "accumulator drained": ONE arrival per warp (256 per-thread arrivals on one mbarrier serialise in the barrier
    // unit - and 256 remote ones per tile made the pair kernel slower than the single-CTA kernel on short-K tiles).
    // Pairs: the leader's MMA warp waits for both CTAs, so the peer arrives on the leader's barrier.
    auto release_accumulator = [&](uint32_t b) {
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[b]), 0));
        else mbar_arrive(&tempty_bar[b]);
      }
    };
    float nb0 = 0.f, nb1 = 0.f;
    // LayerNorm consumer: (mean, rstd) of this thread's row, requested one tile ahead like the bias
    // (ln_parts == 0: finished (mean, rstd); 1..8: the producer's raw partials, folded when the tile starts - the loads
    // are issued a tile ahead and consumed a tile later, so their latency never sits in an epilogue)
    struct RowStatRaw { float2 v[8]; };
    auto rowstat_of = [&](int tt) -> RowStatRaw {
      RowStatRaw o;
      const int mt = p.mcast ? 2 * (tt / p.n_tiles) + rank : tt / p.n_tiles;
      const int64_t row = static_cast<int64_t>(mt) * kBM + r;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o.v[i] = (row < p.M && i < (ep.ln_parts == 0 ? 1 : ep.ln_parts))
                     ? __ldg(ep.ln_rowstat + static_cast<int64_t>(i) * p.M + row)
                     : make_float2(0.f, 0.f);
      return o;
    };
    auto rowstat_fold = [&](const RowStatRaw& o) -> float2 {
      if (ep.ln_parts == 0) return o.v[0];
      const float s = ((o.v[0].x + o.v[1].x) + (o.v[2].x + o.v[3].x)) + ((o.v[4].x + o.v[5].x) + (o.v[6].x + o.v[7].x));
      const float q2 = ((o.v[0].y + o.v[1].y) + (o.v[2].y + o.v[3].y)) + ((o.v[4].y + o.v[5].y) + (o.v[6].y + o.v[7].y));
      const float mean = s * ep.ln_inv_c;
      const float var = fmaxf(fmaf(-mean, mean, q2 * ep.ln_inv_c), 0.f);
      return make_float2(mean, rsqrtf(var + ep.ln_eps));
    };
    RowStatRaw ln_next;
#pragma unroll
    for (int i = 0; i < 8; ++i) ln_next.v[i] = make_float2(0.f, 0.f);
    Item wi, wnext;
    bool have = get_item(0, wi);
    if (have) {
      nb0 = bias_of(wi.t, etid);
      nb1 = bias_of(wi.t, etid + kEpiThreads);
      if constexpr (kLnConsume) ln_next = rowstat_of(wi.t);
    }
    for (int idx = 0; have; ++idx, ++lt) {
      const bool have_next = get_item(idx + 1, wnext);
      const int t = wi.t;
      const int n_tile = t % p.n_tiles;
      const int m_tile = p.mcast ? 2 * (t / p.n_tiles) + rank : t / p.n_tiles;
      int x0 = 0, y0 = 0, n0 = 0;
      bool valid;
      int64_t out_row;
      if (CONV) {
        x0 = (m_tile % p.tiles_x) * p.tile_w;
        y0 = ((m_tile / p.tiles_x) % p.tiles_y) * p.tile_h;
        n0 = (m_tile / (p.tiles_x * p.tiles_y)) * p.tile_n;
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
      const int col_base = n_tile * bn_out;                         // first OUTPUT column of the tile
      const int nsl = (min(bn_out, n_out - col_base) + 31) >> 5;    // 32-column output slices in this tile

      // ---- per-tile column bias (+ the per-sample temb vector when it is uniform over row blocks) -> smem.
      // The values were requested one tile ahead (nb0 / nb1): a global load at this point would put ~1 us of
      // latency into every tile's epilogue, more than the main loop of a short-K tile takes.
      float* sb = sbias + buf * (kMaxBiasGroups * BN);
      if (etid < nbias) sb[etid] = nb0;
      if (etid + kEpiThreads < nbias) sb[etid + kEpiThreads] = nb1;
      float ln_mean = 0.f, ln_rstd = 0.f;
      if constexpr (kLnConsume) {
        const float2 ms = rowstat_fold(ln_next);
        ln_mean = ms.x;
        ln_rstd = ms.y;
      }
      if (have_next) {
        nb0 = bias_of(wnext.t, etid);
        nb1 = bias_of(wnext.t, etid + kEpiThreads);
        if constexpr (kLnConsume) ln_next = rowstat_of(wnext.t);
      }
      const float* sbr = sb + (p.rgb_rows > 0 ? (r / p.rgb_rows) * BN : 0);
      const float* scs = sb + BN;       // LayerNorm consumer: column sums of the gamma-scaled weights
      float2 rs2 = make_float2(0.f, 0.f), rq2 = make_float2(0.f, 0.f);   // row-statistics producer: this thread's share of its row (two lanes)
      (void)ln_mean; (void)ln_rstd; (void)scs;
      const uint32_t lane_addr = tmem_base + buf * Cfg::BUF_COLS + (static_cast<uint32_t>(q * 32) << 16);

      // ---- stream-K bookkeeping of this item
      const int sk_tt = wi.t - p.sk_first;                    // index among the stream-K tiles (kind != 0 only)
      const int npr = p.mcast ? 2 : 1;                        // CTAs per worker: each keeps its own partial slots / flag
      const float* sk_part = nullptr;                         // owner: slot 0 of this tile's partial accumulators
      if (wi.kind == 1) {
        // partial producer: dump the raw fp32 accumulator (all BN columns of my row) and raise the tile's flag
        // slot layout [BN / 4 column quads][128 rows] float4: a warp's 32 rows of one quad are 512 contiguous bytes, so
        // the dump and the owner's loads are fully coalesced (row-major slots cost 32 sectors per warp instruction)
        float* dst = p.sk_ws + ((static_cast<size_t>(sk_tt) * p.sk_maxp + wi.aux) * npr + rank) * (kBM * BN) + r * 4;
        named_bar_sync(1, kEpiThreads);
        mbar_wait(&tfull_bar[buf], use & 1);
        tc_fence_after();
        for (int sl = half; sl < BN / 32; sl += 2) {
          uint32_t ra[32];
          tmem_ld_32x32b_x32(lane_addr + sl * 32, ra);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(dst + (sl * 8 + u) * (kBM * 4)) = make_uint4(ra[4 * u], ra[4 * u + 1], ra[4 * u + 2], ra[4 * u + 3]);
        }
        __threadfence();
        named_bar_sync(1, kEpiThreads);
        if (etid == 0) atomicAdd(p.sk_flags + sk_tt * npr + rank, 1);
        tc_fence_before();
        release_accumulator(buf);
        wi = wnext;
        have = have_next;
        continue;
      }
      if (wi.kind == 2) {
        // owner: the earlier segments of this tile were queued before anything their CTAs could wait on
        if (etid == 0) {
          const volatile int* f = p.sk_flags + sk_tt * npr + rank;
          uint32_t spins = 0;
          while (*f < wi.aux) {
            if (++spins > (1u << 27)) {
              printf("gyre_b200: stream-K flag timeout block %d tile %d\n", blockIdx.x, wi.t);
              __trap();
            }
          }
          __threadfence();
        }
        sk_part = p.sk_ws + ((static_cast<size_t>(sk_tt) * p.sk_maxp) * npr + rank) * (kBM * BN) + r * 4;
      }
      // adds the partial accumulators of this row, columns [c, c + 32), in slot order
      auto add_partials = [&](uint32_t (&acc)[32], int c) {
        for (int s2 = 0; s2 < wi.aux; ++s2) {
          const float* src = sk_part + static_cast<size_t>(s2) * npr * kBM * BN + (c >> 2) * (kBM * 4);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 v4 = __ldcg(reinterpret_cast<const float4*>(src + u * (kBM * 4)));
            acc[4 * u] = __float_as_uint(__uint_as_float(acc[4 * u]) + v4.x);
            acc[4 * u + 1] = __float_as_uint(__uint_as_float(acc[4 * u + 1]) + v4.y);
            acc[4 * u + 2] = __float_as_uint(__uint_as_float(acc[4 * u + 2]) + v4.z);
            acc[4 * u + 3] = __float_as_uint(__uint_as_float(acc[4 * u + 3]) + v4.w);
          }
        }
      };

      if (p.debug & 2) {
        // measurement only: no epilogue work at all (main-loop time of the kernel)
        named_bar_sync(1, kEpiThreads);
        mbar_wait(&tfull_bar[buf], use & 1);
        tc_fence_after();
      } else if (EPI != 1 || p.tma_epi) {
        // ================================================= fast path: smem slices + TMA
        const bool has_res = ep.residual != nullptr;
        auto issue_res = [&](int slice, uint32_t cnt) {       // leader only
          const int slot = cnt & 1;
          mbar_arrive_expect_tx(&res_bar[half * 2 + slot], kSliceBytes);
          if (CONV)
            tma_load_4d(sRes + slot * kSliceBytes, &tmRes, &res_bar[half * 2 + slot], col_base + slice * 32, x0, y0, n0);
          else
            tma_load_2d(sRes + slot * kSliceBytes, &tmRes, &res_bar[half * 2 + slot], col_base + slice * 32,
                        m_tile * kBM);
        };
        // the first residual slice of the tile is requested before the accumulator is ready
        if (has_res && leader && half < nsl) issue_res(half, slice_cnt);
        named_bar_sync(1, kEpiThreads);                        // column bias staged
        mbar_wait(&tfull_bar[buf], use & 1);
        tc_fence_after();
        for (int sl = half; sl < nsl; sl += 2, ++slice_cnt) {
          const int slot = slice_cnt & 1;
          if (leader) {
            tma_store_wait_read<1>();                          // the store that last used this out slot has drained
            if (has_res && sl + 2 < nsl) issue_res(sl + 2, slice_cnt + 1);
          }
          named_bar_sync(2 + half, 128);
          float v[32];
          if constexpr (kGegluCode) {
            if (geglu) {
              uint32_t ra[32], rg[32];
              tmem_ld_32x32b_x32(lane_addr + sl * 32, ra);
              tmem_ld_32x32b_x32(lane_addr + BN / 2 + sl * 32, rg);
              tmem_ld_wait();
              if (sk_part) {
                add_partials(ra, sl * 32);
                add_partials(rg, BN / 2 + sl * 32);
              }
              if constexpr (EPI == 5) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float a = fmaf(ln_rstd, fmaf(-ln_mean, scs[sl * 32 + j], __uint_as_float(ra[j])), sbr[sl * 32 + j]);
                  const float gg = fmaf(ln_rstd, fmaf(-ln_mean, scs[BN / 2 + sl * 32 + j], __uint_as_float(rg[j])),
                                        sbr[BN / 2 + sl * 32 + j]);
                  v[j] = a * gelu_fast(gg);
                }
              } else if (EPI == 2 || p.fast_gelu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float a = __uint_as_float(ra[j]) + sbr[sl * 32 + j];
                  const float gg = __uint_as_float(rg[j]) + sbr[BN / 2 + sl * 32 + j];
                  v[j] = a * gelu_fast(gg);
                }
              } else {
                if constexpr (EPI == 1) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) {
                    const float a = __uint_as_float(ra[j]) + sbr[sl * 32 + j];
                    const float gg = __uint_as_float(rg[j]) + sbr[BN / 2 + sl * 32 + j];
                    v[j] = a * gelu_erf(gg);
                  }
                }
              }
            }
          }
          if constexpr (kPlainCode) {
            if (!geglu) {
              uint32_t ra[32];
              tmem_ld_32x32b_x32(lane_addr + sl * 32, ra);
              tmem_ld_wait();
              if (sk_part) add_partials(ra, sl * 32);
              // packed fp32x2 arithmetic (FADD2 / FFMA2): half the issue slots of the slice loop, same roundings
              if constexpr (kLnConsume) {
                const float2 nm = make_float2(-ln_mean, -ln_mean), rs2v = make_float2(ln_rstd, ln_rstd);
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float2 acc2 = make_float2(__uint_as_float(ra[j]), __uint_as_float(ra[j + 1]));
                  const float2 cs2 = *reinterpret_cast<const float2*>(scs + sl * 32 + j);
                  const float2 b2 = *reinterpret_cast<const float2*>(sbr + sl * 32 + j);
                  const float2 r2 = __ffma2_rn(rs2v, __ffma2_rn(nm, cs2, acc2), b2);
                  v[j] = r2.x;
                  v[j + 1] = r2.y;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float2 r2 = __fadd2_rn(make_float2(__uint_as_float(ra[j]), __uint_as_float(ra[j + 1])),
                                               *reinterpret_cast<const float2*>(sbr + sl * 32 + j));
                  v[j] = r2.x;
                  v[j + 1] = r2.y;
                }
              }
              if constexpr (EPI == 1) {
                if (ep.act == ACT_SILU) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
                } else if (ep.act == ACT_QUICKGELU) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = v[j] * rcp_approx(1.0f + ex2_approx(-2.4554669595930156f * v[j]));
                } else if (ep.act == ACT_GELU) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
                } else if (ep.act == ACT_RELU) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
              }
            }
          }
          if (has_res) {
            mbar_wait(&res_bar[half * 2 + slot], (slice_cnt >> 1) & 1);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const uint4 rv = *reinterpret_cast<const uint4*>(my_res_row + slot * kSliceBytes + ((u ^ swz) << 4));
              const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 r2 = __fadd2_rn(make_float2(v[u * 8 + 2 * i], v[u * 8 + 2 * i + 1]), __half22float2(rh[i]));
                v[u * 8 + 2 * i] = r2.x;
                v[u * 8 + 2 * i + 1] = r2.y;
              }
            }
          }
          if constexpr (kRowStat) {
            // columns past N hold exact zeros (zero-filled B rows, no bias, zero-filled residual): they add nothing
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 v2 = make_float2(v[j], v[j + 1]);
              rs2 = __fadd2_rn(rs2, v2);
              rq2 = __ffma2_rn(v2, v2, rq2);
            }
          }
          if constexpr (kGnStat) {
            // Statistics of the ROUNDED outputs (what the GroupNorm will read).  Per thread: (sum, sum of squares) of every
            // group segment inside this slice of its row, parked in this warp's 32 rows of the still unused output slot
            // [segment][lane]; then four lanes per segment add the 32 rows in a fixed order and the quad's total joins the
            // tile's group totals.  Everything is warp-private up to the tile's last barrier: no atomics, fixed order.
            uint32_t hw[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const __half2 h2 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
              hw[i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            float2* scratch = reinterpret_cast<float2*>(sOut + slot * kSliceBytes + q * (32 * 64));
            const uint32_t mask = p.gn_mask[sl];
            float2 a2 = make_float2(0.f, 0.f), q2 = make_float2(0.f, 0.f);
            int k = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
              a2 = __fadd2_rn(a2, f);
              q2 = __ffma2_rn(f, f, q2);
              if ((mask >> i) & 1u) {
                scratch[k * 32 + lane] = make_float2(a2.x + a2.y, q2.x + q2.y);
                ++k;
                a2 = make_float2(0.f, 0.f);
                q2 = make_float2(0.f, 0.f);
              }
            }
            if (!((mask >> 15) & 1u)) {
              scratch[k * 32 + lane] = make_float2(a2.x + a2.y, q2.x + q2.y);
              ++k;
            }
            __syncwarp();
            {
              const int seg = lane >> 2, part = lane & 3;
              float ts = 0.f, tq = 0.f;
              if (seg < k) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float2 e = scratch[seg * 32 + part * 8 + ((i + lane) & 7)];
                  ts += e.x;
                  tq += e.y;
                }
              }
              ts += __shfl_xor_sync(0xffffffffu, ts, 1);
              tq += __shfl_xor_sync(0xffffffffu, tq, 1);
              ts += __shfl_xor_sync(0xffffffffu, ts, 2);
              tq += __shfl_xor_sync(0xffffffffu, tq, 2);
              if (seg < k && part == 0) {
                float2* t2 = gn_tot + ((q * kGnTileGroups + p.gn_g0[sl] + seg) * 2 + (sl & 1));
                const float2 o = *t2;
                *t2 = make_float2(o.x + ts, o.y + tq);
              }
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < 4; ++u)
              *reinterpret_cast<uint4*>(my_out_row + slot * kSliceBytes + ((u ^ swz) << 4)) =
                  make_uint4(hw[4 * u], hw[4 * u + 1], hw[4 * u + 2], hw[4 * u + 3]);
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              __align__(16) __half2 h[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[u * 8 + 2 * i], v[u * 8 + 2 * i + 1]);
              *reinterpret_cast<uint4*>(my_out_row + slot * kSliceBytes + ((u ^ swz) << 4)) = *reinterpret_cast<uint4*>(h);
            }
          }
          fence_proxy_async();
          named_bar_sync(2 + half, 128);
          if (leader && !(p.debug & 1)) {
            if (CONV)
              tma_store_4d(&tmOut, sOut + slot * kSliceBytes, col_base + sl * 32, x0, y0, n0);
            else
              tma_store_2d(&tmOut, sOut + slot * kSliceBytes, col_base + sl * 32, m_tile * kBM);
            tma_store_commit();
          }
        }
        if constexpr (kRowStat) {
          if (valid)
            ep.rowstat_out[(static_cast<int64_t>(n_tile) * 2 + half) * (ep.rowstat_ld > 0 ? ep.rowstat_ld : p.M) + out_row] =
                make_float2(rs2.x + rs2.y, rq2.x + rq2.y);
        }
        if constexpr (kGnStat) {
          // every warp's slice totals are in: one thread per group of the tile adds the 4 quadrants x 2 slice parities in
          // a fixed order, leaves the entries zero for the next tile and writes the tile's partial
          named_bar_sync(1, kEpiThreads);
          const int ntg = min(bn_out, n_out - col_base) / p.gn_cpg;
          // (n0 >= Bn: the phantom second tile of the last CTA pair when the tile count is odd - its entries are still reset)
          const bool real_tile = n0 < p.Bn;
          if (etid < ntg) {
            float ts = 0.f, tq = 0.f;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              // the two slice parities are added to each other first: which parity a straddling group's halves land in
              // depends on the tile width, the sum of the pair does not (same partials at every BN)
              float2* t2 = gn_tot + (qq * kGnTileGroups + etid) * 2;
              ts += t2[0].x + t2[1].x;
              tq += t2[0].y + t2[1].y;
              t2[0] = make_float2(0.f, 0.f);
              t2[1] = make_float2(0.f, 0.f);
            }
            const int per_sample = p.tiles_x * p.tiles_y;
            const int chunk = m_tile % per_sample;
            float* dst = ep.gn_out + ((static_cast<int64_t>(n0) * ep.gn_nparts + chunk) * ep.gn_groups +
                                      col_base / p.gn_cpg + etid) * 2;
            if (real_tile) *reinterpret_cast<float2*>(dst) = make_float2(ts, tq);
          }
        }
      } else if constexpr (EPI == 1) {
        // ================================================= direct path (fp32 out, unaligned pitches, tiny N)
        const __half* rgb = (p.rgb_rows == 0 && valid && ep.rowgroup_bias)
                                ? ep.rowgroup_bias + (out_row / ep.rows_per_group) * ep.rgb_ld
                                : nullptr;
        const __half* res = (valid && ep.residual) ? ep.residual + out_row * ep.ldr : nullptr;
        named_bar_sync(1, kEpiThreads);
        mbar_wait(&tfull_bar[buf], use & 1);
        tc_fence_after();
        float best = -INFINITY;
        int best_idx = 0x7fffffff;
#pragma unroll 1
        for (int sl = half; sl < nsl; sl += 2) {
          float v[32];
          if (geglu) {
            uint32_t ra[32], rg[32];
            tmem_ld_32x32b_x32(lane_addr + sl * 32, ra);
            tmem_ld_32x32b_x32(lane_addr + BN / 2 + sl * 32, rg);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float a = __uint_as_float(ra[j]) + sbr[sl * 32 + j];
              const float gg = __uint_as_float(rg[j]) + sbr[BN / 2 + sl * 32 + j];
              v[j] = a * gelu_erf(gg);
            }
          } else {
            uint32_t ra[32];
            tmem_ld_32x32b_x32(lane_addr + sl * 32, ra);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]) + sbr[sl * 32 + j];
          }
          if (ep.act == ACT_ROWMAX) {
            const int col0 = col_base + sl * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < n_out && v[j] > best) {     // strict: the first (lowest) column wins ties
                best = v[j];
                best_idx = col0 + j;
              }
          } else if (valid) {
            const int col0 = col_base + sl * 32;
            if (rgb) {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < n_out) v[j] += __half2float(__ldg(rgb + col0 + j));
            }
            if (ep.act == ACT_SILU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
            } else if (ep.act == ACT_QUICKGELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = v[j] * rcp_approx(1.0f + ex2_approx(-2.4554669595930156f * v[j]));
            } else if (ep.act == ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
            } else if (ep.act == ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (res) {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < n_out) v[j] += __half2float(res[col0 + j]);
            }
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8)
              if (col0 + g8 * 8 < n_out) store8(ep, out_row, col0 + g8 * 8, n_out, v + g8 * 8);
          }
        }
        if (ep.act == ACT_ROWMAX && valid) {
          const int64_t o = out_row * ep.rowmax_ld + n_tile * 2 + half;
          ep.rowmax_val[o] = best;
          ep.rowmax_idx[o] = best_idx;
        }
      }
      if (wi.kind == 2 && etid == 0) p.sk_flags[sk_tt * npr + rank] = 0;   // zero again for the next launch
      tc_fence_before();
      release_accumulator(buf);
      wi = wnext;
      have = have_next;
    }
    if (leader) tma_store_wait_all();   // smem must outlive the last bulk stores
  }

  tc_fence_before();
  __syncthreads();
  if (p.mcast) cluster_sync_all();   // neither CTA leaves while the other can still signal its barriers
  if (warp == 1) {
    if constexpr (PAIR2) tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
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

// Kernel image for a launch.  EPI 2 (GEGLU) exists at BN = 256, GEMM only; pair images at BN >= 64.
template <int BN, bool CONV, bool PAIR2>
static auto kernel_for(int epi) -> decltype(&gemm_tc_kernel<BN, CONV, PAIR2, 1>) {
  if (epi == 0) return gemm_tc_kernel<BN, CONV, PAIR2, 0>;
  if constexpr (BN == 256 && !CONV) {
    if (epi == 2) return gemm_tc_kernel<BN, CONV, PAIR2, 2>;
    if (epi == 5) return gemm_tc_kernel<BN, CONV, PAIR2, 5>;
  }
  if constexpr (BN >= 64 && !CONV) {      // LayerNorm-fusion images (row statistics out / in): GEMM only
    if (epi == 3) return gemm_tc_kernel<BN, CONV, PAIR2, 3>;
    if (epi == 4) return gemm_tc_kernel<BN, CONV, PAIR2, 4>;
  }
  if constexpr (BN >= 128 && CONV) {      // GroupNorm statistics of the output: conv only
    if (epi == 6) return gemm_tc_kernel<BN, CONV, PAIR2, 6>;
  }
  return gemm_tc_kernel<BN, CONV, PAIR2, 1>;
}
// shared-memory footprint / ring depth of an image
template <int BN, bool PAIR2>
static void cfg_for(int epi, int* smem, int* stages) {
  if (epi == 1) {
    *smem = GemmCfg<BN, PAIR2, 1>::SMEM;
    *stages = GemmCfg<BN, PAIR2, 1>::STAGES;
  } else if (epi == 6) {
    *smem = GemmCfg<BN, PAIR2, 6>::SMEM;
    *stages = GemmCfg<BN, PAIR2, 6>::STAGES;
  } else {
    *smem = GemmCfg<BN, PAIR2, 0>::SMEM;
    *stages = GemmCfg<BN, PAIR2, 0>::STAGES;
  }
}
static int epi_class(const GemmParams& p) {
  if (!p.tma_epi) return 1;
  if (p.ep.ln_rowstat != nullptr) return p.ep.act == ACT_GEGLU ? 5 : 4;   // validated in gemm2_f16
  if (p.ep.rowstat_out != nullptr) return 3;
  if (p.ep.gn_out != nullptr) return 6;                                   // validated in conv_impl
  if (p.ep.act == ACT_NONE) return 0;
  if (p.ep.act == ACT_GEGLU && p.fast_gelu && !p.conv) return 2;
  return 1;
}

// One-time per tile width: opt in to the large dynamic shared memory and ask how many 2-CTA clusters of this kernel
// the device can hold at once (GPCs with an odd SM count strand one SM).
template <int BN>
static int prepare(int* max_clusters) {
  using Cfg = GemmCfg<BN>;
  static bool done = false;
  static int clusters = 0;
  if (!done) {
    for (int epi = 0; epi < 7; ++epi) {
      if ((epi == 2 || epi == 5) && BN != 256) continue;   // kernel_for aliases the all-variants image elsewhere
      if ((epi == 3 || epi == 4) && BN < 64) continue;
      if (epi == 6 && BN < 128) continue;
      int smem = 0, stages = 0;
      cfg_for<BN, false>(epi, &smem, &stages);
      // (EPI 6 exists for the convolution only: the GEMM lookup would alias the all-variants image)
      if (epi != 6)
        GYRE_CHECK_CUDA(cudaFuncSetAttribute(kernel_for<BN, false, false>(epi), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             smem));
      GYRE_CHECK_CUDA(cudaFuncSetAttribute(kernel_for<BN, true, false>(epi), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           smem));
      if constexpr (BN >= 64) {
        cfg_for<BN, true>(epi, &smem, &stages);
        if (epi != 6)
          GYRE_CHECK_CUDA(cudaFuncSetAttribute(kernel_for<BN, false, true>(epi), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               smem));
        GYRE_CHECK_CUDA(cudaFuncSetAttribute(kernel_for<BN, true, true>(epi), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             smem));
      }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * sm_count());
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = 2;
    a[0].val.clusterDim.y = 1;
    a[0].val.clusterDim.z = 1;
    cfg.attrs = a;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<BN, false, false, 1>, &cfg) == cudaSuccess) clusters = n;
    cudaGetLastError();
    done = true;
  }
  *max_clusters = clusters;
  return 0;
}

static int gemm_max_clusters(int bn, int* n) {
  switch (bn) {
    case 32: return prepare<32>(n);
    case 64: return prepare<64>(n);
    case 128: return prepare<128>(n);
    case 160: return prepare<160>(n);
    case 256: return prepare<256>(n);
  }
  *n = 0;
  return 0;
}

// Pair mode (see the kernel comment) when the tunable allows it, there are at least two M tiles and the device can
// co-schedule enough clusters.
static int want_pair_mode(int bn, long long m_tiles, int k_iters, int* mcast) {
  *mcast = 0;
  // 0 off, 1 multicast pairs, 2 cta_group::2 pairs where they win, 3 cta_group::2 pairs everywhere (measurement)
  int mode = tunable(TUNE_MCAST);
  if (mode <= 0 || mode > 3 || m_tiles < 2 || bn < 64) return 0;
  if (mode == 2 && k_iters < 10) return 0;   // short-K tiles: the pair's per-tile hand-offs cost more than the B half saves
  if (mode == 3) mode = 2;
  int n = 0;
  GYRE_TRY(gemm_max_clusters(bn, &n));
  if (n >= sm_count() / 4) *mcast = mode;
  return 0;
}

// Stream-K over the last, partial wave: worth it when the wave is far from full and K is long enough to cut.
static void plan_stream_k(GemmParams* p, int bn, int max_clusters) {
  p->sk_tiles = 0;
  p->sk_first = 0;
  p->sk_maxp = 0;
  p->sk_ws = nullptr;
  p->sk_flags = nullptr;
  const Epilogue& ep = p->ep;
  if (!tunable(TUNE_STREAMK) || p->mcast == 1 || !p->tma_epi || ep.sk_ws == nullptr || ep.sk_flags == nullptr) return;
  const int npr = p->mcast ? 2 : 1;                          // cta_group::2 pairs: the worker is a cluster
  // The hand-off costs ~5-10 us of epilogue time per launch (partial dump + fence + flag, then latency-bound partial
  // loads in the owner's epilogue - measured on B200), so only tiles whose main loop runs for tens of us qualify:
  // the 3x3 convolutions at 32x32 and below, not the transformer GEMMs (their 3-15 us tiles got slower).
  if (ep.act == ACT_ROWMAX || static_cast<long long>(p->k_iters) * bn < tunable(TUNE_SK_MIN)) return;
  const long long T = static_cast<long long>(npr == 2 ? (p->m_tiles + 1) / 2 : p->m_tiles) * p->n_tiles;
  const int G = npr == 2 ? max_clusters : sm_count();
  if (G <= 0) return;
  const long long full_waves = T / G;
  const int R = static_cast<int>(T % G);
  if (R == 0) return;
  // time in tile units: ceil(T/G) without, T/G with; require >= 6 % gain
  const double without = static_cast<double>(full_waves + 1), with = static_cast<double>(T) / G;
  // measured: the hand-off eats ~8 points of the theoretical gain, and with many waves the partial wave is a small share
  if ((without - with) / without < (npr == 2 ? 0.10 : 0.06) || (npr == 2 && without > 4.0)) return;
  const long long W = static_cast<long long>(R) * p->k_iters;
  const long long per = W / G;
  if (per < 4) return;                                   // segments too short to be worth a hand-off
  const int maxp = static_cast<int>(p->k_iters / per) + 2;
  const size_t need = static_cast<size_t>(R) * maxp * npr * kBM * bn * sizeof(float);
  if (need > ep.sk_ws_bytes || R * npr > ep.sk_flags_count) return;
  p->sk_tiles = R;
  p->sk_first = static_cast<int>(T - R);
  p->sk_maxp = maxp;
  p->sk_ws = static_cast<float*>(ep.sk_ws);
  p->sk_flags = ep.sk_flags;
}

template <int BN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                  const CUtensorMap& tmRes, const GemmParams& p_in, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static_assert(Cfg::STAGES >= 3, "pipeline too shallow");
  int max_clusters = 0;
  GYRE_TRY(prepare<BN>(&max_clusters));
  const int sms = sm_count();
  GemmParams p = p_in;
  p.debug = tunable(TUNE_DEBUG);
  const int epi = (p.debug & 4) ? 1 : epi_class(p);       // DEBUG bit 2: always the all-variants image (A/B)
  const int cap = tunable(TUNE_GEMM_STAGES);
  int smem = 0, max_stages = 0;
  cfg_for<BN, false>(epi, &smem, &max_stages);
  p.stages = (cap >= 2 && cap < max_stages) ? cap : max_stages;
  // stream-K owners spin on flags that other CTAs of the same grid raise: the grid must be co-resident as a whole, which
  // only a cooperative launch guarantees when the device is shared with other streams / processes
  const bool coop = p.sk_tiles > 0;
  if (p.mcast) {
    GYRE_REQUIRE(max_clusters > 0, "gemm: pair mode requested but clusters are unavailable");
    const long long pairs = static_cast<long long>((p.m_tiles + 1) / 2) * p.n_tiles;
    GYRE_REQUIRE(pairs > 0 && pairs < (1ll << 30), "gemm: bad tile count %lld", pairs);
    // stream-K needs every worker resident at once (owners spin on flags other workers raise)
    const unsigned clusters = p.sk_tiles > 0 ? static_cast<unsigned>(max_clusters)
                                             : static_cast<unsigned>(pairs < max_clusters ? pairs : max_clusters);
    if constexpr (BN >= 64) {
      if (p.mcast == 2) {
        cfg_for<BN, true>(epi, &smem, &max_stages);
        p.stages = (cap >= 2 && cap < max_stages) ? cap : max_stages;
        if (p.conv)
          return launch_kernel_cluster(kernel_for<BN, true, true>(epi), dim3(2 * clusters), dim3(kThreads), smem, st, 2,
                                       coop, tmA, tmA2, tmB, tmOut, tmRes, p);
        return launch_kernel_cluster(kernel_for<BN, false, true>(epi), dim3(2 * clusters), dim3(kThreads), smem, st, 2,
                                     coop, tmA, tmA2, tmB, tmOut, tmRes, p);
      }
    }
    GYRE_REQUIRE(p.mcast == 1, "gemm: pair mode %d is not available at BN %d", p.mcast, BN);
    if (p.conv)
      return launch_kernel_cluster(kernel_for<BN, true, false>(epi), dim3(2 * clusters), dim3(kThreads), smem, st, 2,
                                   coop, tmA, tmA2, tmB, tmOut, tmRes, p);
    return launch_kernel_cluster(kernel_for<BN, false, false>(epi), dim3(2 * clusters), dim3(kThreads), smem, st, 2,
                                 coop, tmA, tmA2, tmB, tmOut, tmRes, p);
  }
  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles;
  GYRE_REQUIRE(tiles > 0 && tiles < (1ll << 31), "gemm: bad tile count %lld", tiles);
  // stream-K needs every CTA resident at once (owners spin on flags other CTAs raise): one CTA per SM, whole device
  const unsigned blocks = p.sk_tiles > 0 ? static_cast<unsigned>(sms) : static_cast<unsigned>(tiles < sms ? tiles : sms);
  // the implicit-GEMM convolution is its own instantiation (no run-time branches in the producer / epilogue, and a
  // distinct kernel name in profiles)
  if (p.conv)
    return launch_kernel_cluster(kernel_for<BN, true, false>(epi), dim3(blocks), dim3(kThreads), smem, st, 1, coop, tmA,
                                 tmA2, tmB, tmOut, tmRes, p);
  return launch_kernel_cluster(kernel_for<BN, false, false>(epi), dim3(blocks), dim3(kThreads), smem, st, 1, coop, tmA,
                               tmA2, tmB, tmOut, tmRes, p);
}

// Tile width: maximise (SM wave efficiency) x (1 - N padding) x (per-tile efficiency of the shape).
static int pick_bn(long long m_tiles, int N, int act) {
  if (act == ACT_GEGLU) return 256;
  {
    const int f = tunable(TUNE_FORCE_BN);   // measurement only
    if (f == 32 || f == 64 || f == 128 || f == 160 || f == 256) return f;
  }
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

// Tile width of a convolution.  Small maps with wide channels (the 8x8 / 16x16 levels: <= 32 M tiles, K in the thousands)
// do not fill a wave of 128 x BN tiles whatever BN is, so with stream-K available - it cuts the K range of the tiles over
// all SMs - the wave term of pick_bn is moot and the widest tile that divides N wins (fewest operand bytes per flop:
// measured 16x8x8 1280->1280 45 -> 35 us, 16x16x16 1280->1280 95 -> 85 us, profiles/r02_ab_conv_smallmap.txt).
static int pick_bn_conv(long long m_tiles, int N, int k_iters, bool streamk_avail) {
  const int plain = pick_bn(m_tiles, N, ACT_NONE);
  if (!streamk_avail || !tunable(TUNE_STREAMK) || tunable(TUNE_FORCE_BN) != 0 || tunable(TUNE_MCAST) < 2) return plain;
  if (m_tiles < 4 || m_tiles > 32 || k_iters < 10) return plain;
  const long long pairs = (m_tiles + 1) / 2;
  int cand = 0;
  if (pairs >= 8 && N % 256 == 0) cand = 256;
  else if (N % 160 == 0) cand = 160;
  else if (N % 128 == 0) cand = 128;
  if (cand == 0 || static_cast<long long>(k_iters) * cand < tunable(TUNE_SK_MIN)) return plain;   // plan_stream_k's own floor
  // the wide tile only pays if plan_stream_k will really cut its partial wave: same conditions, evaluated up front
  int G = 0;
  if (gemm_max_clusters(cand, &G) != 0 || G < sm_count() / 4) return plain;
  const long long T = pairs * (N / cand);
  const long long R = T % G;
  if (R != 0) {
    const double without = static_cast<double>(T / G + 1), with = static_cast<double>(T) / G;
    if ((without - with) / without < 0.10 || without > 4.0 || R * k_iters / G < 4) return plain;
  }
  return cand;
}

static int dispatch(int bn, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                    const CUtensorMap& tmOut, const CUtensorMap& tmRes, const GemmParams& p, cudaStream_t st) {
  switch (bn) {
    case 32: return launch<32>(tmA, tmA2, tmB, tmOut, tmRes, p, st);
    case 64: return launch<64>(tmA, tmA2, tmB, tmOut, tmRes, p, st);
    case 128: return launch<128>(tmA, tmA2, tmB, tmOut, tmRes, p, st);
    case 160: return launch<160>(tmA, tmA2, tmB, tmOut, tmRes, p, st);
    case 256: return launch<256>(tmA, tmA2, tmB, tmOut, tmRes, p, st);
  }
  set_last_error("gemm: unsupported BN %d", bn);
  return -2;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Can output / residual go through the TMA slices?  (fp16, 16-byte aligned rows)
static bool tma_epilogue_ok(const Epilogue& ep, int n_out) {
  if (ep.out_mode != OUT_F16 || ep.out == nullptr || (ep.ldo & 7) != 0 || !aligned16(ep.out)) return false;
  if (ep.residual != nullptr && ((ep.ldr & 7) != 0 || !aligned16(ep.residual))) return false;
  return n_out >= 8;
}

int gemm_rowmax_partials(int M, int N) {
  const int bn = pick_bn((M + kBM - 1) / kBM, N, ACT_ROWMAX);
  return 2 * ((N + bn - 1) / bn);
}

int gemm_rowstat_parts(int M, int N) {
  const int bn = pick_bn((M + kBM - 1) / kBM, N, ACT_NONE);
  return 2 * ((N + bn - 1) / bn);
}

int gemm_f16(const __half* A, int lda, const __half* W, int ldw, int M, int N, int K, const Epilogue& ep,
             cudaStream_t st) {
  return gemm2_f16(A, lda, K, nullptr, 0, 0, W, ldw, M, N, ep, st);
}

int gemm2_f16(const __half* A, int lda, int K1, const __half* A2, int lda2, int K2, const __half* W, int ldw, int M,
              int N, const Epilogue& ep, cudaStream_t st) {
  const int K = K1 + K2;
  GYRE_REQUIRE(M > 0 && N > 0 && K1 > 0 && K2 >= 0, "gemm: empty problem %dx%dx%d", M, N, K);
  GYRE_REQUIRE(K2 == 0 || (A2 != nullptr && K1 % kBK == 0 && lda2 % 8 == 0 && aligned16(A2)),
               "gemm: two-source A needs K1 %% 64 == 0 and a 16B-aligned second source (K1=%d)", K1);
  GYRE_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "gemm: lda/ldw must be multiples of 8 halfs (TMA 16B pitch)");
  GYRE_REQUIRE(aligned16(A) && aligned16(W), "gemm: operands must be 16B aligned");
  if (ep.act == ACT_ROWMAX)
    GYRE_REQUIRE(ep.rowmax_val && ep.rowmax_idx && ep.rowmax_ld >= gemm_rowmax_partials(M, N), "gemm: bad rowmax buffers");
  else
    GYRE_REQUIRE(ep.out != nullptr && ep.ldo > 0, "gemm: null output");
  const int bn = pick_bn((M + kBM - 1) / kBM, N, ep.act);
  if (ep.act == ACT_GEGLU) GYRE_REQUIRE(N % 256 == 0, "gemm: GEGLU needs N %% 256 == 0 (got %d)", N);
  const int n_out = ep.act == ACT_GEGLU ? N / 2 : N;
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.k1_iters = (K1 + kBK - 1) / kBK;
  p.k_iters = p.k1_iters + (K2 + kBK - 1) / kBK;
  p.n_tiles = (N + bn - 1) / bn;
  p.m_tiles = (M + kBM - 1) / kBM;
  p.conv = 0;
  p.fast_gelu = tunable(TUNE_GELU_FAST);
  p.mcast = 0;
  p.ep = ep;
  p.rgb_rows = 0;
  p.tma_epi = (tma_epilogue_ok(ep, n_out) && ep.rowgroup_bias == nullptr) ? 1 : 0;
  if (ep.rowstat_out != nullptr)
    GYRE_REQUIRE(p.tma_epi && ep.act == ACT_NONE && bn >= 64 && ep.ln_rowstat == nullptr,
                 "gemm: row statistics need the fp16 TMA epilogue without activation (N=%d, ldo=%d)", N, ep.ldo);
  if (ep.ln_rowstat != nullptr)
    GYRE_REQUIRE(p.tma_epi && ep.ln_colsum != nullptr && bn >= 64 && ep.ln_parts >= 0 && ep.ln_parts <= 8 &&
                     (ep.ln_parts == 0 || ep.ln_inv_c > 0.f) &&
                     (ep.act == ACT_NONE || (ep.act == ACT_GEGLU && p.fast_gelu)),
                 "gemm: the folded LayerNorm needs the fp16 TMA epilogue, column sums and no activation but the fast GEGLU");
  GYRE_TRY(want_pair_mode(bn, p.m_tiles, p.k_iters, &p.mcast));
  {
    int max_clusters = 0;
    if (p.mcast) GYRE_TRY(gemm_max_clusters(bn, &max_clusters));
    plan_stream_k(&p, bn, max_clusters);
  }
  CUtensorMap tmA, tmA2, tmB, tmOut, tmRes;
  uint32_t es[2] = {1, 1};
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K1), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {kBK, kBM};
    GYRE_TRY(encode_tmap_f16(&tmA, A, 2, dims, strides, box, es, true));
    tmA2 = tmA;
  }
  if (K2 > 0) {
    uint64_t dims[2] = {static_cast<uint64_t>(K2), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda2) * 2};
    uint32_t box[2] = {kBK, kBM};
    GYRE_TRY(encode_tmap_f16(&tmA2, A2, 2, dims, strides, box, es, true));
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {kBK, static_cast<uint32_t>(p.mcast ? bn / 2 : bn)};   // pair mode: each CTA fetches half a B tile
    GYRE_TRY(encode_tmap_f16(&tmB, W, 2, dims, strides, box, es, true));
  }
  tmOut = tmA;
  tmRes = tmA;
  if (p.tma_epi) {
    uint64_t dims[2] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(M)};
    uint32_t box[2] = {32, kBM};
    uint64_t so[1] = {static_cast<uint64_t>(ep.ldo) * 2};
    GYRE_TRY(encode_tmap_f16_sw(&tmOut, ep.out, 2, dims, so, box, es, 64));
    tmRes = tmOut;
    if (ep.residual) {
      uint64_t sr[1] = {static_cast<uint64_t>(ep.ldr) * 2};
      GYRE_TRY(encode_tmap_f16_sw(&tmRes, ep.residual, 2, dims, sr, box, es, 64));
    }
  }
  prof::Scope ps(prof::F_GEMM, 2.0 * M * static_cast<double>(N) * K,
                 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K +
                        static_cast<double>(M) * n_out * (ep.residual ? 2 : 1)),
                 st);
  return dispatch(bn, tmA, tmA2, tmB, tmOut, tmRes, p, st);
}

static inline int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// The 128-pixel output patch (tile_w x tile_h pixels of tile_n images) with the least padded work.
static bool conv_tile_shape(int Ho, int Wo, int B, int stride, int* tw_out, int* th_out, int* tn_out) {
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
  *tw_out = best_w;
  *th_out = best_h;
  *tn_out = best_n;
  return best_cost > 0;
}

int conv3x3_out_extent(int n, int stride, int pad) {
  return (stride == 1) ? n : (pad == 1 ? (n - 1) / 2 + 1 : (n + 1 - 3) / 2 + 1);
}

// GroupNorm statistics in the epilogue (Epilogue::gn_out): tiles must partition each sample exactly (one image per tile,
// no overhang - padded rows would carry the bias), group boundaries must coincide with tile boundaries, a slice may not
// hold more than 8 group segments (cpg >= 4, even: column pairs never straddle a group).
static int conv_gn_parts(int B, int Ho, int Wo, int Cout, int stride, int groups, int* bn_out) {
  if (!tunable(TUNE_GN_FUSE) || groups <= 0 || Cout % groups != 0 || (Cout & 7) != 0) return 0;
  const int cpg = Cout / groups;
  if (cpg < 4 || (cpg & 1) != 0) return 0;
  int tw = 0, th = 0, tn = 0;
  if (!conv_tile_shape(Ho, Wo, B, stride, &tw, &th, &tn)) return 0;
  if (tn != 1 || Wo % tw != 0 || Ho % th != 0) return 0;
  const int m_tiles = (Wo / tw) * (Ho / th) * B;
  const int bn = pick_bn(m_tiles, Cout, ACT_NONE);
  if (bn < 128 || bn % cpg != 0 || bn / cpg > kGnTileGroups) return 0;
  if (bn_out) *bn_out = bn;
  return (Wo / tw) * (Ho / th);
}

int conv3x3_gn_parts(int B, int H, int W, int Cout, int stride, int pad, int groups) {
  return conv_gn_parts(B, conv3x3_out_extent(H, stride, pad), conv3x3_out_extent(W, stride, pad), Cout, stride, groups, nullptr);
}

// Geometry of one implicit-GEMM convolution launch: taps_h x taps_w taps whose (0, 0) tap reads input pixel
// (x * stride + off_x, y * stride + off_y); output pixel (x, y) of the Ho x Wo grid is written to pixel
// (x * out_step + out_ox, y * out_step + out_oy) of a [B, Ho*out_step, Wo*out_step, ldo] tensor.
struct ConvGeom {
  int taps_h = 3, taps_w = 3;
  int off_x = -1, off_y = -1;
  int Ho = 0, Wo = 0;
  int out_step = 1, out_ox = 0, out_oy = 0;
  double algo_flops = 0.0, algo_bytes = 0.0;
};

static int conv_impl(const __half* X, int ldx, int B, int H, int W, int Cin, const __half* Wp, int Cout, int stride,
                     const ConvGeom& gm, const Epilogue& ep, cudaStream_t st) {
  const int Ho = gm.Ho, Wo = gm.Wo;
  const int ntaps = gm.taps_h * gm.taps_w;
  int best_w = 0, best_h = 0, best_n = 0;
  const long long best_cost = conv_tile_shape(Ho, Wo, B, stride, &best_w, &best_h, &best_n) ? 1 : -1;
  GYRE_REQUIRE(best_cost > 0, "conv3x3: no tile shape for %dx%d", Ho, Wo);
  const int m_tiles = ((Wo + best_w - 1) / best_w) * ((Ho + best_h - 1) / best_h) * ((B + best_n - 1) / best_n);
  GemmParams p{};
  p.M = 0;
  p.N = Cout;
  p.cin_chunks = (Cin + kBK - 1) / kBK;
  p.cin_pad = p.cin_chunks * kBK;
  p.k_iters = ntaps * p.cin_chunks;
  // (a launch that also produces GroupNorm statistics keeps the width conv3x3_gn_parts planned with)
  const bool sk_avail = ep.sk_ws != nullptr && ep.sk_flags != nullptr && ep.gn_out == nullptr && ep.act != ACT_ROWMAX &&
                        tma_epilogue_ok(ep, Cout);
  const int bn = pick_bn_conv(m_tiles, Cout, p.k_iters, sk_avail);
  p.k1_iters = p.k_iters;
  p.n_tiles = (Cout + bn - 1) / bn;
  p.m_tiles = m_tiles;
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
  p.taps_w = gm.taps_w;
  p.off_x = gm.off_x;
  p.off_y = gm.off_y;
  p.fast_gelu = 0;
  p.mcast = 0;
  p.ep = ep;
  // the per-sample bias (temb projection) is uniform over the rows of one image inside a tile: stage it with
  // the column bias when a tile holds at most kMaxBiasGroups images
  p.rgb_rows = 0;
  if (ep.rowgroup_bias != nullptr && ep.rows_per_group == Ho * Wo && best_n <= kMaxBiasGroups)
    p.rgb_rows = best_w * best_h;
  p.tma_epi = (tma_epilogue_ok(ep, Cout) && (ep.rowgroup_bias == nullptr || p.rgb_rows > 0)) ? 1 : 0;
  if (ep.gn_out != nullptr) {
    int bn_gn = 0;
    const int parts = conv_gn_parts(B, Ho, Wo, Cout, stride, ep.gn_groups, &bn_gn);
    GYRE_REQUIRE(parts > 0 && parts == ep.gn_nparts && bn_gn == bn && p.tma_epi && ep.act == ACT_NONE && gm.out_step == 1,
                 "conv3x3: GroupNorm statistics cannot be produced here (%dx%d, Cout %d, %d groups, %d parts given)", Ho, Wo,
                 Cout, ep.gn_groups, ep.gn_nparts);
    const int cpg = Cout / ep.gn_groups;
    p.gn_cpg = cpg;
    for (int sl = 0; sl < 8; ++sl) {
      const int c0 = sl * 32;                       // tile-relative first column of the slice (tiles start on a group)
      uint32_t mask = 0;
      for (int pos = cpg - c0 % cpg; pos <= 32; pos += cpg) mask |= 1u << (pos / 2 - 1);
      p.gn_mask[sl] = mask;
      p.gn_g0[sl] = static_cast<uint8_t>(c0 / cpg);
    }
  }
  GYRE_TRY(want_pair_mode(bn, m_tiles, p.k_iters, &p.mcast));
  {
    int max_clusters = 0;
    if (p.mcast) GYRE_TRY(gemm_max_clusters(bn, &max_clusters));
    plan_stream_k(&p, bn, max_clusters);
  }
  GYRE_REQUIRE(gm.out_step == 1 || (p.tma_epi && ep.residual == nullptr),
               "conv: a strided output view needs the TMA epilogue (fp16, 16B-aligned rows, no residual)");
  CUtensorMap tmA, tmB, tmOut, tmRes;
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
    uint64_t dims[2] = {static_cast<uint64_t>(ntaps) * p.cin_pad, static_cast<uint64_t>(Cout)};
    uint64_t strides[1] = {static_cast<uint64_t>(ntaps) * p.cin_pad * 2};
    uint32_t box[2] = {kBK, static_cast<uint32_t>(p.mcast ? bn / 2 : bn)};
    uint32_t es[2] = {1, 1};
    GYRE_TRY(encode_tmap_f16(&tmB, Wp, 2, dims, strides, box, es, true));
  }
  tmOut = tmA;
  tmRes = tmA;
  if (p.tma_epi) {
    uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                        static_cast<uint64_t>(B)};
    uint32_t box[4] = {32, static_cast<uint32_t>(best_w), static_cast<uint32_t>(best_h), static_cast<uint32_t>(best_n)};
    uint32_t es[4] = {1, 1, 1, 1};
    const uint64_t step = static_cast<uint64_t>(gm.out_step);
    const uint64_t pix = static_cast<uint64_t>(ep.ldo) * 2;            // bytes per output pixel
    uint64_t so[3] = {pix * step, pix * (Wo * step) * step, pix * (Wo * step) * (Ho * step)};
    const __half* obase = static_cast<const __half*>(ep.out) +
                          (static_cast<size_t>(gm.out_oy) * Wo * step + gm.out_ox) * ep.ldo;
    GYRE_TRY(encode_tmap_f16_sw(&tmOut, obase, 4, dims, so, box, es, 64));
    tmRes = tmOut;
    if (ep.residual) {
      uint64_t sr[3] = {static_cast<uint64_t>(ep.ldr) * 2, static_cast<uint64_t>(ep.ldr) * 2 * Wo,
                        static_cast<uint64_t>(ep.ldr) * 2 * Wo * Ho};
      GYRE_TRY(encode_tmap_f16_sw(&tmRes, ep.residual, 4, dims, sr, box, es, 64));
    }
  }
  prof::Scope ps(prof::F_CONV, gm.algo_flops, gm.algo_bytes, st);
  return dispatch(bn, tmA, tmA, tmB, tmOut, tmRes, p, st);
}

int conv3x3_f16(const __half* X, int ldx, int B, int H, int W, int Cin, const __half* Wp, int Cout, int stride,
                int pad, const Epilogue& ep, cudaStream_t st) {
  GYRE_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3x3: empty problem");
  GYRE_REQUIRE(stride == 1 || stride == 2, "conv3x3: stride %d", stride);
  GYRE_REQUIRE(ldx % 8 == 0 && Cin % 8 == 0, "conv3x3: channel pitch must be a multiple of 8");
  GYRE_REQUIRE(ep.act != ACT_GEGLU, "conv3x3: GEGLU epilogue not supported");
  GYRE_REQUIRE(ep.out != nullptr && ep.ldo > 0, "conv3x3: null output");
  ConvGeom gm;
  gm.off_x = gm.off_y = -pad;
  gm.Ho = conv3x3_out_extent(H, stride, pad);
  gm.Wo = conv3x3_out_extent(W, stride, pad);
  gm.algo_flops = 2.0 * 9 * Cin * Cout * static_cast<double>(B) * gm.Ho * gm.Wo;
  gm.algo_bytes = 2.0 * (static_cast<double>(B) * H * W * Cin + 9.0 * Cin * Cout +
                         static_cast<double>(B) * gm.Ho * gm.Wo * Cout * (ep.residual ? 2 : 1));
  return conv_impl(X, ldx, B, H, W, Cin, Wp, Cout, stride, gm, ep, st);
}

int upconv2x_f16(const __half* X, int ldx, int B, int H, int W, int Cin, const __half* Wp4, int Cout,
                 const Epilogue& ep, cudaStream_t st) {
  GYRE_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "upconv2x: empty problem");
  GYRE_REQUIRE(ldx % 8 == 0 && Cin % 8 == 0, "upconv2x: channel pitch must be a multiple of 8");
  GYRE_REQUIRE(ep.out != nullptr && ep.ldo > 0 && ep.act == ACT_NONE && ep.residual == nullptr &&
                   ep.rowgroup_bias == nullptr,
               "upconv2x: only a bias epilogue is supported");
  const int cin_pad = (Cin + kBK - 1) / kBK * kBK;
  const size_t phase_elems = static_cast<size_t>(Cout) * 4 * cin_pad;
  for (int phase = 0; phase < 4; ++phase) {
    const int dy = phase >> 1, dx = phase & 1;
    ConvGeom gm;
    gm.taps_h = gm.taps_w = 2;
    gm.off_x = dx - 1;
    gm.off_y = dy - 1;
    gm.Ho = H;
    gm.Wo = W;
    gm.out_step = 2;
    gm.out_ox = dx;
    gm.out_oy = dy;
    // algorithmic work of the un-folded op (9 taps on the upsampled grid), a quarter per phase
    gm.algo_flops = 2.0 * 9 * Cin * Cout * static_cast<double>(B) * H * W;
    gm.algo_bytes = 2.0 * (static_cast<double>(B) * H * W * Cin * 0.25 + 9.0 * Cin * Cout * 0.25 +
                           static_cast<double>(B) * H * W * Cout);
    GYRE_TRY(conv_impl(X, ldx, B, H, W, Cin, Wp4 + phase * phase_elems, Cout, 1, gm, ep, st));
  }
  return 0;
}

}  // namespace gyre
