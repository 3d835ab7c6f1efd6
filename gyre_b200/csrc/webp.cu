// Lossless WebP (VP8L) encoding on the device (the other artifact format of the step after the decode).
//
// The reference: `cv.imencode(".webp", image, [cv.IMWRITE_WEBP_QUALITY, 500])` on the host (quality > 100 = lossless;
// gyre/images.py:125-135 toWebpBytes, chosen by gyre/services/generate.py:73-76 when the client accepts image/webp).
// Lossless, so the contract is the decoded image, not libwebp's byte stream.
//
// Stream ("WebP Lossless Bitstream Specification"): RIFF / WEBP / VP8L; 0x2f, 14-bit width - 1, 14-bit height - 1, alpha flag,
// version 0; ONE transform - the predictor transform with the whole image in mode 12, clamp(L + T - TL) per channel (block
// size 2^9: the mode image is a handful of identical pixels behind zero-bit prefix codes); no colour cache, no meta prefix
// codes; prefix codes from the residuals' own histograms (length-limited canonical Huffman, sent with a flat 4-bit
// code-length code; one- and two-symbol alphabets as "simple" codes - a constant alpha channel costs zero bits per pixel); then
// the green, red, blue, alpha codes of every pixel in scan order, LSB-first.  Literal-only (no LZ77 / colour cache): ~15 %
// larger than libwebp's lossless files on photographic content.
//
// VP8L is ONE bit stream without restart points, so unlike the PNG encoder (png.cu) the chunks cannot be byte-aligned:
//   webp_residual_kernel   residuals (packed ARGB) + per-channel histograms (shared-memory counts, integer atomics)
//   webp_codes_kernel      one CTA per (image, channel): sort, two-queue Huffman, length limit, canonical codes
//   webp_count_kernel      bits per chunk of pixels
//   webp_header_kernel     one CTA per image: header bits, exclusive scan of the chunk sizes -> bit offsets, RIFF sizes
//   webp_emit_kernel       codes OR-ed into the zeroed output at their global bit offsets (atomics on the boundary words)
// oracle/webp.py restates this byte for byte and is pinned by Pillow's libwebp decoding its output.
#include "common.cuh"
#include "ops.h"

namespace gyre {

namespace {

constexpr int kThreads = 256;
constexpr int kChunkPixels = 4096;
constexpr int kMaxBits = 15;
constexpr int kPredMode = 12;
constexpr int kSizeBits = 9;
constexpr int kRiffBytes = 20;          // "RIFF" size "WEBP" "VP8L" size

struct CodeTab {
  uint16_t code[256];
  uint8_t len[256];
  int32_t kind;          // 1: one symbol, 2: two symbols, 0: normal
  int32_t sym0, sym1;
  int32_t pad;
};

__device__ __forceinline__ uint32_t rev_bits(uint32_t v, int n) { return n ? (__brev(v) >> (32 - n)) : 0u; }

// OR `nbits` (<= 32) bits into the zero-initialised stream of 32-bit words at bit position `pos`
__device__ __forceinline__ void or_bits(uint32_t* words, uint64_t pos, uint32_t value, int nbits) {
  if (nbits == 0) return;
  const uint32_t sh = static_cast<uint32_t>(pos & 31u);
  const uint64_t v = static_cast<uint64_t>(value) << sh;
  atomicOr(&words[pos >> 5], static_cast<uint32_t>(v));
  if (sh + nbits > 32) atomicOr(&words[(pos >> 5) + 1], static_cast<uint32_t>(v >> 32));
}

__device__ __forceinline__ uint32_t load_px(const uint8_t* img, int64_t idx, int C) {
  // (R, G, B, A) bytes of pixel idx; A = 255 without an alpha channel
  const uint8_t* p = img + idx * C;
  return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) |
         ((C == 4 ? static_cast<uint32_t>(p[3]) : 255u) << 24);
}

__device__ __forceinline__ uint32_t clamp_add_sub(uint32_t l, uint32_t t, uint32_t tl) {
  uint32_t out = 0;
#pragma unroll
  for (int s = 0; s < 32; s += 8) {
    int v = static_cast<int>((l >> s) & 255u) + static_cast<int>((t >> s) & 255u) - static_cast<int>((tl >> s) & 255u);
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    out |= static_cast<uint32_t>(v) << s;
  }
  return out;
}

__device__ __forceinline__ uint32_t sub_bytes(uint32_t a, uint32_t b) {
  uint32_t out = 0;
#pragma unroll
  for (int s = 0; s < 32; s += 8) out |= (((a >> s) - (b >> s)) & 255u) << s;
  return out;
}

// grid (chunks, B): residuals [B][H * W] packed (R, G, B, A) + hist [B][4][256] in the order g, r, b, a
__global__ void __launch_bounds__(kThreads) webp_residual_kernel(const uint8_t* __restrict__ img, int H, int W, int C,
                                                                 uint32_t* __restrict__ res, int* __restrict__ hist) {
  __shared__ int sh[4 * 256];
  for (int i = threadIdx.x; i < 4 * 256; i += kThreads) sh[i] = 0;
  __syncthreads();
  const int b = blockIdx.y;
  const int64_t n = static_cast<int64_t>(H) * W;
  const uint8_t* src = img + static_cast<int64_t>(b) * n * C;
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * kChunkPixels;
  const int64_t p1 = min(n, p0 + kChunkPixels);
  for (int64_t p = p0 + threadIdx.x; p < p1; p += kThreads) {
    const int x = static_cast<int>(p % W), y = static_cast<int>(p / W);
    const uint32_t px = load_px(src, p, C);
    uint32_t pred;
    if (y == 0)
      pred = x == 0 ? 0xFF000000u : load_px(src, p - 1, C);
    else if (x == 0)
      pred = load_px(src, p - W, C);
    else
      pred = clamp_add_sub(load_px(src, p - 1, C), load_px(src, p - W, C), load_px(src, p - W - 1, C));
    const uint32_t r = sub_bytes(px, pred);
    res[static_cast<int64_t>(b) * n + p] = r;
    atomicAdd(&sh[0 * 256 + ((r >> 8) & 255u)], 1);     // green
    atomicAdd(&sh[1 * 256 + (r & 255u)], 1);            // red
    atomicAdd(&sh[2 * 256 + ((r >> 16) & 255u)], 1);    // blue
    atomicAdd(&sh[3 * 256 + (r >> 24)], 1);             // alpha
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 4 * 256; i += kThreads)
    if (sh[i]) atomicAdd(&hist[b * 4 * 256 + i], sh[i]);
}

// grid (4, B): the prefix code of one channel of one image
__global__ void __launch_bounds__(kThreads) webp_codes_kernel(const int* __restrict__ hist, CodeTab* __restrict__ tabs) {
  __shared__ unsigned long long keys[512];
  __shared__ uint32_t leaf_w[260], int_w[260];
  __shared__ uint16_t par_leaf[260], par_int[260], depth_int[260];
  __shared__ int cnt[64];
  __shared__ int n_used;
  const int tid = threadIdx.x;
  const int* f = hist + (blockIdx.y * 4 + blockIdx.x) * 256;
  CodeTab* T = tabs + blockIdx.y * 4 + blockIdx.x;
  if (tid == 0) n_used = 0;
  for (int i = tid; i < 512; i += kThreads) {
    const unsigned long long fr = i < 256 ? static_cast<unsigned long long>(f[i]) : 0ull;
    keys[i] = fr ? ((fr << 9) | static_cast<unsigned long long>(i)) : ~0ull;
  }
  T->code[tid] = 0;
  T->len[tid] = 0;
  __syncthreads();
  if (f[tid]) atomicAdd(&n_used, 1);
  for (int size = 2; size <= 512; size <<= 1) {
    for (int strd = size >> 1; strd > 0; strd >>= 1) {
      const int i = 2 * tid - (tid & (strd - 1));
      const int j = i + strd;
      const bool up_dir = (i & size) == 0;
      const unsigned long long a = keys[i], c = keys[j];
      if ((a > c) == up_dir) {
        keys[i] = c;
        keys[j] = a;
      }
      __syncthreads();
    }
  }
  if (tid != 0) return;
  const int n = n_used;
  T->kind = 0;
  T->sym0 = T->sym1 = 0;
  T->pad = 0;
  if (n == 1) {
    T->kind = 1;
    T->sym0 = static_cast<int>(keys[0] & 511ull);
    return;
  }
  if (n == 2) {
    // the lower symbol value gets code 0 (oracle: used[0] < used[1])
    const int a = static_cast<int>(keys[0] & 511ull), c = static_cast<int>(keys[1] & 511ull);
    T->kind = 2;
    T->sym0 = a < c ? a : c;
    T->sym1 = a < c ? c : a;
    T->len[T->sym0] = 1;
    T->len[T->sym1] = 1;
    T->code[T->sym1] = 1;
    return;
  }
  for (int i = 0; i < n; ++i) leaf_w[i] = static_cast<uint32_t>(keys[i] >> 9);
  for (int i = 0; i < 64; ++i) cnt[i] = 0;
  int li = 0, ii = 0, made = 0;
  for (int kk = 0; kk < n - 1; ++kk) {       // two-queue Huffman; on equal weights the leaf goes first
    uint32_t w = 0;
    for (int t = 0; t < 2; ++t) {
      if (li < n && (ii >= made || leaf_w[li] <= int_w[ii])) {
        w += leaf_w[li];
        par_leaf[li++] = static_cast<uint16_t>(kk);
      } else {
        w += int_w[ii];
        par_int[ii++] = static_cast<uint16_t>(kk);
      }
    }
    int_w[made++] = w;
  }
  depth_int[n - 2] = 0;
  for (int j = n - 3; j >= 0; --j) depth_int[j] = static_cast<uint16_t>(depth_int[par_int[j]] + 1);
  for (int i = 0; i < n; ++i) {
    const int d = depth_int[par_leaf[i]] + 1;
    cnt[d < kMaxBits ? d : kMaxBits] += 1;
  }
  int total = 0;
  for (int l = 1; l <= kMaxBits; ++l) total += cnt[l] << (kMaxBits - l);
  while (total > (1 << kMaxBits)) {
    cnt[kMaxBits] -= 1;
    for (int l = kMaxBits - 1; l > 0; --l)
      if (cnt[l]) {
        cnt[l] -= 1;
        cnt[l + 1] += 2;
        break;
      }
    total -= 1;
  }
  int idx = 0;
  for (int l = kMaxBits; l > 0; --l)
    for (int c = 0; c < cnt[l]; ++c) T->len[keys[idx++] & 511ull] = static_cast<uint8_t>(l);
  uint32_t next[kMaxBits + 2];
  uint32_t code = 0;
  next[0] = 0;
  for (int l = 1; l <= kMaxBits; ++l) {
    code = (code + (l > 1 ? static_cast<uint32_t>(cnt[l - 1]) : 0u)) << 1;
    next[l] = code;
  }
  for (int s = 0; s < 256; ++s) {
    const int l = T->len[s];
    if (l) T->code[s] = static_cast<uint16_t>(rev_bits(next[l]++, l));
  }
}

__device__ __forceinline__ int px_bits(const CodeTab* T, uint32_t r) {
  return T[0].len[(r >> 8) & 255u] + T[1].len[r & 255u] + T[2].len[(r >> 16) & 255u] + T[3].len[r >> 24];
}

// grid (chunks, B): bits of each chunk of pixels
__global__ void __launch_bounds__(kThreads) webp_count_kernel(const uint32_t* __restrict__ res, int64_t n, const CodeTab* __restrict__ tabs,
                                                              unsigned long long* __restrict__ chunk_bits) {
  __shared__ uint8_t lens[4][256];
  __shared__ uint32_t red[kThreads / 32];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 4 * 256; i += kThreads) lens[i >> 8][i & 255] = tabs[b * 4 + (i >> 8)].len[i & 255];
  __syncthreads();
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * kChunkPixels;
  const int64_t p1 = min(n, p0 + kChunkPixels);
  uint32_t bits = 0;
  for (int64_t p = p0 + threadIdx.x; p < p1; p += kThreads) {
    const uint32_t r = res[b * n + p];
    bits += lens[0][(r >> 8) & 255u] + lens[1][r & 255u] + lens[2][(r >> 16) & 255u] + lens[3][r >> 24];
  }
  for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = bits;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    chunk_bits[static_cast<int64_t>(b) * gridDim.x + blockIdx.x] = t;
  }
}

__device__ void put_simple(uint32_t* words, uint64_t& pos, int symbol) {
  or_bits(words, pos, 1u, 1);       // simple code, num_symbols - 1 = 0
  pos += 2;
  if (symbol < 2) {
    or_bits(words, pos + 1, static_cast<uint32_t>(symbol), 1);
    pos += 2;
  } else {
    or_bits(words, pos, 1u, 1);
    or_bits(words, pos + 1, static_cast<uint32_t>(symbol), 8);
    pos += 9;
  }
}

__device__ void put_code(uint32_t* words, uint64_t& pos, const CodeTab& T, int alphabet) {
  if (T.kind == 1) {
    put_simple(words, pos, T.sym0);
    return;
  }
  if (T.kind == 2) {
    or_bits(words, pos, 7u, 3);     // simple, two symbols, first symbol in 8 bits
    or_bits(words, pos + 3, static_cast<uint32_t>(T.sym0), 8);
    or_bits(words, pos + 11, static_cast<uint32_t>(T.sym1), 8);
    pos += 19;
    return;
  }
  // normal code: 0, 4 bits (19 - 4), 19 x 3 bits code-length-code lengths (order 17 18 0 1 2 3 4 5 16 6 ..15), max_symbol bit 0
  or_bits(words, pos + 1, 15u, 4);
  pos += 5;
  for (int i = 0; i < 19; ++i) {
    const bool unused = i == 0 || i == 1 || i == 8;          // 17, 18, 16
    or_bits(words, pos, unused ? 0u : 4u, 3);
    pos += 3;
  }
  pos += 1;
  for (int s = 0; s < alphabet; ++s) {
    or_bits(words, pos, rev_bits(s < 256 ? T.len[s] : 0u, 4), 4);
    pos += 4;
  }
}

// grid (B): header bits, chunk bit offsets (exclusive scan, in place), RIFF framing, file length
__global__ void __launch_bounds__(kThreads) webp_header_kernel(const CodeTab* __restrict__ tabs, unsigned long long* __restrict__ chunk_bits,
                                                               int n_chunks, int H, int W, int C, uint8_t* __restrict__ out,
                                                               size_t out_stride, int64_t* __restrict__ out_len) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  uint8_t* file = out + static_cast<size_t>(b) * out_stride;
  uint32_t* words = reinterpret_cast<uint32_t*>(file + kRiffBytes);
  uint64_t pos = 0;
  or_bits(words, pos, 0x2Fu, 8);
  or_bits(words, pos + 8, static_cast<uint32_t>(W - 1), 14);
  or_bits(words, pos + 22, static_cast<uint32_t>(H - 1), 14);
  or_bits(words, pos + 36, C == 4 ? 1u : 0u, 1);
  pos += 40;                                            // (+ 3 version bits = 0)
  or_bits(words, pos, 1u, 1);                           // a transform follows: predictor (type 0)
  or_bits(words, pos + 3, static_cast<uint32_t>(kSizeBits - 2), 3);
  pos += 6;
  pos += 1;                                             // mode image: no colour cache
  put_simple(words, pos, kPredMode);
  put_simple(words, pos, 0);
  put_simple(words, pos, 0);
  put_simple(words, pos, 255);
  put_simple(words, pos, 0);
  pos += 3;                                             // no more transforms, no colour cache, no meta prefix codes
  const CodeTab* T = tabs + b * 4;
  put_code(words, pos, T[0], 280);
  put_code(words, pos, T[1], 256);
  put_code(words, pos, T[2], 256);
  put_code(words, pos, T[3], 256);
  put_simple(words, pos, 0);                            // distance
  unsigned long long* cb = chunk_bits + static_cast<int64_t>(b) * n_chunks;
  uint64_t run = pos;
  for (int k = 0; k < n_chunks; ++k) {
    const uint64_t t = cb[k];
    cb[k] = run;
    run += t;
  }
  const uint32_t data_bytes = static_cast<uint32_t>((run + 7) >> 3);
  const uint32_t padded = data_bytes + (data_bytes & 1u);
  const uint32_t riff = 4 + 8 + padded;
  const uint8_t hdr[kRiffBytes] = {'R', 'I', 'F', 'F', static_cast<uint8_t>(riff), static_cast<uint8_t>(riff >> 8),
                                   static_cast<uint8_t>(riff >> 16), static_cast<uint8_t>(riff >> 24), 'W', 'E', 'B', 'P', 'V', 'P',
                                   '8', 'L', static_cast<uint8_t>(data_bytes), static_cast<uint8_t>(data_bytes >> 8),
                                   static_cast<uint8_t>(data_bytes >> 16), static_cast<uint8_t>(data_bytes >> 24)};
  for (int i = 0; i < kRiffBytes; ++i) file[i] = hdr[i];
  out_len[b] = static_cast<int64_t>(kRiffBytes) + padded;
}

// grid (chunks, B): the pixel codes at their global bit offsets
__global__ void __launch_bounds__(kThreads) webp_emit_kernel(const uint32_t* __restrict__ res, int64_t n, const CodeTab* __restrict__ tabs,
                                                             const unsigned long long* __restrict__ chunk_off, uint8_t* __restrict__ out,
                                                             size_t out_stride) {
  __shared__ uint16_t code[4][256];
  __shared__ uint8_t lens[4][256];
  __shared__ uint32_t warp_tot[kThreads / 32];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 4 * 256; i += kThreads) {
    code[i >> 8][i & 255] = tabs[b * 4 + (i >> 8)].code[i & 255];
    lens[i >> 8][i & 255] = tabs[b * 4 + (i >> 8)].len[i & 255];
  }
  __syncthreads();
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * kChunkPixels;
  const int64_t p1 = min(n, p0 + kChunkPixels);
  const int per = kChunkPixels / kThreads;                           // consecutive pixels per thread
  const int64_t s0 = min(p1, p0 + static_cast<int64_t>(tid) * per), s1 = min(p1, s0 + per);
  const uint32_t* rr = res + b * n;
  uint32_t my = 0;
  for (int64_t p = s0; p < s1; ++p) {
    const uint32_t r = rr[p];
    my += lens[0][(r >> 8) & 255u] + lens[1][r & 255u] + lens[2][(r >> 16) & 255u] + lens[3][r >> 24];
  }
  uint32_t v = my;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  uint32_t base = 0;
  for (int w = 0; w < warp; ++w) base += warp_tot[w];
  uint64_t pos = chunk_off[static_cast<int64_t>(b) * gridDim.x + blockIdx.x] + base + v - my;
  uint32_t* words = reinterpret_cast<uint32_t*>(out + static_cast<size_t>(b) * out_stride + kRiffBytes);
  uint64_t acc = 0;
  int nb = static_cast<int>(pos & 31u);
  uint64_t wi = pos >> 5;
  bool first = true;
  auto put = [&](uint32_t c, int l) {
    acc |= static_cast<uint64_t>(c) << nb;
    nb += l;
    if (nb >= 32) {
      if (first) {
        atomicOr(&words[wi], static_cast<uint32_t>(acc));
        first = false;
      } else {
        words[wi] = static_cast<uint32_t>(acc);                      // a word this thread's bits cover entirely
      }
      acc >>= 32;
      nb -= 32;
      ++wi;
    }
  };
  for (int64_t p = s0; p < s1; ++p) {
    const uint32_t r = rr[p];
    const uint32_t g = (r >> 8) & 255u, rd = r & 255u, bl = (r >> 16) & 255u, al = r >> 24;
    put(code[0][g], lens[0][g]);
    put(code[1][rd], lens[1][rd]);
    put(code[2][bl], lens[2][bl]);
    put(code[3][al], lens[3][al]);
  }
  if (nb > 0 && my > 0) atomicOr(&words[wi], static_cast<uint32_t>(acc));
}

struct Plan {
  int64_t n;
  int n_chunks;
  size_t out_stride, res_bytes, hist_bytes, tab_bytes, chunk_bytes;
};

int make_plan(int B, int H, int W, int C, Plan* p) {
  GYRE_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && H <= 16384 && W <= 16384 && (C == 3 || C == 4),
               "webp_encode: bad shape %d x %d x %d x %d (RGB or RGBA, at most 16384 a side)", B, H, W, C);
  p->n = static_cast<int64_t>(H) * W;
  p->n_chunks = static_cast<int>((p->n + kChunkPixels - 1) / kChunkPixels);
  GYRE_REQUIRE(p->n_chunks <= 65535 * 16, "webp_encode: image too large");
  // header (< 1 KB) + the worst a Huffman code does on bytes: entropy (<= 8 bits) + 1 per symbol, with room for the length
  // limit's repairs: 10 bits per channel symbol
  const size_t worst = kRiffBytes + 1024 + static_cast<size_t>(p->n) * 4 * 10 / 8 + 16;
  p->out_stride = (worst + 63) & ~size_t(63);
  auto up256 = [](size_t v) { return (v + 255) & ~size_t(255); };          // every section starts 256-byte aligned
  p->res_bytes = up256(static_cast<size_t>(B) * p->n * sizeof(uint32_t));
  p->hist_bytes = up256(static_cast<size_t>(B) * 4 * 256 * sizeof(int));
  p->tab_bytes = up256(static_cast<size_t>(B) * 4 * sizeof(CodeTab));
  p->chunk_bytes = up256(static_cast<size_t>(B) * p->n_chunks * sizeof(unsigned long long));
  return 0;
}

}  // namespace

int webp_sizes(int B, int H, int W, int C, size_t* workspace_bytes, size_t* out_stride) {
  Plan p;
  GYRE_TRY(make_plan(B, H, W, C, &p));
  if (workspace_bytes) *workspace_bytes = p.res_bytes + p.hist_bytes + p.tab_bytes + p.chunk_bytes + 1024;
  if (out_stride) *out_stride = p.out_stride;
  return 0;
}

int webp_encode(const uint8_t* images, int B, int H, int W, int C, uint8_t* out, size_t out_stride, int64_t* out_len,
                void* workspace, size_t workspace_bytes, cudaStream_t st) {
  GYRE_REQUIRE(images && out && out_len && workspace, "webp_encode: null argument");
  Plan p;
  GYRE_TRY(make_plan(B, H, W, C, &p));
  size_t need = 0;
  GYRE_TRY(webp_sizes(B, H, W, C, &need, nullptr));
  GYRE_REQUIRE(workspace_bytes >= need, "webp_encode: workspace %zu < %zu bytes", workspace_bytes, need);
  GYRE_REQUIRE(out_stride >= p.out_stride && out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
               "webp_encode: output stride %zu < %zu bytes (or not 4-byte aligned)", out_stride, p.out_stride);
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  uint32_t* res = reinterpret_cast<uint32_t*>(base);
  int* hist = reinterpret_cast<int*>(base + p.res_bytes);
  CodeTab* tabs = reinterpret_cast<CodeTab*>(base + p.res_bytes + p.hist_bytes);
  unsigned long long* chunk_bits = reinterpret_cast<unsigned long long*>(base + p.res_bytes + p.hist_bytes + p.tab_bytes);
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  GYRE_CHECK_CUDA(cudaMemsetAsync(hist, 0, static_cast<size_t>(B) * 4 * 256 * sizeof(int), st));
  GYRE_CHECK_CUDA(cudaMemsetAsync(out, 0, static_cast<size_t>(B) * out_stride, st));
  const dim3 grid(p.n_chunks, B);
  webp_residual_kernel<<<grid, kThreads, 0, st>>>(images, H, W, C, res, hist);
  GYRE_CHECK_CUDA(cudaGetLastError());
  webp_codes_kernel<<<dim3(4, B), kThreads, 0, st>>>(hist, tabs);
  GYRE_CHECK_CUDA(cudaGetLastError());
  webp_count_kernel<<<grid, kThreads, 0, st>>>(res, p.n, tabs, chunk_bits);
  GYRE_CHECK_CUDA(cudaGetLastError());
  webp_header_kernel<<<B, kThreads, 0, st>>>(tabs, chunk_bits, p.n_chunks, H, W, C, out, out_stride, out_len);
  GYRE_CHECK_CUDA(cudaGetLastError());
  webp_emit_kernel<<<grid, kThreads, 0, st>>>(res, p.n, tabs, chunk_bits, out, out_stride);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gyre
