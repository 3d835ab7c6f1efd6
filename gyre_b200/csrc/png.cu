// PNG encoding on the device (the step after the decode / safety check).
//
// The reference hands every finished image to `torchvision.io.encode_png` on the host (gyre/images.py:93-111 toPngBytes,
// called per artifact by gyre/services/generate.py:79): `.cpu()` of the fp32 image, libpng at compression level 6, one
// image at a time - tens of milliseconds each, serial with the next request at tens of images per second per node.  PNG is
// lossless, so the contract is the decoded image, not libpng's byte stream.
//
// Layout chosen so that nothing in the bit stream depends on another chunk (one CTA per chunk, all chunks of all images in
// flight at once):
//   signature | IHDR | one IDAT per chunk of scanlines | a 4-byte IDAT with the Adler-32 | IEND
//   chunk: adaptive filter per scanline (minimum sum of absolute residuals, ties to the lower type) -> literal-only
//   deflate, ONE dynamic-Huffman block built from the chunk's own histogram (257 code lengths sent with a flat 4-bit
//   code-length code, two 1-bit distance codes so every inflater sees complete codes) or one stored block when that is not
//   smaller, then an empty stored block (BFINAL on the last chunk) that byte-aligns the stream - what zlib's Z_SYNC_FLUSH
//   emits, the trick parallel gzip implementations use.  Chunk 0 starts with the zlib header 78 01.
//   CRC-32 of every IDAT: per-thread table CRCs of 1/256th of the payload combined with x^(8n) mod P (GF(2) polynomial
//   products); Adler-32: per-chunk (sum, weighted sum) pairs combined in order by the assembling kernel.
// Against libpng level 6 on photographic content the files are ~1 % larger (no LZ77 matches; the filters + entropy code
// carry almost all of PNG's gain there); flat synthetic images compress worse (1 bit per byte at best).
// oracle/png.py restates this encoder on the CPU byte for byte and is itself pinned by zlib, Pillow and torchvision decoding
// its output (tests/test_png_cpu.py); tests/test_png_gpu.py compares the two byte streams and decodes both.
#include "common.cuh"
#include "ops.h"

namespace gyre {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunkTarget = 16384;      // filtered bytes per chunk (at least one scanline)
constexpr int kMaxRowBytes = 32768;
constexpr int kMaxBits = 15;
constexpr int kHeaderBits = 3 + 5 + 5 + 4 + 19 * 3 + 259 * 4;
constexpr uint32_t kCrcPoly = 0xEDB88320u;
constexpr uint32_t kCrcIdatTag = 0x35AF061Eu;   // crc32("IDAT")
constexpr uint32_t kAdlerMod = 65521u;

struct ChunkMeta {
  uint32_t size;      // payload bytes of this IDAT
  uint32_t crc;       // CRC-32 over "IDAT" + payload
  uint32_t s1, s2;    // Adler pieces over the chunk's filtered bytes: sum d_i, sum (n - i) d_i   (mod 65521)
  uint32_t raw_len;   // filtered bytes
  uint32_t pad[3];
};

__device__ __forceinline__ uint32_t gf_mul(uint32_t a, uint32_t b) {
  // product of two polynomials over GF(2) modulo the CRC-32 polynomial, reflected bit order (bit 31 = x^0)
  uint32_t p = 0;
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    if (a & 0x80000000u) p ^= b;
    a <<= 1;
    b = (b >> 1) ^ ((b & 1u) ? kCrcPoly : 0u);
  }
  return p;
}

__device__ __forceinline__ uint32_t x_pow_8n(uint32_t n, const uint32_t* pow8) {   // x^(8 n) mod P
  uint32_t r = 0x80000000u;
  for (int k = 0; n; ++k, n >>= 1)
    if (n & 1u) r = gf_mul(r, pow8[k]);
  return r;
}

__device__ __forceinline__ uint32_t crc_bitwise(uint32_t crc, uint8_t byte) {
  crc ^= byte;
#pragma unroll
  for (int i = 0; i < 8; ++i) crc = (crc >> 1) ^ ((crc & 1u) ? kCrcPoly : 0u);
  return crc;
}

__device__ __forceinline__ int paeth(int a, int b, int c) {
  const int p = a + b - c;
  const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

__device__ __forceinline__ int residual(int f, int x, int a, int b, int c) {
  switch (f) {
    case 0: return x;
    case 1: return x - a;
    case 2: return x - b;
    case 3: return x - ((a + b) >> 1);
    default: return x - paeth(a, b, c);
  }
}

__device__ __forceinline__ void put_be32(uint8_t* p, uint32_t v) {
  p[0] = static_cast<uint8_t>(v >> 24);
  p[1] = static_cast<uint8_t>(v >> 16);
  p[2] = static_cast<uint8_t>(v >> 8);
  p[3] = static_cast<uint8_t>(v);
}

// OR `nbits` (<= 32) bits of `value` into the zero-initialised stream at bit position `pos` (shared-memory words)
__device__ __forceinline__ void or_bits(uint32_t* words, uint32_t pos, uint32_t value, int nbits) {
  const uint64_t v = static_cast<uint64_t>(value) << (pos & 31u);
  atomicOr(&words[pos >> 5], static_cast<uint32_t>(v));
  if ((pos & 31u) + nbits > 32) atomicOr(&words[(pos >> 5) + 1], static_cast<uint32_t>(v >> 32));
}

__device__ __forceinline__ uint32_t rev_bits(uint32_t v, int n) { return __brev(v) >> (32 - n); }

struct Smem {
  uint32_t hist[kWarps][260];
  uint32_t keys[512];
  uint32_t leaf_w[260];
  uint32_t int_w[260];
  uint16_t par_leaf[260];
  uint16_t par_int[260];
  uint16_t depth_int[260];
  uint16_t codes[260];
  uint8_t lens[260];
  uint32_t crc_tab[256];
  uint32_t pow8[32];
  uint32_t scan[kThreads];
  uint32_t red[kThreads];
  uint32_t a_s1[kThreads];
  uint32_t a_s2[kThreads];
  int cnt[64];
  int n_syms;
  int mode;             // 1 dynamic block, 0 stored
  uint32_t dyn_bits;    // header + literals + end-of-block
};

// grid (n_chunks, batch).  img u8 [B, H, W * C]; staging [B][n_chunks][stride] bytes; meta [B][n_chunks]
__global__ void __launch_bounds__(kThreads) png_chunk_kernel(const uint8_t* __restrict__ img, int H, int wc, int bpp, int R,
                                                             int chunk_cap, uint8_t* __restrict__ staging, int stride,
                                                             ChunkMeta* __restrict__ meta) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  Smem& S = *reinterpret_cast<Smem*>(dyn_smem);
  uint8_t* filt = dyn_smem + ((sizeof(Smem) + 15) & ~size_t(15));
  uint32_t* outw = reinterpret_cast<uint32_t*>(filt + chunk_cap);
  const int out_words = (chunk_cap + 32) / 4;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k = blockIdx.x, b = blockIdx.y, n_chunks = gridDim.x;
  const int r0 = k * R, r1 = min(H, r0 + R);
  const int row_bytes = 1 + wc;
  const int len = (r1 - r0) * row_bytes;
  const uint8_t* src = img + static_cast<size_t>(b) * H * wc;

  for (int i = tid; i < kWarps * 260; i += kThreads) (&S.hist[0][0])[i] = 0;
  for (int i = tid; i < out_words; i += kThreads) outw[i] = 0;
  {
    uint32_t c = tid;
#pragma unroll
    for (int i = 0; i < 8; ++i) c = (c >> 1) ^ ((c & 1u) ? kCrcPoly : 0u);
    S.crc_tab[tid] = c;
  }
  if (tid == 0) {
    uint32_t p = 0x00800000u;      // x^8
    for (int i = 0; i < 32; ++i) {
      S.pow8[i] = p;
      p = gf_mul(p, p);
    }
  }
  __syncthreads();

  // ---- 1. filter: one warp per scanline, first the five costs, then the chosen residuals + histogram
  for (int r = r0 + warp; r < r1; r += kWarps) {
    const uint8_t* cur = src + static_cast<size_t>(r) * wc;
    const uint8_t* up = r > 0 ? cur - wc : nullptr;
    int cost[5] = {0, 0, 0, 0, 0};
    for (int i = lane; i < wc; i += 32) {
      const int x = cur[i];
      const int a = i >= bpp ? cur[i - bpp] : 0;
      const int bb = up ? up[i] : 0;
      const int c = (up && i >= bpp) ? up[i - bpp] : 0;
#pragma unroll
      for (int f = 0; f < 5; ++f) {
        const int v = residual(f, x, a, bb, c) & 255;
        cost[f] += v < 128 ? v : 256 - v;
      }
    }
#pragma unroll
    for (int f = 0; f < 5; ++f)
      for (int o = 16; o > 0; o >>= 1) cost[f] += __shfl_xor_sync(0xffffffffu, cost[f], o);
    int best = 0;
#pragma unroll
    for (int f = 1; f < 5; ++f)
      if (cost[f] < cost[best]) best = f;
    uint8_t* dst = filt + (r - r0) * row_bytes;
    if (lane == 0) {
      dst[0] = static_cast<uint8_t>(best);
      atomicAdd(&S.hist[warp][best], 1u);
    }
    for (int i = lane; i < wc; i += 32) {
      const int x = cur[i];
      const int a = i >= bpp ? cur[i - bpp] : 0;
      const int bb = up ? up[i] : 0;
      const int c = (up && i >= bpp) ? up[i - bpp] : 0;
      const int v = residual(best, x, a, bb, c) & 255;
      dst[1 + i] = static_cast<uint8_t>(v);
      atomicAdd(&S.hist[warp][v], 1u);
    }
  }
  __syncthreads();

  // ---- 2. symbol frequencies -> keys (freq, symbol), bitonic sort ascending
  for (int i = tid; i < 512; i += kThreads) {
    uint32_t f = 0;
    if (i < 256) {
#pragma unroll
      for (int w = 0; w < kWarps; ++w) f += S.hist[w][i];
    } else if (i == 256) {
      f = 1;          // end of block
    }
    S.keys[i] = f ? ((f << 9) | static_cast<uint32_t>(i)) : 0xFFFFFFFFu;
    S.red[i & (kThreads - 1)] = 0;
  }
  __syncthreads();
  for (int size = 2; size <= 512; size <<= 1) {
    for (int strd = size >> 1; strd > 0; strd >>= 1) {
      const int i = 2 * tid - (tid & (strd - 1));          // lower index of the pair this thread owns
      const int j = i + strd;
      const bool up_dir = (i & size) == 0;
      const uint32_t a = S.keys[i], c = S.keys[j];
      if ((a > c) == up_dir) {
        S.keys[i] = c;
        S.keys[j] = a;
      }
      __syncthreads();
    }
  }
  atomicAdd(&S.red[0], static_cast<uint32_t>((S.keys[tid] != 0xFFFFFFFFu) + (S.keys[tid + 256] != 0xFFFFFFFFu)));
  __syncthreads();

  // ---- 3. code lengths (one thread: a few thousand shared-memory operations), canonical codes
  if (tid == 0) {
    const int n = static_cast<int>(S.red[0]);
    S.n_syms = n;
    for (int i = 0; i < n; ++i) S.leaf_w[i] = S.keys[i] >> 9;
    for (int i = 0; i < 260; ++i) S.lens[i] = 0;
    for (int i = 0; i < 64; ++i) S.cnt[i] = 0;
    // two-queue Huffman: leaves ascending, internal nodes appear in ascending weight; on equal weights the leaf goes first
    int li = 0, ii = 0, made = 0;
    for (int kk = 0; kk < n - 1; ++kk) {
      uint32_t w = 0;
      for (int t = 0; t < 2; ++t) {
        if (li < n && (ii >= made || S.leaf_w[li] <= S.int_w[ii])) {
          w += S.leaf_w[li];
          S.par_leaf[li++] = static_cast<uint16_t>(kk);
        } else {
          w += S.int_w[ii];
          S.par_int[ii++] = static_cast<uint16_t>(kk);
        }
      }
      S.int_w[made++] = w;
    }
    if (n >= 2) S.depth_int[n - 2] = 0;
    for (int j = n - 3; j >= 0; --j) S.depth_int[j] = static_cast<uint16_t>(S.depth_int[S.par_int[j]] + 1);
    for (int i = 0; i < n; ++i) {
      const int d = S.depth_int[S.par_leaf[i]] + 1;
      S.cnt[d < kMaxBits ? d : kMaxBits] += 1;       // depths beyond the limit are folded into the longest length ...
    }
    int total = 0;
    for (int l = 1; l <= kMaxBits; ++l) total += S.cnt[l] << (kMaxBits - l);
    while (total > (1 << kMaxBits)) {                // ... and the Kraft sum repaired one unit at a time
      S.cnt[kMaxBits] -= 1;
      for (int l = kMaxBits - 1; l > 0; --l)
        if (S.cnt[l]) {
          S.cnt[l] -= 1;
          S.cnt[l + 1] += 2;
          break;
        }
      total -= 1;
    }
    int idx = 0;
    for (int l = kMaxBits; l > 0; --l)
      for (int c = 0; c < S.cnt[l]; ++c) S.lens[S.keys[idx++] & 511u] = static_cast<uint8_t>(l);
    uint32_t next[kMaxBits + 2];
    uint32_t code = 0;
    next[0] = 0;
    for (int l = 1; l <= kMaxBits; ++l) {
      code = (code + (l > 1 ? static_cast<uint32_t>(S.cnt[l - 1]) : 0u)) << 1;
      next[l] = code;
    }
    uint32_t bits = kHeaderBits;
    for (int s = 0; s < 257; ++s) {
      const int l = S.lens[s];
      if (l) {
        S.codes[s] = static_cast<uint16_t>(rev_bits(next[l]++, l));
        bits += static_cast<uint32_t>(l) * (s < 256 ? 0u : 1u);     // end of block; the literals are added from the scan
      } else {
        S.codes[s] = 0;
      }
    }
    S.dyn_bits = bits;
  }
  __syncthreads();

  // ---- 4. bit lengths per thread segment, scan, stored / dynamic decision
  const int seg = (len + kThreads - 1) / kThreads;
  const int s0 = min(len, tid * seg), s1 = min(len, s0 + seg);
  uint32_t my_bits = 0;
  for (int i = s0; i < s1; ++i) my_bits += S.lens[filt[i]];
  {
    uint32_t v = my_bits;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) S.red[warp] = v;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += S.red[w];
    S.scan[tid] = base + v - my_bits;                 // exclusive
    __syncthreads();
  }
  uint32_t lit_bits = 0;
  for (int w = 0; w < kWarps; ++w) lit_bits += S.red[w];
  const uint32_t dyn_bits = S.dyn_bits + lit_bits;
  const uint32_t dyn_bytes = (dyn_bits + 3 + 7) / 8 + 4;
  const uint32_t stored_bytes = 5u + static_cast<uint32_t>(len) + 5u;
  const bool dynamic = dyn_bytes < stored_bytes;
  const uint32_t lead = k == 0 ? 2u : 0u;             // zlib header in front of chunk 0
  const bool last = k == n_chunks - 1;
  uint8_t* outb = reinterpret_cast<uint8_t*>(outw);
  uint32_t payload;
  if (dynamic) {
    const uint32_t p0 = lead * 8;
    if (tid == 0) {
      if (lead) or_bits(outw, 0, 0x0178u, 16);
      // BFINAL 0, BTYPE 10, HLIT 0 (257 codes), HDIST 1 (2 codes), HCLEN 15 (19 lengths)
      or_bits(outw, p0, 0u | (2u << 1) | (0u << 3) | (1u << 8) | (15u << 13), 17);
      // code-length-code lengths in the order 16 17 18 0 8 7 ...: 0 for the three run-length symbols, 4 for 0..15
      for (int i = 0; i < 19; ++i) or_bits(outw, p0 + 17 + 3 * i, i < 3 ? 0u : 4u, 3);
      const uint32_t end = p0 + S.dyn_bits + lit_bits - S.lens[256];
      or_bits(outw, end, S.codes[256], S.lens[256]);
      uint32_t q = end + S.lens[256];
      or_bits(outw, q, last ? 1u : 0u, 3);
      q = (q + 3 + 7) & ~7u;
      or_bits(outw, q, 0xFFFF0000u, 32);
    }
    for (int i = tid; i < 259; i += kThreads) {
      const uint32_t l = i < 257 ? S.lens[i] : 1u;
      or_bits(outw, p0 + 74 + 4 * i, rev_bits(l, 4), 4);
    }
    // literals
    uint32_t pos = p0 + kHeaderBits + S.scan[tid];
    uint64_t acc = 0;
    int nb = static_cast<int>(pos & 31u);
    uint32_t wi = pos >> 5;
    bool first = true;
    for (int i = s0; i < s1; ++i) {
      const int sym = filt[i];
      acc |= static_cast<uint64_t>(S.codes[sym]) << nb;
      nb += S.lens[sym];
      if (nb >= 32) {
        if (first) {
          atomicOr(&outw[wi], static_cast<uint32_t>(acc));
          first = false;
        } else {
          outw[wi] = static_cast<uint32_t>(acc);
        }
        acc >>= 32;
        nb -= 32;
        ++wi;
      }
    }
    if (nb > 0 && s1 > s0) atomicOr(&outw[wi], static_cast<uint32_t>(acc));
    payload = lead + dyn_bytes;
  } else {
    if (tid == 0) {
      if (lead) {
        outb[0] = 0x78;
        outb[1] = 0x01;
      }
      uint8_t* p = outb + lead;
      p[0] = 0;
      p[1] = static_cast<uint8_t>(len & 255);
      p[2] = static_cast<uint8_t>(len >> 8);
      p[3] = static_cast<uint8_t>(~len & 255);
      p[4] = static_cast<uint8_t>((~len >> 8) & 255);
      uint8_t* t = p + 5 + len;
      t[0] = last ? 1 : 0;
      t[1] = 0;
      t[2] = 0;
      t[3] = 0xFF;
      t[4] = 0xFF;
    }
    for (int i = tid; i < len; i += kThreads) outb[lead + 5 + i] = filt[i];
    payload = lead + stored_bytes;
  }
  __syncthreads();

  // ---- 5. CRC-32 over "IDAT" + payload; Adler pieces over the filtered bytes
  {
    const uint32_t pseg = (payload + kThreads - 1) / kThreads;
    const uint32_t q0 = min(payload, tid * pseg), q1 = min(payload, q0 + pseg);
    uint32_t crc = 0;
    if (q1 > q0) {
      crc = 0xFFFFFFFFu;
      for (uint32_t i = q0; i < q1; ++i) crc = S.crc_tab[(crc ^ outb[i]) & 255u] ^ (crc >> 8);
      crc = ~crc;
      crc = gf_mul(x_pow_8n(payload - q1, S.pow8), crc);
    }
    if (tid == 0) crc ^= gf_mul(x_pow_8n(payload, S.pow8), kCrcIdatTag);
    for (int o = 16; o > 0; o >>= 1) crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
    uint32_t t1 = 0, t2 = 0;
    const uint32_t m = static_cast<uint32_t>(s1 - s0);
    for (int i = s0; i < s1; ++i) {
      const uint32_t d = filt[i];
      t1 += d;
      t2 += (m - static_cast<uint32_t>(i - s0)) * d;
    }
    S.a_s1[tid] = t1;
    S.a_s2[tid] = t2 % kAdlerMod;
    __syncthreads();           // (everyone is done with S.red from the scan)
    if (lane == 0) S.red[warp] = crc;
    __syncthreads();
    if (tid == 0) {
      uint32_t c = 0;
      for (int w = 0; w < kWarps; ++w) c ^= S.red[w];
      uint64_t A = 0, B = 0;
      for (int t = 0; t < kThreads; ++t) {
        const uint32_t mt = static_cast<uint32_t>(min(len, (t + 1) * seg) - min(len, t * seg));
        B = (B + static_cast<uint64_t>(mt) * A + S.a_s2[t]) % kAdlerMod;
        A = (A + S.a_s1[t]) % kAdlerMod;
      }
      ChunkMeta mm;
      mm.size = payload;
      mm.crc = c;
      mm.s1 = static_cast<uint32_t>(A);
      mm.s2 = static_cast<uint32_t>(B);
      mm.raw_len = static_cast<uint32_t>(len);
      mm.pad[0] = mm.pad[1] = mm.pad[2] = 0;
      meta[static_cast<size_t>(b) * n_chunks + k] = mm;
    }
  }
  // ---- 6. payload -> staging (word copies: both sides are 16-byte aligned)
  uint32_t* dst = reinterpret_cast<uint32_t*>(staging + (static_cast<size_t>(b) * n_chunks + k) * stride);
  for (uint32_t i = tid; i < (payload + 3) / 4; i += kThreads) dst[i] = outw[i];
}

// grid (n_chunks, batch): frames chunk k as an IDAT at its final offset; chunk 0 also writes the signature and IHDR, the last
// chunk the Adler IDAT, IEND and the file length
__global__ void __launch_bounds__(kThreads) png_assemble_kernel(const uint8_t* __restrict__ staging, int stride,
                                                                const ChunkMeta* __restrict__ meta, int H, int W, int color_type,
                                                                uint8_t* __restrict__ out, size_t out_stride,
                                                                int64_t* __restrict__ out_len) {
  __shared__ unsigned long long part[kThreads];
  const int tid = threadIdx.x;
  const int k = blockIdx.x, b = blockIdx.y, n_chunks = gridDim.x;
  const ChunkMeta* mb = meta + static_cast<size_t>(b) * n_chunks;
  unsigned long long s = 0;
  for (int j = tid; j < k; j += kThreads) s += 12ull + mb[j].size;
  part[tid] = s;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (tid < o) part[tid] += part[tid + o];
    __syncthreads();
  }
  const size_t off = 33 + static_cast<size_t>(part[0]);
  uint8_t* dst = out + static_cast<size_t>(b) * out_stride;
  const ChunkMeta mk = mb[k];
  const uint8_t* srcp = staging + (static_cast<size_t>(b) * n_chunks + k) * stride;
  for (uint32_t i = tid; i < mk.size; i += kThreads) dst[off + 8 + i] = srcp[i];
  if (tid == 0) {
    put_be32(dst + off, mk.size);
    dst[off + 4] = 'I';
    dst[off + 5] = 'D';
    dst[off + 6] = 'A';
    dst[off + 7] = 'T';
    put_be32(dst + off + 8 + mk.size, mk.crc);
  }
  if (k == 0 && tid == 32) {
    const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    for (int i = 0; i < 8; ++i) dst[i] = sig[i];
    uint8_t* p = dst + 8;
    put_be32(p, 13);
    p[4] = 'I';
    p[5] = 'H';
    p[6] = 'D';
    p[7] = 'R';
    put_be32(p + 8, static_cast<uint32_t>(W));
    put_be32(p + 12, static_cast<uint32_t>(H));
    p[16] = 8;
    p[17] = static_cast<uint8_t>(color_type);
    p[18] = 0;
    p[19] = 0;
    p[20] = 0;
    uint32_t crc = 0xFFFFFFFFu;
    for (int i = 4; i < 21; ++i) crc = crc_bitwise(crc, p[i]);
    put_be32(p + 21, ~crc);
  }
  if (k == n_chunks - 1 && tid == 64) {
    uint64_t A = 1, B = 0;
    for (int j = 0; j < n_chunks; ++j) {
      B = (B + static_cast<uint64_t>(mb[j].raw_len % kAdlerMod) * A + mb[j].s2) % kAdlerMod;
      A = (A + mb[j].s1) % kAdlerMod;
    }
    const uint32_t adler = static_cast<uint32_t>((B << 16) | A);
    uint8_t* p = dst + off + 12 + mk.size;
    put_be32(p, 4);
    p[4] = 'I';
    p[5] = 'D';
    p[6] = 'A';
    p[7] = 'T';
    put_be32(p + 8, adler);
    uint32_t crc = 0xFFFFFFFFu;
    for (int i = 4; i < 12; ++i) crc = crc_bitwise(crc, p[i]);
    put_be32(p + 12, ~crc);
    p += 16;
    put_be32(p, 0);
    p[4] = 'I';
    p[5] = 'E';
    p[6] = 'N';
    p[7] = 'D';
    put_be32(p + 8, 0xAE426082u);
    out_len[b] = static_cast<int64_t>(off + 12 + mk.size + 16 + 12);
  }
}

struct Plan {
  int row_bytes, R, n_chunks, chunk_cap, stride;
  size_t smem;
};

int make_plan(int H, int W, int C, Plan* p) {
  GYRE_REQUIRE(H > 0 && W > 0 && C >= 1 && C <= 4, "png_encode: bad image shape %d x %d x %d", H, W, C);
  const int64_t rb = 1 + static_cast<int64_t>(W) * C;
  GYRE_REQUIRE(rb <= kMaxRowBytes, "png_encode: scanline of %lld bytes (limit %d)", static_cast<long long>(rb), kMaxRowBytes);
  p->row_bytes = static_cast<int>(rb);
  p->R = std::max(1, std::min(H, kChunkTarget / p->row_bytes));
  p->n_chunks = (H + p->R - 1) / p->R;
  p->chunk_cap = (p->R * p->row_bytes + 15) & ~15;
  p->stride = p->chunk_cap + 32;
  p->smem = ((sizeof(Smem) + 15) & ~size_t(15)) + p->chunk_cap + p->stride;
  return 0;
}

}  // namespace

int png_sizes(int B, int H, int W, int C, size_t* workspace_bytes, size_t* out_stride) {
  Plan p;
  GYRE_TRY(make_plan(H, W, C, &p));
  GYRE_REQUIRE(B > 0, "png_encode: empty batch");
  if (workspace_bytes)
    *workspace_bytes = static_cast<size_t>(B) * p.n_chunks * (static_cast<size_t>(p.stride) + sizeof(ChunkMeta)) + 256;
  // signature + IHDR, per chunk: framing + zlib header + stored-block worst case, Adler IDAT, IEND
  if (out_stride)
    *out_stride = (33 + static_cast<size_t>(p.n_chunks) * (12 + 2 + 10) + static_cast<size_t>(H) * p.row_bytes + 16 + 12 + 63) &
                  ~size_t(63);
  return 0;
}

int png_encode(const uint8_t* images, int B, int H, int W, int C, uint8_t* out, size_t out_stride, int64_t* out_len,
               void* workspace, size_t workspace_bytes, cudaStream_t st) {
  GYRE_REQUIRE(images && out && out_len && workspace, "png_encode: null argument");
  Plan p;
  GYRE_TRY(make_plan(H, W, C, &p));
  size_t need = 0, min_stride = 0;
  GYRE_TRY(png_sizes(B, H, W, C, &need, &min_stride));
  GYRE_REQUIRE(workspace_bytes >= need, "png_encode: workspace %zu < %zu bytes", workspace_bytes, need);
  GYRE_REQUIRE(out_stride >= min_stride, "png_encode: output stride %zu < %zu bytes", out_stride, min_stride);
  GYRE_REQUIRE(B <= 65535, "png_encode: batch %d too large for one launch", B);
  static const int color_type[5] = {0, 0, 4, 2, 6};
  uintptr_t base = (reinterpret_cast<uintptr_t>(workspace) + 15) & ~uintptr_t(15);
  ChunkMeta* meta = reinterpret_cast<ChunkMeta*>(base);
  uint8_t* staging = reinterpret_cast<uint8_t*>(base + static_cast<size_t>(B) * p.n_chunks * sizeof(ChunkMeta));
  GYRE_CHECK_CUDA(cudaFuncSetAttribute(png_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)));
  prof::Scope ps(prof::F_ELEMENTWISE, 0.0, 0.0, st);
  const dim3 grid(p.n_chunks, B);
  png_chunk_kernel<<<grid, kThreads, p.smem, st>>>(images, H, W * C, C, p.R, p.chunk_cap, staging, p.stride, meta);
  GYRE_CHECK_CUDA(cudaGetLastError());
  png_assemble_kernel<<<grid, kThreads, 0, st>>>(staging, p.stride, meta, H, W, color_type[C], out, out_stride, out_len);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gyre
