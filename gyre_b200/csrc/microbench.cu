// Measurement-only kernels (not on any product path): tcgen05.mma issue/latency characteristics on this part.
#include "common.cuh"
#include "ops.h"

namespace gyre {

// One CTA: a single thread issues `reps` kind::f16 MMAs (M=128, K=16, N=n) round-robin over `naccs` accumulators
// (TMEM columns acc * 128), operands either both in shared memory or A in TMEM; returns clocks from first issue
// to completion (commit -> mbarrier), and the issue time alone.  Operand contents are irrelevant.
template <int NACC, int ATMEM>
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int n, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sA = smem;                  // 128 x 64 halfs, 128B-swizzled K-major
  uint8_t* sB = smem + 16384;          // 256 x 64 halfs
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, n);
    const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA));
    const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sB));
    const uint32_t astride = n <= 128 ? 128 : 256;
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 12) {
#pragma unroll
      for (int u = 0; u < 12; ++u) {
        const uint32_t acc = tm + (u % NACC) * astride;
        const int k = u & 3;
        if (ATMEM)
          umma_f16_ts(acc, tm + 448 + k * 8, db + 2 * k, idesc, 1u);
        else
          umma_f16_ss(acc, da + 2 * k, db + 2 * k, idesc, 1u);
      }
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    out[2 * blockIdx.x] = t2 - t0;
    out[2 * blockIdx.x + 1] = t1 - t0;     // issue time alone
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int mma_bench(int n, int naccs, int a_tmem, int reps, int blocks, long long* out_dev, cudaStream_t st) {
  GYRE_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && naccs >= 1 && naccs <= (n <= 128 ? 3 : 1) && reps > 0,
               "mma_bench: bad arguments");
  const int smem = 16384 + 32768 + 1024 + 64;
#define GYRE_MB(NA, AT)                                                                                   \
  if (naccs == NA && (a_tmem != 0) == (AT != 0)) {                                                        \
    GYRE_CHECK_CUDA(cudaFuncSetAttribute(mma_bench_kernel<NA, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    mma_bench_kernel<NA, AT><<<blocks, 128, smem, st>>>(n, reps, out_dev);                                 \
  }
  GYRE_MB(1, 0) GYRE_MB(2, 0) GYRE_MB(3, 0) GYRE_MB(1, 1) GYRE_MB(2, 1) GYRE_MB(3, 1)
#undef GYRE_MB
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gyre
