// Measurement-only kernels (not on any product path): tcgen05.mma issue/latency characteristics on this part.
#include "common.cuh"
#include "ops.h"

namespace gyre {

// One CTA: a single thread issues `reps` kind::f16 MMAs (M=128, K=16, N=n) round-robin over `naccs` accumulators
// (TMEM columns acc * 128), operands either both in shared memory or A in TMEM; returns clocks from first issue
// to completion (commit -> mbarrier).  Operand contents are irrelevant.
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int n, int naccs, int a_tmem, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t acc = tm + (r % naccs) * (n <= 128 ? 128 : 256);
      const int k = r & 3;
      if (a_tmem)
        umma_f16_ts(acc, tm + 448 + k * 8, db + 2 * k, idesc, 1u);
      else
        umma_f16_ss(acc, da + 2 * k, db + 2 * k, idesc, 1u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int mma_bench(int n, int naccs, int a_tmem, int reps, int blocks, long long* out_dev, cudaStream_t st) {
  GYRE_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && naccs >= 1 && naccs <= (n <= 128 ? 3 : 1) && reps > 0,
               "mma_bench: bad arguments");
  const int smem = 16384 + 32768 + 1024 + 64;
  static bool attr = false;
  if (!attr) {
    GYRE_CHECK_CUDA(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  mma_bench_kernel<<<blocks, 128, smem, st>>>(n, naccs, a_tmem, reps, out_dev);
  GYRE_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gyre
