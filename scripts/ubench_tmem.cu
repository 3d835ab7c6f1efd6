// TMEM load / store throughput microbenchmark (sm_100a): clocks per tcgen05.ld / tcgen05.st 32x32b.x32 per warp with
// 1, 2, 4 (one per lane quadrant) and 8 warps (two per quadrant) issuing concurrently.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_tmem scripts/ubench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 512;

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// mode 0: loads, wait after every 4; mode 1: stores, wait after every 4; mode 2: loads with a wait after each
__global__ void k(int nwarps, int mode, uint32_t* out, unsigned long long* clk) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  uint32_t acc = 0;
  unsigned long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    st32(base, r);
    st32(base + 32, r);
    st32(base + 64, r);
    st32(base + 96, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
      if (mode == 0) {
        uint32_t a[32], b[32], c[32], d[32];
        ld32(base, a); ld32(base + 32, b); ld32(base + 64, c); ld32(base + 96, d);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += a[3] ^ b[5] ^ c[7] ^ d[11];
      } else if (mode == 1) {
        r[0] = acc + it;
        st32(base, r); st32(base + 32, r); st32(base + 64, r); st32(base + 96, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else {
        uint32_t a[32];
        for (int q = 0; q < 4; ++q) {
          ld32(base + q * 32, a);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          acc += a[q];
        }
      }
    }
    t1 = clock64();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (warp < nwarps && (threadIdx.x & 31) == 0) clk[blockIdx.x * 8 + warp] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  uint32_t* out;
  unsigned long long* clk;
  cudaMalloc(&out, 148 * 256 * 4);
  cudaMalloc(&clk, 148 * 8 * 8);
  const char* names[3] = {"tcgen05.ld x32 (wait per 4)", "tcgen05.st x32 (wait per 4)", "tcgen05.ld x32 (wait per 1)"};
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 2, 4, 8}) {
      cudaMemset(clk, 0, 148 * 8 * 8);
      k<<<148, 256>>>(nw, mode, out, clk);
      cudaDeviceSynchronize();
      unsigned long long h[148 * 8];
      cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < 148 * 8; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%-30s warps %d: %7.1f clk per x32 access per warp (4 KB each) -> %6.1f B/clk/SM\n", names[mode], nw,
             mx / (ITERS * 4.0), nw * 4096.0 * ITERS * 4.0 / mx);
    }
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
