// Instruction-issue microbenchmark for the softmax inner loop (sm_100a): clocks per warp instruction per SM
// sub-partition for MUFU.EX2, F2FP (cvt.rn.f16x2.f32), FFMA2, FFMA, LOP3, FMNMX and mixes, at 1 / 2 / 4 warps per
// sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_pipes scripts/ubench_pipes.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int U = 8;   // independent chains per thread

template <int OP>
__global__ void k(float* out, unsigned long long* clk, float seed) {
  float a[U], b[U];
  unsigned u[U];
#pragma unroll
  for (int i = 0; i < U; ++i) { a[i] = seed + 0.001f * (i + threadIdx.x); b[i] = seed * 0.5f - 0.002f * i; u[i] = threadIdx.x * 77 + i; }
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (OP == 0) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
      if (OP == 1) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b[i])); a[i] = __uint_as_float(u[i]); }
      if (OP == 2) { asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%1, %0}; fma.rn.f32x2 x, x, y, y; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(b[i])); }
      if (OP == 3) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b[i])); }
      if (OP == 4) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % U]), "r"(u[(i + 3) % U])); }
      if (OP == 5) { asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i])); }
      if (OP == 6) {   // MUFU + F2FP per pair as in the softmax: 2 ex2, 1 cvt
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b[i]));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b[i]));
      }
      if (OP == 7) { asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i])); }
      if (OP == 8) { unsigned short h; asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(a[i])); a[i] = __uint_as_float(h); }
      if (OP == 9) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b[i])); a[i] = __uint_as_float(u[i]); }
      if (OP == 10) { asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) % U])); }
      if (OP == 11) { asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % U])); }
      if (OP == 12) { asm volatile("fma.rn.sat.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b[i])); }
      if (OP == 13) { asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%1, %0}; add.rn.f32x2 x, x, y; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(b[i])); }
      if (OP == 14) { asm volatile("shl.b32 %0, %0, 23;" : "+r"(u[i])); asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % U])); }
      if (OP == 15) { asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % U])); }
      if (OP == 16) {   // the softmax fast path per pair without polynomial: FFMA2, 2 ex2, cvt, (1/2) lop3
        asm volatile("{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%1, %0}; fma.rn.f32x2 x, x, y, y; mov.b64 {%0, %1}, x;}" : "+f"(a[i]), "+f"(b[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b[i]));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(b[i]));
      }
    }
  }
  const unsigned long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < U; ++i) s += a[i] + b[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int inst_per_unit) {
  float* out;
  unsigned long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8);
  for (int wps = 1; wps <= 4; wps *= 2) {
    const int threads = 128 * wps;
    k<OP><<<148, threads>>>(out, clk, 0.37f);
    k<OP><<<148, threads>>>(out, clk, 0.37f);
    cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double per_inst = avg / (double(ITERS) * U * inst_per_unit * wps);
    printf("%-34s warps/SMSP %d  clk per warp-instr per SMSP %6.2f\n", name, wps, per_inst);
  }
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  run<0>("MUFU.EX2 f32", 1);
  run<1>("F2FP cvt.rn.f16x2.f32", 1);
  run<9>("F2FP cvt.rn.bf16x2.f32", 1);
  run<8>("cvt.rn.f16.f32", 1);
  run<7>("MUFU ex2.approx.f16x2", 1);
  run<2>("FFMA2", 1);
  run<13>("FADD2", 1);
  run<3>("FFMA", 1);
  run<12>("FFMA.SAT", 1);
  run<15>("HFMA2", 1);
  run<4>("LOP3", 1);
  run<5>("FMNMX", 1);
  run<10>("PRMT", 1);
  run<11>("IADD", 1);
  run<14>("SHL+IADD (2 instr)", 2);
  run<6>("2 MUFU + 1 F2FP (3 instr)", 3);
  run<16>("FFMA2 + 2 MUFU + F2FP (4 instr)", 4);
  cudaError_t e = cudaGetLastError();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
