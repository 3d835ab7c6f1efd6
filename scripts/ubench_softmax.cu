// The softmax inner loop of the attention kernels in isolation (no TMEM, no barriers): 128 scores per thread in
// registers -> p = 2^(s*c - m) -> packed fp16, OR-reduced.  Reports clocks per 128-score row-tile per warp at 1 / 2 / 4
// warps per SM sub-partition for several polynomial shares and instruction orders.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_softmax scripts/ubench_softmax.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 256;

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fma_sat(float a, float b, float c) { float r; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

__device__ __forceinline__ float2 exp2_poly_sat(float2 xs) {
  const float kM = 12582912.0f - 125.0f;
  const float2 r = __ffma2_rn(xs, make_float2(126.0f, 126.0f), make_float2(kM, kM));
  const float2 m = __fadd2_rn(r, make_float2(-kM, -kM));
  const float2 f = __ffma2_rn(xs, make_float2(126.0f, 126.0f), make_float2(-m.x, -m.y));
  float2 q = __ffma2_rn(make_float2(0.05517084f, 0.05517084f), f, make_float2(0.24260935f, 0.24260935f));
  q = __ffma2_rn(q, f, make_float2(0.69326096f, 0.69326096f));
  q = __ffma2_rn(q, f, make_float2(0.99992818f, 0.99992818f));
  float2 o;
  o.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(r.x) << 23));
  o.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(r.y) << 23));
  return o;
}
// scalar variant: same maths with FFMA / FADD (1 clk each) instead of the 2-clk packed forms
__device__ __forceinline__ float exp2_poly_sat1(float xs) {
  const float kM = 12582912.0f - 125.0f;
  const float r = fmaf(xs, 126.0f, kM);
  const float m = r - kM;
  const float f = fmaf(xs, 126.0f, -m);
  float q = fmaf(0.05517084f, f, 0.24260935f);
  q = fmaf(q, f, 0.69326096f);
  q = fmaf(q, f, 0.99992818f);
  return __int_as_float(__float_as_int(q) + (__float_as_int(r) << 23));
}

// MODE bits: low 3 = polynomial pairs per 8 (0, 1, 2, 3, 4); 8 = scalar polynomial; 16 = scalar scale (FFMA instead of FFMA2)
template <int MODE>
__device__ __forceinline__ void exp32(const uint32_t (&v)[32], float c, float mcs, float cs, float os, uint32_t& ovf) {
  constexpr int NP = MODE & 7;
  uint32_t w[16];
#pragma unroll
  for (int pair = 0; pair < 16; ++pair) {
    const int i = 2 * pair;
    const int k = pair & 7;
    const bool poly = (NP >= 1 && k == 7) || (NP >= 2 && k == 3) || (NP >= 3 && k == 1) || (NP >= 4 && k == 5);
    float2 e;
    if (poly) {
      float2 xs;
      xs.x = fma_sat(__uint_as_float(v[i]), cs, os);
      xs.y = fma_sat(__uint_as_float(v[i + 1]), cs, os);
      if constexpr ((MODE & 8) != 0) { e.x = exp2_poly_sat1(xs.x); e.y = exp2_poly_sat1(xs.y); }
      else e = exp2_poly_sat(xs);
    } else {
      float2 x;
      if constexpr ((MODE & 16) != 0) { x.x = fmaf(__uint_as_float(v[i]), c, -mcs); x.y = fmaf(__uint_as_float(v[i + 1]), c, -mcs); }
      else x = __ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), make_float2(c, c), make_float2(-mcs, -mcs));
      e.x = ex2(x.x);
      e.y = ex2(x.y);
    }
    const __half2 hh = __floats2half2_rn(e.x, e.y);
    w[pair] = *reinterpret_cast<const uint32_t*>(&hh);
  }
  uint32_t a0 = ovf, a1 = 0;
#pragma unroll
  for (int k = 0; k < 16; k += 4) { a0 |= w[k] | w[k + 1]; a1 |= w[k + 2] | w[k + 3]; }
  ovf = a0 | a1;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* in, uint32_t* out, unsigned long long* clk, float c) {
  uint32_t v0[32], v1[32], v2[32], v3[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    v0[i] = __float_as_uint(in[(threadIdx.x * 128 + i) & 4095]);
    v1[i] = __float_as_uint(in[(threadIdx.x * 128 + 32 + i) & 4095]);
    v2[i] = __float_as_uint(in[(threadIdx.x * 128 + 64 + i) & 4095]);
    v3[i] = __float_as_uint(in[(threadIdx.x * 128 + 96 + i) & 4095]);
  }
  uint32_t ovf = 0;
  float mcs = 3.0f;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    const float cs = c * (1.0f / 126.0f), os = (125.0f - mcs) * (1.0f / 126.0f);
    exp32<MODE>(v0, c, mcs, cs, os, ovf);
    exp32<MODE>(v1, c, mcs, cs, os, ovf);
    exp32<MODE>(v2, c, mcs, cs, os, ovf);
    exp32<MODE>(v3, c, mcs, cs, os, ovf);
    mcs += 0.001f;
  }
  const unsigned long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = ovf;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, const float* in, uint32_t* out, unsigned long long* clk) {
  for (int wps : {1, 2, 4}) {
    k<MODE><<<148, 128 * wps>>>(in, out, clk, 0.23f);
    cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("%-44s warps/SMSP %d: %7.1f clk per 128-score tile-row per SMSP (MUFU-only floor 1024)\n", name, wps, avg / (ITERS * wps));
  }
}

int main() {
  float* in; uint32_t* out; unsigned long long* clk;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&clk, 148 * 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = -20.0f + 0.01f * (i % 977);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("no polynomial", in, out, clk);
  run<1>("polynomial 1/8 (packed)", in, out, clk);
  run<2>("polynomial 2/8 (packed)", in, out, clk);
  run<3>("polynomial 3/8 (packed)", in, out, clk);
  run<4>("polynomial 4/8 (packed)", in, out, clk);
  run<8 + 2>("polynomial 2/8 (scalar)", in, out, clk);
  run<8 + 3>("polynomial 3/8 (scalar)", in, out, clk);
  run<8 + 4>("polynomial 4/8 (scalar)", in, out, clk);
  run<16 + 8 + 3>("polynomial 3/8 (scalar), scalar scale", in, out, clk);
  run<16 + 8 + 4>("polynomial 4/8 (scalar), scalar scale", in, out, clk);
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
