// Hand-ordered softmax inner loop (see ubench_softmax.cu for the compiler-ordered baseline): groups of 8 score pairs
// are staged  A: scale (FFMA2 / FFMA.SAT + Cody-Waite split)  ->  B: exponentials (MUFU.EX2, polynomial FMAs)  ->
// C: pack (F2FP) + overflow OR (LOP3), software-pipelined so that group g's MUFUs are in flight while group g-1 is
// packed and group g+1 is scaled.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_softmax2 scripts/ubench_softmax2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 256;

#define V_ASM asm volatile
__device__ __forceinline__ void ffma2(float& x, float& y, float a0, float a1, float b, float c) {
  V_ASM("{.reg .b64 p, q, r; mov.b64 p, {%2, %3}; mov.b64 q, {%4, %4}; mov.b64 r, {%5, %5}; fma.rn.f32x2 p, p, q, r; mov.b64 {%0, %1}, p;}"
        : "=f"(x), "=f"(y) : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void ffma2v(float& x, float& y, float a0, float a1, float b0, float b1, float c0, float c1) {
  V_ASM("{.reg .b64 p, q, r; mov.b64 p, {%2, %3}; mov.b64 q, {%4, %5}; mov.b64 r, {%6, %7}; fma.rn.f32x2 p, p, q, r; mov.b64 {%0, %1}, p;}"
        : "=f"(x), "=f"(y) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void fadd2(float& x, float& y, float a0, float a1, float c) {
  V_ASM("{.reg .b64 p, r; mov.b64 p, {%2, %3}; mov.b64 r, {%4, %4}; add.rn.f32x2 p, p, r; mov.b64 {%0, %1}, p;}"
        : "=f"(x), "=f"(y) : "f"(a0), "f"(a1), "f"(c));
}
__device__ __forceinline__ void mufu(float& x) { V_ASM("ex2.approx.ftz.f32 %0, %0;" : "+f"(x)); }
__device__ __forceinline__ float fsat(float a, float b, float c) { float r; V_ASM("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; V_ASM("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t or3(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; V_ASM("lop3.b32 %0, %1, %2, %3, 0xfe;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ float shl23add(float q, float r) {
  uint32_t o; V_ASM("{.reg .b32 t; shl.b32 t, %2, 23; add.s32 %0, t, %1;}" : "=r"(o) : "r"(__float_as_uint(q)), "r"(__float_as_uint(r))); return __uint_as_float(o);
}

struct Grp { float x[16]; float r[16]; };   // x: exponent argument / result; r: magic + n of the polynomial pairs

template <int NP>
__device__ __forceinline__ constexpr bool is_poly(int k) {   // k: pair index within the group of 8
  return (NP >= 1 && k == 7) || (NP >= 2 && k == 3) || (NP >= 3 && k == 1) || (NP >= 4 && k == 5);
}

// stage A: 16 scores -> exponent arguments (MUFU pairs) or (f, magic + n) of the polynomial pairs
template <int NP>
__device__ __forceinline__ void stageA(const uint32_t* v, Grp& g, float c, float mcs, float cs, float os) {
  const float kM = 12582912.0f - 125.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float s0 = __uint_as_float(v[2 * k]), s1 = __uint_as_float(v[2 * k + 1]);
    if (is_poly<NP>(k)) {
      const float x0 = fsat(s0, cs, os), x1 = fsat(s1, cs, os);
      ffma2(g.r[2 * k], g.r[2 * k + 1], x0, x1, 126.0f, kM);                       // magic + n
      float m0, m1;
      fadd2(m0, m1, g.r[2 * k], g.r[2 * k + 1], -kM);                               // n + 125
      ffma2v(g.x[2 * k], g.x[2 * k + 1], x0, x1, 126.0f, 126.0f, -m0, -m1);        // f = y - n
    } else {
      ffma2(g.x[2 * k], g.x[2 * k + 1], s0, s1, c, -mcs);
    }
  }
}
// stage B (group gb: exponentials) interleaved with stage C (group gc: pack + OR)
template <int NP, bool HAVE_C>
__device__ __forceinline__ void stageBC(Grp& gb, Grp& gc, uint32_t* w, uint32_t& ovf) {
  uint32_t pk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (is_poly<NP>(k)) {
      float q0, q1;
      ffma2(q0, q1, gb.x[2 * k], gb.x[2 * k + 1], 0.05517084f, 0.24260935f);
      ffma2v(q0, q1, q0, q1, gb.x[2 * k], gb.x[2 * k + 1], 0.69326096f, 0.69326096f);
      ffma2v(q0, q1, q0, q1, gb.x[2 * k], gb.x[2 * k + 1], 0.99992818f, 0.99992818f);
      gb.x[2 * k] = shl23add(q0, gb.r[2 * k]);
      gb.x[2 * k + 1] = shl23add(q1, gb.r[2 * k + 1]);
    } else {
      mufu(gb.x[2 * k]);
      mufu(gb.x[2 * k + 1]);
    }
    if (HAVE_C) {
      pk[k] = pack(gc.x[2 * k], gc.x[2 * k + 1]);
      if (k & 1) ovf = or3(ovf, pk[k - 1], pk[k]);
      w[k] = pk[k];
    }
  }
}

template <int NP>
__global__ void __launch_bounds__(256, 1) k(const float* in, uint32_t* out, unsigned long long* clk, float c) {
  uint32_t v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = __float_as_uint(in[(threadIdx.x * 128 + i) & 4095]);
  uint32_t ovf = 0;
  float mcs = 3.0f;
  uint32_t w[64];
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    const float cs = c * (1.0f / 126.0f), os = (125.0f - mcs) * (1.0f / 126.0f);
    Grp gg[2];
    stageA<NP>(v, gg[0], c, mcs, cs, os);
    stageBC<NP, false>(gg[0], gg[1], w, ovf);
    stageA<NP>(v + 16, gg[1], c, mcs, cs, os);
#pragma unroll
    for (int g = 1; g < 8; ++g) {
      stageBC<NP, true>(gg[g & 1], gg[(g - 1) & 1], w + 8 * (g - 1), ovf);     // exps of g, pack of g - 1
      if (g + 1 < 8) stageA<NP>(v + 16 * (g + 1), gg[(g + 1) & 1], c, mcs, cs, os);
    }
    {
      Grp& gc = gg[1];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        w[56 + k] = pack(gc.x[2 * k], gc.x[2 * k + 1]);
        if (k & 1) ovf = or3(ovf, w[56 + k - 1], w[56 + k]);
      }
    }
    mcs += 0.001f;
    ovf ^= w[0] ^ w[63] ^ w[17];
  }
  const unsigned long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = ovf;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int NP>
void run(const char* name, const float* in, uint32_t* out, unsigned long long* clk) {
  for (int wps : {1, 2}) {
    k<NP><<<148, 128 * wps>>>(in, out, clk, 0.23f);
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
  run<0>("staged, no polynomial", in, out, clk);
  run<1>("staged, polynomial 1/8", in, out, clk);
  run<2>("staged, polynomial 2/8", in, out, clk);
  run<3>("staged, polynomial 3/8", in, out, clk);
  run<4>("staged, polynomial 4/8", in, out, clk);
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
