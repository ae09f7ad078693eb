// Microbenchmark: throughput of scalar FADD/FMUL/FFMA vs packed FADD2/FMUL2/FFMA2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o packed_rate packed_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fmul(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  constexpr int N = 16;  // independent chains
  float s[N]; u64 p[N];
  for (int i = 0; i < N; i++) { s[i] = seed + i + threadIdx.x; p[i] = ((u64)__float_as_uint(s[i]) << 32) | __float_as_uint(s[i] + 0.5f); }
  float c = seed * 0.999f; u64 c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      if (MODE == 0) s[i] = fadd(s[i], c);
      if (MODE == 1) s[i] = fmul(s[i], c);
      if (MODE == 2) s[i] = ffma(s[i], c, c);
      if (MODE == 3) p[i] = add2(p[i], c2);
      if (MODE == 4) p[i] = mul2(p[i], c2);
      if (MODE == 5) p[i] = fma2(p[i], c2, c2);
      if (MODE == 6) { s[i] = fadd(s[i], c); p[i] = add2(p[i], c2); }   // mix scalar + packed
    }
  }
  float acc = 0; for (int i = 0; i < N; i++) acc += s[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> void run(const char* name, int ops_per_inst, int insts_per_iter_per_chain) {
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * 8);
  int iters = 4096; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {
    k<MODE><<<148, warps * 32>>>(out, 16, 1.0f);
    cudaEventRecord(e0); k<MODE><<<148, warps * 32>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)iters * 16 * insts_per_iter_per_chain * warps;   // warp-instructions per SM
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%-8s warps/SM %2d: %.3f warp-inst/cycle/SM (%.2f per SMSP), %.1f lane-ops/cycle/SM, %.3f ms\n", name, warps, inst / cyc, inst / cyc / 4,
           inst / cyc * 32 * ops_per_inst, ms);
  }
  cudaFree(out);
}
int main() {
  run<0>("FADD", 1, 1); run<1>("FMUL", 1, 1); run<2>("FFMA", 1, 1);
  run<3>("FADD2", 2, 1); run<4>("FMUL2", 2, 1); run<5>("FFMA2", 2, 1); run<6>("FADD+FADD2", 1, 2);
  return 0;
}
