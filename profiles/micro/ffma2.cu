// Microbenchmark: issue / pipe throughput of packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against scalar FFMA on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
template <int MODE> __global__ void k(float* out, int iters, float s) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  unsigned long long p[8];
  for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i + 1]) << 32) | __float_as_uint(a[2 * i]);
  unsigned long long ss = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
  int lop = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(s));
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, ss);
    } else if (MODE == 2) {  // scalar FFMA + as many integer ops (issue-slot competition)
#pragma unroll
      for (int i = 0; i < 16; ++i) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(s)); asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(lop) : "r"(it)); }
    } else {                 // FFMA2 + the same integer ops
#pragma unroll
      for (int i = 0; i < 8; ++i) { p[i] = fma2(p[i], ss, ss); asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(lop) : "r"(it)); asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(lop) : "r"(it + 1)); }
    }
  }
  float r = lop;
  for (int i = 0; i < 16; ++i) r += a[i];
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* out) {
  const int iters = 4096, grid = 148 * 4, block = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, block>>>(out, 16, 0.999f);
  cudaEventRecord(e0);
  k<MODE><<<grid, block>>>(out, iters, 0.999f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)grid * block * iters * 16;
  printf("%-28s %.3f ms  %.1f TFMA/s (scalar-equivalent FMAs)\n", name, ms, fma / ms / 1e9);
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  run<0>("FFMA x16", out); run<1>("FFMA2 x8", out); run<2>("FFMA x16 + LOP3 x16", out); run<3>("FFMA2 x8 + LOP3 x16", out);
  return 0;
}
