// Probe: register <-> (lane, column) mapping of tcgen05.ld.16x256b.x8 on sm_100a. Each warp writes row*256+col with the
// 32x32b shape (thread = lane/row), then reads 16 lanes x 64 columns with the 16x256b shape and prints what landed where.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(int* out) {
  __shared__ uint32_t s_t;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_t)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_t + ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = row * 256 + c0 + i;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tmem + c0), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(half * 16) << 16))
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; ++i) out[((warp * 2 + half) * 32 + lane) * 32 + i] = (int)r[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_t), "r"(64) : "memory");
}
int main() {
  int* d; cudaMalloc(&d, 4 * 2 * 32 * 32 * 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  static int h[4 * 2 * 32 * 32];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int half = 0; half < 2; ++half) for (int t = 0; t < 32; ++t) for (int i = 0; i < 32; ++i) {
    int v = h[((w * 2 + half) * 32 + t) * 32 + i], row = v / 256, col = v % 256;
    int j = i / 4, k4 = i % 4;
    int erow = w * 32 + half * 16 + t / 4 + (k4 >= 2 ? 8 : 0), ecol = 8 * j + 2 * (t % 4) + (k4 & 1);
    if (row != erow || col != ecol) { if (bad < 20) printf("w%d h%d t%d r%d: got (row %d, col %d) expected (%d, %d)\n", w, half, t, i, row, col, erow, ecol); ++bad; }
  }
  printf("mismatches vs the mma-accumulator-style guess: %d\n", bad);
  for (int t = 0; t < 8; ++t) { printf("t%d:", t); for (int i = 0; i < 8; ++i) { int v = h[t * 32 + i]; printf(" (%d,%d)", v / 256, v % 256); } printf("\n"); }
  return 0;
}
