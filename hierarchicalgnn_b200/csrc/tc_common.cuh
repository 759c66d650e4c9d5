// Blackwell (sm_100a) primitives used by the tensor-core kernels: mbarrier,
// bulk async copy, tcgen05 (UMMA) issue / commit / TMEM alloc / TMEM load,
// shared-memory + instruction descriptors, and the 128B-swizzled K-major
// operand layout shared by the weight packer and the in-kernel A-tile writers.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace hgnn {
namespace tc {

constexpr int KBLK = 64;                 // bf16 elements per K-block = one 128-byte swizzle row
constexpr int ROW_BYTES = 128;           // bytes per operand row inside a K-block
constexpr int TILE_M = 128;              // rows per tile = UMMA M = TMEM lanes
constexpr int A_BLK_BYTES = TILE_M * ROW_BYTES;  // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of the 16-byte chunk `c16` (0..7) of row `r` inside a K-block image (Swizzle<3,4,3>)
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t c16) {
  return r * ROW_BYTES + ((c16 ^ (r & 7u)) << 4);
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (CUDA error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// contiguous global -> shared bulk copy, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tensor core / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05 ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// K-major, 128B-swizzled operand: 8-row groups 1024 B apart (SBO), LBO unused (=1), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;                       // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;  // stride byte offset
  d |= (uint64_t)1 << 46;                       // descriptor version
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// MN-major, 128B-swizzled operand over the SAME tile image, read "transposed": the image holds
// [128 K-rows x 64 MN-elements] per column block (row = 128 B, 8-row swizzle atoms of 1024 B).
// SBO = distance between 8-row K groups (1024 B), LBO = distance between 64-wide MN blocks.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor with both operands MN-major (D[m,n] = sum_k A[k,m] B[k,n])
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// one K-block (64 bf16 = 4 UMMA K-steps) of an M=128 x N tile
__device__ __forceinline__ void umma_kblock(uint32_t tmem_d, uint32_t a_saddr, uint32_t b_saddr, uint32_t idesc, bool first) {
  uint64_t ad = make_smem_desc(a_saddr), bd = make_smem_desc(b_saddr);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    umma_bf16(tmem_d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (first && k == 0) ? 0u : 1u);  // +32 B per K-step
  }
}

// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits)
  return *reinterpret_cast<uint32_t*>(&h);
}

// ---- fast, accurate-enough transcendental pieces for the epilogues ----
// erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7): one rcp + one ex2 + 6 fma
__device__ __forceinline__ float fast_erf(float x) {
  float ax = fabsf(x);
  float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  float r = 1.0f - poly * __expf(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float fast_tanh(float x) {
  float t = __expf(2.0f * fminf(fmaxf(x, -15.0f), 15.0f));  // tanh(15) == 1 in fp32; keeps the divide in range
  return 1.0f - __fdividef(2.0f, t + 1.0f);
}
// compile-time activation: one code path per kernel instantiation (keeps the SASS inside the I-cache)
template <int ACT>
__device__ __forceinline__ float tc_act(float y) {
  if constexpr (ACT == HGNN_ACT_GELU) return 0.5f * y * (1.0f + fast_erf(y * 0.70710678118654752f));
  else if constexpr (ACT == HGNN_ACT_TANH) return fast_tanh(y);
  else if constexpr (ACT == HGNN_ACT_RELU) return fmaxf(y, 0.f);
  else if constexpr (ACT == HGNN_ACT_SILU) return __fdividef(y, 1.0f + __expf(fminf(-y, 80.0f)));
  else if constexpr (ACT == HGNN_ACT_SIGMOID) return __fdividef(1.0f, 1.0f + __expf(fminf(-y, 80.0f)));
  else return y;
}
// derivative w.r.t. the pre-activation, same fast-math pieces
template <int ACT>
__device__ __forceinline__ float tc_act_bwd(float y) {
  if constexpr (ACT == HGNN_ACT_GELU) {
    float ex = __expf(-0.5f * y * y);
    float ax = fabsf(y) * 0.70710678118654752f;
    float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
    float poly = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
    float cdf = 0.5f * (1.0f + copysignf(1.0f - poly * ex, y));
    return fmaf(y * 0.3989422804014327f, ex, cdf);
  } else if constexpr (ACT == HGNN_ACT_TANH) {
    float t = fast_tanh(y);
    return 1.0f - t * t;
  } else if constexpr (ACT == HGNN_ACT_RELU) {
    return y > 0.f ? 1.f : 0.f;
  } else if constexpr (ACT == HGNN_ACT_SILU) {
    float sg = __fdividef(1.0f, 1.0f + __expf(fminf(-y, 80.0f)));
    return sg * (1.0f + y * (1.0f - sg));
  } else if constexpr (ACT == HGNN_ACT_SIGMOID) {
    float sg = __fdividef(1.0f, 1.0f + __expf(fminf(-y, 80.0f)));
    return sg * (1.0f - sg);
  } else {
    return 1.f;
  }
}

}  // namespace tc
}  // namespace hgnn
