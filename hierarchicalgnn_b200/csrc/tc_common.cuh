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
// same, with a suspend-time hint (ns): the hardware parks the thread until the phase completes or the time is up, so a
// waiting role costs (almost) no issue slots — plain try_wait returns after a few tens of cycles and the loop around it
// competes with the warps that have work
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (CUDA error) instead of hanging the GPU
#ifdef HGNN_DEBUG_MBAR
// debug builds (-DHGNN_DEBUG_MBAR): the first wait that times out records who / where, every later wait falls through, the
// kernel ends (with garbage) and the host reads the record (per translation unit: hgnn_tc_debug_mbar_timeout in edge_tc.cu)
static __device__ int g_mbar_dbg[8];
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*(volatile int*)&g_mbar_dbg[0] != 0) return;
    if (clock64() - t0 > 200000000LL) {
      if (atomicCAS(&g_mbar_dbg[0], 0, 1) == 0) {
        g_mbar_dbg[1] = (int)blockIdx.x; g_mbar_dbg[2] = (int)threadIdx.x; g_mbar_dbg[3] = (int)bar; g_mbar_dbg[4] = (int)parity;
        uint32_t dyn; asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn)); g_mbar_dbg[5] = (int)dyn;
        __threadfence();
      }
      return;
    }
  }
}
#else
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t n = 0;
  long long t0 = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {   // parked up to 20 us per attempt
    if ((++n & 255u) == 0) {                           // the clock is read once per 256 attempts
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 8000000000LL) __trap();  // seconds: a protocol bug becomes a CUDA error, not a hang
    }
  }
}
#endif
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// contiguous global -> shared bulk copy, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// L2 residency hints for rows that are read twice per tile with ~100 MB of streaming traffic in between (the fp32 skip
// rows): first read evict_last (stay), second read evict_first (done with it)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float4* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol));
  return v;
}
// non-blocking test of a phase (no suspend): for schedulers that have something else to do
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// 16-byte asynchronous global -> shared copy (LDGSTS): no register staging; src_bytes < 16 zero-fills the rest
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
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

// ---- transcendental pieces for the epilogues (bf16 path: error budget << bf16 rounding of the outputs) ----
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }  // rcp(+inf) = +0

// hardware tanh (MUFU.TANH, one special-function slot; max relative error 2^-11, far below the bf16 rounding of every
// value that passes through it here)
__device__ __forceinline__ float tanh_approx(float x) { float r; asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---- packed fp32 pairs: FFMA2 / FADD2 / FMUL2 issue two lanes of arithmetic per slot on sm_100 (profiles/micro/ffma2.cu);
// the LayerNorm / activation epilogues are issue-bound, so everything that can be paired is ----
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

// GELU(y) = y * Phi(y) with Phi(y) ~= sigmoid(z), z = y * (c0 + c1 y^2 + c2 y^4), y^2 clamped to 64.
// Coefficients fitted against the exact erf form: max |error| of GELU = 2.8e-5 over the reals
// (profiles/fit_gelu.py). The backward's derivative takes sigmoid(z) = 1/2 + 1/2 tanh(z/2) through the hardware tanh (one
// special-function op instead of ex2 + rcp: its 2^-11 lands on gradients that are rounded to bf16 right after).
constexpr float kG0 = 1.5949518066125585f, kG1 = 0.07406933810902123f, kG2 = -0.0007124377042967553f;
constexpr float kLog2e = 1.4426950408889634f;
struct GeluParts { float sig, y2; };
__device__ __forceinline__ GeluParts gelu_sigmoid(float y) {
  GeluParts g;
  g.y2 = fminf(y * y, 64.0f);
  const float zh = y * fmaf(fmaf(0.5f * kG2, g.y2, 0.5f * kG1), g.y2, 0.5f * kG0);  // z / 2
  g.sig = fmaf(0.5f, tanh_approx(zh), 0.5f);
  return g;
}
// forward: ex2 + rcp (1e-6) — g feeds GEMM2 and the LayerNorm behind it amplifies what the hardware tanh would add
__device__ __forceinline__ float fast_gelu(float y) {
  const float y2 = fminf(y * y, 64.0f);
  const float zn = y * fmaf(fmaf(-kG2 * kLog2e, y2, -kG1 * kLog2e), y2, -kG0 * kLog2e);  // -z * log2(e)
  return y * rcp_approx(1.0f + ex2_approx(zn));
}
__device__ __forceinline__ float2 fast_gelu2(float2 y) {
  float2 y2 = mul2(y, y);
  y2.x = fminf(y2.x, 64.0f);
  y2.y = fminf(y2.y, 64.0f);
  const float2 p = fma2(fma2(splat2(-kG2 * kLog2e), y2, splat2(-kG1 * kLog2e)), y2, splat2(-kG0 * kLog2e));
  const float2 zn = mul2(y, p);
  const float2 d = add2(make_float2(ex2_approx(zn.x), ex2_approx(zn.y)), splat2(1.0f));
  return mul2(y, make_float2(rcp_approx(d.x), rcp_approx(d.y)));
}
// exact derivative of fast_gelu: sig * (1 + y (1 - sig) z'(y)),  z' = c0 + 3 c1 y^2 + 5 c2 y^4
__device__ __forceinline__ float fast_gelu_bwd(float y) {
  const GeluParts g = gelu_sigmoid(y);
  const float zp = fmaf(fmaf(5.0f * kG2, g.y2, 3.0f * kG1), g.y2, kG0);
  return g.sig * fmaf(fmaf(-y, g.sig, y), zp, 1.0f);
}
__device__ __forceinline__ float2 fast_gelu_bwd2(float2 y) {
  float2 y2 = mul2(y, y);
  y2.x = fminf(y2.x, 64.0f);
  y2.y = fminf(y2.y, 64.0f);
  const float2 p = fma2(fma2(splat2(0.5f * kG2), y2, splat2(0.5f * kG1)), y2, splat2(0.5f * kG0));
  const float2 zh = mul2(y, p);
  const float2 sig = fma2(splat2(0.5f), make_float2(tanh_approx(zh.x), tanh_approx(zh.y)), splat2(0.5f));
  const float2 zp = fma2(fma2(splat2(5.0f * kG2), y2, splat2(3.0f * kG1)), y2, splat2(kG0));
  const float2 w = fma2(make_float2(-y.x, -y.y), sig, y);  // y (1 - sig)
  return mul2(sig, fma2(w, zp, splat2(1.0f)));
}
// tanh(y) = 1 - 2 / (1 + e^{2y}); saturates correctly through ex2 -> inf / 0. The OUTPUT activation of the edge / node
// networks lands in fp32 latents, so it keeps the ex2 + rcp form (1e-6) instead of the hardware tanh (5e-4).
__device__ __forceinline__ float fast_tanh(float y) { return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(y * (2.0f * kLog2e))), 1.0f); }
__device__ __forceinline__ float2 fast_tanh2(float2 y) {
  const float2 a = mul2(y, splat2(2.0f * kLog2e));
  const float2 d = add2(make_float2(ex2_approx(a.x), ex2_approx(a.y)), splat2(1.0f));
  return fma2(splat2(-2.0f), make_float2(rcp_approx(d.x), rcp_approx(d.y)), splat2(1.0f));
}
// A&S 7.1.26 erf (|err| <= 1.5e-7) kept for activations that need it
__device__ __forceinline__ float fast_erf(float x) {
  float ax = fabsf(x);
  float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
  return copysignf(1.0f - poly * __expf(-ax * ax), x);
}
// compile-time activation: one code path per kernel instantiation (keeps the SASS inside the I-cache)
template <int ACT>
__device__ __forceinline__ float tc_act(float y) {
  if constexpr (ACT == HGNN_ACT_GELU) return fast_gelu(y);
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
    return fast_gelu_bwd(y);
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

// pairs: the GELU / Tanh of the HGNN edge and node networks have packed forms, the rest go lane by lane
template <int ACT>
__device__ __forceinline__ float2 tc_act2(float2 y) {
  if constexpr (ACT == HGNN_ACT_GELU) return fast_gelu2(y);
  else if constexpr (ACT == HGNN_ACT_TANH) return fast_tanh2(y);
  else return make_float2(tc_act<ACT>(y.x), tc_act<ACT>(y.y));
}
template <int ACT>
__device__ __forceinline__ float2 tc_act_bwd2(float2 y) {
  if constexpr (ACT == HGNN_ACT_GELU) {
    return fast_gelu_bwd2(y);
  } else if constexpr (ACT == HGNN_ACT_TANH) {
    const float2 t = make_float2(tanh_approx(y.x), tanh_approx(y.y));
    return fma2(make_float2(-t.x, -t.y), t, splat2(1.0f));
  } else {
    return make_float2(tc_act_bwd<ACT>(y.x), tc_act_bwd<ACT>(y.y));
  }
}
// 8 bf16 (one 16-byte chunk of an image / x-hat stash) -> 4 fp32 pairs: a shift / a mask per value
__device__ __forceinline__ void unpack8_2(const uint4& q, float2 (&f)[4]) {
  f[0] = make_float2(__uint_as_float(q.x << 16), __uint_as_float(q.x & 0xffff0000u));
  f[1] = make_float2(__uint_as_float(q.y << 16), __uint_as_float(q.y & 0xffff0000u));
  f[2] = make_float2(__uint_as_float(q.z << 16), __uint_as_float(q.z & 0xffff0000u));
  f[3] = make_float2(__uint_as_float(q.w << 16), __uint_as_float(q.w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf16(float2 v) { return pack_bf16(v.x, v.y); }

// ---- LayerNorm epilogue pieces over a TMEM row (lane = row, this thread owns NCH*32 consecutive columns) ----
// partial statistics of (accumulator + bias): single pass around a pivot taken from the row itself
// (no catastrophic cancellation: |pivot - mean| is of the order of the row's spread)
template <int NCH>
__device__ __forceinline__ void ln_partial(uint32_t taddr, const float* __restrict__ sbias, float& mean_loc, float& m2_loc) {
  float v[32];
  float pv = 0.f;
  float2 sum = make_float2(0.f, 0.f), sq = make_float2(0.f, 0.f), npv = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    tmem_ld32(taddr + ch * 32, v);
    const float4* b4 = reinterpret_cast<const float4*>(sbias + ch * 32);
    if (ch == 0) { pv = v[0] + sbias[0]; npv = splat2(-pv); }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = b4[i];
      const float2 d0 = add2(add2(make_float2(v[4 * i], v[4 * i + 1]), make_float2(b.x, b.y)), npv);
      const float2 d1 = add2(add2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(b.z, b.w)), npv);
      sum = add2(sum, add2(d0, d1));
      sq = fma2(d0, d0, sq);
      sq = fma2(d1, d1, sq);
    }
  }
  constexpr float inv = 1.0f / (NCH * 32);
  const float sm = sum.x + sum.y, sqs = sq.x + sq.y;
  mean_loc = fmaf(sm, inv, pv);
  m2_loc = fmaxf(sqs - sm * sm * inv, 0.f);
}

// act(LayerNorm(accumulator + bias)) -> bf16, written as the K-major swizzled A-operand image of the next GEMM.
// sb / sg / sbe are the full per-column parameter arrays; c0 = first column owned by this thread.
// xh_out (optional): the normalised pre-affine values xhat = (h - mean) * rstd are also stashed in HBM as bf16, 8 per
// uint4, in [column chunk][row] order (a warp's 32 rows of one chunk are 512 contiguous bytes) for the backward pass.
template <int ACT, int NCH>
__device__ __forceinline__ void ln_act_to_image(uint32_t taddr, const float* __restrict__ sb, const float* __restrict__ sg,
                                                const float* __restrict__ sbe, int c0, float mean, float rstd,
                                                uint8_t* __restrict__ img, int row, uint4* __restrict__ xh_out = nullptr) {
  float v[32];
  const float2 rs2 = splat2(rstd), nmr2 = splat2(-mean * rstd);
#pragma unroll 1
  for (int ch = 0; ch < NCH; ++ch) {
    tmem_ld32(taddr + ch * 32, v);
    const int cb = c0 + ch * 32;
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      const int c = cb + g8 * 8;
      float2 o[4], xh[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 b = *reinterpret_cast<const float4*>(sb + c + 4 * h);
        const float4 g = *reinterpret_cast<const float4*>(sg + c + 4 * h);
        const float4 be = *reinterpret_cast<const float4*>(sbe + c + 4 * h);
        xh[2 * h] = fma2(add2(make_float2(v[g8 * 8 + 4 * h], v[g8 * 8 + 4 * h + 1]), make_float2(b.x, b.y)), rs2, nmr2);
        xh[2 * h + 1] = fma2(add2(make_float2(v[g8 * 8 + 4 * h + 2], v[g8 * 8 + 4 * h + 3]), make_float2(b.z, b.w)), rs2, nmr2);
        o[2 * h] = tc_act2<ACT>(fma2(xh[2 * h], make_float2(g.x, g.y), make_float2(be.x, be.y)));
        o[2 * h + 1] = tc_act2<ACT>(fma2(xh[2 * h + 1], make_float2(g.z, g.w), make_float2(be.z, be.w)));
      }
      *reinterpret_cast<uint4*>(img + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3)) =
          make_uint4(pack_bf16(o[0]), pack_bf16(o[1]), pack_bf16(o[2]), pack_bf16(o[3]));
      if (xh_out)
        xh_out[(size_t)(c >> 3) * TILE_M + row] = make_uint4(pack_bf16(xh[0]), pack_bf16(xh[1]), pack_bf16(xh[2]), pack_bf16(xh[3]));
    }
  }
}

// What the forward edge step leaves in HBM for its backward (one caller-owned buffer, hgnn_tc_edge_stash_bytes):
//   a0  [tiles][L/64][16 KB]   bf16 tile image of the edge-latent columns e of the gathered input (weight-gradient operand;
//                              the x[src] / x[dst] columns are not kept: their weight gradient is formed per node)
//   g   [tiles][2L/64][16 KB]  bf16 tile image of act(LN1(h1))                                (weight-gradient operand)
//   xh1 [tiles][2L/8][128]     uint4 = 8 bf16 of xhat1 = (h1 - mean1) rstd1                   (LN1 / activation adjoint)
//   xh2 [tiles][L/8][128]      uint4 = 8 bf16 of xhat2                                        (LN2 / activation adjoint)
//   rstd [tiles][2][128]       fp32 rstd1, rstd2 per row
struct EdgeStash {
  size_t a0, g, xh1, xh2, rstd, total;
  int64_t tiles;
};
__host__ __device__ inline EdgeStash edge_stash_layout(int64_t n_edges, int L) {
  EdgeStash S{};
  S.tiles = (n_edges + TILE_M - 1) / TILE_M;
  size_t off = 0;
  S.a0 = off;   off += (size_t)S.tiles * (L / KBLK) * A_BLK_BYTES;
  S.g = off;    off += (size_t)S.tiles * (2 * L / KBLK) * A_BLK_BYTES;
  S.xh1 = off;  off += (size_t)S.tiles * (2 * L / 8) * TILE_M * 16;
  S.xh2 = off;  off += (size_t)S.tiles * (L / 8) * TILE_M * 16;
  S.rstd = off; off += (size_t)S.tiles * 2 * TILE_M * 4;
  S.total = off;
  return S;
}

// ---- pieces shared by the fused edge kernels and the per-layer row kernels ----
// (gathers assume 256-thread CTAs: 16 threads per 256 B row piece, 16 rows per pass)
// gather one K-block (64 fp32 columns starting at `col0` of rows rowid[r] of `base`): 16 threads per 256 B row piece,
// 16 rows per pass. Split in two so the loads of block k+1 are in flight while block k is converted, stored and multiplied.
__device__ __forceinline__ void gather_load(float4 (&v)[8], const float* __restrict__ base, int ld, const int* __restrict__ rowid, int col0) {
  const int sub = threadIdx.x & 15, rr = threadIdx.x >> 4;
#pragma unroll
  for (int p = 0; p < 8; ++p) {  // a negative row id marks a padding row of the last tile: zeros (also in the saved image)
    const int rid = rowid[p * 16 + rr];
    v[p] = rid >= 0 ? __ldg(reinterpret_cast<const float4*>(base + (size_t)rid * ld + col0) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// same, with an L2 cache-policy hint on the loads
__device__ __forceinline__ void gather_load_hint(float4 (&v)[8], const float* __restrict__ base, int ld, const int* __restrict__ rowid, int col0,
                                                 uint64_t pol) {
  const int sub = threadIdx.x & 15, rr = threadIdx.x >> 4;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int rid = rowid[p * 16 + rr];
    v[p] = rid >= 0 ? ldg_f4_hint(reinterpret_cast<const float4*>(base + (size_t)rid * ld + col0) + sub, pol) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// gimg (optional): the same swizzled bf16 block is also left in HBM — the backward pass and the weight-gradient GEMM
// read these tile images back with bulk copies instead of re-gathering
__device__ __forceinline__ void gather_store(uint8_t* __restrict__ blk, const float4 (&v)[8], uint8_t* __restrict__ gimg = nullptr) {
  const int sub = threadIdx.x & 15, rr = threadIdx.x >> 4;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const int r = p * 16 + rr;
    const uint2 pk = make_uint2(pack_bf16(v[p].x, v[p].y), pack_bf16(v[p].z, v[p].w));
    const uint32_t off = sw128_off(r, sub >> 1) + (sub & 1) * 8;
    *reinterpret_cast<uint2*>(blk + off) = pk;
    if (gimg) *reinterpret_cast<uint2*>(gimg + off) = pk;
  }
}
__device__ __forceinline__ void gather_a_block(uint8_t* __restrict__ blk, const float* __restrict__ base, int ld,
                                               const int* __restrict__ rowid, int col0) {
  float4 v[8];
  gather_load(v, base, ld, rowid, col0);
  gather_store(blk, v);
}

// gather one K-block (64 bf16 columns starting at col0) of rows rowid[r] of a bf16 [rows, ld] matrix: 8 threads per 128 B row
// piece, 32 rows per pass, 16 registers per thread — two K-blocks fit in flight where one fp32 block did. The chunks land in
// the swizzled operand block as they are (no conversion).
__device__ __forceinline__ void gather_load_bf16(uint4 (&v)[4], const uint16_t* __restrict__ base, int ld, const int* __restrict__ rowid, int col0) {
  const int c16 = threadIdx.x & 7, rr = threadIdx.x >> 3;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int rid = rowid[p * 32 + rr];
    v[p] = rid >= 0 ? __ldg(reinterpret_cast<const uint4*>(base + (size_t)rid * ld + col0) + c16) : make_uint4(0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ void gather_store_bf16(uint8_t* __restrict__ blk, const uint4 (&v)[4]) {
  const int c16 = threadIdx.x & 7, rr = threadIdx.x >> 3;
#pragma unroll
  for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4*>(blk + sw128_off(p * 32 + rr, c16)) = v[p];
}

struct LnStat { float mean, rstd; };

// Chan-combine the two half-row partials (n each): returns mean / rstd of the full row
__device__ __forceinline__ LnStat combine_halves(const float* red, int r, int n_half, float eps) {
  float m0 = red[r * 4 + 0], q0 = red[r * 4 + 1], m1 = red[r * 4 + 2], q1 = red[r * 4 + 3];
  float mean = 0.5f * (m0 + m1);
  float d = m1 - m0;
  float m2 = q0 + q1 + d * d * (0.5f * n_half);  // Chan: M2 = M2a + M2b + delta^2 * na*nb/(na+nb)
  LnStat s;
  s.mean = mean;
  s.rstd = rsqrtf(m2 / (2.0f * n_half) + eps);
  return s;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// column sums over the 32 rows held by a warp: in v[c] = this lane's value for column c;
// returns the sum over lanes of column `lane` (31 shuffles, log-step register transpose)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    bool up = lane & 16;
    float send = up ? v[i] : v[i + 16], keep = up ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    bool up = lane & 8;
    float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bool up = lane & 4;
    float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    bool up = lane & 2;
    float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    bool up = lane & 1;
    float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

// mean / rstd of a row from 4 equal partial (mean, M2) pairs (Chan et al.), n values each
__device__ __forceinline__ void combine4(const float* red, int r, int n, float eps, float& mean, float& rstd) {
  float m[4], q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = red[r * 8 + 2 * i]; q[i] = red[r * 8 + 2 * i + 1]; }
  mean = 0.25f * (m[0] + m[1] + m[2] + m[3]);
  float m2 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { float d = m[i] - mean; m2 += q[i] + (float)n * d * d; }
  rstd = rsqrtf(m2 / (4.0f * n) + eps);
}


// weight-gradient GEMM over tile images (wgrad_tc.cu): dW[CA, CB] (+offsets into a larger matrix) = img_a^T img_b
struct WgradProblem {
  const uint8_t* img_a; int ca_total, ca0, ca;   // columns of A used: [ca0, ca0+ca), ca multiple of 128
  const uint8_t* img_b; int cb_total, cb0, cb;   // columns of B used, cb multiple of 64, <= 256
  float* out; int ld, row_off, col_off, transpose;
};
size_t wgrad_workspace_bytes(const WgradProblem* probs, int n, int splits);
int wgrad_splits(int n_roles, int n_tiles);
// optional passenger of the ordered reduce launch: out[i] = sum_p part[p][i] over n_part partial rows of `width` floats, in order
struct ColumnSums { const float* part; int n_part, width; float* out; };
int launch_wgrad(const WgradProblem* probs, int n, int n_tiles, void* ws, size_t ws_bytes, cudaStream_t st,
                 const ColumnSums* colsums = nullptr);
// fp32 [rows, cols] (cols % 64 == 0) -> bf16 tile image, padding rows of the last tile zeroed
int launch_make_image(const float* src, int64_t rows, int cols, uint8_t* img, cudaStream_t st);

}  // namespace tc
}  // namespace hgnn
