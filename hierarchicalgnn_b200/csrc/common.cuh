// Shared helpers for libhgnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/hgnn_b200.h"

namespace hgnn {

void set_error(const char* fmt, ...);

inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  set_error("%s", buf);
  return code;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(HGNN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return HGNN_OK;
}

#define HGNN_CUDA_TRY(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return ::hgnn::fail(HGNN_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define HGNN_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::hgnn::fail(HGNN_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

// graph_ops.cu: ordered segmented row sums; skip_short = only the hub segments (> 512 rows; one CTA each), for callers
// whose own kernel already produced the short ones
int launch_segment_reduce(const float* src, int64_t width, const int32_t* gather, const float* weight, const int32_t* perm,
                          const int32_t* rowptr, int64_t n_segments, int mean, float* out, bool skip_short, cudaStream_t st, int64_t src_ld = 0);

// simple bump allocator over the caller-provided workspace
struct Workspace {
  char* base;
  size_t size, used;
  Workspace(void* p, size_t n) : base((char*)p), size(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t off = align_up(used, 256);
    size_t end = off + count * sizeof(T);
    if (end > size || base == nullptr) { used = size + 1; return nullptr; }
    used = end;
    return (T*)(base + off);
  }
  bool ok() const { return used <= size; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- activations (forward value and derivative w.r.t. the pre-activation) ----
__device__ __forceinline__ float act_fwd(int act, float y) {
  switch (act) {
    case HGNN_ACT_GELU: return 0.5f * y * (1.0f + erff(y * 0.70710678118654752f));
    case HGNN_ACT_TANH: return tanhf(y);
    case HGNN_ACT_RELU: return y > 0.f ? y : 0.f;
    case HGNN_ACT_SILU: return y / (1.0f + expf(-y));
    case HGNN_ACT_SIGMOID: return 1.0f / (1.0f + expf(-y));
    default: return y;
  }
}
__device__ __forceinline__ float act_bwd(int act, float y) {
  switch (act) {
    case HGNN_ACT_GELU: {
      float cdf = 0.5f * (1.0f + erff(y * 0.70710678118654752f));
      float pdf = 0.3989422804014327f * expf(-0.5f * y * y);
      return cdf + y * pdf;
    }
    case HGNN_ACT_TANH: { float t = tanhf(y); return 1.0f - t * t; }
    case HGNN_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case HGNN_ACT_SILU: { float s = 1.0f / (1.0f + expf(-y)); return s * (1.0f + y * (1.0f - s)); }
    case HGNN_ACT_SIGMOID: { float s = 1.0f / (1.0f + expf(-y)); return s * (1.0f - s); }
    default: return 1.f;
  }
}

}  // namespace hgnn
