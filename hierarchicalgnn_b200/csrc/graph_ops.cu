// Graph-side kernels of the HGNN hot path: segment plan (CSR) build, segmented
// weighted row reductions, row gathers, gathered row dots, radius-kNN,
// kNN->edge-list compaction, symmetrize, radius tracker.
// All HBM-bound integer/float streaming work: coalesced float4 rows, one
// deterministic sequential sum per segment (no float atomics).
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"

using namespace hgnn;

// ---------------------------------------------------------------------------
// CSR build
// ---------------------------------------------------------------------------
namespace {

__global__ void k_narrow_iota(const int64_t* __restrict__ keys, int64_t n, int64_t limit, int32_t* __restrict__ k32,
                              int32_t* __restrict__ iota, int32_t* __restrict__ bad) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t k = keys[i];
  if (k < 0 || k >= limit) {
    if (bad) atomicAdd(bad, 1);
    k = k < 0 ? 0 : limit - 1;
  }
  k32[i] = (int32_t)k;
  if (iota) iota[i] = (int32_t)i;
}

__global__ void k_rowptr(const int32_t* __restrict__ sorted, int64_t n, int64_t n_seg, int32_t* __restrict__ rowptr) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n_seg) return;
  // lower_bound(sorted, s)
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted[mid] < (int32_t)s) lo = mid + 1; else hi = mid;
  }
  rowptr[s] = (int32_t)lo;
}

__global__ void k_gather_i32(const int32_t* __restrict__ src, const int32_t* __restrict__ idx, int64_t n,
                             int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}

inline int bits_for(int64_t n) {
  int b = 1;
  while (b < 63 && ((int64_t)1 << b) < n) ++b;
  return b;
}

size_t csr_cub_bytes(int64_t n) {
  size_t t = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n, 0, 32);
  return t;
}

}  // namespace

extern "C" size_t hgnn_csr_build_workspace_bytes(int64_t n_items) {
  if (n_items <= 0) return 256;
  size_t n = (size_t)n_items;
  return 4 * align_up(n * 4, 256) + align_up(csr_cub_bytes(n_items), 256) + 1024;
}

extern "C" int hgnn_csr_build(const int64_t* keys, int64_t n_items, int64_t n_segments, int32_t* perm, int32_t* rowptr,
                              int32_t* keys32, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  HGNN_REQUIRE(n_items >= 0 && n_segments >= 0 && n_items < INT32_MAX && n_segments < INT32_MAX, "csr_build: sizes out of range");
  HGNN_REQUIRE(rowptr != nullptr, "csr_build: rowptr is NULL");
  if (n_items == 0) {
    HGNN_CUDA_TRY(cudaMemsetAsync(rowptr, 0, (size_t)(n_segments + 1) * 4, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(keys && perm, "csr_build: NULL keys/perm");
  HGNN_REQUIRE(n_segments > 0, "csr_build: %lld items but no segments", (long long)n_items);
  Workspace w(ws, ws_bytes);
  int32_t* k_in = w.take<int32_t>(n_items);
  int32_t* k_out = w.take<int32_t>(n_items);
  int32_t* v_in = w.take<int32_t>(n_items);
  size_t cub_bytes = csr_cub_bytes(n_items);
  char* cub_ws = w.take<char>(cub_bytes);
  if (!w.ok()) return fail(HGNN_ERR_WORKSPACE, "csr_build: workspace too small (%zu given)", ws_bytes);
  int T = 256;
  k_narrow_iota<<<(unsigned)((n_items + T - 1) / T), T, 0, st>>>(keys, n_items, n_segments, k_in, v_in, nullptr);
  HGNN_CUDA_TRY(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, k_in, k_out, v_in, perm, (int)n_items, 0,
                                                bits_for(n_segments), st));
  k_rowptr<<<(unsigned)((n_segments + 1 + T - 1) / T), T, 0, st>>>(k_out, n_items, n_segments, rowptr);
  if (keys32) HGNN_CUDA_TRY(cudaMemcpyAsync(keys32, k_in, (size_t)n_items * 4, cudaMemcpyDeviceToDevice, st));
  return check_launch("csr_build");
}

extern "C" int hgnn_index_to_i32(const int64_t* in, int64_t n, int64_t limit, int32_t* out, int32_t* bad, void* stream) {
  if (n <= 0) return HGNN_OK;
  HGNN_REQUIRE(in && out, "index_to_i32: NULL pointer");
  k_narrow_iota<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, n, limit, out, nullptr, bad);
  return check_launch("index_to_i32");
}

// ---------------------------------------------------------------------------
// segmented reduce / gather / dot
// ---------------------------------------------------------------------------
namespace {

template <int VEC>
struct VecT;
template <>
struct VecT<4> { using type = float4; };
template <>
struct VecT<1> { using type = float; };

__device__ __forceinline__ void fma_acc(float4& a, float w, const float4& v) {
  a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
}
__device__ __forceinline__ void fma_acc(float& a, float w, const float& v) { a = fmaf(w, v, a); }
__device__ __forceinline__ void scale_v(float4& a, float s) { a.x *= s; a.y *= s; a.z *= s; a.w *= s; }
__device__ __forceinline__ void scale_v(float& a, float s) { a *= s; }
template <typename V> __device__ __forceinline__ V zero_v();
template <> __device__ __forceinline__ float4 zero_v<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <> __device__ __forceinline__ float zero_v<float>() { return 0.f; }

// one thread = one (segment, column-chunk); consecutive threads = consecutive chunks
template <int VEC>
__global__ void __launch_bounds__(256) k_segment_reduce(const float* __restrict__ src, int chunks, const int32_t* __restrict__ gather,
                                 const float* __restrict__ weight, const int32_t* __restrict__ perm,
                                 const int32_t* __restrict__ rowptr, int64_t n_seg, int mean, float* __restrict__ out,
                                 int long_threshold, int64_t ldv) {  // ldv: source row stride in vector units (chunks when dense)
  using V = typename VecT<VEC>::type;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t s = t / chunks;
  int c = (int)(t - s * chunks);
  if (s >= n_seg) return;
  const V* __restrict__ rows = reinterpret_cast<const V*>(src);
  int beg = rowptr[s], end = rowptr[s + 1];
  if (long_threshold > 0 && end - beg > long_threshold) return;  // hub segment: k_segment_reduce_long owns it
  V acc = zero_v<V>();
  int j = beg;
  // long segments (hub nodes): 16 independent row loads in flight per step, summed in the same row order
  for (; j + 16 <= end; j += 16) {
    V v[16];
    float w[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int i = perm ? perm[j + u] : j + u;
      const int64_t r = gather ? gather[i] : i;
      w[u] = weight ? weight[i] : 1.f;
      v[u] = rows[r * ldv + c];
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) fma_acc(acc, w[u], v[u]);
  }
  // ordinary segments: predicated batches of 8 — every index load, then every row load of the batch is in flight before the
  // first add (a serial tail of dependent perm -> row round trips cost more than the batch body); summed in row order
  for (; j < end; j += 8) {
    V v[8];
    float w[8];
    int64_t r[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool ok = j + u < end;
      const int i = ok ? (perm ? perm[j + u] : j + u) : 0;
      r[u] = ok ? (gather ? (int64_t)gather[i] : (int64_t)i) : -1;
      w[u] = (ok && weight) ? weight[i] : 1.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = r[u] >= 0 ? rows[r[u] * ldv + c] : zero_v<V>();
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r[u] >= 0) fma_acc(acc, w[u], v[u]);
  }
  if (mean) scale_v(acc, 1.0f / (float)max(end - beg, 1));
  reinterpret_cast<V*>(out)[s * chunks + c] = acc;
}

// Hub segments (more than `long_threshold` rows): one CTA per segment, the rows split into blockDim / chunks contiguous
// parts that are summed in parallel (each part in row order, 16 loads in flight) and then combined in part order —
// a fixed association for a given segment length, so still bit-reproducible.
constexpr int LONG_THREADS = 1024;
template <int VEC>
__global__ void __launch_bounds__(LONG_THREADS) k_segment_reduce_long(const float* __restrict__ src, int chunks,
                                                                      const int32_t* __restrict__ gather, const float* __restrict__ weight,
                                                                      const int32_t* __restrict__ perm, const int32_t* __restrict__ rowptr,
                                                                      int64_t n_seg, int mean, float* __restrict__ out, int long_threshold,
                                                                      int64_t ldv) {
  using V = typename VecT<VEC>::type;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  V* part_sum = reinterpret_cast<V*>(sm_raw);
  const V* __restrict__ rows = reinterpret_cast<const V*>(src);
  const int nparts = blockDim.x / chunks;
  const int part = threadIdx.x / chunks, c = threadIdx.x % chunks;
  // this CTA owns segment ids {b, b + grid, b + 2 grid, ...} (consecutive ids — where power-law hubs cluster — land on
  // different CTAs): one parallel look at rowptr finds the (rare) hub segments among them
  __shared__ int s_queue[LONG_THREADS];
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  {
    const int64_t s = (int64_t)threadIdx.x * gridDim.x + blockIdx.x;
    if (s < n_seg && rowptr[s + 1] - rowptr[s] > long_threshold) s_queue[atomicAdd(&s_count, 1)] = (int)s;
  }
  __syncthreads();
  const int n_long = s_count;  // queue order varies run to run; each segment's sum does not depend on it
  for (int qi = 0; qi < n_long; ++qi) {
    const int64_t s = s_queue[qi];
    const int beg = rowptr[s], end = rowptr[s + 1];
    const int per = (end - beg + nparts - 1) / nparts;
    if (part < nparts) {
      const int b0 = min(end, beg + part * per), b1 = min(end, b0 + per);
      V acc = zero_v<V>();
      int j = b0;
      for (; j + 16 <= b1; j += 16) {
        V v[16];
        float w[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int i = perm ? perm[j + u] : j + u;
          const int64_t r = gather ? gather[i] : i;
          w[u] = weight ? weight[i] : 1.f;
          v[u] = rows[r * ldv + c];
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) fma_acc(acc, w[u], v[u]);
      }
      for (; j < b1; ++j) {
        const int i = perm ? perm[j] : j;
        const int64_t r = gather ? gather[i] : i;
        fma_acc(acc, weight ? weight[i] : 1.f, rows[r * ldv + c]);
      }
      part_sum[part * chunks + c] = acc;
    }
    __syncthreads();
    if (part == 0) {
      V acc = part_sum[c];
      for (int p = 1; p < nparts; ++p) fma_acc(acc, 1.f, part_sum[p * chunks + c]);
      if (mean) scale_v(acc, 1.0f / (float)(end - beg));
      reinterpret_cast<V*>(out)[s * chunks + c] = acc;
    }
    __syncthreads();
  }
}

// four rows per thread: four independent index -> row load chains in flight (one row per thread left the kernel at 45 % of the
// HBM roof: 16 bytes in flight per thread cannot cover the latency)
constexpr int GATHER_ROWS_PER_THREAD = 4;
template <int VEC>
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ src, int chunks, const int32_t* __restrict__ idx,
                              const float* __restrict__ weight, int64_t n, float* __restrict__ out) {
  using V = typename VecT<VEC>::type;
  constexpr int R = GATHER_ROWS_PER_THREAD;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t g = t / chunks;
  const int c = (int)(t - g * chunks);
  const int64_t i0 = g * R;
  if (i0 >= n) return;
  int64_t r[R];
  float w[R];
  V v[R];
#pragma unroll
  for (int u = 0; u < R; ++u) {
    const int64_t i = i0 + u;
    r[u] = i < n ? (idx ? (int64_t)idx[i] : i) : -1;
    w[u] = (i < n && weight) ? weight[i] : 1.f;
  }
#pragma unroll
  for (int u = 0; u < R; ++u) v[u] = r[u] >= 0 ? reinterpret_cast<const V*>(src)[r[u] * chunks + c] : zero_v<V>();
#pragma unroll
  for (int u = 0; u < R; ++u) {
    if (r[u] < 0) continue;
    if (weight) scale_v(v[u], w[u]);
    reinterpret_cast<V*>(out)[(i0 + u) * chunks + c] = v[u];
  }
}

// LANES threads cooperate on one item (LANES = 1, 8 or 32)
template <int LANES>
__global__ void __launch_bounds__(256) k_edge_dot(const float* __restrict__ a, const int32_t* __restrict__ ai, const float* __restrict__ b,
                           const int32_t* __restrict__ bi, int width, int64_t n, float* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t i = t / LANES;
  int l = (int)(t % LANES);
  float acc = 0.f;
  if (i < n) {
    const float* ra = a + (int64_t)(ai ? ai[i] : i) * width;
    const float* rb = b + (int64_t)(bi ? bi[i] : i) * width;
    if ((width & 3) == 0) {
      for (int d = l * 4; d < width; d += LANES * 4) {
        float4 x = *reinterpret_cast<const float4*>(ra + d), y = *reinterpret_cast<const float4*>(rb + d);
        acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
      }
    } else {
      for (int d = l; d < width; d += LANES) acc = fmaf(ra[d], rb[d], acc);
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (i < n && l == 0) out[i] = acc;
}

}  // namespace

namespace hgnn {
// launches the per-thread kernel for ordinary segments and the per-CTA kernel for hub segments (> 512 rows)
int launch_segment_reduce(const float* src, int64_t width, const int32_t* gather, const float* weight, const int32_t* perm,
                          const int32_t* rowptr, int64_t n_segments, int mean, float* out, bool skip_short, cudaStream_t st,
                          int64_t src_ld) {
  if (src_ld <= 0) src_ld = width;  // dense rows
  bool vec = (width % 4 == 0) && (src_ld % 4 == 0) && (((uintptr_t)src | (uintptr_t)out) % 16 == 0);
  int chunks = vec ? (int)(width / 4) : (int)width;
  const int64_t ldv = vec ? src_ld / 4 : src_ld;
  const int long_threshold = chunks <= 256 ? 512 : 0;  // hub path needs >= 4 row parts per 1024-thread CTA
  if (!skip_short) {
    int64_t threads = n_segments * chunks;
    unsigned grid = (unsigned)((threads + 255) / 256);
    if (vec) k_segment_reduce<4><<<grid, 256, 0, st>>>(src, chunks, gather, weight, perm, rowptr, n_segments, mean, out, long_threshold, ldv);
    else k_segment_reduce<1><<<grid, 256, 0, st>>>(src, chunks, gather, weight, perm, rowptr, n_segments, mean, out, long_threshold, ldv);
  }
  if (long_threshold > 0) {
    unsigned grid = (unsigned)((n_segments + LONG_THREADS - 1) / LONG_THREADS);  // one segment id per thread to inspect
    size_t smem = (size_t)LONG_THREADS * (vec ? 16 : 4);
    if (vec) k_segment_reduce_long<4><<<grid, LONG_THREADS, smem, st>>>(src, chunks, gather, weight, perm, rowptr, n_segments, mean, out, long_threshold, ldv);
    else k_segment_reduce_long<1><<<grid, LONG_THREADS, smem, st>>>(src, chunks, gather, weight, perm, rowptr, n_segments, mean, out, long_threshold, ldv);
  }
  return check_launch("segment_reduce");
}
}  // namespace hgnn

extern "C" int hgnn_segment_reduce(const float* src, int64_t width, const int32_t* gather, const float* weight,
                                   const int32_t* perm, const int32_t* rowptr, int64_t n_segments, int mean, float* out,
                                   void* stream) {
  if (n_segments <= 0 || width <= 0) return HGNN_OK;
  HGNN_REQUIRE(src && rowptr && out, "segment_reduce: NULL pointer");
  HGNN_REQUIRE(width <= 65536, "segment_reduce: width too large");
  return hgnn::launch_segment_reduce(src, width, gather, weight, perm, rowptr, n_segments, mean, out, false, (cudaStream_t)stream);
}

extern "C" int hgnn_segment_reduce_ld(const float* src, int64_t width, int64_t src_ld, const int32_t* gather, const float* weight,
                                      const int32_t* perm, const int32_t* rowptr, int64_t n_segments, int mean, float* out,
                                      void* stream) {
  if (n_segments <= 0 || width <= 0) return HGNN_OK;
  HGNN_REQUIRE(src && rowptr && out, "segment_reduce: NULL pointer");
  HGNN_REQUIRE(width <= 65536 && src_ld >= width, "segment_reduce: bad width / row stride");
  return hgnn::launch_segment_reduce(src, width, gather, weight, perm, rowptr, n_segments, mean, out, false, (cudaStream_t)stream, src_ld);
}

extern "C" int hgnn_gather_rows(const float* src, int64_t width, const int32_t* idx, const float* weight, int64_t n_items,
                                float* out, void* stream) {
  if (n_items <= 0 || width <= 0) return HGNN_OK;
  HGNN_REQUIRE(src && out, "gather_rows: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = (width % 4 == 0) && (((uintptr_t)src | (uintptr_t)out) % 16 == 0);
  int chunks = vec ? (int)(width / 4) : (int)width;
  int64_t threads = (n_items + GATHER_ROWS_PER_THREAD - 1) / GATHER_ROWS_PER_THREAD * chunks;
  unsigned grid = (unsigned)((threads + 255) / 256);
  if (vec) k_gather_rows<4><<<grid, 256, 0, st>>>(src, chunks, idx, weight, n_items, out);
  else k_gather_rows<1><<<grid, 256, 0, st>>>(src, chunks, idx, weight, n_items, out);
  return check_launch("gather_rows");
}

extern "C" int hgnn_edge_dot(const float* a, const int32_t* ai, const float* b, const int32_t* bi, int64_t width,
                             int64_t n_items, float* out, void* stream) {
  if (n_items <= 0) return HGNN_OK;
  HGNN_REQUIRE(a && b && out && width > 0, "edge_dot: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if ((width & 3) == 0 && (((uintptr_t)a | (uintptr_t)b) % 16 != 0))
    return fail(HGNN_ERR_BAD_ARG, "edge_dot: rows must be 16-byte aligned when width %% 4 == 0");
  if (width <= 16) k_edge_dot<1><<<(unsigned)((n_items + 255) / 256), 256, 0, st>>>(a, ai, b, bi, (int)width, n_items, out);
  else if (width <= 64) k_edge_dot<8><<<(unsigned)((n_items * 8 + 255) / 256), 256, 0, st>>>(a, ai, b, bi, (int)width, n_items, out);
  else k_edge_dot<32><<<(unsigned)((n_items * 32 + 255) / 256), 256, 0, st>>>(a, ai, b, bi, (int)width, n_items, out);
  return check_launch("edge_dot");
}

// ---------------------------------------------------------------------------
// radius kNN (brute force, shared-memory tiled)
// ---------------------------------------------------------------------------
namespace {

constexpr int KNN_THREADS = 128;
constexpr int KNN_TILE = 128;
constexpr int KNN_KMAX = 32;

// One thread per query, references tiled through shared memory, register top-k with strict-< insertion (an equal distance
// never displaces an earlier = smaller index). KCAP = compile-time capacity of the register list (8 / 16 / 32 >= k): the
// insertion chain and the register footprint follow k instead of the maximum. blockIdx.y = reference split: a CTA scans
// the contiguous reference range [split * ref_per_split, ...) and, when there is more than one split, leaves its sorted
// partial list (distance, index) in the workspace for k_knn_merge — small problems (1 200 x 1 200) otherwise occupy ten SMs.
template <int DIM, int KCAP>  // DIM > 0: compile-time dimension with the query in registers; 0: runtime dim (<= 32)
__global__ void __launch_bounds__(KNN_THREADS) k_knn_radius(const float* __restrict__ query, int64_t nq, const float* __restrict__ ref,
                                                           int64_t nr, int dim_rt, int k, float r2, int64_t ref_per_split,
                                                           int64_t* __restrict__ idx, float* __restrict__ part_d, int32_t* __restrict__ part_i,
                                                           const int32_t* __restrict__ qptr, const int32_t* __restrict__ rptr) {
  extern __shared__ __align__(16) float smem[];
  const int dim = DIM > 0 ? DIM : dim_rt;
  float* s_ref = smem;                           // [KNN_TILE][dim]
  float* s_q = smem + KNN_TILE * dim;            // [KNN_THREADS][dim+1] (runtime-dim path only)
  // batched events (qptr / rptr = row offsets of the events, blockIdx.z = event): a query only ever meets the references of
  // its own event — the block-diagonal search a per-event loop would run, in one launch. Indices stay global row numbers.
  int64_t q_lo = 0, q_hi = nq, r_lo = 0, r_hi = nr;
  if (qptr != nullptr) {
    q_lo = qptr[blockIdx.z]; q_hi = qptr[blockIdx.z + 1];
    r_lo = rptr[blockIdx.z]; r_hi = rptr[blockIdx.z + 1];
    if (q_lo + (int64_t)blockIdx.x * KNN_THREADS >= q_hi) return;  // this event has fewer query blocks (whole CTA, before any barrier)
    const int64_t tiles = (r_hi - r_lo + KNN_TILE - 1) / KNN_TILE;
    ref_per_split = (tiles + gridDim.y - 1) / gridDim.y * KNN_TILE;
  }
  int64_t q = q_lo + (int64_t)blockIdx.x * KNN_THREADS + threadIdx.x;
  bool live = q < q_hi;
  float qreg[DIM > 0 ? DIM : 1];
  if (DIM > 0) {
#pragma unroll
    for (int d = 0; d < (DIM > 0 ? DIM : 1); ++d) qreg[d] = live ? query[q * dim + d] : 0.f;
  } else {
    for (int d = 0; d < dim; ++d) s_q[threadIdx.x * (dim + 1) + d] = live ? query[q * dim + d] : 0.f;
  }
  float bd[KCAP];
  int bi[KCAP];
#pragma unroll
  for (int j = 0; j < KCAP; ++j) { bd[j] = INFINITY; bi[j] = -1; }
  float kth = INFINITY;  // current k-th best distance

  const int64_t ref_lo = r_lo + (int64_t)blockIdx.y * ref_per_split;
  const int64_t ref_hi = min(r_hi, ref_lo + ref_per_split);
  for (int64_t base = ref_lo; base < ref_hi; base += KNN_TILE) {
    int cnt = (int)min((int64_t)KNN_TILE, ref_hi - base);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt * dim; t += KNN_THREADS) s_ref[t] = ref[base * dim + t];
    __syncthreads();
    if (!live) continue;
    for (int j = 0; j < cnt; ++j) {
      float d2 = 0.f;
      if (DIM == 8) {  // two 16-byte broadcast loads per reference instead of eight scalar ones; same fma order
        const float4 a = *reinterpret_cast<const float4*>(s_ref + j * 8), b = *reinterpret_cast<const float4*>(s_ref + j * 8 + 4);
        const float rv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int d = 0; d < 8; ++d) { float df = qreg[d] - rv[d]; d2 = fmaf(df, df, d2); }
      } else if (DIM > 0) {
#pragma unroll
        for (int d = 0; d < (DIM > 0 ? DIM : 1); ++d) { float df = qreg[d] - s_ref[j * dim + d]; d2 = fmaf(df, df, d2); }
      } else {
        for (int d = 0; d < dim; ++d) { float df = s_q[threadIdx.x * (dim + 1) + d] - s_ref[j * dim + d]; d2 = fmaf(df, df, d2); }
      }
      if (d2 < r2 && d2 < kth) {
        float cd = d2;
        int ci = (int)(base + j);
        bool placed = false;
#pragma unroll
        for (int m = 0; m < KCAP; ++m) {
          if (m < k) {
            // strict '<' for the new candidate: an equal distance never displaces an earlier (smaller) index. Once it is
            // placed, everything below shifts down by one unconditionally — a displaced entry compared with '<' would
            // jump over an equal-distance entry that was behind it and reverse their order
            if (placed || cd < bd[m]) { float td = bd[m]; int ti = bi[m]; bd[m] = cd; bi[m] = ci; cd = td; ci = ti; placed = true; }
          }
        }
        // refresh the k-th best (bd[k-1]) without dynamic register indexing
#pragma unroll
        for (int m = 0; m < KCAP; ++m) if (m == k - 1) kth = bd[m];
      }
    }
  }
  if (live) {
    if (gridDim.y == 1) {
#pragma unroll
      for (int m = 0; m < KCAP; ++m) if (m < k) idx[q * k + m] = (int64_t)bi[m];
    } else {
      const int64_t o = ((int64_t)blockIdx.y * nq + q) * k;
#pragma unroll
      for (int m = 0; m < KCAP; ++m) if (m < k) { part_d[o + m] = bd[m]; part_i[o + m] = bi[m]; }
    }
  }
}

// merges the per-split sorted lists of one query: k rounds of "smallest head", ties to the lower split = the smaller index
// (splits are increasing reference ranges), i.e. exactly the list a single scan would have produced
__global__ void k_knn_merge(const float* __restrict__ part_d, const int32_t* __restrict__ part_i, int n_split, int64_t nq, int k,
                            int64_t* __restrict__ idx) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  unsigned char head[64];
  for (int s = 0; s < n_split; ++s) head[s] = 0;
  for (int m = 0; m < k; ++m) {
    float best = INFINITY;
    int bs = -1;
    for (int s = 0; s < n_split; ++s) {
      if (head[s] >= k) continue;
      const float d = part_d[((int64_t)s * nq + q) * k + head[s]];
      if (d < best) { best = d; bs = s; }
    }
    if (bs < 0) { idx[q * k + m] = -1; continue; }
    idx[q * k + m] = (int64_t)part_i[((int64_t)bs * nq + q) * k + head[bs]];
    ++head[bs];
  }
}

__global__ void k_knn_count(const int64_t* __restrict__ idx, int64_t nq, int k, int32_t* __restrict__ cnt) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  int c = 0;
  for (int m = 0; m < k; ++m) c += idx[q * k + m] >= 0;
  cnt[q] = c;
}

__global__ void k_knn_fill(const int64_t* __restrict__ idx, int64_t nq, int k, const int32_t* __restrict__ cnt,
                           const int32_t* __restrict__ off, int64_t ld, int64_t* __restrict__ graph, int64_t* __restrict__ total) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  int o = off[q];
  for (int m = 0; m < k; ++m) {
    int64_t v = idx[q * k + m];
    if (v >= 0) { graph[o] = q; graph[ld + o] = v; ++o; }
  }
  if (q == nq - 1) *total = (int64_t)off[q] + cnt[q];
}

}  // namespace

static int knn_splits(int64_t n_query, int64_t n_ref) {
  const int64_t qblocks = (n_query + KNN_THREADS - 1) / KNN_THREADS;
  const int64_t ref_tiles = (n_ref + KNN_TILE - 1) / KNN_TILE;
  int64_t want = (2 * (int64_t)num_sms() + qblocks - 1) / qblocks;  // aim at two CTAs per SM
  want = std::min<int64_t>(std::min<int64_t>(want, ref_tiles), 64);
  return (int)std::max<int64_t>(want, 1);
}

extern "C" size_t hgnn_knn_radius_workspace_bytes(int64_t n_query, int64_t n_ref, int64_t k) {
  const int sp = knn_splits(n_query, n_ref);
  return sp <= 1 ? 0 : (size_t)sp * (size_t)n_query * (size_t)k * 8 + 256;
}

template <int DIM>
static void launch_knn(int kcap, dim3 grid, size_t smem, cudaStream_t st, const float* query, int64_t nq, const float* ref, int64_t nr,
                       int dim, int k, float r2, int64_t per, int64_t* idx, float* pd, int32_t* pi, const int32_t* qptr,
                       const int32_t* rptr) {
  if (kcap == 8) k_knn_radius<DIM, 8><<<grid, KNN_THREADS, smem, st>>>(query, nq, ref, nr, dim, k, r2, per, idx, pd, pi, qptr, rptr);
  else if (kcap == 16) k_knn_radius<DIM, 16><<<grid, KNN_THREADS, smem, st>>>(query, nq, ref, nr, dim, k, r2, per, idx, pd, pi, qptr, rptr);
  else k_knn_radius<DIM, 32><<<grid, KNN_THREADS, smem, st>>>(query, nq, ref, nr, dim, k, r2, per, idx, pd, pi, qptr, rptr);
}

// ws may be NULL (or too small): the scan then runs unsplit, one CTA per 128 queries. n_events > 0: block-diagonal search,
// event b = queries [query_ptr[b], query_ptr[b+1]) against references [ref_ptr[b], ref_ptr[b+1]) (device arrays of n_events + 1
// int32 offsets; the host never reads them: every event gets the grid of the whole problem and idle CTAs leave at once)
static int knn_radius_impl(const float* query, int64_t n_query, const float* ref, int64_t n_ref, int64_t dim, int64_t k, float radius,
                           int64_t* idx, const int32_t* query_ptr, const int32_t* ref_ptr, int64_t n_events, void* ws,
                           size_t ws_bytes, void* stream) {
  if (n_query <= 0 || k <= 0) return HGNN_OK;
  HGNN_REQUIRE(query && idx, "knn_radius: NULL pointer");
  HGNN_REQUIRE(dim >= 1 && dim <= 32, "knn_radius: dim must be in [1,32], got %lld", (long long)dim);
  HGNN_REQUIRE(k <= KNN_KMAX, "knn_radius: k must be <= %d, got %lld", KNN_KMAX, (long long)k);
  HGNN_REQUIRE(n_ref >= 0 && n_ref < INT32_MAX && n_query < INT32_MAX, "knn_radius: n_query / n_ref out of range");
  HGNN_REQUIRE(n_events == 0 || (query_ptr && ref_ptr && n_events > 0 && n_events <= 65535),
               "knn_radius: batched search needs query_ptr, ref_ptr and 1..65535 events");
  cudaStream_t st = (cudaStream_t)stream;
  // per-event problem size for the split heuristic: the average event (the offsets live on the device)
  const int64_t ev = n_events > 0 ? n_events : 1;
  const int64_t nr_ev = (n_ref + ev - 1) / ev;
  int sp = n_events > 0 ? (int)std::max<int64_t>(1, std::min<int64_t>(knn_splits(n_query, nr_ev), (nr_ev + KNN_TILE - 1) / KNN_TILE))
                        : knn_splits(n_query, n_ref);
  if (sp > 1 && (ws == nullptr || ws_bytes < (size_t)sp * (size_t)n_query * (size_t)k * 8 + 256)) sp = 1;
  const int64_t tiles = (n_ref + KNN_TILE - 1) / KNN_TILE;
  int64_t per = (tiles + sp - 1) / sp * KNN_TILE;  // references per split (whole tiles); recomputed per event in the batched kernel
  if (n_events == 0) sp = (int)std::max<int64_t>(1, (n_ref + per - 1) / std::max<int64_t>(per, 1));
  float* pd = nullptr;
  int32_t* pi = nullptr;
  if (sp > 1) {
    pd = (float*)align_up((uintptr_t)ws, 256);
    pi = (int32_t*)(pd + (size_t)sp * n_query * k);
  }
  dim3 grid((unsigned)((n_query + KNN_THREADS - 1) / KNN_THREADS), (unsigned)sp, (unsigned)ev);
  float r2 = radius * radius;
  size_t smem = (size_t)(KNN_TILE * dim + KNN_THREADS * (dim + 1)) * sizeof(float);
  const int kcap = k <= 8 ? 8 : (k <= 16 ? 16 : 32);
  const int32_t* qp = n_events > 0 ? query_ptr : nullptr;
  const int32_t* rp = n_events > 0 ? ref_ptr : nullptr;
  if (dim == 8) launch_knn<8>(kcap, grid, smem, st, query, n_query, ref, n_ref, 8, (int)k, r2, per, idx, pd, pi, qp, rp);
  else launch_knn<0>(kcap, grid, smem, st, query, n_query, ref, n_ref, (int)dim, (int)k, r2, per, idx, pd, pi, qp, rp);
  if (sp > 1) k_knn_merge<<<(unsigned)((n_query + 127) / 128), 128, 0, st>>>(pd, pi, sp, n_query, (int)k, idx);
  return check_launch("knn_radius");
}

extern "C" int hgnn_knn_radius_ws(const float* query, int64_t n_query, const float* ref, int64_t n_ref, int64_t dim, int64_t k,
                                  float radius, int64_t* idx, void* ws, size_t ws_bytes, void* stream) {
  return knn_radius_impl(query, n_query, ref, n_ref, dim, k, radius, idx, nullptr, nullptr, 0, ws, ws_bytes, stream);
}

extern "C" int hgnn_knn_radius(const float* query, int64_t n_query, const float* ref, int64_t n_ref, int64_t dim, int64_t k,
                               float radius, int64_t* idx, void* stream) {
  return knn_radius_impl(query, n_query, ref, n_ref, dim, k, radius, idx, nullptr, nullptr, 0, nullptr, 0, stream);
}

extern "C" int hgnn_knn_radius_batched(const float* query, int64_t n_query, const float* ref, int64_t n_ref, int64_t dim, int64_t k,
                                       float radius, const int32_t* query_ptr, const int32_t* ref_ptr, int64_t n_events,
                                       int64_t* idx, void* ws, size_t ws_bytes, void* stream) {
  HGNN_REQUIRE(n_events >= 1, "knn_radius_batched: n_events must be >= 1");
  return knn_radius_impl(query, n_query, ref, n_ref, dim, k, radius, idx, query_ptr, ref_ptr, n_events, ws, ws_bytes, stream);
}

extern "C" size_t hgnn_knn_edges_workspace_bytes(int64_t n_query) {
  if (n_query <= 0) return 256;
  size_t t = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, t, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n_query);
  return 2 * align_up((size_t)n_query * 4, 256) + align_up(t, 256) + 1024;
}

extern "C" int hgnn_knn_edges(const int64_t* idx, int64_t n_query, int64_t k, int64_t* graph, int64_t* count, void* ws,
                              size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  HGNN_REQUIRE(count != nullptr, "knn_edges: count is NULL");
  if (n_query <= 0 || k <= 0) {
    HGNN_CUDA_TRY(cudaMemsetAsync(count, 0, 8, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(idx && graph, "knn_edges: NULL pointer");
  HGNN_REQUIRE(n_query * k < INT32_MAX, "knn_edges: too many candidate edges");
  Workspace w(ws, ws_bytes);
  int32_t* cnt = w.take<int32_t>(n_query);
  int32_t* off = w.take<int32_t>(n_query);
  size_t t = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, t, (const int32_t*)nullptr, (int32_t*)nullptr, (int)n_query);
  char* cub_ws = w.take<char>(t);
  if (!w.ok()) return fail(HGNN_ERR_WORKSPACE, "knn_edges: workspace too small");
  unsigned grid = (unsigned)((n_query + 255) / 256);
  k_knn_count<<<grid, 256, 0, st>>>(idx, n_query, (int)k, cnt);
  HGNN_CUDA_TRY(cub::DeviceScan::ExclusiveSum(cub_ws, t, cnt, off, (int)n_query, st));
  k_knn_fill<<<grid, 256, 0, st>>>(idx, n_query, (int)k, cnt, off, n_query * k, graph, count);
  return check_launch("knn_edges");
}

// ---------------------------------------------------------------------------
// symmetrize: pack -> radix sort -> unique -> unpack
// ---------------------------------------------------------------------------
namespace {

__global__ void k_pack_both(const int64_t* __restrict__ g, int64_t ld, int64_t n, int64_t nv, uint64_t* __restrict__ keys) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t s = (uint64_t)g[i], d = (uint64_t)g[ld + i];
  keys[i] = s * (uint64_t)nv + d;
  keys[n + i] = d * (uint64_t)nv + s;
}

__global__ void k_unpack(const uint64_t* __restrict__ keys, const int64_t* __restrict__ count, int64_t ld, int64_t nv,
                         int64_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *count) return;
  uint64_t k = keys[i];
  out[i] = (int64_t)(k / (uint64_t)nv);
  out[ld + i] = (int64_t)(k % (uint64_t)nv);
}

struct SymSizes { size_t sort, uniq; };
SymSizes sym_cub_bytes(int64_t n2) {
  SymSizes s{0, 0};
  cub::DeviceRadixSort::SortKeys(nullptr, s.sort, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)n2, 0, 64);
  cub::DeviceSelect::Unique(nullptr, s.uniq, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int64_t*)nullptr, (int)n2);
  return s;
}

}  // namespace

extern "C" size_t hgnn_symmetrize_workspace_bytes(int64_t n_edges) {
  if (n_edges <= 0) return 256;
  int64_t n2 = 2 * n_edges;
  SymSizes s = sym_cub_bytes(n2);
  return 3 * align_up((size_t)n2 * 8, 256) + align_up(s.sort > s.uniq ? s.sort : s.uniq, 256) + 1024;
}

extern "C" int hgnn_symmetrize(const int64_t* graph_in, int64_t ld_in, int64_t n_edges, int64_t n_vertices, int64_t* graph_out,
                               int64_t* count, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  HGNN_REQUIRE(count != nullptr, "symmetrize: count is NULL");
  if (n_edges <= 0) {
    HGNN_CUDA_TRY(cudaMemsetAsync(count, 0, 8, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(graph_in && graph_out && n_vertices > 0, "symmetrize: bad argument");
  HGNN_REQUIRE(2 * n_edges < INT32_MAX && n_vertices < ((int64_t)1 << 31), "symmetrize: sizes out of range");
  int64_t n2 = 2 * n_edges;
  Workspace w(ws, ws_bytes);
  uint64_t* k0 = w.take<uint64_t>(n2);
  uint64_t* k1 = w.take<uint64_t>(n2);
  uint64_t* k2 = w.take<uint64_t>(n2);
  SymSizes s = sym_cub_bytes(n2);
  size_t tb = s.sort > s.uniq ? s.sort : s.uniq;
  char* cub_ws = w.take<char>(tb);
  if (!w.ok()) return fail(HGNN_ERR_WORKSPACE, "symmetrize: workspace too small");
  k_pack_both<<<(unsigned)((n_edges + 255) / 256), 256, 0, st>>>(graph_in, ld_in, n_edges, n_vertices, k0);
  int bits = 2 * bits_for(n_vertices);
  if (bits > 64) bits = 64;
  size_t t1 = s.sort;
  HGNN_CUDA_TRY(cub::DeviceRadixSort::SortKeys(cub_ws, t1, k0, k1, (int)n2, 0, bits, st));
  size_t t2 = s.uniq;
  HGNN_CUDA_TRY(cub::DeviceSelect::Unique(cub_ws, t2, k1, k2, count, (int)n2, st));
  k_unpack<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(k2, count, n2, n_vertices, graph_out);
  return check_launch("symmetrize");
}

// ---------------------------------------------------------------------------
// radius tracker: max Euclidean edge length
// ---------------------------------------------------------------------------
namespace {
__global__ void k_edge_max_dist(const float* __restrict__ a, const int64_t* __restrict__ ai, const float* __restrict__ b,
                                const int64_t* __restrict__ bi, int width, int64_t n, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float d2 = 0.f;
  if (i < n) {
    const float* ra = a + (ai ? ai[i] : i) * width;
    const float* rb = b + (bi ? bi[i] : i) * width;
    for (int d = 0; d < width; ++d) { float df = ra[d] - rb[d]; d2 = fmaf(df, df, d2); }
  }
  float v = sqrtf(d2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0.f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(v));  // v >= 0: int order == float order
}
}  // namespace

extern "C" int hgnn_edge_max_dist(const float* a, const int64_t* ai, const float* b, const int64_t* bi, int64_t width,
                                  int64_t n_items, float* out, void* stream) {
  if (n_items <= 0) return HGNN_OK;
  HGNN_REQUIRE(a && b && out && width > 0, "edge_max_dist: bad argument");
  k_edge_max_dist<<<(unsigned)((n_items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, ai, b, bi, (int)width, n_items, out);
  return check_launch("edge_max_dist");
}
