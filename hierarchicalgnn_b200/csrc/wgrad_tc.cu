// Tensor-core weight-gradient kernel (sm_100a):  dW[ca, cb] = sum_rows A[row, ca] * B[row, cb]
// with rows = edges (the reduction dimension, K of the MMA).
//
// Operands arrive as "tile images": per 128-row tile, per 64-column block, one 16 KB
// [128 rows x 128 B] bf16 array with the 128B XOR swizzle — exactly the layout the
// backward-data kernel leaves in shared memory, bulk-copied to HBM as is. Here each
// image is bulk-copied back into shared memory (no per-thread gather, no conversion)
// and fed to tcgen05.mma as an MN-major operand, i.e. read transposed by the tensor
// core itself. A CTA accumulates its tile range in TMEM (split-K over edges) and
// writes one fp32 partial; a deterministic second stage sums the partials in order.
#include <algorithm>

#include "tc_common.cuh"

using namespace hgnn;
using namespace hgnn::tc;

namespace {

constexpr int WG_THREADS = 128;
constexpr int WG_MAX_ROLES = 4;

struct WgRole {
  const uint8_t* img_a;  // [tiles][na_total][16 KB]
  const uint8_t* img_b;
  int na_total, a0, m_halves;   // A uses column blocks [a0, a0 + 2*m_halves): M = 128 per half
  int nb_total, b0, nb;         // B uses column blocks [b0, b0 + nb): N = 64*nb <= 256
  float* partial;               // [splits][M_total x N]
};

struct WgArgs {
  WgRole role[WG_MAX_ROLES];
  int n_roles, splits;
  int n_tiles;
};

__global__ void __launch_bounds__(WG_THREADS, 1) k_tc_wgrad(WgArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* const base = smem_raw;
  const int r = blockIdx.x / args.splits, split = blockIdx.x % args.splits;
  const WgRole R = args.role[r];
  const int na = 2 * R.m_halves, nb = R.nb;
  const int N = 64 * nb;
  const uint32_t stage_bytes = (uint32_t)(na + nb) * A_BLK_BYTES;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(base + 2 * stage_bytes);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 8);
  const uint32_t base_u = smem_u32(base), bar0 = smem_u32(s_bar);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto FREE = [&](int s) { return bar0 + 16u + 8u * s; };
  const uint32_t ACC = bar0 + 32u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int tcols = 32;
  while (tcols < N * R.m_halves) tcols <<= 1;

  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(bar0 + 8u * i, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), tcols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const int per = (args.n_tiles + args.splits - 1) / args.splits;
  const int t0 = split * per, t1 = min(args.n_tiles, t0 + per);
  const uint32_t idesc = make_idesc_mn(TILE_M, N);

  if (tid == 0 && t0 < t1) {
    auto load = [&](int t, int s) {
      mbar_expect_tx(FULL(s), stage_bytes);
      const uint32_t dst = base_u + s * stage_bytes;
      bulk_g2s(dst, R.img_a + ((size_t)t * R.na_total + R.a0) * A_BLK_BYTES, na * A_BLK_BYTES, FULL(s));
      bulk_g2s(dst + na * A_BLK_BYTES, R.img_b + ((size_t)t * R.nb_total + R.b0) * A_BLK_BYTES, nb * A_BLK_BYTES, FULL(s));
    };
    load(t0, 0);
    for (int t = t0; t < t1; ++t) {
      const int i = t - t0, s = i & 1;
      if (t + 1 < t1) {  // prefetch the next tile into the other stage once its previous MMAs retired
        const int i2 = i + 1, s2 = i2 & 1;
        mbar_wait(FREE(s2), ((i2 >> 1) & 1) ^ 1);
        load(t + 1, s2);
      }
      mbar_wait(FULL(s), (i >> 1) & 1);
      tc_fence_after();
      const uint32_t a_s = base_u + s * stage_bytes, b_s = a_s + na * A_BLK_BYTES;
      for (int mh = 0; mh < R.m_halves; ++mh) {
#pragma unroll
        for (int k = 0; k < TILE_M / 16; ++k) {  // 16 K-rows (2 swizzle atoms = 2048 B) per MMA
          uint64_t ad = make_smem_desc_mn(a_s + mh * 2 * A_BLK_BYTES + k * 2048, A_BLK_BYTES);
          uint64_t bd = make_smem_desc_mn(b_s + k * 2048, A_BLK_BYTES);
          umma_bf16(tmem + mh * N, ad, bd, idesc, (i == 0 && k == 0) ? 0u : 1u);
        }
      }
      umma_commit(FREE(s));
    }
    umma_commit(ACC);
  }
  const int Mtot = 128 * R.m_halves;
  float* out = R.partial + (size_t)split * Mtot * N;
  if (t0 < t1) {
    mbar_wait(ACC, 0);
    tc_fence_after();
    float v[32];
    for (int mh = 0; mh < R.m_halves; ++mh) {
      const int row = mh * 128 + warp * 32 + lane;
      for (int c0 = 0; c0 < N; c0 += 32) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + mh * N + c0, v);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(out + (size_t)row * N + c0 + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
      }
    }
  } else {
    for (int i = tid; i < Mtot * N; i += WG_THREADS) out[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, tcols);
}

// out[(row_off + m) * ld + col_off + n] (or transposed) = sum_splits partial[s][m][n]; one launch, blockIdx.y = problem
struct WgReduceArgs {
  const float* partial[WG_MAX_ROLES];
  float* out[WG_MAX_ROLES];
  int M[WG_MAX_ROLES], N[WG_MAX_ROLES], ld[WG_MAX_ROLES], row_off[WG_MAX_ROLES], col_off[WG_MAX_ROLES], transpose[WG_MAX_ROLES];
  int splits, n_prob;
  ColumnSums cs;  // blockIdx.y == n_prob: the column-sum passenger
};
__global__ void k_wgrad_reduce(WgReduceArgs a) {
  const int p = blockIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (p == a.n_prob) {
    // column-sum passenger: a block owns 32 columns; its 8 warps sum 8 consecutive ranges of the partial rows (a fixed
    // split for a given n_part), combined in range order through shared memory: bit-reproducible, 8x shorter chains
    __shared__ float s_part[8][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    if (blockIdx.x * 32 >= a.cs.width) return;
    const int per = (a.cs.n_part + 7) / 8;
    const int k0 = min(a.cs.n_part, grp * per), k1 = min(a.cs.n_part, k0 + per);
    float s = 0.f;
    if (col < a.cs.width) {
      const float* part = a.cs.part + col;
      int k = k0;
      for (; k + 8 <= k1; k += 8) {  // eight loads in flight, summed in partial order
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = part[(size_t)(k + u) * a.cs.width];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
      }
      for (; k < k1; ++k) s += part[(size_t)k * a.cs.width];
    }
    s_part[grp][lane] = s;
    __syncthreads();
    if (grp == 0 && col < a.cs.width) {
      float t = s_part[0][lane];
#pragma unroll
      for (int g = 1; g < 8; ++g) t += s_part[g][lane];
      a.cs.out[col] = t;
    }
    return;
  }
  const int M = a.M[p], N = a.N[p];
  if (i >= M * N) return;
  const float* partial = a.partial[p];
  float s = 0.f;
  int k = 0;
  for (; k + 4 <= a.splits; k += 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = partial[(size_t)(k + u) * M * N + i];
#pragma unroll
    for (int u = 0; u < 4; ++u) s += v[u];
  }
  for (; k < a.splits; ++k) s += partial[(size_t)k * M * N + i];
  int m = i / N, n = i % N;
  if (a.transpose[p]) a.out[p][(size_t)(a.row_off[p] + n) * a.ld[p] + a.col_off[p] + m] = s;
  else a.out[p][(size_t)(a.row_off[p] + m) * a.ld[p] + a.col_off[p] + n] = s;
}

// fp32 [rows, cols] -> tile images (test helper / generic producer); rows beyond `rows` are zero
__global__ void k_make_image(const float* __restrict__ src, int64_t rows, int cols, uint8_t* __restrict__ img) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (padded row, 8-column group)
  int groups = cols / 8;
  int64_t prow = t / groups;
  int g = (int)(t % groups);
  int64_t n_tiles = (rows + TILE_M - 1) / TILE_M;
  if (prow >= n_tiles * TILE_M) return;
  int64_t tile = prow / TILE_M;
  int r = (int)(prow % TILE_M);
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (prow < rows)
    for (int i = 0; i < 8; ++i) v[i] = src[prow * cols + g * 8 + i];
  uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  *reinterpret_cast<uint4*>(img + ((size_t)tile * (cols / 64) + g / 8) * A_BLK_BYTES + sw128_off(r, g % 8)) = pk;
}

}  // namespace

namespace hgnn {
namespace tc {

size_t wgrad_workspace_bytes(const WgradProblem* probs, int n, int splits) {
  size_t tot = 0;
  for (int i = 0; i < n; ++i) tot += align_up((size_t)splits * probs[i].ca * probs[i].cb * 4, 256);
  return tot + 256;
}

int launch_make_image(const float* src, int64_t rows, int cols, uint8_t* img, cudaStream_t st) {
  HGNN_REQUIRE(src && img && rows > 0 && cols % KBLK == 0, "make_image: bad argument");
  const int64_t tiles = (rows + TILE_M - 1) / TILE_M, t = tiles * TILE_M * (cols / 8);
  k_make_image<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(src, rows, cols, img);
  return check_launch("make_image");
}

// split-K over the row tiles: at most one CTA per SM over all problems of the launch, and at least WG_MIN_TILES tiles per CTA
// (every split costs an [M, N] fp32 partial written and read back: with one tile per split a 12 000-row layer moved
// 24 MB of partials for a 0.4 MB gradient)
constexpr int WG_MIN_TILES = 4;
int wgrad_splits(int n_roles, int n_tiles) {
  int s = std::max(1, num_sms() / std::max(1, n_roles));
  return std::max(1, std::min(s, (n_tiles + WG_MIN_TILES - 1) / WG_MIN_TILES));
}

int launch_wgrad(const WgradProblem* probs, int n, int n_tiles, void* ws, size_t ws_bytes, cudaStream_t st, const ColumnSums* colsums) {
  HGNN_REQUIRE(n >= 1 && n <= WG_MAX_ROLES, "wgrad: 1..%d problems per launch", WG_MAX_ROLES);
  WgArgs a{};
  a.n_roles = n;
  a.n_tiles = n_tiles;
  a.splits = wgrad_splits(n, n_tiles);
  if (ws_bytes < wgrad_workspace_bytes(probs, n, a.splits)) return fail(HGNN_ERR_WORKSPACE, "wgrad: workspace too small");
  char* w = (char*)ws;
  size_t max_stage = 0;
  for (int i = 0; i < n; ++i) {
    const WgradProblem& p = probs[i];
    HGNN_REQUIRE(p.ca % 128 == 0 && p.ca >= 128 && p.ca <= 256 && p.cb % 64 == 0 && p.cb >= 64 && p.cb <= 256 &&
                 (p.ca / 128) * p.cb <= 512, "wgrad: unsupported tile shape %d x %d", p.ca, p.cb);
    WgRole& R = a.role[i];
    R.img_a = p.img_a; R.na_total = p.ca_total / 64; R.a0 = p.ca0 / 64; R.m_halves = p.ca / 128;
    R.img_b = p.img_b; R.nb_total = p.cb_total / 64; R.b0 = p.cb0 / 64; R.nb = p.cb / 64;
    R.partial = (float*)w;
    w += align_up((size_t)a.splits * p.ca * p.cb * 4, 256);
    max_stage = std::max(max_stage, (size_t)(2 * R.m_halves + R.nb) * A_BLK_BYTES);
  }
  size_t smem = 2 * max_stage + 128;
  HGNN_REQUIRE(smem <= 227 * 1024, "wgrad: stage too large for shared memory");
  HGNN_CUDA_TRY(cudaFuncSetAttribute(k_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_tc_wgrad<<<n * a.splits, WG_THREADS, smem, st>>>(a);
  WgReduceArgs ra{};
  ra.splits = a.splits;
  ra.n_prob = n;
  int max_total = 0;
  if (colsums) { ra.cs = *colsums; max_total = colsums->width * 8; }  // 32 columns per 256-thread block
  for (int i = 0; i < n; ++i) {
    const WgradProblem& p = probs[i];
    ra.partial[i] = a.role[i].partial; ra.out[i] = p.out;
    ra.M[i] = p.ca; ra.N[i] = p.cb; ra.ld[i] = p.ld; ra.row_off[i] = p.row_off; ra.col_off[i] = p.col_off; ra.transpose[i] = p.transpose;
    max_total = std::max(max_total, p.ca * p.cb);
  }
  k_wgrad_reduce<<<dim3((max_total + 255) / 256, n + (colsums ? 1 : 0)), 256, 0, st>>>(ra);
  return check_launch("tc_wgrad");
}

}  // namespace tc
}  // namespace hgnn

// Unit-test entry: out[CA, CB] = bf16(A[rows, CA])^T . bf16(B[rows, CB]) through tile images + MN-major UMMA.
extern "C" size_t hgnn_tc_debug_wgrad_workspace_bytes(int64_t rows, int64_t ca, int64_t cb) {
  int64_t tiles = (rows + TILE_M - 1) / TILE_M;
  return (size_t)tiles * TILE_M * (ca + cb) * 2 + (size_t)num_sms() * ca * cb * 4 + 4096;
}

extern "C" int hgnn_tc_debug_wgrad(const float* A, const float* B, int64_t rows, int64_t ca, int64_t cb, float* out, void* ws,
                                   size_t ws_bytes, void* stream) {
  HGNN_REQUIRE(A && B && out && ws && rows > 0, "tc_debug_wgrad: bad argument");
  HGNN_REQUIRE(ca % 128 == 0 && cb % 64 == 0, "tc_debug_wgrad: ca %% 128 == 0 and cb %% 64 == 0 required");
  HGNN_REQUIRE(ws_bytes >= hgnn_tc_debug_wgrad_workspace_bytes(rows, ca, cb), "tc_debug_wgrad: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t tiles = (rows + TILE_M - 1) / TILE_M;
  uint8_t* img_a = (uint8_t*)ws;
  uint8_t* img_b = img_a + align_up((size_t)tiles * TILE_M * ca * 2, 1024);
  char* rest = (char*)(img_b + align_up((size_t)tiles * TILE_M * cb * 2, 1024));
  int64_t ta = tiles * TILE_M * (ca / 8), tb = tiles * TILE_M * (cb / 8);
  k_make_image<<<(unsigned)((ta + 255) / 256), 256, 0, st>>>(A, rows, (int)ca, img_a);
  k_make_image<<<(unsigned)((tb + 255) / 256), 256, 0, st>>>(B, rows, (int)cb, img_b);
  hgnn::tc::WgradProblem p{img_a, (int)ca, 0, (int)ca, img_b, (int)cb, 0, (int)cb, out, (int)cb, 0, 0, 0};
  return hgnn::tc::launch_wgrad(&p, 1, (int)tiles, rest, ws_bytes - (size_t)(rest - (char*)ws), st);
}

// ---------------------------------------------------------------------------------------------------------------------
// dW[n_out, fan_in] = delta^T A from two tile images (the A-operand images hgnn_tc_gemm leaves behind): tiled into
// <= 256 x 128 / 128 x 256 problems of the split-K kernel above, four per launch.
namespace {
int build_wgrad_problems(const uint8_t* d_img, int n_out, const uint8_t* a_img, int fan_in, float* dW, hgnn::tc::WgradProblem* pr,
                         int max_pr) {
  int n = 0;
  if (n_out % 128 == 0) {  // delta supplies the M side (M = 128 per half)
    for (int ca0 = 0; ca0 < n_out; ca0 += 256) {
      const int ca = std::min(256, n_out - ca0);
      const int cb_max = ca == 256 ? 128 : 256;  // keeps a double-buffered stage inside shared memory
      for (int cb0 = 0; cb0 < fan_in; cb0 += cb_max) {
        const int cb = std::min(cb_max, fan_in - cb0);
        if (n == max_pr) return -1;
        pr[n++] = hgnn::tc::WgradProblem{d_img, n_out, ca0, ca, a_img, fan_in, cb0, cb, dW, fan_in, ca0, cb0, 0};
      }
    }
  } else if (fan_in % 128 == 0) {  // narrow fan-out (64): the input image supplies the M side, result stored transposed
    for (int ca0 = 0; ca0 < fan_in; ca0 += 256) {
      const int ca = std::min(256, fan_in - ca0);
      const int cb_max = ca == 256 ? 128 : 256;
      for (int cb0 = 0; cb0 < n_out; cb0 += cb_max) {
        const int cb = std::min(cb_max, n_out - cb0);
        if (n == max_pr) return -1;
        pr[n++] = hgnn::tc::WgradProblem{a_img, fan_in, ca0, ca, d_img, n_out, cb0, cb, dW, fan_in, cb0, ca0, 1};
      }
    }
  } else {
    return -2;
  }
  return n;
}
constexpr int WG_MAX_PROBLEMS = 32;
}  // namespace

extern "C" int hgnn_tc_wgrad_supported(int64_t n_out, int64_t fan_in) {
  return n_out > 0 && fan_in > 0 && n_out % 64 == 0 && fan_in % 64 == 0 && n_out <= 512 && fan_in <= 768 &&
         (n_out % 128 == 0 || fan_in % 128 == 0);
}

extern "C" size_t hgnn_tc_wgrad_workspace_bytes(int64_t rows, int64_t n_out, int64_t fan_in) {
  (void)n_out; (void)fan_in;
  int64_t tiles = (rows + TILE_M - 1) / TILE_M;
  int splits = hgnn::tc::wgrad_splits(WG_MAX_ROLES, (int)std::max<int64_t>(tiles, 1));
  int splits1 = hgnn::tc::wgrad_splits(1, (int)std::max<int64_t>(tiles, 1));
  // four problems of at most 256 x 128 floats per split, or a single one with all the SMs' splits
  size_t a = (size_t)WG_MAX_ROLES * align_up((size_t)splits * 256 * 128 * 4, 256);
  size_t b = align_up((size_t)splits1 * 256 * 128 * 4, 256);
  return std::max(a, b) + 1024;
}

extern "C" int hgnn_tc_wgrad(const void* delta_img, int64_t n_out, const void* a_img, int64_t fan_in, int64_t rows, float* dW,
                             void* ws, size_t ws_bytes, void* stream) {
  if (!hgnn_tc_wgrad_supported(n_out, fan_in))
    return fail(HGNN_ERR_UNSUPPORTED, "tc_wgrad: need n_out, fan_in multiples of 64 (one of them of 128), n_out <= 512, fan_in <= 768");
  HGNN_REQUIRE(dW != nullptr, "tc_wgrad: dW is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0) {
    HGNN_CUDA_TRY(cudaMemsetAsync(dW, 0, (size_t)n_out * fan_in * 4, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(delta_img && a_img && ws, "tc_wgrad: NULL pointer");
  hgnn::tc::WgradProblem pr[WG_MAX_PROBLEMS];
  int n = build_wgrad_problems((const uint8_t*)delta_img, (int)n_out, (const uint8_t*)a_img, (int)fan_in, dW, pr, WG_MAX_PROBLEMS);
  HGNN_REQUIRE(n > 0, "tc_wgrad: could not tile the problem");
  const int tiles = (int)((rows + TILE_M - 1) / TILE_M);
  uintptr_t base = align_up((uintptr_t)ws, 256);
  size_t avail = ws_bytes - (size_t)(base - (uintptr_t)ws);
  for (int i = 0; i < n; i += WG_MAX_ROLES) {
    const int cnt = std::min(WG_MAX_ROLES, n - i);
    int rc = hgnn::tc::launch_wgrad(pr + i, cnt, tiles, (void*)base, avail, st);  // stream order serialises workspace reuse
    if (rc) return rc;
  }
  return HGNN_OK;
}
