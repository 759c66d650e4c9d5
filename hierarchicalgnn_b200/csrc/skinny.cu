// Skinny make_mlp layers (utils.py:183-196), fp32 SIMT, one warp per row — the two layer shapes that are pure
// bandwidth and have no business in a GEMM kernel:
//   narrow-in :  fan-in <= 8  (the encoders' first Linear on x[N,3] / [x[src] | x[dst]], EC/Models/IN.py:29-33,84-85;
//                BC/Models/HGNN_GMM.py:37-41) -> out[r, :] = act(LayerNorm(W a_r + b)), write-bound
//   narrow-out:  fan-out <= 8 (edge classifier / bipartite score / embedding heads, EC/Models/IN.py:126,
//                BC/Models/HGNN_GMM.py:45-48,342-344) -> out[r, n] = <W[n, :], a_r> + b[n], read-bound
// Backward passes recompute what they need per row, keep the weight / bias / LayerNorm-affine gradient partials in
// registers across the rows a warp owns, and sum them warp by warp, CTA by CTA in a fixed order (bit-reproducible).
#include <algorithm>

#include "common.cuh"
#include "tc_common.cuh"

using namespace hgnn;

namespace {

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int NI_MAX_K = 8;

struct NarrowInArgs {
  const float* seg_ptr[HGNN_MLP_MAX_SEGS];
  const int32_t* seg_idx[HGNN_MLP_MAX_SEGS];
  int seg_width[HGNN_MLP_MAX_SEGS];
  int n_seg, K, N, act;
  float eps;
  const float *W, *bias, *gamma, *beta;  // gamma == NULL: no LayerNorm
  int64_t rows;
};

// lane k < K fetches input column k of row r (through the segment gathers); everyone gets all K values by shuffle
template <int KMAX = NI_MAX_K>
__device__ __forceinline__ void load_row(const NarrowInArgs& A, int64_t r, int lane, int my_seg, int my_col, float (&a)[NI_MAX_K]) {
  float mine = 0.f;
  if (lane < A.K) {
    const int64_t sr = A.seg_idx[my_seg] ? (int64_t)A.seg_idx[my_seg][r] : r;
    mine = __ldg(A.seg_ptr[my_seg] + sr * A.seg_width[my_seg] + my_col);
  }
#pragma unroll
  for (int k = 0; k < NI_MAX_K; ++k) a[k] = k < KMAX ? __shfl_sync(0xffffffffu, mine, k) : 0.f;
}

__device__ __forceinline__ void lane_source(const NarrowInArgs& A, int lane, int& seg, int& col) {
  int c = lane, s = 0;
  while (s + 1 < A.n_seg && c >= A.seg_width[s]) { c -= A.seg_width[s]; ++s; }
  seg = s;
  col = c;
}

// Sums 8 per-lane values across the warp with 9 shuffles instead of 40: each step folds half of the remaining values
// onto the partner lane (log-step transpose-reduce); on return lane l holds the warp total of value (l & 7) in v[0].
__device__ __forceinline__ float warp_sum8(float (&v)[8], int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 8);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
  return v[0];
}

// shared: Wt[K][N] | bias[N] | gamma[N] | beta[N]
template <int NJ>
__device__ __forceinline__ void stage_params(const NarrowInArgs& A, float* sm) {
  const int N = A.N;
  for (int i = threadIdx.x; i < N * A.K; i += SK_THREADS) {
    const int n = i / A.K, k = i % A.K;
    sm[k * N + n] = A.W[i];
  }
  float* sb = sm + NI_MAX_K * N;
  for (int i = threadIdx.x; i < N; i += SK_THREADS) {
    sb[i] = A.bias[i];
    sb[N + i] = A.gamma ? A.gamma[i] : 1.f;
    sb[2 * N + i] = A.beta ? A.beta[i] : 0.f;
  }
}

// KT > 0: compile-time fan-in (3 = hit coordinates, 6 = the two end points of an edge: the encoders' first layers), the
// FMA chains and shuffles follow it instead of the maximum of 8
template <int NJ, int KT>
__global__ void __launch_bounds__(SK_THREADS) k_narrow_in_fwd(NarrowInArgs A, float* __restrict__ out) {
  constexpr int KMAX = KT > 0 ? KT : NI_MAX_K;
  extern __shared__ float sm[];
  const int N = A.N;
  stage_params<NJ>(A, sm);
  __syncthreads();
  const float *sb = sm + NI_MAX_K * N, *sg = sb + N, *sbe = sb + 2 * N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int my_seg, my_col;
  lane_source(A, lane, my_seg, my_col);
  const float invN = 1.0f / N;
  for (int64_t r = (int64_t)blockIdx.x * SK_WARPS + warp; r < A.rows; r += (int64_t)gridDim.x * SK_WARPS) {
    float a[NI_MAX_K];
    load_row<KMAX>(A, r, lane, my_seg, my_col, a);
    float h[NJ];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane + 32 * j;
      float v = sb[c];
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (KT > 0 || k < A.K) v = fmaf(sm[k * N + c], a[k], v);
      h[j] = v;
      sum += v;
    }
    if (A.gamma) {
      const float mean = warp_sum(sum) * invN;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) { const float d = h[j] - mean; sq = fmaf(d, d, sq); }
      const float rstd = rsqrtf(warp_sum(sq) * invN + A.eps);
#pragma unroll
      for (int j = 0; j < NJ; ++j) { const int c = lane + 32 * j; h[j] = fmaf((h[j] - mean) * rstd, sg[c], sbe[c]); }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) out[(size_t)r * N + lane + 32 * j] = act_fwd(A.act, h[j]);
  }
}

// partial layout per CTA: [K + 3][N] = dW^T rows (k-major) | d bias | d gamma | d beta
// DIN: the input gradient is wanted (false for the encoders: hit coordinates carry no gradient — its 2 K FMAs per column go)
template <int NJ, int KT, bool DIN>
__global__ void __launch_bounds__(SK_THREADS, NJ > 8 ? 1 : 2) k_narrow_in_bwd(NarrowInArgs A, const float* __restrict__ gout, float* __restrict__ d_in,
                                                              float* __restrict__ partial) {
  constexpr int KMAX = KT > 0 ? KT : NI_MAX_K;
  extern __shared__ float sm[];
  const int N = A.N;
  stage_params<NJ>(A, sm);
  float* s_acc = sm + (NI_MAX_K + 3) * N;  // [K + 3][N] CTA accumulator
  for (int i = threadIdx.x; i < (NI_MAX_K + 3) * N; i += SK_THREADS) s_acc[i] = 0.f;
  __syncthreads();
  const float *sb = sm + NI_MAX_K * N, *sg = sb + N, *sbe = sb + 2 * N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int my_seg, my_col;
  lane_source(A, lane, my_seg, my_col);
  const float invN = 1.0f / N;
  float aW[NJ][KMAX], ab[NJ], ag[NJ], abe[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    ab[j] = ag[j] = abe[j] = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) aW[j][k] = 0.f;
  }
  for (int64_t r = (int64_t)blockIdx.x * SK_WARPS + warp; r < A.rows; r += (int64_t)gridDim.x * SK_WARPS) {
    float a[NI_MAX_K];
    load_row<KMAX>(A, r, lane, my_seg, my_col, a);
    float h[NJ], go[NJ];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane + 32 * j;
      go[j] = __ldg(gout + (size_t)r * N + c);
      float v = sb[c];
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (KT > 0 || k < A.K) v = fmaf(sm[k * N + c], a[k], v);
      h[j] = v;
      sum += v;
    }
    float delta[NJ];
    if (A.gamma) {
      const float mean = warp_sum(sum) * invN;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) { const float d = h[j] - mean; sq = fmaf(d, d, sq); }
      const float rstd = rsqrtf(warp_sum(sq) * invN + A.eps);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = lane + 32 * j;
        const float xh = (h[j] - mean) * rstd;
        const float d = go[j] * act_bwd(A.act, fmaf(xh, sg[c], sbe[c]));
        ag[j] = fmaf(d, xh, ag[j]);
        abe[j] += d;
        const float gd = sg[c] * d;
        h[j] = xh;
        delta[j] = gd;
        s1 += gd;
        s2 = fmaf(gd, xh, s2);
      }
      s1 = warp_sum(s1) * invN;
      s2 = warp_sum(s2) * invN;
#pragma unroll
      for (int j = 0; j < NJ; ++j) delta[j] = rstd * (delta[j] - s1 - h[j] * s2);
    } else {
#pragma unroll
      for (int j = 0; j < NJ; ++j) delta[j] = go[j] * act_bwd(A.act, h[j]);
    }
    float da[NI_MAX_K];
#pragma unroll
    for (int k = 0; k < NI_MAX_K; ++k) da[k] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane + 32 * j;
      ab[j] += delta[j];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (KT > 0 || k < A.K) {
          aW[j][k] = fmaf(delta[j], a[k], aW[j][k]);
          if (DIN) da[k] = fmaf(delta[j], sm[k * N + c], da[k]);
        }
      }
    }
    if (DIN) {
      // after the transpose-reduce lane l holds column ((l & 4) | (l & 2) | (l & 1)) bit-reversed pairing: value index
      // = 4 * bit2(l) + 2 * bit1(l) + bit0(l) = l & 7
      const float mine = warp_sum8(da, lane);
      if (lane < A.K) d_in[(size_t)r * A.K + lane] = mine;
    }
  }
  // ordered accumulation: warp 0, 1, ... 7 add their registers into the CTA accumulator in turn
  for (int w = 0; w < SK_WARPS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = lane + 32 * j;
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (KT > 0 || k < A.K) s_acc[k * N + c] += aW[j][k];
        s_acc[(NI_MAX_K + 0) * N + c] += ab[j];
        s_acc[(NI_MAX_K + 1) * N + c] += ag[j];
        s_acc[(NI_MAX_K + 2) * N + c] += abe[j];
      }
    }
    __syncthreads();
  }
  float* o = partial + (size_t)blockIdx.x * (NI_MAX_K + 3) * N;
  for (int i = threadIdx.x; i < (NI_MAX_K + 3) * N; i += SK_THREADS) o[i] = s_acc[i];
}

// dW[n, k] = sum_cta partial[cta][k][n];  dvec[3, N] = (d bias, d gamma, d beta)
__global__ void k_narrow_in_reduce(const float* __restrict__ partial, int n_part, int N, int K, float* __restrict__ dW,
                                   float* __restrict__ dvec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (NI_MAX_K + 3) * N) return;
  const int k = i / N, n = i % N;
  if (k < NI_MAX_K && k >= K) return;
  float s = 0.f;
  for (int p = 0; p < n_part; ++p) s += partial[(size_t)p * (NI_MAX_K + 3) * N + i];
  if (k < NI_MAX_K) dW[(size_t)n * K + k] = s;
  else dvec[(size_t)(k - NI_MAX_K) * N + n] = s;
}

// ---------------------------------------------------------------------------------------------------------------------
constexpr int NO_MAX_N = 8;

// out[r, n] = <W[n, :], a[r, :]> + b[n];  KCH = K / 128 float4 chunks per lane
template <int KCH>
__global__ void __launch_bounds__(SK_THREADS) k_narrow_out_fwd(const float* __restrict__ a, int64_t rows, int K, const float* __restrict__ W,
                                                               const float* __restrict__ bias, int n_out, float* __restrict__ out) {
  extern __shared__ float sm[];  // W [n_out][K]
  for (int i = threadIdx.x; i < n_out * K; i += SK_THREADS) sm[i] = W[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * SK_WARPS + warp; r < rows; r += (int64_t)gridDim.x * SK_WARPS) {
    float4 av[KCH];
#pragma unroll
    for (int c = 0; c < KCH; ++c) av[c] = __ldg(reinterpret_cast<const float4*>(a + (size_t)r * K) + lane + 32 * c);
    float acc[NO_MAX_N];
#pragma unroll
    for (int n = 0; n < NO_MAX_N; ++n) {
      acc[n] = 0.f;
      if (n < n_out) {
#pragma unroll
        for (int c = 0; c < KCH; ++c) {
          const float4 w = *reinterpret_cast<const float4*>(sm + (size_t)n * K + (lane + 32 * c) * 4);
          acc[n] = fmaf(av[c].x, w.x, fmaf(av[c].y, w.y, fmaf(av[c].z, w.z, fmaf(av[c].w, w.w, acc[n]))));
        }
        acc[n] = warp_sum(acc[n]);
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < NO_MAX_N; ++n)
        if (n < n_out) out[(size_t)r * n_out + n] = acc[n] + bias[n];
    }
  }
}

// d_a[r, :] = sum_n gout[r, n] W[n, :];  partial per CTA: [n_out][K] dW | [n_out] db
template <int KCH, int NMAX>
__global__ void __launch_bounds__(SK_THREADS) k_narrow_out_bwd(const float* __restrict__ a, int64_t rows, int K, const float* __restrict__ W,
                                                               int n_out, const float* __restrict__ gout, float* __restrict__ d_a,
                                                               float* __restrict__ partial) {
  extern __shared__ float sm[];  // W [n_out][K] | acc [n_out][K] | db [NO_MAX_N]
  float* s_acc = sm + n_out * K;
  for (int i = threadIdx.x; i < n_out * K; i += SK_THREADS) { sm[i] = W[i]; s_acc[i] = 0.f; }
  if (threadIdx.x < NO_MAX_N) s_acc[n_out * K + threadIdx.x] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 aw[NMAX][KCH];
  float ab[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
    ab[n] = 0.f;
#pragma unroll
    for (int c = 0; c < KCH; ++c) aw[n][c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t r = (int64_t)blockIdx.x * SK_WARPS + warp; r < rows; r += (int64_t)gridDim.x * SK_WARPS) {
    float4 av[KCH], da[KCH];
#pragma unroll
    for (int c = 0; c < KCH; ++c) {
      av[c] = __ldg(reinterpret_cast<const float4*>(a + (size_t)r * K) + lane + 32 * c);
      da[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      if (n < n_out) {
        const float g = __ldg(gout + (size_t)r * n_out + n);
        ab[n] += g;
#pragma unroll
        for (int c = 0; c < KCH; ++c) {
          const float4 w = *reinterpret_cast<const float4*>(sm + (size_t)n * K + (lane + 32 * c) * 4);
          da[c].x = fmaf(g, w.x, da[c].x); da[c].y = fmaf(g, w.y, da[c].y);
          da[c].z = fmaf(g, w.z, da[c].z); da[c].w = fmaf(g, w.w, da[c].w);
          aw[n][c].x = fmaf(g, av[c].x, aw[n][c].x); aw[n][c].y = fmaf(g, av[c].y, aw[n][c].y);
          aw[n][c].z = fmaf(g, av[c].z, aw[n][c].z); aw[n][c].w = fmaf(g, av[c].w, aw[n][c].w);
        }
      }
    }
    if (d_a) {
#pragma unroll
      for (int c = 0; c < KCH; ++c) *(reinterpret_cast<float4*>(d_a + (size_t)r * K) + lane + 32 * c) = da[c];
    }
  }
  for (int w = 0; w < SK_WARPS; ++w) {  // ordered accumulation over the warps
    if (warp == w) {
#pragma unroll
      for (int n = 0; n < NMAX; ++n) {
        if (n < n_out) {
#pragma unroll
          for (int c = 0; c < KCH; ++c) {
            float4* p = reinterpret_cast<float4*>(s_acc + (size_t)n * K + (lane + 32 * c) * 4);
            float4 v = *p;
            v.x += aw[n][c].x; v.y += aw[n][c].y; v.z += aw[n][c].z; v.w += aw[n][c].w;
            *p = v;
          }
          if (lane == 0) s_acc[n_out * K + n] += ab[n];
        }
      }
    }
    __syncthreads();
  }
  float* o = partial + (size_t)blockIdx.x * (n_out * K + NO_MAX_N);
  for (int i = threadIdx.x; i < n_out * K + NO_MAX_N; i += SK_THREADS) o[i] = s_acc[i];
}

__global__ void k_narrow_out_reduce(const float* __restrict__ partial, int n_part, int stride, int n_w, int n_out, float* __restrict__ dW,
                                    float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_w + n_out) return;
  float s = 0.f;
  for (int p = 0; p < n_part; ++p) s += partial[(size_t)p * stride + i];
  if (i < n_w) dW[i] = s; else db[i - n_w] = s;
}

// ---------------------------------------------------------------------------------------------------------------------
// Row-wise LayerNorm + activation (+ residual) on an already computed pre-activation h[rows, N] (bias included), and its
// adjoint. The companions of hgnn_tc_gemm for layers too wide to normalise inside the GEMM kernel (fan-out 512 at latent
// 256) or too narrow for it (fan-out 64): one warp per row, lane l owns columns l, l + 32, ...
template <int NJ>
__global__ void __launch_bounds__(SK_THREADS) k_ln_act_fwd(const float* __restrict__ h, int64_t rows, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, int act,
                                                           const float* __restrict__ skip, float* __restrict__ out) {
  constexpr int N = 32 * NJ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float g[NJ], b[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) { g[j] = gamma[lane + 32 * j]; b[j] = beta[lane + 32 * j]; }
  const float invN = 1.0f / N;
  for (int64_t r = (int64_t)blockIdx.x * SK_WARPS + warp; r < rows; r += (int64_t)gridDim.x * SK_WARPS) {
    float v[NJ];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) { v[j] = __ldg(h + (size_t)r * N + lane + 32 * j); sum += v[j]; }
    const float mean = warp_sum(sum) * invN;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) { const float d = v[j] - mean; sq = fmaf(d, d, sq); }
    const float rstd = rsqrtf(warp_sum(sq) * invN + eps);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float y = act_fwd(act, fmaf((v[j] - mean) * rstd, g[j], b[j]));
      if (skip) y += __ldg(skip + (size_t)r * N + lane + 32 * j);
      out[(size_t)r * N + lane + 32 * j] = y;
    }
  }
}

// delta = LayerNorm/activation adjoint of grad_out at h; partial per CTA: [3][N] = sum delta | sum d xhat | sum d
template <int NJ>
__global__ void __launch_bounds__(SK_THREADS) k_ln_act_bwd(const float* __restrict__ h, const float* __restrict__ gout, int64_t rows,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                           int act, float* __restrict__ delta, float* __restrict__ partial) {
  constexpr int N = 32 * NJ;
  extern __shared__ float s_acc[];  // [3][N]
  for (int i = threadIdx.x; i < 3 * N; i += SK_THREADS) s_acc[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float g[NJ], b[NJ], ab[NJ], ag[NJ], abe[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) { g[j] = gamma[lane + 32 * j]; b[j] = beta[lane + 32 * j]; ab[j] = ag[j] = abe[j] = 0.f; }
  const float invN = 1.0f / N;
  for (int64_t r = (int64_t)blockIdx.x * SK_WARPS + warp; r < rows; r += (int64_t)gridDim.x * SK_WARPS) {
    float v[NJ], go[NJ];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      v[j] = __ldg(h + (size_t)r * N + lane + 32 * j);
      go[j] = __ldg(gout + (size_t)r * N + lane + 32 * j);
      sum += v[j];
    }
    const float mean = warp_sum(sum) * invN;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) { const float d = v[j] - mean; sq = fmaf(d, d, sq); }
    const float rstd = rsqrtf(warp_sum(sq) * invN + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float xh = (v[j] - mean) * rstd;
      const float d = go[j] * act_bwd(act, fmaf(xh, g[j], b[j]));
      ag[j] = fmaf(d, xh, ag[j]);
      abe[j] += d;
      const float gd = g[j] * d;
      v[j] = xh;
      go[j] = gd;
      s1 += gd;
      s2 = fmaf(gd, xh, s2);
    }
    s1 = warp_sum(s1) * invN;
    s2 = warp_sum(s2) * invN;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float dl = rstd * (go[j] - s1 - v[j] * s2);
      ab[j] += dl;
      delta[(size_t)r * N + lane + 32 * j] = dl;
    }
  }
  for (int w = 0; w < SK_WARPS; ++w) {  // ordered accumulation over the warps
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c = lane + 32 * j;
        s_acc[c] += ab[j];
        s_acc[N + c] += ag[j];
        s_acc[2 * N + c] += abe[j];
      }
    }
    __syncthreads();
  }
  float* o = partial + (size_t)blockIdx.x * 3 * N;
  for (int i = threadIdx.x; i < 3 * N; i += SK_THREADS) o[i] = s_acc[i];
}

// ---- vectorised variants for widths that are multiples of 128 (the layer shapes of latent 128 / 256): lane l owns the float4
// at columns 4 l + 128 j, so a warp moves 512 contiguous bytes per load / store; the activations are the tensor-core path's
// fast forms (these kernels only run beside tcgen05 GEMMs: bf16 operands, fp32 LayerNorm) selected at compile time.
template <int ACT>
__device__ __forceinline__ float ln_act_f(int act, float y) {
  if constexpr (ACT >= 0) return hgnn::tc::tc_act<ACT>(y);
  else return act_fwd(act, y);
}
template <int ACT>
__device__ __forceinline__ float ln_act_b(int act, float y) {
  if constexpr (ACT >= 0) return hgnn::tc::tc_act_bwd<ACT>(y);
  else return act_bwd(act, y);
}

template <int NV, int ACT>
__global__ void __launch_bounds__(SK_THREADS, 3) k_ln_act_fwd_v(const float* __restrict__ h, int64_t rows, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, int act,
                                                             const float* __restrict__ skip, float* __restrict__ out) {
  constexpr int N = 128 * NV;
  constexpr int RU = 4 / NV;  // rows per warp iteration: every load of RU rows is in flight before the first reduction
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 g[NV], b[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    g[j] = *reinterpret_cast<const float4*>(gamma + 4 * lane + 128 * j);
    b[j] = *reinterpret_cast<const float4*>(beta + 4 * lane + 128 * j);
  }
  const float invN = 1.0f / N;
  for (int64_t r0 = ((int64_t)blockIdx.x * SK_WARPS + warp) * RU; r0 < rows; r0 += (int64_t)gridDim.x * SK_WARPS * RU) {
    float4 v[RU][NV], sk[RU][NV];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int64_t r = min(r0 + u, rows - 1);  // a tail row is recomputed, not stored twice (guard below)
      const float4* hr = reinterpret_cast<const float4*>(h + (size_t)r * N) + lane;
#pragma unroll
      for (int j = 0; j < NV; ++j) v[u][j] = __ldg(hr + 32 * j);
      if (skip) {
        const float4* sr = reinterpret_cast<const float4*>(skip + (size_t)r * N) + lane;
#pragma unroll
        for (int j = 0; j < NV; ++j) sk[u][j] = __ldg(sr + 32 * j);
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) sum += (v[u][j].x + v[u][j].y) + (v[u][j].z + v[u][j].w);
      const float mean = warp_sum(sum) * invN;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float d0 = v[u][j].x - mean, d1 = v[u][j].y - mean, d2 = v[u][j].z - mean, d3 = v[u][j].w - mean;
        sq = fmaf(d0, d0, sq); sq = fmaf(d1, d1, sq); sq = fmaf(d2, d2, sq); sq = fmaf(d3, d3, sq);
      }
      const float rstd = rsqrtf(warp_sum(sq) * invN + eps);
      if (r0 + u >= rows) continue;
      float4* orow = reinterpret_cast<float4*>(out + (size_t)(r0 + u) * N) + lane;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float4 y;
        y.x = ln_act_f<ACT>(act, fmaf((v[u][j].x - mean) * rstd, g[j].x, b[j].x));
        y.y = ln_act_f<ACT>(act, fmaf((v[u][j].y - mean) * rstd, g[j].y, b[j].y));
        y.z = ln_act_f<ACT>(act, fmaf((v[u][j].z - mean) * rstd, g[j].z, b[j].z));
        y.w = ln_act_f<ACT>(act, fmaf((v[u][j].w - mean) * rstd, g[j].w, b[j].w));
        if (skip) { y.x += sk[u][j].x; y.y += sk[u][j].y; y.z += sk[u][j].z; y.w += sk[u][j].w; }
        orow[32 * j] = y;
      }
    }
  }
}

template <int NV, int ACT>
__global__ void __launch_bounds__(SK_THREADS, 2) k_ln_act_bwd_v(const float* __restrict__ h, const float* __restrict__ gout, int64_t rows,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                             int act, float* __restrict__ delta, float* __restrict__ partial) {
  constexpr int N = 128 * NV;
  constexpr int RU = NV == 1 ? 2 : 1;  // rows per warp iteration (loads of both rows in flight together)
  extern __shared__ float s_acc[];  // [3][N]
  for (int i = threadIdx.x; i < 3 * N; i += SK_THREADS) s_acc[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float g[NV][4], b[NV][4], ab[NV][4], ag[NV][4], abe[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float4 gg = *reinterpret_cast<const float4*>(gamma + 4 * lane + 128 * j), bb = *reinterpret_cast<const float4*>(beta + 4 * lane + 128 * j);
    g[j][0] = gg.x; g[j][1] = gg.y; g[j][2] = gg.z; g[j][3] = gg.w;
    b[j][0] = bb.x; b[j][1] = bb.y; b[j][2] = bb.z; b[j][3] = bb.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) ab[j][k] = ag[j][k] = abe[j][k] = 0.f;
  }
  const float invN = 1.0f / N;
  for (int64_t r0 = ((int64_t)blockIdx.x * SK_WARPS + warp) * RU; r0 < rows; r0 += (int64_t)gridDim.x * SK_WARPS * RU) {
    float v[RU][NV][4], go[RU][NV][4];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const bool live = r0 + u < rows;
      const int64_t r = live ? r0 + u : rows - 1;
      const float4* hr = reinterpret_cast<const float4*>(h + (size_t)r * N) + lane;
      const float4* gr = reinterpret_cast<const float4*>(gout + (size_t)r * N) + lane;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 a = __ldg(hr + 32 * j);
        float4 c = __ldg(gr + 32 * j);
        if (!live) c = make_float4(0.f, 0.f, 0.f, 0.f);  // a tail row past the end contributes nothing
        v[u][j][0] = a.x; v[u][j][1] = a.y; v[u][j][2] = a.z; v[u][j][3] = a.w;
        go[u][j][0] = c.x; go[u][j][1] = c.y; go[u][j][2] = c.z; go[u][j][3] = c.w;
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) sum += (v[u][j][0] + v[u][j][1]) + (v[u][j][2] + v[u][j][3]);
      const float mean = warp_sum(sum) * invN;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float d = v[u][j][k] - mean; sq = fmaf(d, d, sq); }
      const float rstd = rsqrtf(warp_sum(sq) * invN + eps);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float xh = (v[u][j][k] - mean) * rstd;
          const float d = go[u][j][k] * ln_act_b<ACT>(act, fmaf(xh, g[j][k], b[j][k]));
          ag[j][k] = fmaf(d, xh, ag[j][k]);
          abe[j][k] += d;
          const float gd = g[j][k] * d;
          v[u][j][k] = xh;
          go[u][j][k] = gd;
          s1 += gd;
          s2 = fmaf(gd, xh, s2);
        }
      s1 = warp_sum(s1) * invN;
      s2 = warp_sum(s2) * invN;
      if (r0 + u >= rows) continue;  // (its addends above were zeros)
      float4* drow = reinterpret_cast<float4*>(delta + (size_t)(r0 + u) * N) + lane;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float dl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { dl[k] = rstd * (go[u][j][k] - s1 - v[u][j][k] * s2); ab[j][k] += dl[k]; }
        drow[32 * j] = make_float4(dl[0], dl[1], dl[2], dl[3]);
      }
    }
  }
  for (int w = 0; w < SK_WARPS; ++w) {  // ordered accumulation over the warps
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = 4 * lane + 128 * j + k;
          s_acc[c] += ab[j][k];
          s_acc[N + c] += ag[j][k];
          s_acc[2 * N + c] += abe[j][k];
        }
    }
    __syncthreads();
  }
  float* o = partial + (size_t)blockIdx.x * 3 * N;
  for (int i = threadIdx.x; i < 3 * N; i += SK_THREADS) o[i] = s_acc[i];
}

// width 64 (the hidden width of latent 32, the latent of latent 64): 16 lanes x float4 per row, two rows per warp
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int ACT>
__global__ void __launch_bounds__(SK_THREADS) k_ln_act_fwd_64(const float* __restrict__ h, int64_t rows, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int act,
                                                              const float* __restrict__ skip, float* __restrict__ out) {
  constexpr int N = 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 4, c4 = lane & 15;
  const float4 g = *reinterpret_cast<const float4*>(gamma + 4 * c4), b = *reinterpret_cast<const float4*>(beta + 4 * c4);
  for (int64_t r0 = ((int64_t)blockIdx.x * SK_WARPS + warp) * 2; r0 < rows; r0 += (int64_t)gridDim.x * SK_WARPS * 2) {
    const int64_t r = r0 + sub;
    const bool live = r < rows;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), sk = v;
    if (live) {
      v = __ldg(reinterpret_cast<const float4*>(h + (size_t)r * N) + c4);
      if (skip) sk = __ldg(reinterpret_cast<const float4*>(skip + (size_t)r * N) + c4);
    }
    const float mean = half_warp_sum((v.x + v.y) + (v.z + v.w)) * (1.0f / N);
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    const float rstd = rsqrtf(half_warp_sum(fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, d3 * d3)))) * (1.0f / N) + eps);
    if (live) {
      float4 y;
      y.x = ln_act_f<ACT>(act, fmaf(d0 * rstd, g.x, b.x)) + sk.x;
      y.y = ln_act_f<ACT>(act, fmaf(d1 * rstd, g.y, b.y)) + sk.y;
      y.z = ln_act_f<ACT>(act, fmaf(d2 * rstd, g.z, b.z)) + sk.z;
      y.w = ln_act_f<ACT>(act, fmaf(d3 * rstd, g.w, b.w)) + sk.w;
      reinterpret_cast<float4*>(out + (size_t)r * N)[c4] = y;
    }
  }
}

template <int ACT>
__global__ void __launch_bounds__(SK_THREADS) k_ln_act_bwd_64(const float* __restrict__ h, const float* __restrict__ gout, int64_t rows,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                              int act, float* __restrict__ delta, float* __restrict__ partial) {
  constexpr int N = 64;
  extern __shared__ float s_acc[];  // [3][N]
  for (int i = threadIdx.x; i < 3 * N; i += SK_THREADS) s_acc[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 4, c4 = lane & 15;
  const float4 g4 = *reinterpret_cast<const float4*>(gamma + 4 * c4), b4 = *reinterpret_cast<const float4*>(beta + 4 * c4);
  const float g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
  float ab[4] = {0.f, 0.f, 0.f, 0.f}, ag[4] = {0.f, 0.f, 0.f, 0.f}, abe[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t r0 = ((int64_t)blockIdx.x * SK_WARPS + warp) * 2; r0 < rows; r0 += (int64_t)gridDim.x * SK_WARPS * 2) {
    const int64_t r = r0 + sub;
    const bool live = r < rows;
    float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), gv = hv;
    if (live) {
      hv = __ldg(reinterpret_cast<const float4*>(h + (size_t)r * N) + c4);
      gv = __ldg(reinterpret_cast<const float4*>(gout + (size_t)r * N) + c4);
    }
    float v[4] = {hv.x, hv.y, hv.z, hv.w}, go[4] = {gv.x, gv.y, gv.z, gv.w};
    const float mean = half_warp_sum((v[0] + v[1]) + (v[2] + v[3])) * (1.0f / N);
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] -= mean; sq = fmaf(v[k], v[k], sq); }
    const float rstd = rsqrtf(half_warp_sum(sq) * (1.0f / N) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = v[k] * rstd;
      const float d = go[k] * ln_act_b<ACT>(act, fmaf(xh, g[k], b[k]));  // padding rows: gout = 0 -> d = 0
      ag[k] = fmaf(d, xh, ag[k]);
      abe[k] += d;
      const float gd = g[k] * d;
      v[k] = xh;
      go[k] = gd;
      s1 += gd;
      s2 = fmaf(gd, xh, s2);
    }
    s1 = half_warp_sum(s1) * (1.0f / N);
    s2 = half_warp_sum(s2) * (1.0f / N);
    float dl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { dl[k] = live ? rstd * (go[k] - s1 - v[k] * s2) : 0.f; ab[k] += dl[k]; }
    if (live) reinterpret_cast<float4*>(delta + (size_t)r * N)[c4] = make_float4(dl[0], dl[1], dl[2], dl[3]);
  }
  for (int w = 0; w < 2 * SK_WARPS; ++w) {  // ordered accumulation over the (warp, row half) pairs
    if (2 * warp + sub == w) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = 4 * c4 + k;
        s_acc[c] += ab[k];
        s_acc[N + c] += ag[k];
        s_acc[2 * N + c] += abe[k];
      }
    }
    __syncthreads();
  }
  float* o = partial + (size_t)blockIdx.x * 3 * N;
  for (int i = threadIdx.x; i < 3 * N; i += SK_THREADS) o[i] = s_acc[i];
}

__global__ void k_partial_reduce(const float* __restrict__ partial, int n_part, int width, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= width) return;
  float s = 0.f;
  for (int p = 0; p < n_part; ++p) s += partial[(size_t)p * width + i];
  out[i] = s;
}

int narrow_in_args(const hgnn_mlp_desc* d, int64_t rows, NarrowInArgs& a, const char* who) {
  HGNN_REQUIRE(d != nullptr, "%s: desc is NULL", who);
  if (!hgnn_narrow_in_supported(d))
    return fail(HGNN_ERR_UNSUPPORTED, "%s: needs one layer, fan-in <= %d, fan-out in {32, 64, 128, 256, 512}", who, NI_MAX_K);
  HGNN_REQUIRE(rows >= 0 && rows < INT32_MAX, "%s: rows out of range", who);
  a = NarrowInArgs{};
  a.n_seg = d->n_seg;
  int k = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    HGNN_REQUIRE(d->seg_ptr[s] != nullptr, "%s: segment %d is NULL", who, s);
    a.seg_ptr[s] = d->seg_ptr[s]; a.seg_idx[s] = d->seg_idx[s]; a.seg_width[s] = d->seg_width[s];
    k += d->seg_width[s];
  }
  a.K = k; a.N = d->out_width[0]; a.act = d->act[0]; a.eps = d->ln_eps;
  a.W = d->W[0]; a.bias = d->b[0]; a.gamma = d->gamma[0]; a.beta = d->beta[0];
  a.rows = rows;
  HGNN_REQUIRE(a.W && a.bias && ((a.gamma == nullptr) == (a.beta == nullptr)), "%s: NULL parameter", who);
  return HGNN_OK;
}

int skinny_grid(int64_t rows, int rows_per_cta) {
  int64_t want = (rows + rows_per_cta - 1) / rows_per_cta;
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, 2 * (int64_t)num_sms()));
}

}  // namespace

extern "C" int hgnn_narrow_in_supported(const hgnn_mlp_desc* d) {
  if (!d || d->n_layers != 1 || d->n_seg < 1 || d->n_seg > HGNN_MLP_MAX_SEGS || d->skip_seg >= 0 || d->out_idx) return 0;
  int k = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    if (d->seg_width[s] <= 0) return 0;
    k += d->seg_width[s];
  }
  const int n = d->out_width[0];
  return k <= NI_MAX_K && (n == 32 || n == 64 || n == 128 || n == 256 || n == 512);
}

extern "C" int hgnn_narrow_in_forward(const hgnn_mlp_desc* d, int64_t rows, float* out, void* stream) {
  NarrowInArgs a;
  int rc = narrow_in_args(d, rows, a, "narrow_in_forward");
  if (rc) return rc;
  if (rows == 0) return HGNN_OK;
  HGNN_REQUIRE(out != nullptr, "narrow_in_forward: out is NULL");
  const size_t smem = (size_t)(NI_MAX_K + 3) * a.N * 4;
  // 40 registers and 11 KB of shared memory per CTA: up to eight CTAs per SM hide the gather -> shuffle -> LayerNorm chain
  // of a row (ncu: two CTAs per SM left the warp slots 25 % occupied)
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows + SK_WARPS * 4 - 1) / (SK_WARPS * 4), 8 * (int64_t)num_sms()));
  cudaStream_t st = (cudaStream_t)stream;
  switch (a.N / 32) {
    case 1: k_narrow_in_fwd<1, 0><<<grid, SK_THREADS, smem, st>>>(a, out); break;
    case 2: k_narrow_in_fwd<2, 0><<<grid, SK_THREADS, smem, st>>>(a, out); break;
    case 4: k_narrow_in_fwd<4, 0><<<grid, SK_THREADS, smem, st>>>(a, out); break;
    case 8:
      if (a.K == 3) k_narrow_in_fwd<8, 3><<<grid, SK_THREADS, smem, st>>>(a, out);
      else if (a.K == 6) k_narrow_in_fwd<8, 6><<<grid, SK_THREADS, smem, st>>>(a, out);
      else k_narrow_in_fwd<8, 0><<<grid, SK_THREADS, smem, st>>>(a, out);
      break;
    default:  // fan-out 512: the encoders of the latent-256 configs (hidden = 512)
      if (a.K == 3) k_narrow_in_fwd<16, 3><<<grid, SK_THREADS, smem, st>>>(a, out);
      else if (a.K == 6) k_narrow_in_fwd<16, 6><<<grid, SK_THREADS, smem, st>>>(a, out);
      else k_narrow_in_fwd<16, 0><<<grid, SK_THREADS, smem, st>>>(a, out);
      break;
  }
  return check_launch("narrow_in_forward");
}

extern "C" size_t hgnn_narrow_in_backward_workspace_bytes(int64_t n_out) {
  return (size_t)2 * num_sms() * (NI_MAX_K + 3) * n_out * 4 + 256;
}

extern "C" int hgnn_narrow_in_backward(const hgnn_mlp_desc* d, int64_t rows, const float* grad_out, float* d_in, float* dW,
                                       float* dvec, void* ws, size_t ws_bytes, void* stream) {
  NarrowInArgs a;
  int rc = narrow_in_args(d, rows, a, "narrow_in_backward");
  if (rc) return rc;
  HGNN_REQUIRE(dW && dvec, "narrow_in_backward: NULL output");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    HGNN_CUDA_TRY(cudaMemsetAsync(dW, 0, (size_t)a.N * a.K * 4, st));
    HGNN_CUDA_TRY(cudaMemsetAsync(dvec, 0, (size_t)3 * a.N * 4, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(grad_out && ws, "narrow_in_backward: NULL pointer");
  const int grid = skinny_grid(rows, SK_WARPS * 8);
  HGNN_REQUIRE(ws_bytes >= (size_t)grid * (NI_MAX_K + 3) * a.N * 4, "narrow_in_backward: workspace too small");
  float* partial = (float*)ws;
  const size_t smem = (size_t)2 * (NI_MAX_K + 3) * a.N * 4;
  switch (a.N / 32) {
    case 1: k_narrow_in_bwd<1, 0, true><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial); break;
    case 2: k_narrow_in_bwd<2, 0, true><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial); break;
    case 4: k_narrow_in_bwd<4, 0, true><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial); break;
    case 16:
      if (d_in != nullptr && a.K == 3) k_narrow_in_bwd<16, 3, true><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      else if (d_in != nullptr && a.K == 6) k_narrow_in_bwd<16, 6, true><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      else if (d_in != nullptr) k_narrow_in_bwd<16, 0, true><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      else if (a.K == 3) k_narrow_in_bwd<16, 3, false><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      else if (a.K == 6) k_narrow_in_bwd<16, 6, false><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      else k_narrow_in_bwd<16, 0, false><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      break;
    default:
      if (d_in != nullptr && a.K == 3) k_narrow_in_bwd<8, 3, true><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      else if (d_in != nullptr && a.K == 6) k_narrow_in_bwd<8, 6, true><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      else if (d_in != nullptr) k_narrow_in_bwd<8, 0, true><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      else if (a.K == 3) k_narrow_in_bwd<8, 3, false><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      else if (a.K == 6) k_narrow_in_bwd<8, 6, false><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      else k_narrow_in_bwd<8, 0, false><<<grid, SK_THREADS, smem, st>>>(a, grad_out, d_in, partial);
      break;
  }
  rc = check_launch("narrow_in_backward");
  if (rc) return rc;
  const int total = (NI_MAX_K + 3) * a.N;
  k_narrow_in_reduce<<<(total + 255) / 256, 256, 0, st>>>(partial, grid, a.N, a.K, dW, dvec);
  if (!a.gamma) {  // no LayerNorm: d gamma / d beta rows are meaningless -> zero
    HGNN_CUDA_TRY(cudaMemsetAsync(dvec + a.N, 0, (size_t)2 * a.N * 4, st));
  }
  return check_launch("narrow_in_backward (reduce)");
}

extern "C" int hgnn_narrow_out_supported(int64_t k, int64_t n_out) {
  return (k == 128 || k == 256 || k == 512) && n_out >= 1 && n_out <= NO_MAX_N && k * n_out <= 2048;
}

extern "C" int hgnn_narrow_out_forward(const float* a, int64_t rows, int64_t k, const float* W, const float* bias, int64_t n_out,
                                       float* out, void* stream) {
  if (!hgnn_narrow_out_supported(k, n_out))
    return fail(HGNN_ERR_UNSUPPORTED, "narrow_out_forward: needs fan-in in {128, 256, 512}, fan-out <= 8, fan-in * fan-out <= 2048");
  if (rows <= 0) return HGNN_OK;
  HGNN_REQUIRE(a && W && bias && out, "narrow_out_forward: NULL pointer");
  const size_t smem = (size_t)n_out * k * 4;
  const int grid = skinny_grid(rows, SK_WARPS * 4);
  cudaStream_t st = (cudaStream_t)stream;
  switch (k / 128) {
    case 1: k_narrow_out_fwd<1><<<grid, SK_THREADS, smem, st>>>(a, rows, (int)k, W, bias, (int)n_out, out); break;
    case 2: k_narrow_out_fwd<2><<<grid, SK_THREADS, smem, st>>>(a, rows, (int)k, W, bias, (int)n_out, out); break;
    default: k_narrow_out_fwd<4><<<grid, SK_THREADS, smem, st>>>(a, rows, (int)k, W, bias, (int)n_out, out); break;
  }
  return check_launch("narrow_out_forward");
}

extern "C" size_t hgnn_narrow_out_backward_workspace_bytes(int64_t k, int64_t n_out) {
  return (size_t)2 * num_sms() * (n_out * k + NO_MAX_N) * 4 + 256;
}

extern "C" int hgnn_narrow_out_backward(const float* a, int64_t rows, int64_t k, const float* W, int64_t n_out, const float* grad_out,
                                        float* d_a, float* dW, float* db, void* ws, size_t ws_bytes, void* stream) {
  if (!hgnn_narrow_out_supported(k, n_out))
    return fail(HGNN_ERR_UNSUPPORTED, "narrow_out_backward: needs fan-in in {128, 256, 512}, fan-out <= 8, fan-in * fan-out <= 2048");
  HGNN_REQUIRE(dW && db, "narrow_out_backward: NULL output");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0) {
    HGNN_CUDA_TRY(cudaMemsetAsync(dW, 0, (size_t)n_out * k * 4, st));
    HGNN_CUDA_TRY(cudaMemsetAsync(db, 0, (size_t)n_out * 4, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(a && W && grad_out && ws, "narrow_out_backward: NULL pointer");
  const int grid = skinny_grid(rows, SK_WARPS * 8);
  const int stride = (int)(n_out * k) + NO_MAX_N;
  HGNN_REQUIRE(ws_bytes >= (size_t)grid * stride * 4, "narrow_out_backward: workspace too small");
  float* partial = (float*)ws;
  const size_t smem = ((size_t)2 * n_out * k + NO_MAX_N) * 4;
  const int K = (int)k, NO = (int)n_out;
  // register accumulators: n_out x K/32 floats per lane (<= 64)
  if (k == 128) {
    k_narrow_out_bwd<1, 8><<<grid, SK_THREADS, smem, st>>>(a, rows, K, W, NO, grad_out, d_a, partial);
  } else if (k == 256) {
    k_narrow_out_bwd<2, 8><<<grid, SK_THREADS, smem, st>>>(a, rows, K, W, NO, grad_out, d_a, partial);
  } else {
    k_narrow_out_bwd<4, 4><<<grid, SK_THREADS, smem, st>>>(a, rows, K, W, NO, grad_out, d_a, partial);
  }
  int rc = check_launch("narrow_out_backward");
  if (rc) return rc;
  const int total = NO * K + NO;
  k_narrow_out_reduce<<<(total + 255) / 256, 256, 0, st>>>(partial, grid, stride, NO * K, NO, dW, db);
  return check_launch("narrow_out_backward (reduce)");
}

extern "C" int hgnn_ln_act_supported(int64_t n) { return n == 64 || n == 128 || n == 256 || n == 512; }

extern "C" int hgnn_ln_act_forward(const float* h, int64_t rows, int64_t n, const float* gamma, const float* beta, float eps, int act,
                                   const float* skip, float* out, void* stream) {
  if (!hgnn_ln_act_supported(n)) return fail(HGNN_ERR_UNSUPPORTED, "ln_act_forward: width must be 64, 128, 256 or 512 (got %lld)", (long long)n);
  if (rows <= 0) return HGNN_OK;
  HGNN_REQUIRE(h && gamma && beta && out, "ln_act_forward: NULL pointer");
  HGNN_REQUIRE(act >= HGNN_ACT_NONE && act <= HGNN_ACT_SIGMOID, "ln_act_forward: unknown activation %d", act);
  // grid-stride kernels without cross-CTA state: three CTAs per SM resident (the vectorised kernels are bounded to 80 registers)
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows + SK_WARPS * 4 - 1) / (SK_WARPS * 4), 6 * (int64_t)num_sms()));
  cudaStream_t st = (cudaStream_t)stream;
  const bool al16 = (((uintptr_t)h | (uintptr_t)out | (uintptr_t)skip | (uintptr_t)gamma | (uintptr_t)beta) % 16) == 0;
  if (n == 64 && al16) {
    if (act == HGNN_ACT_GELU) k_ln_act_fwd_64<HGNN_ACT_GELU><<<grid, SK_THREADS, 0, st>>>(h, rows, gamma, beta, eps, act, skip, out);
    else if (act == HGNN_ACT_TANH) k_ln_act_fwd_64<HGNN_ACT_TANH><<<grid, SK_THREADS, 0, st>>>(h, rows, gamma, beta, eps, act, skip, out);
    else k_ln_act_fwd_64<-1><<<grid, SK_THREADS, 0, st>>>(h, rows, gamma, beta, eps, act, skip, out);
    return check_launch("ln_act_forward");
  }
  const bool vec = n % 128 == 0 && al16;
  if (vec) {
#define HGNN_LN_FWD(NV, ACT) k_ln_act_fwd_v<NV, ACT><<<grid, SK_THREADS, 0, st>>>(h, rows, gamma, beta, eps, act, skip, out)
#define HGNN_LN_FWD_ACT(NV)                                       \
    do {                                                          \
      if (act == HGNN_ACT_GELU) HGNN_LN_FWD(NV, HGNN_ACT_GELU);   \
      else if (act == HGNN_ACT_TANH) HGNN_LN_FWD(NV, HGNN_ACT_TANH); \
      else if (act == HGNN_ACT_NONE) HGNN_LN_FWD(NV, HGNN_ACT_NONE); \
      else HGNN_LN_FWD(NV, -1);                                   \
    } while (0)
    if (n == 128) HGNN_LN_FWD_ACT(1);
    else if (n == 256) HGNN_LN_FWD_ACT(2);
    else HGNN_LN_FWD_ACT(4);
#undef HGNN_LN_FWD_ACT
#undef HGNN_LN_FWD
    return check_launch("ln_act_forward");
  }
  switch (n / 32) {
    case 2: k_ln_act_fwd<2><<<grid, SK_THREADS, 0, st>>>(h, rows, gamma, beta, eps, act, skip, out); break;
    case 4: k_ln_act_fwd<4><<<grid, SK_THREADS, 0, st>>>(h, rows, gamma, beta, eps, act, skip, out); break;
    case 8: k_ln_act_fwd<8><<<grid, SK_THREADS, 0, st>>>(h, rows, gamma, beta, eps, act, skip, out); break;
    default: k_ln_act_fwd<16><<<grid, SK_THREADS, 0, st>>>(h, rows, gamma, beta, eps, act, skip, out); break;
  }
  return check_launch("ln_act_forward");
}

extern "C" size_t hgnn_ln_act_backward_workspace_bytes(int64_t n) { return (size_t)2 * num_sms() * 3 * n * 4 + 256; }

extern "C" int hgnn_ln_act_backward(const float* h, const float* grad_out, int64_t rows, int64_t n, const float* gamma,
                                    const float* beta, float eps, int act, float* delta, float* dvec, void* ws, size_t ws_bytes,
                                    void* stream) {
  if (!hgnn_ln_act_supported(n)) return fail(HGNN_ERR_UNSUPPORTED, "ln_act_backward: width must be 64, 128, 256 or 512 (got %lld)", (long long)n);
  HGNN_REQUIRE(dvec != nullptr, "ln_act_backward: dvec is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0) {
    HGNN_CUDA_TRY(cudaMemsetAsync(dvec, 0, (size_t)3 * n * 4, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(h && grad_out && gamma && beta && delta && ws, "ln_act_backward: NULL pointer");
  HGNN_REQUIRE(act >= HGNN_ACT_NONE && act <= HGNN_ACT_SIGMOID, "ln_act_backward: unknown activation %d", act);
  const int grid = skinny_grid(rows, SK_WARPS * 8);
  HGNN_REQUIRE(ws_bytes >= (size_t)grid * 3 * n * 4, "ln_act_backward: workspace too small");
  float* partial = (float*)ws;
  const size_t smem = (size_t)3 * n * 4;
  const bool al16 = (((uintptr_t)h | (uintptr_t)grad_out | (uintptr_t)delta | (uintptr_t)gamma | (uintptr_t)beta) % 16) == 0;
  const bool vec = n % 128 == 0 && al16;
  if (n == 64 && al16) {
    if (act == HGNN_ACT_GELU) k_ln_act_bwd_64<HGNN_ACT_GELU><<<grid, SK_THREADS, smem, st>>>(h, grad_out, rows, gamma, beta, eps, act, delta, partial);
    else if (act == HGNN_ACT_TANH) k_ln_act_bwd_64<HGNN_ACT_TANH><<<grid, SK_THREADS, smem, st>>>(h, grad_out, rows, gamma, beta, eps, act, delta, partial);
    else k_ln_act_bwd_64<-1><<<grid, SK_THREADS, smem, st>>>(h, grad_out, rows, gamma, beta, eps, act, delta, partial);
  } else
  if (vec) {
#define HGNN_LN_BWD(NV, ACT) k_ln_act_bwd_v<NV, ACT><<<grid, SK_THREADS, smem, st>>>(h, grad_out, rows, gamma, beta, eps, act, delta, partial)
#define HGNN_LN_BWD_ACT(NV)                                       \
    do {                                                          \
      if (act == HGNN_ACT_GELU) HGNN_LN_BWD(NV, HGNN_ACT_GELU);   \
      else if (act == HGNN_ACT_TANH) HGNN_LN_BWD(NV, HGNN_ACT_TANH); \
      else if (act == HGNN_ACT_NONE) HGNN_LN_BWD(NV, HGNN_ACT_NONE); \
      else HGNN_LN_BWD(NV, -1);                                   \
    } while (0)
    if (n == 128) HGNN_LN_BWD_ACT(1);
    else if (n == 256) HGNN_LN_BWD_ACT(2);
    else HGNN_LN_BWD_ACT(4);
#undef HGNN_LN_BWD_ACT
#undef HGNN_LN_BWD
  } else
  switch (n / 32) {
    case 2: k_ln_act_bwd<2><<<grid, SK_THREADS, smem, st>>>(h, grad_out, rows, gamma, beta, eps, act, delta, partial); break;
    case 4: k_ln_act_bwd<4><<<grid, SK_THREADS, smem, st>>>(h, grad_out, rows, gamma, beta, eps, act, delta, partial); break;
    case 8: k_ln_act_bwd<8><<<grid, SK_THREADS, smem, st>>>(h, grad_out, rows, gamma, beta, eps, act, delta, partial); break;
    default: k_ln_act_bwd<16><<<grid, SK_THREADS, smem, st>>>(h, grad_out, rows, gamma, beta, eps, act, delta, partial); break;
  }
  int rc = check_launch("ln_act_backward");
  if (rc) return rc;
  k_partial_reduce<<<(int)((3 * n + 255) / 256), 256, 0, st>>>(partial, grid, (int)(3 * n), dvec);
  return check_launch("ln_act_backward (reduce)");
}
