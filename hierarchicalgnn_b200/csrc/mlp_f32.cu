// Fused gathered-concat MLP, fp32 SIMT ("exact" path).
//
// One row tile per CTA iteration: the concatenated input row [x[src] | x[dst] | e]
// (or [x | agg | ...]) is gathered straight into shared memory, every layer
// (Linear -> LayerNorm -> activation) runs tile-resident, and only the final
// rows (+ skip) go back to HBM. Nothing of the [rows, 3L] concat or the hidden
// activations the reference materialises (gnn_utils.py:61-62) ever touches HBM
// in the forward pass.
//
// Backward recomputes the forward per tile (replacing torch.utils.checkpoint,
// gnn_utils.py:14-15), emits per-row input gradients and the per-layer delta /
// activation rows that the split-K weight-gradient kernel consumes.
// All reductions are ordered => bit-reproducible run to run.
#include <algorithm>

#include "common.cuh"

using namespace hgnn;

namespace {

constexpr int THREADS = 256;
constexpr int KC = 16;    // k-chunk of the weight stream
constexpr int NB = 256;   // output-column block
constexpr int WLD = NB + 1;

struct Plan {
  int K0;                                  // concatenated input width
  int in_width[HGNN_MLP_MAX_LAYERS];       // K_l
  int seg_off[HGNN_MLP_MAX_SEGS + 1];      // column offset of each segment in the concat
  int ld0;                                 // row stride of the input tile
  int ldw;                                 // row stride of hidden tiles (max out width)
  int vec_off[HGNN_MLP_MAX_LAYERS + 1];    // offsets into the packed (db,dgamma,dbeta) accumulator
  // backward smem offsets (floats)
  int h_off[HGNN_MLP_MAX_LAYERS];
  int h_ld[HGNN_MLP_MAX_LAYERS];
};

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

inline Plan make_plan(const hgnn_mlp_desc& d) {
  Plan p{};
  int off = 0;
  for (int s = 0; s < d.n_seg; ++s) { p.seg_off[s] = off; off += d.seg_width[s]; }
  p.seg_off[d.n_seg] = off;
  p.K0 = off;
  p.ld0 = round4(off);
  int wmax = 4;
  int vo = 0;
  for (int l = 0; l < d.n_layers; ++l) {
    p.in_width[l] = l == 0 ? p.K0 : d.out_width[l - 1];
    if (d.out_width[l] > wmax) wmax = d.out_width[l];
    p.vec_off[l] = vo;
    vo += 3 * d.out_width[l];
  }
  p.vec_off[d.n_layers] = vo;
  p.ldw = round4(wmax);
  return p;
}

inline int validate(const hgnn_mlp_desc* d, int64_t rows) {
  HGNN_REQUIRE(d != nullptr, "mlp: desc is NULL");
  HGNN_REQUIRE(d->n_seg >= 1 && d->n_seg <= HGNN_MLP_MAX_SEGS, "mlp: n_seg must be in [1,%d]", HGNN_MLP_MAX_SEGS);
  HGNN_REQUIRE(d->n_layers >= 1 && d->n_layers <= HGNN_MLP_MAX_LAYERS, "mlp: n_layers must be in [1,%d]", HGNN_MLP_MAX_LAYERS);
  HGNN_REQUIRE(rows >= 0 && rows < INT32_MAX, "mlp: rows out of range");
  int k0 = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    HGNN_REQUIRE(d->seg_ptr[s] && d->seg_width[s] > 0, "mlp: segment %d is empty", s);
    k0 += d->seg_width[s];
  }
  HGNN_REQUIRE(k0 <= 1024, "mlp: concatenated input width %d > 1024 unsupported", k0);
  for (int l = 0; l < d->n_layers; ++l) {
    HGNN_REQUIRE(d->W[l] && d->b[l], "mlp: layer %d has NULL weight/bias", l);
    HGNN_REQUIRE(d->out_width[l] > 0 && d->out_width[l] <= 1024, "mlp: layer %d width %d unsupported", l, d->out_width[l]);
    HGNN_REQUIRE((d->gamma[l] == nullptr) == (d->beta[l] == nullptr), "mlp: layer %d has only one of gamma/beta", l);
    HGNN_REQUIRE(d->act[l] >= HGNN_ACT_NONE && d->act[l] <= HGNN_ACT_SIGMOID, "mlp: layer %d unknown activation %d", l, d->act[l]);
  }
  if (d->skip_seg >= 0) {
    HGNN_REQUIRE(d->skip_seg < d->n_seg, "mlp: skip_seg out of range");
    HGNN_REQUIRE(d->seg_width[d->skip_seg] == d->out_width[d->n_layers - 1], "mlp: skip segment width != output width");
  }
  return HGNN_OK;
}

// ---------------------------------------------------------------------------
// device building blocks
// ---------------------------------------------------------------------------

// gather the concatenated input rows of one tile into shared memory
template <int TM>
__device__ void load_input_tile(const hgnn_mlp_desc& d, const Plan& p, int64_t row0, int64_t rows, float* __restrict__ a0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < TM; r += THREADS / 32) {
    int64_t rho = row0 + r;
    float* dst = a0 + (size_t)r * p.ld0;
    if (rho < rows) {
      for (int s = 0; s < d.n_seg; ++s) {
        const int w = d.seg_width[s];
        int64_t srow = d.seg_idx[s] ? (int64_t)d.seg_idx[s][rho] : rho;
        const float* src = d.seg_ptr[s] + srow * w;
        float* o = dst + p.seg_off[s];
        if (((w | p.seg_off[s]) & 3) == 0 && (((uintptr_t)src) & 15) == 0) {
          for (int c = lane * 4; c < w; c += 128) *reinterpret_cast<float4*>(o + c) = *reinterpret_cast<const float4*>(src + c);
        } else {
          for (int c = lane; c < w; c += 32) o[c] = src[c];
        }
      }
      for (int c = p.K0 + lane; c < p.ld0; c += 32) dst[c] = 0.f;
    } else {
      for (int c = lane; c < p.ld0; c += 32) dst[c] = 0.f;
    }
  }
}

// out[r][c] (+)= sum_k in[r][k] * B(k,c) (+ bias[c]);  TRANS: B(k,c) = W[c*ldW + k] (forward, nn.Linear layout)
//                                                      else:  B(k,c) = W[k*ldW + c] (data gradient)
// Tile rows TM in {8,16,32}; 8 warps: TM/4 row groups x NCH column interleaves.
template <int TM, bool TRANS>
__device__ void tile_gemm(const float* __restrict__ in, int ld_in, int K, const float* __restrict__ W, int ldW,
                          const float* __restrict__ bias, int N, float* __restrict__ out, int ld_out,
                          float* __restrict__ wbuf) {
  constexpr int RG = TM / 4;         // row groups of 4 rows
  constexpr int NCH = 8 / RG;        // warps sharing a row group
  constexpr int NJ = NB / (32 * NCH);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = warp % RG, ch = warp / RG;
  const int r0 = rg * 4;
  constexpr int PER_T = KC * NB / THREADS;  // 16 staged floats per thread

  for (int nb = 0; nb < N; nb += NB) {
    float acc[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;

    float stage[PER_T];
    auto fetch = [&](int k0) {
      if (TRANS) {
        const int kk = threadIdx.x % KC, c0 = threadIdx.x / KC;  // 16 consecutive k per column
#pragma unroll
        for (int i = 0; i < PER_T; ++i) {
          int c = nb + c0 + i * (THREADS / KC);
          int k = k0 + kk;
          stage[i] = (c < N && k < K) ? __ldg(W + (size_t)c * ldW + k) : 0.f;
        }
      } else {
        const int c = nb + threadIdx.x;
#pragma unroll
        for (int i = 0; i < PER_T; ++i) {
          int k = k0 + i;
          stage[i] = (c < N && k < K) ? __ldg(W + (size_t)k * ldW + c) : 0.f;
        }
      }
    };
    auto commit = [&]() {
      if (TRANS) {
        const int kk = threadIdx.x % KC, c0 = threadIdx.x / KC;
#pragma unroll
        for (int i = 0; i < PER_T; ++i) wbuf[kk * WLD + c0 + i * (THREADS / KC)] = stage[i];
      } else {
#pragma unroll
        for (int i = 0; i < PER_T; ++i) wbuf[i * WLD + threadIdx.x] = stage[i];
      }
    };

    fetch(0);
    for (int k0 = 0; k0 < K; k0 += KC) {
      __syncthreads();  // previous chunk fully consumed
      commit();
      __syncthreads();
      if (k0 + KC < K) fetch(k0 + KC);  // prefetch next chunk into registers while computing
      const int kmax = min(KC, K - k0);
#pragma unroll 4
      for (int kk = 0; kk < kmax; ++kk) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = in[(size_t)(r0 + i) * ld_in + k0 + kk];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          float w = wbuf[kk * WLD + lane + 32 * (j * NCH + ch)];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(a[i], w, acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      int c = nb + lane + 32 * (j * NCH + ch);
      if (c < N) {
        float bv = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) out[(size_t)(r0 + i) * ld_out + c] = acc[i][j] + bv;
      }
    }
  }
  __syncthreads();
}

// row statistics of h (two-pass, like ATen's LayerNorm): stats[r] = (mean, rstd)
template <int TM>
__device__ void row_stats(const float* __restrict__ h, int ld, int N, float eps, float* __restrict__ stats) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < TM; r += THREADS / 32) {
    const float* x = h + (size_t)r * ld;
    float s = 0.f;
    for (int c = lane; c < N; c += 32) s += x[c];
    float mean = warp_sum(s) / (float)N;
    float v = 0.f;
    for (int c = lane; c < N; c += 32) { float dlt = x[c] - mean; v = fmaf(dlt, dlt, v); }
    float var = warp_sum(v) / (float)N;
    if (lane == 0) { stats[2 * r] = mean; stats[2 * r + 1] = 1.0f / sqrtf(var + eps); }
  }
  __syncthreads();
}

// dst[r][c] = act(LN(h[r][c]))   (dst may alias h)
template <int TM>
__device__ void norm_act(const float* __restrict__ h, int ld, int N, const float* __restrict__ gamma,
                         const float* __restrict__ beta, int act, const float* __restrict__ stats,
                         float* __restrict__ dst, int ld_dst) {
  for (int t = threadIdx.x; t < TM * N; t += THREADS) {
    int r = t / N, c = t - r * N;
    float y = h[(size_t)r * ld + c];
    if (gamma) y = (y - stats[2 * r]) * stats[2 * r + 1] * __ldg(gamma + c) + __ldg(beta + c);
    dst[(size_t)r * ld_dst + c] = act_fwd(act, y);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------
template <int TM>
__global__ void __launch_bounds__(THREADS, 1) k_mlp_forward(hgnn_mlp_desc d, Plan p, int64_t rows, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* a0 = smem;
  float* buf0 = a0 + (size_t)TM * p.ld0;
  float* buf1 = buf0 + (size_t)TM * p.ldw;
  float* stats = buf1 + (size_t)TM * p.ldw;
  float* wbuf = stats + 2 * TM;
  const int n_tiles = (int)((rows + TM - 1) / TM);
  const int last = d.n_layers - 1;
  const int Nout = d.out_width[last];

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * TM;
    __syncthreads();
    load_input_tile<TM>(d, p, row0, rows, a0);
    __syncthreads();
    const float* in = a0;
    int ld_in = p.ld0;
    for (int l = 0; l <= last; ++l) {
      float* o = (l & 1) ? buf1 : buf0;
      tile_gemm<TM, true>(in, ld_in, p.in_width[l], d.W[l], p.in_width[l], d.b[l], d.out_width[l], o, p.ldw, wbuf);
      if (d.gamma[l]) row_stats<TM>(o, p.ldw, d.out_width[l], d.ln_eps, stats);
      if (d.gamma[l] || d.act[l] != HGNN_ACT_NONE)
        norm_act<TM>(o, p.ldw, d.out_width[l], d.gamma[l], d.beta[l], d.act[l], stats, o, p.ldw);
      in = o;
      ld_in = p.ldw;
    }
    const float* skip = d.skip_seg >= 0 ? a0 + p.seg_off[d.skip_seg] : nullptr;
    for (int t = threadIdx.x; t < TM * Nout; t += THREADS) {
      int r = t / Nout, c = t - r * Nout;
      int64_t rho = row0 + r;
      if (rho < rows) {
        float v = in[(size_t)r * ld_in + c];
        if (skip) v += skip[(size_t)r * p.ld0 + c];
        int64_t orow = d.out_idx ? (int64_t)d.out_idx[rho] : rho;
        out[orow * Nout + c] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// backward-data kernel (recompute + deltas + input gradients)
// ---------------------------------------------------------------------------
struct BwdPtrs {
  float* delta[HGNN_MLP_MAX_LAYERS];  // [rows, out_width[l]] pre-LN linear-output gradients
  float* acts[HGNN_MLP_MAX_LAYERS];   // [rows, out_width[l]] post-activation outputs, l < last
  float* dseg[HGNN_MLP_MAX_SEGS];     // per-row input gradients
  float* colacc;                      // [grid, vec_total] partial (db, dgamma, dbeta)
};

template <int TM>
__global__ void __launch_bounds__(THREADS, 1) k_mlp_backward(hgnn_mlp_desc d, Plan p, int64_t rows, const float* __restrict__ gout,
                                                             BwdPtrs bp) {
  extern __shared__ __align__(16) float smem[];
  float* a0 = smem;                                  // [TM][ld0]  (re-used for d(a0) at the end)
  float* hbase = a0 + (size_t)TM * p.ld0;            // per-layer pre-LN outputs
  int hsz = 0;
  for (int l = 0; l < d.n_layers; ++l) hsz += TM * p.h_ld[l];
  float* scrA = hbase + hsz;                         // [TM][ldw]
  float* scrB = scrA + (size_t)TM * p.ldw;           // [TM][ldw]
  float* stats = scrB + (size_t)TM * p.ldw;          // [layers][TM][2]
  float* rowsum = stats + 2 * TM * HGNN_MLP_MAX_LAYERS;  // [TM][2]
  float* colacc = rowsum + 2 * TM;                   // [vec_total]
  float* wbuf = colacc + p.vec_off[d.n_layers];
  const int n_tiles = (int)((rows + TM - 1) / TM);
  const int last = d.n_layers - 1;
  const int Nout = d.out_width[last];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int t = threadIdx.x; t < p.vec_off[d.n_layers]; t += THREADS) colacc[t] = 0.f;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * TM;
    __syncthreads();
    load_input_tile<TM>(d, p, row0, rows, a0);
    __syncthreads();
    // ---- forward recompute, keeping pre-LN outputs and statistics ----
    {
      const float* in = a0;
      int ld_in = p.ld0;
      for (int l = 0; l <= last; ++l) {
        float* h = hbase + p.h_off[l] * TM;
        tile_gemm<TM, true>(in, ld_in, p.in_width[l], d.W[l], p.in_width[l], d.b[l], d.out_width[l], h, p.h_ld[l], wbuf);
        if (d.gamma[l]) row_stats<TM>(h, p.h_ld[l], d.out_width[l], d.ln_eps, stats + 2 * TM * l);
        if (l < last) {
          norm_act<TM>(h, p.h_ld[l], d.out_width[l], d.gamma[l], d.beta[l], d.act[l], stats + 2 * TM * l, scrA, p.ldw);
          // activation rows feed the weight-gradient pass of layer l+1
          const int N = d.out_width[l];
          for (int t = threadIdx.x; t < TM * N; t += THREADS) {
            int r = t / N, c = t - r * N;
            if (row0 + r < rows) bp.acts[l][(row0 + r) * N + c] = scrA[(size_t)r * p.ldw + c];
          }
          in = scrA;
          ld_in = p.ldw;
          __syncthreads();
        }
      }
    }
    // ---- load the output cotangent ----
    for (int t = threadIdx.x; t < TM * Nout; t += THREADS) {
      int r = t / Nout, c = t - r * Nout;
      int64_t rho = row0 + r;
      float g = 0.f;
      if (rho < rows) {
        int64_t orow = d.out_idx ? (int64_t)d.out_idx[rho] : rho;
        g = gout[orow * Nout + c];
      }
      scrB[(size_t)r * p.ldw + c] = g;
    }
    __syncthreads();
    // ---- backward sweep ----
    for (int l = last; l >= 0; --l) {
      const int N = d.out_width[l];
      const float* h = hbase + p.h_off[l] * TM;
      const int hl = p.h_ld[l];
      const float* st = stats + 2 * TM * l;
      const float* gam = d.gamma[l];
      const float* bet = d.beta[l];
      const int act = d.act[l];
      // pass A (warp per row): delta_y = g * act'(y), row sums for the LayerNorm adjoint
      for (int r = warp; r < TM; r += THREADS / 32) {
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < N; c += 32) {
          float hv = h[(size_t)r * hl + c];
          float xh = 0.f, y = hv;
          if (gam) { xh = (hv - st[2 * r]) * st[2 * r + 1]; y = xh * __ldg(gam + c) + __ldg(bet + c); }
          float dy = scrB[(size_t)r * p.ldw + c] * act_bwd(act, y);
          scrB[(size_t)r * p.ldw + c] = dy;
          if (gam) { float gd = __ldg(gam + c) * dy; s1 += gd; s2 = fmaf(gd, xh, s2); }
        }
        if (gam) {
          s1 = warp_sum(s1); s2 = warp_sum(s2);
          if (lane == 0) { rowsum[2 * r] = s1 / (float)N; rowsum[2 * r + 1] = s2 / (float)N; }
        }
      }
      __syncthreads();
      // pass B (thread per column, rows in order): delta_h, ordered column sums, spill delta_h rows
      for (int c = threadIdx.x; c < N; c += THREADS) {
        float db = 0.f, dg = 0.f, dbt = 0.f;
        float gc = gam ? __ldg(gam + c) : 1.f;
        for (int r = 0; r < TM; ++r) {
          float dy = scrB[(size_t)r * p.ldw + c];
          float dh = dy;
          if (gam) {
            float xh = (h[(size_t)r * hl + c] - st[2 * r]) * st[2 * r + 1];
            dg = fmaf(dy, xh, dg);
            dbt += dy;
            dh = st[2 * r + 1] * (gc * dy - rowsum[2 * r] - xh * rowsum[2 * r + 1]);
          }
          db += dh;
          scrB[(size_t)r * p.ldw + c] = dh;
          if (row0 + r < rows) bp.delta[l][(row0 + r) * N + c] = dh;
        }
        float* acc = colacc + p.vec_off[l];
        acc[c] += db; acc[N + c] += dg; acc[2 * N + c] += dbt;
      }
      __syncthreads();
      // data gradient: d(a_{l-1}) = delta_h W_l
      if (l > 0) {
        tile_gemm<TM, false>(scrB, p.ldw, N, d.W[l], p.in_width[l], nullptr, p.in_width[l], scrA, p.ldw, wbuf);
        float* tmp = scrA; scrA = scrB; scrB = tmp;
      } else {
        bool need = false;
        for (int s = 0; s < d.n_seg; ++s) need |= bp.dseg[s] != nullptr;
        if (need) {
          tile_gemm<TM, false>(scrB, p.ldw, N, d.W[0], p.K0, nullptr, p.K0, a0, p.ld0, wbuf);
          for (int s = 0; s < d.n_seg; ++s) {
            if (!bp.dseg[s]) continue;
            const int w = d.seg_width[s];
            for (int t = threadIdx.x; t < TM * w; t += THREADS) {
              int r = t / w, c = t - r * w;
              int64_t rho = row0 + r;
              if (rho < rows) {
                float v = a0[(size_t)r * p.ld0 + p.seg_off[s] + c];
                if (s == d.skip_seg) {
                  int64_t orow = d.out_idx ? (int64_t)d.out_idx[rho] : rho;
                  v += gout[orow * Nout + c];
                }
                bp.dseg[s][rho * w + c] = v;
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < p.vec_off[d.n_layers]; t += THREADS)
    bp.colacc[(size_t)blockIdx.x * p.vec_off[d.n_layers] + t] = colacc[t];
}

__global__ void k_reduce_partials(const float* __restrict__ part, int n_part, int64_t stride, int64_t n, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int b = 0; b < n_part; ++b) s += part[(size_t)b * stride + i];
  out[i] = s;
}

// ---------------------------------------------------------------------------
// weight gradient: dW[n][k] = sum_r delta[r][n] * a[r][k]   (split-K over rows)
// ---------------------------------------------------------------------------
constexpr int WG_T = 64;   // output tile edge
constexpr int WG_R = 32;   // rows per smem stage

__global__ void __launch_bounds__(THREADS) k_wgrad(hgnn_mlp_desc d, Plan p, int layer, int64_t rows, const float* __restrict__ delta,
                                                    const float* __restrict__ act_in, int64_t rows_per_split,
                                                    float* __restrict__ partial) {
  __shared__ __align__(16) float s_d[WG_R][WG_T + 4];
  __shared__ __align__(16) float s_a[WG_R][WG_T + 4];
  const int N = d.out_width[layer], K = p.in_width[layer];
  const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
  const int64_t r_beg = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(rows, r_beg + rows_per_split);
  const int tn = (threadIdx.x / 16) * 4, tk = (threadIdx.x % 16) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t rb = r_beg; rb < r_end; rb += WG_R) {
    __syncthreads();
    for (int t = threadIdx.x; t < WG_R * WG_T; t += THREADS) {
      int rr = t / WG_T, cc = t % WG_T;
      int64_t rho = rb + rr;
      float dv = 0.f, av = 0.f;
      if (rho < r_end) {
        if (n0 + cc < N) dv = delta[rho * N + n0 + cc];
        int k = k0 + cc;
        if (k < K) {
          if (layer > 0) {
            av = act_in[rho * K + k];
          } else {
            int s = 0;
            while (s + 1 < d.n_seg && k >= p.seg_off[s + 1]) ++s;
            int64_t srow = d.seg_idx[s] ? (int64_t)d.seg_idx[s][rho] : rho;
            av = d.seg_ptr[s][srow * d.seg_width[s] + (k - p.seg_off[s])];
          }
        }
      }
      s_d[rr][cc] = dv;
      s_a[rr][cc] = av;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < WG_R; ++rr) {
      float4 dv = *reinterpret_cast<const float4*>(&s_d[rr][tn]);
      float4 av = *reinterpret_cast<const float4*>(&s_a[rr][tk]);
      float dd[4] = {dv.x, dv.y, dv.z, dv.w}, aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dd[i], aa[j], acc[i][j]);
    }
  }
  float* o = partial + (size_t)blockIdx.z * N * K;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tn + i, k = k0 + tk + j;
      if (n < N && k < K) o[(size_t)n * K + k] = acc[i][j];
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
constexpr size_t SMEM_LIMIT = 227 * 1024;

size_t fwd_smem(const Plan& p, int tm) {
  return sizeof(float) * ((size_t)tm * p.ld0 + 2 * (size_t)tm * p.ldw + 2 * tm + (size_t)KC * WLD);
}

size_t bwd_smem(const hgnn_mlp_desc& d, Plan& p, int tm) {
  size_t hs = 0;
  int off = 0;
  for (int l = 0; l < d.n_layers; ++l) { p.h_ld[l] = round4(d.out_width[l]); p.h_off[l] = off; off += p.h_ld[l]; hs += (size_t)tm * p.h_ld[l]; }
  return sizeof(float) * ((size_t)tm * p.ld0 + hs + 2 * (size_t)tm * p.ldw + 2 * tm * HGNN_MLP_MAX_LAYERS + 2 * tm +
                          p.vec_off[d.n_layers] + (size_t)KC * WLD);
}

struct BwdLayout {
  size_t delta_off[HGNN_MLP_MAX_LAYERS], acts_off[HGNN_MLP_MAX_LAYERS], colacc_off, partial_off, total;
  int grid, tm;
  int splits[HGNN_MLP_MAX_LAYERS];
  int64_t rows_per_split[HGNN_MLP_MAX_LAYERS];
};

int pick_bwd_tm(const hgnn_mlp_desc& d, Plan& p) {
  for (int tm : {32, 16, 8}) if (bwd_smem(d, p, tm) <= SMEM_LIMIT) return tm;
  return 0;
}

BwdLayout bwd_layout(const hgnn_mlp_desc& d, Plan& p, int64_t rows) {
  BwdLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = align_up(off, 256); off = o + bytes; return o; };
  size_t r = (size_t)(rows > 0 ? rows : 1);
  for (int l = 0; l < d.n_layers; ++l) L.delta_off[l] = take(r * d.out_width[l] * 4);
  for (int l = 0; l + 1 < d.n_layers; ++l) L.acts_off[l] = take(r * d.out_width[l] * 4);
  L.tm = pick_bwd_tm(d, p);
  int tm = L.tm ? L.tm : 8;
  int64_t tiles = (rows + tm - 1) / tm;
  L.grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, num_sms()));
  L.colacc_off = take((size_t)L.grid * p.vec_off[d.n_layers] * 4);
  size_t pmax = 0;
  for (int l = 0; l < d.n_layers; ++l) {
    int N = d.out_width[l], K = p.in_width[l];
    int64_t tiles_l = (int64_t)((N + WG_T - 1) / WG_T) * ((K + WG_T - 1) / WG_T);
    int64_t want = std::max<int64_t>(1, (2 * (int64_t)num_sms() + tiles_l - 1) / tiles_l);
    int64_t max_splits = std::max<int64_t>(1, (rows + WG_R * 4 - 1) / (WG_R * 4));
    int64_t s = std::min(want, max_splits);
    int64_t rps = (rows + s - 1) / s;
    rps = (rps + WG_R - 1) / WG_R * WG_R;
    if (rps < WG_R) rps = WG_R;
    s = std::max<int64_t>(1, (rows + rps - 1) / rps);
    L.splits[l] = (int)s;
    L.rows_per_split[l] = rps;
    pmax = std::max(pmax, (size_t)s * N * K * 4);
  }
  L.partial_off = take(pmax);
  L.total = align_up(off, 256);
  return L;
}

template <typename KernelT>
int set_smem(KernelT k, size_t bytes) {
  HGNN_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return HGNN_OK;
}

}  // namespace

extern "C" int hgnn_mlp_forward(const hgnn_mlp_desc* d, int64_t rows, float* out, void* stream) {
  int rc = validate(d, rows);
  if (rc) return rc;
  if (rows == 0) return HGNN_OK;
  HGNN_REQUIRE(out != nullptr, "mlp_forward: out is NULL");
  Plan p = make_plan(*d);
  cudaStream_t st = (cudaStream_t)stream;
  int tm = 0;
  for (int t : {32, 16, 8}) if (fwd_smem(p, t) <= SMEM_LIMIT) { tm = t; break; }
  if (!tm) return fail(HGNN_ERR_UNSUPPORTED, "mlp_forward: layer widths need more than 227 KB of shared memory");
  size_t smem = fwd_smem(p, tm);
  int64_t tiles = (rows + tm - 1) / tm;
  int occ = (int)std::max<size_t>(1, std::min<size_t>(4, SMEM_LIMIT / (smem + 1024)));
  unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)num_sms() * occ);
  if (tm == 32) { rc = set_smem(k_mlp_forward<32>, smem); if (rc) return rc; k_mlp_forward<32><<<grid, THREADS, smem, st>>>(*d, p, rows, out); }
  else if (tm == 16) { rc = set_smem(k_mlp_forward<16>, smem); if (rc) return rc; k_mlp_forward<16><<<grid, THREADS, smem, st>>>(*d, p, rows, out); }
  else { rc = set_smem(k_mlp_forward<8>, smem); if (rc) return rc; k_mlp_forward<8><<<grid, THREADS, smem, st>>>(*d, p, rows, out); }
  return check_launch("mlp_forward");
}

extern "C" size_t hgnn_mlp_backward_workspace_bytes(const hgnn_mlp_desc* d, int64_t rows) {
  if (validate(d, rows) != HGNN_OK) return 0;
  Plan p = make_plan(*d);
  return bwd_layout(*d, p, rows).total;
}

extern "C" int hgnn_mlp_backward_data(const hgnn_mlp_desc* d, int64_t rows, const float* grad_out,
                                      float* const dseg[HGNN_MLP_MAX_SEGS], float* const dvec[HGNN_MLP_MAX_LAYERS],
                                      void* ws, size_t ws_bytes, void* stream) {
  int rc = validate(d, rows);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Plan p = make_plan(*d);
  if (rows == 0) {
    for (int l = 0; l < d->n_layers; ++l)
      if (dvec && dvec[l]) HGNN_CUDA_TRY(cudaMemsetAsync(dvec[l], 0, (size_t)3 * d->out_width[l] * 4, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(grad_out && ws, "mlp_backward_data: NULL grad_out/workspace");
  BwdLayout L = bwd_layout(*d, p, rows);
  if (!L.tm) return fail(HGNN_ERR_UNSUPPORTED, "mlp_backward: layer widths need more than 227 KB of shared memory");
  if (ws_bytes < L.total) return fail(HGNN_ERR_WORKSPACE, "mlp_backward_data: workspace %zu < required %zu", ws_bytes, L.total);
  BwdPtrs bp{};
  char* base = (char*)ws;
  for (int l = 0; l < d->n_layers; ++l) bp.delta[l] = (float*)(base + L.delta_off[l]);
  for (int l = 0; l + 1 < d->n_layers; ++l) bp.acts[l] = (float*)(base + L.acts_off[l]);
  for (int s = 0; s < d->n_seg; ++s) bp.dseg[s] = dseg ? dseg[s] : nullptr;
  bp.colacc = (float*)(base + L.colacc_off);
  size_t smem = bwd_smem(*d, p, L.tm);
  if (L.tm == 32) { rc = set_smem(k_mlp_backward<32>, smem); if (rc) return rc; k_mlp_backward<32><<<L.grid, THREADS, smem, st>>>(*d, p, rows, grad_out, bp); }
  else if (L.tm == 16) { rc = set_smem(k_mlp_backward<16>, smem); if (rc) return rc; k_mlp_backward<16><<<L.grid, THREADS, smem, st>>>(*d, p, rows, grad_out, bp); }
  else { rc = set_smem(k_mlp_backward<8>, smem); if (rc) return rc; k_mlp_backward<8><<<L.grid, THREADS, smem, st>>>(*d, p, rows, grad_out, bp); }
  rc = check_launch("mlp_backward_data");
  if (rc) return rc;
  for (int l = 0; l < d->n_layers; ++l) {
    if (!dvec || !dvec[l]) continue;
    int64_t n = 3 * (int64_t)d->out_width[l];
    // partial layout: [grid][vec_total]; reduce this layer's slice in block order
    k_reduce_partials<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(bp.colacc + p.vec_off[l], L.grid, p.vec_off[d->n_layers], n, dvec[l]);
  }
  return check_launch("mlp_backward_data/reduce");
}

extern "C" int hgnn_mlp_backward_weights(const hgnn_mlp_desc* d, int64_t rows, float* const dW[HGNN_MLP_MAX_LAYERS], void* ws,
                                         size_t ws_bytes, void* stream) {
  int rc = validate(d, rows);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Plan p = make_plan(*d);
  HGNN_REQUIRE(dW != nullptr, "mlp_backward_weights: dW is NULL");
  if (rows == 0) {
    for (int l = 0; l < d->n_layers; ++l)
      if (dW[l]) HGNN_CUDA_TRY(cudaMemsetAsync(dW[l], 0, (size_t)d->out_width[l] * p.in_width[l] * 4, st));
    return HGNN_OK;
  }
  BwdLayout L = bwd_layout(*d, p, rows);
  if (!ws || ws_bytes < L.total) return fail(HGNN_ERR_WORKSPACE, "mlp_backward_weights: workspace %zu < required %zu", ws_bytes, L.total);
  char* base = (char*)ws;
  float* partial = (float*)(base + L.partial_off);
  for (int l = 0; l < d->n_layers; ++l) {
    if (!dW[l]) continue;
    const int N = d->out_width[l], K = p.in_width[l];
    const float* delta = (const float*)(base + L.delta_off[l]);
    const float* act_in = l > 0 ? (const float*)(base + L.acts_off[l - 1]) : nullptr;
    dim3 grid((N + WG_T - 1) / WG_T, (K + WG_T - 1) / WG_T, L.splits[l]);
    k_wgrad<<<grid, THREADS, 0, st>>>(*d, p, l, rows, delta, act_in, L.rows_per_split[l], partial);
    int64_t n = (int64_t)N * K;
    k_reduce_partials<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(partial, L.splits[l], n, n, dW[l]);
  }
  return check_launch("mlp_backward_weights");
}
