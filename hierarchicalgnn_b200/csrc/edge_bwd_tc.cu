// Tensor-core fused edge step, BACKWARD-DATA (sm_100a, latent 128): back-propagates through
// Tanh/LayerNorm/Linear/GELU/LayerNorm/Linear of InteractionGNNCell.edge_update (gnn_utils.py:56-64).
//
// The forward kernel (edge_tc.cu) stashes, per 128-edge tile, what the adjoint needs in bf16 — the normalised
// pre-affine activations xhat1 / xhat2, the row rstd's, and the MMA operand images A0[:, e columns] and g — so nothing is
// recomputed here (the reference recomputes the whole cell under torch.utils.checkpoint, gnn_utils.py:14-15; on a
// 180 GB part 1.5 KB per edge-step is the cheaper side of that trade). Per tile:
//
//   LOAD   gout = grad_eout[i] + grad_agg[dst_i] -> bf16 image in shared memory; xhat2 / rstd rows into registers
//   EPI-B  d(y2) = gout * tanh'(gamma2 xhat2 + beta2), LayerNorm-2 adjoint -> delta2 (bf16 image, in place over gout)
//   GEMM3  dG = delta2 W2           (W2^T image, two N = 128 halves of the hidden width, accumulators in TMEM)
//   EPI-C  d(y1) = dG * gelu'(gamma1 xhat1 + beta1) (parked in TMEM), LayerNorm-1 adjoint -> delta1 (bf16 image)
//   GEMM4  d(e)_mlp = delta1 W1c    (the edge-latent columns of W1 only: four 16 KB W1c^T pieces through the slots)
//   EPI-D  d(e) = d(e)_mlp + gout, as coalesced full rows
//
// The node part of the first layer's adjoint is NOT done per edge. W1 acts linearly on [x[src] | x[dst] | e], so
//   d(x)[n]  = (sum_{src_i = n} delta1_i) W1a + (sum_{dst_i = n} delta1_i) W1b,   dW1a = R_src^T X,  dW1b = R_dst^T X
// with R_src / R_dst the per-node sums of delta1: after this kernel a segmented reduce over the delta1 image builds
// R = [R_src | R_dst] ([nodes, 2H]), one node-level GEMM gives d(x) and one node-level weight-gradient GEMM dW1a / dW1b
// (10x fewer rows than edges). Two thirds of GEMM4 / EPI-D, both per-edge d(x) row tensors (1 KB per edge written and
// read back) and two thirds of the per-edge dW1 GEMM disappear.
//
// gout folds the adjoint of the scatter_add that follows the edge step. The delta1 / delta2 images go to HBM with
// bulk copies; with the forward's A0 / g images they are the operands of the weight-gradient kernel (wgrad_tc.cu).
// Bias / LayerNorm-affine gradients are reduced across rows with a register transpose-reduce, summed in fixed order.
// Weight traffic is decoupled from the tile loop: a producer thread (thread 32) refills the six weight slots as the
// MMAs retire and already requests the next tile's first blocks under EPI-D; the MMA thread (thread 0) never waits
// for a retirement.
//
// k_tc_edge_bwd2 (edge_bwd2_tc.cuh) implements this tile program: two 256-thread CTAs per SM, 104 KB of shared memory and
// 256 TMEM columns each (its one-CTA predecessor, 1.27 ms against 1.06 ms per million edges, was removed in round 2).
#include <algorithm>
#include <cstdlib>

#include "tc_common.cuh"

using namespace hgnn;
using namespace hgnn::tc;

namespace {

constexpr int NT = 512;
constexpr int L = 128, H = 256, K1 = 384;
constexpr int NKB2 = H / KBLK, NKBL = L / KBLK;  // 4, 2
constexpr int W2T_BLK = H * ROW_BYTES;       // 32 KB : W2^T image K-block ([H rows] of [H, L]) = two slots
constexpr int W1T_BLK = K1 * ROW_BYTES;      // 48 KB : W1^T image K-block ([3L rows] of [3L, H])
constexpr int SEG_BLK = L * ROW_BYTES;       // 16 KB : one weight slot = one segment's rows inside a W1^T K-block
constexpr int NSLOT = 6;
constexpr int A2_OFF = NSLOT * SEG_BLK;      // delta1 image -> fp32 output staging (64 KB), after the 96 KB of weight slots
constexpr int A2_BYTES = NKB2 * A_BLK_BYTES;
constexpr int GS_OFF = A2_OFF + A2_BYTES;    // bf16 image of the upstream gradient tile -> delta2 image (32 KB)
constexpr int GS_BYTES = NKBL * A_BLK_BYTES;
constexpr int PAR_OFF = GS_OFF + GS_BYTES;
constexpr int PAR_FLOATS = 3 * H + 3 * L;
constexpr int IDS_OFF = PAR_OFF + PAR_FLOATS * 4;
constexpr int RED_OFF = IDS_OFF + 2 * TILE_M * 4;      // [128 rows][4 splits][2]
constexpr int BAR_OFF = RED_OFF + TILE_M * 8 * 4;
constexpr int NBAR = NSLOT + NSLOT + 1;
constexpr int SMEM_BYTES = BAR_OFF + NBAR * 8 + 16;
constexpr int NPIECE = NKB2;                            // 4 W1c^T pieces (K-blocks over the hidden width); piece b lives in slot (b + 4) % 6
constexpr uint32_t TM_DG = 0, TM_DA0 = 0;               // dG / d(y1): hidden unit c at column c; d(e)_mlp re-uses columns [0, L) afterwards

struct BwdArgs {
  hgnn_tc_edge_params P;
  const uint8_t* w1t;   // W1^T image: [3L rows, H cols]
  const uint8_t* w2t;   // W2^T image: [H rows, L cols]
  const int32_t* dst; const int32_t* perm;  // perm: tile row j -> edge id (NULL = identity)
  const float* g_e; const float* g_agg;   // upstream: d/d e_out [E, L], d/d agg [N, L] (may be NULL)
  float* d_e;                              // [E, L]
  const uint4* xh1; const uint4* xh2; const float* rstd;  // forward stash (EdgeStash in tc_common.cuh)
  uint8_t* d1_img; uint8_t* d2_img;
  float* colpart;                          // [grid][4][PAR_FLOATS]
  int64_t n_edges;
  int stagger_cycles;             // start offset between the four CTA groups (0 = none)
  unsigned long long* phase_clk;  // optional [16] per-phase cycle accumulators (CTA 0, thread 0); NULL in production
};

// 8 bf16 (one stashed uint4) -> fp32
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

#include "edge_bwd2_tc.cuh"

// out[i] = sum_p part[p][i] for i < width, in a fixed order: a block owns 32 columns, its 8 warps sum 8 contiguous ranges of
// the partial vectors in parallel, the 8 range sums are combined in range order. Columns >= split go to out1[i - split].
__global__ void __launch_bounds__(256) k_ordered_colsum(const float* __restrict__ part, int n_part, int width, float* __restrict__ out0,
                                                        int split, float* __restrict__ out1) {
  __shared__ float s_sum[8][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int per = (n_part + 7) / 8, p0 = min(n_part, grp * per), p1 = min(n_part, p0 + per);
  float s = 0.f;
  if (i < width) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // eight loads in flight; fixed association
    int p = p0;
    for (; p + 8 <= p1; p += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += part[(size_t)(p + k) * width + i];
    }
    for (; p < p1; ++p) a[0] += part[(size_t)p * width + i];
    s = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
  s_sum[grp][lane] = s;
  __syncthreads();
  if (grp == 0 && i < width) {
    float t = s_sum[0][lane];
#pragma unroll
    for (int g = 1; g < 8; ++g) t += s_sum[g][lane];
    if (i < split) out0[i] = t; else out1[i - split] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// R[side][n, c] = sum over the tile rows p of segment n (side 0: rows with src = n, side 1: rows with dst = n) of
// delta1[p, c], read straight from the bf16 delta1 tile image the kernel above left in HBM, summed in fp32 in plan order
// (ordered: bit-reproducible). One warp per segment, lane = one 16-byte chunk (8 of the 256 hidden columns): a row is
// four full 128 B lines. Hub segments (> IMG_LONG rows) are left to k_img_segment_reduce_long.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int IMG_LONG = 1024;
struct ImgPlan { const int32_t* rows; const int32_t* rowptr; };  // rows == NULL: tile rows are already segment-sorted

// rows [b0, b1) of one plan, summed in order into acc (whole warp; lane = chunk): loads go out in batches of 8 (the last
// one predicated, so a tail of a few rows costs one round trip, not one per row), the adds follow in row order
__device__ __forceinline__ void img_rows_sum(float (&acc)[8], const uint8_t* __restrict__ img, const int32_t* __restrict__ rows,
                                             int b0, int b1, int lane) {
  const int kb = lane >> 3, c16 = lane & 7;
  for (int j0 = b0; j0 < b1; j0 += 32) {
    const int n = min(32, b1 - j0);
    const int mine = (lane < n) ? (rows ? __ldg(rows + j0 + lane) : j0 + lane) : 0;  // one coalesced look at the row list
#pragma unroll 1
    for (int u = 0; u < n; u += 8) {
      uint4 q[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int p = __shfl_sync(0xffffffffu, mine, (u + k) & 31);
        q[k] = make_uint4(0u, 0u, 0u, 0u);
        if (u + k < n) q[k] = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)(p >> 7) * NKB2 + kb) * A_BLK_BYTES + sw128_off(p & 127, c16)));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float f[8];
        unpack8(q[k], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];  // + 0.0f for the predicated-off tail: exact
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_img_segment_reduce(const uint8_t* __restrict__ img, ImgPlan ps, ImgPlan pd, int64_t n_seg,
                                                            float* __restrict__ R) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * 8 + warp;
  if (s >= n_seg) return;
  const ImgPlan P = blockIdx.y == 0 ? ps : pd;
  const int beg = P.rowptr[s], end = P.rowptr[s + 1];
  if (end - beg > IMG_LONG) return;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  img_rows_sum(acc, img, P.rows, beg, end, lane);
  float4* o = reinterpret_cast<float4*>(R + ((size_t)blockIdx.y * n_seg + s) * H + lane * 8);
  o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// hub segments: a CTA per segment, its 8 warps sum 8 contiguous parts in parallel, combined in part order
__global__ void __launch_bounds__(256) k_img_segment_reduce_long(const uint8_t* __restrict__ img, ImgPlan ps, ImgPlan pd, int64_t n_seg,
                                                                 float* __restrict__ R) {
  __shared__ int s_queue[256];
  __shared__ int s_count;
  __shared__ float s_part[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ImgPlan P = blockIdx.y == 0 ? ps : pd;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  {  // this CTA inspects segment ids {b, b + grid, ...}: neighbouring ids (where hubs cluster) land on different CTAs
    const int64_t s = (int64_t)threadIdx.x * gridDim.x + blockIdx.x;
    if (s < n_seg && P.rowptr[s + 1] - P.rowptr[s] > IMG_LONG) s_queue[atomicAdd(&s_count, 1)] = (int)s;
  }
  __syncthreads();
  const int n_long = s_count;
  for (int qi = 0; qi < n_long; ++qi) {
    const int64_t s = s_queue[qi];
    const int beg = P.rowptr[s], end = P.rowptr[s + 1];
    const int per = (end - beg + 7) / 8;
    const int b0 = min(end, beg + warp * per), b1 = min(end, b0 + per);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    img_rows_sum(acc, img, P.rows, b0, b1, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) s_part[warp][lane * 8 + i] = acc[i];
    __syncthreads();
    {
      const int c = threadIdx.x;
      float t = s_part[0][c];
#pragma unroll
      for (int w = 1; w < 8; ++w) t += s_part[w][c];
      R[((size_t)blockIdx.y * n_seg + s) * H + c] = t;
    }
    __syncthreads();
  }
}

// d bias1 = sum over edges of delta1 = sum over nodes of R_dst (every edge has exactly one destination): two-stage ordered
// column sum of the fp32 [n, H] matrix (used with the two-CTA kernel, which does not accumulate d bias1 itself)
__global__ void __launch_bounds__(H) k_rows_colsum_partial(const float* __restrict__ R, int64_t n_rows, float* __restrict__ part) {
  const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = min(n_rows, (int64_t)blockIdx.x * per), r1 = min(n_rows, r0 + per);
  const int c = threadIdx.x;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int64_t r = r0;
  for (; r + 4 <= r1; r += 4) {
    s0 += R[r * H + c]; s1 += R[(r + 1) * H + c]; s2 += R[(r + 2) * H + c]; s3 += R[(r + 3) * H + c];
  }
  for (; r < r1; ++r) s0 += R[r * H + c];
  part[(size_t)blockIdx.x * H + c] = (s0 + s1) + (s2 + s3);
}
struct Layout {
  size_t d1, d2, colpart, wgrad, wgrad_node, r, r_img, x_img, rsum, total;
  int rsum_parts;
  int grid, tiles, node_tiles;
  size_t wgrad_bytes;
};

Layout make_layout(int64_t n_edges, int64_t n_nodes) {
  Layout Y{};
  Y.tiles = (int)((n_edges + TILE_M - 1) / TILE_M);
  Y.node_tiles = (int)((n_nodes + TILE_M - 1) / TILE_M);
  Y.grid = std::max(1, std::min(Y.tiles, 2 * num_sms()));  // two CTAs per SM
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = align_up(off, 1024); off = o + b; return o; };
  Y.d1 = take((size_t)Y.tiles * NKB2 * A_BLK_BYTES);
  Y.d2 = take((size_t)Y.tiles * NKBL * A_BLK_BYTES);
  Y.colpart = take((size_t)Y.grid * 4 * PAR_FLOATS * 4);
  // wgrad partials: two launches of two [256 x 128] problems each (edge level: dW1c, dW2; node level: dW1a, dW1b)
  int splits = std::max(hgnn::tc::wgrad_splits(2, Y.tiles), hgnn::tc::wgrad_splits(2, std::max(1, Y.node_tiles)));
  Y.wgrad_bytes = (size_t)2 * align_up((size_t)splits * 256 * 128 * 4, 256) + 256;
  Y.wgrad = take(Y.wgrad_bytes);
  Y.wgrad_node = take(Y.wgrad_bytes);  // the node-level launch may run concurrently with the edge-level one (aux stream)
  Y.r = take((size_t)n_nodes * 2 * H * 4);                                // R_src [n, H] then R_dst [n, H], fp32
  Y.r_img = take((size_t)Y.node_tiles * (2 * H / KBLK) * A_BLK_BYTES);    // its bf16 tile image (left by the d(x) GEMM)
  Y.x_img = take((size_t)Y.node_tiles * (L / KBLK) * A_BLK_BYTES);        // bf16 tile image of x
  Y.rsum_parts = 2 * num_sms();
  Y.rsum = take((size_t)Y.rsum_parts * H * 4);
  Y.total = align_up(off, 1024);
  return Y;
}

}  // namespace


extern "C" size_t hgnn_tc_edge_backward_workspace_bytes(int64_t n_edges, int64_t n_nodes) {
  return make_layout(n_edges > 0 ? n_edges : 1, n_nodes > 0 ? n_nodes : 1).total + 1024;
}

extern "C" int hgnn_tc_edge_backward(const hgnn_tc_edge_params* p, const void* w1t_packed, const void* w2t_packed,
                                     const void* wx_packed, const void* stash, const float* x, int64_t n_nodes,
                                     const int32_t* dst, const int32_t* perm, const int32_t* src_rows, const int32_t* src_rowptr,
                                     const int32_t* dst_rows, const int32_t* dst_rowptr, int64_t n_edges, const float* grad_eout,
                                     const float* grad_agg, float* d_e, float* d_x, float* dW1, float* dW2, float* dvec1,
                                     float* dvec2, void* ws, size_t ws_bytes, void* stream, void* aux_stream) {
  HGNN_REQUIRE(p != nullptr, "tc_edge_backward: params is NULL");
  HGNN_REQUIRE(p->latent == 128 && p->hidden == 256, "tc_edge_backward: only latent 128 / hidden 256 is built (got %d / %d)",
               p->latent, p->hidden);
  HGNN_REQUIRE(n_nodes > 0 && n_nodes < INT32_MAX, "tc_edge_backward: bad n_nodes");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_edges <= 0) {
    if (dW1) HGNN_CUDA_TRY(cudaMemsetAsync(dW1, 0, (size_t)H * K1 * 4, st));
    if (dW2) HGNN_CUDA_TRY(cudaMemsetAsync(dW2, 0, (size_t)L * H * 4, st));
    if (dvec1) HGNN_CUDA_TRY(cudaMemsetAsync(dvec1, 0, (size_t)3 * H * 4, st));
    if (dvec2) HGNN_CUDA_TRY(cudaMemsetAsync(dvec2, 0, (size_t)3 * L * 4, st));
    if (d_x) HGNN_CUDA_TRY(cudaMemsetAsync(d_x, 0, (size_t)n_nodes * L * 4, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(w1t_packed && w2t_packed && wx_packed && stash && x && dst && src_rows && src_rowptr && dst_rowptr && grad_eout &&
               d_e && d_x && dW1 && dW2 && dvec1 && dvec2 && ws, "tc_edge_backward: NULL pointer");
  HGNN_REQUIRE(n_edges < INT32_MAX, "tc_edge_backward: too many edges");
  HGNN_REQUIRE(p->gamma1 && p->beta1 && p->gamma2 && p->beta2 && p->b1 && p->b2, "tc_edge_backward: NULL parameter pointer");
  Layout Y = make_layout(n_edges, n_nodes);
  uintptr_t base = align_up((uintptr_t)ws, 1024);
  if (ws_bytes < (base - (uintptr_t)ws) + Y.total) return fail(HGNN_ERR_WORKSPACE, "tc_edge_backward: workspace too small");
  uint8_t* w = (uint8_t*)base;
  const EdgeStash SL = edge_stash_layout(n_edges, L);
  const uint8_t* sb = (const uint8_t*)stash;
  BwdArgs A{};
  A.P = *p;
  A.w1t = (const uint8_t*)w1t_packed;
  A.w2t = (const uint8_t*)w2t_packed;
  A.dst = dst; A.perm = perm;
  A.g_e = grad_eout; A.g_agg = grad_agg;
  A.d_e = d_e;
  A.xh1 = (const uint4*)(sb + SL.xh1); A.xh2 = (const uint4*)(sb + SL.xh2); A.rstd = (const float*)(sb + SL.rstd);
  A.d1_img = w + Y.d1; A.d2_img = w + Y.d2;
  A.colpart = (float*)(w + Y.colpart);
  A.n_edges = n_edges;
  A.phase_clk = (unsigned long long*)p->debug_phase_clock;
  {
    static int stagger = -1;
    if (stagger < 0) { const char* e = getenv("HGNN_BWD_STAGGER"); stagger = e ? atoi(e) : 0; }
    A.stagger_cycles = stagger;
  }
  HGNN_REQUIRE(p->act_hidden == HGNN_ACT_GELU && p->act_out == HGNN_ACT_TANH, "tc_edge_backward: only GELU / Tanh is built");
  const int grid = Y.grid;
  {
    auto kern = v2::k_tc_edge_bwd2<HGNN_ACT_GELU, HGNN_ACT_TANH>;
    HGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2::SMEM_BYTES2));
    kern<<<grid, v2::NT2, v2::SMEM_BYTES2, st>>>(A);
  }
  int rc = check_launch("tc_edge_backward");
  if (rc) return rc;
  // dvec1[3, H] = (db1, dgamma1, dbeta1), dvec2[3, L]: ordered sum over the [grid * 4] partial vectors of the kernel
  // Per-edge weight gradients: dW1[:, 2L:3L] = delta1^T A0[:, e columns], dW2 = delta2^T g. They depend only on the kernel above
  // and stream 1.5 KB per edge at the HBM roof, while the node-level chain below is a string of short latency-bound launches:
  // with an auxiliary stream the two run side by side (fork / join with two events created for this call).
  const uint8_t* a0_img = sb + SL.a0;
  const uint8_t* g_img = sb + SL.g;
  hgnn::tc::WgradProblem pe[2], pn[2];
  pe[0] = hgnn::tc::WgradProblem{A.d1_img, H, 0, H, a0_img, L, 0, L, dW1, K1, 0, 2 * L, 0};
  pe[1] = hgnn::tc::WgradProblem{A.d2_img, L, 0, L, g_img, H, 0, H, dW2, H, 0, 0, 0};
  cudaStream_t aux = (cudaStream_t)aux_stream;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  if (aux != nullptr && aux != st) {
    HGNN_CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    HGNN_CUDA_TRY(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    HGNN_CUDA_TRY(cudaEventRecord(ev_fork, st));
    HGNN_CUDA_TRY(cudaStreamWaitEvent(aux, ev_fork, 0));
    rc = hgnn::tc::launch_wgrad(pe, 2, Y.tiles, w + Y.wgrad, Y.wgrad_bytes, aux);
    HGNN_CUDA_TRY(cudaEventRecord(ev_join, aux));
    if (rc) { cudaStreamWaitEvent(st, ev_join, 0); cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); return rc; }
  }
  auto join = [&]() {  // the caller's stream owns every buffer again once it has waited for the auxiliary stream
    if (ev_join) { cudaStreamWaitEvent(st, ev_join, 0); cudaEventDestroy(ev_fork); cudaEventDestroy(ev_join); ev_join = nullptr; }
  };
  k_ordered_colsum<<<(PAR_FLOATS + 31) / 32, 256, 0, st>>>(A.colpart, grid * 4, PAR_FLOATS, dvec1, 3 * H, dvec2);

  // ---- node level: R = per-node sums of delta1 by source / by destination ----
  float* R = (float*)(w + Y.r);
  const ImgPlan ps{src_rows, src_rowptr}, pd{dst_rows, dst_rowptr};
  k_img_segment_reduce<<<dim3((unsigned)((n_nodes + 7) / 8), 2), 256, 0, st>>>(A.d1_img, ps, pd, n_nodes, R);
  k_img_segment_reduce_long<<<dim3((unsigned)((n_nodes + 255) / 256), 2), 256, 0, st>>>(A.d1_img, ps, pd, n_nodes, R);
  {  // d bias1 = column sums of R_dst (written after k_colpart_reduce left zeros there)
    k_rows_colsum_partial<<<Y.rsum_parts, H, 0, st>>>(R + (size_t)n_nodes * H, n_nodes, (float*)(w + Y.rsum));
    k_ordered_colsum<<<H / 32, 256, 0, st>>>((const float*)(w + Y.rsum), Y.rsum_parts, H, dvec1, H, nullptr);
  }
  rc = check_launch("tc_edge_backward (delta1 node sums)");
  if (rc) { join(); return rc; }
  // d(x) = [R_src | R_dst] . [W1a ; W1b]  (wx_packed = image of [W1a^T | W1b^T] as an [L, 2H] Linear weight); the GEMM
  // leaves the bf16 tile image of [R_src | R_dst] behind for the weight-gradient GEMM
  hgnn_tc_row_layer g{};
  g.n_seg = 2; g.n_out = L; g.act = HGNN_ACT_NONE;
  g.seg_ptr[0] = R; g.seg_ptr[1] = R + (size_t)n_nodes * H;
  g.seg_width[0] = H; g.seg_width[1] = H;
  g.w_packed = wx_packed;
  rc = hgnn_tc_gemm(&g, n_nodes, d_x, L, 0, w + Y.r_img, st);
  if (rc) { join(); return rc; }
  rc = hgnn::tc::launch_make_image(x, n_nodes, L, w + Y.x_img, st);
  if (rc) { join(); return rc; }
  // node-level weight gradients: dW1[:, 0:L] = R_src^T X, dW1[:, L:2L] = R_dst^T X
  for (int sd = 0; sd < 2; ++sd)
    pn[sd] = hgnn::tc::WgradProblem{w + Y.r_img, 2 * H, sd * H, H, w + Y.x_img, L, 0, L, dW1, K1, 0, sd * L, 0};
  rc = hgnn::tc::launch_wgrad(pn, 2, Y.node_tiles, w + Y.wgrad_node, Y.wgrad_bytes, st);
  if (rc) { join(); return rc; }
  if (ev_join) join();
  else rc = hgnn::tc::launch_wgrad(pe, 2, Y.tiles, w + Y.wgrad, Y.wgrad_bytes, st);  // no auxiliary stream: in line
  return rc;
}
