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
// Two kernels implement this tile program. k_tc_edge_bwd2 (edge_bwd2_tc.cuh) is the one hgnn_tc_edge_backward launches: two
// 256-thread CTAs per SM, 104 KB of shared memory and 256 TMEM columns each. k_tc_edge_bwd below is its predecessor — one
// 512-thread CTA per SM, ~200 KB of shared memory, six weight slots — kept behind HGNN_BWD_V2=0 as the A/B baseline
// (tests/test_gpu_tc.py::test_two_cta_backward_kernel_agrees_with_the_one_cta_kernel).
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

template <int ACT_H, int ACT_O>
__global__ void __launch_bounds__(NT, 1) k_tc_edge_bwd(BwdArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* const sm = smem_raw;
  float* s_par = reinterpret_cast<float*>(sm + PAR_OFF);
  float *s_b1 = s_par, *s_g1 = s_par + H, *s_be1 = s_par + 2 * H, *s_b2 = s_par + 3 * H, *s_g2 = s_par + 3 * H + L,
        *s_be2 = s_par + 3 * H + 2 * L;
  int* s_eid = reinterpret_cast<int*>(sm + IDS_OFF);
  int* s_dst = s_eid + TILE_M;
  float* s_red = reinterpret_cast<float*>(sm + RED_OFF);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + BAR_OFF + NBAR * 8);
  const uint32_t sm_u = smem_u32(sm), bar0 = sm_u + BAR_OFF;
  if ((sm_u & 1023u) != 0) __trap();
  enum { B_FULL = 0, B_FREE = NSLOT, ACC = 2 * NSLOT };
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cs = warp >> 2;  // TMEM lane quarter, column split
  const int row = q * 32 + lane;
  const hgnn_tc_edge_params& P = A.P;

  if (tid == 0) {
    for (int i = 0; i < NBAR; ++i) mbar_init(BAR(i), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
  for (int i = tid; i < H; i += NT) { s_b1[i] = P.b1[i]; s_g1[i] = P.gamma1[i]; s_be1[i] = P.beta1[i]; }
  for (int i = tid; i < L; i += NT) { s_b2[i] = P.b2[i]; s_g2[i] = P.gamma2[i]; s_be2[i] = P.beta2[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
  const uint32_t idesc_l = make_idesc(TILE_M, L);

  uint32_t acc_par = 0;
  int nx_eid = 0, nx_dst = 0;  // row ids of the next tile, prefetched
  // mbarrier phase bookkeeping. Every completion of a barrier is awaited exactly once, in order, by one thread:
  //   thread 0  consumes B_FULL[slot]  (nf = completions consumed so far -> next wait parity nf & 1)
  //   thread 32 consumes B_FREE[slot]  (nr likewise)
  uint32_t nf[NSLOT] = {0, 0, 0, 0, 0, 0}, nr[NSLOT] = {0, 0, 0, 0, 0, 0};
  // per-lane column-sum accumulators: lane c of warp (q, cs) owns columns cs*64 + {c, 32 + c} of H and cs*32 + c of L
  float acc_db1[2] = {0.f, 0.f}, acc_dg1[2] = {0.f, 0.f}, acc_dbe1[2] = {0.f, 0.f};
  float acc_db2 = 0.f, acc_dg2 = 0.f, acc_dbe2 = 0.f;

  auto full_wait = [&](int slot) { mbar_wait(BAR(B_FULL + slot), nf[slot] & 1); nf[slot]++; tc_fence_after(); };  // thread 0
  auto free_wait = [&](int slot) { mbar_wait(BAR(B_FREE + slot), nr[slot] & 1); nr[slot]++; };                    // thread 32
  auto fill = [&](int slot, const void* src, uint32_t bytes) {                                                    // thread 32
    mbar_expect_tx(BAR(B_FULL + slot), bytes);
    bulk_g2s(sm_u + slot * SEG_BLK, src, bytes, BAR(B_FULL + slot));
  };
  auto piece_src = [&](int kb) {  // piece kb = rows [2L, 3L) (the edge-latent inputs) of K-block kb of the W1^T image
    return A.w1t + (size_t)kb * W1T_BLK + (size_t)2 * SEG_BLK;
  };
  // first blocks of a tile: W2^T K-blocks 0 / 1 (32 KB each, slots 0+1 / 2+3) and W1^T pieces 0 / 1 (slots 4 / 5)
  auto head_fill = [&]() {  // thread 32
    fill(0, A.w2t, W2T_BLK);
    fill(2, A.w2t + W2T_BLK, W2T_BLK);
    fill(4, piece_src(0), SEG_BLK);
    fill(5, piece_src(1), SEG_BLK);
  };

  long long t_prev = clock64();
  auto MARK = [&](int ph) {
    if (A.phase_clk && tid == 0 && blockIdx.x == 0) { long long t = clock64(); atomicAdd(A.phase_clk + ph, (unsigned long long)(t - t_prev)); t_prev = t; }
  };
  const int n_tiles = (int)((A.n_edges + TILE_M - 1) / TILE_M);
  // All CTAs start together and every tile costs the same, so without this the SMs stay in lock-step and their
  // HBM-heavy phases (gradient rows in, gradient rows out) coincide: stagger the start by a fraction of a tile so the
  // memory system sees a steady demand instead of bursts.
  if (A.stagger_cycles > 0 && n_tiles > (int)gridDim.x) {
    const long long t0 = clock64(), wait = (long long)(blockIdx.x % 4) * A.stagger_cycles;
    while (clock64() - t0 < wait) {}
  }
  if (tid == 32 && (int)blockIdx.x < n_tiles) head_fill();

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ================= tile setup: row ids (fetched one tile ahead into registers; see below) =================
    if (tid < TILE_M) {
      if (tile == (int)blockIdx.x) {
        int64_t j = (int64_t)tile * TILE_M + tid;
        if (j >= A.n_edges) j = A.n_edges - 1;
        nx_eid = A.perm ? A.perm[j] : (int)j;
        nx_dst = A.dst[nx_eid];
      }
      s_eid[tid] = nx_eid;
      s_dst[tid] = nx_dst;
    }
    __syncthreads();
    const bool has_next = tile + (int)gridDim.x < n_tiles;
    if (tid < TILE_M && has_next) {  // first link of the dependent chain perm -> src/dst for the next tile
      int64_t j = (int64_t)(tile + gridDim.x) * TILE_M + tid;
      if (j >= A.n_edges) j = A.n_edges - 1;
      nx_eid = A.perm ? A.perm[j] : (int)j;
    }
    MARK(0);
    // ================= LOAD: upstream gradient tile -> bf16 image; stashed xhat2 / rstd of this thread's row ==========
    const float rstd1 = __ldg(A.rstd + (size_t)tile * 2 * TILE_M + row);
    const float rstd2 = __ldg(A.rstd + (size_t)tile * 2 * TILE_M + TILE_M + row);
    uint4 xq2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xq2[j] = __ldg(A.xh2 + ((size_t)tile * (L / 8) + cs * 4 + j) * TILE_M + row);
    {
      const int g_sub = tid & 31, g_rr = tid >> 5;  // 32 threads per row, 16 rows per pass
      float4 gv[TILE_M / 16], ga[TILE_M / 16];
#pragma unroll
      for (int p = 0; p < TILE_M / 16; ++p) {
        const int r = p * 16 + g_rr;
        const bool live = (int64_t)tile * TILE_M + r < A.n_edges;
        gv[p] = live ? __ldg(reinterpret_cast<const float4*>(A.g_e + (size_t)s_eid[r] * L) + g_sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        ga[p] = (live && A.g_agg) ? __ldg(reinterpret_cast<const float4*>(A.g_agg + (size_t)s_dst[r] * L) + g_sub)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int p = 0; p < TILE_M / 16; ++p) {  // zero rows for padding: their delta's vanish
        const int r = p * 16 + g_rr;
        const int c = g_sub * 4;
        *reinterpret_cast<uint2*>(sm + GS_OFF + (c / KBLK) * A_BLK_BYTES + sw128_off(r, (c % KBLK) >> 3) + ((c >> 2) & 1) * 8) =
            make_uint2(pack_bf16(gv[p].x + ga[p].x, gv[p].y + ga[p].y), pack_bf16(gv[p].z + ga[p].z, gv[p].w + ga[p].w));
      }
    }
    __syncthreads();
    if (tid < TILE_M && has_next) nx_dst = A.dst[nx_eid];  // second link, consumed next tile

    MARK(1);
    // ================= EPI-B: d(y2) = gout * act'(y2), LayerNorm-2 adjoint -> delta2 (in place over gout) =============
    {
      const int c0 = cs * 32;
      float v[32], dy[32];
      const float4* g4 = reinterpret_cast<const float4*>(s_g2 + c0);
      const float4* e4 = reinterpret_cast<const float4*>(s_be2 + c0);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        const int c = c0 + g8 * 8;
        const uint4 pk = *reinterpret_cast<const uint4*>(sm + GS_OFF + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3));
        float go[8], xh[8];
        unpack8(pk, go);
        unpack8(xq2[g8], xh);
        const float4 ga = g4[2 * g8], gb = g4[2 * g8 + 1], ea = e4[2 * g8], eb = e4[2 * g8 + 1];
        const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const float ee[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = g8 * 8 + i;
          const float d = go[i] * tc_act_bwd<ACT_O>(fmaf(xh[i], gg[i], ee[i]));
          const float gd = gg[i] * d;
          dy[k] = d;
          v[k] = xh[i];
          s1 += gd;
          s2 = fmaf(gd, xh[i], s2);
        }
      }
      s_red[row * 8 + cs * 2] = s1;
      s_red[row * 8 + cs * 2 + 1] = s2;
      __syncthreads();
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { t1 += s_red[row * 8 + 2 * i]; t2 += s_red[row * 8 + 2 * i + 1]; }
      t1 *= (1.0f / L);
      t2 *= (1.0f / L);
      float tmp[32];
      // d gamma2 += dy * xhat ; d beta2 += dy
#pragma unroll
      for (int i = 0; i < 32; ++i) tmp[i] = dy[i] * v[i];
      acc_dg2 += warp_colsum32(tmp, lane);
#pragma unroll
      for (int i = 0; i < 32; ++i) tmp[i] = dy[i];
      acc_dbe2 += warp_colsum32(tmp, lane);
      // delta2 = rstd (gamma dy - mean(gamma dy) - xhat mean(gamma dy xhat))
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 g = g4[i];
        dy[4 * i] = rstd2 * (g.x * dy[4 * i] - t1 - v[4 * i] * t2);
        dy[4 * i + 1] = rstd2 * (g.y * dy[4 * i + 1] - t1 - v[4 * i + 1] * t2);
        dy[4 * i + 2] = rstd2 * (g.z * dy[4 * i + 2] - t1 - v[4 * i + 2] * t2);
        dy[4 * i + 3] = rstd2 * (g.w * dy[4 * i + 3] - t1 - v[4 * i + 3] * t2);
      }
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {  // each thread overwrites exactly the gout chunks it read above
        const int c = c0 + g8 * 8;
        *reinterpret_cast<uint4*>(sm + GS_OFF + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3)) =
            make_uint4(pack_bf16(dy[g8 * 8], dy[g8 * 8 + 1]), pack_bf16(dy[g8 * 8 + 2], dy[g8 * 8 + 3]),
                       pack_bf16(dy[g8 * 8 + 4], dy[g8 * 8 + 5]), pack_bf16(dy[g8 * 8 + 6], dy[g8 * 8 + 7]));
      }
      acc_db2 += warp_colsum32(dy, lane);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    MARK(2);
    // ================= GEMM3: dG = delta2 W2 (two N = 128 halves) ; delta2 image -> HBM =================
    if (tid == 0) {
      bulk_s2g(A.d2_img + (size_t)tile * GS_BYTES, sm_u + GS_OFF, GS_BYTES);
      bulk_commit();
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < NKBL; ++kb) {
        const int slot = kb * 2;
        full_wait(slot);
        const uint32_t a_s = sm_u + GS_OFF + kb * A_BLK_BYTES, b_s = sm_u + slot * SEG_BLK;
        umma_kblock(tmem + TM_DG, a_s, b_s, idesc_l, kb == 0);                       // hidden units [0, 128)
        umma_kblock(tmem + TM_DG + L, a_s, b_s + L * ROW_BYTES, idesc_l, kb == 0);   // hidden units [128, 256)
        umma_commit(BAR(B_FREE + slot));
      }
      umma_commit(BAR(ACC));
    }
    // stashed xhat1 of this thread's 64 hidden columns: 8 x 8 bf16, in flight under GEMM3
    uint4 xq1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xq1[j] = __ldg(A.xh1 + ((size_t)tile * (H / 8) + cs * 8 + j) * TILE_M + row);
    if (tid == 32) {  // W1c^T pieces 2 / 3 replace W2^T K-block 0 as its MMAs retire
      free_wait(0);
      fill(0, piece_src(2), SEG_BLK);
      fill(1, piece_src(3), SEG_BLK);
    }
    if (warp == 0) mbar_wait(BAR(ACC), acc_par);  // one warp polls the mbarrier; the rest park on the hardware barrier
    __syncthreads();
    acc_par ^= 1;
    tc_fence_after();

    MARK(3);
    // ================= EPI-C: d(y1) = dG * act'(y1), LayerNorm-1 adjoint -> delta1 image =================
    {
      const int c0 = cs * 64;
      const uint32_t t_dg = t_lane + TM_DG + c0;
      float u[32], tmp[32];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        tmem_ld32(t_dg + ch * 32, u);
        const int cb = c0 + ch * 32;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float xh[8];
          unpack8(xq1[ch * 4 + g8], xh);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 g = *reinterpret_cast<const float4*>(s_g1 + cb + g8 * 8 + 4 * h);
            const float4 be = *reinterpret_cast<const float4*>(s_be1 + cb + g8 * 8 + 4 * h);
            const float gg[4] = {g.x, g.y, g.z, g.w}, ee[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int i = g8 * 8 + 4 * h + k;
              const float x = xh[4 * h + k];
              const float d = u[i] * tc_act_bwd<ACT_H>(fmaf(x, gg[k], ee[k]));
              const float gd = gg[k] * d;
              u[i] = d;
              tmp[i] = d * x;
              s1 += gd;
              s2 = fmaf(gd, x, s2);
            }
          }
        }
        tmem_st32(t_dg + ch * 32, u);  // park d(y1) where dG was
        acc_dg1[ch] += warp_colsum32(tmp, lane);
        acc_dbe1[ch] += warp_colsum32(u, lane);
      }
      __syncthreads();  // s_red free (EPI-B readers done)
      s_red[row * 8 + cs * 2] = s1;
      s_red[row * 8 + cs * 2 + 1] = s2;
      __syncthreads();
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { t1 += s_red[row * 8 + 2 * i]; t2 += s_red[row * 8 + 2 * i + 1]; }
      t1 *= (1.0f / H);
      t2 *= (1.0f / H);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        tmem_ld32(t_dg + ch * 32, u);
        const int cb = c0 + ch * 32;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float xh[8];
          unpack8(xq1[ch * 4 + g8], xh);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 g = *reinterpret_cast<const float4*>(s_g1 + cb + g8 * 8 + 4 * h);
            const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int i = g8 * 8 + 4 * h + k;
              u[i] = rstd1 * (gg[k] * u[i] - t1 - xh[4 * h + k] * t2);
            }
          }
          const int c = cb + g8 * 8;
          *reinterpret_cast<uint4*>(sm + A2_OFF + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3)) =
              make_uint4(pack_bf16(u[g8 * 8], u[g8 * 8 + 1]), pack_bf16(u[g8 * 8 + 2], u[g8 * 8 + 3]),
                         pack_bf16(u[g8 * 8 + 4], u[g8 * 8 + 5]), pack_bf16(u[g8 * 8 + 6], u[g8 * 8 + 7]));
        }
        acc_db1[ch] += warp_colsum32(u, lane);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    MARK(4);
    // ================= GEMM4: d(e)_mlp = delta1 W1c, one N = 128 accumulator =================
    if (tid == 0) {
      bulk_s2g(A.d1_img + (size_t)tile * A2_BYTES, sm_u + A2_OFF, A2_BYTES);
      bulk_commit();
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < NPIECE; ++kb) {
        const int slot = (kb + 4) % NSLOT;
        full_wait(slot);
        umma_kblock(tmem + TM_DA0, sm_u + A2_OFF + kb * A_BLK_BYTES, sm_u + slot * SEG_BLK, idesc_l, kb == 0);
        umma_commit(BAR(B_FREE + slot));
      }
      umma_commit(BAR(ACC));
      bulk_wait_read0();  // delta1 / delta2 images have left shared memory before their regions are reused
    }
    float4 skipg[8];  // fp32 upstream gradient of this lane's 8 output chunks (skip path of d(e)), in flight under GEMM4
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = warp * 8 + k;
      const bool live = (int64_t)tile * TILE_M + r < A.n_edges;
      float4 go = live ? __ldg(reinterpret_cast<const float4*>(A.g_e + (size_t)s_eid[r] * L) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (live && A.g_agg) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(A.g_agg + (size_t)s_dst[r] * L) + lane);
        go.x += a.x; go.y += a.y; go.z += a.z; go.w += a.w;
      }
      skipg[k] = go;
    }
    if (warp == 0) mbar_wait(BAR(ACC), acc_par);  // one warp polls the mbarrier; the rest park on the hardware barrier
    __syncthreads();  // also orders thread 0's bulk_wait_read0 before the staging writes below
    acc_par ^= 1;
    tc_fence_after();
    // every MMA has retired: request the next tile's first weight blocks under EPI-D
    if (tid == 32) {  // consume the remaining retirements: W2^T K-block 1 (slot 2), pieces 0 / 1 (slots 4 / 5), 2 / 3 (slots 0 / 1)
      free_wait(2); free_wait(4); free_wait(5); free_wait(0); free_wait(1);
      if (has_next) head_fill();
    }

    MARK(5);
    // ================= EPI-D: rows of d(e) through a swizzled fp32 staging tile =================
    {
      float v[32];
      tmem_ld32(t_lane + TM_DA0 + cs * 32, v);
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        const int c4 = cs * 8 + g4;
        *reinterpret_cast<float4*>(sm + A2_OFF + (size_t)row * (L * 4) + ((c4 ^ (row & 7)) << 4)) =
            make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // 8 rows per warp, one float4 chunk per lane; skip connection: d(e) += gout (fp32)
      const int r = warp * 8 + k, c4 = lane;
      const int64_t j = (int64_t)tile * TILE_M + r;
      if (j < A.n_edges) {
        float4 y = *reinterpret_cast<const float4*>(sm + A2_OFF + (size_t)r * (L * 4) + ((c4 ^ (r & 7)) << 4));
        y.x += skipg[k].x; y.y += skipg[k].y; y.z += skipg[k].z; y.w += skipg[k].w;
        *reinterpret_cast<float4*>(A.d_e + (size_t)s_eid[r] * L + c4 * 4) = y;
      }
    }
    fence_proxy_async();  // staging (generic proxy) precedes the next tile's bulk store from / writes into these bytes
    tc_fence_before();
    __syncthreads();
    MARK(6);
  }

  // ---- ordered hand-off of the column sums: [cta][q][PAR_FLOATS], lane c owns its columns ----
  {
    float* o = A.colpart + ((size_t)blockIdx.x * 4 + q) * PAR_FLOATS;
    // layout mirrors s_par: db1 | dgamma1 | dbeta1 | db2 | dgamma2 | dbeta2
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int c = cs * 64 + ch * 32 + lane;
      o[c] = acc_db1[ch];
      o[H + c] = acc_dg1[ch];
      o[2 * H + c] = acc_dbe1[ch];
    }
    const int c = cs * 32 + lane;
    o[3 * H + c] = acc_db2;
    o[3 * H + L + c] = acc_dg2;
    o[3 * H + 2 * L + c] = acc_dbe2;
  }
  if (tid == 0) bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
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
  Y.grid = std::max(1, std::min(Y.tiles, 2 * num_sms()));  // sized for the two-CTA kernel; the one-CTA kernel uses half
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

static void* g_phase_clk = nullptr;
// debug hook (not part of the stable ABI): device buffer of 16 uint64 that CTA 0 fills with per-phase cycle counts
extern "C" void hgnn_tc_debug_set_phase_clock(void* dev_u64x16) { g_phase_clk = dev_u64x16; }

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
  A.phase_clk = (unsigned long long*)g_phase_clk;
  {
    static int stagger = -1;
    if (stagger < 0) { const char* e = getenv("HGNN_BWD_STAGGER"); stagger = e ? atoi(e) : 0; }
    A.stagger_cycles = stagger;
  }
  HGNN_REQUIRE(p->act_hidden == HGNN_ACT_GELU && p->act_out == HGNN_ACT_TANH, "tc_edge_backward: only GELU / Tanh is built");
  static int two_cta = -1;  // HGNN_BWD_V2=0 selects the one-CTA-per-SM kernel (A/B comparisons)
  if (two_cta < 0) { const char* e = getenv("HGNN_BWD_V2"); two_cta = e ? atoi(e) : 1; }
  int grid = Y.grid;
  if (two_cta) {
    auto kern = v2::k_tc_edge_bwd2<HGNN_ACT_GELU, HGNN_ACT_TANH>;
    HGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2::SMEM_BYTES2));
    kern<<<grid, v2::NT2, v2::SMEM_BYTES2, st>>>(A);
  } else {
    grid = std::max(1, std::min(Y.tiles, num_sms()));
    auto kern = k_tc_edge_bwd<HGNN_ACT_GELU, HGNN_ACT_TANH>;
    HGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    kern<<<grid, NT, SMEM_BYTES, st>>>(A);
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
  if (two_cta) {  // d bias1 = column sums of R_dst (written after k_colpart_reduce left zeros there)
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
