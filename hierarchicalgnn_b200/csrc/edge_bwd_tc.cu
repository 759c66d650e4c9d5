// Tensor-core fused edge step, BACKWARD-DATA (sm_100a, latent 128):
// recomputes the forward per 128-edge tile (replacing torch.utils.checkpoint,
// gnn_utils.py:14-15) and back-propagates through Tanh/LayerNorm/Linear/GELU/
// LayerNorm/Linear of InteractionGNNCell.edge_update (gnn_utils.py:56-64):
//
//   GEMM1  h1 = [x[src] | x[dst] | e] W1^T          (gather -> bf16 ring -> tcgen05, D1 kept in TMEM)
//   EPI-A  LN1 stats, g = act(LN1(h1 + b1))  -> bf16 image (A operand of GEMM2, bulk-stored for wgrad)
//   GEMM2  h2 = g W2^T
//   EPI-B  LN2/Tanh forward, d(y2) = gout * act'(y2), LN2 adjoint -> delta2 (bf16 image), column sums
//   GEMM3  dG = delta2 W2        (W2^T image as B operand)
//   EPI-C  d(y1) = dG * act'(y1) (parked in TMEM), LN1 adjoint -> delta1 (bf16 image), column sums
//   GEMM4  dA0 = delta1 W1       (W1^T image streamed in 16 KB (segment, K-block) pieces)
//   EPI-D  per-edge rows d(x[src]), d(x[dst]) and d(e) = dA0_e + gout, as coalesced full rows
//
// gout = grad_eout[i] + grad_agg[dst_i] folds the adjoint of the scatter_add that follows the
// edge step. The bf16 tile images of A0, g, delta1, delta2 go to HBM with bulk copies and are the
// operands of the weight-gradient kernel (wgrad_tc.cu). Bias / LayerNorm-affine gradients are
// reduced across rows with a register transpose-reduce and summed in a fixed order.
// One CTA per SM (all 512 TMEM columns, ~205 KB of shared memory), 16 warps.
#include <algorithm>
#include <cstdlib>

#include "tc_common.cuh"

using namespace hgnn;
using namespace hgnn::tc;

namespace {

constexpr int NT = 512;
constexpr int L = 128, H = 256, K1 = 384;
constexpr int NKB1 = K1 / KBLK, NKB2 = H / KBLK, NKBL = L / KBLK;  // 6, 4, 2
constexpr int W1_BLK = H * ROW_BYTES;        // 32 KB : W1 image K-block ([H rows] of the [H, 3L] matrix)
constexpr int W2_BLK = L * ROW_BYTES;        // 16 KB : W2 image K-block ([L rows] of [L, H])
constexpr int W2T_BLK = H * ROW_BYTES;       // 32 KB : W2^T image K-block ([H rows] of [H, L])
constexpr int W1T_BLK = K1 * ROW_BYTES;      // 48 KB : W1^T image K-block ([3L rows] of [3L, H])
constexpr int SEG_BLK = L * ROW_BYTES;       // 16 KB : one segment's rows inside a W1^T K-block
constexpr int STAGE = A_BLK_BYTES + W1_BLK;  // 48 KB
constexpr int RING = 2 * STAGE;              // 96 KB
constexpr int D2IMG_OFF = NKB2 * W2_BLK;     // delta2 image sits after the W2 / W2^T area: 64 KB
constexpr int A2_OFF = RING;                 // g image -> delta1 image -> fp32 output staging (64 KB)
constexpr int A2_BYTES = NKB2 * A_BLK_BYTES;
constexpr int GS_OFF = A2_OFF + A2_BYTES;    // bf16 image of the upstream gradient tile (32 KB)
constexpr int GS_BYTES = NKBL * A_BLK_BYTES;
constexpr int PAR_OFF = GS_OFF + GS_BYTES;
constexpr int PAR_FLOATS = 3 * H + 3 * L;
constexpr int IDS_OFF = PAR_OFF + PAR_FLOATS * 4;
constexpr int RED_OFF = IDS_OFF + 3 * TILE_M * 4;      // [128 rows][4 splits][2]
constexpr int BAR_OFF = RED_OFF + TILE_M * 8 * 4;
constexpr int NBAR = 2 + 2 + 6 + 6 + 1 + 1 + 1 + 1;
constexpr int SMEM_BYTES = BAR_OFF + NBAR * 8 + 16;
constexpr int NSLOT = 6;                                // 16 KB slots over the ring for GEMM4's weight stream
constexpr uint32_t TM_D1 = 0, TM_D2 = H, TM_DGHI = H, TM_DGLO = H + L, TM_DA0 = 0;

struct BwdArgs {
  hgnn_tc_edge_params P;
  const uint8_t* w1t;   // W1^T image: [3L rows, H cols]
  const uint8_t* w2t;   // W2^T image: [H rows, L cols]
  const int32_t* src; const int32_t* dst; const int32_t* perm;  // perm: tile row j -> edge id (NULL = identity)
  const float* g_e; const float* g_agg;   // upstream: d/d e_out [E, L], d/d agg [N, L] (may be NULL)
  float* d_e; float* d_xs; float* d_xd;   // [E, L] each
  const uint8_t* a0_img;  // [tiles][6][16 KB] bf16 image of [x[src] | x[dst] | e], written by the forward kernel in the same row order
  uint8_t* g_img; uint8_t* d1_img; uint8_t* d2_img;
  float* colpart;                          // [grid][4][PAR_FLOATS]
  int64_t n_edges;
  int stagger_cycles;             // start offset between the four CTA groups (0 = none)
  unsigned long long* phase_clk;  // optional [16] per-phase cycle accumulators (CTA 0, thread 0); NULL in production
};

template <int ACT_H, int ACT_O>
__global__ void __launch_bounds__(NT, 1) k_tc_edge_bwd(BwdArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* const sm = smem_raw;
  float* s_par = reinterpret_cast<float*>(sm + PAR_OFF);
  float *s_b1 = s_par, *s_g1 = s_par + H, *s_be1 = s_par + 2 * H, *s_b2 = s_par + 3 * H, *s_g2 = s_par + 3 * H + L,
        *s_be2 = s_par + 3 * H + 2 * L;
  int* s_eid = reinterpret_cast<int*>(sm + IDS_OFF);
  int* s_src = s_eid + TILE_M;
  int* s_dst = s_src + TILE_M;
  float* s_red = reinterpret_cast<float*>(sm + RED_OFF);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + BAR_OFF + NBAR * 8);
  const uint32_t sm_u = smem_u32(sm), bar0 = sm_u + BAR_OFF;
  if ((sm_u & 1023u) != 0) __trap();
  enum { W_FULL = 0, ST_FREE = 2, B_FULL = 4, B_FREE = 10, ACC = 16, A_REST = 17, W_KB2 = 18, KB2_DONE = 19 };
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
  const uint32_t idesc_h = make_idesc(TILE_M, H), idesc_l = make_idesc(TILE_M, L);
  const uint8_t* w1p = reinterpret_cast<const uint8_t*>(P.w1_packed);
  const uint8_t* w2p = reinterpret_cast<const uint8_t*>(P.w2_packed);

  uint32_t it1 = 0, acc_par = 0, tl = 0;  // ring use counter, accumulator-barrier parity, tiles done by this CTA
  int nx_eid = 0, nx_src = 0, nx_dst = 0;  // row ids of the next tile, prefetched
  uint32_t n_fill[NSLOT] = {0, 0, 0, 0, 0, 0}, n_commit[NSLOT] = {0, 0, 0, 0, 0, 0};  // thread 0 bookkeeping
  // per-lane column-sum accumulators: lane c of warp (q, cs) owns columns cs*64 + {c, 32 + c} of H and cs*32 + c of L
  float acc_db1[2] = {0.f, 0.f}, acc_dg1[2] = {0.f, 0.f}, acc_dbe1[2] = {0.f, 0.f};
  float acc_db2 = 0.f, acc_dg2 = 0.f, acc_dbe2 = 0.f;

  auto slot_fill = [&](int slot, uint32_t dst_off, const void* src, uint32_t bytes) {  // thread 0
    if (n_commit[slot] > 0) mbar_wait(BAR(B_FREE + slot), (n_commit[slot] - 1) & 1);
    mbar_expect_tx(BAR(B_FULL + slot), bytes);
    bulk_g2s(sm_u + dst_off, src, bytes, BAR(B_FULL + slot));
    n_fill[slot]++;
  };
  auto slot_wait_full = [&](int slot) { mbar_wait(BAR(B_FULL + slot), (n_fill[slot] - 1) & 1); tc_fence_after(); };
  auto slot_commit = [&](int slot) { umma_commit(BAR(B_FREE + slot)); n_commit[slot]++; };

  long long t_prev = clock64();
  auto MARK = [&](int ph) {
    if (A.phase_clk && tid == 0 && blockIdx.x == 0) { long long t = clock64(); atomicAdd(A.phase_clk + ph, (unsigned long long)(t - t_prev)); t_prev = t; }
  };
  long long t_sub = 0;
  auto SUB0 = [&]() { if (A.phase_clk && tid == 0 && blockIdx.x == 0) t_sub = clock64(); };
  auto SUB = [&](int i) {
    if (A.phase_clk && tid == 0 && blockIdx.x == 0) { long long t = clock64(); atomicAdd(A.phase_clk + 9 + i, (unsigned long long)(t - t_sub)); t_sub = t; }
  };
  // GEMM1 operand requests (thread 0). K-blocks 0/1: A0 image block + W1 block into ring stage (it & 1), one
  // transaction barrier for both; they are requested a tile ahead. K-blocks 2..5: the four A0 blocks land together in
  // the A2 region (idle until EPI-A) at tile start, only their W1 blocks go through the ring.
  auto g1_issue = [&](int t, int kb, uint32_t it) {
    const int s = it & 1;
    mbar_wait(BAR(ST_FREE + s), ((it >> 1) & 1) ^ 1);
    if (kb < 2) {
      mbar_expect_tx(BAR(W_FULL + s), A_BLK_BYTES + W1_BLK);
      bulk_g2s(sm_u + s * STAGE, A.a0_img + ((size_t)t * NKB1 + kb) * A_BLK_BYTES, A_BLK_BYTES, BAR(W_FULL + s));
    } else {
      mbar_expect_tx(BAR(W_FULL + s), W1_BLK);
    }
    bulk_g2s(sm_u + s * STAGE + A_BLK_BYTES, w1p + (size_t)kb * W1_BLK, W1_BLK, BAR(W_FULL + s));
  };
  // W1 K-block 2 is parked in the upstream-gradient staging region (idle from EPI-B of one tile to the end of the next
  // tile's GEMM1) so that GEMM1 itself only streams K-blocks 3..5
  auto w1_kb2_prefetch = [&]() {  // thread 0
    mbar_expect_tx(BAR(W_KB2), W1_BLK);
    bulk_g2s(sm_u + GS_OFF, w1p + (size_t)2 * W1_BLK, W1_BLK, BAR(W_KB2));
  };
  uint32_t rest_par = 0;
  const int n_tiles = (int)((A.n_edges + TILE_M - 1) / TILE_M);
  // All CTAs start together and every tile costs the same, so without this the SMs stay in lock-step and their
  // HBM-heavy phases (operand images + gradient rows in, gradient rows out) coincide: stagger the start by a fraction
  // of a tile so the memory system sees a steady demand instead of bursts.
  if (A.stagger_cycles > 0 && n_tiles > (int)gridDim.x) {
    const long long t0 = clock64(), wait = (long long)(blockIdx.x % 4) * A.stagger_cycles;
    while (clock64() - t0 < wait) {}
  }
  if (tid == 0 && (int)blockIdx.x < n_tiles) {
    g1_issue(blockIdx.x, 0, 0);
    g1_issue(blockIdx.x, 1, 1);
    w1_kb2_prefetch();
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ================= tile setup: row ids (fetched one tile ahead into registers; see the end of the loop) ==========
    if (tid < TILE_M) {
      if (tile == (int)blockIdx.x) {
        int64_t j = (int64_t)tile * TILE_M + tid;
        if (j >= A.n_edges) j = A.n_edges - 1;
        nx_eid = A.perm ? A.perm[j] : (int)j;
        nx_src = A.src[nx_eid];
        nx_dst = A.dst[nx_eid];
      }
      s_eid[tid] = nx_eid;
      s_src[tid] = nx_src;
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
    // ================= GEMM1 (recompute): D1 = A0 W1^T, operands by bulk copy only =================
    // Two single-thread roles in different warps: thread 0 only waits for operands and issues MMAs; thread 32 (the
    // "producer") waits for ring stages to retire and refills them, so the issuing thread never blocks on an MMA's
    // completion. K-blocks 0/1 were requested a tile ahead and W1 K-block 2 is parked in the gradient staging region.
    SUB0();
    if (tid == 0) {
      // the A2 region is free (previous tile's EPI-D staging was drained before the closing barrier)
      fence_proxy_async();
      mbar_expect_tx(BAR(A_REST), 4 * A_BLK_BYTES);
      bulk_g2s(sm_u + A2_OFF, A.a0_img + ((size_t)tile * NKB1 + 2) * A_BLK_BYTES, 4 * A_BLK_BYTES, BAR(A_REST));
      for (int kb = 0; kb < 2; ++kb) {
        const uint32_t u = it1 + kb;
        const int s = u & 1;
        mbar_wait(BAR(W_FULL + s), (u >> 1) & 1);
        tc_fence_after();
        umma_kblock(tmem + TM_D1, sm_u + s * STAGE, sm_u + s * STAGE + A_BLK_BYTES, idesc_h, kb == 0);
        umma_commit(BAR(ST_FREE + s));
      }
    }
    if (tid == 32) {  // ring uses of this tile: K-blocks 0, 1, 3, 4, 5 -> it1 .. it1 + 4
      g1_issue(tile, 3, it1 + 2);
      g1_issue(tile, 4, it1 + 3);
    }
    SUB(0);
    // upstream-gradient loads (every warp): 128 KB of register-staged loads per tile, in flight under the MMA loop
    const int g_sub = tid & 31, g_rr = tid >> 5;
    float4 gv[TILE_M / 16], ga[TILE_M / 16];
    auto gout_loads = [&]() {
#pragma unroll
      for (int p = 0; p < TILE_M / 16; ++p) {
        const int r = p * 16 + g_rr;
        const bool live = (int64_t)tile * TILE_M + r < A.n_edges;
        gv[p] = live ? __ldg(reinterpret_cast<const float4*>(A.g_e + (size_t)s_eid[r] * L) + g_sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        ga[p] = (live && A.g_agg) ? __ldg(reinterpret_cast<const float4*>(A.g_agg + (size_t)s_dst[r] * L) + g_sub)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    gout_loads();
    SUB(1);
    if (tid == 0) {
      mbar_wait(BAR(A_REST), rest_par);
      SUB(2);
      mbar_wait(BAR(W_KB2), rest_par);
      tc_fence_after();
      umma_kblock(tmem + TM_D1, sm_u + A2_OFF, sm_u + GS_OFF, idesc_h, false);
      umma_commit(BAR(KB2_DONE));
      SUB(3);
      for (int kb = 3; kb < NKB1; ++kb) {
        const uint32_t u = it1 + kb - 1;
        const int s = u & 1;
        mbar_wait(BAR(W_FULL + s), (u >> 1) & 1);
        tc_fence_after();
        umma_kblock(tmem + TM_D1, sm_u + A2_OFF + (kb - 2) * A_BLK_BYTES, sm_u + s * STAGE + A_BLK_BYTES, idesc_h, false);
        umma_commit(BAR(ST_FREE + s));
        if (kb == NKB1 - 1) umma_commit(BAR(ACC));
      }
    }
    if (tid == 32) g1_issue(tile, 5, it1 + 4);  // as soon as K-block 3's MMAs retire
    SUB(4);
    it1 += NKB1 - 1;
    rest_par ^= 1;
    if (warp == 0) mbar_wait(BAR(ACC), acc_par);  // one warp polls the mbarrier; the rest park on the hardware barrier
    __syncthreads();
    acc_par ^= 1;
    tc_fence_after();
    SUB(5);
    // GEMM1 has retired (W1 K-block 2 no longer needed in this region): stage the upstream gradient tile as a bf16 image (zero for padding rows): 32 threads per row, 16 rows per pass
    {
#pragma unroll
      for (int p = 0; p < TILE_M / 16; ++p) {
        const int r = p * 16 + g_rr;
        const int c = g_sub * 4;
        *reinterpret_cast<uint2*>(sm + GS_OFF + (c / KBLK) * A_BLK_BYTES + sw128_off(r, (c % KBLK) >> 3) + ((c >> 2) & 1) * 8) =
            make_uint2(pack_bf16(gv[p].x + ga[p].x, gv[p].y + ga[p].y), pack_bf16(gv[p].z + ga[p].z, gv[p].w + ga[p].w));
      }
    }
    // ring is idle: bring all of W2 in (4 x 16 KB slots) behind EPI-A
    if (tid == 0) {
      for (int j = 0; j < NKB2; ++j) slot_fill(j, j * W2_BLK, w2p + (size_t)j * W2_BLK, W2_BLK);
    }

    MARK(1);
    // ================= EPI-A: LN1 statistics, g = act(LN1(h1 + b1)) -> A2 image =================
    float mean1, rstd1;
    {
      const int c0 = cs * 64;
      float mloc, m2;
      ln_partial<2>(t_lane + TM_D1 + c0, s_b1 + c0, mloc, m2);
      s_red[row * 8 + cs * 2] = mloc;
      s_red[row * 8 + cs * 2 + 1] = m2;
      __syncthreads();
      combine4(s_red, row, 64, P.ln_eps, mean1, rstd1);
      ln_act_to_image<ACT_H, 2>(t_lane + TM_D1 + c0, s_b1, s_g1, s_be1, c0, mean1, rstd1, sm + A2_OFF, row);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    MARK(2);
    // ================= GEMM2 (recompute): D2 = g W2^T ; g image -> HBM =================
    if (tid == 0) {
      bulk_s2g(A.g_img + (size_t)tile * A2_BYTES, sm_u + A2_OFF, A2_BYTES);
      bulk_commit();
      tc_fence_after();
      for (int j = 0; j < NKB2; ++j) {
        slot_wait_full(j);
        umma_kblock(tmem + TM_D2, sm_u + A2_OFF + j * A_BLK_BYTES, sm_u + j * W2_BLK, idesc_l, j == 0);
        slot_commit(j);
      }
      umma_commit(BAR(ACC));
    }
    if (warp == 0) mbar_wait(BAR(ACC), acc_par);  // one warp polls the mbarrier; the rest park on the hardware barrier
    __syncthreads();
    acc_par ^= 1;
    tc_fence_after();
    // W2^T (2 x 32 KB) replaces W2 in the ring behind EPI-B
    if (tid == 0) {
      slot_fill(0, 0, A.w2t, W2T_BLK);
      slot_fill(2, W2T_BLK, A.w2t + W2T_BLK, W2T_BLK);
    }

    MARK(3);
    if (tid < TILE_M && has_next) { nx_src = A.src[nx_eid]; nx_dst = A.dst[nx_eid]; }  // second link, consumed next tile
    // ================= EPI-B: LN2 + act forward, adjoint down to delta2 =================
    {
      const int c0 = cs * 32;
      float v[32], dy[32];
      tmem_ld32(t_lane + TM_D2 + c0, v);
      const float4* b4 = reinterpret_cast<const float4*>(s_b2 + c0);
      const float4* g4 = reinterpret_cast<const float4*>(s_g2 + c0);
      const float4* e4 = reinterpret_cast<const float4*>(s_be2 + c0);
      float sum = 0.f, sq = 0.f;
      const float pv = v[0] + s_b2[c0];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b = b4[i];
        v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float d = v[4 * i + k] - pv; sum += d; sq = fmaf(d, d, sq); }
      }
      s_red[row * 8 + cs * 2] = fmaf(sum, 1.0f / 32, pv);      // EPI-A readers are past the pre-GEMM2 barrier
      s_red[row * 8 + cs * 2 + 1] = fmaxf(sq - sum * sum * (1.0f / 32), 0.f);
      __syncthreads();
      float mean2, rstd2;
      combine4(s_red, row, 32, P.ln_eps, mean2, rstd2);
      const float nmr = -mean2 * rstd2;
      // upstream gradient (bf16 image) for this quarter row: 4 chunks of 8
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        const int c = c0 + g8 * 8;
        const uint4 pk = *reinterpret_cast<const uint4*>(sm + GS_OFF + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3));
        const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
        const float4 ga = g4[2 * g8], gb = g4[2 * g8 + 1], ea = e4[2 * g8], eb = e4[2 * g8 + 1];
        const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const float ee[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float go = __uint_as_float((i & 1) ? (w[i >> 1] & 0xffff0000u) : (w[i >> 1] << 16));
          const int k = g8 * 8 + i;
          const float xh = fmaf(v[k], rstd2, nmr);
          const float d = go * tc_act_bwd<ACT_O>(fmaf(xh, gg[i], ee[i]));
          const float gd = gg[i] * d;
          dy[k] = d;
          v[k] = xh;
          s1 += gd;
          s2 = fmaf(gd, xh, s2);
        }
      }
      __syncthreads();  // everyone has read the LN2 statistics
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
      for (int g8 = 0; g8 < 4; ++g8) {
        const int c = c0 + g8 * 8;
        *reinterpret_cast<uint4*>(sm + D2IMG_OFF + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3)) =
            make_uint4(pack_bf16(dy[g8 * 8], dy[g8 * 8 + 1]), pack_bf16(dy[g8 * 8 + 2], dy[g8 * 8 + 3]),
                       pack_bf16(dy[g8 * 8 + 4], dy[g8 * 8 + 5]), pack_bf16(dy[g8 * 8 + 6], dy[g8 * 8 + 7]));
      }
      acc_db2 += warp_colsum32(dy, lane);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    MARK(4);
    // ================= GEMM3: dG = delta2 W2 (two N = 128 halves) ; delta2 image -> HBM =================
    if (tid == 0) {
      if (has_next) { fence_proxy_async(); w1_kb2_prefetch(); }  // gradient image consumed by EPI-B: region free again
      bulk_s2g(A.d2_img + (size_t)tile * GS_BYTES, sm_u + D2IMG_OFF, NKBL * A_BLK_BYTES);
      bulk_commit();
      tc_fence_after();
      for (int kb = 0; kb < NKBL; ++kb) {
        const int slot = kb * 2;
        slot_wait_full(slot);
        const uint32_t a_s = sm_u + D2IMG_OFF + kb * A_BLK_BYTES, b_s = sm_u + kb * W2T_BLK;
        umma_kblock(tmem + TM_DGLO, a_s, b_s, idesc_l, kb == 0);                      // hidden units [0, 128)
        umma_kblock(tmem + TM_DGHI, a_s, b_s + L * ROW_BYTES, idesc_l, kb == 0);      // hidden units [128, 256)
        slot_commit(slot);
      }
      umma_commit(BAR(ACC));
    }
    if (warp == 0) mbar_wait(BAR(ACC), acc_par);  // one warp polls the mbarrier; the rest park on the hardware barrier
    __syncthreads();
    acc_par ^= 1;
    tc_fence_after();
    // start streaming W1^T (segment, K-block) pieces into the six 16 KB slots behind EPI-C
    if (tid == 0) {
      bulk_wait_read0();  // g / delta2 images have left shared memory
      for (int b = 0; b < NSLOT; ++b) {
        const int sg = b / NKB2, kb = b % NKB2;
        slot_fill(b, b * SEG_BLK, A.w1t + (size_t)kb * W1T_BLK + (size_t)sg * SEG_BLK, SEG_BLK);
      }
    }

    MARK(5);
    // ================= EPI-C: d(y1) = dG * act'(y1), LN1 adjoint -> delta1 image =================
    {
      const int c0 = cs * 64;
      const uint32_t t_dg = t_lane + (cs < 2 ? TM_DGLO + c0 : TM_DGHI + (c0 - 128));
      const float nmr1 = -mean1 * rstd1;
      float v[32], u[32], tmp[32];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        tmem_ld32(t_lane + TM_D1 + c0 + ch * 32, v);
        tmem_ld32(t_dg + ch * 32, u);
        const int cb = c0 + ch * 32;
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b = *reinterpret_cast<const float4*>(s_b1 + cb + 4 * i4);
          const float4 g = *reinterpret_cast<const float4*>(s_g1 + cb + 4 * i4);
          const float4 be = *reinterpret_cast<const float4*>(s_be1 + cb + 4 * i4);
          const float bb[4] = {b.x, b.y, b.z, b.w}, gg[4] = {g.x, g.y, g.z, g.w}, ee[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 4 * i4 + k;
            const float xh = fmaf(v[i] + bb[k], rstd1, nmr1);
            const float d = u[i] * tc_act_bwd<ACT_H>(fmaf(xh, gg[k], ee[k]));
            const float gd = gg[k] * d;
            u[i] = d;
            v[i] = xh;
            s1 += gd;
            s2 = fmaf(gd, xh, s2);
          }
        }
        tmem_st32(t_dg + ch * 32, u);  // park d(y1) where dG was
#pragma unroll
        for (int i = 0; i < 32; ++i) tmp[i] = u[i] * v[i];
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
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        tmem_ld32(t_lane + TM_D1 + c0 + ch * 32, v);
        tmem_ld32(t_dg + ch * 32, u);
        const int cb = c0 + ch * 32;
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b = *reinterpret_cast<const float4*>(s_b1 + cb + 4 * i4);
          const float4 g = *reinterpret_cast<const float4*>(s_g1 + cb + 4 * i4);
          const float bb[4] = {b.x, b.y, b.z, b.w}, gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 4 * i4 + k;
            const float xh = fmaf(v[i] + bb[k], rstd1, nmr1);
            u[i] = rstd1 * (gg[k] * u[i] - t1 - xh * t2);
          }
        }
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
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

    MARK(6);
    // ================= GEMM4: dA0 = delta1 W1, one N = 128 accumulator per input segment =================
    if (tid == 0) {
      bulk_s2g(A.d1_img + (size_t)tile * A2_BYTES, sm_u + A2_OFF, A2_BYTES);
      bulk_commit();
      tc_fence_after();
      constexpr int NB = 3 * NKB2;  // 12 pieces, segment-major
      for (int b = 0; b < NB; ++b) {
        const int slot = b % NSLOT, sg = b / NKB2, kb = b % NKB2;
        slot_wait_full(slot);
        umma_kblock(tmem + TM_DA0 + sg * L, sm_u + A2_OFF + kb * A_BLK_BYTES, sm_u + slot * SEG_BLK, idesc_l, kb == 0);
        slot_commit(slot);
        if (b < NSLOT) n_fill[slot]++;  // pieces 6..11 are requested by the producer thread below
      }
      umma_commit(BAR(ACC));
      bulk_wait_read0();  // delta1 image has left shared memory before EPI-D reuses the region
    }
    if (tid == 32) {
      // producer: refill slot sl with piece sl + 6 once the MMAs of piece sl have retired. Completions of B_FREE[sl]
      // before this point: 4 / 3 / 4 / 3 / 2 / 2 per earlier tile (GEMM2 + GEMM3 + 2 x GEMM4) plus 2 / 1 / 2 / 1 / 0 / 0 in
      // this tile, so the awaited completion has parity 0 except on slots 1 and 3, where it alternates with the tile.
#pragma unroll 1
      for (int sl = 0; sl < NSLOT; ++sl) {
        const int b2 = sl + NSLOT, sg2 = b2 / NKB2, kb2 = b2 % NKB2;
        const uint32_t par = (sl == 1 || sl == 3) ? ((tl + 1) & 1u) : 0u;
        mbar_wait(BAR(B_FREE + sl), par);
        mbar_expect_tx(BAR(B_FULL + sl), SEG_BLK);
        bulk_g2s(sm_u + sl * SEG_BLK, A.w1t + (size_t)kb2 * W1T_BLK + (size_t)sg2 * SEG_BLK, SEG_BLK, BAR(B_FULL + sl));
      }
    }
    if (warp == 0) mbar_wait(BAR(ACC), acc_par);  // one warp polls the mbarrier; the rest park on the hardware barrier
    __syncthreads();  // also orders thread 0's bulk_wait_read0 before the staging writes below
    acc_par ^= 1;
    tc_fence_after();
    // every MMA that read the ring has retired: request the next tile's first two GEMM1 K-blocks under EPI-D
    if (tid == 0 && tile + (int)gridDim.x < n_tiles) {
      g1_issue(tile + gridDim.x, 0, it1);
      g1_issue(tile + gridDim.x, 1, it1 + 1);
    }

    MARK(7);
    // ================= EPI-D: rows of d(x[src]), d(x[dst]), d(e) through a swizzled fp32 staging tile =================
    float4 skipg[8];  // fp32 upstream gradient of this lane's 8 output chunks (skip path of d(e)), loaded early
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
#pragma unroll 1
    for (int sg = 0; sg < 3; ++sg) {
      {
        float v[32];
        tmem_ld32(t_lane + TM_DA0 + sg * L + cs * 32, v);
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          const int c4 = cs * 8 + g4;
          *reinterpret_cast<float4*>(sm + A2_OFF + (size_t)row * (L * 4) + ((c4 ^ (row & 7)) << 4)) =
              make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
        }
      }
      __syncthreads();
      float* outp = sg == 0 ? A.d_xs : (sg == 1 ? A.d_xd : A.d_e);
#pragma unroll
      for (int k = 0; k < 8; ++k) {  // 8 rows per warp, one float4 chunk per lane
        const int r = warp * 8 + k, c4 = lane;
        const int64_t j = (int64_t)tile * TILE_M + r;
        if (j < A.n_edges) {
          float4 y = *reinterpret_cast<const float4*>(sm + A2_OFF + (size_t)r * (L * 4) + ((c4 ^ (r & 7)) << 4));
          if (sg == 2) {  // skip connection: d(e) += gout (fp32)
            y.x += skipg[k].x; y.y += skipg[k].y; y.z += skipg[k].z; y.w += skipg[k].w;
          }
          *reinterpret_cast<float4*>(outp + (size_t)s_eid[r] * L + c4 * 4) = y;
        }
      }
      __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    ++tl;
    MARK(8);
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

// dvec1[3, H] = (db1, dgamma1, dbeta1), dvec2[3, L]: ordered sum over [grid * 4] partial vectors
__global__ void k_colpart_reduce(const float* __restrict__ part, int n_part, float* __restrict__ dvec1, float* __restrict__ dvec2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= PAR_FLOATS) return;
  float s = 0.f;
  for (int p = 0; p < n_part; ++p) s += part[(size_t)p * PAR_FLOATS + i];
  if (i < 3 * H) dvec1[i] = s; else dvec2[i - 3 * H] = s;
}

struct Layout {
  size_t g, d1, d2, colpart, wgrad, total;
  int grid, tiles;
  size_t wgrad_bytes;
};

Layout make_layout(int64_t n_edges) {
  Layout Y{};
  Y.tiles = (int)((n_edges + TILE_M - 1) / TILE_M);
  Y.grid = std::max(1, std::min(Y.tiles, num_sms()));
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = align_up(off, 1024); off = o + b; return o; };
  Y.g = take((size_t)Y.tiles * NKB2 * A_BLK_BYTES);
  Y.d1 = take((size_t)Y.tiles * NKB2 * A_BLK_BYTES);
  Y.d2 = take((size_t)Y.tiles * NKBL * A_BLK_BYTES);
  Y.colpart = take((size_t)Y.grid * 4 * PAR_FLOATS * 4);
  // wgrad partials: 3 roles of [256 x 128] + 1 role of [128 x 256]
  int splits = hgnn::tc::wgrad_splits(4, Y.tiles);
  Y.wgrad_bytes = (size_t)4 * align_up((size_t)splits * 256 * 128 * 4, 256) + 256;
  Y.wgrad = take(Y.wgrad_bytes);
  Y.total = align_up(off, 1024);
  return Y;
}

}  // namespace

static void* g_phase_clk = nullptr;
// debug hook (not part of the stable ABI): device buffer of 16 uint64 that CTA 0 fills with per-phase cycle counts
extern "C" void hgnn_tc_debug_set_phase_clock(void* dev_u64x16) { g_phase_clk = dev_u64x16; }

extern "C" size_t hgnn_tc_edge_backward_workspace_bytes(int64_t n_edges) {
  return make_layout(n_edges > 0 ? n_edges : 1).total + 1024;
}

extern "C" int hgnn_tc_edge_backward(const hgnn_tc_edge_params* p, const void* w1t_packed, const void* w2t_packed,
                                     const void* a0_img, const int32_t* src, const int32_t* dst, const int32_t* perm,
                                     int64_t n_edges, const float* grad_eout, const float* grad_agg, float* d_e, float* d_xsrc_rows,
                                     float* d_xdst_rows, float* dW1, float* dW2, float* dvec1, float* dvec2, void* ws,
                                     size_t ws_bytes, void* stream) {
  HGNN_REQUIRE(p != nullptr, "tc_edge_backward: params is NULL");
  HGNN_REQUIRE(p->latent == 128 && p->hidden == 256, "tc_edge_backward: only latent 128 / hidden 256 is built (got %d / %d)",
               p->latent, p->hidden);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_edges <= 0) {
    if (dW1) HGNN_CUDA_TRY(cudaMemsetAsync(dW1, 0, (size_t)H * K1 * 4, st));
    if (dW2) HGNN_CUDA_TRY(cudaMemsetAsync(dW2, 0, (size_t)L * H * 4, st));
    if (dvec1) HGNN_CUDA_TRY(cudaMemsetAsync(dvec1, 0, (size_t)3 * H * 4, st));
    if (dvec2) HGNN_CUDA_TRY(cudaMemsetAsync(dvec2, 0, (size_t)3 * L * 4, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(w1t_packed && w2t_packed && a0_img && src && dst && grad_eout && d_e && d_xsrc_rows && d_xdst_rows && dW1 && dW2 &&
               dvec1 && dvec2 && ws, "tc_edge_backward: NULL pointer");
  HGNN_REQUIRE(n_edges < INT32_MAX, "tc_edge_backward: too many edges");
  Layout Y = make_layout(n_edges);
  uintptr_t base = align_up((uintptr_t)ws, 1024);
  if (ws_bytes < (base - (uintptr_t)ws) + Y.total) return fail(HGNN_ERR_WORKSPACE, "tc_edge_backward: workspace too small");
  uint8_t* w = (uint8_t*)base;
  BwdArgs A{};
  A.P = *p;
  A.w1t = (const uint8_t*)w1t_packed;
  A.w2t = (const uint8_t*)w2t_packed;
  A.src = src; A.dst = dst; A.perm = perm;
  A.g_e = grad_eout; A.g_agg = grad_agg;
  A.d_e = d_e; A.d_xs = d_xsrc_rows; A.d_xd = d_xdst_rows;
  A.a0_img = (const uint8_t*)a0_img; A.g_img = w + Y.g; A.d1_img = w + Y.d1; A.d2_img = w + Y.d2;
  A.colpart = (float*)(w + Y.colpart);
  A.n_edges = n_edges;
  A.phase_clk = (unsigned long long*)g_phase_clk;
  {
    static int stagger = -1;
    if (stagger < 0) { const char* e = getenv("HGNN_BWD_STAGGER"); stagger = e ? atoi(e) : 17000; }
    A.stagger_cycles = stagger;
  }
  HGNN_REQUIRE(p->act_hidden == HGNN_ACT_GELU && p->act_out == HGNN_ACT_TANH, "tc_edge_backward: only GELU / Tanh is built");
  size_t smem = SMEM_BYTES;
  auto kern = k_tc_edge_bwd<HGNN_ACT_GELU, HGNN_ACT_TANH>;
  HGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<Y.grid, NT, smem, st>>>(A);
  int rc = check_launch("tc_edge_backward");
  if (rc) return rc;
  k_colpart_reduce<<<(PAR_FLOATS + 255) / 256, 256, 0, st>>>(A.colpart, Y.grid * 4, dvec1, dvec2);
  // weight gradients: dW1[:, seg] = delta1^T A0[:, seg] (3 problems), dW2 = delta2^T g
  hgnn::tc::WgradProblem pr[4];
  for (int s = 0; s < 3; ++s)
    pr[s] = hgnn::tc::WgradProblem{A.d1_img, H, 0, H, A.a0_img, K1, s * L, L, dW1, K1, 0, s * L, 0};
  pr[3] = hgnn::tc::WgradProblem{A.d2_img, L, 0, L, A.g_img, H, 0, H, dW2, H, 0, 0, 0};
  return hgnn::tc::launch_wgrad(pr, 4, Y.tiles, w + Y.wgrad, Y.wgrad_bytes, st);
}
