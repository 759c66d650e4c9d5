// Tensor-core fused edge step (forward), sm_100a:
//   e'[i] = Tanh(LN2(W2 . GELU(LN1(W1 . [x[src_i] | x[dst_i] | e_i] + b1)) + b2)) + e_i
// (InteractionGNNCell.edge_update, gnn_utils.py:56-64, with make_mlp's
// LayerNorm/activation layout, utils.py:183-196.)
//
// One CTA owns a 128-edge tile (UMMA M = 128 = TMEM lanes); two CTAs are
// resident per SM so that one tile's CUDA-core epilogue overlaps the other's
// gathers and MMAs. Per tile:
//   GEMM1  K-block by K-block (64 bf16): all threads gather the fp32 rows of
//          x[src] / x[dst] / e (coalesced 256 B row pieces), convert to bf16 and
//          write them 128B-swizzled into a 2-stage ring; the matching W1 K-block
//          (pre-packed bf16 UMMA image) arrives by one bulk async copy; one thread
//          issues 4 tcgen05.mma (128 x H x 16) into TMEM and commits to the
//          stage's mbarrier, so block k's MMAs run under block k+1's gather.
//   EPI1   each thread owns half a TMEM lane (row): bias + two-pass LayerNorm
//          (Chan-combined across the two half-row threads) + activation, written
//          as the bf16 K-major A operand of GEMM2 over the (now idle) ring.
//   GEMM2  W2 K-blocks streamed through the tail of the ring; accumulator
//          aliases the first L TMEM columns.
//   EPI2   bias + LayerNorm + Tanh to a swizzled fp32 staging tile, then a
//          coalesced pass adds the fp32 skip row and stores e' as full 512 B rows.
// Weights are never resident in full: smem per CTA is ~105 KB at L = 128.
// Training (stash != NULL) also leaves in HBM what the backward needs, so that it recomputes nothing: the two MMA
// operand images (the e columns of A0 written while gathering, g by one bulk copy), the normalised pre-affine activations of both
// LayerNorms as bf16 and the row rstd's (tc_common.cuh: EdgeStash, 1.5 KB per edge at L = 128).
#include <algorithm>
#include <cstdlib>

#include "tc_common.cuh"

using namespace hgnn;
using namespace hgnn::tc;

namespace {

constexpr int TC_THREADS = 256;

template <int L>
struct Cfg {
  static constexpr int H = 2 * L;
  static constexpr int K1 = 3 * L;
  static constexpr int NKB1 = K1 / KBLK;
  static constexpr int NKB2 = H / KBLK;
  static constexpr int W1_BLK = H * ROW_BYTES;
  static constexpr int W2_BLK = L * ROW_BYTES;
  static constexpr int STAGE = A_BLK_BYTES + W1_BLK;
  static constexpr int NSTAGE = 2;
  static constexpr int A2_BYTES = NKB2 * A_BLK_BYTES;
  static constexpr int SOUT_BYTES = TILE_M * L * 4;
  static constexpr int R0 = NSTAGE * STAGE;
  static constexpr int R1 = A2_BYTES + 2 * W2_BLK;
  static constexpr int REGION = (R0 > R1 ? R0 : R1) > SOUT_BYTES ? (R0 > R1 ? R0 : R1) : SOUT_BYTES;
  static constexpr int PARAM_FLOATS = 3 * H + 3 * L;
  // region | params | row ids (3 x 128 int) | LN exchange (128 x 2 x 2 float) | 16 barriers | tmem slot
  static constexpr int SMEM = REGION + PARAM_FLOATS * 4 + 4 * TILE_M * 4 + TILE_M * 4 * 4 + 16 * 8 + 16;
  static constexpr int TMEM_COLS = H;  // power of two >= 32 for L in {64,128}
  static_assert(L == 64 || L == 128, "tensor-core edge step is instantiated for latent 64 and 128");
};

template <int L, int ACT_H, int ACT_O>
__global__ void __launch_bounds__(TC_THREADS, 2)
k_tc_edge_fwd(hgnn_tc_edge_params P, const uint16_t* __restrict__ xb16, const float* __restrict__ e, const int32_t* __restrict__ src,
              const int32_t* __restrict__ dst, const int32_t* __restrict__ perm, int64_t n_edges, float* __restrict__ e_out,
              const int32_t* __restrict__ rowptr, float* __restrict__ agg, uint8_t* __restrict__ stash, EdgeStash SL,
              unsigned long long* __restrict__ phase_clk, int stagger_cycles) {
  using C = Cfg<L>;
  constexpr int H = C::H;
  extern __shared__ __align__(1024) uint8_t smem_raw[];  // declared alignment keeps the shared address space visible (LDS/STS)
  uint8_t* const region = smem_raw;
  float* s_par = reinterpret_cast<float*>(region + C::REGION);
  float *s_b1 = s_par, *s_g1 = s_par + H, *s_be1 = s_par + 2 * H, *s_b2 = s_par + 3 * H, *s_g2 = s_par + 3 * H + L,
        *s_be2 = s_par + 3 * H + 2 * L;
  int* s_eid = reinterpret_cast<int*>(s_par + C::PARAM_FLOATS);
  int* s_src = s_eid + TILE_M;
  int* s_dst = s_src + TILE_M;
  int* s_flag = s_dst + TILE_M;  // per row: 1 = last row of a segment that lies inside this row group (store the sum),
                                 //          2 = last row of a run that continues elsewhere (drop it; fix-up pass owns it), 0 = inside a run
  float* s_red = reinterpret_cast<float*>(s_flag + TILE_M);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + TILE_M * 4);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t region_u = smem_u32(region);
  if ((region_u & 1023u) != 0) __trap();  // swizzled operand images need 1024 B alignment
  const uint32_t bar0 = smem_u32(s_bar);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  enum { W_FULL = 0, ST_FREE = 2, W2_FULL = 4, W2_FREE = 6, ACC_FULL = 8 };

  if (tid == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(BAR(i), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), C::TMEM_COLS);
  for (int i = tid; i < H; i += TC_THREADS) { s_b1[i] = P.b1[i]; s_g1[i] = P.gamma1[i]; s_be1[i] = P.beta1[i]; }
  for (int i = tid; i < L; i += TC_THREADS) { s_b2[i] = P.b2[i]; s_g2[i] = P.gamma2[i]; s_be2[i] = P.beta2[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t idesc1 = make_idesc(TILE_M, H), idesc2 = make_idesc(TILE_M, L);
  const uint8_t* w1p = reinterpret_cast<const uint8_t*>(P.w1_packed);
  const uint8_t* w2p = reinterpret_cast<const uint8_t*>(P.w2_packed);
  // activations kept for the backward pass (NULL in inference): see EdgeStash in tc_common.cuh
  uint8_t* const a0_img = stash ? stash + SL.a0 : nullptr;
  uint8_t* const g_img = stash ? stash + SL.g : nullptr;
  uint4* const xh1_st = stash ? reinterpret_cast<uint4*>(stash + SL.xh1) : nullptr;
  uint4* const xh2_st = stash ? reinterpret_cast<uint4*>(stash + SL.xh2) : nullptr;
  float* const rstd_st = stash ? reinterpret_cast<float*>(stash + SL.rstd) : nullptr;

  // the fp32 edge rows are read twice per tile (GEMM1 operand, then the skip connection) with the tile's whole stash
  // streaming out in between: keep them in L2 after the first read, release them on the second
  const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  uint32_t it1 = 0, it2 = 0, acc_par = 0;
  const int q = warp & 3, hsel = warp >> 2;
  const int row = q * 32 + lane;
  const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
  const int n_tiles = (int)((n_edges + TILE_M - 1) / TILE_M);

  // desynchronise the CTAs' HBM-heavy phases (see the backward kernel): a fraction-of-a-tile start offset per CTA group
  if (stagger_cycles > 0 && n_tiles > (int)gridDim.x) {
    const long long t0 = clock64(), wait = (long long)(blockIdx.x % 4) * stagger_cycles;
    while (clock64() - t0 < wait) {}
  }
  long long t_prev = clock64();
  auto MARK = [&](int ph) {  // optional per-phase cycle accounting (CTA 0, thread 0)
    if (phase_clk && tid == 0 && blockIdx.x == 0) { long long t = clock64(); atomicAdd(phase_clk + ph, (unsigned long long)(t - t_prev)); t_prev = t; }
  };
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- row ids of this tile ----
    if (tid < TILE_M) {
      int64_t j = (int64_t)tile * TILE_M + tid;
      if (j >= n_edges) j = n_edges - 1;  // padding rows recompute the last edge; never stored
      int eid = perm ? perm[j] : (int)j;
      s_eid[tid] = eid;
      s_src[tid] = src[eid];
      const int d = dst[eid];
      s_dst[tid] = d;
      if (agg) {  // rows arrive destination-sorted (perm = the by-destination plan): classify run ends
        constexpr int G = TILE_M / (TC_THREADS / L);           // rows per reduction group
        const int64_t j0 = (int64_t)tile * TILE_M + tid;       // true sorted position (may be >= n_edges: padding)
        const int64_t gbeg = j0 / G * G, gend = gbeg + G < n_edges ? gbeg + G : n_edges;
        int flag = 0;
        if (j0 < n_edges) {
          const bool last = (j0 + 1 >= gend) || (dst[perm ? perm[j0 + 1] : (int)(j0 + 1)] != d);
          if (last) flag = (rowptr[d] >= gbeg && rowptr[d + 1] <= gend) ? 1 : 2;
        } else {
          flag = 2;
        }
        s_flag[tid] = flag;
      }
    }
    __syncthreads();
    MARK(0);

    // ---- GEMM1: D1[128, H] = [x[src] | x[dst] | e] . W1^T, K-blocks visited e first (its fp32 rows come from HBM: their
    // loads are issued at tile start), then x[src], x[dst] from the bf16 shadow copy of the node rows, two blocks in flight ----
    constexpr int KPS = L / KBLK;  // K-blocks per segment
    auto blk_seg = [&](int i) { return i < KPS ? 2 : (i < 2 * KPS ? 0 : 1); };
    auto blk_kb = [&](int i) { return blk_seg(i) * KPS + i % KPS; };  // K-block index in the concatenation / W1 image order
    auto blk_rows = [&](int i) { return blk_seg(i) == 0 ? s_src : s_dst; };
    float4 pre[8];
    uint4 xv[2][4];
    gather_load_hint(pre, e, L, s_eid, 0, pol_keep);
    if constexpr (KPS == 1) gather_load_bf16(xv[0], xb16, L, blk_rows(KPS), 0);
#pragma unroll
    for (int i = 0; i < C::NKB1; ++i, ++it1) {
      const int s = it1 % C::NSTAGE;
      const uint32_t ph = (it1 / C::NSTAGE) & 1;
      uint8_t* stage = region + s * C::STAGE;
      mbar_wait(BAR(ST_FREE + s), ph ^ 1);  // MMAs that last read this stage are done
      if (tid == 0) {
        mbar_expect_tx(BAR(W_FULL + s), C::W1_BLK);
        bulk_g2s(region_u + s * C::STAGE + A_BLK_BYTES, w1p + (size_t)blk_kb(i) * C::W1_BLK, C::W1_BLK, BAR(W_FULL + s));
      }
      if (i < KPS) {  // the edge-latent K-blocks are also left in HBM as the weight-gradient operand (x columns: handled per node)
        gather_store(stage, pre, a0_img ? a0_img + ((size_t)tile * KPS + i) * A_BLK_BYTES : nullptr);
        if (i + 1 < KPS) gather_load_hint(pre, e, L, s_eid, (i + 1) * KBLK, pol_keep);
      } else {
        gather_store_bf16(stage, xv[(i - KPS) & 1]);
      }
      if (i + 2 >= KPS && i + 2 < C::NKB1)  // node rows two blocks ahead
        gather_load_bf16(xv[(i + 2 - KPS) & 1], xb16, L, blk_rows(i + 2), ((i + 2) % KPS) * KBLK);
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        mbar_wait(BAR(W_FULL + s), ph);
        tc_fence_after();
        umma_kblock(tmem, region_u + s * C::STAGE, region_u + s * C::STAGE + A_BLK_BYTES, idesc1, i == 0);
        umma_commit(BAR(ST_FREE + s));
        if (i == C::NKB1 - 1) umma_commit(BAR(ACC_FULL));
      }
    }
    if (warp == 0) mbar_wait(BAR(ACC_FULL), acc_par);  // one warp polls the mbarrier; the rest park on the hardware barrier
    __syncthreads();
    acc_par ^= 1;
    tc_fence_after();
    MARK(1);

    // W2 K-blocks 0/1 stream in behind EPI1 (their slots lie past the A2 image; GEMM1 no longer reads them)
    if (tid == 0) {
      for (int j = 0; j < 2 && j < C::NKB2; ++j) {
        const uint32_t u = it2 + j, sl = u & 1, ph = (u >> 1) & 1;
        mbar_wait(BAR(W2_FREE + sl), ph ^ 1);
        mbar_expect_tx(BAR(W2_FULL + sl), C::W2_BLK);
        bulk_g2s(region_u + C::A2_BYTES + sl * C::W2_BLK, w2p + (size_t)j * C::W2_BLK, C::W2_BLK, BAR(W2_FULL + sl));
      }
    }

    // ---- EPI1: bias + LayerNorm + activation -> bf16 A2 (K-major, swizzled) ----
    {
      constexpr int NC = H / 2;  // columns per thread
      const int c0 = hsel * NC;
      float mloc, m2;
      ln_partial<NC / 32>(t_lane + c0, s_b1 + c0, mloc, m2);
      s_red[row * 4 + hsel * 2] = mloc;
      s_red[row * 4 + hsel * 2 + 1] = m2;
      __syncthreads();
      const LnStat st = combine_halves(s_red, row, NC, P.ln_eps);
      ln_act_to_image<ACT_H, NC / 32>(t_lane + c0, s_b1, s_g1, s_be1, c0, st.mean, st.rstd, region, row,
                                      xh1_st ? xh1_st + (size_t)tile * (H / 8) * TILE_M : nullptr);
      if (rstd_st && hsel == 0) rstd_st[(size_t)tile * 2 * TILE_M + row] = st.rstd;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    MARK(2);

    // ---- GEMM2: D2[128, L] = A2 . W2^T (accumulator aliases TMEM columns [0, L)) ----
    if (tid == 0) {
      if (g_img) {  // the finished A2 image is also the weight-gradient operand of layer 2: one bulk copy to HBM
        bulk_s2g(g_img + (size_t)tile * C::A2_BYTES, region_u, C::A2_BYTES);
        bulk_commit();
      }
      tc_fence_after();
      for (int j = 0; j < C::NKB2; ++j) {
        const uint32_t u = it2 + j, sl = u & 1, ph = (u >> 1) & 1;
        mbar_wait(BAR(W2_FULL + sl), ph);
        tc_fence_after();
        umma_kblock(tmem, region_u + j * A_BLK_BYTES, region_u + C::A2_BYTES + sl * C::W2_BLK, idesc2, j == 0);
        umma_commit(BAR(W2_FREE + sl));
        if (j + 2 < C::NKB2) {  // refill this slot with block j+2 once its MMAs retire
          const uint32_t u2 = u + 2, ph2 = (u2 >> 1) & 1;
          mbar_wait(BAR(W2_FREE + sl), ph2 ^ 1);
          mbar_expect_tx(BAR(W2_FULL + sl), C::W2_BLK);
          bulk_g2s(region_u + C::A2_BYTES + sl * C::W2_BLK, w2p + (size_t)(j + 2) * C::W2_BLK, C::W2_BLK, BAR(W2_FULL + sl));
        }
      }
      umma_commit(BAR(ACC_FULL));
    }
    it2 += C::NKB2;
    if (warp == 0) mbar_wait(BAR(ACC_FULL), acc_par);  // one warp polls the mbarrier; the rest park on the hardware barrier
    __syncthreads();
    acc_par ^= 1;
    tc_fence_after();
    MARK(3);

    // ---- EPI2: bias + LayerNorm + activation -> fp32 staging tile (swizzled 16 B chunks) ----
    {
      constexpr int NC = L / 2;
      const int c0 = hsel * NC;
      float mloc, m2;
      ln_partial<NC / 32>(t_lane + c0, s_b2 + c0, mloc, m2);
      s_red[row * 4 + hsel * 2] = mloc;   // EPI1's readers passed the pre-GEMM2 barrier long ago
      s_red[row * 4 + hsel * 2 + 1] = m2;
      if (tid == 0 && g_img) bulk_wait_read0();  // the g image has left shared memory before the region becomes staging
      __syncthreads();
      const LnStat st = combine_halves(s_red, row, NC, P.ln_eps);
      const float nmr = -st.mean * st.rstd;
      if (rstd_st && hsel == 0) rstd_st[(size_t)tile * 2 * TILE_M + TILE_M + row] = st.rstd;
      uint4* const xh2_t = xh2_st ? xh2_st + (size_t)tile * (L / 8) * TILE_M : nullptr;
      float v[32];
#pragma unroll 1
      for (int ch = 0; ch < NC / 32; ++ch) {
        tmem_ld32(t_lane + c0 + ch * 32, v);
        const int cb = c0 + ch * 32;
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float xh[8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = cb + g8 * 8 + 4 * h;
            const float4 b = *reinterpret_cast<const float4*>(s_b2 + c);
            const float4 g = *reinterpret_cast<const float4*>(s_g2 + c);
            const float4 be = *reinterpret_cast<const float4*>(s_be2 + c);
            const float2 x0 = fma2(add2(make_float2(v[g8 * 8 + 4 * h], v[g8 * 8 + 4 * h + 1]), make_float2(b.x, b.y)), splat2(st.rstd), splat2(nmr));
            const float2 x1 = fma2(add2(make_float2(v[g8 * 8 + 4 * h + 2], v[g8 * 8 + 4 * h + 3]), make_float2(b.z, b.w)), splat2(st.rstd), splat2(nmr));
            xh[4 * h + 0] = x0.x; xh[4 * h + 1] = x0.y; xh[4 * h + 2] = x1.x; xh[4 * h + 3] = x1.y;
            const float2 o0 = tc_act2<ACT_O>(fma2(x0, make_float2(g.x, g.y), make_float2(be.x, be.y)));
            const float2 o1 = tc_act2<ACT_O>(fma2(x1, make_float2(g.z, g.w), make_float2(be.z, be.w)));
            const float4 o = make_float4(o0.x, o0.y, o1.x, o1.y);
            *reinterpret_cast<float4*>(region + (size_t)row * (L * 4) + (((c >> 2) ^ (row & 7)) << 4)) = o;
          }
          if (xh2_t)
            xh2_t[(size_t)((cb + g8 * 8) >> 3) * TILE_M + row] =
                make_uint4(pack_bf16(xh[0], xh[1]), pack_bf16(xh[2], xh[3]), pack_bf16(xh[4], xh[5]), pack_bf16(xh[6], xh[7]));
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    MARK(4);
    // ---- coalesced pass: + fp32 skip row, full-row stores. The skip rows are fetched eight at a time into registers
    // before anything depends on them (one L2 round trip per batch instead of one per row) ----
    {
      constexpr int CPR = L / 4;                       // float4 chunks per row
      constexpr int ROWS_PER_WARP = TILE_M / (TC_THREADS / 32);
      constexpr int ITERS = ROWS_PER_WARP * CPR / 32;
      constexpr int BATCH = 8;
      static_assert(ITERS % BATCH == 0, "store pass batches");
#pragma unroll 1
      for (int b0 = 0; b0 < ITERS; b0 += BATCH) {
        float4 sk[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {  // padding rows carry the (valid) id of the last edge: load unconditionally
          const int idx = lane + (b0 + u) * 32;
          const int r = warp * ROWS_PER_WARP + idx / CPR, c4 = idx % CPR;
          sk[u] = ldg_f4_hint(reinterpret_cast<const float4*>(e + (size_t)s_eid[r] * L + c4 * 4), pol_drop);
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
          const int idx = lane + (b0 + u) * 32;
          const int r = warp * ROWS_PER_WARP + idx / CPR, c4 = idx % CPR;
          float4* sp = reinterpret_cast<float4*>(region + (size_t)r * (L * 4) + ((c4 ^ (r & 7)) << 4));
          const float4 y = *sp;
          const float4 o = make_float4(y.x + sk[u].x, y.y + sk[u].y, y.z + sk[u].z, y.w + sk[u].w);
          if ((int64_t)tile * TILE_M + r < n_edges) {
            *reinterpret_cast<float4*>(e_out + (size_t)s_eid[r] * L + c4 * 4) = o;
            if (agg) *sp = o;
          }
        }
      }
    }
    MARK(5);
    if (agg) {
      // ---- fused scatter_add: destination-sorted segmented reduce of the finished tile, ordered, no atomics.
      // One thread per (row group, column): runs of equal destination are summed in row order; a run that lies
      // inside the group is stored, runs crossing a group boundary are left to the fix-up pass.
      __syncthreads();
      constexpr int G = TILE_M / (TC_THREADS / L);
      const int grp = tid / L, col = tid % L;
      // run-end / store masks of this row group, gathered once (warp-uniform control flow below)
      uint64_t endm = 0, storem = 0;
#pragma unroll
      for (int w = 0; w < G / 32; ++w) {
        const int f = s_flag[grp * G + w * 32 + lane];
        endm |= (uint64_t)__ballot_sync(0xffffffffu, f != 0) << (32 * w);
        storem |= (uint64_t)__ballot_sync(0xffffffffu, f == 1) << (32 * w);
      }
      const uint8_t* colp = region + (((col >> 2) << 4) | ((col & 3) << 2));
      float acc = 0.f;
#pragma unroll 8
      for (int i = 0; i < G; ++i) {
        const int r = grp * G + i;
        acc += *reinterpret_cast<const float*>(colp + (size_t)r * (L * 4) - (((col >> 2) << 4)) + ((((col >> 2) ^ (r & 7)) << 4)));
        if ((endm >> i) & 1) {
          if ((storem >> i) & 1) agg[(size_t)s_dst[r] * L + col] = acc;
          acc = 0.f;
        }
      }
    }
    __syncthreads();  // staging tile / row ids free for the next tile
    MARK(6);
  }

  if (tid == 0 && g_img) bulk_wait0();  // outstanding image stores
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, C::TMEM_COLS);
}

#include "edge_fwd_pp.cuh"

// bf16 shadow copy of the node rows (each row is gathered ~2 E/N times per step: convert once, gather half the bytes)
__global__ void __launch_bounds__(256) k_rows_to_bf16(const float* __restrict__ x, int64_t n8, uint4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
  out[i] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
}

// segments the fused reduce did not finish: empty (-> zeros) or spanning a row-group boundary (-> ordered sum of e_out rows)
template <int L>
__global__ void __launch_bounds__(256) k_agg_fixup(const float* __restrict__ e_out, const int32_t* __restrict__ perm,
                                                   const int32_t* __restrict__ rowptr, int64_t n_nodes, float* __restrict__ agg) {
  constexpr int G = TILE_M / (TC_THREADS / L);
  constexpr int CH = L / 4;
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t s = t / CH;
  int c = (int)(t % CH);
  if (s >= n_nodes) return;
  const int beg = rowptr[s], end = rowptr[s + 1];
  if (end > beg && beg / G == (end - 1) / G) return;  // finished inside the edge kernel
  if (end - beg > 512) return;                        // hub segment: summed by the per-CTA long-segment kernel
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int j = beg;
  for (; j + 16 <= end; j += 16) {  // hub nodes: 16 row loads in flight, summed in row order
    float4 v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = *(reinterpret_cast<const float4*>(e_out + (size_t)(perm ? perm[j + u] : j + u) * L) + c);
#pragma unroll
    for (int u = 0; u < 16; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  for (; j < end; ++j) {
    const float4 v = *(reinterpret_cast<const float4*>(e_out + (size_t)(perm ? perm[j] : j) * L) + c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4*>(agg + (size_t)s * L)[c] = acc;
}

// ---------------------------------------------------------------------------
// weight packing: fp32 [N, K] (nn.Linear layout) -> bf16 K-blocks of [N, 64], 128B-swizzled
// ---------------------------------------------------------------------------
__global__ void k_pack_weights(const float* __restrict__ W, int N, int K, uint8_t* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (n, 8-column group)
  int groups = K / 8;
  if (t >= N * groups) return;
  int n = t / groups, g = t % groups;
  int kb = g / 8, c16 = g % 8;
  const float* s = W + (size_t)n * K + g * 8;
  uint4 pk = make_uint4(pack_bf16(s[0], s[1]), pack_bf16(s[2], s[3]), pack_bf16(s[4], s[5]), pack_bf16(s[6], s[7]));
  *reinterpret_cast<uint4*>(out + (size_t)kb * N * ROW_BYTES + sw128_off(n, c16)) = pk;
}

// ---------------------------------------------------------------------------
// debug / unit-test GEMM: C[M, N] = bf16(A[M, K]) . bf16(W[N, K])^T with the same building blocks
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_debug_gemm(const float* __restrict__ A, const uint8_t* __restrict__ Wp, int64_t M, int N, int K, float* __restrict__ Cout) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* const region = smem_raw;
  const int wblk = N * ROW_BYTES;
  int* s_rid = reinterpret_cast<int*>(region + A_BLK_BYTES + wblk);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_rid + TILE_M);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t region_u = smem_u32(region), bar_w = smem_u32(s_bar), bar_mma = bar_w + 8;
  int tcols = 32;
  while (tcols < N) tcols <<= 1;
  if (tid == 0) { mbar_init(bar_w, 1); mbar_init(bar_mma, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), tcols);
  if (tid < TILE_M) {
    int64_t r = (int64_t)blockIdx.x * TILE_M + tid;
    s_rid[tid] = (int)(r < M ? r : M - 1);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t idesc = make_idesc(TILE_M, N);
  uint32_t par = 0;
  for (int kb = 0; kb < K / KBLK; ++kb) {
    if (tid == 0) {
      mbar_expect_tx(bar_w, wblk);
      bulk_g2s(region_u + A_BLK_BYTES, Wp + (size_t)kb * wblk, wblk, bar_w);
    }
    gather_a_block(region, A, K, s_rid, kb * KBLK);
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_w, par);
      tc_fence_after();
      umma_kblock(tmem, region_u, region_u + A_BLK_BYTES, idesc, kb == 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, par);  // fully serialised: this kernel only checks layouts/descriptors
    par ^= 1;
    tc_fence_after();
  }
  const int q = warp & 3, hsel = warp >> 2;
  const int row = q * 32 + lane;
  const int64_t grow = (int64_t)blockIdx.x * TILE_M + row;
  float v[32];
  for (int c0 = hsel * 32; c0 < N; c0 += 64) {
    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
    if (grow < M)
      for (int i = 0; i < 32; ++i) Cout[grow * N + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, tcols);
}

}  // namespace

extern "C" int hgnn_tc_supported(int64_t latent, int64_t hidden, int64_t n_layers, int layer_norm, int act_hidden, int act_out) {
  return (latent == 64 || latent == 128) && hidden == 2 * latent && n_layers == 2 && layer_norm &&
         act_hidden == HGNN_ACT_GELU && act_out == HGNN_ACT_TANH ? 1 : 0;
}

extern "C" size_t hgnn_tc_packed_weight_bytes(int64_t out_features, int64_t in_features) {
  return (size_t)out_features * in_features * 2;
}

extern "C" int hgnn_tc_pack_weights(const float* W, int64_t out_features, int64_t in_features, void* packed, void* stream) {
  HGNN_REQUIRE(W && packed, "tc_pack_weights: NULL pointer");
  HGNN_REQUIRE(in_features % KBLK == 0 && out_features % 8 == 0 && out_features <= 256,
               "tc_pack_weights: need in_features %% 64 == 0, out_features %% 8 == 0 and <= 256 (got %lld x %lld)",
               (long long)out_features, (long long)in_features);
  int total = (int)(out_features * in_features / 8);
  k_pack_weights<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, (int)out_features, (int)in_features, (uint8_t*)packed);
  return check_launch("tc_pack_weights");
}

extern "C" int hgnn_tc_debug_gemm(const float* A, const void* w_packed, int64_t M, int64_t N, int64_t K, float* C, void* stream) {
  HGNN_REQUIRE(A && w_packed && C && M > 0, "tc_debug_gemm: bad argument");
  HGNN_REQUIRE(K % KBLK == 0 && N % 32 == 0 && N >= 32 && N <= 256, "tc_debug_gemm: need K %% 64 == 0, N %% 32 == 0, N <= 256");
  size_t smem = A_BLK_BYTES + (size_t)N * ROW_BYTES + TILE_M * 4 + 64;
  HGNN_CUDA_TRY(cudaFuncSetAttribute(k_tc_debug_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned grid = (unsigned)((M + TILE_M - 1) / TILE_M);
  k_tc_debug_gemm<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(A, (const uint8_t*)w_packed, M, (int)N, (int)K, C);
  return check_launch("tc_debug_gemm");
}

static int fwd_pp_enabled() {  // HGNN_FWD_PP=1 selects the persistent warp-specialised ping-pong kernel (A/B comparisons;
  static int v = -1;           // measured equal to the two-CTA kernel within 5 %, see profiles/r02_fwd_pingpong.md)
  if (v < 0) { const char* e = getenv("HGNN_FWD_PP"); v = e ? atoi(e) : 0; }
  return v;
}
#ifdef HGNN_DEBUG_MBAR
// debug builds only: {timed out?, block, thread, barrier smem address, parity, dynamic smem bytes, -, -} of the first stuck wait
extern "C" int hgnn_tc_debug_mbar_timeout(int* out8) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out8, g_mbar_dbg, sizeof(int) * 8);
}
#endif
#ifdef HGNN_TRACE
// trace builds only: copies the event records of CTA 0 (3 x uint64 each) and resets the counter; returns the record count
extern "C" int hgnn_tc_debug_trace(unsigned long long* out, int max_records) {
  cudaDeviceSynchronize();
  unsigned int n = 0;
  cudaMemcpyFromSymbol(&n, pp::g_trace_n, sizeof(n));
  if ((int)n > max_records) n = (unsigned)max_records;
  if (n > 8192) n = 8192;
  cudaMemcpyFromSymbol(out, pp::g_trace, sizeof(unsigned long long) * 3 * n);
  unsigned int z = 0;
  cudaMemcpyToSymbol(pp::g_trace_n, &z, sizeof(z));
  return (int)n;
}
#endif
static int fwd_stagger() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("HGNN_FWD_STAGGER"); v = e ? atoi(e) : 0; }
  return v;
}

extern "C" size_t hgnn_tc_edge_forward_workspace_bytes(int64_t n_edges, int64_t n_nodes, int64_t latent) {
  (void)n_edges;
  return align_up((size_t)(n_nodes > 0 ? n_nodes : 1) * (size_t)latent * 2, 256) + 256;  // bf16 shadow copy of x
}
extern "C" size_t hgnn_tc_edge_stash_bytes(int64_t n_edges, int64_t latent) {
  return edge_stash_layout(n_edges > 0 ? n_edges : 1, (int)latent).total;
}

template <int L>
static int launch_edge_fwd(const hgnn_tc_edge_params* p, const float* x, const float* e, const int32_t* src, const int32_t* dst,
                           const int32_t* perm, int64_t n_edges, float* e_out, const int32_t* rowptr, int64_t n_nodes, float* agg,
                           uint8_t* stash, void* ws, size_t ws_bytes, cudaStream_t st) {
  uint16_t* xb16 = (uint16_t*)align_up((uintptr_t)ws, 256);
  if (ws == nullptr || ws_bytes < ((uintptr_t)xb16 - (uintptr_t)ws) + (size_t)n_nodes * L * 2)
    return fail(HGNN_ERR_WORKSPACE, "tc_edge_forward: workspace too small (hgnn_tc_edge_forward_workspace_bytes)");
  {
    const int64_t n8 = n_nodes * L / 8;
    k_rows_to_bf16<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(x, n8, reinterpret_cast<uint4*>(xb16));
  }
  int64_t tiles = (n_edges + TILE_M - 1) / TILE_M;
  if (fwd_pp_enabled() && L == 128) {  // persistent ping-pong kernel: one CTA per SM, two tiles in flight
    if constexpr (L == 128) {
      size_t smem = pp::PpCfg<L>::SMEM;
      auto kern = pp::k_tc_edge_fwd_pp<L, HGNN_ACT_GELU, HGNN_ACT_TANH>;
      HGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)num_sms());
      kern<<<grid, pp::PP_THREADS, smem, st>>>(*p, xb16, e, src, dst, perm, n_edges, e_out, rowptr, agg, stash, edge_stash_layout(n_edges, L));
    }
  } else {
    size_t smem = Cfg<L>::SMEM;
    auto kern = k_tc_edge_fwd<L, HGNN_ACT_GELU, HGNN_ACT_TANH>;
    HGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)std::min<int64_t>(tiles, 2 * (int64_t)num_sms());
    kern<<<grid, TC_THREADS, smem, st>>>(*p, xb16, e, src, dst, perm, n_edges, e_out, rowptr, agg, stash, edge_stash_layout(n_edges, L),
                                          (unsigned long long*)p->debug_phase_clock, fwd_stagger());
  }
  if (agg) {
    int64_t threads = n_nodes * (L / 4);
    k_agg_fixup<L><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(e_out, perm, rowptr, n_nodes, agg);
    int rc = launch_segment_reduce(e_out, L, nullptr, nullptr, perm, rowptr, n_nodes, 0, agg, /*skip_short=*/true, st);
    if (rc) return rc;
  }
  return check_launch("tc_edge_forward");
}

extern "C" int hgnn_tc_edge_forward(const hgnn_tc_edge_params* p, const float* x, const float* e, const int32_t* src,
                                    const int32_t* dst, const int32_t* perm, const int32_t* rowptr, int64_t n_edges,
                                    int64_t n_nodes, float* e_out, float* agg, void* stash, void* ws, size_t ws_bytes,
                                    void* stream) {
  HGNN_REQUIRE(agg == nullptr || (rowptr != nullptr && n_nodes > 0),
               "tc_edge_forward: the fused aggregate needs the destination-sorted plan (rowptr; perm unless the edges are stored sorted) and n_nodes");
  if (n_edges <= 0 && agg != nullptr) {
    HGNN_CUDA_TRY(cudaMemsetAsync(agg, 0, (size_t)n_nodes * p->latent * 4, (cudaStream_t)stream));
  }
  HGNN_REQUIRE(p != nullptr, "tc_edge_forward: params is NULL");
  if (n_edges <= 0) return HGNN_OK;
  HGNN_REQUIRE(x && e && src && dst && e_out, "tc_edge_forward: NULL pointer");
  HGNN_REQUIRE(n_nodes > 0 && n_nodes < INT32_MAX, "tc_edge_forward: n_nodes (rows of x) is required");
  HGNN_REQUIRE(p->w1_packed && p->w2_packed && p->b1 && p->gamma1 && p->beta1 && p->b2 && p->gamma2 && p->beta2,
               "tc_edge_forward: NULL parameter pointer");
  HGNN_REQUIRE(n_edges < INT32_MAX, "tc_edge_forward: too many edges");
  if (!hgnn_tc_supported(p->latent, p->hidden, 2, 1, p->act_hidden, p->act_out))
    return fail(HGNN_ERR_UNSUPPORTED,
                "tc_edge_forward: latent %d / hidden %d / activations (%d, %d) not built (need latent in {64,128}, hidden = 2*latent, GELU/Tanh)",
                p->latent, p->hidden, p->act_hidden, p->act_out);
  cudaStream_t st = (cudaStream_t)stream;
  if (p->latent == 128) return launch_edge_fwd<128>(p, x, e, src, dst, perm, n_edges, e_out, rowptr, n_nodes, agg, (uint8_t*)stash, ws, ws_bytes, st);
  return launch_edge_fwd<64>(p, x, e, src, dst, perm, n_edges, e_out, rowptr, n_nodes, agg, (uint8_t*)stash, ws, ws_bytes, st);
}
