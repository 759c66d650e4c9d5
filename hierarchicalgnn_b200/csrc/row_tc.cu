// Tensor-core row layer (sm_100a): one make_mlp layer (utils.py:183-196) applied to a gathered concatenation
//   out[r] = act(LayerNorm(W . [seg0[i0(r)] | seg1[i1(r)] | seg2[i2(r)]] + b)) (+ skip row)
// This is the building block of the node / supernode updates (gnn_utils.py:45-54,119-127,137-145), of the encoder
// layers past the first and of the classifier heads' hidden layers (EC/Models/IN.py:29-48,126; BC/Models/HGNN_GMM.py:
// 37-82,342-344): every Linear whose fan-in is a multiple of 128 (<= 384) and whose fan-out is 128 or 256.
//
// Forward  (256 threads, 2 CTAs / SM): gather fp32 rows -> bf16 128B-swizzled K-blocks (2-stage ring, also left in HBM
//          as the "A image" when a backward will follow) -> tcgen05.mma into TMEM -> bias + LayerNorm + activation
//          straight from TMEM -> swizzled fp32 staging -> coalesced full-row stores (+ fp32 skip row).
// Backward (512 threads, 1 CTA / SM, all 512 TMEM columns):
//   GEMM-re  h = A W^T from the saved A image (bulk copies only, no gather)       [recompute: no activations are kept]
//   EPI      LayerNorm statistics, d(y) = gout * act'(y) parked in TMEM, LayerNorm adjoint -> delta (bf16 image that
//            overwrites the staged upstream gradient in place), ordered column sums for d bias / d gamma / d beta
//   GEMM-d   dA[128, K] = delta W, one N = 128 accumulator per 128 input columns, W^T pieces streamed through 6 slots
//   EPI-D    rows of dA through a swizzled fp32 staging tile (the skip path's gradient is gout itself: no kernel work)
//   wgrad    dW = delta^T A by the MN-major split-K kernel of wgrad_tc.cu over the two images.
#include <algorithm>

#include "tc_common.cuh"

using namespace hgnn;
using namespace hgnn::tc;

namespace {

constexpr int RF_THREADS = 256;
constexpr int RB_THREADS = 512;
constexpr int MAX_SEG = 3;

struct RowFwdArgs {
  const float* seg_ptr[MAX_SEG];
  const int32_t* seg_idx[MAX_SEG];
  int seg_width[MAX_SEG];
  int n_seg, nkb;
  const uint8_t* w_packed;
  const float *bias, *gamma, *beta;
  float eps;
  const float* skip;  // [rows, N] rows added after the activation (NULL: none)
  float* out;
  int64_t ld_out;     // row stride of out (floats); the N output columns start at col0 (plain-GEMM mode writes a column block)
  int col0;
  uint8_t* a_img;     // optional [tiles][nkb][16 KB]
  int64_t rows;
};

template <int N>
struct RowCfg {
  static constexpr int W_BLK = N * ROW_BYTES;
  static constexpr int STAGE = A_BLK_BYTES + W_BLK;
  static constexpr int RING = 2 * STAGE;
  static constexpr int PC = N / 2 < 64 ? N / 2 : 64;     // columns a thread stages per output pass
  static constexpr int SW = 2 * PC;                      // columns staged per output pass (both column halves)
  static constexpr int STAGING = TILE_M * SW * 4;
  static constexpr int REGION = RING > STAGING ? RING : STAGING;
  // region | bias, gamma, beta | row ids (3 x 128) | LN exchange (128 x 4) | 8 barriers | tmem slot
  static constexpr int SMEM = REGION + 3 * N * 4 + MAX_SEG * TILE_M * 4 + TILE_M * 4 * 4 + 8 * 8 + 16;
};

// RAW = plain GEMM: out[:, col0 : col0 + N] = [segments] . W^T (+ bias), no LayerNorm / activation / residual
template <int N, int ACT, bool RAW>
__global__ void __launch_bounds__(RF_THREADS, 2) k_tc_row_fwd(RowFwdArgs A) {
  using C = RowCfg<N>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* const region = smem_raw;
  float* s_b = reinterpret_cast<float*>(region + C::REGION);
  float* s_g = s_b + N;
  float* s_be = s_g + N;
  int* s_ids = reinterpret_cast<int*>(s_be + N);
  float* s_red = reinterpret_cast<float*>(s_ids + MAX_SEG * TILE_M);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + TILE_M * 4);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t region_u = smem_u32(region);
  if ((region_u & 1023u) != 0) __trap();
  const uint32_t bar0 = smem_u32(s_bar);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  enum { W_FULL = 0, ST_FREE = 2, ACC_FULL = 4 };

  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(BAR(i), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), N);
  for (int i = tid; i < N; i += RF_THREADS) {
    s_b[i] = A.bias ? A.bias[i] : 0.f;
    s_g[i] = RAW ? 1.f : A.gamma[i];
    s_be[i] = RAW ? 0.f : A.beta[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t idesc = make_idesc(TILE_M, N);

  uint32_t it = 0, acc_par = 0;
  const int q = warp & 3, hsel = warp >> 2;
  const int row = q * 32 + lane;
  const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
  const int n_tiles = (int)((A.rows + TILE_M - 1) / TILE_M);
  // K-block -> (segment, first column inside the segment)
  auto kb_seg = [&](int kb, int& col) {
    int c = kb * KBLK, s = 0;
    while (s + 1 < A.n_seg && c >= A.seg_width[s]) { c -= A.seg_width[s]; ++s; }
    col = c;
    return s;
  };

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (tid < TILE_M) {
      const int64_t j = (int64_t)tile * TILE_M + tid;
      // padding rows of the last tile gather zeros (id -1): never stored, and zero in the saved image so that the
      // weight-gradient GEMM over the images sees no phantom rows
      for (int s = 0; s < A.n_seg; ++s) s_ids[s * TILE_M + tid] = j < A.rows ? (A.seg_idx[s] ? A.seg_idx[s][j] : (int)j) : -1;
    }
    __syncthreads();

    // ---- GEMM: D[128, N] = [segments] . W^T ----
    float4 pre[8];
    {
      int col;
      const int s = kb_seg(0, col);
      gather_load(pre, A.seg_ptr[s], A.seg_width[s], s_ids + s * TILE_M, col);
    }
    for (int kb = 0; kb < A.nkb; ++kb, ++it) {
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      uint8_t* stage = region + st * C::STAGE;
      mbar_wait(BAR(ST_FREE + st), ph ^ 1);
      if (tid == 0) {
        mbar_expect_tx(BAR(W_FULL + st), C::W_BLK);
        bulk_g2s(region_u + st * C::STAGE + A_BLK_BYTES, A.w_packed + (size_t)kb * C::W_BLK, C::W_BLK, BAR(W_FULL + st));
      }
      gather_store(stage, pre, A.a_img ? A.a_img + ((size_t)tile * A.nkb + kb) * A_BLK_BYTES : nullptr);
      if (kb + 1 < A.nkb) {
        int col;
        const int s = kb_seg(kb + 1, col);
        gather_load(pre, A.seg_ptr[s], A.seg_width[s], s_ids + s * TILE_M, col);
      }
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        mbar_wait(BAR(W_FULL + st), ph);
        tc_fence_after();
        umma_kblock(tmem, region_u + st * C::STAGE, region_u + st * C::STAGE + A_BLK_BYTES, idesc, kb == 0);
        umma_commit(BAR(ST_FREE + st));
        if (kb == A.nkb - 1) umma_commit(BAR(ACC_FULL));
      }
    }
    if (warp == 0) mbar_wait(BAR(ACC_FULL), acc_par);
    __syncthreads();
    acc_par ^= 1;
    tc_fence_after();

    // ---- epilogue: bias + LayerNorm + activation from TMEM, staged 128 columns at a time ----
    {
      constexpr int NC = N / 2;           // columns per thread
      constexpr int PC = C::PC;           // columns a thread stages per pass
      constexpr int PASSES = NC / PC;
      const int c0 = hsel * NC;
      float rstd = 1.f, nmr = 0.f;
      if constexpr (!RAW) {
        float mloc, m2;
        ln_partial<NC / 32>(t_lane + c0, s_b + c0, mloc, m2);
        s_red[row * 4 + hsel * 2] = mloc;
        s_red[row * 4 + hsel * 2 + 1] = m2;
        __syncthreads();
        const LnStat st = combine_halves(s_red, row, NC, A.eps);
        rstd = st.rstd;
        nmr = -st.mean * st.rstd;
      }
      float v[32];
#pragma unroll 1
      for (int p = 0; p < PASSES; ++p) {
#pragma unroll 1
        for (int ch = 0; ch < PC / 32; ++ch) {
          const int cb = c0 + p * PC + ch * 32;            // parameter / accumulator column
          const int sc4 = (hsel * PC + ch * 32) >> 2;      // first float4 chunk inside the staging row
          tmem_ld32(t_lane + cb, v);
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const int c = cb + g4 * 4;
            const float4 b = *reinterpret_cast<const float4*>(s_b + c);
            const float4 g = *reinterpret_cast<const float4*>(s_g + c);
            const float4 be = *reinterpret_cast<const float4*>(s_be + c);
            float4 o;
            if constexpr (RAW) {
              o = make_float4(v[g4 * 4 + 0] + b.x, v[g4 * 4 + 1] + b.y, v[g4 * 4 + 2] + b.z, v[g4 * 4 + 3] + b.w);
            } else {
              const float2 rs2 = splat2(rstd), nmr2 = splat2(nmr);
              const float2 o0 = tc_act2<ACT>(fma2(fma2(add2(make_float2(v[g4 * 4 + 0], v[g4 * 4 + 1]), make_float2(b.x, b.y)), rs2, nmr2),
                                                  make_float2(g.x, g.y), make_float2(be.x, be.y)));
              const float2 o1 = tc_act2<ACT>(fma2(fma2(add2(make_float2(v[g4 * 4 + 2], v[g4 * 4 + 3]), make_float2(b.z, b.w)), rs2, nmr2),
                                                  make_float2(g.z, g.w), make_float2(be.z, be.w)));
              o = make_float4(o0.x, o0.y, o1.x, o1.y);
            }
            *reinterpret_cast<float4*>(region + (size_t)row * (C::SW * 4) + (((sc4 + g4) ^ (row & 7)) << 4)) = o;
          }
        }
        __syncthreads();
        {
          constexpr int CPR = C::SW / 4;                       // float4 chunks per staged row
          constexpr int ROWS_PER_WARP = TILE_M / (RF_THREADS / 32);
          constexpr int ITERS = ROWS_PER_WARP * CPR / 32;
          constexpr int BATCH = ITERS % 8 == 0 ? 8 : 4;
          static_assert(ITERS % BATCH == 0, "store pass batches");
          const bool has_skip = !RAW && A.skip != nullptr;
#pragma unroll 1
          for (int b0 = 0; b0 < ITERS; b0 += BATCH) {
            float4 sk[BATCH];  // residual rows of the whole batch in flight before anything depends on them
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
              const int idx = lane + (b0 + u) * 32;
              const int r = warp * ROWS_PER_WARP + idx / CPR, c4 = idx % CPR;
              const int64_t j = (int64_t)tile * TILE_M + r;
              const int sc = c4 * 4;
              const size_t g = (size_t)j * A.ld_out + A.col0 + (sc / PC) * NC + p * PC + (sc % PC);
              sk[u] = (has_skip && j < A.rows) ? __ldg(reinterpret_cast<const float4*>(A.skip + g)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
              const int idx = lane + (b0 + u) * 32;
              const int r = warp * ROWS_PER_WARP + idx / CPR, c4 = idx % CPR;
              const int64_t j = (int64_t)tile * TILE_M + r;
              if (j < A.rows) {
                float4 y = *reinterpret_cast<const float4*>(region + (size_t)r * (C::SW * 4) + ((c4 ^ (r & 7)) << 4));
                const int sc = c4 * 4;
                const size_t g = (size_t)j * A.ld_out + A.col0 + (sc / PC) * NC + p * PC + (sc % PC);
                y.x += sk[u].x; y.y += sk[u].y; y.z += sk[u].z; y.w += sk[u].w;
                *reinterpret_cast<float4*>(A.out + g) = y;
              }
            }
          }
        }
        __syncthreads();
      }
    }
    fence_proxy_async();  // staging (generic proxy) precedes the next tile's bulk copies into the same bytes
    tc_fence_before();
    __syncthreads();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, N);
}

// ---------------------------------------------------------------------------------------------------------------------
struct RowBwdArgs {
  int nkb, K;
  const uint8_t* w_packed;   // W image: K/64 blocks of [N rows x 128 B]
  const uint8_t* wt_packed;  // W^T image: N/64 blocks of [K rows x 128 B]
  const float *bias, *gamma, *beta;
  float eps;
  const uint8_t* a_img;
  const float* gout;         // [rows, N]
  float* d_piece[3];         // d(input) columns [128 p, 128 p + 128): base pointer of the piece (NULL = not wanted) ...
  int d_ld[3];               // ... and its row stride in floats (one [rows, K] matrix, or one dense matrix per gathered segment)
  uint8_t* d_img;            // [tiles][N/64][16 KB] delta image (operand of the weight-gradient GEMM)
  float* colpart;            // [grid][4][3 N]
  int64_t rows;
};

template <int N>
struct RowBCfg {
  static constexpr int W_BLK = N * ROW_BYTES;
  static constexpr int STAGE = A_BLK_BYTES + W_BLK;
  static constexpr int NSLOT = 6;
  static constexpr int SEG_BLK = 128 * ROW_BYTES;  // one 128-row piece of a W^T K-block
  static constexpr int RING = 2 * STAGE > NSLOT * SEG_BLK ? 2 * STAGE : NSLOT * SEG_BLK;
  static constexpr int GS_OFF = RING;
  static constexpr int GS_BYTES = (N / KBLK) * A_BLK_BYTES;
  static constexpr int PAR_OFF = GS_OFF + GS_BYTES;
  static constexpr int RED_OFF = PAR_OFF + 3 * N * 4;
  static constexpr int BAR_OFF = RED_OFF + TILE_M * 8 * 4;
  static constexpr int NBAR = 2 + 2 + NSLOT + NSLOT + 1;
  static constexpr int SMEM = BAR_OFF + NBAR * 8 + 16;
};

template <int N, int ACT>
__global__ void __launch_bounds__(RB_THREADS, 1) k_tc_row_bwd(RowBwdArgs A) {
  using C = RowBCfg<N>;
  constexpr int NSLOT = C::NSLOT;
  constexpr int NQ = N / 4;       // columns per thread
  constexpr int NCH = NQ / 32;    // 32-column chunks per thread
  constexpr int NKBN = N / KBLK;  // K-blocks of the data-gradient GEMM (reduction over N)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* const sm = smem_raw;
  float* s_b = reinterpret_cast<float*>(sm + C::PAR_OFF);
  float* s_g = s_b + N;
  float* s_be = s_g + N;
  float* s_red = reinterpret_cast<float*>(sm + C::RED_OFF);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + C::BAR_OFF + C::NBAR * 8);
  const uint32_t sm_u = smem_u32(sm), bar0 = sm_u + C::BAR_OFF;
  if ((sm_u & 1023u) != 0) __trap();
  enum { W_FULL = 0, ST_FREE = 2, B_FULL = 4, B_FREE = 4 + NSLOT, ACC = 4 + 2 * NSLOT };
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cs = warp >> 2;
  const int row = q * 32 + lane;

  if (tid == 0) {
    for (int i = 0; i < C::NBAR; ++i) mbar_init(BAR(i), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
  for (int i = tid; i < N; i += RB_THREADS) { s_b[i] = A.bias[i]; s_g[i] = A.gamma[i]; s_be[i] = A.beta[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
  const uint32_t idesc_n = make_idesc(TILE_M, N), idesc_128 = make_idesc(TILE_M, 128);
  const uint32_t TM_DY = N;  // d(y) parked next to h

  uint32_t it = 0, acc_par = 0;
  uint32_t n_fill[NSLOT] = {0, 0, 0, 0, 0, 0}, n_commit[NSLOT] = {0, 0, 0, 0, 0, 0};  // thread 0 bookkeeping
  float acc_db[NCH], acc_dg[NCH], acc_dbe[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) acc_db[i] = acc_dg[i] = acc_dbe[i] = 0.f;

  auto slot_fill = [&](int slot, const void* src) {  // thread 0
    if (n_commit[slot] > 0) mbar_wait(BAR(B_FREE + slot), (n_commit[slot] - 1) & 1);
    mbar_expect_tx(BAR(B_FULL + slot), C::SEG_BLK);
    bulk_g2s(sm_u + slot * C::SEG_BLK, src, C::SEG_BLK, BAR(B_FULL + slot));
    n_fill[slot]++;
  };
  auto piece_src = [&](int b) {  // piece b = (128 input columns sg, K-block kb over N)
    const int sg = b / NKBN, kb = b % NKBN;
    return A.wt_packed + (size_t)kb * A.K * ROW_BYTES + (size_t)sg * C::SEG_BLK;
  };
  auto re_issue = [&](int t, int kb, uint32_t u) {  // thread 0: A image block + W block of K-block kb into ring stage u & 1
    const int s = u & 1;
    mbar_wait(BAR(ST_FREE + s), ((u >> 1) & 1) ^ 1);
    mbar_expect_tx(BAR(W_FULL + s), A_BLK_BYTES + C::W_BLK);
    bulk_g2s(sm_u + s * C::STAGE, A.a_img + ((size_t)t * A.nkb + kb) * A_BLK_BYTES, A_BLK_BYTES, BAR(W_FULL + s));
    bulk_g2s(sm_u + s * C::STAGE + A_BLK_BYTES, A.w_packed + (size_t)kb * C::W_BLK, C::W_BLK, BAR(W_FULL + s));
  };

  const int n_tiles = (int)((A.rows + TILE_M - 1) / TILE_M);
  const int n_piece = A.K / 128, NB = n_piece * NKBN;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    if (tid == 0) {
      fence_proxy_async();  // the previous tile's staging traffic (generic proxy) precedes these bulk writes
      for (int kb = 0; kb < 2 && kb < A.nkb; ++kb) re_issue(tile, kb, it + kb);
    }
    // ---- upstream gradient tile -> bf16 image (zero rows past the end): coalesced row pieces ----
    {
      constexpr int TPR = N / 4;                 // threads per row (one float4 each)
      constexpr int RPP = RB_THREADS / TPR;      // rows per pass
      constexpr int NP = TILE_M / RPP;
      const int sub = tid % TPR, rr = tid / TPR;
      const int c = sub * 4;
      const uint32_t coff = (uint32_t)(c / KBLK) * A_BLK_BYTES + (uint32_t)((c >> 2) & 1) * 8;
#pragma unroll 1
      for (int p0 = 0; p0 < NP; p0 += 8) {
        float4 gv[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int r = (p0 + p) * RPP + rr;
          const int64_t j = (int64_t)tile * TILE_M + r;
          gv[p] = j < A.rows ? __ldg(reinterpret_cast<const float4*>(A.gout + (size_t)j * N) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int r = (p0 + p) * RPP + rr;
          *reinterpret_cast<uint2*>(sm + C::GS_OFF + coff + sw128_off(r, (c % KBLK) >> 3)) =
              make_uint2(pack_bf16(gv[p].x, gv[p].y), pack_bf16(gv[p].z, gv[p].w));
        }
      }
    }
    // ---- GEMM-re: h = A W^T (operands by bulk copy) ----
    if (tid == 0) {
      for (int kb = 0; kb < A.nkb; ++kb) {
        const uint32_t u = it + kb;
        const int s = u & 1;
        mbar_wait(BAR(W_FULL + s), (u >> 1) & 1);
        tc_fence_after();
        umma_kblock(tmem, sm_u + s * C::STAGE, sm_u + s * C::STAGE + A_BLK_BYTES, idesc_n, kb == 0);
        umma_commit(BAR(ST_FREE + s));
        if (kb == A.nkb - 1) umma_commit(BAR(ACC));
        if (kb + 2 < A.nkb) re_issue(tile, kb + 2, u + 2);
      }
    }
    it += A.nkb;
    if (warp == 0) mbar_wait(BAR(ACC), acc_par);
    __syncthreads();
    acc_par ^= 1;
    tc_fence_after();
    // ring idle: first W^T pieces stream in behind the epilogue
    if (tid == 0) {
      for (int b = 0; b < NSLOT && b < NB; ++b) slot_fill(b, piece_src(b));
    }

    // ---- EPI: LayerNorm statistics; d(y) = gout * act'(y); LayerNorm adjoint -> delta image ----
    {
      const int c0 = cs * NQ;
      float mean, rstd;
      {
        float mloc, m2;
        ln_partial<NCH>(t_lane + c0, s_b + c0, mloc, m2);
        s_red[row * 8 + cs * 2] = mloc;
        s_red[row * 8 + cs * 2 + 1] = m2;
        __syncthreads();
        combine4(s_red, row, NQ, A.eps, mean, rstd);
      }
      const float nmr = -mean * rstd;
      float v[32], u[32], tmp[32];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        const int cb = c0 + ch * 32;
        tmem_ld32(t_lane + cb, v);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int c = cb + g8 * 8;
          const uint4 pk = *reinterpret_cast<const uint4*>(sm + C::GS_OFF + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3));
          const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 b = *reinterpret_cast<const float4*>(s_b + c + 4 * h);
            const float4 g = *reinterpret_cast<const float4*>(s_g + c + 4 * h);
            const float4 be = *reinterpret_cast<const float4*>(s_be + c + 4 * h);
            const float bb[4] = {b.x, b.y, b.z, b.w}, gg[4] = {g.x, g.y, g.z, g.w}, ee[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int i8 = 4 * h + k, i = g8 * 8 + i8;
              const float go = __uint_as_float((i8 & 1) ? (w[i8 >> 1] & 0xffff0000u) : (w[i8 >> 1] << 16));
              const float xh = fmaf(v[i] + bb[k], rstd, nmr);
              const float d = go * tc_act_bwd<ACT>(fmaf(xh, gg[k], ee[k]));
              const float gd = gg[k] * d;
              u[i] = d;
              v[i] = xh;
              s1 += gd;
              s2 = fmaf(gd, xh, s2);
            }
          }
        }
        tmem_st32(t_lane + TM_DY + cb, u);  // park d(y)
#pragma unroll
        for (int i = 0; i < 32; ++i) tmp[i] = u[i] * v[i];
        acc_dg[ch] += warp_colsum32(tmp, lane);
        acc_dbe[ch] += warp_colsum32(u, lane);
      }
      __syncthreads();  // every thread has read the LayerNorm statistics
      s_red[row * 8 + cs * 2] = s1;
      s_red[row * 8 + cs * 2 + 1] = s2;
      __syncthreads();
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { t1 += s_red[row * 8 + 2 * i]; t2 += s_red[row * 8 + 2 * i + 1]; }
      t1 *= (1.0f / N);
      t2 *= (1.0f / N);
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        const int cb = c0 + ch * 32;
        tmem_ld32(t_lane + cb, v);
        tmem_ld32(t_lane + TM_DY + cb, u);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b = *reinterpret_cast<const float4*>(s_b + cb + 4 * i4);
          const float4 g = *reinterpret_cast<const float4*>(s_g + cb + 4 * i4);
          const float bb[4] = {b.x, b.y, b.z, b.w}, gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 4 * i4 + k;
            const float xh = fmaf(v[i] + bb[k], rstd, nmr);
            u[i] = rstd * (gg[k] * u[i] - t1 - xh * t2);
          }
        }
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int c = cb + g8 * 8;
          *reinterpret_cast<uint4*>(sm + C::GS_OFF + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3)) =
              make_uint4(pack_bf16(u[g8 * 8], u[g8 * 8 + 1]), pack_bf16(u[g8 * 8 + 2], u[g8 * 8 + 3]),
                         pack_bf16(u[g8 * 8 + 4], u[g8 * 8 + 5]), pack_bf16(u[g8 * 8 + 6], u[g8 * 8 + 7]));
        }
        acc_db[ch] += warp_colsum32(u, lane);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    // ---- GEMM-d: dA = delta W, 128 input columns per accumulator; delta image -> HBM ----
    if (tid == 0) {
      bulk_s2g(A.d_img + (size_t)tile * C::GS_BYTES, sm_u + C::GS_OFF, C::GS_BYTES);
      bulk_commit();
      tc_fence_after();
      for (int b = 0; b < NB; ++b) {
        const int slot = b % NSLOT, sg = b / NKBN, kb = b % NKBN;
        mbar_wait(BAR(B_FULL + slot), (n_fill[slot] - 1) & 1);
        tc_fence_after();
        umma_kblock(tmem + sg * 128, sm_u + C::GS_OFF + kb * A_BLK_BYTES, sm_u + slot * C::SEG_BLK, idesc_128, kb == 0);
        umma_commit(BAR(B_FREE + slot));
        n_commit[slot]++;
        if (b >= 1 && b - 1 + NSLOT < NB) {  // refill the slot consumed one step ago
          const int b2 = b - 1 + NSLOT;
          slot_fill(b2 % NSLOT, piece_src(b2));
        }
      }
      umma_commit(BAR(ACC));
      bulk_wait_read0();  // delta image has left shared memory
    }
    if (warp == 0) mbar_wait(BAR(ACC), acc_par);
    __syncthreads();
    acc_par ^= 1;
    tc_fence_after();

    // ---- EPI-D: rows of dA (128 columns per pass) through a swizzled fp32 staging tile over the idle ring ----
#pragma unroll 1
    for (int sg = 0; sg < n_piece; ++sg) {
      {
        float v[32];
        tmem_ld32(t_lane + sg * 128 + cs * 32, v);
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          const int c4 = cs * 8 + g4;
          *reinterpret_cast<float4*>(sm + (size_t)row * 512 + ((c4 ^ (row & 7)) << 4)) =
              make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k) {  // 8 rows per warp, one float4 chunk per lane
        const int r = warp * 8 + k, c4 = lane;
        const int64_t j = (int64_t)tile * TILE_M + r;
        if (j < A.rows && A.d_piece[sg] != nullptr) {
          float4 y = *reinterpret_cast<const float4*>(sm + (size_t)r * 512 + ((c4 ^ (r & 7)) << 4));
          *reinterpret_cast<float4*>(A.d_piece[sg] + (size_t)j * A.d_ld[sg] + c4 * 4) = y;
        }
      }
      __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
  }

  // ---- ordered hand-off of the column sums: [cta][q][3 N] = d bias | d gamma | d beta ----
  {
    float* o = A.colpart + ((size_t)blockIdx.x * 4 + q) * (3 * N);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int c = cs * NQ + ch * 32 + lane;
      o[c] = acc_db[ch];
      o[N + c] = acc_dg[ch];
      o[2 * N + c] = acc_dbe[ch];
    }
  }
  if (tid == 0) bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

bool act_built(int act) { return act == HGNN_ACT_GELU || act == HGNN_ACT_TANH || act == HGNN_ACT_RELU || act == HGNN_ACT_SILU; }

int layer_k(const hgnn_tc_row_layer* d) {
  int k = 0;
  for (int s = 0; s < d->n_seg; ++s) k += d->seg_width[s];
  return k;
}

struct BwdLayout {
  size_t d_img, colpart, wgrad, total, wgrad_bytes;
  int grid, tiles, n_prob;
};

BwdLayout bwd_layout(int64_t rows, int K, int N) {
  BwdLayout Y{};
  Y.tiles = (int)((rows + TILE_M - 1) / TILE_M);
  Y.grid = std::max(1, std::min(Y.tiles, num_sms()));
  Y.n_prob = K / 128;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = align_up(off, 1024); off = o + b; return o; };
  Y.d_img = take((size_t)Y.tiles * (N / KBLK) * A_BLK_BYTES);
  Y.colpart = take((size_t)Y.grid * 4 * 3 * N * 4);
  int splits = wgrad_splits(Y.n_prob, Y.tiles);
  Y.wgrad_bytes = (size_t)Y.n_prob * align_up((size_t)splits * N * 128 * 4, 256) + 256;
  Y.wgrad = take(Y.wgrad_bytes);
  Y.total = align_up(off, 1024);
  return Y;
}

template <int N, int ACT, bool RAW = false>
int launch_fwd(const RowFwdArgs& a, cudaStream_t st) {
  auto kern = k_tc_row_fwd<N, ACT, RAW>;
  size_t smem = RowCfg<N>::SMEM;
  HGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t tiles = (a.rows + TILE_M - 1) / TILE_M;
  unsigned grid = (unsigned)std::min<int64_t>(tiles, 2 * (int64_t)num_sms());
  kern<<<grid, RF_THREADS, smem, st>>>(a);
  return check_launch("tc_row_forward");
}

template <int N, int ACT>
int launch_bwd(const RowBwdArgs& a, int grid, cudaStream_t st) {
  auto kern = k_tc_row_bwd<N, ACT>;
  size_t smem = RowBCfg<N>::SMEM;
  HGNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, RB_THREADS, smem, st>>>(a);
  return check_launch("tc_row_backward");
}

#define ROW_DISPATCH(FN, N, ACT, ...)                                                        \
  ((N) == 128 ? ((ACT) == HGNN_ACT_GELU   ? FN<128, HGNN_ACT_GELU>(__VA_ARGS__)              \
                 : (ACT) == HGNN_ACT_TANH ? FN<128, HGNN_ACT_TANH>(__VA_ARGS__)              \
                 : (ACT) == HGNN_ACT_RELU ? FN<128, HGNN_ACT_RELU>(__VA_ARGS__)              \
                                          : FN<128, HGNN_ACT_SILU>(__VA_ARGS__))             \
              : ((ACT) == HGNN_ACT_GELU   ? FN<256, HGNN_ACT_GELU>(__VA_ARGS__)              \
                 : (ACT) == HGNN_ACT_TANH ? FN<256, HGNN_ACT_TANH>(__VA_ARGS__)              \
                 : (ACT) == HGNN_ACT_RELU ? FN<256, HGNN_ACT_RELU>(__VA_ARGS__)              \
                                          : FN<256, HGNN_ACT_SILU>(__VA_ARGS__)))

int validate(const hgnn_tc_row_layer* d, const char* who) {
  HGNN_REQUIRE(d != nullptr, "%s: layer descriptor is NULL", who);
  if (!hgnn_tc_row_supported(d)) {
    return fail(HGNN_ERR_UNSUPPORTED,
                "%s: layer shape not built (need 1..3 segments of width %% 64 == 0, fan-in %% 128 == 0 and <= 384, fan-out 128 or 256, "
                "LayerNorm, activation in GELU/Tanh/ReLU/SiLU)", who);
  }
  return HGNN_OK;
}

}  // namespace

extern "C" int hgnn_tc_row_supported(const hgnn_tc_row_layer* d) {
  if (!d || d->n_seg < 1 || d->n_seg > MAX_SEG) return 0;
  int k = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    if (d->seg_width[s] <= 0 || d->seg_width[s] % KBLK != 0) return 0;
    k += d->seg_width[s];
  }
  if (k % 128 != 0 || k > 384) return 0;
  if (d->n_out != 128 && d->n_out != 256) return 0;
  if (!act_built(d->act)) return 0;
  return 1;
}

extern "C" size_t hgnn_tc_row_image_bytes(int64_t rows, int64_t k) {
  int64_t tiles = (rows + TILE_M - 1) / TILE_M;
  return (size_t)tiles * (k / KBLK) * A_BLK_BYTES;
}

extern "C" int hgnn_tc_row_forward(const hgnn_tc_row_layer* d, int64_t rows, float* out, void* a_img, void* stream) {
  int rc = validate(d, "tc_row_forward");
  if (rc) return rc;
  if (rows <= 0) return HGNN_OK;
  HGNN_REQUIRE(out && d->w_packed && d->bias && d->gamma && d->beta, "tc_row_forward: NULL pointer");
  HGNN_REQUIRE(rows < INT32_MAX, "tc_row_forward: too many rows");
  RowFwdArgs a{};
  a.n_seg = d->n_seg;
  for (int s = 0; s < d->n_seg; ++s) {
    HGNN_REQUIRE(d->seg_ptr[s] != nullptr, "tc_row_forward: segment %d is NULL", s);
    a.seg_ptr[s] = d->seg_ptr[s]; a.seg_idx[s] = d->seg_idx[s]; a.seg_width[s] = d->seg_width[s];
  }
  a.nkb = layer_k(d) / KBLK;
  a.w_packed = (const uint8_t*)d->w_packed;
  a.bias = d->bias; a.gamma = d->gamma; a.beta = d->beta; a.eps = d->ln_eps;
  a.skip = d->skip;
  a.out = out; a.ld_out = d->n_out; a.col0 = 0; a.a_img = (uint8_t*)a_img; a.rows = rows;
  cudaStream_t st = (cudaStream_t)stream;
  return ROW_DISPATCH(launch_fwd, d->n_out, d->act, a, st);
}

extern "C" int hgnn_tc_gemm_supported(const hgnn_tc_row_layer* d) {
  if (!d || d->n_seg < 1 || d->n_seg > MAX_SEG) return 0;
  int k = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    if (d->seg_width[s] <= 0 || d->seg_width[s] % KBLK != 0) return 0;
    k += d->seg_width[s];
  }
  return k <= 768 && (d->n_out == 64 || d->n_out == 128 || d->n_out == 256);
}

extern "C" int hgnn_tc_gemm(const hgnn_tc_row_layer* d, int64_t rows, float* out, int64_t ld_out, int64_t col0, void* a_img,
                            void* stream) {
  HGNN_REQUIRE(d != nullptr, "tc_gemm: descriptor is NULL");
  if (!hgnn_tc_gemm_supported(d))
    return fail(HGNN_ERR_UNSUPPORTED, "tc_gemm: needs 1..3 segments of width %% 64 == 0, fan-in <= 768, n_out in {64, 128, 256}");
  if (rows <= 0) return HGNN_OK;
  HGNN_REQUIRE(out && d->w_packed, "tc_gemm: NULL pointer");
  HGNN_REQUIRE(rows < INT32_MAX && ld_out >= col0 + d->n_out && col0 >= 0 && col0 % 4 == 0 && ld_out % 4 == 0,
               "tc_gemm: bad output geometry (ld_out %lld, col0 %lld)", (long long)ld_out, (long long)col0);
  RowFwdArgs a{};
  a.n_seg = d->n_seg;
  for (int s = 0; s < d->n_seg; ++s) {
    HGNN_REQUIRE(d->seg_ptr[s] != nullptr, "tc_gemm: segment %d is NULL", s);
    a.seg_ptr[s] = d->seg_ptr[s]; a.seg_idx[s] = d->seg_idx[s]; a.seg_width[s] = d->seg_width[s];
  }
  a.nkb = layer_k(d) / KBLK;
  a.w_packed = (const uint8_t*)d->w_packed;
  a.bias = d->bias; a.gamma = nullptr; a.beta = nullptr; a.eps = 0.f; a.skip = nullptr;
  a.out = out; a.ld_out = ld_out; a.col0 = (int)col0; a.a_img = (uint8_t*)a_img; a.rows = rows;
  cudaStream_t st = (cudaStream_t)stream;
  if (d->n_out == 64) return launch_fwd<64, HGNN_ACT_NONE, true>(a, st);
  if (d->n_out == 128) return launch_fwd<128, HGNN_ACT_NONE, true>(a, st);
  return launch_fwd<256, HGNN_ACT_NONE, true>(a, st);
}

extern "C" size_t hgnn_tc_row_backward_workspace_bytes(int64_t rows, int64_t k, int64_t n_out) {
  return bwd_layout(rows > 0 ? rows : 1, (int)k, (int)n_out).total + 1024;
}

namespace {

// d_seg == nullptr: d(input) goes to d_in as one [rows, K] matrix; else segment s goes to d_seg[s] as a dense
// [rows, seg_width[s]] matrix (NULL entry = that segment's gradient is not wanted: its columns are never stored)
int row_backward(const hgnn_tc_row_layer* d, const void* wt_packed, const void* a_img, int64_t rows, const float* grad_out,
                 float* d_in, float* const* d_seg, float* dW, float* dvec, void* ws, size_t ws_bytes, void* stream, const char* who) {
  int rc = validate(d, who);
  if (rc) return rc;
  const int K = layer_k(d), N = d->n_out;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0) {
    if (dW) HGNN_CUDA_TRY(cudaMemsetAsync(dW, 0, (size_t)N * K * 4, st));
    if (dvec) HGNN_CUDA_TRY(cudaMemsetAsync(dvec, 0, (size_t)3 * N * 4, st));
    return HGNN_OK;
  }
  HGNN_REQUIRE(wt_packed && a_img && grad_out && (d_in || d_seg) && dW && dvec && ws && d->w_packed && d->bias && d->gamma && d->beta,
               "%s: NULL pointer", who);
  HGNN_REQUIRE(rows < INT32_MAX, "%s: too many rows", who);
  BwdLayout Y = bwd_layout(rows, K, N);
  uintptr_t base = align_up((uintptr_t)ws, 1024);
  if (ws_bytes < (base - (uintptr_t)ws) + Y.total) return fail(HGNN_ERR_WORKSPACE, "%s: workspace too small", who);
  uint8_t* w = (uint8_t*)base;
  RowBwdArgs a{};
  a.nkb = K / KBLK; a.K = K;
  a.w_packed = (const uint8_t*)d->w_packed; a.wt_packed = (const uint8_t*)wt_packed;
  a.bias = d->bias; a.gamma = d->gamma; a.beta = d->beta; a.eps = d->ln_eps;
  a.a_img = (const uint8_t*)a_img; a.gout = grad_out;
  if (d_seg == nullptr) {
    for (int p = 0; p < K / 128; ++p) { a.d_piece[p] = d_in + p * 128; a.d_ld[p] = K; }
  } else {
    int col = 0;
    for (int s = 0; s < d->n_seg; ++s) {
      const int wdt = d->seg_width[s];
      HGNN_REQUIRE(wdt % 128 == 0, "%s: per-segment gradients need segment widths that are multiples of 128 (segment %d is %d wide)", who, s, wdt);
      HGNN_REQUIRE(d_seg[s] == nullptr || ((uintptr_t)d_seg[s] % 16) == 0, "%s: segment gradient %d is not 16-byte aligned", who, s);
      for (int c = 0; c < wdt; c += 128) {
        a.d_piece[(col + c) / 128] = d_seg[s] ? d_seg[s] + c : nullptr;
        a.d_ld[(col + c) / 128] = wdt;
      }
      col += wdt;
    }
  }
  a.d_img = w + Y.d_img; a.colpart = (float*)(w + Y.colpart); a.rows = rows;
  rc = ROW_DISPATCH(launch_bwd, N, d->act, a, Y.grid, st);
  if (rc) return rc;
  // dW[:, 128 p .. 128 p + 128) = delta^T A[:, same columns]; the launch that sums the split-K partials in order also sums
  // the per-CTA column-sum partials (d bias | d gamma | d beta)
  WgradProblem pr[3];
  for (int p = 0; p < Y.n_prob; ++p)
    pr[p] = WgradProblem{a.d_img, N, 0, N, a.a_img, K, p * 128, 128, dW, K, 0, p * 128, 0};
  ColumnSums cs{a.colpart, Y.grid * 4, 3 * N, dvec};
  return launch_wgrad(pr, Y.n_prob, Y.tiles, w + Y.wgrad, Y.wgrad_bytes, st, &cs);
}

}  // namespace

extern "C" int hgnn_tc_row_backward(const hgnn_tc_row_layer* d, const void* wt_packed, const void* a_img, int64_t rows,
                                    const float* grad_out, float* d_in, float* dW, float* dvec, void* ws, size_t ws_bytes,
                                    void* stream) {
  return row_backward(d, wt_packed, a_img, rows, grad_out, d_in, nullptr, dW, dvec, ws, ws_bytes, stream, "tc_row_backward");
}

extern "C" int hgnn_tc_row_backward_split(const hgnn_tc_row_layer* d, const void* wt_packed, const void* a_img, int64_t rows,
                                          const float* grad_out, float* const* d_seg, float* dW, float* dvec, void* ws,
                                          size_t ws_bytes, void* stream) {
  HGNN_REQUIRE(d_seg != nullptr, "tc_row_backward_split: d_seg is NULL");
  return row_backward(d, wt_packed, a_img, rows, grad_out, nullptr, d_seg, dW, dvec, ws, ws_bytes, stream, "tc_row_backward_split");
}
