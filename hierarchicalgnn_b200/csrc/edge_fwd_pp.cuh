// Persistent, warp-specialised, two-tiles-in-flight ("ping-pong") forward kernel of the tensor-core edge step
// (included by edge_tc.cu).
//
//   e'[i] = Tanh(LN2(W2 . GELU(LN1(W1 . [x[src_i] | x[dst_i] | e_i] + b1)) + b2)) + e_i
//   (InteractionGNNCell.edge_update, gnn_utils.py:56-64; make_mlp layout utils.py:183-196)
//
// Same mathematics, same HBM inputs / outputs / stash as k_tc_edge_fwd. One 768-thread CTA per SM walks its tiles (128 edges
// each); the CUDA-core work of a tile (two LayerNorm + activation epilogues, the skip / store pass, the fused segmented
// reduce) is a serial chain of ~30 k cycles that no single group of warps can overlap with itself, so TWO epilogue groups
// each own every other tile and run half a period apart: one group's TMEM reads / special-function math fills the other's
// global-load and barrier waits, and both hide the gathers and MMAs, which belong to other warps altogether:
//
//   warp 0       row ids of the coming tiles (edge id, src, dst, run-end flags) into a 4-deep ring
//   warp 1       one lane schedules and issues every tcgen05.mma and streams the weight pieces they read (16 KB = 128 output
//                rows x one 64-wide K-block, bulk copies through a 3-deep ring). G1 = [x|x|e] . W1^T goes into the tile's
//                group's TMEM half as two N = 128 accumulators, G2 = A2 . W2^T into the first L columns of the same half
//                after the group has read D1. The order is decided piece by piece: a G2 whose A2 image is complete goes
//                first (it is short and an epilogue group is waiting for it), otherwise the open G1 continues, otherwise
//                the next tile's G1 opens once its group has released the TMEM half
//   warp 3       one lane bulk-stores the finished hidden images g (stash, training only)
//   warps 4-7    gather into a 2-deep ring of swizzled A0 K-blocks, in the order e0 xs0 e1 xs1 xd0 xd1: fp32 e rows through
//                registers (a whole K-block of loads in flight, converted to bf16), bf16 node rows (shadow copy) by
//                cp.async straight into the swizzled block — the e loads travel while the node blocks are copied
//   warps 8-15   epilogue group 0 (tiles 0, 2, 4, ... of this CTA)
//   warps 16-23  epilogue group 1 (tiles 1, 3, 5, ...)
//
// Shared memory (L = 128): 2 x 64 KB A2 image / fp32 staging tile (one per group) | A0 ring 2 x 16 KB | weight ring
// 3 x 16 KB | parameters, id ring, LayerNorm exchange, barriers = 225 KB. TMEM: 2 x 256 columns.
#pragma once

namespace pp {

constexpr int PP_THREADS = 768;
constexpr int GAT_WARP0 = 4, GAT_THREADS = 128;
constexpr int EPI_WARP0 = 8, EPI_THREADS = 256;  // per group
constexpr int NIDS = 4;                          // id-ring depth
constexpr int PIECE = 16384;                     // weight-ring slot

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// barrier among the 256 threads of one epilogue group (hardware named barriers 1 and 2)
__device__ __forceinline__ void grp_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(EPI_THREADS) : "memory"); }
// one arrive per warp: every lane's writes (and proxy fences) are ordered before it by the warp barrier
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// -DHGNN_TRACE: CTA 0 appends {clock, role << 16 | event << 8 | group, tile iteration} records to a device buffer
// (hgnn_tc_debug_trace in edge_tc.cu); profiles/pp_trace.py prints the timeline
#ifdef HGNN_TRACE
static __device__ unsigned long long g_trace[3 * 8192];
static __device__ unsigned int g_trace_n;
__device__ __forceinline__ void trace_ev(int role, int ev, int g, int it) {
  if (blockIdx.x != 0) return;
  const unsigned int i = atomicAdd(&g_trace_n, 1u);
  if (i < 8192) { g_trace[3 * i] = (unsigned long long)clock64(); g_trace[3 * i + 1] = (unsigned long long)((role << 16) | (ev << 8) | g); g_trace[3 * i + 2] = (unsigned long long)it; }
}
#define PP_TRACE(role, ev, g, it) trace_ev(role, ev, g, it)
#else
#define PP_TRACE(role, ev, g, it) ((void)0)
#endif

template <int L>
struct PpCfg {
  static constexpr int H = 2 * L;
  static constexpr int KPS = L / KBLK;
  static constexpr int NKB1 = 3 * KPS;
  static constexpr int NKB2 = H / KBLK;
  static constexpr int NH = H / 128;                 // N = 128 accumulator halves of GEMM1
  static constexpr int W1_BLK = H * ROW_BYTES;
  static constexpr int W2_BLK = L * ROW_BYTES;
  static constexpr int NA = 2, NW = 3;
  static constexpr int A2_BYTES = NKB2 * A_BLK_BYTES;
  static constexpr int SOUT_BYTES = TILE_M * L * 4;
  static constexpr int A2_REGION = A2_BYTES > SOUT_BYTES ? A2_BYTES : SOUT_BYTES;
  static constexpr int A2_OFF = 0;                               // two regions, one per group
  static constexpr int A0_OFF = A2_OFF + 2 * A2_REGION;
  static constexpr int W_OFF = A0_OFF + NA * A_BLK_BYTES;
  static constexpr int PAR_OFF = W_OFF + NW * PIECE;
  static constexpr int PARAM_FLOATS = 3 * H + 3 * L;
  static constexpr int IDS_OFF = PAR_OFF + PARAM_FLOATS * 4;      // NIDS x {eid, src, dst, flag} x 128 int
  static constexpr int RED_OFF = IDS_OFF + NIDS * 4 * TILE_M * 4;  // 2 groups x [128 rows][2 halves][2]
  static constexpr int BAR_OFF = RED_OFF + 2 * TILE_M * 4 * 4;
  // barrier indices
  static constexpr int IDS_FULL = 0, IDS_EMPTY = IDS_FULL + NIDS, A0_FULL = IDS_EMPTY + NIDS, A0_EMPTY = A0_FULL + NA,
                       W_FULL = A0_EMPTY + NA, W_EMPTY = W_FULL + NW, D1_FULL = W_EMPTY + NW, A2_FULL = D1_FULL + 2,
                       D2_FULL = A2_FULL + 2, TM_FREE = D2_FULL + 2, G_READ = TM_FREE + 2, NBAR = G_READ + 2;
  static constexpr int SMEM = BAR_OFF + NBAR * 8 + 16 + 16;  // barriers | TMEM base | piece FIFO
  static constexpr int TMEM_COLS = 2 * H;  // 512 (L = 128) / 256 (L = 64)
  static_assert(L == 128, "the ping-pong edge step is instantiated for latent 128 (block order e0 xs0 e1 xs1 xd0 xd1)");
  static_assert(W2_BLK <= PIECE && 128 * ROW_BYTES == PIECE, "weight pieces are one ring slot each");
  static_assert(SMEM <= 232448, "shared memory budget");
};

template <int L, int ACT_H, int ACT_O>
__global__ void __launch_bounds__(PP_THREADS, 1)
k_tc_edge_fwd_pp(hgnn_tc_edge_params P, const uint16_t* __restrict__ xb16, const float* __restrict__ e, const int32_t* __restrict__ src,
                 const int32_t* __restrict__ dst, const int32_t* __restrict__ perm, int64_t n_edges, float* __restrict__ e_out,
                 const int32_t* __restrict__ rowptr, float* __restrict__ agg, uint8_t* __restrict__ stash, EdgeStash SL) {
  using C = PpCfg<L>;
  constexpr int H = C::H, KPS = C::KPS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* const sm = smem_raw;
  float* s_par = reinterpret_cast<float*>(sm + C::PAR_OFF);
  float *s_b1 = s_par, *s_g1 = s_par + H, *s_be1 = s_par + 2 * H, *s_b2 = s_par + 3 * H, *s_g2 = s_par + 3 * H + L,
        *s_be2 = s_par + 3 * H + 2 * L;
  int* s_ids = reinterpret_cast<int*>(sm + C::IDS_OFF);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + C::BAR_OFF + C::NBAR * 8);
  const uint32_t sm_u = smem_u32(sm), bar0 = sm_u + C::BAR_OFF;
  if ((sm_u & 1023u) != 0) __trap();  // swizzled operand images need 1024 B alignment
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  auto IDS = [&](int buf, int which) { return s_ids + (buf * 4 + which) * TILE_M; };  // which: 0 eid, 1 src, 2 dst, 3 flag

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < NIDS; ++i) { mbar_init(BAR(C::IDS_FULL + i), 1); mbar_init(BAR(C::IDS_EMPTY + i), 12); }
    for (int i = 0; i < C::NA; ++i) { mbar_init(BAR(C::A0_FULL + i), 4); mbar_init(BAR(C::A0_EMPTY + i), 1); }
    for (int i = 0; i < C::NW; ++i) { mbar_init(BAR(C::W_FULL + i), 1); mbar_init(BAR(C::W_EMPTY + i), 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(BAR(C::D1_FULL + g), 1);
      mbar_init(BAR(C::A2_FULL + g), 8);
      mbar_init(BAR(C::D2_FULL + g), 1);
      mbar_init(BAR(C::TM_FREE + g), 8);
      mbar_init(BAR(C::G_READ + g), 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), C::TMEM_COLS);
  for (int i = tid; i < H; i += PP_THREADS) { s_b1[i] = P.b1[i]; s_g1[i] = P.gamma1[i]; s_be1[i] = P.beta1[i]; }
  for (int i = tid; i < L; i += PP_THREADS) { s_b2[i] = P.b2[i]; s_g2[i] = P.gamma2[i]; s_be2[i] = P.beta2[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const int n_tiles = (int)((n_edges + TILE_M - 1) / TILE_M);
  const int n_it = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA (>= 1: grid <= tiles)
  const uint8_t* w1p = reinterpret_cast<const uint8_t*>(P.w1_packed);
  const uint8_t* w2p = reinterpret_cast<const uint8_t*>(P.w2_packed);
  uint8_t* const a0_img = stash ? stash + SL.a0 : nullptr;
  uint8_t* const g_img = stash ? stash + SL.g : nullptr;
  uint4* const xh1_st = stash ? reinterpret_cast<uint4*>(stash + SL.xh1) : nullptr;
  uint4* const xh2_st = stash ? reinterpret_cast<uint4*>(stash + SL.xh2) : nullptr;
  float* const rstd_st = stash ? reinterpret_cast<float*>(stash + SL.rstd) : nullptr;
  // K-blocks of the concatenation [x[src] | x[dst] | e] are visited in the order e0 xs0 e1 xs1 xd0 xd1: the fp32 e rows come
  // from HBM through registers, the node rows from the bf16 shadow copy (L2) by cp.async; interleaved, one hides the other
  auto blk_seg = [&](int i) { return i < 4 ? ((i & 1) ? 0 : 2) : 1; };      // 2 = e, 0 = x[src], 1 = x[dst]
  auto blk_sub = [&](int i) { return i < 4 ? (i >> 1) : (i - 4); };         // K-block inside the segment
  auto blk_kb = [&](int i) { return blk_seg(i) * KPS + blk_sub(i); };       // K-block index in the concatenation / W1 image order
  auto tile_of = [&](int it) { return (int)blockIdx.x + it * (int)gridDim.x; };

  // Register budget (61 440 allocated at launch: 768 x 80): the control warps keep 32, the gather warps take 96, the
  // epilogue warps 88. Each role's code sits inside the branch that sets its budget (ptxas allocates per region).
  const int wg = warp >> 2;
  if (wg == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
      // ======================= row ids (whole warp) =======================
      for (int it = 0; it < n_it; ++it) {
        const int tile = tile_of(it);
        const int buf = it % NIDS, k = it / NIDS;
        if (k >= 1) mbar_wait(BAR(C::IDS_EMPTY + buf), (k - 1) & 1);
        int *s_eid = IDS(buf, 0), *s_src = IDS(buf, 1), *s_dst = IDS(buf, 2), *s_flag = IDS(buf, 3);
#pragma unroll 1
        for (int rr = 0; rr < TILE_M / 32; ++rr) {
          const int r = rr * 32 + lane;
          const int64_t j0 = (int64_t)tile * TILE_M + r;           // true sorted position (may be >= n_edges: padding)
          const int64_t j = j0 < n_edges ? j0 : n_edges - 1;       // padding rows recompute the last edge; never stored
          const int eid = perm ? perm[j] : (int)j;
          s_eid[r] = eid;
          s_src[r] = src[eid];
          const int d = dst[eid];
          s_dst[r] = d;
          int flag = 0;
          if (agg) {  // rows arrive destination-sorted: classify run ends (1 = store the sum, 2 = run continues elsewhere)
            constexpr int G = TILE_M / (EPI_THREADS / L);
            const int64_t gbeg = j0 / G * G, gend = gbeg + G < n_edges ? gbeg + G : n_edges;
            if (j0 < n_edges) {
              const bool last = (j0 + 1 >= gend) || (dst[perm ? perm[j0 + 1] : (int)(j0 + 1)] != d);
              if (last) flag = (rowptr[d] >= gbeg && rowptr[d + 1] <= gend) ? 1 : 2;
            } else {
              flag = 2;
            }
          }
          s_flag[r] = flag;
        }
        warp_arrive(BAR(C::IDS_FULL + buf), lane);
      }
    } else if (warp == 2) {
      // ======================= scheduler: decides the order of the weight pieces (= of the MMAs) and streams them =======================
      if (lane == 0) {
        uint32_t* const s_code = reinterpret_cast<uint32_t*>(sm + C::BAR_OFF + C::NBAR * 8 + 16);  // what each ring slot holds
        constexpr int NP1 = C::NKB1 * C::NH, NP2 = C::NKB2;
        uint32_t n_sched = 0;
        int g1_it = 0, g1_p = 0, g2_it = 0, g2_p = 0;  // next piece of the G1 / G2 being scheduled (p == 0: not open)
        while (g1_it < n_it || g2_it < n_it) {
          uint32_t code, bytes;
          const uint8_t* srcp;
          bool g2 = g2_p > 0;
          if (!g2 && g2_it < g1_it) g2 = mbar_test(BAR(C::A2_FULL + (g2_it & 1)), (g2_it >> 1) & 1);
          if (g2) {
            code = ((uint32_t)g2_it << 8) | 0x80u | (uint32_t)g2_p;
            srcp = w2p + (size_t)g2_p * C::W2_BLK;
            bytes = C::W2_BLK;
            if (++g2_p == NP2) { g2_p = 0; ++g2_it; }
          } else {
            if (g1_p == 0) {
              const int n = g1_it >> 1;
              const bool can_open = g1_it < n_it && (n == 0 || mbar_test(BAR(C::TM_FREE + (g1_it & 1)), (n - 1) & 1));
              if (!can_open) {  // nothing to schedule now: park on the event that can come next, then look again
                if (g2_it < g1_it) mbar_try_wait_hint(BAR(C::A2_FULL + (g2_it & 1)), (g2_it >> 1) & 1, 2000u);
                else mbar_try_wait_hint(BAR(C::TM_FREE + (g1_it & 1)), (n - 1) & 1, 2000u);
                continue;
              }
            }
            code = ((uint32_t)g1_it << 8) | (uint32_t)g1_p;
            srcp = w1p + (size_t)blk_kb(g1_p / C::NH) * C::W1_BLK + (size_t)(g1_p % C::NH) * PIECE;
            bytes = PIECE;
            if (++g1_p == NP1) { g1_p = 0; ++g1_it; }
          }
          const uint32_t sl = n_sched % C::NW, k = n_sched / C::NW;
          if (k >= 1) mbar_wait(BAR(C::W_EMPTY + sl), (k - 1) & 1);
          s_code[sl] = code;  // ordered before the arrive below (release), read by the MMA thread after its wait (acquire)
#ifdef PP_EXPERIMENT_NO_W  // timing experiment (wrong results): 16 bytes per piece instead of 16 KB
          bytes = 16;
#endif
          mbar_expect_tx(BAR(C::W_FULL + sl), bytes);
          bulk_g2s(sm_u + C::W_OFF + sl * PIECE, srcp, bytes, BAR(C::W_FULL + sl));
          ++n_sched;
        }
      }
    } else if (warp == 1) {
      // ======================= MMA issuer: follows the ring =======================
      if (lane == 0) {
        const uint32_t idesc = make_idesc(TILE_M, 128), idesc2 = make_idesc(TILE_M, L);
        const volatile uint32_t* const s_code = reinterpret_cast<const volatile uint32_t*>(sm + C::BAR_OFF + C::NBAR * 8 + 16);
        constexpr int NP1 = C::NKB1 * C::NH, NP2 = C::NKB2;
        const uint32_t n_pieces = (uint32_t)n_it * (NP1 + NP2);
        uint32_t ca = 0;
        for (uint32_t n = 0; n < n_pieces; ++n) {
          const uint32_t sl = n % C::NW;
          mbar_wait(BAR(C::W_FULL + sl), (n / C::NW) & 1);
          const uint32_t code = s_code[sl];
          const int it = (int)(code >> 8), g = it & 1, pc = (int)(code & 0x7fu);
          const uint32_t td = tmem + g * H;
          if (code & 0x80u) {
            if (pc == 0) PP_TRACE(1, 3, g, it);
            tc_fence_after();
            umma_kblock(td, sm_u + C::A2_OFF + g * C::A2_REGION + pc * A_BLK_BYTES, sm_u + C::W_OFF + sl * PIECE, idesc2, pc == 0);
            umma_commit(BAR(C::W_EMPTY + sl));
            if (pc == NP2 - 1) { umma_commit(BAR(C::D2_FULL + g)); PP_TRACE(1, 4, g, it); }
          } else {
            const int i = pc / C::NH, h = pc % C::NH;
            const uint32_t sa = ca % C::NA;
            if (pc == 0) PP_TRACE(1, 1, g, it);
            if (h == 0) mbar_wait(BAR(C::A0_FULL + sa), (ca / C::NA) & 1);
            tc_fence_after();
            umma_kblock(td + h * 128, sm_u + C::A0_OFF + sa * A_BLK_BYTES, sm_u + C::W_OFF + sl * PIECE, idesc, i == 0);
            umma_commit(BAR(C::W_EMPTY + sl));
            if (h == C::NH - 1) { umma_commit(BAR(C::A0_EMPTY + sa)); ++ca; }
            if (pc == NP1 - 1) { umma_commit(BAR(C::D1_FULL + g)); PP_TRACE(1, 2, g, it); }
          }
        }
      }
    } else if (warp == 3) {
      // ======================= stash: the finished A2 image is the weight-gradient operand of layer 2 =======================
      if (lane == 0 && g_img) {
        for (int it = 0; it < n_it; ++it) {
          const int g = it & 1, n = it >> 1;
          mbar_wait(BAR(C::A2_FULL + g), n & 1);
          bulk_s2g(g_img + (size_t)tile_of(it) * C::A2_BYTES, sm_u + C::A2_OFF + g * C::A2_REGION, C::A2_BYTES);
          bulk_commit();
          bulk_wait_read0();           // the image has left shared memory: the region may become the fp32 staging tile
          mbar_arrive(BAR(C::G_READ + g));
        }
        bulk_wait0();
      }
    }
  } else if (wg == 1) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    // ======================= gather: A0 K-blocks =======================
    // fp32 rows: 16 threads per 256 B row piece, 8 rows per pass, 16 passes (a whole K-block in flight in registers);
    // bf16 rows: 8 threads per 128 B piece, 16 rows per pass, 8 cp.async of 16 B per thread
    const int gt = tid - GAT_WARP0 * 32;  // 0..127
    const uint64_t pol_keep = l2_policy_evict_last();
    const int e_sub = gt & 15, e_rr = gt >> 4, x_c16 = gt & 7, x_rr = gt >> 3;
    float4 ev[16];
    auto load_e = [&](const int* s_eid, int col0) {
#pragma unroll
      for (int p = 0; p < 16; ++p)
        ev[p] = ldg_f4_hint(reinterpret_cast<const float4*>(e + (size_t)s_eid[p * 8 + e_rr] * L + col0) + e_sub, pol_keep);
    };
    auto store_e = [&](uint8_t* blk, uint8_t* gimg, bool pad_tile, int64_t row0) {
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const int r = p * 8 + e_rr;
        uint2 pk = make_uint2(pack_bf16(ev[p].x, ev[p].y), pack_bf16(ev[p].z, ev[p].w));
        if (pad_tile && row0 + r >= n_edges) pk = make_uint2(0u, 0u);  // padding rows: zeros (also in the saved image)
        const uint32_t off = sw128_off(r, e_sub >> 1) + (e_sub & 1) * 8;
        *reinterpret_cast<uint2*>(blk + off) = pk;
        if (gimg) *reinterpret_cast<uint2*>(gimg + off) = pk;
      }
    };
    auto copy_x = [&](uint32_t blk_u, const int* rows, int col0, bool pad_tile, int64_t row0) {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int r = p * 16 + x_rr;
        const uint32_t nbytes = (pad_tile && row0 + r >= n_edges) ? 0u : 16u;  // padding rows: zero fill
        cp_async16(blk_u + sw128_off(r, x_c16), reinterpret_cast<const uint4*>(xb16 + (size_t)rows[r] * L + col0) + x_c16, nbytes);
      }
      cp_async_commit();
    };
    uint32_t ca = 0;
    auto slot_acquire = [&]() -> uint32_t {  // the ring slot of the next block, free to be written
      const uint32_t sa = ca % C::NA, k = ca / C::NA;
      if (k >= 1) mbar_wait(BAR(C::A0_EMPTY + sa), (k - 1) & 1);
      return sa;
    };
    auto publish = [&](uint32_t sa) {  // this thread's writes of the block (generic proxy or completed cp.async) -> async proxy
      fence_proxy_async();
      warp_arrive(BAR(C::A0_FULL + sa), lane);
    };
    mbar_wait(BAR(C::IDS_FULL + 0), 0);
    load_e(IDS(0, 0), 0);
    for (int it = 0; it < n_it; ++it) {
      const int tile = tile_of(it);
      const int ibuf = it % NIDS;
      const int *s_eid = IDS(ibuf, 0), *s_src = IDS(ibuf, 1), *s_dst = IDS(ibuf, 2);
      const bool pad_tile = (int64_t)(tile + 1) * TILE_M > n_edges;
      const int64_t row0 = (int64_t)tile * TILE_M;
      uint8_t* const img = a0_img ? a0_img + (size_t)tile * KPS * A_BLK_BYTES : nullptr;
      // e0 (its loads were issued one tile ago)
      if (gt == 0) PP_TRACE(2, 1, 0, it);
      uint32_t s0 = slot_acquire(); ++ca;
      if (gt == 0) PP_TRACE(2, 10, 0, it);
      store_e(sm + C::A0_OFF + s0 * A_BLK_BYTES, img, pad_tile, row0);
      publish(s0);
      if (gt == 0) PP_TRACE(2, 20, 0, it);
      load_e(s_eid, KBLK);  // e1 travels under xs0
      // xs0
      uint32_t s1 = slot_acquire(); ++ca;
      if (gt == 0) PP_TRACE(2, 11, 0, it);
      copy_x(sm_u + C::A0_OFF + s1 * A_BLK_BYTES, s_src, 0, pad_tile, row0);
      cp_async_wait<0>();
      publish(s1);
      if (gt == 0) PP_TRACE(2, 21, 0, it);
      // e1
      s0 = slot_acquire(); ++ca;
      if (gt == 0) PP_TRACE(2, 12, 0, it);
      store_e(sm + C::A0_OFF + s0 * A_BLK_BYTES, img ? img + A_BLK_BYTES : nullptr, pad_tile, row0);
      publish(s0);
      if (gt == 0) PP_TRACE(2, 22, 0, it);
      if (it + 1 < n_it) {  // the next tile's e0 travels under xs1 xd0 xd1
        const int nbuf = (it + 1) % NIDS;
        mbar_wait(BAR(C::IDS_FULL + nbuf), ((it + 1) / NIDS) & 1);
        load_e(IDS(nbuf, 0), 0);
      }
      // xs1, xd0 back to back (two copies in flight), then xd1 behind xs1
      s1 = slot_acquire(); ++ca;
      if (gt == 0) PP_TRACE(2, 13, 0, it);
      copy_x(sm_u + C::A0_OFF + s1 * A_BLK_BYTES, s_src, KBLK, pad_tile, row0);
      s0 = slot_acquire(); ++ca;
      if (gt == 0) PP_TRACE(2, 14, 0, it);
      copy_x(sm_u + C::A0_OFF + s0 * A_BLK_BYTES, s_dst, 0, pad_tile, row0);
      cp_async_wait<1>();
      publish(s1);
      if (gt == 0) PP_TRACE(2, 23, 0, it);
      s1 = slot_acquire(); ++ca;
      if (gt == 0) PP_TRACE(2, 15, 0, it);
      copy_x(sm_u + C::A0_OFF + s1 * A_BLK_BYTES, s_dst, KBLK, pad_tile, row0);
      cp_async_wait<1>();
      publish(s0);
      if (gt == 0) PP_TRACE(2, 24, 0, it);
      cp_async_wait<0>();
      publish(s1);
      if (gt == 0) PP_TRACE(2, 2, 0, it);
      warp_arrive(BAR(C::IDS_EMPTY + ibuf), lane);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    // ======================= epilogue groups (warps 8..15 and 16..23) =======================
    const int grp = (warp - EPI_WARP0) >> 3;
    const int et = tid - (EPI_WARP0 + 8 * grp) * 32;  // 0..255 inside the group
    const int ew = et >> 5;                           // epilogue warp 0..7
    const int q = warp & 3, hsel = ew >> 2;           // TMEM lane quarter (= warp id % 4), column half
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + grp * H + ((uint32_t)(q * 32) << 16);
    float* const s_red = reinterpret_cast<float*>(sm + C::RED_OFF) + grp * TILE_M * 4;
    uint8_t* const region = sm + C::A2_OFF + grp * C::A2_REGION;  // A2 image, then fp32 staging tile
    const uint64_t pol_drop = l2_policy_evict_first();
    for (int it = grp; it < n_it; it += 2) {
      const int tile = tile_of(it);
      const int ibuf = it % NIDS;
      const int *s_eid = IDS(ibuf, 0), *s_dst = IDS(ibuf, 2), *s_flag = IDS(ibuf, 3);
      const uint32_t par = (it >> 1) & 1;
      mbar_wait(BAR(C::IDS_FULL + ibuf), (it / NIDS) & 1);  // (long complete: the gather warps needed it first) acquire the ids
      // ---- EPI1: bias + LayerNorm + activation -> bf16 A2 (K-major, swizzled) ----
      if (et == 0) PP_TRACE(3, 0, grp, it);
      mbar_wait(BAR(C::D1_FULL + grp), par);
      tc_fence_after();
      if (et == 0) PP_TRACE(3, 1, grp, it);
      {
        constexpr int NC = H / 2;
        const int c0 = hsel * NC;
        float mloc, m2;
        ln_partial<NC / 32>(t_lane + c0, s_b1 + c0, mloc, m2);
        s_red[row * 4 + hsel * 2] = mloc;
        s_red[row * 4 + hsel * 2 + 1] = m2;
        grp_sync(grp);  // also: every thread of the group has left the previous tile's aggregate -> the region is free
        const LnStat st = combine_halves(s_red, row, NC, P.ln_eps);
        if (rstd_st && hsel == 0) rstd_st[(size_t)tile * 2 * TILE_M + row] = st.rstd;
        ln_act_to_image<ACT_H, NC / 32>(t_lane + c0, s_b1, s_g1, s_be1, c0, st.mean, st.rstd, region, row,
                                        xh1_st ? xh1_st + (size_t)tile * (H / 8) * TILE_M : nullptr);
      }
      fence_proxy_async();
      tc_fence_before();
      warp_arrive(BAR(C::A2_FULL + grp), lane);
      if (et == 0) PP_TRACE(3, 2, grp, it);
      // ---- EPI2: bias + LayerNorm + activation -> fp32 staging tile (swizzled 16 B chunks) over the A2 region ----
      mbar_wait(BAR(C::D2_FULL + grp), par);  // every MMA reading A2 has retired
      tc_fence_after();
      if (et == 0) PP_TRACE(3, 3, grp, it);
      {
        constexpr int NC = L / 2, NCH = NC / 32;
        const int c0 = hsel * NC;
        float mloc, m2;
        ln_partial<NCH>(t_lane + c0, s_b2 + c0, mloc, m2);
        s_red[row * 4 + hsel * 2] = mloc;  // EPI1's readers of s_red are past: D2_FULL needs every warp's A2_FULL arrival
        s_red[row * 4 + hsel * 2 + 1] = m2;
        grp_sync(grp);
        const LnStat st = combine_halves(s_red, row, NC, P.ln_eps);
        const float2 rs2 = splat2(st.rstd), nmr2 = splat2(-st.mean * st.rstd);
        if (rstd_st && hsel == 0) rstd_st[(size_t)tile * 2 * TILE_M + TILE_M + row] = st.rstd;
        uint4* const xh2_t = xh2_st ? xh2_st + (size_t)tile * (L / 8) * TILE_M : nullptr;
        if (g_img) mbar_wait(BAR(C::G_READ + grp), par);  // the g image has left shared memory
        float v[32];
#pragma unroll 1
        for (int ch = 0; ch < NCH; ++ch) {
          tmem_ld32(t_lane + c0 + ch * 32, v);
          if (ch == NCH - 1) {  // last read of D2 by this warp: GEMM1 of the group's next tile may overwrite the TMEM half
            tc_fence_before();
            warp_arrive(BAR(C::TM_FREE + grp), lane);
          }
          const int cb = c0 + ch * 32;
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            float2 xh[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c = cb + g8 * 8 + 4 * h;
              const float4 b = *reinterpret_cast<const float4*>(s_b2 + c);
              const float4 g = *reinterpret_cast<const float4*>(s_g2 + c);
              const float4 be = *reinterpret_cast<const float4*>(s_be2 + c);
              xh[2 * h] = fma2(add2(make_float2(v[g8 * 8 + 4 * h], v[g8 * 8 + 4 * h + 1]), make_float2(b.x, b.y)), rs2, nmr2);
              xh[2 * h + 1] = fma2(add2(make_float2(v[g8 * 8 + 4 * h + 2], v[g8 * 8 + 4 * h + 3]), make_float2(b.z, b.w)), rs2, nmr2);
              const float2 o0 = tc_act2<ACT_O>(fma2(xh[2 * h], make_float2(g.x, g.y), make_float2(be.x, be.y)));
              const float2 o1 = tc_act2<ACT_O>(fma2(xh[2 * h + 1], make_float2(g.z, g.w), make_float2(be.z, be.w)));
              *reinterpret_cast<float4*>(region + (size_t)row * (L * 4) + (((c >> 2) ^ (row & 7)) << 4)) = make_float4(o0.x, o0.y, o1.x, o1.y);
            }
            if (xh2_t)
              xh2_t[(size_t)((cb + g8 * 8) >> 3) * TILE_M + row] = make_uint4(pack_bf16(xh[0]), pack_bf16(xh[1]), pack_bf16(xh[2]), pack_bf16(xh[3]));
          }
        }
      }
      grp_sync(grp);
      if (et == 0) PP_TRACE(3, 4, grp, it);
      // ---- coalesced pass: + fp32 skip row, full-row stores; skip rows fetched eight at a time before anything depends on them ----
      {
        constexpr int CPR = L / 4;                       // float4 chunks per row
        constexpr int ROWS_PER_WARP = TILE_M / (EPI_THREADS / 32);
        constexpr int ITERS = ROWS_PER_WARP * CPR / 32;
        constexpr int BATCH = 8;
        static_assert(ITERS % BATCH == 0, "store pass batches");
#pragma unroll 1
        for (int b0 = 0; b0 < ITERS; b0 += BATCH) {
          float4 sk[BATCH];
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {  // padding rows carry the (valid) id of the last edge: load unconditionally
            const int idx = lane + (b0 + u) * 32;
            const int r = ew * ROWS_PER_WARP + idx / CPR, c4 = idx % CPR;
            sk[u] = ldg_f4_hint(reinterpret_cast<const float4*>(e + (size_t)s_eid[r] * L + c4 * 4), pol_drop);
          }
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            const int idx = lane + (b0 + u) * 32;
            const int r = ew * ROWS_PER_WARP + idx / CPR, c4 = idx % CPR;
            float4* sp = reinterpret_cast<float4*>(region + (size_t)r * (L * 4) + ((c4 ^ (r & 7)) << 4));
            const float4 y = *sp;
            const float4 o = make_float4(y.x + sk[u].x, y.y + sk[u].y, y.z + sk[u].z, y.w + sk[u].w);
            if ((int64_t)tile * TILE_M + r < n_edges) {
              *reinterpret_cast<float4*>(e_out + (size_t)s_eid[r] * L + c4 * 4) = o;
              if (agg) *sp = o;
            }
          }
        }
        if (agg) {
          // ---- fused scatter_add: destination-sorted segmented reduce of the finished tile, ordered, no atomics.
          // One thread per (row group, column): runs of equal destination are summed in row order; a run that lies
          // inside the group is stored, runs crossing a group boundary are left to the fix-up pass.
          grp_sync(grp);
          if (et == 0) PP_TRACE(3, 5, grp, it);
          constexpr int G = TILE_M / (EPI_THREADS / L);
          const int rg = et / L, col = et % L;
          uint64_t endm = 0, storem = 0;
#pragma unroll
          for (int w = 0; w < G / 32; ++w) {
            const int f = s_flag[rg * G + w * 32 + lane];
            endm |= (uint64_t)__ballot_sync(0xffffffffu, f != 0) << (32 * w);
            storem |= (uint64_t)__ballot_sync(0xffffffffu, f == 1) << (32 * w);
          }
          const uint8_t* colp = region + ((col & 3) << 2);
          float acc = 0.f;
#pragma unroll 8
          for (int i = 0; i < G; ++i) {
            const int r = rg * G + i;
            acc += *reinterpret_cast<const float*>(colp + (size_t)r * (L * 4) + ((((col >> 2) ^ (r & 7)) << 4)));
            if ((endm >> i) & 1) {
              if ((storem >> i) & 1) agg[(size_t)s_dst[r] * L + col] = acc;
              acc = 0.f;
            }
          }
        }
      }
      if (et == 0) PP_TRACE(3, 6, grp, it);
      warp_arrive(BAR(C::IDS_EMPTY + ibuf), lane);
      // the staging tile is released to the group's next A2 image by the grp_sync inside the next EPI1
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, C::TMEM_COLS);
}

}  // namespace pp
