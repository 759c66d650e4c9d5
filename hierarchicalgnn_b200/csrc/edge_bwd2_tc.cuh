// Backward-data kernel of the tensor-core edge step, two-CTAs-per-SM variant (included by edge_bwd_tc.cu).
//
// Same mathematics and the same HBM inputs / outputs as k_tc_edge_bwd, re-tiled so that TWO 256-thread CTAs share an SM:
// one CTA's memory-bound phases (gradient rows in, weight pieces, d(e) rows out) then run under the other CTA's
// issue-bound LayerNorm / activation adjoints, which a single 512-thread CTA executes strictly one after the other.
// What makes two CTAs fit:
//   * shared memory 104 KB: the delta2 image lives in the first half of the delta1 image's bytes (delta2 is dead — consumed
//     by GEMM3 and bulk-stored — before the second EPI-C pass writes delta1), and the weights stream through TWO 16 KB
//     slots (8 pieces per tile: W2^T as four [128 hidden rows x 64] pieces, W1c^T as four K-blocks);
//   * TMEM 256 columns: dG [128 x 256] with d(y1) parked in place; d(e)_mlp re-uses columns [0, 128) afterwards; EPI-B parks
//     d(y2) there before GEMM3 (so its second pass does not have to keep 64 + 64 values in registers);
//   * registers <= 128 at 256 threads: a thread owns half a row (64 latent / 128 hidden columns) and walks it in
//     32-column chunks; the stashed x-hat chunks are fetched just in time (16 registers) and re-read (L1 / L2) in pass 2.
// d bias1 is not accumulated here: hgnn_tc_edge_backward takes it from the per-node sums of delta1.
#pragma once

namespace v2 {

constexpr int NT2 = 256;
constexpr int NSLOT2 = 2;
constexpr int A2_OFF2 = 0;                               // delta1 image (64 KB) | gout -> delta2 image (first 32 KB) | fp32 d(e) staging
constexpr int SLOT_OFF2 = A2_BYTES;                      // two 16 KB weight slots
constexpr int PAR_OFF2 = SLOT_OFF2 + NSLOT2 * SEG_BLK;
constexpr int IDS_OFF2 = PAR_OFF2 + PAR_FLOATS * 4;
constexpr int RED_OFF2 = IDS_OFF2 + 2 * TILE_M * 4;      // [128 rows][2 halves][2]
constexpr int BAR_OFF2 = RED_OFF2 + TILE_M * 4 * 4;
constexpr int NBAR2 = 2 * NSLOT2 + 1;
constexpr int SMEM_BYTES2 = BAR_OFF2 + NBAR2 * 8 + 16;
constexpr int NPIECE2 = 8;                               // 0..3: W2^T (K-block g >> 1, hidden half g & 1); 4..7: W1c^T K-blocks

template <int ACT_H, int ACT_O>
__global__ void __launch_bounds__(NT2, 2) k_tc_edge_bwd2(BwdArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* const sm = smem_raw;
  float* s_par = reinterpret_cast<float*>(sm + PAR_OFF2);
  float *s_g1 = s_par + H, *s_be1 = s_par + 2 * H, *s_g2 = s_par + 3 * H + L, *s_be2 = s_par + 3 * H + 2 * L;
  int* s_eid = reinterpret_cast<int*>(sm + IDS_OFF2);
  int* s_dst = s_eid + TILE_M;
  float* s_red = reinterpret_cast<float*>(sm + RED_OFF2);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(sm + BAR_OFF2 + NBAR2 * 8);
  const uint32_t sm_u = smem_u32(sm), bar0 = sm_u + BAR_OFF2;
  if ((sm_u & 1023u) != 0) __trap();
  enum { B_FULL = 0, B_FREE = NSLOT2, ACC = 2 * NSLOT2 };
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hs = warp >> 2;  // TMEM lane quarter, column half
  const int row = q * 32 + lane;
  const hgnn_tc_edge_params& P = A.P;

  if (tid == 0) {
    for (int i = 0; i < NBAR2; ++i) mbar_init(BAR(i), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), 256);
  for (int i = tid; i < H; i += NT2) { s_par[i] = P.b1[i]; s_g1[i] = P.gamma1[i]; s_be1[i] = P.beta1[i]; }
  for (int i = tid; i < L; i += NT2) { s_par[3 * H + i] = P.b2[i]; s_g2[i] = P.gamma2[i]; s_be2[i] = P.beta2[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
  const uint32_t idesc_l = make_idesc(TILE_M, L);

  // the fp32 upstream-gradient rows are read twice per tile (bf16 image, then the skip path of d(e)): keep / release in L2
  const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  uint32_t acc_par = 0;
  int nx_eid = 0, nx_dst = 0;
  // every completion of a barrier is awaited exactly once, in order: thread 0 consumes B_FULL, thread 32 consumes B_FREE
  uint32_t nf[NSLOT2] = {0, 0}, nr[NSLOT2] = {0, 0};
  // per-lane column sums: lane c of warp (q, hs) owns hidden columns hs*128 + 32 ch + c (ch < 4), latent columns hs*64 + 32 ch + c (ch < 2)
  float acc_dg1[4] = {0.f, 0.f, 0.f, 0.f}, acc_dbe1[4] = {0.f, 0.f, 0.f, 0.f};
  float acc_db2[2] = {0.f, 0.f}, acc_dg2[2] = {0.f, 0.f}, acc_dbe2[2] = {0.f, 0.f};

  auto full_wait = [&](int slot) { mbar_wait(BAR(B_FULL + slot), nf[slot] & 1); nf[slot]++; tc_fence_after(); };  // thread 0
  auto free_wait = [&](int slot) { mbar_wait(BAR(B_FREE + slot), nr[slot] & 1); nr[slot]++; };                    // thread 32
  auto piece_src = [&](int g) -> const uint8_t* {
    if (g < 4) return A.w2t + (size_t)(g >> 1) * W2T_BLK + (size_t)(g & 1) * SEG_BLK;  // rows [128 (g&1), +128) of K-block g >> 1
    return A.w1t + (size_t)(g - 4) * W1T_BLK + (size_t)2 * SEG_BLK;                    // rows [2L, 3L) of K-block g - 4
  };
  auto fill = [&](int g) {  // thread 32: piece g -> slot g & 1
    const int slot = g & 1;
    mbar_expect_tx(BAR(B_FULL + slot), SEG_BLK);
    bulk_g2s(sm_u + SLOT_OFF2 + slot * SEG_BLK, piece_src(g), SEG_BLK, BAR(B_FULL + slot));
  };

  long long t_prev = clock64();
  auto MARK = [&](int ph) {
    if (A.phase_clk && tid == 0 && blockIdx.x == 0) { long long t = clock64(); atomicAdd(A.phase_clk + ph, (unsigned long long)(t - t_prev)); t_prev = t; }
  };
  const int n_tiles = (int)((A.n_edges + TILE_M - 1) / TILE_M);
  if (tid == 32 && (int)blockIdx.x < n_tiles) { fill(0); fill(1); }

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ================= tile setup: row ids, fetched one tile ahead into registers =================
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
    if (tid < TILE_M && has_next) {
      int64_t j = (int64_t)(tile + gridDim.x) * TILE_M + tid;
      if (j >= A.n_edges) j = A.n_edges - 1;
      nx_eid = A.perm ? A.perm[j] : (int)j;
    }
    MARK(0);
    // ================= LOAD: gout = grad_eout[i] + grad_agg[dst_i] -> bf16 image (32 threads per row, 8 rows per pass) ==========
    {
      const int g_sub = tid & 31, g_rr = tid >> 5;
#pragma unroll 1
      for (int p0 = 0; p0 < TILE_M / 8; p0 += 8) {
        float4 gv[8], ga[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int r = (p0 + p) * 8 + g_rr;
          const bool live = (int64_t)tile * TILE_M + r < A.n_edges;
          gv[p] = live ? ldg_f4_hint(reinterpret_cast<const float4*>(A.g_e + (size_t)s_eid[r] * L) + g_sub, pol_keep) : make_float4(0.f, 0.f, 0.f, 0.f);
          ga[p] = (live && A.g_agg) ? __ldg(reinterpret_cast<const float4*>(A.g_agg + (size_t)s_dst[r] * L) + g_sub)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) {  // zero rows for padding: their delta's vanish
          const int r = (p0 + p) * 8 + g_rr;
          const int c = g_sub * 4;
          *reinterpret_cast<uint2*>(sm + A2_OFF2 + (c / KBLK) * A_BLK_BYTES + sw128_off(r, (c % KBLK) >> 3) + ((c >> 2) & 1) * 8) =
              make_uint2(pack_bf16(gv[p].x + ga[p].x, gv[p].y + ga[p].y), pack_bf16(gv[p].z + ga[p].z, gv[p].w + ga[p].w));
        }
      }
    }
    const float rstd1 = __ldg(A.rstd + (size_t)tile * 2 * TILE_M + row);
    const float rstd2 = __ldg(A.rstd + (size_t)tile * 2 * TILE_M + TILE_M + row);
    const uint4* xh2 = A.xh2 + ((size_t)tile * (L / 8) + hs * 8) * TILE_M + row;  // this thread's 8 x 8 bf16 of xhat2
    uint4 xq2_nx[4];  // first chunk in flight across the barrier
#pragma unroll
    for (int j = 0; j < 4; ++j) xq2_nx[j] = __ldg(xh2 + (size_t)j * TILE_M);
    __syncthreads();
    if (tid < TILE_M && has_next) nx_dst = A.dst[nx_eid];
    MARK(1);

    // ================= EPI-B: d(y2) = gout * act'(y2) (parked in TMEM), LayerNorm-2 adjoint -> delta2 (in place over gout) =====
    {
      float2 s1v = make_float2(0.f, 0.f), s2v = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        const int cb = hs * 64 + ch * 32;
        uint4 xq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xq[j] = xq2_nx[j];
        if (ch == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) xq2_nx[j] = __ldg(xh2 + (size_t)(4 + j) * TILE_M);
        }
        float u[32], tmp[32];
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int c = cb + g8 * 8;
          const uint4 pk = *reinterpret_cast<const uint4*>(sm + A2_OFF2 + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3));
          float2 go[4], xh[4];
          unpack8_2(pk, go);
          unpack8_2(xq[g8], xh);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 g = *reinterpret_cast<const float4*>(s_g2 + c + 4 * h);
            const float4 be = *reinterpret_cast<const float4*>(s_be2 + c + 4 * h);
            const float2 gg[2] = {make_float2(g.x, g.y), make_float2(g.z, g.w)}, ee[2] = {make_float2(be.x, be.y), make_float2(be.z, be.w)};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int i = g8 * 8 + 4 * h + 2 * k;
              const float2 x = xh[2 * h + k];
              const float2 d = mul2(go[2 * h + k], tc_act_bwd2<ACT_O>(fma2(x, gg[k], ee[k])));
              const float2 gd = mul2(gg[k], d), dx = mul2(d, x);
              u[i] = d.x; u[i + 1] = d.y;
              tmp[i] = dx.x; tmp[i + 1] = dx.y;
              s1v = add2(s1v, gd);
              s2v = fma2(gd, x, s2v);
            }
          }
        }
        tmem_st32(t_lane + cb, u);  // park d(y2): TMEM is idle until GEMM3
        acc_dg2[ch] += warp_colsum32(tmp, lane);
        acc_dbe2[ch] += warp_colsum32(u, lane);
      }
      s_red[row * 4 + hs * 2] = s1v.x + s1v.y;
      s_red[row * 4 + hs * 2 + 1] = s2v.x + s2v.y;
      __syncthreads();
      const float t1 = (s_red[row * 4] + s_red[row * 4 + 2]) * (1.0f / L);
      const float t2 = (s_red[row * 4 + 1] + s_red[row * 4 + 3]) * (1.0f / L);
      const float2 rs = splat2(rstd2), nt1r = splat2(-t1 * rstd2), nt2r = splat2(-t2 * rstd2);
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        const int cb = hs * 64 + ch * 32;
        uint4 xq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xq[j] = __ldg(xh2 + (size_t)(ch * 4 + j) * TILE_M);
        float u[32];
        tmem_ld32(t_lane + cb, u);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int c = cb + g8 * 8;
          float2 xh[4];
          unpack8_2(xq[g8], xh);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 g = *reinterpret_cast<const float4*>(s_g2 + c + 4 * h);
            const float2 gg[2] = {make_float2(g.x, g.y), make_float2(g.z, g.w)};
#pragma unroll
            for (int k = 0; k < 2; ++k) {  // delta = rstd (gamma d - t1 - xhat t2)
              const int i = g8 * 8 + 4 * h + 2 * k;
              const float2 w = fma2(xh[2 * h + k], nt2r, fma2(gg[k], mul2(make_float2(u[i], u[i + 1]), rs), nt1r));
              u[i] = w.x; u[i + 1] = w.y;
            }
          }
          *reinterpret_cast<uint4*>(sm + A2_OFF2 + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3)) =
              make_uint4(pack_bf16(u[g8 * 8], u[g8 * 8 + 1]), pack_bf16(u[g8 * 8 + 2], u[g8 * 8 + 3]),
                         pack_bf16(u[g8 * 8 + 4], u[g8 * 8 + 5]), pack_bf16(u[g8 * 8 + 6], u[g8 * 8 + 7]));
        }
        acc_db2[ch] += warp_colsum32(u, lane);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    MARK(2);

    // ================= GEMM3: dG = delta2 W2 (hidden halves as separate N = 128 accumulators) ; delta2 image -> HBM ==========
    if (tid == 0) {
      bulk_s2g(A.d2_img + (size_t)tile * GS_BYTES, sm_u + A2_OFF2, GS_BYTES);
      bulk_commit();
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int slot = g & 1, kb = g >> 1, half = g & 1;
        full_wait(slot);
        umma_kblock(tmem + half * L, sm_u + A2_OFF2 + kb * A_BLK_BYTES, sm_u + SLOT_OFF2 + slot * SEG_BLK, idesc_l, kb == 0);
        umma_commit(BAR(B_FREE + slot));
      }
      umma_commit(BAR(ACC));
    }
    if (tid == 32) {  // pieces 2 / 3 follow 0 / 1, then GEMM4's first two pieces follow those
      free_wait(0); fill(2);
      free_wait(1); fill(3);
      free_wait(0); fill(4);
      free_wait(1); fill(5);
    }
    // stashed xhat1 of this thread's first 32 hidden columns: in flight under GEMM3
    const uint4* xh1 = A.xh1 + ((size_t)tile * (H / 8) + hs * 16) * TILE_M + row;  // this thread's 16 x 8 bf16 of xhat1
    uint4 xq_nx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xq_nx[j] = __ldg(xh1 + (size_t)j * TILE_M);
    if (warp == 0) mbar_wait(BAR(ACC), acc_par);
    __syncthreads();
    acc_par ^= 1;
    tc_fence_after();
    MARK(3);

    // ================= EPI-C: d(y1) = dG * act'(y1) (parked in place), LayerNorm-1 adjoint -> delta1 image =================
    {
      const uint32_t t_dg = t_lane + hs * 128;
      float2 s1v = make_float2(0.f, 0.f), s2v = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        const int cb = hs * 128 + ch * 32;
        uint4 xq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xq[j] = xq_nx[j];
        if (ch + 1 < 4) {  // next chunk's xhat1 while this one is processed
#pragma unroll
          for (int j = 0; j < 4; ++j) xq_nx[j] = __ldg(xh1 + (size_t)((ch + 1) * 4 + j) * TILE_M);
        }
        float u[32], tmp[32];
        tmem_ld32(t_dg + ch * 32, u);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float2 xh[4];
          unpack8_2(xq[g8], xh);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 g = *reinterpret_cast<const float4*>(s_g1 + cb + g8 * 8 + 4 * h);
            const float4 be = *reinterpret_cast<const float4*>(s_be1 + cb + g8 * 8 + 4 * h);
            const float2 gg[2] = {make_float2(g.x, g.y), make_float2(g.z, g.w)}, ee[2] = {make_float2(be.x, be.y), make_float2(be.z, be.w)};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int i = g8 * 8 + 4 * h + 2 * k;
              const float2 x = xh[2 * h + k];
              const float2 d = mul2(make_float2(u[i], u[i + 1]), tc_act_bwd2<ACT_H>(fma2(x, gg[k], ee[k])));
              const float2 gd = mul2(gg[k], d), dx = mul2(d, x);
              u[i] = d.x; u[i + 1] = d.y;
              tmp[i] = dx.x; tmp[i + 1] = dx.y;
              s1v = add2(s1v, gd);
              s2v = fma2(gd, x, s2v);
            }
          }
        }
        tmem_st32(t_dg + ch * 32, u);  // park d(y1) where dG was
        acc_dg1[ch] += warp_colsum32(tmp, lane);
        acc_dbe1[ch] += warp_colsum32(u, lane);
      }
      if (tid == 0) bulk_wait_read0();  // the delta2 image has left shared memory: pass 2 overwrites its bytes with delta1
      s_red[row * 4 + hs * 2] = s1v.x + s1v.y;     // (EPI-B's readers of s_red passed the barriers of GEMM3)
      s_red[row * 4 + hs * 2 + 1] = s2v.x + s2v.y;
      __syncthreads();
      const float t1 = (s_red[row * 4] + s_red[row * 4 + 2]) * (1.0f / H);
      const float t2 = (s_red[row * 4 + 1] + s_red[row * 4 + 3]) * (1.0f / H);
      const float2 rs = splat2(rstd1), nt1r = splat2(-t1 * rstd1), nt2r = splat2(-t2 * rstd1);
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        const int cb = hs * 128 + ch * 32;
        uint4 xq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xq[j] = __ldg(xh1 + (size_t)(ch * 4 + j) * TILE_M);
        float u[32];
        tmem_ld32(t_dg + ch * 32, u);
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float2 xh[4];
          unpack8_2(xq[g8], xh);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 g = *reinterpret_cast<const float4*>(s_g1 + cb + g8 * 8 + 4 * h);
            const float2 gg[2] = {make_float2(g.x, g.y), make_float2(g.z, g.w)};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int i = g8 * 8 + 4 * h + 2 * k;
              const float2 w = fma2(xh[2 * h + k], nt2r, fma2(gg[k], mul2(make_float2(u[i], u[i + 1]), rs), nt1r));
              u[i] = w.x; u[i + 1] = w.y;
            }
          }
          const int c = cb + g8 * 8;
          *reinterpret_cast<uint4*>(sm + A2_OFF2 + (c / KBLK) * A_BLK_BYTES + sw128_off(row, (c % KBLK) >> 3)) =
              make_uint4(pack_bf16(u[g8 * 8], u[g8 * 8 + 1]), pack_bf16(u[g8 * 8 + 2], u[g8 * 8 + 3]),
                         pack_bf16(u[g8 * 8 + 4], u[g8 * 8 + 5]), pack_bf16(u[g8 * 8 + 6], u[g8 * 8 + 7]));
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    MARK(4);

    // ================= GEMM4: d(e)_mlp = delta1 W1c ; delta1 image -> HBM =================
    if (tid == 0) {
      bulk_s2g(A.d1_img + (size_t)tile * A2_BYTES, sm_u + A2_OFF2, A2_BYTES);
      bulk_commit();
      tc_fence_after();
#pragma unroll
      for (int g = 4; g < NPIECE2; ++g) {
        const int slot = g & 1, kb = g - 4;
        full_wait(slot);
        umma_kblock(tmem, sm_u + A2_OFF2 + kb * A_BLK_BYTES, sm_u + SLOT_OFF2 + slot * SEG_BLK, idesc_l, kb == 0);
        umma_commit(BAR(B_FREE + slot));
      }
      umma_commit(BAR(ACC));
      bulk_wait_read0();  // the delta1 image has left shared memory before it becomes the d(e) staging tile
    }
    if (tid == 32) {
      free_wait(0); fill(6);
      free_wait(1); fill(7);
    }
    float4 skipg[16];  // fp32 upstream gradient (skip path of d(e)) of this lane's 16 output rows, in flight under GEMM4
#pragma unroll         // (the adjoint passes' registers are dead here: 64 registers of loads cost nothing)
    for (int k = 0; k < 16; ++k) {
      const int r = warp * 16 + k;
      const bool live = (int64_t)tile * TILE_M + r < A.n_edges;
      float4 go = live ? ldg_f4_hint(reinterpret_cast<const float4*>(A.g_e + (size_t)s_eid[r] * L) + lane, pol_drop) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (live && A.g_agg) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(A.g_agg + (size_t)s_dst[r] * L) + lane);
        go.x += a.x; go.y += a.y; go.z += a.z; go.w += a.w;
      }
      skipg[k] = go;
    }
    if (warp == 0) mbar_wait(BAR(ACC), acc_par);
    __syncthreads();  // also orders thread 0's bulk_wait_read0 before the staging writes below
    acc_par ^= 1;
    tc_fence_after();
    if (tid == 32) {  // every MMA has retired: the next tile's first two pieces travel under EPI-D
      free_wait(0); free_wait(1);
      if (has_next) { fill(0); fill(1); }
    }
    MARK(5);

    // ================= EPI-D: rows of d(e) through a swizzled fp32 staging tile =================
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      float v[32];
      tmem_ld32(t_lane + hs * 64 + ch * 32, v);
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        const int c4 = hs * 16 + ch * 8 + g4;
        *reinterpret_cast<float4*>(sm + A2_OFF2 + (size_t)row * (L * 4) + ((c4 ^ (row & 7)) << 4)) =
            make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
      }
    }
    tc_fence_before();
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {  // 16 rows per warp, one float4 chunk per lane; skip connection: d(e) += gout (fp32)
      const int r = warp * 16 + k, c4 = lane;
      if ((int64_t)tile * TILE_M + r < A.n_edges) {
        float4 y = *reinterpret_cast<const float4*>(sm + A2_OFF2 + (size_t)r * (L * 4) + ((c4 ^ (r & 7)) << 4));
        y.x += skipg[k].x; y.y += skipg[k].y; y.z += skipg[k].z; y.w += skipg[k].w;
        *reinterpret_cast<float4*>(A.d_e + (size_t)s_eid[r] * L + c4 * 4) = y;
      }
    }
    fence_proxy_async();  // staging (generic proxy) precedes the next tile's bulk store from these bytes
    __syncthreads();
    MARK(6);
  }

  // ---- ordered hand-off of the column sums: [cta][q][PAR_FLOATS]; layout mirrors s_par: db1 | dgamma1 | dbeta1 | db2 | dgamma2 | dbeta2
  {
    float* o = A.colpart + ((size_t)blockIdx.x * 4 + q) * PAR_FLOATS;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      const int c = hs * 128 + ch * 32 + lane;
      o[c] = 0.f;  // d bias1: formed per node by the caller of this kernel
      o[H + c] = acc_dg1[ch];
      o[2 * H + c] = acc_dbe1[ch];
    }
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int c = hs * 64 + ch * 32 + lane;
      o[3 * H + c] = acc_db2[ch];
      o[3 * H + L + c] = acc_dg2[ch];
      o[3 * H + 2 * L + c] = acc_dbe2[ch];
    }
  }
  if (tid == 0) bulk_wait0();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace v2
