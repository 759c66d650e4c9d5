// Row collectives of the destination-partitioned event over NVLink / NVSwitch peer memory:
//
//   all-gather      every rank's block of updated node rows -> every rank's replicated [world * rows, width] table
//   reduce-scatter  sum over ranks of the [world * rows, width] node-gradient partials -> the owned block
//
// The tables live in SYMMETRIC buffers (same size on every rank, mapped into every peer: CUDA IPC / fabric handles; the
// host side obtains them with torch.distributed._symmetric_memory, parallel.py:SymmetricRows). Two data paths:
//   * multicast (NVLS): one `multimem.st` stores a 16-byte chunk into ALL ranks' copies at once, one
//     `multimem.ld_reduce.add` returns the sum of the chunk over all ranks, added inside the switch — a rank moves its
//     block once instead of world - 1 times, and the reduce-scatter needs no temporary;
//   * peer pointers: plain 16-byte stores to / loads from each peer's buffer (boxes without multicast support).
// Neither kernel synchronises ranks: the caller brackets them with the symmetric-memory barrier (all blocks written
// before anybody reads the table; all partials written before anybody reduces; nobody overwrites a table a peer still reads).
// NCCL needed 0.20 ms (all-gather) and 0.29 ms (reduce-scatter, RING_LL) for the 61 MB table of the full pile-up
// event at 8 GPUs — latency / protocol bound, not link bound (each rank owns 7.7 MB).
#include "common.cuh"

using namespace hgnn;

namespace {

constexpr int P2P_THREADS = 256;
constexpr int MAX_WORLD = 16;

struct PeerPtrs { float* p[MAX_WORLD]; };

__device__ __forceinline__ void multimem_st4(float* mc, const float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_add4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}

// block of this rank: chunks [0, n4) of `local` -> slot `rank` of the table on every rank
__global__ void __launch_bounds__(P2P_THREADS) k_all_gather_rows(const float4* __restrict__ local, int64_t n4, float* mc_base, PeerPtrs peers,
                                                                int world, int rank) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t slot = (int64_t)rank * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(local + i);
    if (mc_base) {
      multimem_st4(mc_base + (slot + i) * 4, v);
    } else {
#pragma unroll 1
      for (int w = 0; w < world; ++w) reinterpret_cast<float4*>(peers.p[w])[slot + i] = v;
    }
  }
}

// owned block of the sum over ranks of the partial tables: out[i] = sum_w table_w[rank * n4 + i]
__global__ void __launch_bounds__(P2P_THREADS) k_reduce_scatter_rows(float4* __restrict__ out, int64_t n4, const float* mc_base, PeerPtrs peers,
                                                                    int world, int rank) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t slot = (int64_t)rank * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc;
    if (mc_base) {
      acc = multimem_ld_add4(mc_base + (slot + i) * 4);
    } else {  // fixed rank order: bit-reproducible run to run
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int w = 0; w < world; ++w) {
        const float4 v = reinterpret_cast<const float4*>(peers.p[w])[slot + i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    out[i] = acc;
  }
}

// in-place sum over ranks of the first n4 chunks of the symmetric buffer, left in every rank's copy: rank r reduces the
// r-th slice (inside the switch with multimem.ld_reduce, or peer by peer in rank order) and stores the result to all
__global__ void __launch_bounds__(P2P_THREADS) k_all_reduce(int64_t n4, float* mc_base, PeerPtrs peers, int world, int rank) {
  const int64_t per = (n4 + world - 1) / world;
  const int64_t lo = (int64_t)rank * per, hi = min(n4, lo + per);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    if (mc_base) {
      multimem_st4(mc_base + i * 4, multimem_ld_add4(mc_base + i * 4));
    } else {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int w = 0; w < world; ++w) {
        const float4 v = reinterpret_cast<const float4*>(peers.p[w])[i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
#pragma unroll 1
      for (int w = 0; w < world; ++w) reinterpret_cast<float4*>(peers.p[w])[i] = acc;
    }
  }
}

int fill_peers(PeerPtrs& P, const uint64_t* peer_bases, int world) {
  for (int w = 0; w < MAX_WORLD; ++w) P.p[w] = nullptr;
  if (peer_bases)
    for (int w = 0; w < world; ++w) P.p[w] = reinterpret_cast<float*>(peer_bases[w]);
  return 0;
}

unsigned p2p_grid(int64_t n4) {
  const int64_t want = (n4 + P2P_THREADS - 1) / P2P_THREADS;
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, 4 * (int64_t)num_sms()));
}

}  // namespace

extern "C" int hgnn_p2p_all_gather_rows(const float* local, int64_t rows, int64_t width, void* mc_base, const uint64_t* peer_bases,
                                        int world, int rank, void* stream) {
  HGNN_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "p2p_all_gather_rows: bad world / rank");
  HGNN_REQUIRE(local && (mc_base || peer_bases), "p2p_all_gather_rows: NULL pointer (need a multicast base or the peer bases)");
  HGNN_REQUIRE((rows * width) % 4 == 0 && ((uintptr_t)local % 16) == 0, "p2p_all_gather_rows: blocks must be whole 16-byte chunks");
  if (rows <= 0 || width <= 0) return HGNN_OK;
  PeerPtrs P;
  fill_peers(P, peer_bases, world);
  const int64_t n4 = rows * width / 4;
  k_all_gather_rows<<<p2p_grid(n4), P2P_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(local), n4, (float*)mc_base, P, world, rank);
  return check_launch("p2p_all_gather_rows");
}

extern "C" int hgnn_p2p_reduce_scatter_rows(float* out, int64_t rows, int64_t width, const void* mc_base, const uint64_t* peer_bases,
                                            int world, int rank, void* stream) {
  HGNN_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "p2p_reduce_scatter_rows: bad world / rank");
  HGNN_REQUIRE(out && (mc_base || peer_bases), "p2p_reduce_scatter_rows: NULL pointer (need a multicast base or the peer bases)");
  HGNN_REQUIRE((rows * width) % 4 == 0 && ((uintptr_t)out % 16) == 0, "p2p_reduce_scatter_rows: blocks must be whole 16-byte chunks");
  if (rows <= 0 || width <= 0) return HGNN_OK;
  PeerPtrs P;
  fill_peers(P, peer_bases, world);
  const int64_t n4 = rows * width / 4;
  k_reduce_scatter_rows<<<p2p_grid(n4), P2P_THREADS, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(out), n4, (const float*)mc_base, P, world, rank);
  return check_launch("p2p_reduce_scatter_rows");
}

extern "C" int hgnn_p2p_all_reduce(int64_t n_floats, void* mc_base, const uint64_t* peer_bases, int world, int rank, void* stream) {
  HGNN_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "p2p_all_reduce: bad world / rank");
  HGNN_REQUIRE(mc_base || peer_bases, "p2p_all_reduce: need a multicast base or the peer bases");
  HGNN_REQUIRE(n_floats % 4 == 0, "p2p_all_reduce: the buffer must be whole 16-byte chunks");
  if (n_floats <= 0) return HGNN_OK;
  PeerPtrs P;
  fill_peers(P, peer_bases, world);
  const int64_t n4 = n_floats / 4;
  k_all_reduce<<<p2p_grid((n4 + world - 1) / world), P2P_THREADS, 0, (cudaStream_t)stream>>>(n4, (float*)mc_base, P, world, rank);
  return check_launch("p2p_all_reduce");
}
