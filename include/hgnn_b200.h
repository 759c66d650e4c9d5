/*
 * hgnn_b200.h — C ABI of libhgnn_b200.so: hand-written sm_100a kernels for the
 * HierarchicalGNN message-passing hot path.
 *
 * Conventions ("Third-party op boundary being replaced"):
 *   - plain C: raw device pointers + explicit int64 sizes; no torch types;
 *   - every call launches on the caller's `stream` (a cudaStream_t passed as
 *     void*), never allocates device memory, never synchronises unless stated,
 *     and keeps no global mutable state; scratch comes from the caller through
 *     `ws`/`ws_bytes`, sized by the matching *_workspace_bytes() query;
 *   - returns 0 (HGNN_OK) or a negative hgnn_status; hgnn_last_error() gives a
 *     thread-local message for the last non-zero status on this host thread;
 *   - feature rows are fp32, row-major, contiguous; graphs cross the boundary
 *     as int64 (as the reference passes them) and are int32 internally.
 *
 * Each entry point cites the reference interface (file:line under
 * /root/reference/Modules) that it replaces.
 */
#ifndef HGNN_B200_H
#define HGNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGNN_ABI_VERSION 1

typedef enum {
  HGNN_OK = 0,
  HGNN_ERR_BAD_ARG = -1,
  HGNN_ERR_UNSUPPORTED = -2,
  HGNN_ERR_WORKSPACE = -3,
  HGNN_ERR_CUDA = -4
} hgnn_status;

/* activation codes: the torch.nn names accepted by make_mlp (utils.py:178-181) */
typedef enum {
  HGNN_ACT_NONE = 0,
  HGNN_ACT_GELU = 1, /* exact erf form, nn.GELU() default */
  HGNN_ACT_TANH = 2,
  HGNN_ACT_RELU = 3,
  HGNN_ACT_SILU = 4,
  HGNN_ACT_SIGMOID = 5
} hgnn_act;

int hgnn_abi_version(void);
const char* hgnn_last_error(void);

/* ------------------------------------------------------------------------
 * Destination-sorted CSR ("segment plan") — built once per graph and reused by
 * every cell, forward and backward. Replaces the per-call float-atomic
 * scatter inside torch_scatter.scatter_add (gnn_utils.py:50,124,125,142,143).
 *   keys[n_items]  int64 segment id of each item (e.g. graph[1]), 0 <= key < n_segments
 *   perm[n_items]  int32 out: item ids ordered by (key, item id)  (stable)
 *   rowptr[n_segments+1] int32 out: segment s owns perm[rowptr[s] .. rowptr[s+1])
 *   keys32[n_items] int32 out (optional, may be NULL): keys narrowed to int32
 * ------------------------------------------------------------------------ */
size_t hgnn_csr_build_workspace_bytes(int64_t n_items);
int hgnn_csr_build(const int64_t* keys, int64_t n_items, int64_t n_segments, int32_t* perm, int32_t* rowptr,
                   int32_t* keys32, void* ws, size_t ws_bytes, void* stream);

/* Narrow an int64 index array to int32 (range-checked on device: out-of-range
 * entries are clamped and counted into *bad if bad != NULL). */
int hgnn_index_to_i32(const int64_t* in, int64_t n, int64_t limit, int32_t* out, int32_t* bad, void* stream);

/* ------------------------------------------------------------------------
 * Segmented (weighted, gathered) row reduction — deterministic, no atomics.
 *   out[s, :] = scale_s * sum_{j in [rowptr[s], rowptr[s+1])} w[i] * src[g(i), :],  i = perm[j]
 *   g(i) = gather ? gather[i] : i ;  w[i] = weight ? weight[i] : 1 ;
 *   scale_s = mean ? 1/max(count_s,1) : 1
 * Replaces scatter_add / scatter_mean and the "w * X[idx] -> scatter" bipartite
 * aggregations (gnn_utils.py:50,124-125,142-143; BC/Models/HGNN_GMM.py:251,269).
 * Its adjoint w.r.t. src is the same call on the transposed plan.
 * ------------------------------------------------------------------------ */
int hgnn_segment_reduce(const float* src, int64_t width, const int32_t* gather, const float* weight,
                        const int32_t* perm, const int32_t* rowptr, int64_t n_segments, int mean, float* out,
                        void* stream);
/* Same with an explicit source row stride (floats): rows of a column slice of a wider matrix are reduced in place
 * (the column blocks of a layer's input gradient, one per gathered segment) instead of being copied out first. */
int hgnn_segment_reduce_ld(const float* src, int64_t width, int64_t src_ld, const int32_t* gather, const float* weight,
                           const int32_t* perm, const int32_t* rowptr, int64_t n_segments, int mean, float* out,
                           void* stream);

/* out[i, :] = w[i] * src[idx[i], :]  (idx NULL = identity). Adjoint of an
 * ungathered segment reduce; also the row gather used by the encoders. */
int hgnn_gather_rows(const float* src, int64_t width, const int32_t* idx, const float* weight, int64_t n_items,
                     float* out, void* stream);

/* out[i] = sum_d a[ai[i], d] * b[bi[i], d]   (ai / bi NULL = identity).
 * Replaces einsum('ij,ij->i', src[g0], dst[g1]) (gnn_utils.py:208;
 * BC/Models/HGNN_GMM.py:188) and gives d(weight) of a weighted segment reduce. */
int hgnn_edge_dot(const float* a, const int32_t* ai, const float* b, const int32_t* bi, int64_t width,
                  int64_t n_items, float* out, void* stream);

/* ------------------------------------------------------------------------
 * Fused gathered-concat MLP (make_mlp, utils.py:169-196) applied row-wise:
 *   in_row(r)  = concat_s seg_ptr[s][ seg_idx[s] ? seg_idx[s][r] : r , : ]
 *   per layer l: h = a W_l^T + b_l ; optional LayerNorm(gamma_l, beta_l, eps) ; act_l
 *   out[ out_idx ? out_idx[r] : r , : ] = a_last (+ seg_ptr[skip_seg] row if skip_seg >= 0)
 * One kernel covers the edge step (gnn_utils.py:56-64,129-135,147-153: segments
 * x[src] | x[dst] | e, skip = e), the node / supernode steps (gnn_utils.py:45-54,
 * 119-127,137-145: segments x | agg [| agg2], skip = x), the encoders and heads
 * (EC/Models/IN.py:84-85,126; BC/Models/HGNN_GMM.py:270-271,342-344).
 * W_l is [out_width[l], in_width] row-major exactly as nn.Linear stores it.
 * ------------------------------------------------------------------------ */
#define HGNN_MLP_MAX_SEGS 3
#define HGNN_MLP_MAX_LAYERS 4

typedef struct {
  int32_t n_seg;
  int32_t n_layers;
  int32_t skip_seg; /* -1: no skip connection */
  float ln_eps;
  const float* seg_ptr[HGNN_MLP_MAX_SEGS];
  const int32_t* seg_idx[HGNN_MLP_MAX_SEGS];
  int32_t seg_width[HGNN_MLP_MAX_SEGS];
  int32_t out_width[HGNN_MLP_MAX_LAYERS];
  int32_t act[HGNN_MLP_MAX_LAYERS];
  const float* W[HGNN_MLP_MAX_LAYERS];
  const float* b[HGNN_MLP_MAX_LAYERS];
  const float* gamma[HGNN_MLP_MAX_LAYERS]; /* NULL: no LayerNorm after this layer */
  const float* beta[HGNN_MLP_MAX_LAYERS];
  const int32_t* out_idx; /* optional output row scatter */
} hgnn_mlp_desc;

int hgnn_mlp_forward(const hgnn_mlp_desc* d, int64_t rows, float* out, void* stream);

/* Backward with in-kernel recompute (replaces torch.utils.checkpoint around
 * every update, gnn_utils.py:14-15). Two launches:
 *  (1) hgnn_mlp_backward_data: recomputes the forward per row tile, produces
 *      - dseg[s]: per-ROW input gradients [rows, seg_width[s]] (NULL = skip);
 *        for gathered segments the caller reduces them with the transposed
 *        plan (hgnn_segment_reduce), keeping the result deterministic;
 *      - per-layer row buffers in ws for the weight-gradient pass;
 *      - dvec[l]: packed [3, out_width[l]] = (d bias, d gamma, d beta).
 *  (2) hgnn_mlp_backward_weights: dW[l] = delta_l^T a_{l-1} as a split-K
 *      reduction over rows with a deterministic second stage.
 * grad_out is [rows, out_width[last]] indexed like `out` (through out_idx). */
size_t hgnn_mlp_backward_workspace_bytes(const hgnn_mlp_desc* d, int64_t rows);
int hgnn_mlp_backward_data(const hgnn_mlp_desc* d, int64_t rows, const float* grad_out,
                           float* const dseg[HGNN_MLP_MAX_SEGS], float* const dvec[HGNN_MLP_MAX_LAYERS], void* ws,
                           size_t ws_bytes, void* stream);
int hgnn_mlp_backward_weights(const hgnn_mlp_desc* d, int64_t rows, float* const dW[HGNN_MLP_MAX_LAYERS], void* ws,
                              size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * Radius-limited k nearest neighbours, brute force, tiled through shared
 * memory. Replaces frnn.frnn_grid_points as called by find_neighbors
 * (utils.py:228-239): for each query row the up-to-k reference rows with
 * squared distance < radius^2, ascending (ties: smaller index), -1 padded.
 * Distances are direct fp32 sums of squared differences.
 *   idx[n_query, k] int64 out;  dim <= 32, k <= 32.
 * ------------------------------------------------------------------------ */
int hgnn_knn_radius(const float* query, int64_t n_query, const float* ref, int64_t n_ref, int64_t dim, int64_t k,
                    float radius, int64_t* idx, void* stream);
/* Same result, bit for bit. With a workspace (hgnn_knn_radius_workspace_bytes; 0 bytes for large problems) a small
 * problem is split over contiguous reference ranges (grid = query blocks x splits, two CTAs per SM) and the per-split
 * sorted lists are merged (ties to the lower range = the smaller index); without one it runs as hgnn_knn_radius. */
size_t hgnn_knn_radius_workspace_bytes(int64_t n_query, int64_t n_ref, int64_t k);
int hgnn_knn_radius_ws(const float* query, int64_t n_query, const float* ref, int64_t n_ref, int64_t dim, int64_t k,
                       float radius, int64_t* idx, void* ws, size_t ws_bytes, void* stream);
/* A batch of events in one launch: query rows [query_ptr[b], query_ptr[b+1]) only meet reference rows
 * [ref_ptr[b], ref_ptr[b+1]) — the result of calling hgnn_knn_radius once per event (the reference's loaders hand the
 * model one event at a time, bipartite_classification_base.py:42, so its find_neighbors never sees two), with the
 * neighbour ids kept as row numbers of the whole `ref` matrix. query_ptr / ref_ptr: DEVICE arrays of n_events + 1
 * ascending int32 offsets (first 0, last n_query / n_ref); they are never read on the host. Workspace as above (optional). */
int hgnn_knn_radius_batched(const float* query, int64_t n_query, const float* ref, int64_t n_ref, int64_t dim, int64_t k,
                            float radius, const int32_t* query_ptr, const int32_t* ref_ptr, int64_t n_events, int64_t* idx,
                            void* ws, size_t ws_bytes, void* stream);

/* Compacts idx (>= 0 entries, query-major, rank-minor: gnn_utils.py:195-202)
 * into graph[2, n_query*k] (row 0 = query id, row 1 = neighbour id; only the
 * first *count columns are valid). `count` is a device int64; ld = n_query*k. */
size_t hgnn_knn_edges_workspace_bytes(int64_t n_query);
int hgnn_knn_edges(const int64_t* idx, int64_t n_query, int64_t k, int64_t* graph, int64_t* count, void* ws,
                   size_t ws_bytes, void* stream);

/* Union of an edge list with its transpose, duplicates removed, columns in
 * lexicographic order. Replaces cugraph symmetrize (gnn_utils.py:198-199).
 *   graph_in[2, n_edges] (row stride ld_in) -> graph_out[2, 2*n_edges] (row
 *   stride 2*n_edges), *count valid columns (device int64). */
size_t hgnn_symmetrize_workspace_bytes(int64_t n_edges);
int hgnn_symmetrize(const int64_t* graph_in, int64_t ld_in, int64_t n_edges, int64_t n_vertices, int64_t* graph_out,
                    int64_t* count, void* ws, size_t ws_bytes, void* stream);

/* max_i || a[ai[i]] - b[bi[i]] ||_2 over the edge list -> *out (device float);
 * the radius tracker at gnn_utils.py:203-205. *out must be zeroed by caller. */
int hgnn_edge_max_dist(const float* a, const int64_t* ai, const float* b, const int64_t* bi, int64_t width,
                       int64_t n_items, float* out, void* stream);

/* Weakly connected components over the kept edges (keep == NULL: all edges):
 * labels[v] = smallest vertex id of v's component, -1 for vertices touched by
 * no kept edge. Replaces cugraph.components.connected_components as used by
 * HierarchicalGNNBlock.clustering (BC/Models/HGNN_GMM.py:215-232). */
size_t hgnn_connected_components_workspace_bytes(int64_t n_vertices);
int hgnn_connected_components(const int64_t* graph, int64_t ld, int64_t n_edges, const uint8_t* keep,
                              int64_t n_vertices, int32_t* labels, void* ws, size_t ws_bytes, void* stream);

/* 1-D two-component Gaussian mixture by EM, entirely on device (replaces the
 * sklearn fit at BC/Models/HGNN_GMM.py:192). params[6] out (device):
 * (pi0, mu0, var0, pi1, mu1, var1). ws: hgnn_gmm1d_workspace_bytes(). */
size_t hgnn_gmm1d_workspace_bytes(void);
int hgnn_gmm1d_fit(const float* x, int64_t n, int32_t max_iter, float tol, float* params, void* ws, size_t ws_bytes,
                   void* stream);

/* ------------------------------------------------------------------------
 * Tensor-core (tcgen05 / TMEM) fused edge step — bf16 operands, fp32
 * accumulate, fp32 storage. See hgnn_tc.h section below; available for
 * latent in {64, 128} with the 2-layer LayerNorm GELU/Tanh edge network of the HGNN configs.
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t latent;     /* L */
  int32_t hidden;     /* H */
  int32_t act_hidden; /* hgnn_act of layer 0 */
  int32_t act_out;    /* hgnn_act of layer 1 */
  float ln_eps;
  const void* w1_packed; /* bf16 UMMA smem images, from hgnn_tc_pack_weights */
  const void* w2_packed;
  const float* b1;
  const float* gamma1;
  const float* beta1;
  const float* b2;
  const float* gamma2;
  const float* beta2;
  /* Profiling, per call (NULL in production): device buffer of 16 uint64 in which CTA 0 of the forward / backward kernel
   * accumulates the cycles of each phase of its tile loop (forward: setup, GEMM1 + gather, EPI1, GEMM2, EPI2, store pass,
   * fused aggregate; backward: setup, LOAD, EPI-B, GEMM3, EPI-C, GEMM4, EPI-D). The library keeps no global state. */
  void* debug_phase_clock;
} hgnn_tc_edge_params;

int hgnn_tc_supported(int64_t latent, int64_t hidden, int64_t n_layers, int layer_norm, int act_hidden, int act_out);
size_t hgnn_tc_packed_weight_bytes(int64_t out_features, int64_t in_features);
/* fp32 nn.Linear weight [out, in] -> bf16, K-major, 128B-swizzled UMMA smem image */
int hgnn_tc_pack_weights(const float* W, int64_t out_features, int64_t in_features, void* packed, void* stream);
/* Unit-test entry: C[M,N] = bf16(A[M,K]) . bf16(W[N,K])^T, fp32 accumulate, built from
 * the same gather / descriptor / tcgen05 / TMEM pieces as the fused kernels. */
int hgnn_tc_debug_gemm(const float* A, const void* w_packed, int64_t M, int64_t N, int64_t K, float* C, void* stream);
/* Unit-test entry for the weight-gradient primitive: out[ca, cb] = bf16(A[rows, ca])^T . bf16(B[rows, cb]),
 * through bf16 tile images read back as MN-major UMMA operands (split-K over rows, ordered reduce). */
size_t hgnn_tc_debug_wgrad_workspace_bytes(int64_t rows, int64_t ca, int64_t cb);
int hgnn_tc_debug_wgrad(const float* A, const float* B, int64_t rows, int64_t ca, int64_t cb, float* out, void* ws,
                        size_t ws_bytes, void* stream);
/* The fused edge step:  e_out[i] = MLP([x[src_i] | x[dst_i] | e_i]) + e_i, edges visited in row order `perm`
 * (NULL = identity). If agg != NULL the same launch also produces agg[n] = sum_{dst_i = n} e_out[i] — the
 * scatter_add that opens the next cell (gnn_utils.py:50) — by a destination-sorted, ordered segmented reduce of
 * each finished tile in shared memory (no atomics); this needs perm/rowptr = the by-destination plan of
 * hgnn_csr_build (perm may be NULL when the edges are already stored destination-sorted: rows then stream in place). Segments crossing a row-group boundary (hub nodes) and empty segments are completed by a small
 * second kernel inside the same call.
 * stash (optional, hgnn_tc_edge_stash_bytes): pass it when a backward will follow. The kernel then also leaves in HBM,
 * per 128-edge tile, the bf16 tile images of the edge-latent columns of its first MMA operand and of the hidden activation,
 * the normalised pre-affine activations of both LayerNorms as bf16 and the row rstd's (1.5 KB per edge at latent 128):
 * everything hgnn_tc_edge_backward and the weight-gradient GEMM need, so the backward recomputes nothing and
 * never gathers again. */
/* ws (hgnn_tc_edge_forward_workspace_bytes) holds the bf16 shadow copy of x the call makes first: every node row is gathered
 * ~2 E / N times, so it is converted once and gathered as 256 B rows (n_nodes = rows of x, always required). */
size_t hgnn_tc_edge_forward_workspace_bytes(int64_t n_edges, int64_t n_nodes, int64_t latent);
size_t hgnn_tc_edge_stash_bytes(int64_t n_edges, int64_t latent);
int hgnn_tc_edge_forward(const hgnn_tc_edge_params* p, const float* x, const float* e, const int32_t* src,
                         const int32_t* dst, const int32_t* perm, const int32_t* rowptr, int64_t n_edges, int64_t n_nodes,
                         float* e_out, float* agg, void* stash, void* ws, size_t ws_bytes, void* stream);

/* Backward of the tensor-core edge step (latent 128) from the forward's stash (no recompute, no gathers).
 * Per edge it back-propagates through Tanh/LN/Linear/GELU/LN and the edge-latent columns of the first Linear only
 * (d_e, final). The node part of the first layer is linear in x, so it is done per NODE: a segmented reduce over the
 * delta1 tile image gives R_src[n] / R_dst[n] = sum of delta1 over the edges leaving / entering n, then
 *   d_x = R_src W1[:, 0:L] + R_dst W1[:, L:2L]   (one node-level tcgen05 GEMM, d_x is [n_nodes, L], complete)
 *   dW1[:, 0:L] = R_src^T x,  dW1[:, L:2L] = R_dst^T x   (node-level weight-gradient GEMM)
 * (the adjoint of nodes[graph[0]] / nodes[graph[1]] in gnn_utils.py:61, which autograd does as two index_add's over
 * [E, L] rows). dW1[:, 2L:3L], dW2 come from the per-edge tcgen05 split-K kernel; dvec{1,2} = [3, width] (d bias,
 * d gamma, d beta). stash / perm must be the buffer and row order of the matching hgnn_tc_edge_forward call.
 * {src,dst}_rows / {src,dst}_rowptr: CSR over the forward's TILE ROWS (position j of the row order, not edge ids) grouped
 * by source / destination node; dst_rows may be NULL when the row order is already destination-sorted.
 * grad_agg (optional, [n_nodes, L]) is the cotangent of agg = scatter_add(e_out, dst): the kernel uses
 * grad_eout[i] + grad_agg[dst_i]. w1t/w2t_packed are hgnn_tc_pack_weights images of W1^T / W2^T; wx_packed is the image
 * of the [L, 2H] matrix [W1[:, 0:L]^T | W1[:, L:2L]^T].
 * aux_stream (optional, NULL = none): a second stream of the same device. The per-edge weight-gradient GEMM then runs on it,
 * beside the node-level chain on `stream` (fork after the data kernel, join before returning: when the call returns, `stream`
 * is ordered after everything the call enqueued on either stream, so the caller needs no extra synchronisation). */
size_t hgnn_tc_edge_backward_workspace_bytes(int64_t n_edges, int64_t n_nodes);
int hgnn_tc_edge_backward(const hgnn_tc_edge_params* p, const void* w1t_packed, const void* w2t_packed, const void* wx_packed,
                          const void* stash, const float* x, int64_t n_nodes, const int32_t* dst, const int32_t* perm,
                          const int32_t* src_rows, const int32_t* src_rowptr, const int32_t* dst_rows,
                          const int32_t* dst_rowptr, int64_t n_edges, const float* grad_eout, const float* grad_agg, float* d_e,
                          float* d_x, float* dW1, float* dW2, float* dvec1, float* dvec2, void* ws, size_t ws_bytes,
                          void* stream, void* aux_stream);

/* ------------------------------------------------------------------------
 * Tensor-core row layer: ONE make_mlp layer (utils.py:183-196) on a gathered concatenation,
 *   out[r] = act(LayerNorm(W . [seg0[i0(r)] | seg1[i1(r)] | seg2[i2(r)]] + b)) (+ skip[r])
 * bf16 operands / fp32 accumulate in TMEM / fp32 storage. Covers the node and supernode updates
 * (gnn_utils.py:45-54,119-127,137-145), the encoder layers past the first and the hidden layers of the heads
 * (EC/Models/IN.py:29-48,126; BC/Models/HGNN_GMM.py:37-82,342-344). Built for: 1..3 segments of width % 64 == 0,
 * fan-in % 128 == 0 and <= 384, fan-out 128 or 256, LayerNorm present, activation GELU/Tanh/ReLU/SiLU. `skip`
 * (optional) is any fp32 [rows, fan_out] matrix added after the activation (the residual of the node updates);
 * its gradient is grad_out itself, so the backward entry does not return it.
 * hgnn_tc_row_supported() answers exactly that; other layers run on the fp32 kernels above.
 * a_img (optional, hgnn_tc_row_image_bytes(rows, fan_in)): bf16 tile image of the gathered input, left in HBM for
 * the backward (which recomputes the layer from it and never touches the fp32 inputs again).
 * Backward: d_in[rows, fan_in] per-row input gradients (the caller splits columns per segment and reduces gathered
 * segments with hgnn_segment_reduce), dW[fan_out, fan_in], dvec[3, fan_out] = (d bias, d gamma, d beta).
 * wt_packed = hgnn_tc_pack_weights image of W^T (fan-in rows; > 256 rows: 128-row packs interleaved per K-block).
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t n_seg;
  int32_t n_out;
  int32_t act; /* hgnn_act after the LayerNorm */
  float ln_eps;
  const float* seg_ptr[HGNN_MLP_MAX_SEGS];
  const int32_t* seg_idx[HGNN_MLP_MAX_SEGS]; /* NULL: rows used in place */
  int32_t seg_width[HGNN_MLP_MAX_SEGS];
  const void* w_packed; /* hgnn_tc_pack_weights image of W [n_out, fan_in] */
  const float* bias;
  const float* gamma;
  const float* beta;
  const float* skip; /* NULL: no residual */
} hgnn_tc_row_layer;

int hgnn_tc_row_supported(const hgnn_tc_row_layer* d);
size_t hgnn_tc_row_image_bytes(int64_t rows, int64_t fan_in);
int hgnn_tc_row_forward(const hgnn_tc_row_layer* d, int64_t rows, float* out, void* a_img, void* stream);
size_t hgnn_tc_row_backward_workspace_bytes(int64_t rows, int64_t fan_in, int64_t n_out);
int hgnn_tc_row_backward(const hgnn_tc_row_layer* d, const void* wt_packed, const void* a_img, int64_t rows,
                         const float* grad_out, float* d_in, float* dW, float* dvec, void* ws, size_t ws_bytes,
                         void* stream);
/* Same backward with d(input) delivered per gathered segment instead of as one [rows, fan_in] matrix: d_seg[s] is a dense
 * [rows, seg_width[s]] matrix, or NULL when that segment needs no gradient (its columns are then never stored). Segment widths
 * must be multiples of 128. What the autograd of torch.cat([a, b, c], dim=-1) feeding a Linear hands back to a, b and c
 * (reference gnn_utils.py:101,124-127,139-142) without the strided views of one wide matrix. */
int hgnn_tc_row_backward_split(const hgnn_tc_row_layer* d, const void* wt_packed, const void* a_img, int64_t rows,
                               const float* grad_out, float* const* d_seg, float* dW, float* dvec, void* ws, size_t ws_bytes,
                               void* stream);

/* ------------------------------------------------------------------------
 * Skinny layers (fp32, one warp per row) — the make_mlp layers that are pure bandwidth:
 *  narrow-in : a one-layer hgnn_mlp_desc with fan-in <= 8 and fan-out in {32, 64, 128, 256, 512} (the encoders' first
 *              Linear on x[N,3] / [x[src] | x[dst]]: EC/Models/IN.py:29-33,84-85; BC/Models/HGNN_GMM.py:37-41),
 *              optional LayerNorm + activation; backward gives d_in[rows, fan_in] (optional), dW[fan_out, fan_in],
 *              dvec[3, fan_out] = (d bias, d gamma, d beta).
 *  narrow-out: out[r, n] = <W[n, :], a[r, :]> + b[n] with fan-out <= 8, fan-in in {128, 256, 512}, fan-in * fan-out
 *              <= 2048 and no LayerNorm / activation (the score and embedding heads' last Linear: EC/Models/IN.py:126;
 *              BC/Models/HGNN_GMM.py:45-48,342-344); backward gives d_a[rows, fan_in] (optional), dW, db.
 * All reductions over rows are ordered (warp by warp, CTA by CTA): bit-reproducible.
 * ------------------------------------------------------------------------ */
int hgnn_narrow_in_supported(const hgnn_mlp_desc* d);
int hgnn_narrow_in_forward(const hgnn_mlp_desc* d, int64_t rows, float* out, void* stream);
size_t hgnn_narrow_in_backward_workspace_bytes(int64_t n_out);
int hgnn_narrow_in_backward(const hgnn_mlp_desc* d, int64_t rows, const float* grad_out, float* d_in, float* dW, float* dvec,
                            void* ws, size_t ws_bytes, void* stream);
int hgnn_narrow_out_supported(int64_t fan_in, int64_t n_out);
int hgnn_narrow_out_forward(const float* a, int64_t rows, int64_t fan_in, const float* W, const float* bias, int64_t n_out,
                            float* out, void* stream);
size_t hgnn_narrow_out_backward_workspace_bytes(int64_t fan_in, int64_t n_out);
int hgnn_narrow_out_backward(const float* a, int64_t rows, int64_t fan_in, const float* W, int64_t n_out, const float* grad_out,
                             float* d_a, float* dW, float* db, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * Generic tensor-core layer pieces for the shapes the fused kernels above do not cover (latent 64 / 256: fan-out 64 or
 * 512, fan-in up to 768) — a make_mlp layer then runs as  hgnn_tc_gemm (one call per <= 256 output columns)  ->
 * hgnn_ln_act_forward;  its backward as  hgnn_ln_act_backward -> hgnn_tc_gemm on W^T (data gradient)  ->
 * hgnn_tc_wgrad over the two A-operand images the GEMM calls leave behind.
 *  hgnn_tc_gemm: out[r, col0 : col0 + n_out] = [seg0[i0(r)] | seg1[i1(r)] | seg2[i2(r)]] . W^T (+ bias); the descriptor's
 *    act / gamma / beta / skip are ignored; n_out in {64, 128, 256}, segment widths % 64 == 0, fan-in <= 768;
 *    a_img (optional) receives the bf16 tile image of the gathered input.
 *  hgnn_ln_act_forward / backward: row-wise LayerNorm + activation (+ residual) on h[rows, n] (bias included), n in
 *    {64, 128, 256, 512}; backward gives delta[rows, n] and dvec[3, n] = (d bias, d gamma, d beta), ordered sums.
 *  hgnn_tc_wgrad: dW[n_out, fan_in] = delta^T A from the two images (split-K over rows, ordered reduce).
 * ------------------------------------------------------------------------ */
int hgnn_tc_gemm_supported(const hgnn_tc_row_layer* d);
int hgnn_tc_gemm(const hgnn_tc_row_layer* d, int64_t rows, float* out, int64_t ld_out, int64_t col0, void* a_img, void* stream);
int hgnn_ln_act_supported(int64_t n);
int hgnn_ln_act_forward(const float* h, int64_t rows, int64_t n, const float* gamma, const float* beta, float eps, int act,
                        const float* skip, float* out, void* stream);
size_t hgnn_ln_act_backward_workspace_bytes(int64_t n);
int hgnn_ln_act_backward(const float* h, const float* grad_out, int64_t rows, int64_t n, const float* gamma, const float* beta,
                         float eps, int act, float* delta, float* dvec, void* ws, size_t ws_bytes, void* stream);
int hgnn_tc_wgrad_supported(int64_t n_out, int64_t fan_in);
size_t hgnn_tc_wgrad_workspace_bytes(int64_t rows, int64_t n_out, int64_t fan_in);
int hgnn_tc_wgrad(const void* delta_img, int64_t n_out, const void* a_img, int64_t fan_in, int64_t rows, float* dW, void* ws,
                  size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------
 * Row collectives of the destination-partitioned event over NVLink / NVSwitch
 * peer memory (the reference has no distributed code).
 * The [world * rows, width] fp32 tables are SYMMETRIC buffers: the same
 * allocation on every rank, mapped into every peer. `mc_base` is the multicast
 * (NVLS) address of the table — multimem.st / multimem.ld_reduce move each
 * block once and add inside the switch — or NULL, in which case `peer_bases`
 * (host array of `world` device addresses, one per rank's copy) is used with
 * plain peer stores / loads summed in rank order. Neither call synchronises
 * ranks: bracket them with a cross-rank barrier.
 *   all_gather     local[rows, width] -> slot `rank` of every rank's table
 *   reduce_scatter out[rows, width]   =  sum over ranks of slot `rank` of their tables
 * ------------------------------------------------------------------------ */
int hgnn_p2p_all_gather_rows(const float* local, int64_t rows, int64_t width, void* mc_base, const uint64_t* peer_bases,
                             int world, int rank, void* stream);
int hgnn_p2p_reduce_scatter_rows(float* out, int64_t rows, int64_t width, const void* mc_base, const uint64_t* peer_bases,
                                 int world, int rank, void* stream);
/* In-place sum over ranks of the first n_floats (multiple of 4) of the symmetric buffer, left in every rank's copy
 * (weight-gradient all-reduce): rank r reduces the r-th slice and stores the result to all ranks. */
int hgnn_p2p_all_reduce(int64_t n_floats, void* mc_base, const uint64_t* peer_bases, int world, int rank, void* stream);

/* ------------------------------------------------------------------------
 * Loss side, HOST function (no device work, plain host pointers): maximum-weight matching that covers every row of a
 * block-diagonal sparse score table, the blocks solved side by side on host threads. Replaces
 * scipy.sparse.csgraph.min_weight_full_bipartite_matching(table, maximize=True) (reference
 * bipartite_classification_base.py:173, one event per call) for a collated batch of events: block b = rows
 * [row_ptr[b], row_ptr[b+1]) of the CSR table (indptr / indices / data, n_rows rows; a block only meets the columns
 * its rows list). col_of_row[n_rows] out = the matched column of every row. n_threads <= 0: one per hardware thread.
 * Error (HGNN_ERR_BAD_ARG) when a block has no matching that covers its rows.
 * ------------------------------------------------------------------------ */
int hgnn_match_blocks_max(const int32_t* indptr, const int32_t* indices, const float* data, int64_t n_rows,
                          const int64_t* row_ptr, int64_t n_blocks, int64_t* col_of_row, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* HGNN_B200_H */
