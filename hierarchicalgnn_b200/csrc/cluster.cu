// Supernode clustering support: single-pass GPU union-find
// connected components and an on-device 1-D two-component Gaussian-mixture EM.
// Replaces cugraph.components.connected_components and the CPU sklearn fit in
// HierarchicalGNNBlock.clustering (BC/Models/HGNN_GMM.py:184-234) without any
// device->host round trip inside the calls.
#include "common.cuh"

using namespace hgnn;

namespace {

// ---------------- union-find connected components ----------------
__global__ void k_cc_init(int32_t* __restrict__ parent, uint8_t* __restrict__ present, int64_t n) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) { parent[v] = (int32_t)v; present[v] = 0; }
}

__device__ __forceinline__ int32_t uf_find(volatile int32_t* parent, int32_t v) {
  // path halving; racy writes only ever replace a parent by one of its ancestors
  while (true) {
    int32_t p = parent[v];
    if (p == v) return v;
    int32_t gp = parent[p];
    if (gp != p) parent[v] = gp;
    v = p;
  }
}

__global__ void k_cc_union(const int64_t* __restrict__ g, int64_t ld, int64_t n_edges, const uint8_t* __restrict__ keep,
                           int32_t* parent, uint8_t* __restrict__ present) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  if (keep && !keep[e]) return;
  int32_t u = (int32_t)g[e], v = (int32_t)g[ld + e];
  present[u] = 1;
  present[v] = 1;
  while (true) {
    u = uf_find(parent, u);
    v = uf_find(parent, v);
    if (u == v) break;
    if (u > v) { int32_t t = u; u = v; v = t; }
    // hook the larger root under the smaller one => every root is its component's minimum id
    int32_t old = atomicCAS(&parent[v], v, u);
    if (old == v) break;
  }
}

__global__ void k_cc_finalize(int32_t* parent, const uint8_t* __restrict__ present, int64_t n, int32_t* __restrict__ tmp) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  tmp[v] = present[v] ? uf_find(parent, (int32_t)v) : -1;
}

// ---------------- 1-D two-component GMM (EM) ----------------
constexpr int GMM_BLOCKS = 296;
constexpr int GMM_THREADS = 256;
constexpr int GMM_STATS = 6;  // sum r0, r0*x, r0*x^2, sum x, sum x^2, loglik

struct GmmState {
  double prev_ll;
  int done;
  int iters;
  float params[6];
};

__device__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < GMM_THREADS / 32) t = sh[threadIdx.x];
  if (w == 0) for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;  // valid in thread 0
}

__global__ void __launch_bounds__(GMM_THREADS) k_gmm_estep(const float* __restrict__ x, int64_t n, const GmmState* __restrict__ st,
                                                           double* __restrict__ part, int first) {
  __shared__ double sh[GMM_THREADS / 32];
  if (!first && st->done) return;
  float pi0 = st->params[0], mu0 = st->params[1], v0 = st->params[2];
  float pi1 = st->params[3], mu1 = st->params[4], v1 = st->params[5];
  float c0 = logf(pi0) - 0.5f * logf(6.283185307179586f * v0), c1 = logf(pi1) - 0.5f * logf(6.283185307179586f * v1);
  float i0 = 0.5f / v0, i1 = 0.5f / v1;
  double s[GMM_STATS] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * GMM_THREADS + threadIdx.x; i < n; i += (int64_t)GMM_BLOCKS * GMM_THREADS) {
    float xv = x[i];
    double xd = xv;
    s[3] += xd;
    s[4] += xd * xd;
    if (!first) {
      float l0 = c0 - (xv - mu0) * (xv - mu0) * i0, l1 = c1 - (xv - mu1) * (xv - mu1) * i1;
      float m = fmaxf(l0, l1);
      float ll = m + logf(expf(l0 - m) + expf(l1 - m));
      double r0 = expf(l0 - ll);
      s[0] += r0;
      s[1] += r0 * xd;
      s[2] += r0 * xd * xd;
      s[5] += ll;
    }
  }
  for (int k = 0; k < GMM_STATS; ++k) {
    double t = block_sum(s[k], sh);
    if (threadIdx.x == 0) part[blockIdx.x * GMM_STATS + k] = t;
  }
}

__global__ void k_gmm_mstep(const double* __restrict__ part, int64_t n, GmmState* st, float tol, int first, float* __restrict__ params_out) {
  if (threadIdx.x != 0) return;
  if (!first && st->done) return;
  double s[GMM_STATS] = {0, 0, 0, 0, 0, 0};
  for (int b = 0; b < GMM_BLOCKS; ++b)
    for (int k = 0; k < GMM_STATS; ++k) s[k] += part[b * GMM_STATS + k];
  const double reg = 1e-6, nn = (double)n;
  if (first) {
    double mean = s[3] / nn, var = fmax(s[4] / nn - mean * mean, 1e-12);
    double sd = sqrt(var);
    st->params[0] = 0.5f; st->params[1] = (float)(mean - sd); st->params[2] = (float)(var + reg);
    st->params[3] = 0.5f; st->params[4] = (float)(mean + sd); st->params[5] = (float)(var + reg);
    st->prev_ll = -1e300;
    st->done = 0;
    st->iters = 0;
  } else {
    double n0 = fmax(s[0], 1e-10), n1 = fmax(nn - s[0], 1e-10);
    double m0 = s[1] / n0, m1 = (s[3] - s[1]) / n1;
    double q0 = s[2] / n0 - m0 * m0, q1 = (s[4] - s[2]) / n1 - m1 * m1;
    st->params[0] = (float)(n0 / nn); st->params[1] = (float)m0; st->params[2] = (float)(fmax(q0, 0.0) + reg);
    st->params[3] = (float)(n1 / nn); st->params[4] = (float)m1; st->params[5] = (float)(fmax(q1, 0.0) + reg);
    double ll = s[5] / nn;
    st->iters += 1;
    if (fabs(ll - st->prev_ll) < (double)tol) st->done = 1;
    st->prev_ll = ll;
  }
  for (int k = 0; k < 6; ++k) params_out[k] = st->params[k];
}

}  // namespace

extern "C" size_t hgnn_connected_components_workspace_bytes(int64_t n_vertices) {
  return align_up((size_t)(n_vertices > 0 ? n_vertices : 1), 256) + align_up((size_t)(n_vertices > 0 ? n_vertices : 1) * 4, 256) + 512;
}

extern "C" int hgnn_connected_components(const int64_t* graph, int64_t ld, int64_t n_edges, const uint8_t* keep,
                                         int64_t n_vertices, int32_t* labels, void* ws, size_t ws_bytes, void* stream) {
  if (n_vertices <= 0) return HGNN_OK;
  HGNN_REQUIRE(labels != nullptr && n_vertices < INT32_MAX && n_edges >= 0, "connected_components: bad argument");
  HGNN_REQUIRE(n_edges == 0 || graph != nullptr, "connected_components: graph is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  uint8_t* present = w.take<uint8_t>(n_vertices);
  int32_t* parent = w.take<int32_t>(n_vertices);
  if (!w.ok()) return fail(HGNN_ERR_WORKSPACE, "connected_components: workspace too small");
  unsigned gv = (unsigned)((n_vertices + 255) / 256);
  k_cc_init<<<gv, 256, 0, st>>>(parent, present, n_vertices);
  if (n_edges > 0) k_cc_union<<<(unsigned)((n_edges + 255) / 256), 256, 0, st>>>(graph, ld, n_edges, keep, parent, present);
  k_cc_finalize<<<gv, 256, 0, st>>>(parent, present, n_vertices, labels);
  return check_launch("connected_components");
}

extern "C" size_t hgnn_gmm1d_workspace_bytes(void) {
  return align_up(sizeof(GmmState), 256) + align_up((size_t)GMM_BLOCKS * GMM_STATS * sizeof(double), 256) + 512;
}

extern "C" int hgnn_gmm1d_fit(const float* x, int64_t n, int32_t max_iter, float tol, float* params, void* ws, size_t ws_bytes,
                              void* stream) {
  HGNN_REQUIRE(x && params && n >= 2, "gmm1d_fit: need at least 2 samples");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace w(ws, ws_bytes);
  GmmState* state = w.take<GmmState>(1);
  double* part = w.take<double>((size_t)GMM_BLOCKS * GMM_STATS);
  if (!w.ok()) return fail(HGNN_ERR_WORKSPACE, "gmm1d_fit: workspace too small");
  k_gmm_estep<<<GMM_BLOCKS, GMM_THREADS, 0, st>>>(x, n, state, part, 1);
  k_gmm_mstep<<<1, 32, 0, st>>>(part, n, state, tol, 1, params);
  for (int it = 0; it < max_iter; ++it) {
    k_gmm_estep<<<GMM_BLOCKS, GMM_THREADS, 0, st>>>(x, n, state, part, 0);
    k_gmm_mstep<<<1, 32, 0, st>>>(part, n, state, tol, 0, params);
  }
  return check_launch("gmm1d_fit");
}
