// Tensor-core (tcgen05) fused edge step — placeholder entry points until the
// kernel lands; they report "unsupported" so callers take the fp32 SIMT path.
#include "tc_common.cuh"

using namespace hgnn;

extern "C" int hgnn_tc_supported(int64_t, int64_t, int64_t, int) { return 0; }
extern "C" size_t hgnn_tc_packed_weight_bytes(int64_t out_features, int64_t in_features) {
  return (size_t)out_features * in_features * 2;
}
extern "C" int hgnn_tc_pack_weights(const float*, int64_t, int64_t, void*, void*) {
  return fail(HGNN_ERR_UNSUPPORTED, "tc_pack_weights: tensor-core path not built");
}
extern "C" size_t hgnn_tc_edge_forward_workspace_bytes(int64_t) { return 256; }
extern "C" int hgnn_tc_edge_forward(const hgnn_tc_edge_params*, const float*, const float*, const int32_t*, const int32_t*,
                                    const int32_t*, int64_t, int64_t, float*, void*, size_t, void*) {
  return fail(HGNN_ERR_UNSUPPORTED, "tc_edge_forward: tensor-core path not built");
}
