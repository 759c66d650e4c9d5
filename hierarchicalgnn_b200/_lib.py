"""ctypes binding of libhgnn_b200.so (the C ABI declared in include/hgnn_b200.h).

There is deliberately no fallback: if the shared library is missing or a call
fails, the product path raises. Build with ``python -c "import __graft_entry__
as g; g.build()"`` or ``make -C hierarchicalgnn_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhgnn_b200.so")

MAX_SEGS = 3
MAX_LAYERS = 4

ACT_CODES = {None: 0, "Identity": 0, "GELU": 1, "Tanh": 2, "ReLU": 3, "SiLU": 4, "Sigmoid": 5}

c_f32p = C.c_void_p
c_i32p = C.c_void_p
c_i64p = C.c_void_p
c_vp = C.c_void_p


class MlpDesc(C.Structure):
    _fields_ = [
        ("n_seg", C.c_int32),
        ("n_layers", C.c_int32),
        ("skip_seg", C.c_int32),
        ("ln_eps", C.c_float),
        ("seg_ptr", C.c_void_p * MAX_SEGS),
        ("seg_idx", C.c_void_p * MAX_SEGS),
        ("seg_width", C.c_int32 * MAX_SEGS),
        ("out_width", C.c_int32 * MAX_LAYERS),
        ("act", C.c_int32 * MAX_LAYERS),
        ("W", C.c_void_p * MAX_LAYERS),
        ("b", C.c_void_p * MAX_LAYERS),
        ("gamma", C.c_void_p * MAX_LAYERS),
        ("beta", C.c_void_p * MAX_LAYERS),
        ("out_idx", C.c_void_p),
    ]


class TcEdgeParams(C.Structure):
    _fields_ = [
        ("latent", C.c_int32),
        ("hidden", C.c_int32),
        ("act_hidden", C.c_int32),
        ("act_out", C.c_int32),
        ("ln_eps", C.c_float),
        ("w1_packed", C.c_void_p),
        ("w2_packed", C.c_void_p),
        ("b1", C.c_void_p),
        ("gamma1", C.c_void_p),
        ("beta1", C.c_void_p),
        ("b2", C.c_void_p),
        ("gamma2", C.c_void_p),
        ("beta2", C.c_void_p),
        ("debug_phase_clock", C.c_void_p),
    ]


class TcRowLayer(C.Structure):
    _fields_ = [
        ("n_seg", C.c_int32),
        ("n_out", C.c_int32),
        ("act", C.c_int32),
        ("ln_eps", C.c_float),
        ("seg_ptr", C.c_void_p * MAX_SEGS),
        ("seg_idx", C.c_void_p * MAX_SEGS),
        ("seg_width", C.c_int32 * MAX_SEGS),
        ("w_packed", C.c_void_p),
        ("bias", C.c_void_p),
        ("gamma", C.c_void_p),
        ("beta", C.c_void_p),
        ("skip", C.c_void_p),
    ]


i64, i32, f32, sz, vp = C.c_int64, C.c_int32, C.c_float, C.c_size_t, C.c_void_p

# name -> (restype, argtypes); must list every symbol include/hgnn_b200.h declares
SIGNATURES = {
    "hgnn_abi_version": (C.c_int, []),
    "hgnn_last_error": (C.c_char_p, []),
    "hgnn_csr_build_workspace_bytes": (sz, [i64]),
    "hgnn_csr_build": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, sz, vp]),
    "hgnn_index_to_i32": (C.c_int, [vp, i64, i64, vp, vp, vp]),
    "hgnn_segment_reduce": (C.c_int, [vp, i64, vp, vp, vp, vp, i64, C.c_int, vp, vp]),
    "hgnn_segment_reduce_ld": (C.c_int, [vp, i64, i64, vp, vp, vp, vp, i64, C.c_int, vp, vp]),
    "hgnn_gather_rows": (C.c_int, [vp, i64, vp, vp, i64, vp, vp]),
    "hgnn_edge_dot": (C.c_int, [vp, vp, vp, vp, i64, i64, vp, vp]),
    "hgnn_mlp_forward": (C.c_int, [C.POINTER(MlpDesc), i64, vp, vp]),
    "hgnn_mlp_backward_workspace_bytes": (sz, [C.POINTER(MlpDesc), i64]),
    "hgnn_mlp_backward_data": (C.c_int, [C.POINTER(MlpDesc), i64, vp, C.POINTER(vp * MAX_SEGS),
                                         C.POINTER(vp * MAX_LAYERS), vp, sz, vp]),
    "hgnn_mlp_backward_weights": (C.c_int, [C.POINTER(MlpDesc), i64, C.POINTER(vp * MAX_LAYERS), vp, sz, vp]),
    "hgnn_p2p_all_gather_rows": (C.c_int, [vp, i64, i64, vp, vp, C.c_int, C.c_int, vp]),
    "hgnn_p2p_reduce_scatter_rows": (C.c_int, [vp, i64, i64, vp, vp, C.c_int, C.c_int, vp]),
    "hgnn_p2p_all_reduce": (C.c_int, [i64, vp, vp, C.c_int, C.c_int, vp]),
    "hgnn_knn_radius": (C.c_int, [vp, i64, vp, i64, i64, i64, f32, vp, vp]),
    "hgnn_knn_radius_workspace_bytes": (sz, [i64, i64, i64]),
    "hgnn_knn_radius_ws": (C.c_int, [vp, i64, vp, i64, i64, i64, f32, vp, vp, sz, vp]),
    "hgnn_match_blocks_max": (C.c_int, [vp, vp, vp, i64, vp, i64, vp, C.c_int]),
    "hgnn_knn_radius_batched": (C.c_int, [vp, i64, vp, i64, i64, i64, f32, vp, vp, i64, vp, vp, sz, vp]),
    "hgnn_knn_edges_workspace_bytes": (sz, [i64]),
    "hgnn_knn_edges": (C.c_int, [vp, i64, i64, vp, vp, vp, sz, vp]),
    "hgnn_symmetrize_workspace_bytes": (sz, [i64]),
    "hgnn_symmetrize": (C.c_int, [vp, i64, i64, i64, vp, vp, vp, sz, vp]),
    "hgnn_edge_max_dist": (C.c_int, [vp, vp, vp, vp, i64, i64, vp, vp]),
    "hgnn_connected_components_workspace_bytes": (sz, [i64]),
    "hgnn_connected_components": (C.c_int, [vp, i64, i64, vp, i64, vp, vp, sz, vp]),
    "hgnn_gmm1d_workspace_bytes": (sz, []),
    "hgnn_gmm1d_fit": (C.c_int, [vp, i64, i32, f32, vp, vp, sz, vp]),
    "hgnn_tc_supported": (C.c_int, [i64, i64, i64, C.c_int, C.c_int, C.c_int]),
    "hgnn_tc_packed_weight_bytes": (sz, [i64, i64]),
    "hgnn_tc_pack_weights": (C.c_int, [vp, i64, i64, vp, vp]),
    "hgnn_tc_debug_gemm": (C.c_int, [vp, vp, i64, i64, i64, vp, vp]),
    "hgnn_tc_debug_wgrad_workspace_bytes": (sz, [i64, i64, i64]),
    "hgnn_tc_debug_wgrad": (C.c_int, [vp, vp, i64, i64, i64, vp, vp, sz, vp]),
    "hgnn_tc_edge_forward_workspace_bytes": (sz, [i64, i64, i64]),
    "hgnn_tc_edge_backward_workspace_bytes": (sz, [i64, i64]),
    # (p, w1t, w2t, wx, stash, x, n_nodes, dst, perm, src_rows, src_rowptr, dst_rows, dst_rowptr, n_edges, g_e, g_agg, d_e, d_x,
    #  dW1, dW2, dv1, dv2, ws, ws_bytes, stream, aux_stream)
    "hgnn_tc_edge_backward": (C.c_int, [C.POINTER(TcEdgeParams), vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp,
                                        vp, vp, vp, vp, vp, sz, vp, vp]),
    "hgnn_narrow_in_supported": (C.c_int, [C.POINTER(MlpDesc)]),
    "hgnn_narrow_in_forward": (C.c_int, [C.POINTER(MlpDesc), i64, vp, vp]),
    "hgnn_narrow_in_backward_workspace_bytes": (sz, [i64]),
    "hgnn_narrow_in_backward": (C.c_int, [C.POINTER(MlpDesc), i64, vp, vp, vp, vp, vp, sz, vp]),
    "hgnn_narrow_out_supported": (C.c_int, [i64, i64]),
    "hgnn_narrow_out_forward": (C.c_int, [vp, i64, i64, vp, vp, i64, vp, vp]),
    "hgnn_narrow_out_backward_workspace_bytes": (sz, [i64, i64]),
    "hgnn_narrow_out_backward": (C.c_int, [vp, i64, i64, vp, i64, vp, vp, vp, vp, vp, sz, vp]),
    "hgnn_tc_gemm_supported": (C.c_int, [C.POINTER(TcRowLayer)]),
    "hgnn_tc_gemm": (C.c_int, [C.POINTER(TcRowLayer), i64, vp, i64, i64, vp, vp]),
    "hgnn_ln_act_supported": (C.c_int, [i64]),
    "hgnn_ln_act_forward": (C.c_int, [vp, i64, i64, vp, vp, f32, C.c_int, vp, vp, vp]),
    "hgnn_ln_act_backward_workspace_bytes": (sz, [i64]),
    "hgnn_ln_act_backward": (C.c_int, [vp, vp, i64, i64, vp, vp, f32, C.c_int, vp, vp, vp, sz, vp]),
    "hgnn_tc_wgrad_supported": (C.c_int, [i64, i64]),
    "hgnn_tc_wgrad_workspace_bytes": (sz, [i64, i64, i64]),
    "hgnn_tc_wgrad": (C.c_int, [vp, i64, vp, i64, i64, vp, vp, sz, vp]),
    "hgnn_tc_row_supported": (C.c_int, [C.POINTER(TcRowLayer)]),
    "hgnn_tc_row_image_bytes": (sz, [i64, i64]),
    "hgnn_tc_row_forward": (C.c_int, [C.POINTER(TcRowLayer), i64, vp, vp, vp]),
    "hgnn_tc_row_backward_workspace_bytes": (sz, [i64, i64, i64]),
    "hgnn_tc_row_backward": (C.c_int, [C.POINTER(TcRowLayer), vp, vp, i64, vp, vp, vp, vp, vp, sz, vp]),
    "hgnn_tc_row_backward_split": (C.c_int, [C.POINTER(TcRowLayer), vp, vp, i64, vp, C.POINTER(vp), vp, vp, vp, sz, vp]),
    "hgnn_tc_edge_stash_bytes": (sz, [i64, i64]),
    "hgnn_tc_edge_forward": (C.c_int, [C.POINTER(TcEdgeParams), vp, vp, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, sz, vp]),
}

_lib = None


class HgnnError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HgnnError(
                f"{LIB_PATH} is missing: the CUDA extension is not built and this package has no "
                "CPU/PyTorch fallback. Run `make -C hierarchicalgnn_b200/csrc` (needs nvcc, sm_100a).")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        if handle.hgnn_abi_version() != 1:
            raise HgnnError("libhgnn_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().hgnn_last_error()
        raise HgnnError(f"{what or 'hgnn call'} failed ({rc}): {msg.decode() if msg else ''}")
