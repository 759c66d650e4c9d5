"""PyTorch-facing operators over the C ABI (include/hgnn_b200.h).

Every function here launches hand-written sm_100a kernels from
libhgnn_b200.so on ``torch.cuda.current_stream()``; torch only owns the device
memory and the autograd graph. There is no CPU / eager fallback: CPU tensors
raise.

Reference call sites replaced (under /root/reference/Modules):
  scatter_add / scatter_mean     gnn_utils.py:50,124-125,142-143; BC/Models/HGNN_GMM.py:251,269
  gathered-concat MLP + skip     gnn_utils.py:45-64,119-153 (edge/node/supernode/superedge updates)
  find_neighbors (frnn)          utils.py:228-239
  symmetrize (cugraph)           gnn_utils.py:198-199
  einsum('ij,ij->i') edge dots   gnn_utils.py:208; BC/Models/HGNN_GMM.py:188
  connected components + GMM     BC/Models/HGNN_GMM.py:184-234
"""
from __future__ import annotations

import ctypes as C
import functools
import os
from collections import OrderedDict
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import MAX_LAYERS, MAX_SEGS, ACT_CODES, MlpDesc, check

Tensor = torch.Tensor

# kernel-launch counter (bench.py reports it as gpu_launches)
LAUNCHES = {"count": 0}


def _count(n=1):
    LAUNCHES["count"] += n


# optional per-call CUDA-event profile: name -> [(start, end)], filled when not None
PROFILE = None


class _timed:
    """Brackets one C-ABI call with CUDA events on the launch stream (bench.py)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.b.record()
            PROFILE.setdefault(self.name, []).append((self.a, self.b))
        return False


# Precision of the MLP contractions: "fp32" = SIMT FFMA everywhere (bit-level parity mode);
# "bf16" / "auto" = tcgen05 tensor cores (bf16 operands, fp32 accumulate, fp32 storage) wherever
# a tensor-core kernel exists for the shape, fp32 SIMT elsewhere.
_PRECISION = {"mode": os.environ.get("HGNN_PRECISION", "auto")}
TC_CALLS = {"count": 0}       # fused tensor-core edge-step launches (forward / backward)
TC_ROW_CALLS = {"count": 0}   # tensor-core row-layer launches


def set_precision(mode: str) -> str:
    if mode not in ("auto", "fp32", "bf16"):
        raise _lib.HgnnError("precision must be one of auto / fp32 / bf16")
    old = _PRECISION["mode"]
    _PRECISION["mode"] = mode
    return old


def get_precision() -> str:
    return _PRECISION["mode"]


def compute_dtype(module=None) -> str:
    """Arithmetic type of the MLP contractions on the path that actually ran."""
    return "bf16" if (TC_CALLS["count"] or TC_ROW_CALLS["count"]) else "f32"


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (every launch goes there). The raw getter avoids
    building a torch.cuda.Stream object per call (18 us -> <1 us: ~1300 calls per BC-HGNN step)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


_AUX_STREAMS = {}


def _aux_stream(device) -> Optional[int]:
    """A second stream per device for the calls that can run two independent launch chains side by side (the backward of the
    edge step: per-edge weight-gradient GEMM beside the node-level chain). The C call forks and joins by itself; HGNN_AUX_STREAM=0
    disables it."""
    if os.environ.get("HGNN_AUX_STREAM", "1") == "0":
        return None
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _AUX_STREAMS.get(idx)
    if st is None:
        st = _AUX_STREAMS[idx] = torch.cuda.Stream(device=idx)
    return st.cuda_stream


def _need_cuda(*tensors):
    """Every launch goes to the CURRENT device's current stream and workspaces are allocated beside the tensors, so all
    tensor arguments must live on the current device (one process per GPU; use torch.cuda.set_device / torch.cuda.device)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.HgnnError("hierarchicalgnn_b200 ops need CUDA tensors (no CPU fallback); got a %s tensor" % t.device)
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise _lib.HgnnError("tensor on %s but the current CUDA device is cuda:%d: kernels launch on the current device's "
                                 "stream (call torch.cuda.set_device(%d) or wrap the call in torch.cuda.device)"
                                 % (t.device, cur, t.device.index))


def _f32(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        raise _lib.HgnnError(f"expected float32 features, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _save_with_scratch(ctx, tensors, scratch: Optional[Tensor]):
    """save_for_backward(*tensors [, scratch]). Kernel scratch a backward needs (the edge-step stash, tile images) goes
    through autograd's saved-tensor mechanism rather than a ctx attribute, so it is released with the graph's buffers
    right after the backward has run even when something — a logged loss, an in-place copy of a metric — keeps the
    graph NODES alive (1.5 KB per edge per cell otherwise stays pinned for as long as those references live)."""
    ctx.has_scratch = scratch is not None
    if scratch is not None:
        ctx.save_for_backward(*tensors, scratch)
    else:
        ctx.save_for_backward(*tensors)


def _saved_and_scratch(ctx):
    saved = ctx.saved_tensors
    if ctx.has_scratch:
        return saved[:-1], saved[-1]
    return saved, None


def _workspace(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# ---------------------------------------------------------------------------
# segment plans (destination-sorted CSR), cached per index tensor
# ---------------------------------------------------------------------------
class SegmentPlan:
    """CSR over an int64 key vector: items ordered by (key, item id)."""

    def __init__(self, keys: Tensor, n_segments: int):
        _need_cuda(keys)
        if keys.dtype != torch.int64 or keys.dim() != 1:
            raise _lib.HgnnError("SegmentPlan needs a 1-D int64 key tensor")
        keys = keys.contiguous()
        n = keys.numel()
        self.n_items = n
        self.n_segments = int(n_segments)
        if n and self.n_segments <= 0:
            raise _lib.HgnnError(f"SegmentPlan: {n} keys but n_segments = {n_segments}")
        if n and os.environ.get("HGNN_CHECK_INDICES", "0") != "0":  # one host sync per plan: off by default (the kernel clamps)
            lo, hi = int(keys.min()), int(keys.max())
            if lo < 0 or hi >= self.n_segments:
                raise _lib.HgnnError(f"SegmentPlan: index out of range: keys span [{lo}, {hi}], n_segments = {self.n_segments} "
                                     "(torch_scatter would raise here; without HGNN_CHECK_INDICES the index is clamped)")
        dev = keys.device
        self.perm = torch.empty(n, dtype=torch.int32, device=dev)
        self.rowptr = torch.empty(self.n_segments + 1, dtype=torch.int32, device=dev)
        self.keys32 = torch.empty(n, dtype=torch.int32, device=dev)
        L = _lib.lib()
        ws = _workspace(L.hgnn_csr_build_workspace_bytes(n), dev)
        check(L.hgnn_csr_build(_ptr(keys), n, self.n_segments, _ptr(self.perm), _ptr(self.rowptr), _ptr(self.keys32),
                               _ptr(ws), ws.numel(), _stream()), "csr_build")
        _count(4)
        self._counts_inv = None
        self._identity = None

    def is_identity(self) -> bool:
        """True when the items already arrive segment-sorted (perm == arange): kernels then stream rows in place
        instead of following perm (one host sync, cached)."""
        if self._identity is None:
            k = self.keys32
            self._identity = bool(self.n_items < 2 or bool((k[1:] >= k[:-1]).all()))
        return self._identity

    def inv_counts(self) -> Tensor:
        if self._counts_inv is None:
            cnt = (self.rowptr[1:] - self.rowptr[:-1]).clamp(min=1).to(torch.float32)
            self._counts_inv = 1.0 / cnt
        return self._counts_inv


_PLAN_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()
_PLAN_CACHE_MAX = 32


def plan_for(keys: Tensor, n_segments: int) -> SegmentPlan:
    """Cached SegmentPlan for an index tensor. The cache entry keeps the key
    tensor alive, so its storage address cannot be recycled while cached, and is
    invalidated by in-place modification (version counter)."""
    k = (keys.data_ptr(), keys.numel(), int(n_segments), keys._version, keys.stride(0) if keys.numel() else 1)
    hit = _PLAN_CACHE.get(k)
    if hit is not None:
        _PLAN_CACHE.move_to_end(k)
        return hit[1]
    plan = SegmentPlan(keys, n_segments)
    _PLAN_CACHE[k] = (keys, plan)
    while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
        _PLAN_CACHE.popitem(last=False)
    return plan


def clear_plan_cache():
    _PLAN_CACHE.clear()


# ---------------------------------------------------------------------------
# raw launches
# ---------------------------------------------------------------------------
def _rows_view(src: Tensor):
    """(tensor, row stride in floats) for the kernels that take a source row stride: a column slice of a wider fp32 matrix
    (unit column stride) is used in place, anything else is made dense."""
    if src.dtype != torch.float32:
        raise _lib.HgnnError(f"expected float32 features, got {src.dtype}")
    if src.dim() == 2 and src.stride(1) == 1 and src.stride(0) >= src.shape[1] and src.shape[0] > 0:
        return src, src.stride(0)
    src = src.contiguous()
    return src, src.shape[1]


def segment_reduce_raw(src: Tensor, plan: SegmentPlan, gather32: Optional[Tensor] = None,
                       weight: Optional[Tensor] = None, mean: bool = False) -> Tensor:
    src, ld = _rows_view(src)
    width = src.shape[1]
    if plan.n_items == 0:
        return torch.zeros((plan.n_segments, width), dtype=torch.float32, device=src.device)
    out = torch.empty((plan.n_segments, width), dtype=torch.float32, device=src.device)
    if plan.n_segments and width:
        with _timed("segment_reduce"):
            check(_lib.lib().hgnn_segment_reduce_ld(_ptr(src), width, ld, _ptr(gather32), _ptr(weight), _ptr(plan.perm),
                                                    _ptr(plan.rowptr), plan.n_segments, int(mean), _ptr(out), _stream()),
                  "segment_reduce")
        _count(2)  # per-thread kernel for ordinary segments + per-CTA kernel for hub segments
    return out


def gather_rows_raw(src: Tensor, idx32: Optional[Tensor], weight: Optional[Tensor], n_items: int) -> Tensor:
    src = _f32(src)
    width = src.shape[1]
    out = torch.empty((n_items, width), dtype=torch.float32, device=src.device)
    if n_items and width:
        with _timed("gather_rows"):
            check(_lib.lib().hgnn_gather_rows(_ptr(src), width, _ptr(idx32), _ptr(weight), n_items, _ptr(out), _stream()),
                  "gather_rows")
        _count()
    return out


def edge_dot_raw(a: Tensor, ai32: Optional[Tensor], b: Tensor, bi32: Optional[Tensor], n_items: int) -> Tensor:
    a, b = _f32(a), _f32(b)
    out = torch.empty(n_items, dtype=torch.float32, device=a.device)
    if n_items:
        check(_lib.lib().hgnn_edge_dot(_ptr(a), _ptr(ai32), _ptr(b), _ptr(bi32), a.shape[1], n_items, _ptr(out), _stream()),
              "edge_dot")
        _count()
    return out


# ---------------------------------------------------------------------------
# autograd: weighted gather -> segmented sum
# ---------------------------------------------------------------------------
class _GatherScatter(torch.autograd.Function):
    """out[s] = sum_{i: seg[i]=s} w[i] * src[gather[i]]  (gather / w optional)."""

    @staticmethod
    def forward(ctx, src, weight, gather_plan, seg_plan, mean):
        _need_cuda(src, weight)
        src = _f32(src)
        w = None if weight is None else _f32(weight).reshape(-1)
        g32 = None if gather_plan is None else gather_plan.keys32
        out = segment_reduce_raw(src, seg_plan, g32, w, mean)
        ctx.save_for_backward(src, w)
        ctx.plans = (gather_plan, seg_plan)
        ctx.mean = mean
        ctx.wshape = None if weight is None else weight.shape
        return out

    @staticmethod
    def backward(ctx, gout):
        src, w = ctx.saved_tensors
        gather_plan, seg_plan = ctx.plans
        gout = _f32(gout)
        d_src = d_w = None
        w_eff = w
        if ctx.mean:
            inv = seg_plan.inv_counts()[seg_plan.keys32.long()]
            w_eff = inv if w is None else w * inv
        if ctx.needs_input_grad[0]:
            if gather_plan is None:
                d_src = gather_rows_raw(gout, seg_plan.keys32, w_eff, seg_plan.n_items)
            else:
                d_src = segment_reduce_raw(gout, gather_plan, seg_plan.keys32, w_eff, False)
        if w is not None and ctx.needs_input_grad[1]:
            g32 = None if gather_plan is None else gather_plan.keys32
            d_w = edge_dot_raw(src, g32, gout, seg_plan.keys32, seg_plan.n_items)
            if ctx.mean:
                d_w = d_w * seg_plan.inv_counts()[seg_plan.keys32.long()]
            d_w = d_w.reshape(ctx.wshape)
        return d_src, d_w, None, None, None


def scatter_add(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
                plan: Optional[SegmentPlan] = None) -> Tensor:
    """Drop-in for torch_scatter.scatter_add(src, index, dim=0, dim_size=n) on 2-D rows."""
    if dim != 0 or src.dim() != 2:
        raise _lib.HgnnError("scatter_add: only dim=0 on [items, width] rows is supported")
    if plan is None:
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() else 0
        plan = plan_for(index, int(dim_size))
    return _GatherScatter.apply(src, None, None, plan, False)


def scatter_mean(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
                 plan: Optional[SegmentPlan] = None) -> Tensor:
    if dim != 0 or src.dim() != 2:
        raise _lib.HgnnError("scatter_mean: only dim=0 on [items, width] rows is supported")
    if plan is None:
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() else 0
        plan = plan_for(index, int(dim_size))
    return _GatherScatter.apply(src, None, None, plan, True)


def gather_scatter(src: Tensor, weight: Optional[Tensor], gather_plan: SegmentPlan, seg_plan: SegmentPlan) -> Tensor:
    """scatter_add(weight * src[gather], seg) without materialising the product
    (gnn_utils.py:124,142; BC/Models/HGNN_GMM.py:269). ``gather_plan`` is the plan
    over the gather index (n_segments = src rows), ``seg_plan`` over the segment index."""
    return _GatherScatter.apply(src, weight, gather_plan, seg_plan, False)


# ---------------------------------------------------------------------------
# autograd: value / mean of its segment
# ---------------------------------------------------------------------------
class _SegmentMeanNormalize(torch.autograd.Function):
    """out[i] = w[i] / mean(w over the segment of i). The per-event form of ``edge_weights / edge_weights.mean()``
    (gnn_utils.py:213-214) for a batch of events; the adjoint is two ordered segment sums (autograd through
    ``mean[segment]`` would scatter-add hundreds of thousands of gradients into a handful of rows)."""

    @staticmethod
    def forward(ctx, w, plan: SegmentPlan):
        w = _f32(w).reshape(-1)
        seg = plan.keys32.long()
        inv_n = plan.inv_counts()
        mean = segment_reduce_raw(w.unsqueeze(1), plan)[:, 0] * inv_n
        inv_mean = 1.0 / mean
        ctx.save_for_backward(w, inv_mean, inv_n, seg)
        ctx.plan = plan
        return w * inv_mean[seg]

    @staticmethod
    def backward(ctx, g):
        w, inv_mean, inv_n, seg = ctx.saved_tensors
        g = _f32(g).reshape(-1)
        t = segment_reduce_raw((g * w).unsqueeze(1), ctx.plan)[:, 0]  # sum_i g_i w_i per segment
        return g * inv_mean[seg] - (t * inv_n * inv_mean * inv_mean)[seg], None


def segment_mean_normalize(w: Tensor, segment: Tensor, n_segments: int) -> Tensor:
    """w / (mean of w over its segment); ``segment``: int64 segment id per element."""
    return _SegmentMeanNormalize.apply(w, plan_for(segment, n_segments))


# ---------------------------------------------------------------------------
# autograd: gathered row dot
# ---------------------------------------------------------------------------
class _EdgeDot(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, plan_a, plan_b):
        _need_cuda(a, b)
        a, b = _f32(a), _f32(b)
        out = edge_dot_raw(a, plan_a.keys32, b, plan_b.keys32, plan_a.n_items)
        ctx.save_for_backward(a, b)
        ctx.plans = (plan_a, plan_b)
        return out

    @staticmethod
    def backward(ctx, gout):
        a, b = ctx.saved_tensors
        plan_a, plan_b = ctx.plans
        gout = _f32(gout)
        d_a = d_b = None
        if ctx.needs_input_grad[0]:
            d_a = segment_reduce_raw(b, plan_a, plan_b.keys32, gout, False)
        if ctx.needs_input_grad[1]:
            d_b = segment_reduce_raw(a, plan_b, plan_a.keys32, gout, False)
        return d_a, d_b, None, None


def edge_dot(a: Tensor, b: Tensor, plan_a: SegmentPlan, plan_b: SegmentPlan) -> Tensor:
    """out[i] = <a[ia[i]], b[ib[i]]> where plan_a / plan_b are the plans over ia / ib."""
    return _EdgeDot.apply(a, b, plan_a, plan_b)


# ---------------------------------------------------------------------------
# fused gathered-concat MLP
# ---------------------------------------------------------------------------
class MlpMeta:
    """Static description of one fused MLP call."""

    def __init__(self, seg_plans: Sequence[Optional[SegmentPlan]], acts: Sequence[Optional[str]],
                 has_ln: Sequence[bool], skip_seg: int = -1, eps: float = 1e-5, tc_pack=None):
        self.tc_pack = tc_pack  # callable -> (w1_packed, w2_packed) when the tensor-core edge kernel applies
        self.seg_plans = list(seg_plans)
        self.acts = [ACT_CODES[a] for a in acts]
        self.has_ln = list(has_ln)
        self.skip_seg = int(skip_seg)
        self.eps = float(eps)
        if len(self.seg_plans) > MAX_SEGS or len(self.acts) > MAX_LAYERS:
            raise _lib.HgnnError("fused MLP supports at most %d segments and %d layers" % (MAX_SEGS, MAX_LAYERS))


def _build_desc(meta: MlpMeta, segs: List[Tensor], params: List[Tensor]):
    d = MlpDesc()
    d.n_seg = len(segs)
    d.n_layers = len(meta.acts)
    d.skip_seg = meta.skip_seg
    d.ln_eps = meta.eps
    rows = None
    for s, t in enumerate(segs):
        d.seg_ptr[s] = t.data_ptr()
        d.seg_width[s] = t.shape[1]
        plan = meta.seg_plans[s]
        if plan is not None:
            d.seg_idx[s] = plan.keys32.data_ptr()
            n = plan.n_items
        else:
            d.seg_idx[s] = None
            n = t.shape[0]
        if rows is None:
            rows = n
        elif rows != n:
            raise _lib.HgnnError(f"fused MLP: segment {s} yields {n} rows, expected {rows}")
    i = 0
    layer_params = []
    for l in range(d.n_layers):
        W, b = params[i], params[i + 1]
        i += 2
        g = bt = None
        if meta.has_ln[l]:
            g, bt = params[i], params[i + 1]
            i += 2
        d.W[l], d.b[l] = W.data_ptr(), b.data_ptr()
        d.gamma[l] = None if g is None else g.data_ptr()
        d.beta[l] = None if bt is None else bt.data_ptr()
        d.out_width[l] = W.shape[0]
        d.act[l] = meta.acts[l]
        layer_params.append((W, b, g, bt))
    d.out_idx = None
    return d, rows, layer_params


class _FusedMLP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, meta: MlpMeta, n_seg: int, *tensors):
        _need_cuda(*tensors)
        segs = [_f32(t) for t in tensors[:n_seg]]
        params = [_f32(t) for t in tensors[n_seg:]]
        d, rows, layers = _build_desc(meta, segs, params)
        out = torch.empty((rows, layers[-1][0].shape[0]), dtype=torch.float32, device=segs[0].device)
        stash = None
        if rows and meta.tc_pack is not None:
            stash = tc_edge_forward_raw(meta, segs, layers, out,
                                        save_image=any(ctx.needs_input_grad[2:]) and tc_backward_available(meta, layers))
        elif rows:
            with _timed("mlp_forward"):
                check(_lib.lib().hgnn_mlp_forward(C.byref(d), rows, _ptr(out), _stream()), "mlp_forward")
            _count()
        ctx.meta, ctx.n_seg = meta, n_seg
        _save_with_scratch(ctx, [*segs, *params], stash)
        return out

    @staticmethod
    def backward(ctx, gout):
        meta, n_seg = ctx.meta, ctx.n_seg
        saved, stash = _saved_and_scratch(ctx)
        segs, params = list(saved[:n_seg]), list(saved[n_seg:])
        d, rows, layers = _build_desc(meta, segs, params)
        dev = segs[0].device
        gout = _f32(gout)
        if rows and tc_backward_available(meta, layers) and stash is not None:
            d_x, d_e, dW1, dW2, dv1, dv2 = tc_edge_backward_raw(meta, segs, layers, gout, stash)
            need = ctx.needs_input_grad[2:]
            # segments 0 and 1 are the same tensor (tc path precondition): its whole gradient goes out through segment 0
            return (None, None, d_x if need[0] else None, None, d_e if need[2] else None,
                    dW1, dv1[0], dv1[1], dv1[2], dW2, dv2[0], dv2[1], dv2[2])
        L = _lib.lib()
        need = ctx.needs_input_grad[2:]
        dseg_rows: List[Optional[Tensor]] = []
        dseg_arr = (C.c_void_p * MAX_SEGS)()
        for s in range(n_seg):
            if need[s]:
                t = torch.empty((rows, segs[s].shape[1]), dtype=torch.float32, device=dev)
                dseg_rows.append(t)
                dseg_arr[s] = t.data_ptr()
            else:
                dseg_rows.append(None)
                dseg_arr[s] = None
        dvec, dW = [], []
        dvec_arr = (C.c_void_p * MAX_LAYERS)()
        dW_arr = (C.c_void_p * MAX_LAYERS)()
        for l, (W, b, g, bt) in enumerate(layers):
            v = torch.empty((3, W.shape[0]), dtype=torch.float32, device=dev)
            w = torch.empty_like(W)
            dvec.append(v)
            dW.append(w)
            dvec_arr[l] = v.data_ptr()
            dW_arr[l] = w.data_ptr()
        if rows:
            nbytes = L.hgnn_mlp_backward_workspace_bytes(C.byref(d), rows)
            ws = _workspace(nbytes, dev)
            with _timed("mlp_backward_data"):
                check(L.hgnn_mlp_backward_data(C.byref(d), rows, _ptr(gout), C.byref(dseg_arr), C.byref(dvec_arr), _ptr(ws),
                                               ws.numel(), _stream()), "mlp_backward_data")
            with _timed("mlp_backward_weights"):
                check(L.hgnn_mlp_backward_weights(C.byref(d), rows, C.byref(dW_arr), _ptr(ws), ws.numel(), _stream()),
                      "mlp_backward_weights")
            _count(1 + len(layers) + 2 * len(layers))
        else:
            for v in dvec:
                v.zero_()
            for w in dW:
                w.zero_()
        grads: List[Optional[Tensor]] = [None, None]
        for s in range(n_seg):
            g = dseg_rows[s]
            if g is not None and meta.seg_plans[s] is not None:
                plan = meta.seg_plans[s]
                if plan.n_segments != segs[s].shape[0]:
                    raise _lib.HgnnError("fused MLP: gather plan does not cover the gathered tensor")
                g = segment_reduce_raw(g, plan)
            grads.append(g)
        for l, (W, b, gm, bt) in enumerate(layers):
            grads += [dW[l], dvec[l][0]]
            if gm is not None:
                grads += [dvec[l][1], dvec[l][2]]
        return tuple(grads)


def tc_supported(latent: int, hidden: int, n_layers: int, layer_norm: bool, act_hidden="GELU", act_out="Tanh") -> bool:
    return bool(_lib.lib().hgnn_tc_supported(int(latent), int(hidden), int(n_layers), int(bool(layer_norm)),
                                             ACT_CODES.get(act_hidden, -1), ACT_CODES.get(act_out, -1)))


def tc_pack_weight(W: Tensor) -> Tensor:
    """fp32 nn.Linear weight [out, in] -> bf16 K-major 128B-swizzled UMMA image (uint8 tensor)."""
    _need_cuda(W)
    W = _f32(W.detach())
    out = torch.empty(_lib.lib().hgnn_tc_packed_weight_bytes(W.shape[0], W.shape[1]), dtype=torch.uint8, device=W.device)
    check(_lib.lib().hgnn_tc_pack_weights(_ptr(W), W.shape[0], W.shape[1], _ptr(out), _stream()), "tc_pack_weights")
    _count()
    return out


def tc_pack_weight_t(W: Tensor) -> Tensor:
    """UMMA image of W^T ([in, out] rows), the B operand of the data-gradient GEMMs. in-features > 256
    (W1^T: 3L rows) exceed one MMA's N and are consumed in 128-row pieces, so pack in row chunks."""
    Wt = W.detach().t().contiguous()
    if Wt.shape[0] <= 256:
        return tc_pack_weight(Wt)
    # image layout is per K-block [rows x 128 B]; build it from <=256-row packs and interleave per K-block
    rows, cols = Wt.shape
    nkb = cols // 64
    parts = [tc_pack_weight(Wt[r0:r0 + 128].contiguous()).reshape(nkb, 128 * 128) for r0 in range(0, rows, 128)]
    return torch.cat(parts, dim=1).reshape(-1).contiguous()


def tc_debug_gemm(A: Tensor, W: Tensor) -> Tensor:
    A = _f32(A)
    Wp = tc_pack_weight(W)
    out = torch.empty((A.shape[0], W.shape[0]), dtype=torch.float32, device=A.device)
    check(_lib.lib().hgnn_tc_debug_gemm(_ptr(A), _ptr(Wp), A.shape[0], W.shape[0], W.shape[1], _ptr(out), _stream()),
          "tc_debug_gemm")
    _count()
    return out


def tc_debug_wgrad(A: Tensor, B: Tensor) -> Tensor:
    A, B = _f32(A), _f32(B)
    L = _lib.lib()
    ws = _workspace(L.hgnn_tc_debug_wgrad_workspace_bytes(A.shape[0], A.shape[1], B.shape[1]), A.device)
    out = torch.empty((A.shape[1], B.shape[1]), dtype=torch.float32, device=A.device)
    check(L.hgnn_tc_debug_wgrad(_ptr(A), _ptr(B), A.shape[0], A.shape[1], B.shape[1], _ptr(out), _ptr(ws), ws.numel(), _stream()),
          "tc_debug_wgrad")
    _count(4)
    return out


DEBUG_PHASE_CLOCK = {"ptr": None}  # host-side switch of the per-call profiling pointer in hgnn_tc_edge_params


def _tc_params(meta: MlpMeta, layers, w1p, w2p):
    (W1, b1, g1, be1), (W2, b2, g2, be2) = layers
    p = _lib.TcEdgeParams()
    p.latent, p.hidden = W2.shape[0], W1.shape[0]
    p.act_hidden, p.act_out = meta.acts[0], meta.acts[1]
    p.ln_eps = meta.eps
    p.w1_packed, p.w2_packed = w1p.data_ptr(), w2p.data_ptr()
    p.b1, p.gamma1, p.beta1 = b1.data_ptr(), g1.data_ptr(), be1.data_ptr()
    p.b2, p.gamma2, p.beta2 = b2.data_ptr(), g2.data_ptr(), be2.data_ptr()
    p.debug_phase_clock = DEBUG_PHASE_CLOCK["ptr"]  # None in production; profiles/bwd_phase_clock.py sets it per call
    return p


def tc_backward_available(meta: MlpMeta, layers) -> bool:
    return meta.tc_pack is not None and layers[1][0].shape[0] == 128 and layers[0][0].shape[0] == 256


def _tile_row_plans(plan_s: SegmentPlan, plan_d: SegmentPlan, dst_sorted: bool):
    """CSR over the forward kernel's TILE ROWS grouped by source / by destination node, for the per-node sums of delta1
    in the backward: (src_rows, src_rowptr, dst_rows | None, dst_rowptr). Tile row j holds edge plan_d.perm[j] when the
    forward ran destination-sorted (the fused-aggregate path), edge j otherwise. Cached on the source plan."""
    key = (id(plan_d), bool(dst_sorted))
    cache = plan_s.__dict__.setdefault("_tile_rows", {})
    hit = cache.get(key)
    if hit is None:
        if not dst_sorted:
            hit = (plan_s.perm, plan_s.rowptr, plan_d.perm, plan_d.rowptr, plan_d)
        elif plan_d.is_identity():
            hit = (plan_s.perm, plan_s.rowptr, None, plan_d.rowptr, plan_d)
        else:
            pos = torch.empty_like(plan_d.perm)  # edge id -> tile row
            pos[plan_d.perm.long()] = torch.arange(plan_d.n_items, dtype=torch.int32, device=pos.device)
            hit = (pos[plan_s.perm.long()].contiguous(), plan_s.rowptr, None, plan_d.rowptr, plan_d)
        cache.clear()  # one entry: the plans of one graph are used together
        cache[key] = hit
    return hit[:4]


def tc_edge_backward_raw(meta: MlpMeta, segs, layers, gout: Tensor, stash: Tensor, perm: Optional[Tensor] = None,
                         grad_agg: Optional[Tensor] = None, dst_sorted: bool = False):
    """Backward of the tensor-core edge step from the forward's stash (same row order ``perm``; ``dst_sorted`` says that
    order is the by-destination plan's). Returns (d_x, d_e, dW1, dW2, dvec1, dvec2): d_x is the complete node gradient
    (both gathers), computed per node from the segment sums of delta1."""
    x, e = segs[0], segs[2]
    plan_s, plan_d = meta.seg_plans[0], meta.seg_plans[1]
    w1p, w2p, w1tp, w2tp, wxp = meta.tc_pack()
    p = _tc_params(meta, layers, w1p, w2p)
    E, Lw = e.shape
    N = x.shape[0]
    dev = e.device
    d_e = torch.empty_like(e)
    d_x = torch.empty((N, Lw), dtype=torch.float32, device=dev)
    dW1, dW2 = torch.empty_like(layers[0][0]), torch.empty_like(layers[1][0])
    dv1 = torch.empty((3, layers[0][0].shape[0]), dtype=torch.float32, device=dev)
    dv2 = torch.empty((3, layers[1][0].shape[0]), dtype=torch.float32, device=dev)
    src_rows, src_rowptr, dst_rows, dst_rowptr = _tile_row_plans(plan_s, plan_d, dst_sorted)
    L_ = _lib.lib()
    ws = _workspace(L_.hgnn_tc_edge_backward_workspace_bytes(E, N), dev)
    with _timed("tc_edge_backward"):
        check(L_.hgnn_tc_edge_backward(C.byref(p), _ptr(w1tp), _ptr(w2tp), _ptr(wxp), _ptr(stash), _ptr(x), N,
                                       _ptr(plan_d.keys32), _ptr(perm), _ptr(src_rows), _ptr(src_rowptr), _ptr(dst_rows),
                                       _ptr(dst_rowptr), E, _ptr(gout), _ptr(grad_agg), _ptr(d_e), _ptr(d_x),
                                       _ptr(dW1), _ptr(dW2), _ptr(dv1), _ptr(dv2), _ptr(ws), ws.numel(), _stream(), _aux_stream(dev)),
              "tc_edge_backward")
    # data-gradient kernel, column-sum reduce, delta1 node sums (2) + their column sum (2), d(x) GEMM, x image,
    # 2 x (weight-gradient GEMM + ordered reduce)
    _count(12)
    TC_CALLS["count"] += 1
    return d_x, d_e, dW1, dW2, dv1, dv2


def tc_edge_forward_raw(meta: MlpMeta, segs, layers, out: Tensor, agg: Optional[Tensor] = None, save_image: bool = False):
    """e' = MLP([x[src] | x[dst] | e]) + e on tcgen05 tensor cores (segments: x|by_src, x|by_dst, e); with ``agg``
    the same launch also leaves scatter_add(e', dst) there (edges visited in destination-sorted order)."""
    x, e = segs[0], segs[2]
    plan_s, plan_d = meta.seg_plans[0], meta.seg_plans[1]
    w1p, w2p = meta.tc_pack()[:2]
    p = _tc_params(meta, layers, w1p, w2p)
    n_edges = e.shape[0]
    perm = rowptr = None
    if agg is not None:
        perm, rowptr = (None if plan_d.is_identity() else plan_d.perm), plan_d.rowptr
    a0 = None
    if save_image:  # the forward's stash for the backward pass: operand images + bf16 xhat's + rstd (1.5 KB/edge at L = 128)
        a0 = _workspace(_lib.lib().hgnn_tc_edge_stash_bytes(n_edges, e.shape[1]), e.device)
    ws = _workspace(_lib.lib().hgnn_tc_edge_forward_workspace_bytes(n_edges, x.shape[0], e.shape[1]), e.device)
    with _timed("tc_edge_forward"):
        check(_lib.lib().hgnn_tc_edge_forward(C.byref(p), _ptr(x), _ptr(e), _ptr(plan_s.keys32), _ptr(plan_d.keys32), _ptr(perm),
                                              _ptr(rowptr), n_edges, x.shape[0], _ptr(out), _ptr(agg), _ptr(a0), _ptr(ws), ws.numel(),
                                              _stream()), "tc_edge_forward")
    _count(2 if agg is None else 4)  # bf16 node-row copy, edge kernel (+ aggregate fix-up and hub-segment kernels)
    TC_CALLS["count"] += 1
    return a0


class _TcEdgeStepAgg(torch.autograd.Function):
    """(e', agg) = (MLP([x[src] | x[dst] | e]) + e,  scatter_add(e', dst)) as ONE autograd node on the tensor-core
    path: the backward kernel consumes grad_e' and grad_agg[dst] together (no gather / add kernels in between) —
    the edge step of cell i and the aggregation that opens cell i+1 (gnn_utils.py:68-69 then :50)."""

    @staticmethod
    def forward(ctx, meta: MlpMeta, x, e, *params):
        _need_cuda(x, e, *params)
        segs = [_f32(x), _f32(x), _f32(e)]
        segs[1] = segs[0]
        ps = [_f32(t) for t in params]
        d, rows, layers = _build_desc(meta, segs, ps)
        out = torch.empty((rows, layers[-1][0].shape[0]), dtype=torch.float32, device=e.device)
        agg = torch.empty((segs[0].shape[0], out.shape[1]), dtype=torch.float32, device=e.device)
        need_bwd = any(ctx.needs_input_grad[1:])
        # one launch: edge MLP + skip + destination-sorted reduce (+ the bf16 input image the backward will stream)
        a0 = tc_edge_forward_raw(meta, segs, layers, out, agg, save_image=need_bwd)
        ctx.meta = meta
        _save_with_scratch(ctx, [segs[0], segs[2], *ps], a0)
        return out, agg

    @staticmethod
    def backward(ctx, g_out, g_agg):
        meta = ctx.meta
        saved, stash = _saved_and_scratch(ctx)
        x, e, ps = saved[0], saved[1], list(saved[2:])
        segs = [x, x, e]
        d, rows, layers = _build_desc(meta, segs, ps)
        g_out = torch.zeros_like(e) if g_out is None else _f32(g_out)
        g_agg = None if g_agg is None else _f32(g_agg)
        plan_d = meta.seg_plans[1]
        d_x, d_e, dW1, dW2, dv1, dv2 = tc_edge_backward_raw(meta, segs, layers, g_out, stash,
                                                            None if plan_d.is_identity() else plan_d.perm, g_agg, dst_sorted=True)
        gx = d_x if ctx.needs_input_grad[1] else None
        return (None, gx, d_e if ctx.needs_input_grad[2] else None, dW1, dv1[0], dv1[1], dv1[2], dW2, dv2[0], dv2[1], dv2[2])


def tc_edge_step_with_agg(meta: MlpMeta, x: Tensor, e: Tensor, params: Sequence[Tensor]):
    return _TcEdgeStepAgg.apply(meta, x, e, *params)


# ---------------------------------------------------------------------------
# skinny layers: fan-in <= 8 (encoder first layers) and fan-out <= 8 (heads' last layers), fp32, one warp per row
# ---------------------------------------------------------------------------
@functools.lru_cache(maxsize=256)
def narrow_in_supported(seg_widths, n_out: int) -> bool:
    return 1 <= len(seg_widths) <= MAX_SEGS and sum(seg_widths) <= 8 and n_out in (32, 64, 128, 256, 512)


@functools.lru_cache(maxsize=256)
def narrow_out_supported(fan_in: int, n_out: int) -> bool:
    return bool(_lib.lib().hgnn_narrow_out_supported(int(fan_in), int(n_out)))


class _NarrowIn(torch.autograd.Function):
    """act(LayerNorm(W . concat(gathered segments) + b)) for fan-in <= 8 (one-layer MlpMeta)."""

    @staticmethod
    def forward(ctx, meta: MlpMeta, n_seg: int, *tensors):
        _need_cuda(*tensors)
        segs = [_f32(t) for t in tensors[:n_seg]]
        params = [_f32(t) for t in tensors[n_seg:]]
        d, rows, layers = _build_desc(meta, segs, params)
        out = torch.empty((rows, layers[0][0].shape[0]), dtype=torch.float32, device=params[0].device)
        if rows:
            with _timed("narrow_in_forward"):
                check(_lib.lib().hgnn_narrow_in_forward(C.byref(d), rows, _ptr(out), _stream()), "narrow_in_forward")
            _count()
        ctx.meta, ctx.n_seg = meta, n_seg
        ctx.save_for_backward(*segs, *params)
        return out

    @staticmethod
    def backward(ctx, gout):
        meta, n_seg = ctx.meta, ctx.n_seg
        saved = ctx.saved_tensors
        segs, params = list(saved[:n_seg]), list(saved[n_seg:])
        d, rows, layers = _build_desc(meta, segs, params)
        W = layers[0][0]
        dev = W.device
        N, K = W.shape
        gout = _f32(gout)
        need = ctx.needs_input_grad[2:2 + n_seg]
        d_in = torch.empty((rows, K), dtype=torch.float32, device=dev) if any(need) else None
        dW = torch.empty_like(W)
        dvec = torch.empty((3, N), dtype=torch.float32, device=dev)
        L_ = _lib.lib()
        ws = _workspace(L_.hgnn_narrow_in_backward_workspace_bytes(N), dev)
        with _timed("narrow_in_backward"):
            check(L_.hgnn_narrow_in_backward(C.byref(d), rows, _ptr(gout), _ptr(d_in), _ptr(dW), _ptr(dvec), _ptr(ws), ws.numel(),
                                             _stream()), "narrow_in_backward")
        _count(2)
        grads: List[Optional[Tensor]] = [None, None]
        off = 0
        for s in range(n_seg):
            w = segs[s].shape[1]
            gs = None
            if need[s]:
                gs = d_in if n_seg == 1 else d_in[:, off:off + w]
                plan = meta.seg_plans[s]
                if plan is not None:
                    if plan.n_segments != segs[s].shape[0]:
                        raise _lib.HgnnError("narrow-in layer: gather plan does not cover the gathered tensor")
                    gs = segment_reduce_raw(gs, plan)  # column slice of d_in, reduced in place (row stride)
            grads.append(gs)
            off += w
        grads += [dW, dvec[0]]
        if meta.has_ln[0]:
            grads += [dvec[1], dvec[2]]
        return tuple(grads)


def narrow_in(meta: MlpMeta, segs: Sequence[Tensor], params: Sequence[Tensor]) -> Tensor:
    return _NarrowIn.apply(meta, len(segs), *segs, *params)


class _NarrowOut(torch.autograd.Function):
    """out = a W^T + b for fan-out <= 8."""

    @staticmethod
    def forward(ctx, a, W, b):
        _need_cuda(a, W, b)
        a, W, b = _f32(a), _f32(W), _f32(b)
        rows, K = a.shape
        out = torch.empty((rows, W.shape[0]), dtype=torch.float32, device=a.device)
        if rows:
            with _timed("narrow_out_forward"):
                check(_lib.lib().hgnn_narrow_out_forward(_ptr(a), rows, K, _ptr(W), _ptr(b), W.shape[0], _ptr(out), _stream()),
                      "narrow_out_forward")
            _count()
        ctx.save_for_backward(a, W)
        return out

    @staticmethod
    def backward(ctx, gout):
        a, W = ctx.saved_tensors
        gout = _f32(gout)
        rows, K = a.shape
        n_out = W.shape[0]
        d_a = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        dW = torch.empty_like(W)
        db = torch.empty(n_out, dtype=torch.float32, device=a.device)
        L_ = _lib.lib()
        ws = _workspace(L_.hgnn_narrow_out_backward_workspace_bytes(K, n_out), a.device)
        with _timed("narrow_out_backward"):
            check(L_.hgnn_narrow_out_backward(_ptr(a), rows, K, _ptr(W), n_out, _ptr(gout), _ptr(d_a), _ptr(dW), _ptr(db), _ptr(ws),
                                              ws.numel(), _stream()), "narrow_out_backward")
        _count(2)
        return d_a, dW, db


def narrow_out(a: Tensor, W: Tensor, b: Tensor) -> Tensor:
    return _NarrowOut.apply(a, W, b)


# ---------------------------------------------------------------------------
# tensor-core row layer (one Linear + LayerNorm + activation on a gathered concatenation)
# ---------------------------------------------------------------------------
class RowLayerMeta:
    """Static description of one tensor-core row-layer call: gather plans per segment, activation, LayerNorm eps,
    whether a residual tensor follows the segments, and the provider of the packed bf16 weight images."""

    def __init__(self, seg_plans, act: Optional[str], eps: float, has_skip: bool, pack):
        self.seg_plans = list(seg_plans)
        self.act = ACT_CODES[act]
        self.eps = float(eps)
        self.has_skip = bool(has_skip)
        self.pack = pack  # callable -> (W image, W^T image)


def _row_desc(meta: RowLayerMeta, segs, W, b, g, be, w_packed, skip=None):
    d = _lib.TcRowLayer()
    d.n_seg, d.n_out, d.act, d.ln_eps = len(segs), W.shape[0], meta.act, meta.eps
    rows = None
    for s, t in enumerate(segs):
        plan = meta.seg_plans[s]
        d.seg_ptr[s] = t.data_ptr() if t is not None else None
        d.seg_width[s] = t.shape[1] if t is not None else 0
        d.seg_idx[s] = plan.keys32.data_ptr() if plan is not None else None
        n = plan.n_items if plan is not None else (t.shape[0] if t is not None else None)
        if rows is None:
            rows = n
        elif n is not None and rows != n:
            raise _lib.HgnnError(f"row layer: segment {s} yields {n} rows, expected {rows}")
    d.w_packed = _ptr(w_packed)
    d.bias, d.gamma, d.beta = b.data_ptr(), g.data_ptr(), be.data_ptr()
    d.skip = _ptr(skip)
    return d, rows


def tc_row_supported(seg_widths: Sequence[int], n_out: int, act: Optional[str]) -> bool:
    """Whether hgnn_tc_row_forward/backward are built for this layer shape (include/hgnn_b200.h)."""
    return _tc_row_supported(tuple(int(w) for w in seg_widths), int(n_out), act)


@functools.lru_cache(maxsize=256)
def _tc_row_supported(seg_widths, n_out, act) -> bool:
    if len(seg_widths) < 1 or len(seg_widths) > MAX_SEGS or act not in ACT_CODES:
        return False
    d = _lib.TcRowLayer()
    d.n_seg, d.n_out, d.act = len(seg_widths), int(n_out), ACT_CODES[act]
    for s, w in enumerate(seg_widths):
        d.seg_width[s] = int(w)
    return bool(_lib.lib().hgnn_tc_row_supported(C.byref(d)))


class _TcRowLayer(torch.autograd.Function):
    """out = act(LayerNorm(W . concat(gathered segments) + b)) (+ skip) on tcgen05 tensor cores. The forward leaves the
    bf16 image of its gathered input in HBM; the backward recomputes the layer from that image (no fp32 inputs or
    activations are kept alive by this node) and emits d_in, dW, d bias / gamma / beta."""

    @staticmethod
    def forward(ctx, meta: RowLayerMeta, n_seg: int, *tensors):
        _need_cuda(*tensors)
        segs = [_f32(t) for t in tensors[:n_seg]]
        skip = _f32(tensors[n_seg]) if meta.has_skip else None
        W, b, g, be = [_f32(t) for t in tensors[n_seg + int(meta.has_skip):]]
        w_packed, _ = meta.pack()
        d, rows = _row_desc(meta, segs, W, b, g, be, w_packed, skip)
        K = sum(t.shape[1] for t in segs)
        out = torch.empty((rows, W.shape[0]), dtype=torch.float32, device=W.device)
        if skip is not None and tuple(skip.shape) != tuple(out.shape):
            raise _lib.HgnnError(f"row layer: residual has shape {tuple(skip.shape)}, output {tuple(out.shape)}")
        need_bwd = any(ctx.needs_input_grad[2:])
        a_img = None
        if need_bwd and rows:
            # = hgnn_tc_row_image_bytes(rows, K): one 16 KB block per 128-row tile and 64 input columns
            a_img = torch.empty(((rows + 127) // 128) * (K // 64) * 16384, dtype=torch.uint8, device=W.device)
        if rows:
            if PROFILE is None:
                check(_lib.lib().hgnn_tc_row_forward(C.byref(d), rows, _ptr(out), _ptr(a_img), _stream()), "tc_row_forward")
            else:
                with _timed("tc_row_forward"):
                    check(_lib.lib().hgnn_tc_row_forward(C.byref(d), rows, _ptr(out), _ptr(a_img), _stream()), "tc_row_forward")
            _count()
            TC_ROW_CALLS["count"] += 1
        ctx.meta, ctx.n_seg, ctx.rows = meta, n_seg, rows
        ctx.widths = [t.shape[1] for t in segs]
        ctx.seg_rows = [t.shape[0] for t in segs]
        ctx.desc = d  # the backward reuses the descriptor (weights / LayerNorm pointers: saved tensors, same storage)
        _save_with_scratch(ctx, [W, b, g, be], a_img)
        return out

    @staticmethod
    def backward(ctx, gout):
        meta, n_seg, rows = ctx.meta, ctx.n_seg, ctx.rows
        (W, b, g, be), a_img = _saved_and_scratch(ctx)
        gout = _f32(gout)
        dev = W.device
        N, K = W.shape
        w_packed, wt_packed = meta.pack()
        d = ctx.desc
        d.w_packed = _ptr(w_packed)
        widths = ctx.widths
        need = ctx.needs_input_grad[2:]
        # one dense gradient matrix per segment (what the consumers upstream want: no strided views of a [rows, K] matrix,
        # nothing stored for segments without a gradient) whenever the kernel's 128-column pieces line up with the segments
        split = n_seg > 1 and all(w % 128 == 0 for w in widths)
        dW = torch.empty_like(W)
        dvec = torch.empty((3, N), dtype=torch.float32, device=dev)
        L_ = _lib.lib()
        d_in = None
        d_segs: List[Optional[Tensor]] = [None] * n_seg
        if split:
            for s in range(n_seg):
                if need[s]:
                    d_segs[s] = torch.empty((rows, widths[s]), dtype=torch.float32, device=dev)
        else:
            d_in = torch.empty((rows, K), dtype=torch.float32, device=dev)
        if rows:
            ws = _workspace(L_.hgnn_tc_row_backward_workspace_bytes(rows, K, N), dev)
            with _timed("tc_row_backward"):
                if split:
                    ptrs = (C.c_void_p * MAX_SEGS)(*[_ptr(t) for t in d_segs])
                    check(L_.hgnn_tc_row_backward_split(C.byref(d), _ptr(wt_packed), _ptr(a_img), rows, _ptr(gout), ptrs, _ptr(dW),
                                                        _ptr(dvec), _ptr(ws), ws.numel(), _stream()), "tc_row_backward_split")
                else:
                    check(L_.hgnn_tc_row_backward(C.byref(d), _ptr(wt_packed), _ptr(a_img), rows, _ptr(gout), _ptr(d_in), _ptr(dW),
                                                  _ptr(dvec), _ptr(ws), ws.numel(), _stream()), "tc_row_backward")
            _count(3)  # data-gradient kernel, weight-gradient GEMM, ordered reduce of its partials and of the column sums
            TC_ROW_CALLS["count"] += 1
        else:
            dW.zero_()
            dvec.zero_()
        grads: List[Optional[Tensor]] = [None, None]
        off = 0
        for s in range(n_seg):
            w = widths[s]
            gs = None
            if need[s]:
                gs = d_segs[s] if split else (d_in if n_seg == 1 else d_in[:, off:off + w])
                plan = meta.seg_plans[s]
                if plan is not None:
                    if plan.n_segments != ctx.seg_rows[s]:
                        raise _lib.HgnnError("row layer: gather plan does not cover the gathered tensor")
                    gs = segment_reduce_raw(gs, plan)  # (a column slice of d_in is reduced in place: row stride)
            grads.append(gs)
            off += w
        if meta.has_skip:
            grads.append(gout if need[n_seg] else None)  # d(out)/d(skip) = identity
        grads += [dW, dvec[0], dvec[1], dvec[2]]
        return tuple(grads)


def tc_row_layer(meta: RowLayerMeta, segs: Sequence[Tensor], skip: Optional[Tensor], W, b, gamma, beta) -> Tensor:
    extra = [skip] if meta.has_skip else []
    return _TcRowLayer.apply(meta, len(segs), *segs, *extra, W, b, gamma, beta)


# ---------------------------------------------------------------------------
# generic tensor-core layer: plain tcgen05 GEMM(s) + row-wise LayerNorm/activation kernels (latent 64 / 256 shapes)
# ---------------------------------------------------------------------------
def _width_chunks(n: int):
    """[(offset, width)] with widths in {256, 128, 64} covering n (a multiple of 64)."""
    out, off = [], 0
    while off < n:
        rest = n - off
        w = 256 if rest >= 256 else (128 if rest >= 128 else 64)
        out.append((off, w))
        off += w
    return out


@functools.lru_cache(maxsize=256)
def tc_split_supported(seg_widths, n_out: int, act: Optional[str]) -> bool:
    k = sum(seg_widths)
    return (1 <= len(seg_widths) <= MAX_SEGS and all(w > 0 and w % 64 == 0 for w in seg_widths) and k <= 768
            and n_out in (64, 128, 256, 512) and act in ACT_CODES and (n_out % 128 == 0 or k % 128 == 0))


def tc_pack_split(W: Tensor):
    """Weight images of one Linear for the generic layer: W row chunks (forward GEMMs, one per <= 256 output columns) and
    W^T row chunks (data-gradient GEMMs, one per <= 256 input columns): ([(col0, n, image)], [(col0, n, image)])."""
    Wd = W.detach()
    fwd = [(c0, n, tc_pack_weight(Wd[c0:c0 + n].contiguous())) for c0, n in _width_chunks(Wd.shape[0])]
    Wt = Wd.t().contiguous()
    bwd = [(c0, n, tc_pack_weight(Wt[c0:c0 + n].contiguous())) for c0, n in _width_chunks(Wt.shape[0])]
    return fwd, bwd


def _gemm_desc(seg_tensors, seg_plans, n_out, w_packed, bias_ptr):
    d = _lib.TcRowLayer()
    d.n_seg, d.n_out, d.act, d.ln_eps = len(seg_tensors), n_out, 0, 0.0
    for s, t in enumerate(seg_tensors):
        d.seg_ptr[s] = t.data_ptr()
        d.seg_width[s] = t.shape[1]
        d.seg_idx[s] = seg_plans[s].keys32.data_ptr() if seg_plans[s] is not None else None
    d.w_packed = _ptr(w_packed)
    d.bias = bias_ptr
    return d


class _TcSplitLayer(torch.autograd.Function):
    """out = act(LayerNorm(W . concat(gathered segments) + b)) (+ skip) as tcgen05 GEMM(s) + a row-wise LayerNorm kernel:
    the layer shapes of latent 64 / 256 (fan-out 64 or 512, fan-in up to 768) that the fused kernels do not cover.
    Keeps the pre-activation h (fp32) and the bf16 image of the gathered input for the backward."""

    @staticmethod
    def forward(ctx, meta: RowLayerMeta, n_seg: int, *tensors):
        _need_cuda(*tensors)
        segs = [_f32(t) for t in tensors[:n_seg]]
        skip = _f32(tensors[n_seg]) if meta.has_skip else None
        W, b, g, be = [_f32(t) for t in tensors[n_seg + int(meta.has_skip):]]
        N, K = W.shape
        rows = meta.seg_plans[0].n_items if meta.seg_plans[0] is not None else segs[0].shape[0]
        dev = W.device
        fwd_chunks, _ = meta.pack()
        L_ = _lib.lib()
        h = torch.empty((rows, N), dtype=torch.float32, device=dev)
        out = torch.empty((rows, N), dtype=torch.float32, device=dev)
        if skip is not None and tuple(skip.shape) != (rows, N):
            raise _lib.HgnnError(f"generic layer: residual has shape {tuple(skip.shape)}, output {(rows, N)}")
        need_bwd = any(ctx.needs_input_grad[2:])
        a_img = torch.empty(L_.hgnn_tc_row_image_bytes(rows, K), dtype=torch.uint8, device=dev) if (need_bwd and rows) else None
        if rows:
            with _timed("tc_split_forward"):
                for i, (c0, n, img) in enumerate(fwd_chunks):
                    d = _gemm_desc(segs, meta.seg_plans, n, img, b.data_ptr() + 4 * c0)
                    check(L_.hgnn_tc_gemm(C.byref(d), rows, _ptr(h), N, c0, _ptr(a_img) if i == 0 else None, _stream()), "tc_gemm")
                check(L_.hgnn_ln_act_forward(_ptr(h), rows, N, _ptr(g), _ptr(be), meta.eps, meta.act, _ptr(skip), _ptr(out), _stream()),
                      "ln_act_forward")
            _count(len(fwd_chunks) + 1)
            TC_ROW_CALLS["count"] += 1
        ctx.meta, ctx.n_seg, ctx.rows = meta, n_seg, rows
        ctx.widths = [t.shape[1] for t in segs]
        ctx.seg_rows = [t.shape[0] for t in segs]
        _save_with_scratch(ctx, [W, g, be, h], a_img)
        return out

    @staticmethod
    def backward(ctx, gout):
        meta, n_seg, rows = ctx.meta, ctx.n_seg, ctx.rows
        (W, g, be, h), a_img = _saved_and_scratch(ctx)
        gout = _f32(gout)
        dev = W.device
        N, K = W.shape
        L_ = _lib.lib()
        dW = torch.empty_like(W)
        dvec = torch.empty((3, N), dtype=torch.float32, device=dev)
        need = ctx.needs_input_grad[2:]
        widths = ctx.widths
        offs = [sum(widths[:s]) for s in range(n_seg)]
        _, bwd_chunks = meta.pack()

        def seg_of(c0, n):  # the segment that holds input columns [c0, c0 + n) entirely, or None
            for s in range(n_seg):
                if offs[s] <= c0 and c0 + n <= offs[s] + widths[s]:
                    return s
            return None
        # one dense gradient matrix per segment when every W^T chunk lies inside one segment (latent 256: 256-wide segments and
        # chunks): no strided slices for the consumers to copy, no GEMM for a segment that needs no gradient
        split = n_seg > 1 and all(seg_of(c0, n) is not None for c0, n, _ in bwd_chunks)
        d_in = None
        d_segs: List[Optional[Tensor]] = [None] * n_seg
        if split:
            for s in range(n_seg):
                if need[s]:
                    d_segs[s] = torch.empty((rows, widths[s]), dtype=torch.float32, device=dev)
        else:
            d_in = torch.empty((rows, K), dtype=torch.float32, device=dev)
        if rows:
            delta = torch.empty((rows, N), dtype=torch.float32, device=dev)
            d_img = torch.empty(L_.hgnn_tc_row_image_bytes(rows, N), dtype=torch.uint8, device=dev)
            ws = _workspace(max(L_.hgnn_ln_act_backward_workspace_bytes(N), L_.hgnn_tc_wgrad_workspace_bytes(rows, N, K)), dev)
            with _timed("tc_split_backward"):
                check(L_.hgnn_ln_act_backward(_ptr(h), _ptr(gout), rows, N, _ptr(g), _ptr(be), meta.eps, meta.act, _ptr(delta),
                                              _ptr(dvec), _ptr(ws), ws.numel(), _stream()), "ln_act_backward")
                image_done = False  # the first GEMM that runs also leaves the bf16 image of delta (weight-gradient operand)
                for c0, n, img in bwd_chunks:  # d_in[:, c0 : c0 + n] = delta . W[:, c0 : c0 + n]
                    out, ld, col = d_in, K, c0
                    if split:
                        sg = seg_of(c0, n)
                        if d_segs[sg] is None:
                            continue
                        out, ld, col = d_segs[sg], widths[sg], c0 - offs[sg]
                    d = _gemm_desc([delta], [None], n, img, None)
                    check(L_.hgnn_tc_gemm(C.byref(d), rows, _ptr(out), ld, col, None if image_done else _ptr(d_img), _stream()), "tc_gemm")
                    image_done = True
                if not image_done:  # no input gradient wanted at all: one GEMM into scratch for the image
                    c0, n, img = bwd_chunks[0]
                    d = _gemm_desc([delta], [None], n, img, None)
                    scratch = torch.empty((rows, n), dtype=torch.float32, device=dev)
                    check(L_.hgnn_tc_gemm(C.byref(d), rows, _ptr(scratch), n, 0, _ptr(d_img), _stream()), "tc_gemm")
                check(L_.hgnn_tc_wgrad(_ptr(d_img), N, _ptr(a_img), K, rows, _ptr(dW), _ptr(ws), ws.numel(), _stream()), "tc_wgrad")
            _count(2 + len(bwd_chunks) + 2 * ((N // 128 or 1) * (K // 128 or 1) // 4 + 1))
            TC_ROW_CALLS["count"] += 1
        else:
            dW.zero_()
            dvec.zero_()
        grads: List[Optional[Tensor]] = [None, None]
        off = 0
        for s in range(n_seg):
            w = widths[s]
            gs = None
            if need[s]:
                gs = d_segs[s] if split else (d_in if n_seg == 1 else d_in[:, off:off + w])
                plan = meta.seg_plans[s]
                if plan is not None:
                    if plan.n_segments != ctx.seg_rows[s]:
                        raise _lib.HgnnError("generic layer: gather plan does not cover the gathered tensor")
                    gs = segment_reduce_raw(gs, plan)  # (a column slice of d_in is reduced in place: row stride)
            grads.append(gs)
            off += w
        if meta.has_skip:
            grads.append(gout if need[n_seg] else None)
        grads += [dW, dvec[0], dvec[1], dvec[2]]
        return tuple(grads)


def tc_split_layer(meta: RowLayerMeta, segs: Sequence[Tensor], skip: Optional[Tensor], W, b, gamma, beta) -> Tensor:
    extra = [skip] if meta.has_skip else []
    return _TcSplitLayer.apply(meta, len(segs), *segs, *extra, W, b, gamma, beta)


def fused_mlp(meta: MlpMeta, segs: Sequence[Tensor], params: Sequence[Tensor]) -> Tensor:
    """Row-wise MLP on the concatenation of (optionally gathered) segments, with
    LayerNorm/activation per layer and an optional skip connection — one kernel
    forward, recompute-in-backward."""
    return _FusedMLP.apply(meta, len(segs), *segs, *params)


# ---------------------------------------------------------------------------
# graph construction
# ---------------------------------------------------------------------------
def knn_radius(query: Tensor, ref: Tensor, k: int, radius: float, query_ptr: Optional[Tensor] = None,
               ref_ptr: Optional[Tensor] = None) -> Tensor:
    """[n_query, k] int64 neighbour ids, ascending distance, d < radius, -1 padded
    (frnn.frnn_grid_points as used by find_neighbors, utils.py:228-239).
    ``query_ptr`` / ``ref_ptr`` (int32 CUDA tensors of n_events + 1 row offsets): a batch of events searched in one launch,
    every query against the references of its own event only; ids are rows of ``ref``."""
    _need_cuda(query, ref, query_ptr, ref_ptr)
    q, r = _f32(query.detach()), _f32(ref.detach())
    idx = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
    if (query_ptr is None) != (ref_ptr is None):
        raise _lib.HgnnError("knn_radius: query_ptr and ref_ptr come together")
    if q.shape[0] and k:
        L = _lib.lib()
        nb = L.hgnn_knn_radius_workspace_bytes(q.shape[0], r.shape[0], k)  # > 0: small problem, split over reference ranges
        ws = _workspace(nb, q.device) if nb else None
        if query_ptr is None:
            check(L.hgnn_knn_radius_ws(_ptr(q), q.shape[0], _ptr(r), r.shape[0], q.shape[1], k, float(radius), _ptr(idx),
                                       _ptr(ws), ws.numel() if ws is not None else 0, _stream()), "knn_radius")
        else:
            if (query_ptr.dtype != torch.int32 or ref_ptr.dtype != torch.int32 or query_ptr.dim() != 1
                    or query_ptr.shape != ref_ptr.shape or query_ptr.numel() < 2):
                raise _lib.HgnnError("knn_radius: query_ptr / ref_ptr must be int32 vectors of n_events + 1 offsets")
            check(L.hgnn_knn_radius_batched(_ptr(q), q.shape[0], _ptr(r), r.shape[0], q.shape[1], k, float(radius),
                                            _ptr(query_ptr.contiguous()), _ptr(ref_ptr.contiguous()), query_ptr.numel() - 1,
                                            _ptr(idx), _ptr(ws), ws.numel() if ws is not None else 0, _stream()), "knn_radius_batched")
        _count(2 if nb else 1)
    return idx


def knn_edges(idx: Tensor) -> Tensor:
    """Compacts a -1 padded [n_query, k] neighbour table into graph[2, E'] in
    query-major, rank-minor order (gnn_utils.py:195-202). One host sync (E')."""
    _need_cuda(idx)
    nq, k = idx.shape
    dev = idx.device
    graph = torch.empty((2, max(nq * k, 1)), dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    L = _lib.lib()
    ws = _workspace(L.hgnn_knn_edges_workspace_bytes(nq), dev)
    check(L.hgnn_knn_edges(_ptr(idx), nq, k, _ptr(graph), _ptr(count), _ptr(ws), ws.numel(), _stream()), "knn_edges")
    _count(3)
    n = int(count.item())
    return graph[:, :n]


def symmetrize(graph: Tensor, n_vertices: int) -> Tensor:
    """Union with the transpose, de-duplicated, lexicographic column order
    (cugraph symmetrize, gnn_utils.py:198-199). One host sync (edge count)."""
    _need_cuda(graph)
    E = graph.shape[1]
    dev = graph.device
    if E == 0:
        return graph.new_zeros((2, 0))
    g = graph.contiguous()
    out = torch.empty((2, 2 * E), dtype=torch.int64, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    L = _lib.lib()
    ws = _workspace(L.hgnn_symmetrize_workspace_bytes(E), dev)
    check(L.hgnn_symmetrize(_ptr(g), g.stride(0), E, int(n_vertices), _ptr(out), _ptr(count), _ptr(ws), ws.numel(),
                            _stream()), "symmetrize")
    _count(4)
    n = int(count.item())
    return out[:, :n]


def edge_max_dist(a: Tensor, b: Tensor, graph: Tensor) -> Tensor:
    """max_e ||a[graph[0,e]] - b[graph[1,e]]||_2 as a 0-d device tensor (gnn_utils.py:204)."""
    _need_cuda(a, b, graph)
    a, b = _f32(a.detach()), _f32(b.detach())
    out = torch.zeros(1, dtype=torch.float32, device=a.device)
    E = graph.shape[1]
    if E:
        g0, g1 = graph[0].contiguous(), graph[1].contiguous()
        check(_lib.lib().hgnn_edge_max_dist(_ptr(a), _ptr(g0), _ptr(b), _ptr(g1), a.shape[1], E, _ptr(out), _stream()),
              "edge_max_dist")
        _count()
    return out[0]


def connected_components(graph: Tensor, n_vertices: int, keep: Optional[Tensor] = None) -> Tensor:
    """labels[v] = min vertex id of v's component over kept edges; -1 if v touches
    no kept edge (cugraph connected_components as consumed at BC/Models/HGNN_GMM.py:215-232)."""
    _need_cuda(graph, keep)
    dev = graph.device
    g = graph.contiguous()
    labels = torch.empty(n_vertices, dtype=torch.int32, device=dev)
    k8 = None
    if keep is not None:
        k8 = keep.to(torch.uint8).contiguous()
    L = _lib.lib()
    ws = _workspace(L.hgnn_connected_components_workspace_bytes(n_vertices), dev)
    check(L.hgnn_connected_components(_ptr(g), g.stride(0) if g.shape[1] else 0, g.shape[1], _ptr(k8), n_vertices,
                                      _ptr(labels), _ptr(ws), ws.numel(), _stream()), "connected_components")
    _count(3)
    return labels


def gmm1d_fit(x: Tensor, max_iter: int = 100, tol: float = 1e-3) -> Tensor:
    """Two-component 1-D Gaussian mixture by EM on device; returns a device tensor
    (pi0, mu0, var0, pi1, mu1, var1). Replaces sklearn GaussianMixture.fit
    (BC/Models/HGNN_GMM.py:192)."""
    _need_cuda(x)
    x = _f32(x.detach().reshape(-1))
    params = torch.empty(6, dtype=torch.float32, device=x.device)
    L = _lib.lib()
    ws = _workspace(L.hgnn_gmm1d_workspace_bytes(), x.device)
    check(L.hgnn_gmm1d_fit(_ptr(x), x.numel(), int(max_iter), float(tol), _ptr(params), _ptr(ws), ws.numel(), _stream()),
          "gmm1d_fit")
    _count(2 * (max_iter + 1))
    return params
