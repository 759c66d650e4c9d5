"""Host-side mirror of the hot-path pieces of the reference's Modules/utils.py:
``make_mlp`` (utils.py:169-196) and ``find_neighbors`` (utils.py:228-239).

``make_mlp`` returns an ``nn.Sequential`` subclass whose children are the very
same ``nn.Linear`` / ``nn.LayerNorm`` / activation modules at the very same
indices, so ``state_dict()`` keys and order are identical to the reference and
reference checkpoints load with ``strict=True``. Execution, however, never
walks the children: the whole stack runs as one fused CUDA kernel
(``ops.fused_mlp``), with the concatenation / gathers of its input folded in.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import ops

_SUPPORTED_ACTS = ("GELU", "Tanh", "ReLU", "SiLU", "Sigmoid")


class FusedMLP(nn.Sequential):
    """nn.Sequential-compatible container executed as a single fused kernel."""

    def _layers(self):
        """[(linear, layernorm|None, act_name|None)] parsed from the children."""
        cached = getattr(self, "_layer_cache", None)
        if cached is not None:
            return cached
        out, cur = [], None
        for m in self:
            if isinstance(m, nn.Linear):
                if cur is not None:
                    out.append(tuple(cur))
                cur = [m, None, None]
            elif isinstance(m, nn.LayerNorm):
                cur[1] = m
            else:
                name = type(m).__name__
                if name not in _SUPPORTED_ACTS:
                    raise NotImplementedError(f"activation {name} has no CUDA implementation in hgnn_b200")
                cur[2] = name
        out.append(tuple(cur))
        object.__setattr__(self, "_layer_cache", out)
        return out

    def _drop_packed(self):
        for k in ("_tc_cache", "_row_cache", "_split_cache"):
            self.__dict__.pop(k, None)

    def invalidate_packed_weights(self):
        """Forget the packed bf16 weight images (they are keyed on (data_ptr, version counter) of the fp32 weights; a write
        through ``weight.data`` does not move the counter — call this after such a write)."""
        self._drop_packed()

    def _params(self):
        ps, acts, lns = [], [], []
        eps = 1e-5
        for lin, ln, act in self._layers():
            ps += [lin.weight, lin.bias]
            if ln is not None:
                ps += [ln.weight, ln.bias]
                eps = ln.eps
            acts.append(act)
            lns.append(ln is not None)
        return ps, acts, lns, eps

    def fused(self, segs: Sequence[torch.Tensor], plans: Sequence[Optional[ops.SegmentPlan]] = None, skip: int = -1):
        """MLP(concat_s segs[s][plans[s].keys]) (+ segs[skip] rows). ``plans[s]`` is the
        SegmentPlan over the gather index of segment s (None = rows used as they are)."""
        if plans is None:
            plans = [None] * len(segs)
        ps, acts, lns, eps = self._params()
        packer = self._tc_packer(segs, plans, skip, lns)
        if packer is not None and self._layers()[1][0].out_features != 128 and torch.is_grad_enabled() and (
                any(t.requires_grad for t in segs) or any(p.requires_grad for p in ps)):
            packer = None  # the fused edge kernel has a tensor-core backward at latent 128 only: train layer by layer
        if packer is None:
            # the split into kernel groups depends on shapes / gather pattern / precision only: decided once per pattern
            key = (ops.get_precision(), tuple(t.shape[1] for t in segs), tuple(p is None for p in plans), skip,
                   all(t.is_cuda for t in segs))
            cache = self.__dict__.setdefault("_groups_cache", {})
            if key not in cache:
                cache[key] = self._row_groups(segs, plans, skip)
            groups = cache[key]
            if groups is not None:
                return self._run_groups(groups, list(segs), list(plans), skip)
            packer = self._tc_packer(segs, plans, skip, lns)
        meta = ops.MlpMeta(plans, acts, lns, skip, eps, tc_pack=packer)
        return ops.fused_mlp(meta, list(segs), ps)

    # ---- layer-wise execution: tensor-core row layers where a kernel exists, fp32 SIMT groups elsewhere ----
    def _row_groups(self, segs, plans, skip):
        """[("tc", l) | ("tcs", l) | ("nin", l) | ("nout", l) | ("simt", l0, l1)] covering the layers in order, or None when no layer
        has a specialised kernel (precision fp32, CPU tensors, odd shapes): the whole stack then runs as one fp32 kernel.
        tc = tensor-core row layer, nin / nout = skinny fan-in / fan-out layers, simt = generic fused fp32 group."""
        if ops.get_precision() == "fp32" or not all(t.is_cuda for t in segs):
            return None
        if skip >= 0 and plans[skip] is not None:
            return None  # a gathered residual only exists in the fused fp32 kernel
        layers = self._layers()
        last = len(layers) - 1
        groups, special, cur = [], False, None
        widths = tuple(t.shape[1] for t in segs)
        for l, (lin, ln, act) in enumerate(layers):
            w_in = widths if l == 0 else (layers[l - 1][0].out_features,)
            has_res = l == last and skip >= 0
            kind = None
            if ln is not None and ops.tc_row_supported(w_in, lin.out_features, act):
                kind = "tc"
            elif ln is not None and ops.tc_split_supported(w_in, lin.out_features, act):
                kind = "tcs"  # plain tcgen05 GEMM(s) + row-wise LayerNorm kernel (latent 64 / 256 shapes)
            elif l == 0 and not has_res and ops.narrow_in_supported(w_in, lin.out_features):
                kind = "nin"
            elif (not has_res and ln is None and act is None and len(w_in) == 1 and (l > 0 or plans[0] is None)
                  and ops.narrow_out_supported(w_in[0], lin.out_features)):
                kind = "nout"
            if kind is not None:
                if cur is not None:
                    groups.append(("simt", cur, l))
                    cur = None
                groups.append((kind, l))
                special = True
            elif cur is None:
                cur = l
        if cur is not None:
            groups.append(("simt", cur, len(layers)))
        return groups if special else None

    def _split_pack(self, l):
        lin = self._layers()[l][0]

        def pack():
            w = lin.weight
            key = (w.data_ptr(), w._version)
            cache = self.__dict__.setdefault("_split_cache", {})
            hit = cache.get(l)
            if hit is None or hit[0] != key:
                hit = (key,) + ops.tc_pack_split(w)
                cache[l] = hit
            return hit[1], hit[2]
        return pack

    def _row_pack(self, l):
        lin = self._layers()[l][0]

        def pack():
            w = lin.weight
            key = (w.data_ptr(), w._version)
            cache = self.__dict__.setdefault("_row_cache", {})
            hit = cache.get(l)
            if hit is None or hit[0] != key:  # bf16 shadow images, rebuilt lazily when the fp32 parameter changes
                hit = (key, ops.tc_pack_weight(w), ops.tc_pack_weight_t(w))
                cache[l] = hit
            return hit[1], hit[2]
        return pack

    def _run_groups(self, groups, segs, plans, skip):
        layers = self._layers()
        last = len(layers) - 1
        cur_segs, cur_plans = segs, plans
        residual = segs[skip] if skip >= 0 else None
        for grp in groups:
            if grp[0] == "tc":
                l = grp[1]
                lin, ln, act = layers[l]
                res = residual if l == last else None
                meta = ops.RowLayerMeta(cur_plans, act, ln.eps, res is not None, self._row_pack(l))
                y = ops.tc_row_layer(meta, cur_segs, res, lin.weight, lin.bias, ln.weight, ln.bias)
            elif grp[0] == "tcs":
                l = grp[1]
                lin, ln, act = layers[l]
                res = residual if l == last else None
                meta = ops.RowLayerMeta(cur_plans, act, ln.eps, res is not None, self._split_pack(l))
                y = ops.tc_split_layer(meta, cur_segs, res, lin.weight, lin.bias, ln.weight, ln.bias)
            elif grp[0] == "nin":
                lin, ln, act = layers[grp[1]]
                ps = [lin.weight, lin.bias] + ([ln.weight, ln.bias] if ln is not None else [])
                meta = ops.MlpMeta(cur_plans, [act], [ln is not None], -1, ln.eps if ln is not None else 1e-5)
                y = ops.narrow_in(meta, cur_segs, ps)
            elif grp[0] == "nout":
                lin = layers[grp[1]][0]
                y = ops.narrow_out(cur_segs[0], lin.weight, lin.bias)
            else:
                l0, l1 = grp[1], grp[2]
                ps, acts, lns, eps = [], [], [], 1e-5
                for lin, ln, act in layers[l0:l1]:
                    ps += [lin.weight, lin.bias]
                    if ln is not None:
                        ps += [ln.weight, ln.bias]
                        eps = ln.eps
                    acts.append(act)
                    lns.append(ln is not None)
                sk = -1
                call_segs, call_plans = list(cur_segs), list(cur_plans)
                if l1 - 1 == last and residual is not None:
                    if l0 == 0:
                        sk = skip
                    else:  # the residual is not an input of this trailing fp32 group: added after it
                        sk = -2
                meta = ops.MlpMeta(call_plans, acts, lns, sk if sk >= 0 else -1, eps)
                y = ops.fused_mlp(meta, call_segs, ps)
                if sk == -2:
                    y = y + residual
            cur_segs, cur_plans = [y], [None]
        return y

    def edge_step(self, nodes, edges, plan_src, plan_dst):
        """e' = MLP([x[src] | x[dst] | e]) + e. Returns (e', agg) where agg = scatter_add(e', dst) comes out of the
        same autograd node when the tensor-core forward+backward kernels cover this network, else (e', None)."""
        segs, plans = [nodes, nodes, edges], [plan_src, plan_dst, None]
        ps, acts, lns, eps = self._params()
        packer = self._tc_packer(segs, plans, 2, lns)
        if packer is not None and self._layers()[1][0].out_features == 128:
            meta = ops.MlpMeta(plans, acts, lns, 2, eps, tc_pack=packer)
            return ops.tc_edge_step_with_agg(meta, nodes, edges, ps)
        return self.fused(segs, plans, skip=2), None

    def _tc_packer(self, segs, plans, skip, lns):
        """Returns a weight-image provider when this call is an edge step the tcgen05 kernel covers:
        segments (x | by_src, x | by_dst, e), skip = e, two LayerNorm layers, latent in {64, 128}."""
        if ops.get_precision() == "fp32" or len(segs) != 3 or skip != 2:
            return None
        layers = self._layers()
        if len(layers) != 2 or not all(lns) or plans[0] is None or plans[1] is None or plans[2] is not None:
            return None
        if segs[0] is not segs[1] or not segs[0].is_cuda:
            return None
        L, H = layers[1][0].out_features, layers[0][0].out_features
        if segs[0].shape[1] != L or segs[2].shape[1] != L or not ops.tc_supported(L, H, 2, True, layers[0][2], layers[1][2]):
            return None

        def pack():
            w1, w2 = layers[0][0].weight, layers[1][0].weight
            key = (w1.data_ptr(), w1._version, w2.data_ptr(), w2._version)
            cached = getattr(self, "_tc_cache", None)
            if cached is None or cached[0] != key:
                # bf16 shadow copies, rebuilt lazily when the fp32 parameters change; never registered
                w1d = w1.detach()
                # [W1[:, 0:L]^T | W1[:, L:2L]^T] as an [L, 2H] Linear weight: the node-level d(x) GEMM of the backward
                wx = torch.cat([w1d[:, :L].t(), w1d[:, L:2 * L].t()], dim=1).contiguous()
                cached = (key, ops.tc_pack_weight(w1), ops.tc_pack_weight(w2),
                          ops.tc_pack_weight_t(w1), ops.tc_pack_weight_t(w2), ops.tc_pack_weight(wx))
                object.__setattr__(self, "_tc_cache", cached)
            return cached[1:]
        return pack

    def forward(self, x):
        lead = x.shape[:-1]
        y = self.fused([x.reshape(-1, x.shape[-1])])
        return y.reshape(*lead, y.shape[-1])


def make_mlp(input_size, hidden_size, output_size, hidden_layers, hidden_activation="GELU",
             output_activation="GELU", layer_norm=False):
    """Same signature, same Sequential index layout as the reference factory:
    ``[Linear, (LayerNorm), act] x (hidden_layers-1)``, then ``Linear`` and, only
    when ``output_activation`` is given, ``(LayerNorm), act``."""
    hidden_cls = getattr(nn, hidden_activation)
    out_cls = getattr(nn, output_activation) if output_activation is not None else None
    widths = [input_size] + [hidden_size] * (hidden_layers - 1) + [output_size]
    mods = []
    n = len(widths) - 1
    for i in range(n):
        mods.append(nn.Linear(widths[i], widths[i + 1]))
        final = i == n - 1
        act_cls = out_cls if final else hidden_cls
        if act_cls is None:
            continue
        if layer_norm:
            mods.append(nn.LayerNorm(widths[i + 1]))
        mods.append(act_cls())
    return FusedMLP(*mods)


def find_neighbors(embedding1, embedding2, r_max=1.0, k_max=10, ptr1=None, ptr2=None):
    """[P1, k_max] int64 neighbour table, -1 padded (brute-force tiled CUDA kernel
    in place of frnn.frnn_grid_points). ``ptr1`` / ``ptr2``: int32 row offsets of the events of a batch — rows of
    ``embedding1`` then only find neighbours of their own event (ops.knn_radius)."""
    r = float(r_max.reshape(-1)[0]) if torch.is_tensor(r_max) else float(r_max)
    return ops.knn_radius(embedding1, embedding2, int(k_max), r, ptr1, ptr2)


def event_offsets(event_of_row, n_events):
    """int32 [n_events + 1] row offsets of the events of a batch from the ascending event id of every row (the ``batch``
    vector of a torch_geometric Batch); stays on the device."""
    edges = torch.arange(n_events + 1, device=event_of_row.device, dtype=event_of_row.dtype)
    return torch.searchsorted(event_of_row.contiguous(), edges).to(torch.int32)
