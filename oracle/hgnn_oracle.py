"""ORACLE — TEST INFRASTRUCTURE ONLY.

A CPU restatement, in plain functional torch, of the reference's HGNN
message-passing hot path. Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module; the product package (``hierarchicalgnn_b200``) never does and has no
CPU fallback.

Pinning: the reference ships no tests or golden vectors, so this
restatement is pinned against *outputs of the reference itself*: the fixtures in
``tests/golden/`` were produced by ``oracle/make_golden.py``, which imports the
unmodified reference modules from /root/reference (third-party CUDA packages
stubbed by ``oracle/stubs``) and records inputs, state dicts, outputs and
gradients. ``tests/test_oracle_golden.py`` checks every function here against
them.

Everything operates on a flat ``state_dict`` (same keys as the reference
modules) + the hparams dict, so no nn.Module of this repo is involved.
Citations are into /root/reference/Modules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------
# third-party primitives (torch_scatter / frnn / cugraph semantics)
# --------------------------------------------------------------------------


def scatter_add(src: Tensor, index: Tensor, n: int) -> Tensor:
    """torch_scatter.scatter_add(src, index, dim=0, dim_size=n)."""
    out = src.new_zeros((n,) + tuple(src.shape[1:]))
    return out.index_add_(0, index, src)


def scatter_mean(src: Tensor, index: Tensor, n: int) -> Tensor:
    """torch_scatter.scatter_mean: sum / clamp(count, 1)."""
    total = scatter_add(src, index, n)
    cnt = torch.bincount(index, minlength=n).clamp(min=1).to(src.dtype)
    return total / cnt.reshape([n] + [1] * (src.dim() - 1))


def knn_radius(queries: Tensor, refs: Tensor, k: int, radius: float, chunk: int = 1024) -> Tensor:
    """frnn.frnn_grid_points as consumed at utils.py:228-239: [P1, k] int64,
    ascending distance, Euclidean distance < radius, -1 padded, ties -> smaller
    index. Distances in float64, direct sum of squared differences."""
    q = queries.detach().double()
    r = refs.detach().double()
    r2 = float(radius) ** 2
    kk = min(k, r.shape[0])
    out = torch.full((q.shape[0], k), -1, dtype=torch.long)
    for s in range(0, q.shape[0], chunk):
        d2 = (q[s:s + chunk, None, :] - r[None, :, :]).square().sum(-1)
        srt = torch.sort(d2, dim=1, stable=True)
        keep = srt.values[:, :kk] < r2
        out[s:s + chunk, :kk] = torch.where(keep, srt.indices[:, :kk], torch.full_like(srt.indices[:, :kk], -1))
    return out


def knn_margin(queries: Tensor, refs: Tensor, k: int, radius: float) -> float:
    """Smallest relative gap that could flip the kNN answer under fp32 rounding:
    between rank k-1 and rank k distances, between any two retained neighbours,
    and between any retained distance and the radius. Used to certify that a
    seed has no near-ties."""
    q, r = queries.double(), refs.double()
    d2 = (q[:, None, :] - r[None, :, :]).square().sum(-1)
    srt = torch.sort(d2, dim=1).values[:, : min(k + 1, r.shape[0])]
    gaps = (srt[:, 1:] - srt[:, :-1]) / srt[:, 1:].clamp(min=1e-30)
    rad = ((srt - float(radius) ** 2).abs() / float(radius) ** 2).min()
    return float(min(gaps.min(), rad))


def symmetrize(graph: Tensor) -> Tensor:
    """cugraph symmetrize as used at gnn_utils.py:198-199 — union with the
    transpose, de-duplicated; canonical lexicographic column order."""
    both = torch.cat([graph, graph.flip(0)], dim=1)
    return torch.unique(both, dim=1)


def canonical_edge_order(graph: Tensor) -> Tensor:
    """Permutation sorting columns lexicographically on (row0, row1)."""
    key = graph[0].to(torch.int64) * (int(graph[1].max()) + 1 if graph.numel() else 1) + graph[1]
    return torch.argsort(key, stable=True)


def connected_component_labels(graph: Tensor, n: int) -> Tensor:
    """Weakly connected components; label = min vertex id; vertices that occur
    in no edge get -1 (cugraph returns only vertices present in the edge list,
    BC/Models/HGNN_GMM.py:215-221)."""
    import numpy as np
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    src = graph[0].cpu().numpy()
    dst = graph[1].cpu().numpy()
    adj = coo_matrix((np.ones(len(src), dtype=np.int8), (src, dst)), shape=(n, n))
    _, lab = connected_components(adj, directed=False)
    first = np.full(lab.max() + 1, n, dtype=np.int64)
    np.minimum.at(first, lab, np.arange(n))
    out = torch.from_numpy(first[lab])
    present = torch.zeros(n, dtype=torch.bool)
    present[graph[0]] = True
    present[graph[1]] = True
    out[~present] = -1
    return out


def cluster_labels_from_components(labels: Tensor, min_cluster_size: int) -> Tensor:
    """get_cluster_labels (BC/Models/HGNN_GMM.py:172-181): drop components with
    fewer than min_cluster_size members, renumber survivors by ascending label."""
    clusters = torch.full_like(labels, -1)
    valid = labels >= 0
    uniq, inv, cnt = labels[valid].unique(return_inverse=True, return_counts=True)
    big = cnt[inv] >= min_cluster_size
    idx = valid.nonzero().squeeze(1)[big]
    clusters[idx] = labels[idx].unique(return_inverse=True)[1]
    return clusters


# --------------------------------------------------------------------------
# make_mlp layout (utils.py:169-196)
# --------------------------------------------------------------------------

_ACT = {
    "GELU": lambda t: F.gelu(t),  # nn.GELU() default = exact erf
    "Tanh": torch.tanh,
    "ReLU": torch.relu,
    "SiLU": F.silu,
    "Sigmoid": torch.sigmoid,
    None: lambda t: t,
}


def mlp_layout(n_layers: int, hidden_act: str, output_act: Optional[str], layer_norm: bool):
    """Sequential indices produced by make_mlp: per layer
    (linear_index, layernorm_index_or_None, activation_name_or_None)."""
    out, i = [], 0
    for _ in range(n_layers - 1):
        lin = i
        i += 1
        ln = None
        if layer_norm:
            ln = i
            i += 1
        i += 1  # the activation module occupies an index
        out.append((lin, ln, hidden_act))
    lin = i
    i += 1
    ln = None
    if output_act is not None and layer_norm:
        ln = i
        i += 1
    out.append((lin, ln, output_act))
    return out


def mlp_apply(sd: Dict[str, Tensor], prefix: str, x: Tensor, n_layers: int, hidden_act: str,
              output_act: Optional[str], layer_norm: bool, eps: float = 1e-5) -> Tensor:
    for lin, ln, act in mlp_layout(n_layers, hidden_act, output_act, layer_norm):
        w, b = sd[f"{prefix}.{lin}.weight"], sd[f"{prefix}.{lin}.bias"]
        x = x @ w.t() + b
        if ln is not None:
            x = F.layer_norm(x, (x.shape[-1],), sd[f"{prefix}.{ln}.weight"], sd[f"{prefix}.{ln}.bias"], eps)
        x = _ACT[act](x)
    return x


# --------------------------------------------------------------------------
# cells (gnn_utils.py)
# --------------------------------------------------------------------------


def edge_step(sd, prefix, hp, nodes, edges, graph):
    """InteractionGNNCell.edge_update / HierarchicalGNNCell.edge_update /
    superedge_update (gnn_utils.py:56-64, 129-135, 147-153)."""
    inp = torch.cat([nodes[graph[0]], nodes[graph[1]], edges], dim=-1)
    return mlp_apply(sd, prefix, inp, hp["nb_edge_layer"], hp["hidden_activation"], "Tanh", hp["layernorm"]) + edges


def interaction_cell(sd, prefix, hp, nodes, edges, graph):
    """InteractionGNNCell.forward (gnn_utils.py:45-71): node step on the
    incoming-edge sum, then edge step on the updated nodes."""
    agg = scatter_add(edges, graph[1], nodes.shape[0])
    nodes = mlp_apply(sd, prefix + ".node_network", torch.cat([nodes, agg], -1), hp["nb_node_layer"],
                      hp["hidden_activation"], hp["hidden_activation"], hp["layernorm"]) + nodes
    edges = edge_step(sd, prefix + ".edge_network", hp, nodes, edges, graph)
    return nodes, edges


def hierarchical_cell(sd, prefix, hp, nodes, edges, supernodes, superedges, graph, bgraph, bweights,
                      sgraph, sweights):
    """HierarchicalGNNCell.forward (gnn_utils.py:119-169): supernode -> node ->
    superedge -> edge, each consuming the freshest upstream tensors."""
    S, N = supernodes.shape[0], nodes.shape[0]
    act, ln = hp["hidden_activation"], hp["layernorm"]
    up = scatter_add(bweights * nodes[bgraph[0]], bgraph[1], S)
    att = scatter_add(superedges * sweights, sgraph[1], S)
    supernodes = mlp_apply(sd, prefix + ".supernode_network", torch.cat([supernodes, att, up], -1),
                           hp["nb_node_layer"], act, act, ln) + supernodes
    down = scatter_add(bweights * supernodes[bgraph[1]], bgraph[0], N)
    agg = scatter_add(edges, graph[1], N)
    nodes = mlp_apply(sd, prefix + ".node_network", torch.cat([nodes, agg, down], -1),
                      hp["nb_node_layer"], act, act, ln) + nodes
    superedges = edge_step(sd, prefix + ".superedge_network", hp, supernodes, superedges, sgraph)
    edges = edge_step(sd, prefix + ".edge_network", hp, nodes, edges, graph)
    return nodes, edges, supernodes, superedges


def dynamic_graph(sd, prefix, src_emb, dst_emb, *, weighting: str, sym: bool, norm: bool, k: int,
                  training: bool, graph: Optional[Tensor] = None, momentum: float = 0.1, eps: float = 1e-5):
    """DynamicGraphConstruction.forward (gnn_utils.py:183-218).

    Returns (graph, edge_weights[E,1], logits[E], new_buffers). ``graph`` may be
    injected to compare the differentiable half on a fixed edge list."""
    bufs = {}
    radius = sd[prefix + ".knn_radius"]
    with torch.no_grad():
        if graph is None:
            idx = knn_radius(src_emb, dst_emb, k, float(radius))
            rows = torch.arange(idx.shape[0]).unsqueeze(1).expand_as(idx)
            ok = idx >= 0
            graph = torch.stack([rows[ok], idx[ok]], 0)
            if sym:
                graph = symmetrize(graph)
        if training:
            dmax = (src_emb[graph[0]] - dst_emb[graph[1]]).square().sum(-1).sqrt().max()
            bufs[prefix + ".knn_radius"] = (0.9 * radius + 0.11 * dmax).to(radius.dtype)
    dots = (src_emb[graph[0]] * dst_emb[graph[1]]).sum(-1)
    bn = prefix + ".weight_normalization"
    gamma, beta = sd[bn + ".weight"], sd[bn + ".bias"]
    if training:
        mean = dots.mean()
        var_b = dots.var(unbiased=False)
        n = dots.numel()
        with torch.no_grad():
            bufs[bn + ".running_mean"] = (1 - momentum) * sd[bn + ".running_mean"] + momentum * mean.detach()
            bufs[bn + ".running_var"] = (1 - momentum) * sd[bn + ".running_var"] + momentum * dots.detach().var(unbiased=True) if n > 1 else sd[bn + ".running_var"]
            bufs[bn + ".num_batches_tracked"] = sd[bn + ".num_batches_tracked"] + 1
    else:
        mean, var_b = sd[bn + ".running_mean"][0], sd[bn + ".running_var"][0]
    logits = (dots - mean) / torch.sqrt(var_b + eps) * gamma[0] + beta[0]
    w = torch.sigmoid(logits) if weighting == "sigmoid" else torch.exp(logits)
    if norm:
        w = w / w.mean()
    return graph, w.unsqueeze(1), logits, bufs


# --------------------------------------------------------------------------
# blocks and models
# --------------------------------------------------------------------------


def ignn_block(sd, prefix, hp, x, graph, n_iters: int, emb: bool):
    """InteractionGNNBlock.forward (EC/Models/IN.py:80-95; BC/Models/HGNN_GMM.py:86-99)."""
    act, ln = hp["hidden_activation"], hp["layernorm"]
    nodes = mlp_apply(sd, prefix + ".node_encoder", x, hp["nb_node_layer"], act, act, ln)
    edges = mlp_apply(sd, prefix + ".edge_encoder", torch.cat([x[graph[0]], x[graph[1]]], 1),
                      hp["nb_edge_layer"], act, act, ln)
    for i in range(n_iters):
        cell = f"{prefix}.ignn_cells.{i}"
        nodes, edges = interaction_cell(sd, cell, hp, nodes, edges, graph)
    if not emb:
        return nodes, edges
    e = mlp_apply(sd, prefix + ".output_layer", nodes, hp["output_layers"], hp["hidden_output_activation"], None, ln)
    return F.normalize(e), nodes, edges


def ec_forward(sd, hp, x, graph):
    """EC_InteractionGNN.forward (EC/Models/IN.py:118-128)."""
    E = graph.shape[1]
    directed = torch.cat([graph, graph.flip(0)], 1)
    nodes, edges = ignn_block(sd, "ignn_block", hp, x, directed, hp["n_interaction_graph_iters"], emb=False)
    pair = torch.cat([edges[:E], edges[E:]], 1)
    s = mlp_apply(sd, "edge_classifier", pair, hp["output_layers"], hp["hidden_output_activation"], None, hp["layernorm"])
    return torch.sigmoid(s.squeeze(-1))


def gmm_clustering(hp, embeddings, graph, score_cut: Tensor, training: bool, random_state=0):
    """HierarchicalGNNBlock.clustering (BC/Models/HGNN_GMM.py:162-234) restated.
    Returns (clusters[N], new_score_cut). Parity through this function is
    statistical only (sklearn GMM has random_state=None in the reference)."""
    import numpy as np
    from scipy.optimize import fsolve
    from sklearn.mixture import GaussianMixture
    with torch.no_grad():
        lik = (embeddings[graph[0]] * embeddings[graph[1]]).sum(-1)
        lik = torch.atanh(lik.clamp(-1 + 1e-7, 1 - 1e-7))
        gmm = GaussianMixture(n_components=2, random_state=random_state)
        gmm.fit(lik.unsqueeze(1).cpu().numpy())
        lo, hi = int(gmm.means_.argmin()), int(gmm.means_.argmax())
        mid = float(gmm.means_.mean())
        if bool(torch.isinf(score_cut).all()):
            score_cut = torch.tensor([mid], dtype=score_cut.dtype)
        gran = hp["cluster_granularity"]
        sg = lambda v: 1.0 / (1.0 + math.exp(-v))

        def balance(t):
            p = gmm.predict_proba(np.asarray(t).reshape(-1, 1))
            return sg(gran) * p[:, lo] - sg(-gran) * p[:, hi]

        def solve(t0):
            return float(fsolve(balance, t0).item())

        inside = lambda c: float(gmm.means_.min()) < c < float(gmm.means_.max())
        cut = solve(float(score_cut))
        if training and inside(cut):
            score_cut = 0.95 * score_cut + 0.05 * cut
        else:
            cut = solve(mid)
            if training and inside(cut):
                score_cut = 0.95 * score_cut + 0.05 * cut
        N = embeddings.shape[0]
        keep = lik >= score_cut.to(lik.dtype)
        clusters = None
        if bool(keep.any()):
            clusters = cluster_labels_from_components(
                connected_component_labels(graph[:, keep], N), hp["min_cluster_size"])
        if clusters is None or int(clusters.max()) <= 2:
            clusters = cluster_labels_from_components(connected_component_labels(graph, N), hp["min_cluster_size"])
        return clusters, score_cut


def hgnn_block(sd, prefix, hp, embeddings, nodes, edges, graph, clusters, training: bool,
               super_graph=None, bipartite_graph=None):
    """HierarchicalGNNBlock.forward after clustering (BC/Models/HGNN_GMM.py:247-298).
    ``clusters`` is injected (parity must fix the clustering input)."""
    act, ln = hp["hidden_activation"], hp["layernorm"]
    sel = clusters >= 0
    S = int(clusters.max()) + 1
    means = F.normalize(scatter_mean(embeddings[sel], clusters[sel], S))
    sg, sw, _, b1 = dynamic_graph(sd, prefix + ".super_graph_construction", means, means, weighting="sigmoid",
                                  sym=True, norm=True, k=hp["supergraph_sparsity"], training=training, graph=super_graph)
    bg, bw, blog, b2 = dynamic_graph(sd, prefix + ".bipartite_graph_construction", embeddings, means, weighting="exp",
                                     sym=False, norm=True, k=hp["bipartitegraph_sparsity"], training=training,
                                     graph=bipartite_graph)
    sn0 = scatter_add(F.normalize(nodes, p=1)[bg[0]] * bw, bg[1], S)
    enc = mlp_apply(sd, prefix + ".supernode_encoder", sn0, hp["nb_node_layer"], act, act, ln)
    supernodes = torch.cat([means, enc], -1)
    superedges = mlp_apply(sd, prefix + ".superedge_encoder", torch.cat([supernodes[sg[0]], supernodes[sg[1]]], 1),
                           hp["nb_edge_layer"], act, act, ln)
    for i in range(hp["n_hierarchical_graph_iters"]):
        nodes, edges, supernodes, superedges = hierarchical_cell(
            sd, f"{prefix}.hgnn_cells.{i}", hp, nodes, edges, supernodes, superedges, graph, bg, bw, sg, sw)
    bufs = dict(b1)
    bufs.update(b2)
    return nodes, supernodes, bg, bufs, dict(super_graph=sg, super_weights=sw, bipartite_weights=bw,
                                              bipartite_logits=blog, means=means)


def bc_forward(sd, hp, x, graph, clusters=None, training=True, super_graph=None, bipartite_graph=None,
               return_aux=False):
    """BC_HierarchicalGNN_GMM.forward (BC/Models/HGNN_GMM.py:323-346)."""
    directed = torch.cat([graph, graph.flip(0)], 1)
    emb, nodes, edges = ignn_block(sd, "ignn_block", hp, x, directed, hp["n_interaction_graph_iters"], emb=True)
    bufs = {}
    if clusters is None:
        clusters, cut = gmm_clustering(hp, emb, directed, sd["hgnn_block.score_cut"], training)
        bufs["hgnn_block.score_cut"] = cut
    nodes, supernodes, bg, b, aux = hgnn_block(sd, "hgnn_block", hp, emb, nodes, edges, directed, clusters, training,
                                               super_graph, bipartite_graph)
    bufs.update(b)
    pair = torch.cat([nodes[bg[0]], supernodes[bg[1]]], 1)
    s = mlp_apply(sd, "bipartite_output_layer", pair, hp["output_layers"], hp["hidden_output_activation"], None,
                  hp["layernorm"])
    out = (bg, torch.sigmoid(s.squeeze(-1)), emb)
    if return_aux:
        aux.update(buffers=bufs, clusters=clusters, nodes=nodes, supernodes=supernodes)
        return out + (aux,)
    return out


# --------------------------------------------------------------------------
# helpers for tests / baselines
# --------------------------------------------------------------------------


def cast_state(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def leaf_state(sd):
    """Detach + requires_grad on every floating tensor that is a parameter-like
    entry (buffers of BN / radius / score_cut are left alone)."""
    out = {}
    for k, v in sd.items():
        is_buf = any(t in k for t in ("running_", "num_batches", "knn_radius", "score_cut"))
        out[k] = v.detach().clone().requires_grad_(True) if (v.is_floating_point() and not is_buf) else v.detach().clone()
    return out


def edge_step_cell_fwd_bwd(sd, prefix, hp, nodes, edges, graph):
    """The config-2 microbenchmark unit on CPU: edge step + scatter_add of the
    new edges, forward and backward (used as the timed CPU baseline)."""
    nodes = nodes.detach().requires_grad_(True)
    edges = edges.detach().requires_grad_(True)
    e2 = edge_step(sd, prefix, hp, nodes, edges, graph)
    agg = scatter_add(e2, graph[1], nodes.shape[0])
    (e2.sum() + agg.square().sum()).backward()
    return e2.detach(), agg.detach(), nodes.grad, edges.grad


def roc_auc(scores: Tensor, labels: Tensor) -> float:
    """Rank-based AUC (Mann-Whitney), ties averaged."""
    s = scores.double()
    order = torch.argsort(s)
    ranks = torch.empty_like(s)
    ranks[order] = torch.arange(1, len(s) + 1, dtype=torch.double)
    uniq, inv = torch.unique(s, return_inverse=True)
    mean_rank = scatter_mean(ranks, inv, len(uniq))[inv]
    pos = labels.bool()
    n1, n0 = int(pos.sum()), int((~pos).sum())
    if n1 == 0 or n0 == 0:
        return float("nan")
    return float((mean_rank[pos].sum() - n1 * (n1 + 1) / 2) / (n1 * n0))


def eval_metrics_dense(bipartite_graph: Tensor, pid_all: Tensor, pt_all: Tensor, pt_cut=1.0, nhits_cut=5, majority_cut=0.5):
    """Tracking metrics, restated line by line from the reference (Modules/tracking_utils.py:18-83, primary=False as both
    shared_evaluation callers pass) with a DENSE numpy particle x candidate matrix in place of the cupy sparse one —
    including its ``cluster_hashing`` tie-break (a factor rising from 1 to 1 + 1e-12 over the candidates). Small cases only."""
    import numpy as np
    bg = bipartite_graph.clone().long()
    _, clusters, counts = bg[1].unique(return_inverse=True, return_counts=True)
    bg = bg[:, counts[clusters] >= (nhits_cut * majority_cut)]
    zero = {"track_eff": 0, "track_pur": 0, "hit_eff": 0, "hit_pur": 0}
    if bg.shape[1] == 0:
        return zero
    bg[1] = bg[1].unique(return_inverse=True)[1]
    original_pid, pid, nhits = torch.unique(pid_all, return_inverse=True, return_counts=True)
    n_p, n_c = int(pid.max()) + 1, int(bg[1].max()) + 1
    pt = np.full(n_p, np.inf)
    for h in range(pid.numel()):
        pt[int(pid[h])] = min(pt[int(pid[h])], float(pt_all[h]))
    M = np.zeros((n_p, n_c))
    for h, c in zip(pid[bg[0]].tolist(), bg[1].tolist()):
        M[h, c] += 1.0
    original_pid, nhits_np = original_pid.numpy(), nhits.numpy().astype(np.float64)
    hashing = np.linspace(1, 1 + 1e-12, n_c).reshape(1, -1)
    Mh = M * hashing
    matching = (M >= majority_cut * M.sum(0, keepdims=True)) & (M >= majority_cut * nhits_np.reshape(-1, 1)) & (Mh == Mh.max(1, keepdims=True))
    row, col = np.where(matching)
    if row.shape[0] == 0:
        return zero
    matching_mask = (M[row, col] > majority_cut * nhits_cut) & (original_pid[row] != 0)
    row, col = row[matching_mask], col[matching_mask]
    if row.shape[0] == 0:
        return zero
    mask = (pt[row] > pt_cut) & (nhits_np[row] >= nhits_cut)
    truth_mask = (pt > pt_cut) & (nhits_np >= nhits_cut)
    track_eff = mask.sum() / truth_mask.sum()
    hit_pur = (M[row, col] / M[:, col].sum(0)).mean()
    track_pur = mask.sum() / (M.shape[1] - (~matching_mask).sum() - (~mask).sum())
    hit_eff = (M[row, col][mask] / nhits_np[row][mask]).mean()
    return {"track_eff": float(track_eff), "track_pur": float(track_pur), "hit_eff": float(hit_eff), "hit_pur": float(hit_pur)}
