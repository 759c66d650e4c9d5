"""Hierarchical GNN with GMM clustering (reference:
BipartiteClassification/Models/HGNN_GMM.py), on the hgnn_b200 kernels.

Same classes, constructor arguments, forward signatures, return values and
state-dict keys. What changed underneath:
  * cells / encoders / heads: fused gather-MLP kernels and segmented reductions;
  * supergraph + bipartite graph: brute-force radius-kNN, sort/unique symmetrize;
  * clustering: on-device 1-D GMM (EM) + single-pass union-find components, in
    place of sklearn-on-CPU + cugraph (the cut equation, which the reference
    solves with scipy.fsolve, has a closed form for two 1-D Gaussians);
  * the last cell's edge / superedge updates, whose results the reference
    discards (HGNN_GMM.py:298) and which receive no gradient, are skipped.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from ... import ops
from ...gnn_utils import (DynamicGraphConstruction, GraphPlans, HierarchicalGNNCell, InteractionGNNCell,
                          sort_edges_by_destination)
from ...utils import event_offsets, make_mlp
from ..bipartite_classification_base import BipartiteClassificationBase


class InteractionGNNBlock(nn.Module):
    """Interaction network with the embedding head always on (HGNN_GMM.py:23-99)."""

    def __init__(self, hparams, iterations):
        super().__init__()
        act, ln = hparams["hidden_activation"], hparams["layernorm"]
        C, L, H = hparams["spatial_channels"], hparams["latent"], hparams["hidden"]
        self.node_encoder = make_mlp(C, H, L, hparams["nb_node_layer"], output_activation=act, hidden_activation=act,
                                     layer_norm=ln)
        self.edge_encoder = make_mlp(2 * C, H, L, hparams["nb_edge_layer"], layer_norm=ln, output_activation=act,
                                     hidden_activation=act)
        if hparams["share_weight"]:
            shared = InteractionGNNCell(hparams)
            cells = [shared] * iterations
        else:
            cells = [InteractionGNNCell(hparams) for _ in range(iterations)]
        self.ignn_cells = nn.ModuleList(cells)
        self.output_layer = make_mlp(L, H, hparams["emb_dim"], hparams["output_layers"], layer_norm=ln,
                                     output_activation=None, hidden_activation=hparams["hidden_output_activation"])
        self.hparams = hparams

    def forward(self, x, graph):
        gp = graph if isinstance(graph, GraphPlans) else GraphPlans(graph, x.shape[0], x.shape[0])
        if torch.is_grad_enabled() and x.is_leaf:
            x.requires_grad = True
        nodes = self.node_encoder(x)
        edges = self.edge_encoder.fused([x, x], [gp.by_src, gp.by_dst])
        for cell in self.ignn_cells:
            nodes, edges = cell(nodes, edges, gp)
        embeddings = nn.functional.normalize(self.output_layer(nodes))
        return embeddings, nodes, edges


def gaussian_cut(params, granularity):
    """Point between the two component means where
    sigmoid(g) * P(low | x) == sigmoid(-g) * P(high | x)  (HGNN_GMM.py:162-170).
    For two 1-D Gaussians this is a quadratic in x; returns (cut, found)."""
    pi0, mu0, v0, pi1, mu1, v1 = [float(p) for p in params]
    if mu0 > mu1:
        pi0, mu0, v0, pi1, mu1, v1 = pi1, mu1, v1, pi0, mu0, v0
    sg = lambda t: 1.0 / (1.0 + math.exp(-t))
    a = -0.5 / v0 + 0.5 / v1
    b = mu0 / v0 - mu1 / v1
    c = (-0.5 * mu0 * mu0 / v0 + 0.5 * mu1 * mu1 / v1
         + math.log(max(sg(granularity) * pi0, 1e-300) / math.sqrt(v0))
         - math.log(max(sg(-granularity) * pi1, 1e-300) / math.sqrt(v1)))
    roots = []
    if abs(a) < 1e-12:
        if abs(b) > 1e-30:
            roots = [-c / b]
    else:
        disc = b * b - 4 * a * c
        if disc >= 0:
            s = math.sqrt(disc)
            roots = [(-b + s) / (2 * a), (-b - s) / (2 * a)]
    inside = [r for r in roots if mu0 < r < mu1]
    if inside:
        mid = 0.5 * (mu0 + mu1)
        return min(inside, key=lambda r: abs(r - mid)), True
    return 0.5 * (mu0 + mu1), False


class HierarchicalGNNBlock(nn.Module):
    def __init__(self, hparams, logging):
        super().__init__()
        act, ln = hparams["hidden_activation"], hparams["layernorm"]
        L, H, D = hparams["latent"], hparams["hidden"], hparams["emb_dim"]
        self.supernode_encoder = make_mlp(L, H, L - D, hparams["nb_node_layer"], output_activation=act,
                                          hidden_activation=act, layer_norm=ln)
        self.superedge_encoder = make_mlp(2 * L, H, L, hparams["nb_edge_layer"], layer_norm=ln, output_activation=act,
                                          hidden_activation=act)
        n = hparams["n_hierarchical_graph_iters"]
        if hparams["share_weight"]:
            shared = HierarchicalGNNCell(hparams)
            cells = [shared] * n
        else:
            cells = [HierarchicalGNNCell(hparams) for _ in range(n)]
        self.hgnn_cells = nn.ModuleList(cells)
        self.super_graph_construction = DynamicGraphConstruction("sigmoid", hparams)
        self.bipartite_graph_construction = DynamicGraphConstruction("exp", hparams)
        self.register_buffer("score_cut", torch.tensor([float("inf")]))
        self.log = logging
        self.hparams = hparams

    def get_cluster_labels(self, labels, n):
        """Drop components smaller than min_cluster_size and renumber the survivors by
        ascending component label (HGNN_GMM.py:172-181). ``labels``: int32 [n], -1 = absent."""
        labels = labels.long()
        clusters = torch.full((n,), -1, dtype=torch.long, device=labels.device)
        vertex = (labels >= 0).nonzero().squeeze(1)
        if vertex.numel() == 0:
            return clusters
        lab = labels[vertex]
        _, inverse, counts = lab.unique(return_inverse=True, return_counts=True)
        big = counts[inverse] >= self.hparams["min_cluster_size"]
        if bool(big.any()):
            clusters[vertex[big]] = lab[big].unique(return_inverse=True)[1]
        return clusters

    def clustering(self, x, embeddings, graph):
        gp = graph if isinstance(graph, GraphPlans) else GraphPlans(graph, x.shape[0], x.shape[0])
        g = gp.graph
        with torch.no_grad():
            emb = embeddings.detach()
            likelihood = ops.edge_dot_raw(emb, gp.by_src.keys32, emb, gp.by_dst.keys32, g.shape[1])
            likelihood = torch.atanh(likelihood.clamp(-1 + 1e-7, 1 - 1e-7))
            params = ops.gmm1d_fit(likelihood).tolist()  # one 24-byte D2H copy
            mu_lo, mu_hi = min(params[1], params[4]), max(params[1], params[4])
            if bool(torch.isinf(self.score_cut).all()):
                self.score_cut = torch.full_like(self.score_cut, 0.5 * (mu_lo + mu_hi))
            cut, found = gaussian_cut(params, self.hparams["cluster_granularity"])
            if self.training and found:
                self.score_cut = 0.95 * self.score_cut + 0.05 * cut
            self.log("score_cut", self.score_cut.item())
            keep = likelihood >= self.score_cut.to(likelihood.device)
            clusters = self.get_cluster_labels(ops.connected_components(g, x.shape[0], keep), x.shape[0])
            if int(clusters.max()) <= 2:
                # every edge cut away (or nearly): fall back to the uncut graph (HGNN_GMM.py:224-232)
                clusters = self.get_cluster_labels(ops.connected_components(g, x.shape[0], None), x.shape[0])
            return clusters

    def forward(self, x, embeddings, nodes, edges, graph, clusters=None, batch=None, n_events=None):
        """``batch`` (ascending int64 event id of every hit) + ``n_events``: several events as one disjoint graph. Supernodes
        never mix events (the clustering follows the edges), the two kNN graphs are searched event by event and the edge
        weights normalised per event, so every hit and supernode sees what a call on its own event would have shown it.
        Injected ``clusters`` must number the supernodes event by event (ascending)."""
        N = x.shape[0]
        gp = graph if isinstance(graph, GraphPlans) else GraphPlans(graph, N, N)
        if clusters is None:
            clusters = self.clustering(x, embeddings, gp)
        member = clusters >= 0
        S = int(clusters.max()) + 1
        means = ops.scatter_mean(embeddings[member], clusters[member], dim_size=S)
        means = nn.functional.normalize(means)

        hp = self.hparams
        sup_kw, bip_kw = {}, {}
        if batch is not None:
            if n_events is None:
                n_events = int(batch[-1]) + 1
            with torch.no_grad():
                hit_ptr = event_offsets(batch, n_events)
                sn_event = torch.zeros(S, dtype=batch.dtype, device=batch.device).scatter_(0, clusters[member], batch[member])
                if os.environ.get("HGNN_CHECK_INDICES", "0") != "0" and S > 1 and not bool((sn_event[1:] >= sn_event[:-1]).all()):
                    raise ValueError("batched events: supernode (cluster) ids must ascend with the event id")
                sn_ptr = event_offsets(sn_event, n_events)
            sup_kw = dict(src_ptr=sn_ptr, dst_ptr=sn_ptr, src_event=sn_event)
            bip_kw = dict(src_ptr=hit_ptr, dst_ptr=sn_ptr, src_event=batch)
        super_graph, super_edge_weights = self.super_graph_construction(
            means, means, sym=True, norm=True, k=hp["supergraph_sparsity"], **sup_kw)
        bipartite_graph, bipartite_edge_weights, _logits = self.bipartite_graph_construction(
            embeddings, means, sym=False, norm=True, k=hp["bipartitegraph_sparsity"], logits=True, **bip_kw)
        self.log("clusters", len(means))
        bp = GraphPlans(bipartite_graph, N, S)
        sp = GraphPlans(super_graph, S, S)

        pooled = ops.gather_scatter(nn.functional.normalize(nodes, p=1), bipartite_edge_weights, bp.by_src, bp.by_dst)
        supernodes = torch.cat([means, self.supernode_encoder(pooled)], dim=-1)
        superedges = self.superedge_encoder.fused([supernodes, supernodes], [sp.by_src, sp.by_dst])
        last = len(self.hgnn_cells) - 1
        for i, cell in enumerate(self.hgnn_cells):
            nodes, edges, supernodes, superedges = cell(nodes, edges, supernodes, superedges, gp, bp,
                                                        bipartite_edge_weights, sp, super_edge_weights,
                                                        skip_edge_updates=(i == last))
        return nodes, supernodes, bipartite_graph


class BC_HierarchicalGNN_GMM(BipartiteClassificationBase):
    def __init__(self, hparams):
        super().__init__(hparams)
        self.ignn_block = InteractionGNNBlock(hparams, hparams["n_interaction_graph_iters"])
        self.hgnn_block = HierarchicalGNNBlock(hparams, self.log)
        self.bipartite_output_layer = make_mlp(2 * hparams["latent"], hparams["hidden"], 1, hparams["output_layers"],
                                               layer_norm=hparams["layernorm"], output_activation=None,
                                               hidden_activation=hparams["hidden_output_activation"])

    def forward(self, x, graph, clusters=None, batch=None, n_events=None):
        """``batch`` / ``n_events``: a torch_geometric-style batch of events (x rows event by event, ``graph`` already offset,
        ``batch`` = event id per hit) processed as one disjoint graph — see HierarchicalGNNBlock.forward."""
        N = x.shape[0]
        # destination-sorted once per event; nothing downstream depends on the edge order (HGNN_GMM.py:328-346)
        directed = GraphPlans(sort_edges_by_destination(torch.cat([graph, graph.flip(0)], dim=1))[0], N, N, dst_sorted=True)
        embeddings, nodes, edges = self.ignn_block(x, directed)
        nodes, supernodes, bipartite_graph = self.hgnn_block(x, embeddings, nodes, edges, directed, clusters=clusters,
                                                             batch=batch, n_events=n_events)
        bp = GraphPlans(bipartite_graph, N, supernodes.shape[0])
        scores = self.bipartite_output_layer.fused([nodes, supernodes], [bp.by_src, bp.by_dst]).squeeze()
        return bipartite_graph, torch.sigmoid(scores), embeddings
