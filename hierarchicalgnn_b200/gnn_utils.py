"""Drop-in message-passing cells (reference: Modules/gnn_utils.py).

Same class names, constructor arguments, forward signatures, return order and
state-dict keys as the reference; the arithmetic runs in hand-written CUDA:

* every ``scatter_add`` is a deterministic destination-sorted segmented
  reduction over a CSR plan built once per graph (no float atomics);
* every ``network(torch.cat([...gathers...]))`` is one fused gather -> MLP ->
  LayerNorm -> activation -> skip kernel (no [E,3L] concat in HBM);
* ``torch.utils.checkpoint`` is replaced by recompute inside the fused backward.
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from . import ops
from .utils import make_mlp, find_neighbors


class GraphPlans:
    """Segment plans of one edge list [2, E]: by destination (row 1, over n_dst
    segments) and by source (row 0, over n_src segments). Built lazily, cached."""

    def __init__(self, graph: torch.Tensor, n_src: int, n_dst: int, dst_sorted: bool = False):
        self.graph, self.n_src, self.n_dst = graph, int(n_src), int(n_dst)
        self._src = self._dst = None
        self.dst_sorted = bool(dst_sorted)  # caller guarantees graph[1] ascends (sort_edges_by_destination): no host check

    @property
    def by_src(self) -> ops.SegmentPlan:
        if self._src is None:
            self._src = ops.plan_for(self.graph[0], self.n_src)
        return self._src

    @property
    def by_dst(self) -> ops.SegmentPlan:
        if self._dst is None:
            self._dst = ops.plan_for(self.graph[1], self.n_dst)
            if self.dst_sorted:
                self._dst._identity = True
        return self._dst


def sort_edges_by_destination(graph: torch.Tensor):
    """Returns (sorted_graph, order, inverse): the edge list reordered so destinations ascend (stable), the permutation
    that produced it and its inverse. The models do this once per event: every cell then streams edge rows in place
    (identity plan permutation) and the fused segmented reduce sees contiguous runs."""
    order = torch.argsort(graph[1], stable=True)
    inverse = torch.empty_like(order)
    inverse[order] = torch.arange(order.numel(), device=order.device)
    return graph[:, order].contiguous(), order, inverse


def _plans(graph, n_src, n_dst):
    return graph if isinstance(graph, GraphPlans) else GraphPlans(graph, n_src, n_dst)


# agg = scatter_add(e', dst) produced as a by-product of the edge step that created e' (cell i), consumed by the node
# update that opens cell i+1 (SURVEY §3.2 "fusion crosses the cell boundary"). Keyed on the identity of e'.
_AGG_CACHE = {}
_AGG_CACHE_MAX = 4


def _edge_update(network, nodes, edges, gp: GraphPlans):
    # e' = MLP([x[src] | x[dst] | e]) + e      (gnn_utils.py:56-64)
    new_edges, agg = network.edge_step(nodes, edges, gp.by_src, gp.by_dst)
    if agg is not None:
        while len(_AGG_CACHE) >= _AGG_CACHE_MAX:
            _AGG_CACHE.pop(next(iter(_AGG_CACHE)))
        key = id(new_edges)
        # the entry dies with e' (an aggregate nobody consumed, e.g. the last cell's, must not pin device memory)
        _AGG_CACHE[key] = (weakref.ref(new_edges, lambda _r, k=key: _AGG_CACHE.pop(k, None)), gp.by_dst, agg)
    return new_edges


def _incoming_sum(edges, gp: GraphPlans, n_nodes):
    hit = _AGG_CACHE.pop(id(edges), None)
    if hit is not None and hit[0]() is edges and hit[1] is gp.by_dst:
        return hit[2]
    return ops.scatter_add(edges, gp.graph[1], dim_size=n_nodes, plan=gp.by_dst)


class InteractionGNNCell(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        L, H = hparams["latent"], hparams["hidden"]
        act, ln = hparams["hidden_activation"], hparams["layernorm"]
        self.edge_network = make_mlp(3 * L, H, L, hparams["nb_edge_layer"], layer_norm=ln,
                                     output_activation="Tanh", hidden_activation=act)
        self.node_network = make_mlp(2 * L, H, L, hparams["nb_node_layer"], layer_norm=ln,
                                     output_activation=act, hidden_activation=act)
        self.hparams = hparams

    def node_update(self, nodes, edges, graph):
        gp = _plans(graph, nodes.shape[0], nodes.shape[0])
        messages = _incoming_sum(edges, gp, nodes.shape[0])
        return self.node_network.fused([nodes, messages], skip=0)

    def edge_update(self, nodes, edges, graph):
        return _edge_update(self.edge_network, nodes, edges, _plans(graph, nodes.shape[0], nodes.shape[0]))

    def forward(self, nodes, edges, graph):
        gp = _plans(graph, nodes.shape[0], nodes.shape[0])
        nodes = self.node_update(nodes, edges, gp)
        edges = self.edge_update(nodes, edges, gp)
        return nodes, edges


class HierarchicalGNNCell(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        L, H = hparams["latent"], hparams["hidden"]
        act, ln = hparams["hidden_activation"], hparams["layernorm"]
        ne, nn_ = hparams["nb_edge_layer"], hparams["nb_node_layer"]
        self.edge_network = make_mlp(3 * L, H, L, ne, layer_norm=ln, output_activation="Tanh", hidden_activation=act)
        self.node_network = make_mlp(3 * L, H, L, nn_, layer_norm=ln, output_activation=act, hidden_activation=act)
        self.supernode_network = make_mlp(3 * L, H, L, nn_, layer_norm=ln, output_activation=act, hidden_activation=act)
        self.superedge_network = make_mlp(3 * L, H, L, ne, layer_norm=ln, output_activation="Tanh", hidden_activation=act)
        self.hparams = hparams

    def node_update(self, nodes, edges, supernodes, graph, bipartite_graph, bipartite_edge_weights):
        gp = _plans(graph, nodes.shape[0], nodes.shape[0])
        bp = _plans(bipartite_graph, nodes.shape[0], supernodes.shape[0])
        down = ops.gather_scatter(supernodes, bipartite_edge_weights, bp.by_dst, bp.by_src)
        messages = _incoming_sum(edges, gp, nodes.shape[0])
        return self.node_network.fused([nodes, messages, down], skip=0)

    def edge_update(self, nodes, edges, graph):
        return _edge_update(self.edge_network, nodes, edges, _plans(graph, nodes.shape[0], nodes.shape[0]))

    def supernode_update(self, nodes, supernodes, superedges, bipartite_graph, bipartite_edge_weights, super_graph,
                         super_edge_weights):
        S = supernodes.shape[0]
        bp = _plans(bipartite_graph, nodes.shape[0], S)
        sp = _plans(super_graph, S, S)
        up = ops.gather_scatter(nodes, bipartite_edge_weights, bp.by_src, bp.by_dst)
        attention = ops._GatherScatter.apply(superedges, super_edge_weights, None, sp.by_dst, False)
        return self.supernode_network.fused([supernodes, attention, up], skip=0)

    def superedge_update(self, supernodes, superedges, super_graph, super_edge_weights):
        S = supernodes.shape[0]
        return _edge_update(self.superedge_network, supernodes, superedges, _plans(super_graph, S, S))

    def forward(self, nodes, edges, supernodes, superedges, graph, bipartite_graph, bipartite_edge_weights,
                super_graph, super_edge_weights, skip_edge_updates=False):
        N, S = nodes.shape[0], supernodes.shape[0]
        gp, bp, sp = _plans(graph, N, N), _plans(bipartite_graph, N, S), _plans(super_graph, S, S)
        supernodes = self.supernode_update(nodes, supernodes, superedges, bp, bipartite_edge_weights, sp, super_edge_weights)
        nodes = self.node_update(nodes, edges, supernodes, gp, bp, bipartite_edge_weights)
        if not skip_edge_updates:
            superedges = self.superedge_update(supernodes, superedges, sp, super_edge_weights)
            edges = self.edge_update(nodes, edges, gp)
        return nodes, edges, supernodes, superedges


class DynamicGraphConstruction(nn.Module):
    def __init__(self, weighting_function, hparams):
        super().__init__()
        self.hparams = hparams
        self.weight_normalization = nn.BatchNorm1d(1)
        if weighting_function not in ("sigmoid", "exp"):
            getattr(torch, weighting_function)  # same AttributeError as the reference for unknown names
        self.weighting_function = getattr(torch, weighting_function)
        self.register_buffer("knn_radius", torch.ones(1), persistent=True)

    def build_graph(self, src_embeddings, dst_embeddings, sym, k, src_ptr=None, dst_ptr=None):
        """The no-grad half (gnn_utils.py:193-205): radius-kNN, optional symmetrize,
        radius tracking. Returns graph[2, E'] int64. ``src_ptr`` / ``dst_ptr``: event offsets of a batch of events
        (neighbours are only sought inside the row's own event)."""
        with torch.no_grad():
            idx = find_neighbors(src_embeddings, dst_embeddings, r_max=self.knn_radius, k_max=k, ptr1=src_ptr, ptr2=dst_ptr)
            graph = ops.knn_edges(idx)
            if sym:
                graph = ops.symmetrize(graph, max(src_embeddings.shape[0], dst_embeddings.shape[0]))
            if self.training and graph.shape[1] > 0:
                dmax = ops.edge_max_dist(src_embeddings, dst_embeddings, graph)
                self.knn_radius = 0.9 * self.knn_radius + 0.11 * dmax
        return graph

    def forward(self, src_embeddings, dst_embeddings, sym=False, norm=False, k=10, logits=False, graph=None,
                src_ptr=None, dst_ptr=None, src_event=None):
        """``src_ptr`` / ``dst_ptr`` / ``src_event`` (event offsets of the source and destination rows, event id of every
        source row) describe a batch of events: the graph is built event by event and ``norm`` divides by the mean weight
        of the edge's own event, as a loop over single events would (the reference sees one event per call)."""
        if graph is None:
            graph = self.build_graph(src_embeddings, dst_embeddings, sym, k, src_ptr, dst_ptr)
        gp = GraphPlans(graph, src_embeddings.shape[0], dst_embeddings.shape[0])
        likelihood = ops.edge_dot(src_embeddings, dst_embeddings, gp.by_src, gp.by_dst)
        edge_weights_logits = self.weight_normalization(likelihood.unsqueeze(1)).squeeze()
        edge_weights = self.weighting_function(edge_weights_logits)
        if norm and src_event is not None and graph.shape[1] > 0:
            edge_weights = ops.segment_mean_normalize(edge_weights.reshape(-1), src_event[graph[0]], src_ptr.numel() - 1)
        elif norm:
            edge_weights = edge_weights / edge_weights.mean()
        edge_weights = edge_weights.unsqueeze(1)
        if logits:
            return graph, edge_weights, edge_weights_logits
        return graph, edge_weights
