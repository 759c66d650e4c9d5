"""Interaction-network edge classifier (reference: EdgeClassifier/Models/IN.py).

``InteractionGNNBlock`` / ``EC_InteractionGNN`` keep the reference constructor
arguments, forward signature and state-dict layout; every tensor op on the path
is a kernel from libhgnn_b200.so. The graph's segment plans are built once per
forward and shared by all cells.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ...gnn_utils import GraphPlans, InteractionGNNCell, sort_edges_by_destination
from ...utils import make_mlp
from ..edge_classifier_base import EdgeClassifierBase


class InteractionGNNBlock(nn.Module):
    def __init__(self, hparams, iterations, emb=True):
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
        if emb:
            self.output_layer = make_mlp(L, H, hparams["emb_dim"], hparams["output_layers"], layer_norm=ln,
                                         output_activation=None, hidden_activation=hparams["hidden_output_activation"])
        self.emb = emb
        self.hparams = hparams

    def forward(self, x, graph):
        gp = graph if isinstance(graph, GraphPlans) else GraphPlans(graph, x.shape[0], x.shape[0])
        if torch.is_grad_enabled() and x.is_leaf:
            x.requires_grad = True  # kept from the reference (it needed it for reentrant checkpointing)
        nodes = self.node_encoder(x)
        edges = self.edge_encoder.fused([x, x], [gp.by_src, gp.by_dst])
        for cell in self.ignn_cells:
            nodes, edges = cell(nodes, edges, gp)
        if self.emb:
            embeddings = nn.functional.normalize(self.output_layer(nodes))
            return embeddings, nodes, edges
        return nodes, edges


class EC_InteractionGNN(EdgeClassifierBase):
    def __init__(self, hparams):
        super().__init__(hparams)
        self.ignn_block = InteractionGNNBlock(hparams, hparams["n_interaction_graph_iters"], emb=False)
        self.edge_classifier = make_mlp(2 * hparams["latent"], hparams["hidden"], 1, hparams["output_layers"],
                                        layer_norm=hparams["layernorm"], output_activation=None,
                                        hidden_activation=hparams["hidden_output_activation"])

    def forward(self, x, graph):
        E = graph.shape[1]
        directed = torch.cat([graph, graph.flip(0)], dim=1)          # edge k and k+E are mutual reverses (IN.py:122)
        directed, _, where = sort_edges_by_destination(directed)     # one sort per event; cells stream rows in place
        nodes, edges = self.ignn_block(x, GraphPlans(directed, x.shape[0], x.shape[0], dst_sorted=True))
        # classifier input = [e_forward | e_reverse] (IN.py:126): rows where[k] and where[k+E] of the sorted edge
        # latents, gathered inside the fused MLP kernel — no concat, no un-sort pass
        fwd = ops.plan_for(where[:E].contiguous(), 2 * E)
        rev = ops.plan_for(where[E:].contiguous(), 2 * E)
        scores = self.edge_classifier.fused([edges, edges], [fwd, rev]).squeeze()
        return torch.sigmoid(scores)
