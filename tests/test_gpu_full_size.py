"""BASELINE-size checks of the fused edge step (config 2: L = 128, E = 1e6, N = 1e5) through size-independent
properties, since the oracle cannot run a million edges in seconds:
  * a random sample of edge rows against the fp64 oracle (the edge step is row-independent given the node table),
  * the fused aggregate against an fp64 scatter of the kernel's own e' (ordered-sum accuracy, empty segments = 0),
  * run-to-run bit identity of every output and gradient,
  * exact homogeneity of the backward in the cotangent (scaling by 2 commutes with every rounding in the kernel),
  * the ragged last tile (E is not a multiple of 128)."""
import pytest
import torch

from oracle import hgnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _problem(E, L, seed):
    from hierarchicalgnn_b200.gnn_utils import GraphPlans, InteractionGNNCell
    from hierarchicalgnn_b200.synth import synth_edge_problem
    from hierarchicalgnn_b200.training_utils import kaiming_init
    hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
    torch.manual_seed(0)
    cell = InteractionGNNCell(hp)
    kaiming_init(cell)
    nodes, edges, graph = synth_edge_problem(E, L, seed=seed)
    order = torch.argsort(graph[1], stable=True)
    graph, edges = graph[:, order].contiguous(), edges[order].contiguous()
    return hp, cell, nodes, edges, graph


def test_edge_step_at_baseline_size_properties():
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200.gnn_utils import GraphPlans
    L, E = 128, 1_000_000
    hp, cell, nodes, edges, graph = _problem(E, L, seed=42)
    N = nodes.shape[0]
    assert E % 128 != 0  # the last tile is ragged
    sd = {"edge_network." + k: v.detach().clone() for k, v in cell.edge_network.state_dict().items()}
    cell.to(DEV)
    net = cell.edge_network
    params = list(net.parameters())
    nd, ed, gd = nodes.to(DEV).requires_grad_(True), edges.to(DEV).requires_grad_(True), graph.to(DEV)
    gp = GraphPlans(gd, N, N, dst_sorted=True)
    g = torch.Generator().manual_seed(5)
    cot_e, cot_a = torch.randn(E, L, generator=g).to(DEV), torch.randn(N, L, generator=g).to(DEV)

    def run(scale):
        e2, agg = net.edge_step(nd, ed, gp.by_src, gp.by_dst)
        assert agg is not None, "tensor-core edge step with fused aggregate was not taken"
        grads = torch.autograd.grad([e2, agg], [nd, ed] + params, [cot_e * scale, cot_a * scale])
        return [e2.detach(), agg.detach()] + [t.detach() for t in grads]

    old = ops.set_precision("auto")
    try:
        n0 = ops.TC_CALLS["count"]
        a = run(1.0)
        assert ops.TC_CALLS["count"] - n0 == 2
        b = run(1.0)
        c = run(2.0)
    finally:
        ops.set_precision(old)
    for x, y in zip(a, b):
        assert torch.equal(x, y)  # ordered reductions only: bit-identical run to run
    for x, y in zip(a[2:], c[2:]):
        assert torch.equal(2.0 * x, y)  # backward is exactly homogeneous in the cotangent
    e2, agg = a[0], a[1]
    assert bool(torch.isfinite(e2).all()) and all(bool(torch.isfinite(t).all()) for t in a[2:])
    # sampled rows vs the fp64 oracle (bf16 tolerance on O(1) latents), including the ragged tail
    idx = torch.cat([torch.randint(0, E, (3000,), generator=g), torch.arange(E - 200, E)])
    sub_graph = graph[:, idx]
    want = O.edge_step(O.cast_state(sd, torch.float64), "edge_network", hp, nodes.double(), edges[idx].double(), sub_graph)
    got = e2[idx.to(DEV)].cpu().double()
    assert float((got - want).abs().max()) < 2e-2
    assert float((got - want).abs().mean()) < 2e-3
    # fused aggregate = ordered sum of the kernel's own rows
    ref = torch.zeros(N, L, dtype=torch.float64, device=DEV).index_add_(0, gd[1], e2.double())
    deg = torch.bincount(gd[1], minlength=N)
    assert float((agg.double() - ref).abs().max()) < 1e-5 * float(deg.max()) ** 0.5 + 1e-5
    assert bool((agg[deg == 0] == 0).all())
