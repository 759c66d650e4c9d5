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


def test_edge_step_at_baseline_size_all_gradients_vs_fp64_reference():
    """E = 1e6, L = 128: e', the fused aggregate and EVERY gradient (d_x through the node-level adjoint, d_e, dW1 / dW2 from
    the split-K over 7 813 tiles, biases, LayerNorm affine) against the oracle evaluated in fp64 on the GPU. Tolerance:
    relative Frobenius 1.5e-2 (bf16 MMA operands), rms 6e-3 / max-abs 4e-2 on the O(1) outputs."""
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200.gnn_utils import GraphPlans
    L, E = 128, 1_000_000
    hp, cell, nodes, edges, graph = _problem(E, L, seed=43)
    N = nodes.shape[0]
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():  # non-trivial biases / LayerNorm affine so that their gradients carry signal
        for k, p in cell.edge_network.named_parameters():
            if p.dim() == 1:
                p.add_(0.2 * torch.randn(p.shape, generator=g))
    sd = {"edge_network." + k: v.detach().clone() for k, v in cell.edge_network.state_dict().items()}
    cell.to(DEV)
    net = cell.edge_network
    names = [k for k, _ in net.named_parameters()]
    params = [p for _, p in net.named_parameters()]
    nd, ed, gd = nodes.to(DEV).requires_grad_(True), edges.to(DEV).requires_grad_(True), graph.to(DEV)
    gp = GraphPlans(gd, N, N, dst_sorted=True)
    cot_e, cot_a = torch.randn(E, L, generator=g).to(DEV), torch.randn(N, L, generator=g).to(DEV)
    old = ops.set_precision("auto")
    try:
        e2, agg = net.edge_step(nd, ed, gp.by_src, gp.by_dst)
        assert agg is not None
        got = torch.autograd.grad([e2, agg], [nd, ed] + params, [cot_e, cot_a])
    finally:
        ops.set_precision(old)
    sd64 = {k: v.double().to(DEV).requires_grad_(True) for k, v in sd.items()}
    n64, e64 = nodes.double().to(DEV).requires_grad_(True), edges.double().to(DEV).requires_grad_(True)
    e2_o = O.edge_step(sd64, "edge_network", hp, n64, e64, gd)
    agg_o = O.scatter_add(e2_o, gd[1], N)
    want = torch.autograd.grad([e2_o, agg_o], [n64, e64] + [sd64["edge_network." + k] for k in names],
                               [cot_e.double(), cot_a.double()])
    d = e2.detach().double() - e2_o.detach()
    # one latent's error is ~N(0, 5e-3) (bf16 operands through K = 384 / 256): rms bound, and a 6-sigma bound on the maximum over 1.28e8 values
    assert float(d.square().mean().sqrt()) < 6e-3 and float(d.abs().max()) < 4e-2, (float(d.square().mean().sqrt()), float(d.abs().max()))
    deg = float(torch.bincount(gd[1], minlength=N).max())
    assert float((agg.detach().double() - agg_o.detach()).abs().max()) < 4e-2 * deg ** 0.5 + 1e-3

    def rel(a, b):
        return float((a.double() - b).norm() / b.norm())
    errs = {"d_nodes": rel(got[0], want[0]), "d_edges": rel(got[1], want[1])}
    for k, a, b in zip(names, got[2:], want[2:]):
        errs[k] = rel(a, b)
    bad = {k: v for k, v in errs.items() if not v < 1.5e-2}
    assert not bad, errs


@pytest.mark.parametrize("P1,P2,k,sym", [(120_000, 12_000, 5, False), (12_000, 12_000, 10, True)])
def test_knn_bit_exact_at_full_pileup_sizes(P1, P2, k, sym):
    """The two kNN shapes of a full pile-up event (bipartite: 120 000 hits x 12 000 supernodes, k = 5; supergraph:
    12 000^2, k = 10, self matches kept) against an fp64 brute force evaluated on the GPU in query chunks: identical
    neighbour table (ties -> smaller index) wherever the fp64 distances have no fp32-level near-tie at the decision
    boundaries; the seed is required to have (almost) none."""
    from hierarchicalgnn_b200 import ops
    g = torch.Generator().manual_seed(P1 + k)
    centres = torch.nn.functional.normalize(torch.randn(P2, 8, generator=g))
    if sym:
        q = centres.clone()
    else:
        q = torch.nn.functional.normalize(centres[torch.randint(0, P2, (P1,), generator=g)] + 0.3 * torch.randn(P1, 8, generator=g))
    radius = 0.9
    qd, rd = q.to(DEV), centres.to(DEV)
    got = ops.knn_radius(qd, rd, k, radius)
    assert got.shape == (P1, k)
    q64, r64 = qd.double(), rd.double()
    r2 = radius * radius
    n_bad = n_amb = 0
    for s in range(0, P1, 4096):
        qc = q64[s:s + 4096]
        d2 = (qc[:, None, :] - r64[None, :, :]).square().sum(-1)                    # direct sum of squared differences
        srt = torch.sort(d2, dim=1, stable=True)                                    # ties -> smaller index
        top, idx = srt.values[:, :k + 1], srt.indices[:, :k]
        want = torch.where(top[:, :k] < r2, idx, torch.full_like(idx, -1))
        # rows whose answer could flip under fp32 rounding of the distances: a relative gap < 2e-6 (3x the fp32 error bound of an 8-term sum of squares) between consecutive
        # ranks (up to rank k+1) or between a kept distance and the radius
        gaps = (top[:, 1:] - top[:, :-1]) / top[:, 1:].clamp(min=1e-30)
        amb = (gaps < 2e-6).any(1) | (((top[:, :k] - r2).abs() / r2) < 2e-6).any(1)
        if sym:  # the self match at distance exactly 0 is rank 0 by construction: its gap to rank 1 is never a tie
            amb = (gaps[:, 1:] < 2e-6).any(1) | (((top[:, :k] - r2).abs() / r2) < 2e-6).any(1)
        neq = (got[s:s + 4096] != want).any(1)
        n_bad += int((neq & ~amb).sum())
        n_amb += int(amb.sum())
    assert n_bad == 0
    assert n_amb <= P1 // 100  # the certificate covers >= 99 % of the queries
