"""GPU parity of the tensor-core row layer (hgnn_tc_row_forward / hgnn_tc_row_backward) and of the layer-wise
executor that routes the node / supernode networks, encoder tails and classifier hidden layers through it.
Stated tolerances: against an fp64 reference with the SAME bf16 operand rounding emulated, 4e-3 max / 2e-4 mean on
O(1) outputs; against the plain fp64 reference the bf16 tolerance of SURVEY.md §8c (2e-2 on latents); gradients
relative-Frobenius < 1.5e-2."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ACTS = {"GELU": torch.nn.functional.gelu, "Tanh": torch.tanh, "ReLU": torch.relu, "SiLU": torch.nn.functional.silu}


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _layer_case(widths, n_out, rows, n_src, seed, gathered):
    g = torch.Generator().manual_seed(seed)
    K = sum(widths)
    W = torch.randn(n_out, K, generator=g) / K ** 0.5
    b = 0.3 * torch.randn(n_out, generator=g)
    gamma = 1.0 + 0.2 * torch.randn(n_out, generator=g)
    beta = 0.2 * torch.randn(n_out, generator=g)
    segs, idx = [], []
    for s, w in enumerate(widths):
        if gathered[s]:
            segs.append(torch.randn(n_src, w, generator=g))
            idx.append(torch.randint(0, n_src, (rows,), generator=g))
        else:
            segs.append(torch.randn(rows, w, generator=g))
            idx.append(None)
    return W, b, gamma, beta, segs, idx


def _reference(W, b, gamma, beta, segs, idx, act, skip, emulate):
    rows = [t if i is None else t[i] for t, i in zip(segs, idx)]
    a = torch.cat(rows, 1)
    if emulate:
        h = bf16r(a).double() @ bf16r(W).double().t() + b.double()
    else:
        h = a.double() @ W.double().t() + b.double()
    y = ACTS[act](torch.nn.functional.layer_norm(h, (W.shape[0],), gamma.double(), beta.double(), 1e-5))
    return y if skip is None else y + skip.double()


CASES = [
    # widths, n_out, rows, n_src, gathered, act, skip
    ([128, 128], 256, 1000, 0, [False, False], "GELU", False),       # node network layer 1 (IN)
    ([256], 256, 777, 0, [False], "GELU", False),                    # hidden layer
    ([256], 128, 300, 0, [False], "GELU", True),                     # last layer + residual
    ([128, 128, 128], 256, 515, 0, [False, False, False], "GELU", False),   # HGNN node network layer 1
    ([128, 128], 256, 900, 70, [True, True], "Tanh", False),         # classifier layer 1 on gathered rows
    ([128], 128, 5, 0, [False], "ReLU", True),                       # fewer rows than one tile
    ([64, 64], 128, 129, 33, [True, False], "SiLU", False),          # 64-wide segments, one gathered
    ([256], 256, 128 * 150 + 1, 0, [False], "GELU", False),          # more tiles than SMs
]


@pytest.mark.parametrize("widths,n_out,rows,n_src,gathered,act,skip", CASES)
def test_tc_row_layer_forward_and_backward(widths, n_out, rows, n_src, gathered, act, skip):
    from hierarchicalgnn_b200 import ops
    assert ops.tc_row_supported(widths, n_out, act)
    W, b, gamma, beta, segs, idx = _layer_case(widths, n_out, rows, n_src, seed=rows + n_out, gathered=gathered)
    g = torch.Generator().manual_seed(99)
    res = torch.randn(rows, n_out, generator=g) if skip else None
    cot = torch.randn(rows, n_out, generator=g)
    # fp64 reference (plain and with operand rounding emulated)
    leaves = [t.clone().double().requires_grad_(True) for t in segs]
    Wr, br, gr, ber = [t.clone().double().requires_grad_(True) for t in (W, b, gamma, beta)]
    resr = None if res is None else res.clone().double().requires_grad_(True)
    want = _reference(Wr, br, gr, ber, leaves, idx, act, resr, emulate=False)
    (want * cot.double()).sum().backward()
    want_emul = _reference(W, b, gamma, beta, segs, idx, act, res, emulate=True)

    Wd, bd, gd, bed = [t.to(DEV).requires_grad_(True) for t in (W, b, gamma, beta)]
    segs_d = [t.to(DEV).requires_grad_(True) for t in segs]
    res_d = None if res is None else res.to(DEV).requires_grad_(True)
    plans = [None if i is None else ops.plan_for(i.to(DEV), n_src) for i in idx]
    cache = {}

    def pack():
        if "w" not in cache:
            cache["w"] = (ops.tc_pack_weight(Wd), ops.tc_pack_weight_t(Wd))
        return cache["w"]
    meta = ops.RowLayerMeta(plans, act, 1e-5, res is not None, pack)
    n0 = ops.TC_ROW_CALLS["count"]
    got = ops.tc_row_layer(meta, segs_d, res_d, Wd, bd, gd, bed)
    (got * cot.to(DEV)).sum().backward()
    assert ops.TC_ROW_CALLS["count"] == n0 + 2
    out = got.detach().cpu().double()
    assert float((out - want_emul).abs().max()) < 4e-3
    assert float((out - want_emul).abs().mean()) < 2e-4
    assert float((out - want.detach()).abs().max()) < 3e-2

    def rel(x, y):
        return float((x.detach().cpu().double() - y).norm() / y.norm().clamp_min(1e-30))
    # ReLU's derivative is a step: bf16 rounding of the operands flips a few gates near zero, which the smooth
    # activations do not suffer from -> wider gradient tolerance for that case only
    tol = 4e-2 if act == "ReLU" else 1.5e-2
    for t_d, t_r in zip(segs_d, leaves):
        assert rel(t_d.grad, t_r.grad) < tol
    assert rel(Wd.grad, Wr.grad) < tol
    assert rel(bd.grad, br.grad) < tol and rel(gd.grad, gr.grad) < tol and rel(bed.grad, ber.grad) < tol
    if res is not None:
        assert torch.equal(res_d.grad.cpu(), cot)  # residual gradient is the cotangent itself


def test_tc_row_layer_per_segment_gradients_skip_segments_without_grad():
    """hgnn_tc_row_backward_split: 128-wide segments get one dense gradient matrix each; a segment that needs no gradient is
    never stored (NULL piece) and the others are bit-identical to the run where every segment needs one. A 256-wide segment
    spans two of the kernel's 128-column pieces."""
    from hierarchicalgnn_b200 import ops
    for widths, n_out, rows in (([128, 128, 128], 256, 1300), ([256, 128], 128, 700)):
        W, b, gamma, beta, segs, idx = _layer_case(widths, n_out, rows, 0, seed=11, gathered=[False] * len(widths))
        cot = torch.randn(rows, n_out, generator=torch.Generator().manual_seed(2)).to(DEV)
        grads = []
        for frozen in (None, 0, len(widths) - 1):
            Wd, bd, gd, bed = [t.to(DEV).requires_grad_(True) for t in (W, b, gamma, beta)]
            segs_d = [t.to(DEV).requires_grad_(s != frozen) for s, t in enumerate(segs)]
            packed = (ops.tc_pack_weight(Wd), ops.tc_pack_weight_t(Wd))
            meta = ops.RowLayerMeta([None] * len(widths), "GELU", 1e-5, False, lambda: packed)
            y = ops.tc_row_layer(meta, segs_d, None, Wd, bd, gd, bed)
            (y * cot).sum().backward()
            for t in segs_d:
                assert t.grad is None or (t.grad.is_contiguous() and t.grad.shape == t.shape)
            grads.append([t.grad for t in segs_d] + [Wd.grad, bd.grad, gd.grad, bed.grad])
        full = grads[0]
        # fp64 reference of the input gradients (relative Frobenius, bf16 operands)
        leaves = [t.clone().double().requires_grad_(True) for t in segs]
        want = _reference(W.double(), b.double(), gamma.double(), beta.double(), leaves, idx, "GELU", None, emulate=False)
        (want * cot.cpu().double()).sum().backward()
        for got, ref in zip(full, leaves):
            assert float((got.cpu().double() - ref.grad).norm() / ref.grad.norm()) < 1.5e-2
        for part, frozen in ((grads[1], 0), (grads[2], len(widths) - 1)):
            for s, (a, b_) in enumerate(zip(part, full)):
                if s == frozen:
                    assert a is None
                else:
                    assert torch.equal(a, b_)


def test_tc_row_layer_is_deterministic():
    from hierarchicalgnn_b200 import ops
    widths, n_out, rows = [128, 128], 256, 5000
    W, b, gamma, beta, segs, idx = _layer_case(widths, n_out, rows, 0, seed=3, gathered=[False, False])
    cot = torch.randn(rows, n_out, generator=torch.Generator().manual_seed(1)).to(DEV)
    outs = []
    for _ in range(2):
        Wd, bd, gd, bed = [t.to(DEV).requires_grad_(True) for t in (W, b, gamma, beta)]
        segs_d = [t.to(DEV).requires_grad_(True) for t in segs]
        packed = (ops.tc_pack_weight(Wd), ops.tc_pack_weight_t(Wd))
        meta = ops.RowLayerMeta([None, None], "GELU", 1e-5, False, lambda: packed)
        y = ops.tc_row_layer(meta, segs_d, None, Wd, bd, gd, bed)
        (y * cot).sum().backward()
        outs.append([y.detach(), Wd.grad, bd.grad, gd.grad, bed.grad, segs_d[0].grad, segs_d[1].grad])
    for a, b_ in zip(*outs):
        assert torch.equal(a, b_)  # ordered reductions everywhere: bit-identical run to run


def test_node_network_routes_through_row_layers_and_matches_fp32_path():
    """InteractionGNNCell.node_update (gnn_utils.py:45-54) at latent 128: the three layers run as tensor-core row
    layers on the default path; outputs and all gradients agree with the fp32 SIMT path within bf16 tolerance."""
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200.utils import make_mlp
    from hierarchicalgnn_b200.training_utils import kaiming_init
    L, N = 128, 3000
    torch.manual_seed(0)
    net = make_mlp(2 * L, 2 * L, L, 3, layer_norm=True, output_activation="GELU", hidden_activation="GELU")
    kaiming_init(net)
    net.to(DEV)
    g = torch.Generator().manual_seed(4)
    x, m, cot = (torch.randn(N, L, generator=g).to(DEV) for _ in range(3))
    res = {}
    for mode in ("fp32", "auto"):
        old = ops.set_precision(mode)
        try:
            net.zero_grad(set_to_none=True)
            xd, md = x.clone().requires_grad_(True), m.clone().requires_grad_(True)
            n0 = ops.TC_ROW_CALLS["count"]
            y = net.fused([xd, md], skip=0)
            (y * cot).sum().backward()
            used = ops.TC_ROW_CALLS["count"] - n0
            res[mode] = [y.detach(), xd.grad, md.grad] + [p.grad.clone() for p in net.parameters()]
        finally:
            ops.set_precision(old)
        assert used == (6 if mode == "auto" else 0)

    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-30))
    assert float((res["auto"][0] - res["fp32"][0]).abs().max()) < 3e-2
    for a, b in zip(res["auto"][1:], res["fp32"][1:]):
        assert rel(a, b) < 2e-2


SPLIT_CASES = [
    # widths, n_out, rows, n_src, gathered, act, skip     (latent 256 / 64 layer shapes)
    ([256, 256, 256], 512, 700, 60, [True, True, False], "GELU", False),   # edge network layer 1 at latent 256
    ([512], 256, 1000, 0, [False], "Tanh", True),                          # edge network layer 2 at latent 256 (+ residual)
    ([256, 256], 512, 300, 0, [False, False], "GELU", False),              # node network layer 1 at latent 256
    ([64, 64, 64], 128, 900, 80, [True, True, False], "GELU", False),      # edge network layer 1 at latent 64
    ([128], 64, 515, 0, [False], "Tanh", True),                            # edge network layer 2 at latent 64 (+ residual)
    ([128], 64, 3, 0, [False], "GELU", False),
    ([256, 256, 256], 512, 128 * 150 + 7, 0, [False, False, False], "GELU", False),  # more tiles than SMs
]


@pytest.mark.parametrize("widths,n_out,rows,n_src,gathered,act,skip", SPLIT_CASES)
def test_tc_split_layer_forward_and_backward(widths, n_out, rows, n_src, gathered, act, skip):
    """Generic tensor-core layer (plain tcgen05 GEMMs + row-wise LayerNorm kernels) for the latent 64 / 256 shapes."""
    from hierarchicalgnn_b200 import ops
    assert ops.tc_split_supported(tuple(widths), n_out, act) and not ops.tc_row_supported(widths, n_out, act)
    W, b, gamma, beta, segs, idx = _layer_case(widths, n_out, rows, n_src, seed=rows + n_out + 1, gathered=gathered)
    g = torch.Generator().manual_seed(98)
    res = torch.randn(rows, n_out, generator=g) if skip else None
    cot = torch.randn(rows, n_out, generator=g)
    leaves = [t.clone().double().requires_grad_(True) for t in segs]
    Wr, br, gr, ber = [t.clone().double().requires_grad_(True) for t in (W, b, gamma, beta)]
    resr = None if res is None else res.clone().double().requires_grad_(True)
    want = _reference(Wr, br, gr, ber, leaves, idx, act, resr, emulate=False)
    (want * cot.double()).sum().backward()
    want_emul = _reference(W, b, gamma, beta, segs, idx, act, res, emulate=True)

    Wd, bd, gd, bed = [t.to(DEV).requires_grad_(True) for t in (W, b, gamma, beta)]
    segs_d = [t.to(DEV).requires_grad_(True) for t in segs]
    res_d = None if res is None else res.to(DEV).requires_grad_(True)
    plans = [None if i is None else ops.plan_for(i.to(DEV), n_src) for i in idx]
    packed = ops.tc_pack_split(Wd)
    meta = ops.RowLayerMeta(plans, act, 1e-5, res is not None, lambda: packed)
    outs = []
    for _ in range(2):
        for t in segs_d + [Wd, bd, gd, bed] + ([res_d] if res_d is not None else []):
            t.grad = None
        got = ops.tc_split_layer(meta, segs_d, res_d, Wd, bd, gd, bed)
        (got * cot.to(DEV)).sum().backward()
        outs.append([got.detach().clone(), Wd.grad.clone(), bd.grad.clone(), gd.grad.clone(), bed.grad.clone()] + [t.grad.clone() for t in segs_d])
    for x, y in zip(*outs):
        assert torch.equal(x, y)  # bit-identical run to run
    out = outs[0][0].cpu().double()
    assert float((out - want_emul).abs().max()) < 4e-3
    assert float((out - want_emul).abs().mean()) < 2e-4
    assert float((out - want.detach()).abs().max()) < 3e-2

    def rel(x, y):
        return float((x.detach().cpu().double() - y).norm() / y.norm().clamp_min(1e-30))
    for t_d, t_r in zip(segs_d, leaves):
        assert rel(t_d.grad, t_r.grad) < 1.5e-2
    assert rel(Wd.grad, Wr.grad) < 1.5e-2
    assert rel(bd.grad, br.grad) < 1.5e-2 and rel(gd.grad, gr.grad) < 1.5e-2 and rel(bed.grad, ber.grad) < 1.5e-2
    if res is not None:
        assert torch.equal(res_d.grad.cpu(), cot)


@pytest.mark.parametrize("L", [64, 256])
def test_interaction_cell_other_latents_route_through_tensor_cores(L):
    """InteractionGNNCell at latent 64 / 256 (the BC config's default latent is 256): every layer of both networks runs
    on tensor cores on the default path (row layers where they apply, generic GEMM + LayerNorm layers elsewhere) and the
    cell agrees with the fp32 path within the bf16 tolerance, outputs and gradients."""
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200.gnn_utils import InteractionGNNCell
    from hierarchicalgnn_b200.synth import synth_edge_problem
    from hierarchicalgnn_b200.training_utils import kaiming_init
    hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
    torch.manual_seed(0)
    cell = InteractionGNNCell(hp)
    kaiming_init(cell)
    cell.to(DEV)
    nodes, edges, graph = synth_edge_problem(3000, L, seed=3)
    g = torch.Generator().manual_seed(2)
    cn, ce = torch.randn(nodes.shape, generator=g).to(DEV), torch.randn(edges.shape, generator=g).to(DEV)
    res = {}
    for mode in ("fp32", "auto"):
        old = ops.set_precision(mode)
        try:
            cell.zero_grad(set_to_none=True)
            nd, ed = nodes.to(DEV).requires_grad_(True), edges.to(DEV).requires_grad_(True)
            l0 = ops.TC_ROW_CALLS["count"]
            a, b = cell(nd, ed, graph.to(DEV))
            ((a * cn).sum() + (b * ce).sum()).backward()
            used = ops.TC_ROW_CALLS["count"] - l0
            res[mode] = [a.detach(), b.detach(), nd.grad, ed.grad] + [p.grad.clone() for p in cell.parameters()]
        finally:
            ops.set_precision(old)
        assert used == (10 if mode == "auto" else 0)  # 5 layers, forward + backward each

    def rel(x, y):
        return float((x - y).norm() / y.norm().clamp_min(1e-30))
    assert float((res["auto"][0] - res["fp32"][0]).abs().max()) < 3e-2 and float((res["auto"][1] - res["fp32"][1]).abs().max()) < 3e-2
    for x, y in zip(res["auto"][2:], res["fp32"][2:]):
        assert rel(x, y) < 3e-2
