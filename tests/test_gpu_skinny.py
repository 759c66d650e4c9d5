"""GPU parity of the skinny fp32 layers (hgnn_narrow_in_* / hgnn_narrow_out_*): the encoders' first Linear (fan-in 3 / 6)
and the heads' last Linear (fan-out 1 / 8). fp32 arithmetic -> fp32 tolerances vs an fp64 reference: 2e-5 on outputs,
relative-Frobenius 1e-5 on gradients; run-to-run bit identity."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ACTS = {"GELU": torch.nn.functional.gelu, "Tanh": torch.tanh, "ReLU": torch.relu, None: lambda t: t}


def rel(x, y):
    return float((x.detach().cpu().double() - y).norm() / y.norm().clamp_min(1e-30))


@pytest.mark.parametrize("widths,gathered,n_out,rows,n_src,ln,act", [
    ([3], [False], 256, 5000, 0, True, "GELU"),          # node encoder layer 1
    ([3, 3], [True, True], 256, 20000, 700, True, "GELU"),  # edge encoder layer 1 on [x[src] | x[dst]]
    ([6], [False], 128, 77, 0, False, "Tanh"),
    ([8], [False], 32, 1, 0, True, None),
    ([2, 1], [True, False], 64, 300, 10, True, "ReLU"),
    ([3], [False], 512, 3000, 0, True, "GELU"),          # node encoder layer 1 of the latent-256 configs (hidden 512)
    ([3, 3], [True, True], 512, 9000, 500, True, "GELU"),  # edge encoder layer 1, hidden 512
    ([5], [False], 512, 130, 0, True, "Tanh"),           # fan-out 512, generic fan-in
])
def test_narrow_in_layer(widths, gathered, n_out, rows, n_src, ln, act):
    from hierarchicalgnn_b200 import ops
    g = torch.Generator().manual_seed(rows + n_out)
    K = sum(widths)
    W, b = torch.randn(n_out, K, generator=g) / K ** 0.5, 0.3 * torch.randn(n_out, generator=g)
    gamma, beta = 1 + 0.2 * torch.randn(n_out, generator=g), 0.2 * torch.randn(n_out, generator=g)
    segs, idx = [], []
    for w, ga in zip(widths, gathered):
        segs.append(torch.randn(n_src if ga else rows, w, generator=g))
        idx.append(torch.randint(0, n_src, (rows,), generator=g) if ga else None)
    cot = torch.randn(rows, n_out, generator=g)
    # fp64 reference
    lv = [t.clone().double().requires_grad_(True) for t in segs]
    pr = [t.clone().double().requires_grad_(True) for t in (W, b, gamma, beta)]
    a = torch.cat([t if i is None else t[i] for t, i in zip(lv, idx)], 1)
    h = a @ pr[0].t() + pr[1]
    if ln:
        h = torch.nn.functional.layer_norm(h, (n_out,), pr[2], pr[3], 1e-5)
    want = ACTS[act](h)
    (want * cot.double()).sum().backward()

    seg_d = [t.to(DEV).requires_grad_(True) for t in segs]
    par_d = [t.to(DEV).requires_grad_(True) for t in ((W, b, gamma, beta) if ln else (W, b))]
    plans = [None if i is None else ops.plan_for(i.to(DEV), n_src) for i in idx]
    assert ops.narrow_in_supported(tuple(widths), n_out)
    meta = ops.MlpMeta(plans, [act], [ln], -1, 1e-5)
    outs = []
    for _ in range(2):
        for t in seg_d + par_d:
            t.grad = None
        got = ops.narrow_in(meta, seg_d, par_d)
        (got * cot.to(DEV)).sum().backward()
        outs.append([got.detach().clone()] + [t.grad.clone() for t in seg_d + par_d])
    for x, y in zip(*outs):
        assert torch.equal(x, y)
    assert float((outs[0][0].cpu().double() - want.detach()).abs().max()) < 2e-5
    for t_d, t_r in zip(seg_d, lv):
        assert rel(t_d.grad, t_r.grad) < 2e-5
    for t_d, t_r in zip(par_d, pr):
        assert rel(t_d.grad, t_r.grad) < 2e-5


@pytest.mark.parametrize("K,n_out,rows", [(256, 1, 60000), (256, 8, 12000), (128, 3, 500), (512, 1, 777), (512, 4, 33), (128, 8, 1)])
def test_narrow_out_layer(K, n_out, rows):
    from hierarchicalgnn_b200 import ops
    assert ops.narrow_out_supported(K, n_out)
    g = torch.Generator().manual_seed(K + n_out + rows)
    a, W, b = torch.randn(rows, K, generator=g), torch.randn(n_out, K, generator=g) / K ** 0.5, torch.randn(n_out, generator=g)
    cot = torch.randn(rows, n_out, generator=g)
    ar, Wr, br = [t.clone().double().requires_grad_(True) for t in (a, W, b)]
    want = ar @ Wr.t() + br
    (want * cot.double()).sum().backward()
    outs = []
    for _ in range(2):
        ad, Wd, bd = [t.to(DEV).requires_grad_(True) for t in (a, W, b)]
        got = ops.narrow_out(ad, Wd, bd)
        (got * cot.to(DEV)).sum().backward()
        outs.append([got.detach(), ad.grad, Wd.grad, bd.grad])
    for x, y in zip(*outs):
        assert torch.equal(x, y)
    got, da, dW, db = outs[0]
    assert float((got.cpu().double() - want.detach()).abs().max()) < 2e-5
    assert rel(da, ar.grad) < 1e-5 and rel(dW, Wr.grad) < 1e-5 and rel(db, br.grad) < 1e-5


def test_encoder_and_head_route_through_skinny_and_row_layers():
    """node encoder 3 -> H -> H -> L (EC/Models/IN.py:29-33) = narrow-in + 2 tensor-core row layers; edge classifier
    2L -> H -> H -> 1 (IN.py:44-48) = 2 row layers + narrow-out. Both agree with the fp32 path within bf16 tolerance."""
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200.utils import make_mlp
    from hierarchicalgnn_b200.training_utils import kaiming_init
    L, H, N = 128, 256, 4000
    torch.manual_seed(0)
    enc = make_mlp(3, H, L, 3, layer_norm=True, output_activation="GELU", hidden_activation="GELU")
    head = make_mlp(2 * L, H, 1, 3, layer_norm=True, output_activation=None, hidden_activation="GELU")
    kaiming_init(enc); kaiming_init(head)
    enc.to(DEV); head.to(DEV)
    assert [g[0] for g in enc._row_groups([torch.empty(1, 3, device=DEV)], [None], -1)] == ["nin", "tc", "tc"]
    assert [g[0] for g in head._row_groups([torch.empty(1, 2 * L, device=DEV)], [None], -1)] == ["tc", "tc", "nout"]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, 3, generator=g).to(DEV)
    z = torch.randn(N, 2 * L, generator=g).to(DEV)
    res = {}
    for mode in ("fp32", "auto"):
        old = ops.set_precision(mode)
        try:
            enc.zero_grad(set_to_none=True); head.zero_grad(set_to_none=True)
            xd, zd = x.clone().requires_grad_(True), z.clone().requires_grad_(True)
            y1, y2 = enc(xd), head(zd)
            (y1.sum() + y2.sum()).backward()
            res[mode] = [y1.detach(), y2.detach(), xd.grad, zd.grad] + [p.grad.clone() for p in list(enc.parameters()) + list(head.parameters())]
        finally:
            ops.set_precision(old)
    assert float((res["auto"][0] - res["fp32"][0]).abs().max()) < 3e-2
    assert float((res["auto"][1] - res["fp32"][1]).abs().max()) < 3e-2
    for a, b in zip(res["auto"][2:], res["fp32"][2:]):
        assert float((a - b).norm() / b.norm().clamp_min(1e-30)) < 3e-2
