"""GPU parity: every C-ABI operator (called through hierarchicalgnn_b200.ops)
against the CPU oracle on the same seeded inputs. Bit-exact for index work;
fp32 tolerance (stated per test) for floating point."""
import pytest
import torch

from oracle import hgnn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fp32_mode():
    """These tests state fp32 tolerances: pin the SIMT fp32 path (tensor-core parity lives in test_gpu_tc.py)."""
    from hierarchicalgnn_b200 import ops as _o
    old = _o.set_precision("fp32")
    yield
    _o.set_precision(old)

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from hierarchicalgnn_b200 import ops as _ops
    return _ops


def _rand_index(n_items, n_seg, g, hub=True):
    idx = torch.randint(0, max(n_seg - 3, 1), (n_items,), generator=g)  # last segments stay empty
    if hub and n_items > 40:
        idx[5:40] = 1
    return idx


@pytest.mark.parametrize("n_items,n_seg,width", [(0, 5, 16), (1, 1, 4), (1000, 37, 128), (5000, 700, 32), (777, 50, 8),
                                                   (300, 20, 5), (4096, 4096, 256)])
def test_scatter_add_matches_oracle_and_is_deterministic(ops, n_items, n_seg, width):
    g = torch.Generator().manual_seed(n_items + width)
    src = torch.randn(n_items, width, generator=g)
    idx = _rand_index(n_items, n_seg, g)
    want = O.scatter_add(src.double(), idx, n_seg)
    s = src.to(DEV).requires_grad_(True)
    got = ops.scatter_add(s, idx.to(DEV), dim_size=n_seg)
    # tolerance: deg_max * 2^-23 * max|src|
    deg = int(torch.bincount(idx, minlength=1).max()) if n_items else 1
    tol = max(deg, 1) * 2 ** -23 * 6.0
    assert float((got.detach().cpu().double() - want).abs().max() if n_items else 0.0) <= tol
    again = ops.scatter_add(s, idx.to(DEV), dim_size=n_seg)
    assert torch.equal(got, again)  # ordered sums => run-to-run bit identity
    cot = torch.randn(n_seg, width, generator=g)
    (got * cot.to(DEV)).sum().backward()
    if n_items:
        torch.testing.assert_close(s.grad.cpu(), cot[idx], rtol=0, atol=0)


def test_scatter_mean(ops):
    g = torch.Generator().manual_seed(3)
    src, idx = torch.randn(500, 8, generator=g), _rand_index(500, 40, g)
    s = src.to(DEV).requires_grad_(True)
    got = ops.scatter_mean(s, idx.to(DEV), dim_size=40)
    torch.testing.assert_close(got.detach().cpu(), O.scatter_mean(src, idx, 40), rtol=1e-5, atol=1e-6)
    cot = torch.randn(40, 8, generator=g)
    (got * cot.to(DEV)).sum().backward()
    sr = src.clone().requires_grad_(True)
    (O.scatter_mean(sr, idx, 40) * cot).sum().backward()
    torch.testing.assert_close(s.grad.cpu(), sr.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("width", [8, 32, 128])
def test_weighted_gather_scatter_fwd_bwd(ops, width):
    g = torch.Generator().manual_seed(width)
    n_src, n_seg, n_items = 90, 33, 700
    x = torch.randn(n_src, width, generator=g)
    w = torch.rand(n_items, 1, generator=g)
    gi = torch.randint(0, n_src, (n_items,), generator=g)
    si = _rand_index(n_items, n_seg, g)
    cot = torch.randn(n_seg, width, generator=g)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    want = O.scatter_add(wr * xr[gi], si, n_seg)
    (want * cot).sum().backward()
    xd, wd = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    got = ops.gather_scatter(xd, wd, ops.plan_for(gi.to(DEV), n_src), ops.plan_for(si.to(DEV), n_seg))
    (got * cot.to(DEV)).sum().backward()
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(xd.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(wd.grad.cpu(), wr.grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("width", [5, 8, 64, 128])
def test_edge_dot_fwd_bwd(ops, width):
    g = torch.Generator().manual_seed(width + 1)
    na, nb, n = 60, 45, 500
    a, b = torch.randn(na, width, generator=g), torch.randn(nb, width, generator=g)
    ia, ib = torch.randint(0, na, (n,), generator=g), torch.randint(0, nb, (n,), generator=g)
    cot = torch.randn(n, generator=g)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    want = (ar[ia] * br[ib]).sum(-1)
    (want * cot).sum().backward()
    ad, bd = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    got = ops.edge_dot(ad, bd, ops.plan_for(ia.to(DEV), na), ops.plan_for(ib.to(DEV), nb))
    (got * cot.to(DEV)).sum().backward()
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ad.grad.cpu(), ar.grad, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(bd.grad.cpu(), br.grad, rtol=1e-5, atol=1e-5)


MLP_CASES = [
    # (segment widths, gathered?, hidden, out, layers, hidden_act, out_act, layer_norm, skip, rows)
    ([3], [False], 16, 8, 3, "GELU", "GELU", True, -1, 77),
    ([3, 3], [True, True], 32, 16, 2, "GELU", "GELU", True, -1, 130),
    ([32, 32, 32], [True, True, False], 64, 32, 2, "GELU", "Tanh", True, 2, 301),
    ([32, 32], [False, False], 64, 32, 3, "GELU", "GELU", True, 0, 65),
    ([128, 128, 128], [True, True, False], 256, 128, 2, "GELU", "Tanh", True, 2, 200),
    ([64, 64, 64], [False, False, False], 128, 64, 3, "SiLU", "SiLU", False, 0, 97),
    ([64, 64], [True, True], 128, 1, 3, "Tanh", None, True, -1, 150),
    ([40], [False], 24, 8, 3, "ReLU", None, False, -1, 33),
    ([256, 256, 256], [True, True, False], 512, 256, 2, "GELU", "Tanh", True, 2, 70),
    ([16], [False], 16, 16, 1, "GELU", "Sigmoid", True, 0, 9),
    ([32, 32, 32], [True, True, False], 64, 32, 4, "GELU", "Tanh", True, 2, 0),
]


@pytest.mark.parametrize("case", MLP_CASES, ids=[str(i) for i in range(len(MLP_CASES))])
def test_fused_mlp_fwd_bwd_matches_oracle(ops, case):
    from hierarchicalgnn_b200.utils import make_mlp
    widths, gathered, hidden, out, layers, hact, oact, ln, skip, rows = case
    g = torch.Generator().manual_seed(sum(widths) + rows)
    torch.manual_seed(rows + 1)
    net = make_mlp(sum(widths), hidden, out, layers, hidden_activation=hact, output_activation=oact, layer_norm=ln)
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.1 * torch.randn(p.shape, generator=g))  # non-trivial LN gamma/beta and biases
    sd = {"m." + k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    n_src = 41
    segs, idxs = [], []
    for w, gth in zip(widths, gathered):
        segs.append(torch.randn(n_src if gth else rows, w, generator=g))
        idxs.append(torch.randint(0, n_src, (rows,), generator=g) if gth else None)
    cot = torch.randn(rows, out, generator=g)
    # oracle (fp32 CPU autograd over the restated MLP)
    segs_r = [s.clone().requires_grad_(True) for s in segs]
    inp = torch.cat([s[i] if i is not None else s for s, i in zip(segs_r, idxs)], -1)
    want = O.mlp_apply(sd, "m", inp, layers, hact, oact, ln)
    if skip >= 0:
        want = want + (segs_r[skip][idxs[skip]] if idxs[skip] is not None else segs_r[skip])
    (want * cot).sum().backward()
    # device under test
    net = net.to(DEV)
    segs_d = [s.to(DEV).requires_grad_(True) for s in segs]
    plans = [ops.plan_for(i.to(DEV), n_src) if i is not None else None for i in idxs]
    got = net.fused(segs_d, plans, skip=skip)
    assert got.shape == want.shape
    (got * cot.to(DEV)).sum().backward()
    tol = dict(rtol=2e-4, atol=2e-5)  # fp32 SIMT path; accumulation order differs from MKL
    torch.testing.assert_close(got.detach().cpu(), want.detach(), **tol)
    for s_d, s_r in zip(segs_d, segs_r):
        torch.testing.assert_close(s_d.grad.cpu(), s_r.grad if s_r.grad is not None else torch.zeros_like(s_r), rtol=2e-4, atol=1e-4)
    for k, p in net.named_parameters():
        ref = sd["m." + k].grad
        torch.testing.assert_close(p.grad.cpu(), ref if ref is not None else torch.zeros_like(p.grad.cpu()), rtol=5e-4, atol=2e-4,
                                   msg=lambda m: f"{k}: {m}")


def test_fused_mlp_is_deterministic(ops):
    from hierarchicalgnn_b200.utils import make_mlp
    torch.manual_seed(0)
    net = make_mlp(96, 64, 32, 2, output_activation="Tanh", layer_norm=True).to(DEV)
    x = torch.randn(50, 32, device=DEV)
    e = torch.randn(999, 32, device=DEV, requires_grad=True)
    idx = torch.randint(0, 50, (2, 999), device=DEV)
    outs = []
    for _ in range(2):
        net.zero_grad()
        e.grad = None
        y = net.fused([x, x, e], [ops.plan_for(idx[0], 50), ops.plan_for(idx[1], 50), None], skip=2)
        y.square().sum().backward()
        outs.append((y.detach().clone(), e.grad.clone(), net[0].weight.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("nq,nr,dim,k,radius", [(150, 23, 8, 5, 0.9), (400, 400, 8, 10, 0.6), (64, 7, 8, 10, 10.0),
                                                  (300, 90, 3, 4, 0.5), (100, 50, 16, 32, 2.0), (5, 0, 8, 3, 1.0)])
def test_knn_radius_bit_exact_after_canonical_sort(ops, nq, nr, dim, k, radius):
    for seed in range(nq * 31 + nr, nq * 31 + nr + 200):
        g = torch.Generator().manual_seed(seed)
        q = torch.nn.functional.normalize(torch.randn(nq, dim, generator=g)) * 0.8
        r = q.clone() if nq == nr else torch.nn.functional.normalize(torch.randn(nr, dim, generator=g)) * 0.8
        # pin a seed certified to have no fp32 near-ties (rank gaps and radius gap)
        if not nr or O.knn_margin(q, r, k, radius) > 2e-5:
            break
    else:
        pytest.fail("no tie-free seed found")
    want = O.knn_radius(q, r, k, radius)
    got = ops.knn_radius(q.to(DEV), r.to(DEV), k, radius).cpu()
    assert torch.equal(got, want)
    edges = ops.knn_edges(got.to(DEV)).cpu()
    rows = torch.arange(nq).unsqueeze(1).expand_as(want)
    ok = want >= 0
    assert torch.equal(edges, torch.stack([rows[ok], want[ok]], 0))
    if nq == nr and nr:
        sym = ops.symmetrize(edges.to(DEV), nq).cpu()
        assert torch.equal(sym, O.symmetrize(edges))


def test_knn_tie_rule_smaller_index_first(ops):
    q = torch.tensor([[0.0, 0.0]])
    r = torch.tensor([[0.0, 0.5], [1.0, 0.0], [0.0, -0.5], [0.5, 0.0], [3.0, 0.0]])
    got = ops.knn_radius(q.to(DEV), r.to(DEV), 4, 2.0).cpu()
    assert got.tolist() == [[0, 2, 3, 1]]
    got = ops.knn_radius(q.to(DEV), r.to(DEV), 2, 2.0).cpu()
    assert got.tolist() == [[0, 2]]
    got = ops.knn_radius(q.to(DEV), r.to(DEV), 3, 0.75).cpu()
    assert got.tolist() == [[0, 2, 3]]
    got = ops.knn_radius(q.to(DEV), r.to(DEV), 3, 0.5).cpu()  # strict '<' radius
    assert got.tolist() == [[-1, -1, -1]]
    # a closer point arriving AFTER a tied pair pushes both down without reversing them
    r2 = torch.tensor([[0.0, 0.5], [0.0, -0.5], [0.0, 0.1], [0.5, 0.0]])
    got = ops.knn_radius(q.to(DEV), r2.to(DEV), 4, 2.0).cpu()
    assert got.tolist() == [[2, 0, 1, 3]]


@pytest.mark.parametrize("nq,nr,dim,k", [(1200, 1200, 8, 10), (300, 1000, 8, 7), (77, 5000, 8, 5), (130, 300, 5, 12), (1, 129, 8, 3)])
def test_knn_reference_split_path_is_bit_identical_to_the_single_scan(ops, nq, nr, dim, k):
    """Small problems are split over contiguous reference ranges (grid = query blocks x splits) and the per-split sorted
    lists merged; the result must equal the unsplit scan bit for bit, including ties (duplicated reference rows: the
    smaller index first) and rows with fewer than k neighbours inside the radius."""
    from hierarchicalgnn_b200 import _lib
    g = torch.Generator().manual_seed(nq + nr)
    q = torch.randn(nq, dim, generator=g)
    r = torch.randn(nr, dim, generator=g)
    r[nr // 2:nr // 2 + nr // 8] = r[:nr // 8]          # exact duplicates in different reference ranges: distance ties
    radius = 1.9
    L = _lib.lib()
    assert L.hgnn_knn_radius_workspace_bytes(nq, nr, k) > 0, "shape does not take the split path"
    qd, rd = q.to(DEV), r.to(DEV)
    one = torch.empty((nq, k), dtype=torch.int64, device=DEV)
    assert L.hgnn_knn_radius(qd.data_ptr(), nq, rd.data_ptr(), nr, dim, k, radius, one.data_ptr(), None) == 0
    split = ops.knn_radius(qd, rd, k, radius)
    torch.cuda.synchronize()
    assert torch.equal(split, one)
    if nq >= 50:
        assert bool((split < 0).any()) and bool((split >= 0).any())  # padded rows and real neighbours both occur
    want = O.knn_radius(q, r, k, radius)
    agree = (split.cpu() == want).float().mean()
    assert agree > 0.999  # fp64 oracle vs fp32 sums: only near-ties may swap


def test_out_of_range_segment_ids_raise_when_checking_is_on(ops, monkeypatch):
    """torch_scatter raises on an out-of-range index; the CSR build clamps it (no host sync by default).
    HGNN_CHECK_INDICES=1 restores the check at one sync per plan; zero segments with items is always an error."""
    from hierarchicalgnn_b200 import _lib
    idx = torch.tensor([0, 3, 7, 2], device=DEV)
    monkeypatch.setenv("HGNN_CHECK_INDICES", "1")
    with pytest.raises(_lib.HgnnError):
        ops.SegmentPlan(idx, 5)
    with pytest.raises(_lib.HgnnError):
        ops.SegmentPlan(torch.tensor([0, -1], device=DEV), 5)
    ops.SegmentPlan(idx, 8)
    monkeypatch.setenv("HGNN_CHECK_INDICES", "0")
    with pytest.raises(_lib.HgnnError):
        ops.SegmentPlan(idx, 0)


def test_edge_max_dist(ops):
    g = torch.Generator().manual_seed(9)
    a, b = torch.randn(40, 8, generator=g), torch.randn(30, 8, generator=g)
    graph = torch.stack([torch.randint(0, 40, (200,), generator=g), torch.randint(0, 30, (200,), generator=g)])
    want = (a[graph[0]] - b[graph[1]]).square().sum(-1).sqrt().max()
    got = ops.edge_max_dist(a.to(DEV), b.to(DEV), graph.to(DEV)).cpu()
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n,e,with_keep", [(50, 30, False), (2000, 1500, True), (300, 2000, False), (10, 0, False)])
def test_connected_components_labels(ops, n, e, with_keep):
    g = torch.Generator().manual_seed(n + e)
    graph = torch.randint(0, n, (2, e), generator=g)
    keep = (torch.rand(e, generator=g) < 0.6) if with_keep else None
    kept = graph[:, keep] if keep is not None else graph
    want = O.connected_component_labels(kept, n) if e else torch.full((n,), -1, dtype=torch.long)
    got = ops.connected_components(graph.to(DEV), n, None if keep is None else keep.to(DEV)).cpu().long()
    assert torch.equal(got, want)


def test_gmm1d_separates_two_modes(ops):
    g = torch.Generator().manual_seed(4)
    x = torch.cat([0.2 + 0.3 * torch.randn(7000, generator=g), 2.5 + 0.5 * torch.randn(3000, generator=g)])
    p = ops.gmm1d_fit(x.to(DEV)).cpu()
    from sklearn.mixture import GaussianMixture
    sk = GaussianMixture(2, random_state=0).fit(x.numpy().reshape(-1, 1))
    lo, hi = (0, 3) if p[1] < p[4] else (3, 0)
    order = sk.means_.ravel().argsort()
    # EM fixed point is the same; different init/stop => loose tolerance
    assert abs(float(p[lo + 1]) - sk.means_.ravel()[order[0]]) < 2e-2
    assert abs(float(p[hi + 1]) - sk.means_.ravel()[order[1]]) < 2e-2
    assert abs(float(p[lo]) - sk.weights_[order[0]]) < 2e-2
    assert abs(float(p[lo + 2]) - sk.covariances_.ravel()[order[0]]) < 2e-2


def test_cpu_tensors_fail_loudly(ops):
    from hierarchicalgnn_b200._lib import HgnnError
    with pytest.raises(HgnnError):
        ops.scatter_add(torch.randn(4, 4), torch.tensor([0, 1, 1, 0]), dim_size=2)


def test_bad_arguments_surface_last_error(ops):
    from hierarchicalgnn_b200._lib import HgnnError
    with pytest.raises(HgnnError, match="dim must be in"):
        ops.knn_radius(torch.randn(4, 40, device=DEV), torch.randn(4, 40, device=DEV), 3, 1.0)
    with pytest.raises(HgnnError, match="k must be"):
        ops.knn_radius(torch.randn(4, 8, device=DEV), torch.randn(4, 8, device=DEV), 64, 1.0)


@pytest.mark.parametrize("width", [128, 3, 64])
def test_scatter_add_hub_segments(ops, width):
    """Power-law degrees: segments longer than 512 rows take the per-CTA long-segment kernel (rows split into parts summed
    in parallel, combined in part order). Same fp64 tolerance as ordinary segments, bit-identical run to run, means too."""
    g = torch.Generator().manual_seed(7)
    n_items, n_seg = 40_000, 300
    u = torch.rand(n_items, generator=g)
    idx = (n_seg * u.pow(4.0)).long().clamp_(max=n_seg - 1)      # segment 0 holds ~ n_items * 300^-0.25 = 9.6k rows
    deg = torch.bincount(idx, minlength=n_seg)
    assert int(deg.max()) > 5000 and int((deg > 512).sum()) >= 3 and int((deg == 0).sum()) >= 0
    src = torch.randn(n_items, width, generator=g)
    want = torch.zeros(n_seg, width, dtype=torch.float64).index_add_(0, idx, src.double())
    a = ops.scatter_add(src.to(DEV), idx.to(DEV), dim_size=n_seg)
    b = ops.scatter_add(src.to(DEV), idx.to(DEV), dim_size=n_seg)
    assert torch.equal(a, b)
    assert float((a.cpu().double() - want).abs().max()) < 2e-5 * float(deg.max()) ** 0.5
    m = ops.scatter_mean(src.to(DEV), idx.to(DEV), dim_size=n_seg)
    assert float((m.cpu().double() - want / deg.clamp(min=1).unsqueeze(1)).abs().max()) < 1e-5
    # weighted gather + reduce through the same kernels, gradient = transposed plan
    w = torch.rand(n_items, 1, generator=g)
    table = torch.randn(500, width, generator=g)
    gi = torch.randint(0, 500, (n_items,), generator=g)
    td = table.to(DEV).requires_grad_(True)
    out = ops.gather_scatter(td, w.to(DEV), ops.plan_for(gi.to(DEV), 500), ops.plan_for(idx.to(DEV), n_seg))
    want2 = torch.zeros(n_seg, width, dtype=torch.float64).index_add_(0, idx, (w.double() * table.double()[gi]))
    assert float((out.detach().cpu().double() - want2).abs().max()) < 2e-5 * float(deg.max()) ** 0.5


def test_segment_mean_normalize_forward_and_adjoint_vs_torch():
    """w / mean(w over its segment) — the per-event form of edge_weights / edge_weights.mean() (gnn_utils.py:213-214) —
    against plain torch autograd in fp64, including an empty segment and a one-element segment."""
    from hierarchicalgnn_b200 import ops
    g = torch.Generator().manual_seed(8)
    seg = torch.cat([torch.zeros(700), torch.full((1,), 2.0), torch.full((1300,), 3.0), torch.full((40,), 5.0)]).long()
    seg = seg[torch.randperm(seg.numel(), generator=g)]
    w = (0.2 + torch.rand(seg.numel(), generator=g))
    cot = torch.randn(seg.numel(), generator=g)
    wr = w.double().requires_grad_(True)
    sums = torch.zeros(6, dtype=torch.float64).index_add(0, seg, wr)
    cnt = torch.bincount(seg, minlength=6).clamp(min=1)
    want = wr / (sums / cnt)[seg]
    (want * cot.double()).sum().backward()
    wd = w.cuda().requires_grad_(True)
    got = ops.segment_mean_normalize(wd, seg.cuda(), 6)
    (got * cot.cuda()).sum().backward()
    assert float((got.detach().cpu().double() - want.detach()).abs().max()) < 1e-5
    assert float((wd.grad.cpu().double() - wr.grad).abs().max()) < 1e-5 * float(wr.grad.abs().max())
