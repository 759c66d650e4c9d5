"""GPU parity of the drop-in modules against fixtures recorded from the
UNMODIFIED reference (tests/golden, oracle/make_golden.py): same state dict in,
same outputs / gradients / buffer updates out, at fp32 tolerance."""
import numpy as np
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
FWD = dict(rtol=1e-4, atol=2e-5)
BWD = dict(rtol=1e-3, atol=1e-4)


def _to(x):
    return x.to(DEV) if torch.is_tensor(x) else x


def _check_param_grads(module, want, tol=BWD):
    got = dict(module.named_parameters())
    for k, w in want.items():
        g = got[k].grad
        if w is None:
            assert g is None or float(g.abs().max()) == 0.0, k
        else:
            assert g is not None, k
            torch.testing.assert_close(g.cpu(), w, msg=lambda m: f"{k}: {m}", **tol)


@pytest.mark.parametrize("tag", ["ln_gelu", "noln_relu", "silu"])
def test_interaction_cell_vs_reference(golden, tag):
    from hierarchicalgnn_b200.gnn_utils import InteractionGNNCell
    r = golden("cell_interaction.pt")[tag]
    cell = InteractionGNNCell(r["hparams"])
    cell.load_state_dict(r["state"], strict=True)
    cell.to(DEV)
    nodes, edges = r["nodes"].to(DEV).requires_grad_(True), r["edges"].to(DEV).requires_grad_(True)
    n2, e2 = cell(nodes, edges, r["graph"].to(DEV))
    torch.testing.assert_close(n2.detach().cpu(), r["out_nodes"], **FWD)
    torch.testing.assert_close(e2.detach().cpu(), r["out_edges"], **FWD)
    ((n2 * r["w_nodes"].to(DEV)).sum() + (e2 * r["w_edges"].to(DEV)).sum()).backward()
    torch.testing.assert_close(nodes.grad.cpu(), r["grad_nodes"], **BWD)
    torch.testing.assert_close(edges.grad.cpu(), r["grad_edges"], **BWD)
    _check_param_grads(cell, r["grad_params"])


def test_hierarchical_cell_vs_reference(golden):
    from hierarchicalgnn_b200.gnn_utils import HierarchicalGNNCell
    r = golden("cell_hierarchical.pt")
    cell = HierarchicalGNNCell(r["hparams"])
    cell.load_state_dict(r["state"], strict=True)
    cell.to(DEV)
    names = ["nodes", "edges", "supernodes", "superedges", "bipartite_weights", "super_weights"]
    t = {k: r[k].to(DEV).requires_grad_(True) for k in names}
    outs = cell(t["nodes"], t["edges"], t["supernodes"], t["superedges"], r["graph"].to(DEV),
                r["bipartite_graph"].to(DEV), t["bipartite_weights"], r["super_graph"].to(DEV), t["super_weights"])
    for o, w in zip(outs, r["outs"]):
        torch.testing.assert_close(o.detach().cpu(), w, **FWD)
    sum((o * w.to(DEV)).sum() for o, w in zip(outs, r["ws"])).backward()
    for k in names:
        torch.testing.assert_close(t[k].grad.cpu(), r["grads"][k], msg=lambda m: f"{k}: {m}", **BWD)
    _check_param_grads(cell, r["grad_params"])


@pytest.mark.parametrize("tag", ["bip_train", "bip_eval", "sup_train", "sup_eval"])
def test_dynamic_graph_construction_vs_reference(golden, tag):
    from hierarchicalgnn_b200.gnn_utils import DynamicGraphConstruction
    r = golden("dynamic_graph.pt")[tag]
    m = DynamicGraphConstruction(r["weighting"], {})
    m.load_state_dict(r["state_before"], strict=True)
    m.to(DEV).train(r["training"])
    src = r["src"].to(DEV).requires_grad_(True)
    dst = src if r["sym"] else r["dst"].to(DEV).requires_grad_(True)
    graph, w, logits = m(src, dst, sym=r["sym"], norm=True, k=r["k"], logits=True)
    graph_c = graph.cpu()
    po, pr = O.canonical_edge_order(graph_c), O.canonical_edge_order(r["graph"])
    assert torch.equal(graph_c[:, po], r["graph"][:, pr])  # kNN edge list: bit-exact after canonical sort
    if not r["sym"]:
        assert torch.equal(graph_c, r["graph"])  # query-major, rank-minor order reproduced as is
    torch.testing.assert_close(w.detach().cpu()[po], r["weights"][pr], **FWD)
    torch.testing.assert_close(logits.detach().cpu()[po], r["logits"][pr], rtol=1e-4, atol=1e-4)
    inv = torch.empty_like(po)
    inv[po] = torch.arange(len(po))
    (w * r["wt"][pr][inv].to(DEV)).sum().backward()
    torch.testing.assert_close(src.grad.cpu(), r["grad_src"], **BWD)
    if not r["sym"]:
        torch.testing.assert_close(dst.grad.cpu(), r["grad_dst"], **BWD)
    for key, want in r["state_after"].items():
        torch.testing.assert_close(m.state_dict()[key].cpu(), want, rtol=1e-5, atol=1e-6, msg=lambda s: f"{key}: {s}")


@pytest.mark.parametrize("tag", ["default", "shared_noln"])
def test_ec_model_vs_reference(golden, tag):
    from hierarchicalgnn_b200.EdgeClassifier.Models.IN import EC_InteractionGNN
    r = golden("ec_model.pt")[tag]
    model = EC_InteractionGNN(r["hparams"])
    assert list(model.state_dict().keys()) == r["keys"]
    model.load_state_dict(r["state"], strict=True)
    model.to(DEV)
    x = r["x"].to(DEV)
    scores = model(x, r["graph"].to(DEV))
    torch.testing.assert_close(scores.detach().cpu(), r["scores"], rtol=1e-4, atol=5e-6)
    loss = torch.nn.functional.binary_cross_entropy(scores, r["y"].float().to(DEV))
    torch.testing.assert_close(loss.detach().cpu(), r["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    torch.testing.assert_close(x.grad.cpu(), r["grad_x"], rtol=1e-3, atol=1e-6)
    _check_param_grads(model, r["grad_params"], dict(rtol=1e-3, atol=2e-5))
    auc_ref, auc_got = O.roc_auc(r["scores"], r["y"]), O.roc_auc(scores.detach().cpu(), r["y"])
    assert abs(auc_ref - auc_got) <= 1e-3


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_bc_model_vs_reference_with_injected_clusters(golden, mode):
    from hierarchicalgnn_b200.BipartiteClassification.Models.HGNN_GMM import BC_HierarchicalGNN_GMM
    G = golden("bc_model.pt")
    r = G[mode]
    state = dict(G["train"]["state_before"])
    if mode == "eval":
        state.update(G["train"]["state_after"])
    model = BC_HierarchicalGNN_GMM(G["hparams"])
    assert list(model.state_dict().keys()) == G["keys"]
    model.load_state_dict(state, strict=True)
    model.to(DEV).train(mode == "train")
    x = G["x"].to(DEV)
    bg, scores, emb = model(x, G["graph"].to(DEV), clusters=r["clusters"].to(DEV))
    bg_c = bg.cpu()
    po, pr = O.canonical_edge_order(bg_c), O.canonical_edge_order(r["bipartite_graph"])
    assert torch.equal(bg_c[:, po], r["bipartite_graph"][:, pr])
    torch.testing.assert_close(emb.detach().cpu(), r["embeddings"], **FWD)
    torch.testing.assert_close(scores.detach().cpu()[po], r["scores"][pr], rtol=1e-3, atol=1e-4)
    inv = torch.empty_like(po)
    inv[po] = torch.arange(len(po))
    ((scores * r["ws"][pr][inv].to(DEV)).sum() + (emb * r["we"].to(DEV)).sum()).backward()
    torch.testing.assert_close(x.grad.cpu(), r["grad_x"], rtol=5e-3, atol=5e-4)
    _check_param_grads(model, r["grad_params"], dict(rtol=5e-3, atol=5e-4))
    sd = model.state_dict()
    for key, want in r["state_after"].items():
        if key == "hgnn_block.score_cut":
            continue  # owned by the (injected) clustering stage
        torch.testing.assert_close(sd[key].cpu(), want, rtol=1e-4, atol=1e-5, msg=lambda s: f"{key}: {s}")


def _gmm_loglik(x, p):
    import math
    comps = []
    for k in (0, 3):
        comps.append(math.log(p[k]) - 0.5 * math.log(2 * math.pi * p[k + 2]) - (x - p[k + 1]) ** 2 / (2 * p[k + 2]))
    return float(torch.logsumexp(torch.stack(comps), 0).mean())


def test_bc_own_clustering_on_reference_embeddings(golden):
    """Clustering parity is statistical (the reference's sklearn GMM is unseeded, so any
    EM optimum is a legitimate reference outcome): on the recorded model the on-device
    EM must reach a likelihood no worse than sklearn's, the cut must lie between the two
    modes, and the full forward through own clustering must run."""
    from sklearn.mixture import GaussianMixture
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200.BipartiteClassification.Models.HGNN_GMM import BC_HierarchicalGNN_GMM
    G = golden("bc_model.pt")
    model = BC_HierarchicalGNN_GMM(G["hparams"])
    model.load_state_dict(G["train"]["state_before"], strict=True)
    model.to(DEV).train()
    x, graph = G["x"].to(DEV), G["graph"].to(DEV)
    directed = torch.cat([graph, graph.flip(0)], 1)
    with torch.no_grad():
        emb, _, _ = model.ignn_block(x, directed)
        lik = torch.atanh((emb[directed[0]] * emb[directed[1]]).sum(-1).clamp(-1 + 1e-7, 1 - 1e-7))
        p = ops.gmm1d_fit(lik).cpu().double().tolist()
        clusters = model.hgnn_block.clustering(x, emb, directed).cpu()
    sk = GaussianMixture(2, random_state=0).fit(lik.cpu().double().numpy().reshape(-1, 1))
    assert _gmm_loglik(lik.cpu().double(), p) >= sk.score(lik.cpu().double().numpy().reshape(-1, 1)) - 5e-3
    cut = float(model.hgnn_block.score_cut)
    assert min(p[1], p[4]) < cut < max(p[1], p[4])
    assert clusters.shape == G["train"]["clusters"].shape and int(clusters.max()) >= 0
    if int(clusters.max()) >= 2:  # an untrained model may cluster degenerately; BatchNorm then rejects a 1-edge graph
        bg, scores, emb2 = model(x, graph)
        assert bg.shape[0] == 2 and scores.shape[0] == bg.shape[1] and emb2.shape == (x.shape[0], G["hparams"]["emb_dim"])
        scores.sum().backward()


def test_clustering_partition_equals_oracle_on_separated_embeddings(golden):
    """With well-separated modes every EM start reaches the same optimum, so the GPU
    partition (EM + closed-form cut + union-find + relabel) must equal the oracle's
    restatement of HierarchicalGNNBlock.clustering exactly."""
    from hierarchicalgnn_b200.BipartiteClassification.Models.HGNN_GMM import BC_HierarchicalGNN_GMM
    from hierarchicalgnn_b200.synth import synth_event, direction_embeddings
    G = golden("bc_model.pt")
    ev = synth_event(120, 8, 0.05, 3.0, seed=1234)
    emb = direction_embeddings(ev, dim=8, noise=0.03, seed=1)
    directed = torch.cat([ev.edge_index, ev.edge_index.flip(0)], 1)
    want, want_cut = O.gmm_clustering(G["hparams"], emb, directed, torch.tensor([float("inf")]), training=True)
    model = BC_HierarchicalGNN_GMM(G["hparams"]).to(DEV).train()
    with torch.no_grad():
        got = model.hgnn_block.clustering(ev.x.to(DEV), emb.to(DEV), directed.to(DEV)).cpu()
    torch.testing.assert_close(model.hgnn_block.score_cut.cpu(), want_cut.float(), rtol=2e-2, atol=2e-2)
    assert torch.equal(got, want)
    assert int(got.max()) + 1 >= 100  # ~one supernode per particle


def test_ec_model_medium_event_auc_and_scores_vs_oracle():
    """Config-1 shaped check at a size the oracle finishes in seconds: N=3000, E~13.5k, latent 64."""
    from hierarchicalgnn_b200.synth import synth_event
    from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
    torch.manual_seed(0)
    model = model_selector("EC-IN", dict(latent=64, n_interaction_graph_iters=4))
    kaiming_init(model)
    ev = synth_event(300, 10, 0.0, 4.0, seed=1000)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = O.ec_forward(O.cast_state(sd, torch.float64), dict(model.hparams), ev.x.double(), ev.edge_index).float()
    model.to(DEV).eval()
    with torch.no_grad():
        got = model(ev.x.to(DEV), ev.edge_index.to(DEV)).cpu()
    assert float((got - want).abs().max()) < 5e-5  # fp32 path vs fp64 oracle
    assert abs(O.roc_auc(got, ev.y_pid) - O.roc_auc(want, ev.y_pid)) <= 1e-3


def test_ec_full_config_tensor_core_path_scores_and_auc_vs_fp64_oracle():
    """BASELINE config 1 shape at half size (N=6000, E~27k; full EC-IN config: latent 128, 14 cells) on the DEFAULT
    (tensor-core) path: scores within the bf16 tolerance of the fp64 oracle (1e-2 on sigmoid scores),
    AUC difference <= 1e-3."""
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200.synth import synth_event
    from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
    torch.manual_seed(0)
    model = model_selector("EC-IN")
    kaiming_init(model)
    ev = synth_event(600, 10, 0.0, 4.0, seed=1000)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        want = O.ec_forward(O.cast_state(sd, torch.float64), dict(model.hparams), ev.x.double(), ev.edge_index).float()
    model.to(DEV).eval()
    old = ops.set_precision("auto")
    try:
        n0 = ops.TC_CALLS["count"]
        with torch.no_grad():
            got = model(ev.x.to(DEV), ev.edge_index.to(DEV)).cpu()
        assert ops.TC_CALLS["count"] - n0 == 14  # every cell's edge step ran on tensor cores
    finally:
        ops.set_precision(old)
    assert float((got - want).abs().max()) < 1e-2
    assert abs(O.roc_auc(got, ev.y_pid) - O.roc_auc(want, ev.y_pid)) <= 1e-3


def test_matching_score_table_device_assembly_matches_scipy_coo_path():
    """The assignment loss hands scipy a CSR table assembled on the device (BipartiteClassificationBase._score_table); it has
    to be the table scipy builds itself from the COO triplets (reference bipartite_classification_base.py:163-173):
    same pattern, sorted column indices, duplicates summed (fp32 sum order may differ: 1e-6 relative)."""
    from scipy.sparse import csr_matrix
    from hierarchicalgnn_b200.BipartiteClassification.bipartite_classification_base import BipartiteClassificationBase
    g = torch.Generator().manual_seed(5)
    n_p, n_s, nnz = 700, 300, 40000  # many duplicate (particle, supernode) pairs, as hits of one particle share supernodes
    rows = torch.cat([torch.randint(0, n_p, (nnz,), generator=g), torch.arange(n_p)])
    cols = torch.cat([torch.randint(0, n_s, (nnz,), generator=g), torch.arange(n_s, n_s + n_p)])
    vals = torch.cat([torch.rand(nnz, generator=g), torch.full((n_p,), 1e-12)])
    want = csr_matrix((vals.numpy(), (rows.numpy(), cols.numpy())), shape=(n_p, n_s + n_p))
    want.sum_duplicates()
    got = BipartiteClassificationBase._score_table(rows.cuda(), cols.cuda(), vals.cuda(), n_p, n_s + n_p)
    assert got.shape == want.shape and got.nnz == want.nnz
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    np.testing.assert_allclose(got.data, want.data, rtol=2e-6, atol=0)
    # an empty row set and a single entry
    one = BipartiteClassificationBase._score_table(torch.tensor([0]).cuda(), torch.tensor([1]).cuda(), torch.tensor([0.5]).cuda(), 2, 3)
    assert one.toarray().tolist() == [[0.0, 0.5, 0.0], [0.0, 0.0, 0.0]]


def test_batched_knn_equals_per_event_knn_bit_exact():
    """hgnn_knn_radius_batched: every query meets the references of its own event only; the table is the per-event tables
    (hgnn_knn_radius) with the neighbour ids offset — bit for bit, ties included, for ragged and empty events."""
    from hierarchicalgnn_b200 import ops
    g = torch.Generator().manual_seed(3)
    for dim, k, r in ((8, 5, 0.9), (8, 10, 0.7), (3, 16, 0.5)):
        nq = [300, 0, 1, 777, 130]
        nr = [200, 50, 0, 300, 129]
        qs = [torch.nn.functional.normalize(torch.randn(n, dim, generator=g)) for n in nq]
        rs = [torch.nn.functional.normalize(torch.randn(n, dim, generator=g)) for n in nr]
        rs[3][5] = rs[3][7]  # an exact tie
        qptr = torch.tensor([0] + list(torch.tensor(nq).cumsum(0)), dtype=torch.int32).cuda()
        rptr = torch.tensor([0] + list(torch.tensor(nr).cumsum(0)), dtype=torch.int32).cuda()
        got = ops.knn_radius(torch.cat(qs).cuda(), torch.cat(rs).cuda(), k, r, qptr, rptr).cpu()
        want = []
        for b in range(len(nq)):
            if nq[b] == 0:
                continue
            t = ops.knn_radius(qs[b].cuda(), rs[b].cuda(), k, r).cpu() if nr[b] else torch.full((nq[b], k), -1, dtype=torch.long)
            want.append(torch.where(t >= 0, t + int(rptr[b]), t))
        assert torch.equal(got, torch.cat(want))
        assert bool((got >= 0).any())


def test_bc_model_batched_events_equal_single_event_calls():
    """Three events collated into one disjoint graph (torch_geometric Batch layout) through BC_HierarchicalGNN_GMM in eval
    mode, supernodes = particles: per-hit embeddings, the bipartite graph and its scores equal what three single-event calls
    return (same kernels row by row; the kNN graphs are searched per event and the edge weights normalised per event)."""
    from hierarchicalgnn_b200.synth import collate_events, synth_event
    from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
    torch.manual_seed(0)
    model = model_selector("BC-HGNN-GMM", dict(latent=128))
    kaiming_init(model)
    model.cuda().eval()
    evs = [synth_event(150, 8, 0.05, 3.0, seed=7), synth_event(90, 10, 0.0, 4.0, seed=8), synth_event(200, 6, 0.1, 2.0, seed=9)]
    big = collate_events(evs)
    with torch.no_grad():
        singles = [model(ev.x.cuda(), ev.edge_index.cuda(), clusters=(ev.pid - 1).cuda()) for ev in evs]
        bg, sc, emb = model(big.x.cuda(), big.edge_index.cuda(), clusters=big.clusters.cuda(), batch=big.batch.cuda(),
                            n_events=big.num_graphs)
    assert torch.equal(emb, torch.cat([s[2] for s in singles]))
    hit0 = sn0 = 0
    want_g, want_s = [], []
    for ev, (g1, s1, _e) in zip(evs, singles):
        want_g.append(g1 + torch.tensor([[hit0], [sn0]], device=g1.device))
        want_s.append(s1.reshape(-1))
        hit0 += ev.x.shape[0]
        sn0 += ev.n_particles
    want_g, want_s = torch.cat(want_g, 1), torch.cat(want_s)
    assert torch.equal(bg, want_g)  # same edges in the same (query-major) order
    assert float((sc.reshape(-1) - want_s).abs().max()) < 1e-5


def test_clustering_of_batched_events_equals_per_event_clustering():
    """HierarchicalGNNBlock.clustering on collated events (HGNN_GMM.py:183-232): connected components never cross events, so
    with the score cut fixed the cluster labels are the per-event labels, numbered event by event."""
    from hierarchicalgnn_b200.gnn_utils import GraphPlans
    from hierarchicalgnn_b200.synth import collate_events, direction_embeddings, synth_event
    from hierarchicalgnn_b200.training_utils import model_selector
    model = model_selector("BC-HGNN-GMM", dict(latent=128)).cuda().eval()
    blk = model.hgnn_block
    blk.score_cut.fill_(1.5)
    evs = [synth_event(150, 8, 0.05, 3.0, seed=7), synth_event(90, 10, 0.0, 4.0, seed=8), synth_event(200, 6, 0.1, 2.0, seed=9)]
    embs = [direction_embeddings(ev, seed=i) for i, ev in enumerate(evs)]
    big = collate_events(evs)

    def run(x, emb, graph):
        g = torch.cat([graph, graph.flip(0)], 1).cuda()
        return blk.clustering(x.cuda(), emb.cuda(), GraphPlans(g, x.shape[0], x.shape[0]))
    got = run(big.x, torch.cat(embs), big.edge_index)
    want, off = [], 0
    for ev, emb in zip(evs, embs):
        c = run(ev.x, emb, ev.edge_index)
        assert int(c.max()) > 10
        want.append(torch.where(c >= 0, c + off, c))
        off += int(c.max()) + 1
    assert torch.equal(got, torch.cat(want))


def test_training_step_on_collated_events():
    """BipartiteClassificationBase.training_step on a torch_geometric-style batch: a batch of ONE event carrying the event
    vector gives exactly the loss of the plain single-event step (same graphs, same matching); a batch of three events gives a
    finite loss and a gradient for every parameter that gets one on a single event."""
    from types import SimpleNamespace
    from hierarchicalgnn_b200.synth import collate_events, synth_event
    from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
    evs = [synth_event(150, 8, 0.05, 3.0, seed=17), synth_event(90, 10, 0.0, 4.0, seed=18), synth_event(200, 6, 0.1, 2.0, seed=19)]

    def fresh():
        torch.manual_seed(0)
        m = model_selector("BC-HGNN-GMM", dict(latent=128, loss_schedule=0.5))
        kaiming_init(m)
        return m.cuda().train()

    def batch_of(events, with_vector):
        b = collate_events(events)
        ns = SimpleNamespace(x=b.x.cuda(), edge_index=b.edge_index.cuda(), pid=b.pid.cuda(), pt=b.pt.cuda())
        if with_vector:
            ns.batch, ns.num_graphs = b.batch.cuda(), b.num_graphs
        return ns, b.clusters.cuda()

    losses, grads = [], []
    for events, vec in (([evs[0]], False), ([evs[0]], True), (evs, True)):
        m = fresh()
        b, clusters = batch_of(events, vec)
        m.hgnn_block.clustering = lambda x, emb, graph: clusters  # supernodes = particles
        loss = m.training_step(b, 0)
        loss.backward()
        assert torch.isfinite(loss)
        losses.append(float(loss.detach()))
        grads.append({k: p.grad for k, p in m.named_parameters()})
    # the per-event mean of the edge weights is an ordered segment sum, torch.mean a tree: 1e-7 relative on the weights
    assert abs(losses[0] - losses[1]) < 1e-5 * abs(losses[0])
    for k, g in grads[0].items():
        if g is not None and float(g.abs().max()) > 0:
            assert float((g - grads[1][k]).norm()) < 1e-3 * float(g.norm()) + 1e-6, k  # (biases in front of a LayerNorm: analytically 0)
            assert grads[2][k] is not None and bool(torch.isfinite(grads[2][k]).all()), k



def test_validation_steps_report_tracking_metrics():
    """shared_evaluation / validation_step of both task bases (edge_classifier_base.py:135-197,
    bipartite_classification_base.py:226-301): the EC step with a perfect scorer (score = truth) reconstructs every track of
    a clean synthetic event — efficiency and purity 1 — and its loss equals the training loss of the same scores; the BC
    step returns a finite loss, logs the four metrics, and with the event vector accepts a collated batch."""
    from types import SimpleNamespace
    from hierarchicalgnn_b200.synth import collate_events, synth_event
    from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
    ev = synth_event(120, 8, 0.0, 3.0, seed=31)
    b = SimpleNamespace(x=ev.x.cuda(), edge_index=ev.edge_index.cuda(), pid=ev.pid.cuda(), pt=(ev.pt + 1.0).cuda(),
                        y=ev.y.cuda(), y_pid=ev.y_pid.cuda())
    ec = model_selector("EC-IN", dict(latent=128, n_interaction_graph_iters=1)).cuda().eval()
    logged = {}
    ec.log_dict = lambda d, *a, **k: logged.update({k_: (float(v) if torch.is_tensor(v) else v) for k_, v in d.items()})
    ec.forward = lambda x, graph: b.y_pid.float().clamp(1e-4, 1 - 1e-4)  # a perfect scorer
    graph, loss = ec.shared_evaluation(b, 0, log=True)
    assert graph.shape[0] == 2 and torch.isfinite(loss)
    assert logged["track_eff"] == 1.0 and logged["track_pur"] == 1.0 and logged["hit_eff"] == 1.0 and logged["hit_pur"] == 1.0
    assert abs(float(ec.validation_step(b)) - float(ec.training_step(b))) < 1e-6

    torch.manual_seed(0)
    bc = model_selector("BC-HGNN-GMM", dict(latent=128, loss_schedule=0.5))
    kaiming_init(bc)
    bc.cuda().eval()
    seen = {}
    bc.log_dict = lambda d, *a, **k: seen.update(d)
    evs = [synth_event(150, 8, 0.05, 3.0, seed=41), synth_event(90, 10, 0.0, 4.0, seed=42)]
    for events in ([evs[0]], evs):
        big = collate_events(events)
        nb = SimpleNamespace(x=big.x.cuda(), edge_index=big.edge_index.cuda(), pid=big.pid.cuda(), pt=big.pt.cuda())
        if len(events) > 1:
            nb.batch, nb.num_graphs = big.batch.cuda(), big.num_graphs
        clusters = big.clusters.cuda()
        bc.hgnn_block.clustering = lambda x, emb, graph: clusters
        seen.clear()
        loss = bc.validation_step(nb)
        assert torch.isfinite(loss) and not loss.requires_grad
        assert {"val_loss", "val_embedding_loss", "val_assignment_loss", "track_eff", "track_pur", "hit_eff", "hit_pur"} <= set(seen)
