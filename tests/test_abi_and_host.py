"""CPU-side checks: the C-ABI library loads and exports every symbol declared
in include/hgnn_b200.h (no compute calls), and the host-side mirror of the
reference interface (state-dict layout, init, hparams, cut solver) is right."""
import ctypes
import math
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hgnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hgnn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from hierarchicalgnn_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 25
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in hgnn_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.lib().hgnn_abi_version() == 1


def test_mlp_desc_struct_layout_matches_header():
    from hierarchicalgnn_b200 import _lib
    # 4 scalars (16 B) + 3+3 pointers + 3 ints (+pad) + 4+4 ints + 16 pointers + 1 pointer
    assert ctypes.sizeof(_lib.MlpDesc) == 16 + 6 * 8 + 3 * 4 + 8 * 4 + 4 + 17 * 8


def test_no_cpu_fallback():
    from hierarchicalgnn_b200 import ops
    from hierarchicalgnn_b200._lib import HgnnError
    with pytest.raises(HgnnError, match="no CPU fallback"):
        ops.scatter_add(torch.randn(4, 4), torch.tensor([0, 1, 1, 0]), dim_size=2)
    from hierarchicalgnn_b200.utils import make_mlp
    with pytest.raises(HgnnError):
        make_mlp(4, 8, 4, 2)(torch.randn(3, 4))


def test_make_mlp_state_dict_layout_matches_reference(golden):
    from hierarchicalgnn_b200.utils import make_mlp
    for (n, ln, oa), rec in golden("make_mlp.pt").items():
        net = make_mlp(6, 10, 4, n, hidden_activation="GELU", output_activation=oa, layer_norm=ln)
        assert list(net.state_dict().keys()) == rec["keys"], (n, ln, oa)
        net.load_state_dict(rec["state"], strict=True)


def test_model_state_dict_keys_and_order_match_reference(golden):
    from hierarchicalgnn_b200.training_utils import model_selector
    ec = golden("ec_model.pt")
    for tag in ("default", "shared_noln"):
        hp = ec[tag]["hparams"]
        over = {k: hp[k] for k in ("latent", "n_interaction_graph_iters", "share_weight", "layernorm", "hidden_output_activation")}
        m = model_selector("EC-IN", over)
        assert list(m.state_dict().keys()) == ec[tag]["keys"]
        m.load_state_dict(ec[tag]["state"], strict=True)
    bc = golden("bc_model.pt")
    hp = bc["hparams"]
    m = model_selector("4", {k: hp[k] for k in ("latent", "n_interaction_graph_iters", "n_hierarchical_graph_iters")})
    assert list(m.state_dict().keys()) == bc["keys"]
    m.load_state_dict(bc["train"]["state_before"], strict=True)
    assert torch.isinf(m.hgnn_block.score_cut).all() and float(m.hgnn_block.super_graph_construction.knn_radius) == 1.0


def test_full_size_parameter_counts_match_survey():
    from hierarchicalgnn_b200.training_utils import model_selector
    ec = model_selector("EC-IN")
    assert len(ec.state_dict()) == 310 and sum(p.numel() for p in ec.parameters()) == 4441089
    bc = model_selector("BC-HGNN-GMM", dict(latent=128))
    assert len(bc.state_dict()) == 433 and sum(p.numel() for p in bc.parameters()) == 6358517


def test_kaiming_init_matches_reference_rule():
    from hierarchicalgnn_b200.training_utils import kaiming_init, model_selector
    from oracle import reference_harness as rh
    a = model_selector("EC-IN", dict(latent=16, n_interaction_graph_iters=1))
    b = model_selector("EC-IN", dict(latent=16, n_interaction_graph_iters=1))
    torch.manual_seed(5)
    kaiming_init(a)
    torch.manual_seed(5)
    rh.kaiming_init(b)
    for (k, p), q in zip(a.named_parameters(), b.parameters()):
        assert torch.equal(p, q), k
    w0 = a.ignn_block.node_encoder[0].weight
    assert abs(float(w0.std()) - 1 / math.sqrt(3)) < 0.2
    assert float(a.ignn_block.node_encoder[0].bias.abs().max()) == 0.0


def test_process_hparams_and_aliases():
    from hierarchicalgnn_b200.training_utils import load_hparams
    hp = load_hparams("1")
    assert hp["hidden"] == 256 and hp["latent"] == 128 and hp["cluster_granularity"] == 0
    hp = load_hparams("BC-HGNN-GMM", dict(latent=128))
    assert hp["hidden"] == 256 and hp["emb_dim"] == 8 and hp["cluster_granularity"] == 5
    with pytest.raises(ValueError):
        load_hparams("nope")


def test_gaussian_cut_agrees_with_reference_fsolve_formulation():
    import numpy as np
    from scipy.optimize import fsolve
    from sklearn.mixture import GaussianMixture
    from hierarchicalgnn_b200.BipartiteClassification.Models.HGNN_GMM import gaussian_cut
    rng = np.random.RandomState(0)
    x = np.concatenate([rng.normal(0.2, 0.3, 6000), rng.normal(2.4, 0.5, 4000)]).reshape(-1, 1)
    gmm = GaussianMixture(2, random_state=0).fit(x)
    for gran in (0, 5, -2):
        sg = lambda v: 1 / (1 + np.exp(-v))
        lo, hi = gmm.means_.argmin(), gmm.means_.argmax()
        f = lambda t: sg(gran) * gmm.predict_proba(t.reshape(-1, 1))[:, lo] - sg(-gran) * gmm.predict_proba(t.reshape(-1, 1))[:, hi]
        want = float(fsolve(f, gmm.means_.mean()).item())
        params = []
        for k in (0, 1):
            params += [gmm.weights_[k], gmm.means_[k, 0], gmm.covariances_[k, 0, 0]]
        cut, found = gaussian_cut(params, gran)
        assert found and abs(cut - want) < 1e-4, (gran, cut, want)


def test_synthetic_event_is_deterministic_and_shaped():
    from hierarchicalgnn_b200.synth import synth_event, synth_edge_problem
    a, b = synth_event(50, 10, 0.1, 4.0, seed=7), synth_event(50, 10, 0.1, 4.0, seed=7)
    assert torch.equal(a.x, b.x) and torch.equal(a.edge_index, b.edge_index)
    assert a.x.shape == (550, 3) and a.edge_index.shape[0] == 2 and a.x.dtype == torch.float32
    assert int(a.y_pid.sum()) >= 50 * 9  # random fakes may join same-particle hits
    n, e, g = synth_edge_problem(1000, 32)
    assert n.shape == (100, 32) and e.shape == (1000, 32) and torch.equal(g[:, :500], g[:, 500:].flip(0))


def test_lightning_compat_surface():
    from hierarchicalgnn_b200.training_utils import model_selector
    m = model_selector("EC-IN", dict(latent=16, n_interaction_graph_iters=1))
    assert m.hparams["lr"] == 0.001 and "loss_schedule" not in m.hparams
    m.log("a", 1.0)
    m.log_dict({"b": 2.0})
    assert m.trainer.current_epoch == 0 and m.trainer.global_step == 0
    (opt,), (sched,) = m.configure_optimizers()
    assert isinstance(opt, torch.optim.AdamW) and opt.defaults["amsgrad"] and sched["interval"] == "epoch"
    for hook in ("training_step", "validation_step", "test_step", "optimizer_step", "configure_optimizers"):
        assert callable(getattr(m, hook))


def test_kernel_scratch_is_an_autograd_saved_tensor_released_by_backward():
    """ops._save_with_scratch / _saved_and_scratch (host logic, CPU tensors): scratch a backward needs travels as a saved
    tensor, so it is dropped when the backward has run even while the outputs keep the graph nodes alive; without scratch
    the helpers degrade to plain save_for_backward."""
    import gc
    import weakref
    from hierarchicalgnn_b200 import ops

    refs = {}

    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, with_scratch):
            scratch = torch.full((1024,), 3.0) if with_scratch else None
            if scratch is not None:
                refs["scratch"] = weakref.ref(scratch)
            ops._save_with_scratch(ctx, [x], scratch)
            return x * 2

        @staticmethod
        def backward(ctx, g):
            (x,), scratch = ops._saved_and_scratch(ctx)
            refs["seen"] = None if scratch is None else float(scratch[0])
            return g * 2, None

    x = torch.ones(4, requires_grad=True)
    y = F.apply(x, True)
    loss = y.sum()
    assert refs["scratch"]() is not None            # alive while the backward is still to come
    loss.backward()
    gc.collect()
    assert refs["seen"] == 3.0
    assert y.grad_fn is not None and refs["scratch"]() is None   # graph node still referenced, scratch gone
    y2 = F.apply(x, False)
    y2.sum().backward()
    assert refs["seen"] is None


def test_tile_row_plans_group_the_forward_rows_by_source_and_destination():
    """ops._tile_row_plans (host logic of the node-level adjoint, CPU tensors through duck-typed plans): the CSR handed to
    hgnn_tc_edge_backward must list, per node, exactly the forward's TILE ROWS whose edge leaves / enters that node — for
    rows in edge-id order (.fused), for rows in the by-destination plan's order (.edge_step), and for edges stored sorted."""
    from hierarchicalgnn_b200 import ops

    class Plan:  # what SegmentPlan exposes: items ordered by (key, item id)
        def __init__(self, keys, n):
            self.n_items, self.n_segments = keys.numel(), n
            self.perm = torch.argsort(keys, stable=True).to(torch.int32)
            self.rowptr = torch.zeros(n + 1, dtype=torch.int32)
            self.rowptr[1:] = torch.cumsum(torch.bincount(keys, minlength=n), 0)
            self.keys32 = keys.to(torch.int32)

        def is_identity(self):
            return bool((self.keys32[1:] >= self.keys32[:-1]).all())

    g = torch.Generator().manual_seed(5)
    N, E = 13, 200
    for presorted in (False, True):
        graph = torch.randint(0, N, (2, E), generator=g)
        if presorted:
            graph = graph[:, torch.argsort(graph[1], stable=True)]
        ps, pd = Plan(graph[0], N), Plan(graph[1], N)
        for dst_sorted in (False, True):
            src_rows, src_ptr, dst_rows, dst_ptr = ops._tile_row_plans(ps, pd, dst_sorted)
            # tile row j holds edge row_edge[j]
            row_edge = pd.perm.long() if (dst_sorted and not pd.is_identity()) else torch.arange(E)
            if dst_rows is None:
                dst_rows = torch.arange(E, dtype=torch.int32)
            for rows, ptr, key in ((src_rows, src_ptr, graph[0]), (dst_rows, dst_ptr, graph[1])):
                assert sorted(rows.tolist()) == list(range(E))            # every tile row exactly once
                for n in range(N):
                    mine = rows[int(ptr[n]):int(ptr[n + 1])].long()
                    assert bool((key[row_edge[mine]] == n).all())          # ... in the segment of its node
                    assert int(ptr[n + 1] - ptr[n]) == int((key == n).sum())
            ps.__dict__.pop("_tile_rows", None)


def test_oracle_side_input_generators_equal_the_package_generators():
    """bench.py's CPU arms build their inputs from oracle/inputs.py (no product import); same seeds -> same tensors."""
    from hierarchicalgnn_b200 import synth as S
    from oracle import inputs as I
    a, b = I.synth_edge_problem(3000, 32, seed=5), S.synth_edge_problem(3000, 32, seed=5)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    a, b = I.synth_edge_problem(2000, 16, seed=6, nodes_per_edge=0.04, power_law=True), S.synth_edge_problem(2000, 16, seed=6, nodes_per_edge=0.04, power_law=True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    ea, eb = I.synth_event(50, 7, 0.1, 3.0, seed=9), S.synth_event(50, 7, 0.1, 3.0, seed=9)
    for k in ("x", "pid", "pt", "edge_index", "y_pid"):
        assert torch.equal(getattr(ea, k), getattr(eb, k)), k


def test_ops_refuse_tensors_of_another_device_than_the_current_one():
    """Launches go to the current device's stream: a tensor living elsewhere must raise, not fault."""
    from hierarchicalgnn_b200 import ops, _lib

    class Fake:
        is_cuda = True
        device = torch.device("cuda", 1)
    import unittest.mock as mock
    with mock.patch.object(torch.cuda, "current_device", return_value=0):
        with pytest.raises(_lib.HgnnError, match="current CUDA device"):
            ops._need_cuda(Fake())


def test_kaiming_init_moves_the_version_counter_and_drops_packed_images():
    from hierarchicalgnn_b200.training_utils import invalidate_packed_weights, kaiming_init
    from hierarchicalgnn_b200.utils import make_mlp
    net = make_mlp(12, 8, 4, 2, layer_norm=True, output_activation="Tanh")
    v0 = net[0].weight._version
    object.__setattr__(net, "_tc_cache", ("stale",))
    net.__dict__["_row_cache"] = {0: "stale"}
    kaiming_init(net)
    assert net[0].weight._version > v0
    assert "_tc_cache" not in net.__dict__ and "_row_cache" not in net.__dict__
    net.__dict__["_split_cache"] = {0: "stale"}
    invalidate_packed_weights(net)
    assert "_split_cache" not in net.__dict__


def test_assignment_loss_on_collated_events_treats_particles_per_event():
    """A torch_geometric-style batch keeps the per-event particle ids, so the same id names different particles in different
    events. With the ``batch`` vector the assignment loss (reference bipartite_classification_base.py:152-191, which only
    ever sees one event) matches particles per event: the loss equals the one computed after relabelling the particles with
    globally unique ids, and a single event gives the same loss with and without the vector. CPU tensors: scipy matching."""
    from types import SimpleNamespace
    from hierarchicalgnn_b200.training_utils import model_selector
    model = model_selector("BC-HGNN-GMM", dict(latent=128))
    g = torch.Generator().manual_seed(4)
    hits, parts, sn = [60, 45], [7, 5], [6, 4]
    pid, pt, ev, graph, h0, s0 = [], [], [], [], 0, 0
    for b in range(2):
        p = torch.randint(0, parts[b] + 1, (hits[b],), generator=g)  # 0 = noise, ids collide between the events
        pid.append(p)
        ppt = 0.5 + 2.0 * torch.rand(parts[b] + 1, generator=g)
        pt.append(torch.where(p > 0, ppt[p], torch.zeros(())))
        ev.append(torch.full((hits[b],), b))
        src = torch.arange(hits[b]).repeat_interleave(3)
        dst = torch.randint(0, sn[b], (src.numel(),), generator=g)
        graph.append(torch.stack([src + h0, dst + s0]))
        h0 += hits[b]
        s0 += sn[b]
    pid, pt, ev, graph = torch.cat(pid), torch.cat(pt), torch.cat(ev), torch.unique(torch.cat(graph, 1), dim=1)
    scores = torch.rand(graph.shape[1], generator=g).clamp(0.05, 0.95)
    collated = SimpleNamespace(pid=pid, pt=pt, batch=ev)
    unique_ids = SimpleNamespace(pid=torch.where(pid > 0, pid + 100 * ev, pid), pt=pt)
    a = model.assignment_loss(collated, graph, scores)
    b = model.assignment_loss(unique_ids, graph, scores)
    assert torch.isfinite(a) and abs(float(a) - float(b)) < 1e-6 * max(1.0, abs(float(b)))
    # colliding ids WITHOUT the vector merge particles of different events: a different (wrong) loss
    c = model.assignment_loss(SimpleNamespace(pid=pid, pt=pt), graph, scores)
    assert abs(float(c) - float(b)) > 1e-4
    one = slice(0, hits[0])
    e0 = graph[0] < hits[0]
    x = model.assignment_loss(SimpleNamespace(pid=pid[one], pt=pt[one]), graph[:, e0], scores[e0])
    y = model.assignment_loss(SimpleNamespace(pid=pid[one], pt=pt[one], batch=ev[one]), graph[:, e0], scores[e0])
    assert float(x) == float(y)


def test_native_block_matching_equals_scipy_per_block():
    """hgnn_match_blocks_max (host threads, csrc/matching.cu) against scipy.sparse.csgraph.min_weight_full_bipartite_matching
    (what the reference calls, bipartite_classification_base.py:173): same matched columns block by block on random
    particle x (supernode + virtual supernode) score tables, an empty block in the middle, and an error for a block
    that has no matching covering its rows."""
    import numpy as np
    from scipy.sparse import block_diag, csr_matrix
    from scipy.sparse.csgraph import min_weight_full_bipartite_matching
    from hierarchicalgnn_b200 import _lib
    from hierarchicalgnn_b200.BipartiteClassification.bipartite_classification_base import BipartiteClassificationBase as Base
    rng = np.random.default_rng(5)

    def table(n_p, n_s, nnz):
        rows = np.concatenate([rng.integers(0, n_p, nnz), np.arange(n_p)])
        cols = np.concatenate([rng.integers(0, n_s, nnz), np.arange(n_s, n_s + n_p)])
        vals = np.concatenate([rng.random(nnz), np.full(n_p, 1e-12)]).astype(np.float32)
        m = csr_matrix((vals, (rows, cols)), shape=(n_p, n_s + n_p))
        m.sum_duplicates()
        return m
    blocks = [table(40, 25, 300), table(1, 1, 1), table(300, 280, 6000), table(90, 5, 400)]
    big = block_diag(blocks, format="csr")
    ptr = [0, 40, 40, 41, 341, 431]  # an empty block between the first two
    rows, cols = Base._match_blocks(big, ptr)
    assert rows.tolist() == list(range(431)) and len(set(cols.tolist())) == 431
    r0 = c0 = 0
    for b in blocks:
        r, c = min_weight_full_bipartite_matching(b, maximize=True)
        assert np.array_equal(cols[r0:r0 + b.shape[0]] - c0, c[np.argsort(r)])
        r0, c0 = r0 + b.shape[0], c0 + b.shape[1]
    # two rows that can only take the same column: no covering matching -> error, not a wrong answer
    bad = csr_matrix((np.ones(2, dtype=np.float32), (np.array([0, 1]), np.array([0, 0]))), shape=(2, 3))
    with pytest.raises(_lib.HgnnError):
        Base._match_blocks(bad, [0, 2])


def test_tracking_metrics_equal_the_dense_restatement_of_the_reference():
    """hierarchicalgnn_b200.tracking_utils.eval_metrics (sorted list of the non-zero particle x candidate cells, torch ops)
    against oracle.eval_metrics_dense (the reference's tracking_utils.py:18-83 with a dense matrix): noisy track candidates
    built from the truth — merged tracks, split tracks, stolen hits, noise hits, tiny candidates — over several seeds and cuts."""
    from types import SimpleNamespace
    from hierarchicalgnn_b200.tracking_utils import eval_metrics
    from oracle import hgnn_oracle as O
    checked = 0
    for seed in range(8):
        g = torch.Generator().manual_seed(100 + seed)
        n_part, hpp = 40, int(torch.randint(4, 10, (1,), generator=g))
        pid = torch.arange(1, n_part + 1).repeat_interleave(hpp)
        pid = torch.cat([pid, torch.zeros(30, dtype=torch.long)])  # noise hits
        pt_p = 0.3 + 2.0 * torch.rand(n_part + 1, generator=g)
        pt = torch.where(pid > 0, pt_p[pid], torch.zeros(()))
        n = pid.numel()
        cand = pid.clone() * 3  # candidate = particle, then damage
        cand[pid == 0] = torch.randint(0, 3 * n_part, (30,), generator=g)
        steal = torch.rand(n, generator=g) < 0.15
        cand[steal] = torch.randint(0, 3 * n_part, (int(steal.sum()),), generator=g)
        cand[(pid == 5) | (pid == 6)] = 15                   # two particles merged into one candidate
        half = (pid == 9) & (torch.arange(n) % 2 == 0)
        cand[half] = 1000                                     # a particle split over two candidates
        hit_ids = torch.arange(n)
        drop = torch.rand(n, generator=g) < 0.1              # unassigned hits
        bg = torch.stack([hit_ids[~drop], cand[~drop]])
        event = SimpleNamespace(pid=pid, pt=pt)
        for pt_cut, nhits_cut, maj in ((1.0, 5, 0.5), (0.5, 3, 0.5), (1.0, 4, 0.7)):
            want = O.eval_metrics_dense(bg, pid, pt, pt_cut, nhits_cut, maj)
            got = eval_metrics(bg.clone(), event, pt_cut=pt_cut, nhits_cut=nhits_cut, majority_cut=maj, primary=False)
            for k in want:
                assert abs(float(got[k]) - float(want[k])) < 1e-9, (seed, pt_cut, k, got, want)
            checked += want["track_eff"] > 0
    assert checked >= 12
    assert eval_metrics(torch.zeros(2, 0, dtype=torch.long), SimpleNamespace(pid=torch.ones(3, dtype=torch.long), pt=torch.ones(3))) == \
        {"track_eff": 0, "track_pur": 0, "hit_eff": 0, "hit_pur": 0}


def test_collate_events_and_event_offsets():
    """synth.collate_events lays several events out like a torch_geometric Batch (rows event by event, edge_index offset,
    batch vector, ptr, per-event particle ids kept, cluster labels offset event by event); utils.event_offsets recovers the
    row offsets from the batch vector, empty events included."""
    from hierarchicalgnn_b200.synth import collate_events, synth_event
    from hierarchicalgnn_b200.utils import event_offsets
    evs = [synth_event(20, 5, 0.1, 2.0, seed=1), synth_event(7, 4, 0.0, 3.0, seed=2), synth_event(11, 6, 0.2, 1.0, seed=3)]
    b = collate_events(evs)
    sizes = [e.x.shape[0] for e in evs]
    assert b.num_graphs == 3 and b.ptr.tolist() == [0, sizes[0], sizes[0] + sizes[1], sum(sizes)]
    assert torch.equal(torch.bincount(b.batch), torch.tensor(sizes))
    off = 0
    lo = 0
    for i, e in enumerate(evs):
        n, m = e.x.shape[0], e.edge_index.shape[1]
        assert torch.equal(b.x[off:off + n], e.x) and torch.equal(b.pid[off:off + n], e.pid)
        assert torch.equal(b.edge_index[:, lo:lo + m] - off, e.edge_index)
        assert bool((b.batch[b.edge_index[0, lo:lo + m]] == i).all()) and bool((b.batch[b.edge_index[1, lo:lo + m]] == i).all())
        off, lo = off + n, lo + m
    real = b.clusters >= 0
    assert torch.equal(b.clusters[~real], torch.full((int((~real).sum()),), -1))
    assert int(b.clusters.max()) + 1 == sum(e.n_particles for e in evs)
    ev_of_cluster = torch.zeros(int(b.clusters.max()) + 1, dtype=torch.long).scatter_(0, b.clusters[real], b.batch[real])
    assert bool((ev_of_cluster[1:] >= ev_of_cluster[:-1]).all())  # supernode ids ascend with the event id
    assert event_offsets(b.batch, 3).tolist() == b.ptr.tolist()
    assert event_offsets(torch.tensor([0, 0, 2, 2, 2]), 4).tolist() == [0, 2, 2, 5, 5]  # events 1 and 3 are empty
