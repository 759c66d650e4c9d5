"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: gradient all-reduce for the data-parallel mode and
the destination-partitioned interaction cell (all-gather forward / reduce-scatter backward) against the
single-process oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hgnn_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world=2):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] is None, r[1]
    return dict(res)


def _entry(fn, rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.set_num_threads(1)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        out = fn(rank, world)
        dist.destroy_process_group()
        q.put((rank, None, out) if False else (rank, None))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


def _dp_case(rank, world):
    from hierarchicalgnn_b200.parallel import allreduce_gradients, clip_grad_norm_
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    dead = torch.nn.Linear(3, 3)  # never used: grad stays None on every rank
    data = torch.randn(world, 7, 6, generator=torch.Generator().manual_seed(1))
    net(data[rank]).square().mean().backward()
    params = list(net.parameters()) + list(dead.parameters())
    allreduce_gradients(params, bucket_bytes=64)  # tiny buckets: exercises the flush path
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    ref.load_state_dict(net.state_dict())
    sum(ref(data[r]).square().mean() for r in range(world)).div(world).backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        torch.testing.assert_close(p.grad, q.grad, rtol=1e-5, atol=1e-6)
    for p in dead.parameters():
        assert p.grad is None  # no gradient on any rank: stays None, the optimizer skips it as on one GPU
    total = clip_grad_norm_(list(net.parameters()), 1e-3)
    after = torch.sqrt(sum(p.grad.square().sum() for p in net.parameters()))
    assert float(after) <= 1e-3 * 1.001 and float(total) > 0


class _ToyTask(torch.nn.Module):
    """LightningModule-shaped toy: training_step / optimizer_step, a dead sub-network, a BatchNorm buffer."""

    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.Tanh(), torch.nn.Linear(32, 32), torch.nn.Tanh(),
                                       torch.nn.Linear(32, 1))
        self.dead = torch.nn.Linear(3, 3)
        self.bn = torch.nn.BatchNorm1d(6)

    def training_step(self, batch, batch_idx=0):
        return self.net(self.bn(batch)).square().mean()

    def optimizer_step(self, optimizer=None, **kw):
        optimizer.step()
        optimizer.zero_grad()


def _dp_trainer_case(rank, world):
    """DataParallelTrainer (bucketed all-reduce launched from backward hooks, clip, buffer broadcast) == one process that
    averages the two ranks' gradients by hand, for three optimizer steps."""
    from hierarchicalgnn_b200.parallel import DataParallelTrainer, clip_grad_norm_
    torch.manual_seed(0)
    model = _ToyTask()
    ref = _ToyTask()
    ref.load_state_dict(model.state_dict())
    opt = torch.optim.AdamW(model.parameters(), lr=1e-2)
    opt_ref = torch.optim.AdamW(ref.parameters(), lr=1e-2)
    tr = DataParallelTrainer(model, opt, clip=0.05, bucket_bytes=256)  # several buckets
    assert len(tr.buckets.buckets) >= 3
    data = torch.randn(3, world, 9, 6, generator=torch.Generator().manual_seed(1))
    for step in range(3):
        loss = tr.step(data[step, rank])
        # reference: mean of the per-rank losses; BatchNorm statistics as rank 0 sees them (buffers are broadcast from rank 0)
        import copy
        opt_ref.zero_grad(set_to_none=True)
        live = [p for k, p in ref.named_parameters() if not k.startswith("dead.")]
        for r in range(world):
            m = copy.deepcopy(ref)
            gs = torch.autograd.grad(m.training_step(data[step, r]), [p for k, p in m.named_parameters() if not k.startswith("dead.")])
            for p, g in zip(live, gs):
                p.grad = g / world if p.grad is None else p.grad + g / world
            if r == 0:
                ref.bn.load_state_dict(m.bn.state_dict())
        clip_grad_norm_(list(ref.parameters()), 0.05)
        opt_ref.step()
        for (k, p), q in zip(model.named_parameters(), ref.parameters()):
            torch.testing.assert_close(p, q, rtol=1e-5, atol=1e-6, msg=lambda m: f"step {step} {k}: {m}")
        for (k, b), c in zip(model.named_buffers(), ref.buffers()):
            if b.is_floating_point():
                torch.testing.assert_close(b, c, rtol=1e-5, atol=1e-6, msg=lambda m: f"step {step} buffer {k}: {m}")
    assert all(p.grad is None for p in model.dead.parameters())
    assert tr.global_step == 3
    tr.buckets.remove()


def _partition_case(rank, world, carry=False):
    from hierarchicalgnn_b200.parallel import (allreduce_gradients, pad_rows, partition_by_destination,
                                               partitioned_interaction_cell)
    from hierarchicalgnn_b200.synth import synth_edge_problem
    from oracle.reference_harness import kaiming_init
    from hierarchicalgnn_b200.utils import make_mlp
    L, E = 16, 600
    hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
    torch.manual_seed(0)
    nets = torch.nn.ModuleDict({"edge_network": make_mlp(3 * L, 2 * L, L, 2, layer_norm=True, output_activation="Tanh"),
                                "node_network": make_mlp(2 * L, 2 * L, L, 3, layer_norm=True)})
    kaiming_init(nets)
    nodes, edges, graph = synth_edge_problem(E, L, seed=3)
    N = nodes.shape[0] - 1                     # odd node count: the last block is padded
    nodes = nodes[:N]
    graph = graph.clamp(max=N - 1)
    cot_n, cot_e = torch.randn(N, L, generator=torch.Generator().manual_seed(5)), torch.randn(E, L, generator=torch.Generator().manual_seed(6))

    # single-process oracle, two stacked cells with shared weights
    sd_ref = O.leaf_state({"c." + k: v for k, v in nets.state_dict().items()})
    n_ref, e_ref = nodes.clone().requires_grad_(True), edges.clone().requires_grad_(True)
    n2, e2 = n_ref, e_ref
    for _ in range(2):
        n2, e2 = O.interaction_cell(sd_ref, "c", hp, n2, e2, graph)
    ((n2 * cot_n).sum() + (e2 * cot_e).sum()).backward()

    # partitioned run
    part = partition_by_destination(graph, N, world, rank)
    assert int(part.dst_local.min()) >= 0 and int(part.dst_local.max()) < part.block
    sd = O.leaf_state({"c." + k: v for k, v in nets.state_dict().items()})
    node_fn = lambda x, agg: O.mlp_apply(sd, "c.node_network", torch.cat([x, agg], -1), 3, "GELU", "GELU", True) + x
    edge_fn = lambda x, e, g: O.edge_step(sd, "c.edge_network", hp, x, e, g)
    n_full = pad_rows(nodes, world * part.block).clone().requires_grad_(True)
    e_loc = edges[part.edge_ids].clone().requires_grad_(True)
    pn, pe = n_full, e_loc
    if carry:
        # the edge step also returns scatter_add(e', dst) over all node rows (as the CUDA edge kernel's fused segmented
        # reduce does); the driver hands its owned block to the next cell instead of running a segment sum there
        n_seg_calls = []
        def seg(rows, ids, n):
            n_seg_calls.append(1)
            return O.scatter_add(rows, ids, n)
        def edge_fn_agg(x, e, g):
            e2_ = edge_fn(x, e, g)
            return e2_, O.scatter_add(e2_, g[1], x.shape[0])
        agg = xo = None
        for _ in range(2):  # the owned block is carried too (its gradient stays a [B, L] block)
            pn, pe, agg, xo = partitioned_interaction_cell(part, pn, pe, node_fn, edge_fn_agg, seg, agg_owned=agg, return_agg=True,
                                                           x_owned=xo, return_owned=True)
        assert len(n_seg_calls) == 1 and agg is not None and agg.shape[0] == part.block
        assert torch.equal(xo, pn[part.node_lo:part.node_lo + part.block])
    else:
        for _ in range(2):
            pn, pe = partitioned_interaction_cell(part, pn, pe, node_fn, edge_fn, O.scatter_add)
    torch.testing.assert_close(pn[:N], n2.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(pe, e2.detach()[part.edge_ids], rtol=1e-5, atol=1e-6)
    # each rank back-propagates its share of the objective: owned nodes + owned edges
    own = slice(part.node_lo, part.node_hi)
    loss = (pn[own] * cot_n[own]).sum() + (pe * cot_e[part.edge_ids]).sum()
    loss.backward()
    torch.testing.assert_close(e_loc.grad, e_ref.grad[part.edge_ids], rtol=1e-4, atol=1e-5)
    # the input-node gradient is a sum over ranks (x is replicated): reduce and compare
    g = n_full.grad.clone()
    dist.all_reduce(g)
    torch.testing.assert_close(g[:N], n_ref.grad, rtol=1e-4, atol=1e-5)
    params = [v for v in sd.values() if v.requires_grad]
    allreduce_gradients(params, average=False)
    for k, v in sd.items():
        if v.requires_grad:
            torch.testing.assert_close(v.grad, sd_ref[k].grad, rtol=1e-4, atol=1e-5, msg=lambda m: f"{k}: {m}")


def _hier_partition_case(rank, world):
    """Two stacked HierarchicalGNNCells, destination-partitioned (hits / hit edges partitioned, supernode side replicated)
    against the single-process oracle: outputs, every input gradient, every weight gradient."""
    from hierarchicalgnn_b200.parallel import (pad_rows, partition_by_destination, partition_bipartite,
                                               partitioned_hierarchical_cell)
    from hierarchicalgnn_b200.synth import synth_edge_problem
    from oracle.reference_harness import kaiming_init
    from hierarchicalgnn_b200.utils import make_mlp
    L, E, S, ES = 16, 500, 9, 40
    hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
    torch.manual_seed(0)
    nets = torch.nn.ModuleDict({
        "edge_network": make_mlp(3 * L, 2 * L, L, 2, layer_norm=True, output_activation="Tanh"),
        "node_network": make_mlp(3 * L, 2 * L, L, 3, layer_norm=True),
        "supernode_network": make_mlp(3 * L, 2 * L, L, 3, layer_norm=True),
        "superedge_network": make_mlp(3 * L, 2 * L, L, 2, layer_norm=True, output_activation="Tanh")})
    kaiming_init(nets)
    nodes, edges, graph = synth_edge_problem(E, L, seed=3)
    N = nodes.shape[0] - 1
    nodes, graph = nodes[:N], graph.clamp(max=N - 1)
    g = torch.Generator().manual_seed(7)
    supernodes, superedges = torch.randn(S, L, generator=g), torch.randn(ES, L, generator=g)
    sgraph = torch.randint(0, S, (2, ES), generator=g)
    sweights = torch.rand(ES, 1, generator=g)
    bgraph = torch.stack([torch.arange(N).repeat(2), torch.randint(0, S, (2 * N,), generator=g)])  # two supernodes per hit
    bweights = torch.rand(2 * N, 1, generator=g)
    cots = [torch.randn(*shape, generator=g) for shape in ((N, L), (E, L), (S, L), (ES, L))]

    def leaves(*ts):
        return [t.clone().requires_grad_(True) for t in ts]

    # single-process oracle
    sd_ref = O.leaf_state({"c." + k: v for k, v in nets.state_dict().items()})
    r_n, r_e, r_s, r_se, r_bw, r_sw = leaves(nodes, edges, supernodes, superedges, bweights, sweights)
    a, b, c, d = r_n, r_e, r_s, r_se
    for _ in range(2):
        a, b, c, d = O.hierarchical_cell(sd_ref, "c", hp, a, b, c, d, graph, bgraph, r_bw, sgraph, r_sw)
    sum((o * ct).sum() for o, ct in zip((a, b, c, d), cots)).backward()

    # partitioned run
    part = partition_by_destination(graph, N, world, rank)
    bpart = partition_bipartite(bgraph, part)
    sd = O.leaf_state({"c." + k: v for k, v in nets.state_dict().items()})
    mlp = lambda name, x, n: O.mlp_apply(sd, "c." + name, x, n, "GELU", "GELU", True)
    wsum = lambda rows, w, gi, si, n: O.scatter_add(w * (rows if gi is None else rows[gi]), si, n)
    fns = dict(supernode=lambda sn, att, up: mlp("supernode_network", torch.cat([sn, att, up], -1), 3) + sn,
               node=lambda x, agg, down: mlp("node_network", torch.cat([x, agg, down], -1), 3) + x,
               superedge=lambda sn, se, sg: O.edge_step(sd, "c.superedge_network", hp, sn, se, sg),
               edge=lambda x, e, gr: O.edge_step(sd, "c.edge_network", hp, x, e, gr),
               weighted_sum=wsum, segment_sum=O.scatter_add)
    p_n, p_s, p_se, p_sw = leaves(pad_rows(nodes, world * part.block), supernodes, superedges, sweights)
    p_e, p_bw = leaves(edges[part.edge_ids], bweights[bpart.ids])
    st = dict(nodes=p_n, edges=p_e, supernodes=p_s, superedges=p_se, agg_owned=None, x_owned=None)
    for _ in range(2):
        st = partitioned_hierarchical_cell(part, bpart, st["nodes"], st["edges"], st["supernodes"], st["superedges"], p_bw,
                                           sgraph, p_sw, fns, x_owned=st["x_owned"])
    torch.testing.assert_close(st["nodes"][:N], a.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(st["edges"], b.detach()[part.edge_ids], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(st["supernodes"], c.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(st["superedges"], d.detach(), rtol=1e-5, atol=1e-6)
    # objective: owned hits + owned edges on this rank; the replicated supernode side is counted once (rank 0)
    own = slice(part.node_lo, part.node_hi)
    loss = (st["nodes"][own] * cots[0][own]).sum() + (st["edges"] * cots[1][part.edge_ids]).sum()
    rep = (st["supernodes"] * cots[2]).sum() + (st["superedges"] * cots[3]).sum()
    # every replica must back-propagate the same replicated objective to stay identical: all ranks add it
    (loss + rep).backward()
    torch.testing.assert_close(p_e.grad, r_e.grad[part.edge_ids], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(p_bw.grad, r_bw.grad[bpart.ids], rtol=1e-4, atol=1e-5)
    gn = p_n.grad.clone()
    dist.all_reduce(gn)
    torch.testing.assert_close(gn[:N], r_n.grad, rtol=1e-4, atol=1e-5)
    # replicated inputs: complete and identical on every rank, no reduction
    for got, want in ((p_s.grad, r_s.grad), (p_se.grad, r_se.grad), (p_sw.grad, r_sw.grad)):
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
    for k, v in sd.items():
        if not v.requires_grad:
            continue
        gk = v.grad.clone()
        if ".node_network." in k or ".edge_network." in k:   # per-rank partials
            dist.all_reduce(gk)
        # (fp32 sums over the edges in a different order: 5e-5 absolute on O(1) weight gradients)
        torch.testing.assert_close(gk, sd_ref[k].grad, rtol=1e-4, atol=5e-5, msg=lambda m: f"{k}: {m}")


def test_destination_partitioned_hierarchical_cells_match_single_process_world2():
    _run(_hier_partition_case)


class _BranchyTask(torch.nn.Module):
    """``side`` only enters the loss on rank 0 (a branch taken per rank): its gradient exists on one rank only."""

    def __init__(self):
        super().__init__()
        self.main = torch.nn.Linear(4, 1)
        self.side = torch.nn.Linear(4, 1)
        self.use_side = False

    def training_step(self, batch, batch_idx=0):
        y = self.main(batch)
        if self.use_side:
            y = y + self.side(batch)
        return y.square().mean()


def _one_sided_gradient_case(rank, world):
    """GradientBuckets with a parameter whose gradient arrives on ONE rank only: the other rank contributes zeros to its bucket,
    and after finish() every rank holds the same averaged gradient for it (not None, not stale)."""
    from hierarchicalgnn_b200.parallel import GradientBuckets
    torch.manual_seed(0)
    m = _BranchyTask()
    m.use_side = rank == 0
    gb = GradientBuckets(list(m.parameters()), bucket_bytes=8)  # one parameter per bucket
    data = torch.randn(world, 5, 4, generator=torch.Generator().manual_seed(2))
    for _ in range(2):  # the second pass starts from the first one's views
        gb.prepare()
        assert all(p.grad is None for p in m.parameters())
        m.training_step(data[rank]).backward()
        gb.finish()
        ref = _BranchyTask()
        ref.load_state_dict(m.state_dict())
        ref.use_side = True
        g0 = torch.autograd.grad(ref.training_step(data[0]), list(ref.parameters()))
        ref.use_side = False
        g1 = torch.autograd.grad(ref.training_step(data[1]), list(ref.main.parameters()))
        want = [(g0[0] + g1[0]) / 2, (g0[1] + g1[1]) / 2, g0[2] / 2, g0[3] / 2]
        for p, w in zip(m.parameters(), want):
            torch.testing.assert_close(p.grad, w, rtol=1e-5, atol=1e-7)
    gb.remove()


def test_gradient_buckets_one_sided_gradient_world2():
    _run(_one_sided_gradient_case)


def test_data_parallel_gradient_allreduce_world2():
    _run(_dp_case)


def test_data_parallel_trainer_overlapped_buckets_world2():
    _run(_dp_trainer_case)


def test_destination_partitioned_cells_match_single_process_world2():
    _run(_partition_case)


def _partition_case_carry(rank, world):
    _partition_case(rank, world, carry=True)


def test_destination_partitioned_cells_with_carried_aggregate_world2():
    _run(_partition_case_carry)


def test_partition_covers_every_edge_once():
    from hierarchicalgnn_b200.parallel import partition_by_destination
    g = torch.Generator().manual_seed(0)
    graph = torch.randint(0, 101, (2, 5000), generator=g)
    for world in (1, 2, 4, 8):
        parts = [partition_by_destination(graph, 101, world, r) for r in range(world)]
        ids = torch.cat([p.edge_ids for p in parts])
        assert torch.equal(ids.sort().values, torch.arange(5000))
        assert all(p.block == parts[0].block for p in parts) and sum(p.n_owned for p in parts) == 101
        for p in parts:
            assert torch.equal(p.graph, graph[:, p.edge_ids])
