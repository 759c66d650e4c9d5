"""ORACLE / TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.pt by running the UNMODIFIED reference modules
(/root/reference/Modules, third-party CUDA packages stubbed by oracle/stubs) on
CPU. Run in the build container:  python oracle/make_golden.py
The fixtures pin oracle/hgnn_oracle.py and are what the GPU parity tests
compare the CUDA path against.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_harness as rh  # noqa: E402
from oracle.seeded_state import seeded_init  # noqa: E402
from hierarchicalgnn_b200.synth import synth_event  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SMALL = dict(latent=32, hidden="ratio", hidden_ratio=2)


def grads_of(model):
    return {k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()}


def clone_sd(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def rand_graph(n_src, n_dst, e, g):
    return torch.stack([torch.randint(0, n_src, (e,), generator=g), torch.randint(0, n_dst, (e,), generator=g)], 0)


def golden_cells(C):
    g = torch.Generator().manual_seed(7)
    out = {}
    for tag, over in {"ln_gelu": dict(), "noln_relu": dict(layernorm=False, hidden_activation="ReLU", nb_edge_layer=3),
                      "silu": dict(hidden_activation="SiLU", nb_node_layer=2)}.items():
        hp = rh.load_yaml_hparams("EC", **SMALL, **over)
        torch.manual_seed(3)
        cell = C["gnn_utils"].InteractionGNNCell(hp)
        rh.kaiming_init(cell)
        N, E, L = 57, 301, hp["latent"]
        graph = rand_graph(N - 4, N - 4, E, g)  # last 4 nodes isolated; duplicates + self loops likely
        graph[:, :5] = graph[:, 5:10]           # force duplicate edges
        graph[1, 10:40] = 3                     # a hub node
        nodes = torch.randn(N, L, generator=g).requires_grad_(True)
        edges = torch.randn(E, L, generator=g).requires_grad_(True)
        n2, e2 = cell(nodes, edges, graph)
        wn, we = torch.randn(N, L, generator=g), torch.randn(E, L, generator=g)
        ((n2 * wn).sum() + (e2 * we).sum()).backward()
        out[tag] = dict(hparams=hp, state=clone_sd(cell), nodes=nodes.detach(), edges=edges.detach(), graph=graph,
                        out_nodes=n2.detach(), out_edges=e2.detach(), w_nodes=wn, w_edges=we,
                        grad_nodes=nodes.grad.clone(), grad_edges=edges.grad.clone(), grad_params=grads_of(cell))
    torch.save(out, os.path.join(OUT, "cell_interaction.pt"))

    hp = rh.load_yaml_hparams("BC", **SMALL)
    torch.manual_seed(4)
    cell = C["gnn_utils"].HierarchicalGNNCell(hp)
    rh.kaiming_init(cell)
    N, E, S, L = 83, 400, 9, hp["latent"]
    graph = rand_graph(N, N, E, g)
    bg = torch.stack([torch.arange(N).repeat_interleave(3), torch.randint(0, S, (3 * N,), generator=g)], 0)
    sg = torch.unique(rand_graph(S, S, 40, g), dim=1)
    t = lambda *s: torch.randn(*s, generator=g).requires_grad_(True)
    nodes, edges, sn, se = t(N, L), t(E, L), t(S, L), t(sg.shape[1], L)
    bw = torch.rand(bg.shape[1], 1, generator=g).requires_grad_(True)
    sw = torch.rand(sg.shape[1], 1, generator=g).requires_grad_(True)
    outs = cell(nodes, edges, sn, se, graph, bg, bw, sg, sw)
    ws = [torch.randn(o.shape, generator=g) for o in outs]
    sum((o * w).sum() for o, w in zip(outs, ws)).backward()
    torch.save(dict(hparams=hp, state=clone_sd(cell), nodes=nodes.detach(), edges=edges.detach(), supernodes=sn.detach(),
                    superedges=se.detach(), graph=graph, bipartite_graph=bg, bipartite_weights=bw.detach(),
                    super_graph=sg, super_weights=sw.detach(), outs=[o.detach() for o in outs], ws=ws,
                    grads=dict(nodes=nodes.grad, edges=edges.grad, supernodes=sn.grad, superedges=se.grad,
                               bipartite_weights=bw.grad, super_weights=sw.grad), grad_params=grads_of(cell)),
               os.path.join(OUT, "cell_hierarchical.pt"))


def golden_dgc(C):
    g = torch.Generator().manual_seed(11)
    hp = rh.load_yaml_hparams("BC", **SMALL)
    out = {}
    P1, P2, D = 150, 23, 8
    centres = torch.nn.functional.normalize(torch.randn(P2, D, generator=g))
    for tag, (weighting, sym, k, train) in {"bip_train": ("exp", False, 5, True), "bip_eval": ("exp", False, 5, False),
                                            "sup_train": ("sigmoid", True, 10, True),
                                            "sup_eval": ("sigmoid", True, 10, False)}.items():
        torch.manual_seed(5)
        m = C["gnn_utils"].DynamicGraphConstruction(weighting, hp)
        m.knn_radius.fill_(0.9 if not sym else 1.25)
        m.weight_normalization.running_mean.fill_(0.3)
        m.weight_normalization.running_var.fill_(0.7)
        m.weight_normalization.weight.data.fill_(1.3)
        m.weight_normalization.bias.data.fill_(-0.2)
        m.train(train)
        before = clone_sd(m)
        if sym:
            src = centres.clone().requires_grad_(True)
            res = m(src, src, sym=True, norm=True, k=k, logits=True)
        else:
            src = torch.nn.functional.normalize(centres[torch.randint(0, P2, (P1,), generator=g)]
                                                + 0.35 * torch.randn(P1, D, generator=g)).requires_grad_(True)
            dst = centres.clone().requires_grad_(True)
            res = m(src, dst, sym=False, norm=True, k=k, logits=True)
        graph, w, logits = res
        wt = torch.randn(w.shape, generator=g)
        (w * wt).sum().backward()
        rec = dict(weighting=weighting, sym=sym, k=k, training=train, state_before=before, state_after=clone_sd(m),
                   src=src.detach(), graph=graph, weights=w.detach(), logits=logits.detach(), wt=wt,
                   grad_src=src.grad.clone())
        if not sym:
            rec.update(dst=dst.detach(), grad_dst=dst.grad.clone())
        out[tag] = rec
    torch.save(out, os.path.join(OUT, "dynamic_graph.pt"))


def golden_ec(C):
    out = {}
    for tag, over in {"default": dict(n_interaction_graph_iters=3),
                      "shared_noln": dict(n_interaction_graph_iters=2, share_weight=True, layernorm=False,
                                          hidden_output_activation="Tanh")}.items():
        hp = rh.load_yaml_hparams("EC", **SMALL, **over)
        torch.manual_seed(0)
        m = C["EC_InteractionGNN"](hp)
        rh.kaiming_init(m)
        ev = synth_event(24, 6, 0.1, 2.0, seed=1001)
        x = ev.x.clone()
        scores = m(x, ev.edge_index)
        loss = torch.nn.functional.binary_cross_entropy(scores, ev.y_pid.float())
        loss.backward()
        out[tag] = dict(hparams=hp, state=clone_sd(m), keys=list(m.state_dict().keys()), x=ev.x, graph=ev.edge_index,
                        y=ev.y_pid, scores=scores.detach(), loss=loss.detach(), grad_x=x.grad.clone(),
                        grad_params=grads_of(m))
    torch.save(out, os.path.join(OUT, "ec_model.pt"))


def golden_bc(C):
    hp = rh.load_yaml_hparams("BC", **SMALL, n_interaction_graph_iters=2, n_hierarchical_graph_iters=2)
    torch.manual_seed(0)
    m = C["BC_HierarchicalGNN_GMM"](hp)
    rh.kaiming_init(m)
    m.hgnn_block.GMM_model.set_params(random_state=0)
    ev = synth_event(40, 8, 0.05, 2.0, seed=1002)
    rec = {}
    orig = m.hgnn_block.clustering

    def spy(x, emb, graph):
        c = orig(x, emb, graph)
        rec["clusters"] = c.clone()
        return c
    m.hgnn_block.clustering = spy
    out = {}
    for mode in ("train", "eval"):
        m.train(mode == "train")
        m.zero_grad()
        before = clone_sd(m)
        x = ev.x.clone()
        bg, scores, emb = m(x, ev.edge_index)
        g = torch.Generator().manual_seed(5)
        ws, we = torch.randn(scores.shape, generator=g), torch.randn(emb.shape, generator=g)
        ((scores * ws).sum() + (emb * we).sum()).backward()
        after = {k: v for k, v in clone_sd(m).items() if not torch.equal(v, before[k])}
        out[mode] = dict(state_before=before if mode == "train" else None, state_after=after, clusters=rec["clusters"], bipartite_graph=bg,
                         scores=scores.detach(), embeddings=emb.detach(), ws=ws, we=we, grad_x=x.grad.clone(),
                         grad_params=grads_of(m), logged={k: float(v) for k, v in m.logged.items()})
    out.update(hparams=hp, keys=list(m.state_dict().keys()), x=ev.x, graph=ev.edge_index, pid=ev.pid)
    torch.save(out, os.path.join(OUT, "bc_model.pt"))


def golden_mlp_layout(C):
    mk = C["utils"].make_mlp
    out = {}
    for n in (1, 2, 3, 4):
        for ln in (False, True):
            for oa in (None, "Tanh"):
                torch.manual_seed(n)
                net = mk(6, 10, 4, n, hidden_activation="GELU", output_activation=oa, layer_norm=ln)
                x = torch.randn(5, 6)
                out[(n, ln, oa)] = dict(keys=list(net.state_dict().keys()), state=clone_sd(net), x=x, y=net(x).detach())
    torch.save(out, os.path.join(OUT, "make_mlp.pt"))


def bf16_grads(model):
    """Parameter gradients stored as bf16 (2^-9 relative; the tensor-core tests state 1.5e-2 rel-Frobenius)."""
    return {k: (p.grad.to(torch.bfloat16) if p.grad is not None else None) for k, p in model.named_parameters()}


def golden_latent128(C):
    """Latent-128 fixtures for the DEFAULT (tensor-core) path: the shapes the tcgen05 kernels accept. State dicts are
    not stored: both sides call oracle.seeded_state.seeded_init(module, seed) and compare checksums."""
    BIG = dict(latent=128, hidden="ratio", hidden_ratio=2)
    g = torch.Generator().manual_seed(17)
    out = {}
    # --- InteractionGNNCell (gnn_utils.py:18-71)
    hp = rh.load_yaml_hparams("EC", **BIG)
    cell = C["gnn_utils"].InteractionGNNCell(hp)
    ck = seeded_init(cell, 101)
    N, E, L = 211, 1500, 128
    graph = rand_graph(N - 6, N - 6, E, g)      # last 6 nodes isolated
    graph[:, :7] = graph[:, 7:14]               # duplicate edges
    graph[1, 20:180] = 5                        # a hub spanning more than one 128-row tile
    nodes = torch.randn(N, L, generator=g).requires_grad_(True)
    edges = torch.randn(E, L, generator=g).requires_grad_(True)
    n2, e2 = cell(nodes, edges, graph)
    wn, we = torch.randn(N, L, generator=g), torch.randn(E, L, generator=g)
    ((n2 * wn).sum() + (e2 * we).sum()).backward()
    out["cell"] = dict(hparams=hp, seed=101, checksum=ck, nodes=nodes.detach(), edges=edges.detach(), graph=graph,
                       out_nodes=n2.detach(), out_edges=e2.detach(), w_nodes=wn, w_edges=we,
                       grad_nodes=nodes.grad.clone(), grad_edges=edges.grad.clone(), grad_params=bf16_grads(cell))
    # --- HierarchicalGNNCell (gnn_utils.py:74-169)
    hp = rh.load_yaml_hparams("BC", **BIG)
    cell = C["gnn_utils"].HierarchicalGNNCell(hp)
    ck = seeded_init(cell, 102)
    N, E, S = 300, 1300, 24
    graph = rand_graph(N, N, E, g)
    bg = torch.stack([torch.arange(N).repeat_interleave(3), torch.randint(0, S, (3 * N,), generator=g)], 0)
    sg = torch.unique(rand_graph(S, S, 150, g), dim=1)
    t = lambda *s: torch.randn(*s, generator=g).requires_grad_(True)
    nodes, edges, sn, se = t(N, L), t(E, L), t(S, L), t(sg.shape[1], L)
    bw = torch.rand(bg.shape[1], 1, generator=g).requires_grad_(True)
    sw = torch.rand(sg.shape[1], 1, generator=g).requires_grad_(True)
    outs = cell(nodes, edges, sn, se, graph, bg, bw, sg, sw)
    ws = [torch.randn(o.shape, generator=g) for o in outs]
    sum((o * w).sum() for o, w in zip(outs, ws)).backward()
    out["hcell"] = dict(hparams=hp, seed=102, checksum=ck, nodes=nodes.detach(), edges=edges.detach(),
                        supernodes=sn.detach(), superedges=se.detach(), graph=graph, bipartite_graph=bg,
                        bipartite_weights=bw.detach(), super_graph=sg, super_weights=sw.detach(),
                        outs=[o.detach() for o in outs], ws=ws,
                        grads=dict(nodes=nodes.grad, edges=edges.grad, supernodes=sn.grad, superedges=se.grad,
                                   bipartite_weights=bw.grad, super_weights=sw.grad), grad_params=bf16_grads(cell))
    # --- EC_InteractionGNN, 2 cells (EC/Models/IN.py:118-128): model-level gradients
    hp = rh.load_yaml_hparams("EC", **BIG, n_interaction_graph_iters=2)
    m = C["EC_InteractionGNN"](hp)
    ck = seeded_init(m, 103)
    ev = synth_event(60, 8, 0.05, 3.0, seed=1003)
    x = ev.x.clone()
    scores = m(x, ev.edge_index)
    loss = torch.nn.functional.binary_cross_entropy(scores, ev.y_pid.float())
    loss.backward()
    out["ec"] = dict(hparams=hp, seed=103, checksum=ck, x=ev.x, graph=ev.edge_index, y=ev.y_pid, scores=scores.detach(),
                     loss=loss.detach(), grad_x=x.grad.clone(), grad_params=bf16_grads(m))
    # --- BC_HierarchicalGNN_GMM, 1 + 2 cells (BC/Models/HGNN_GMM.py:323-346), clusters recorded for injection
    hp = rh.load_yaml_hparams("BC", **BIG, n_interaction_graph_iters=1, n_hierarchical_graph_iters=2)
    m = C["BC_HierarchicalGNN_GMM"](hp)
    ck = seeded_init(m, 104)
    m.hgnn_block.GMM_model.set_params(random_state=0)
    ev = synth_event(60, 8, 0.05, 2.0, seed=1004)
    rec = {}
    orig = m.hgnn_block.clustering

    def spy(x, emb, graph):
        c = orig(x, emb, graph)
        rec["clusters"] = c.clone()
        return c
    m.hgnn_block.clustering = spy
    sgc = m.hgnn_block.super_graph_construction
    sgc_fwd = sgc.forward

    def spy_sg(*a, **k):
        r = sgc_fwd(*a, **k)
        rec["super_graph"] = r[0].clone()
        return r
    sgc.forward = spy_sg
    m.train()
    before = {k: v.detach().clone() for k, v in m.state_dict().items() if "running_" in k or "knn_radius" in k}
    x = ev.x.clone()
    bgr, scores, emb = m(x, ev.edge_index)
    gg = torch.Generator().manual_seed(5)
    wsc, wem = torch.randn(scores.shape, generator=gg), torch.randn(emb.shape, generator=gg)
    ((scores * wsc).sum() + (emb * wem).sum()).backward()
    out["bc"] = dict(hparams=hp, seed=104, checksum=ck, x=ev.x, graph=ev.edge_index, clusters=rec["clusters"],
                     super_graph=rec["super_graph"], bipartite_graph=bgr, scores=scores.detach(), embeddings=emb.detach(),
                     ws=wsc, we=wem, grad_x=x.grad.clone(), grad_params=bf16_grads(m))
    torch.save(out, os.path.join(OUT, "latent128.pt"))


def main():
    os.makedirs(OUT, exist_ok=True)
    C = rh.reference_classes()
    if "--only-latent128" in sys.argv:
        golden_latent128(C)
        print("latent128.pt", os.path.getsize(os.path.join(OUT, "latent128.pt")))
        return
    golden_latent128(C)
    golden_mlp_layout(C)
    golden_cells(C)
    golden_dgc(C)
    golden_ec(C)
    golden_bc(C)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
