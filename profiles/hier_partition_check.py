"""Multi-GPU check of the destination-partitioned HierarchicalGNNCell (CUDA kernels + peer-memory collectives) against the
same cell run un-partitioned on one GPU; prints max differences and the step times.
torchrun --nproc-per-node N --master-addr 127.0.0.1 --master-port 29515 profiles/hier_partition_check.py [E]"""
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.gnn_utils import HierarchicalGNNCell, GraphPlans
from hierarchicalgnn_b200.parallel import (SymmetricRows, cuda_hier_cell_callables, pad_rows, partition_by_destination,
                                           partition_bipartite, partitioned_hierarchical_cell)
from hierarchicalgnn_b200.synth import synth_edge_problem
from hierarchicalgnn_b200.training_utils import kaiming_init

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
L = 128
E = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
torch.manual_seed(0)
cells = [HierarchicalGNNCell(hp) for _ in range(2)]
for c in cells:
    kaiming_init(c); c.to(dev)
nodes_h, edges_h, graph_h = synth_edge_problem(E, L, seed=2000, nodes_per_edge=0.04)
N = nodes_h.shape[0]
g = torch.Generator().manual_seed(5)
S, ES = max(N // 10, 8), max(N, 64)
sn_h, se_h = torch.randn(S, L, generator=g), torch.randn(ES, L, generator=g)
sg_h = torch.randint(0, S, (2, ES), generator=g)
sw_h = torch.rand(ES, 1, generator=g)
bg_h = torch.stack([torch.arange(N).repeat(3), torch.randint(0, S, (3 * N,), generator=g)])
bw_h = torch.rand(3 * N, 1, generator=g) / 3
cots = [torch.randn(*s, generator=g) for s in ((N, L), (E, L), (S, L), (ES, L))]
order = torch.argsort(graph_h[1], stable=True)
graph_h, edges_h, cots[1] = graph_h[:, order].contiguous(), edges_h[order].contiguous(), cots[1][order].contiguous()


def leaf(t):
    return t.to(dev).clone().requires_grad_(True)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


# ---- un-partitioned (every rank, redundantly) ----
r = dict(n=leaf(nodes_h), e=leaf(edges_h), s=leaf(sn_h), se=leaf(se_h), bw=leaf(bw_h), sw=leaf(sw_h))
graph, bg, sg = graph_h.to(dev), bg_h.to(dev), sg_h.to(dev)
gp, bp, sp = GraphPlans(graph, N, N, dst_sorted=True), GraphPlans(bg, N, S), GraphPlans(sg, S, S)
cd = [c.to(dev) for c in cots]
params = [p for c in cells for p in c.parameters()]


def solo():
    a, b, c, d = r["n"], r["e"], r["s"], r["se"]
    for i, cell in enumerate(cells):
        a, b, c, d = cell(a, b, c, d, gp, bp, r["bw"], sp, r["sw"], skip_edge_updates=(i == 1))
    outs = (a, b, c, d)
    grads = torch.autograd.grad(sum((o * ct).sum() for o, ct in zip(outs, cd)), list(r.values()) + params, allow_unused=True)
    return outs, grads


o_ref, g_ref = solo()
t1 = timed(solo)

# ---- partitioned ----
part = partition_by_destination(graph_h, N, world, rank)
bpart = partition_bipartite(bg_h, part)
own = slice(part.node_lo, part.node_hi)
p = dict(n=leaf(pad_rows(nodes_h, world * part.block)), e=leaf(edges_h[part.edge_ids]), s=leaf(sn_h), se=leaf(se_h),
         bw=leaf(bw_h[bpart.ids]), sw=leaf(sw_h))
for k in ("graph", "dst_local", "edge_ids"):
    setattr(part, k, getattr(part, k).to(dev))
bpart.ids, bpart.node_local, bpart.supernode = bpart.ids.to(dev), bpart.node_local.to(dev), bpart.supernode.to(dev)
fns = [cuda_hier_cell_callables(c, fuse_aggregate=(i == 0)) for i, c in enumerate(cells)]
sr = SymmetricRows(part.block, L, dev, slots=2) if world > 1 else None
cn, ce = cd[0][own], cd[1][part.edge_ids]


def parted():
    st = dict(nodes=p["n"], edges=p["e"], supernodes=p["s"], superedges=p["se"], agg_owned=None, x_owned=None)
    for i, f in enumerate(fns):
        st = partitioned_hierarchical_cell(part, bpart, st["nodes"], st["edges"], st["supernodes"], st["superedges"], p["bw"],
                                           sg, p["sw"], f, symmetric=sr, agg_owned=st["agg_owned"], x_owned=st["x_owned"],
                                           slot=i, skip_edge_updates=(i == 1))
    loss = (st["nodes"][own] * cn).sum() + (st["edges"] * ce).sum() + (st["supernodes"] * cd[2]).sum() + (st["superedges"] * cd[3]).sum()
    grads = torch.autograd.grad(loss, list(p.values()) + params, allow_unused=True)
    flat = torch.cat([x.reshape(-1) for x in grads[6:] if x is not None])
    dist.all_reduce(flat)  # (a real driver sums only the partitioned networks' gradients; here it is the step-end collective)
    return st, grads


st, g_p = parted()
torch.cuda.synchronize()


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


msgs = [f"nodes {rel(st['nodes'][own], o_ref[0][own]):.2e}", f"edges {rel(st['edges'], o_ref[1][part.edge_ids]):.2e}",
        f"supernodes {rel(st['supernodes'], o_ref[2]):.2e}", f"superedges {rel(st['superedges'], o_ref[3]):.2e}",
        f"d_edges {rel(g_p[1], g_ref[1][part.edge_ids]):.2e}", f"d_supernodes {rel(g_p[2], g_ref[2]):.2e}",
        f"d_bweights {rel(g_p[4], g_ref[4][bpart.ids]):.2e}"]
gn = g_p[0].clone(); dist.all_reduce(gn)
msgs.append(f"d_nodes {rel(gn[:N], g_ref[0]):.2e}")
tp = timed(parted)
if rank == 0:
    print("relative Frobenius differences, partitioned vs one GPU (both on the bf16 tensor-core path):", ", ".join(msgs))
    print(f"E={E} N={N} S={S}: one GPU {t1:.3f} ms, {world} GPUs {tp:.3f} ms -> {t1 / tp:.2f}x  (2 HierarchicalGNNCells fwd+bwd)")
dist.destroy_process_group()
