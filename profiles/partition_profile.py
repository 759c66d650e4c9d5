"""torch.profiler view (rank 0) of one destination-partitioned InteractionGNNCell fwd+bwd step (bench.py --mode partition)
under torchrun: which kernels / collectives make up the step at N ranks."""
import os, sys, torch
sys.path.insert(0, '.')
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from hierarchicalgnn_b200.gnn_utils import InteractionGNNCell
from hierarchicalgnn_b200.parallel import SymmetricRows, cuda_cell_callables, pad_rows, partition_by_destination, partitioned_interaction_cell
from hierarchicalgnn_b200.synth import synth_edge_problem
from hierarchicalgnn_b200.training_utils import kaiming_init

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
L, E = 128, 3_000_000
hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
NC = int(os.environ.get("HGNN_PART_CELLS", "2"))
torch.manual_seed(0)
cells = []
for _ in range(NC):
    c = InteractionGNNCell(hp); kaiming_init(c); cells.append(c.to(dev))
nodes_h, edges_h, graph_h = synth_edge_problem(E, L, seed=2000, nodes_per_edge=0.04)
N = nodes_h.shape[0]
part = partition_by_destination(graph_h, N, world, rank)
g = torch.Generator().manual_seed(11)
cot_n, cot_e = torch.randn(N, L, generator=g), torch.randn(E, L, generator=g)
own = slice(part.node_lo, part.node_hi)
cot_n_d, cot_e_d = cot_n[own].to(dev), cot_e[part.edge_ids].to(dev)
nodes = pad_rows(nodes_h, world * part.block).to(dev).requires_grad_(True)
e_loc = edges_h[part.edge_ids].to(dev).requires_grad_(True)
part.graph, part.dst_local, part.edge_ids = part.graph.to(dev), part.dst_local.to(dev), part.edge_ids.to(dev)
calls = [cuda_cell_callables(c, fuse_aggregate=(i + 1 < NC)) for i, c in enumerate(cells)]
params = [p for c in cells for p in c.parameters()]
sr = SymmetricRows(part.block, L, dev, slots=NC) if (world > 1 and os.environ.get("HGNN_PART_SYMM", "1") != "0") else None

def step():
    x, e, agg, xo = nodes, e_loc, None, None
    for ci, (node_fn, edge_fn, seg) in enumerate(calls):
        x, e, agg, xo = partitioned_interaction_cell(part, x, e, node_fn, edge_fn, seg, symmetric=sr, agg_owned=agg, return_agg=True,
                                                     x_owned=xo, slot=ci, return_owned=True)
    grads = torch.autograd.grad([x[own], e], [nodes, e_loc] + params, [cot_n_d, cot_e_d])
    if world > 1:
        flat = torch.cat([t.reshape(-1) for t in grads[2:]]); dist.all_reduce(flat)
        blk = grads[0][part.node_lo:part.node_lo + part.block].contiguous()
        if sr is not None:
            sr.all_gather(blk)
        else:
            gfull = torch.empty_like(grads[0]); dist.all_gather_into_tensor(gfull, blk)

for _ in range(5): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): step()
b.record(); torch.cuda.synchronize()
if rank == 0: print(f"world {world}: {a.elapsed_time(b) / 10:.3f} ms/step, local edges {int(part.edge_ids.numel())}")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
if rank == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=55))
if world > 1: dist.destroy_process_group()
