"""Runs the fused edge step (forward in training mode = with stash, then backward) a few times on a mid-size problem:
the command line profiled under ncu for the per-kernel counters / source-level stall reasons of the edge kernels.
Usage: python profiles/edge_step_once.py [E] [iters] [fwd|infer|full] [nodes_per_edge]"""
import sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.gnn_utils import GraphPlans, InteractionGNNCell
from hierarchicalgnn_b200.synth import synth_edge_problem
from hierarchicalgnn_b200.training_utils import kaiming_init
E = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 16 * 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
fwd_only = len(sys.argv) > 3 and sys.argv[3] in ("fwd", "infer")
infer = len(sys.argv) > 3 and sys.argv[3] == "infer"
L = 128
hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
torch.manual_seed(0)
cell = InteractionGNNCell(hp); kaiming_init(cell); cell.cuda()
npe = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
n, e, g = synth_edge_problem(E, L, nodes_per_edge=npe)
order = torch.argsort(g[1], stable=True)
g, e = g[:, order].contiguous(), e[order].contiguous()
n, e, g = n.cuda().requires_grad_(True), e.cuda().requires_grad_(True), g.cuda()
gp = GraphPlans(g, n.shape[0], n.shape[0], dst_sorted=True); gp.by_src; gp.by_dst
params = list(cell.edge_network.parameters())
ce, ca = torch.randn_like(e), torch.randn_like(n)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def step():
    if infer:
        with torch.no_grad():
            return cell.edge_network.edge_step(n, e, gp.by_src, gp.by_dst)
    e2, agg = cell.edge_network.edge_step(n, e, gp.by_src, gp.by_dst)
    if not fwd_only:
        torch.autograd.grad([e2, agg], [n, e] + params, [ce, ca])
for i in range(3):
    step()
torch.cuda.synchronize(); a.record()
for i in range(iters):  # back to back: the queue stays full, host launch time is hidden
    step()
b.record(); torch.cuda.synchronize()
print(f"E={E} nodes/edge={npe} {a.elapsed_time(b) / iters:.3f} ms per iteration ({iters} iterations)")
