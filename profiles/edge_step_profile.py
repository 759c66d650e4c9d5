"""torch.profiler kernel table of the edge step (forward + backward) at a given latent: python profiles/edge_step_profile.py [L] [E]"""
import sys, torch
sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.gnn_utils import GraphPlans, InteractionGNNCell
from hierarchicalgnn_b200.synth import synth_edge_problem
from hierarchicalgnn_b200.training_utils import kaiming_init
L = int(sys.argv[1]) if len(sys.argv) > 1 else 256
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
torch.manual_seed(0)
cell = InteractionGNNCell(hp); kaiming_init(cell); cell.cuda()
n, e, g = synth_edge_problem(E, L)
order = torch.argsort(g[1], stable=True); g, e = g[:, order].contiguous(), e[order].contiguous()
n, e, g = n.cuda().requires_grad_(True), e.cuda().requires_grad_(True), g.cuda()
N = n.shape[0]
gp = GraphPlans(g, N, N, dst_sorted=True); gp.by_src; gp.by_dst
params = list(cell.edge_network.parameters())
ce, ca = torch.randn_like(e), torch.randn_like(n)
def step():
    e2, agg = cell.edge_network.edge_step(n, e, gp.by_src, gp.by_dst)
    if agg is None:
        agg = ops.scatter_add(e2, g[1], dim_size=N, plan=gp.by_dst)
    torch.autograd.grad([e2, agg], [n, e] + params, [ce, ca])
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as p:
    step(); torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=80))
