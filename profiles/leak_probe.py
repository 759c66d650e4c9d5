"""Does an end-to-end style step (fresh input tensors + fresh plans per step) leave device memory behind?"""
import gc, sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.gnn_utils import GraphPlans, InteractionGNNCell
from hierarchicalgnn_b200.synth import synth_edge_problem
from hierarchicalgnn_b200.training_utils import kaiming_init
L, E = 128, 1_000_000
hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
torch.manual_seed(0); cell = InteractionGNNCell(hp); kaiming_init(cell); cell.cuda(); net = cell.edge_network
params = list(net.parameters())
nodes_h, edges_h, graph_h = synth_edge_problem(E, L, seed=42)
N = nodes_h.shape[0]
order = torch.argsort(graph_h[1], stable=True); graph_h, edges_h = graph_h[:, order].contiguous(), edges_h[order].contiguous()
cot_e, cot_a = torch.randn(E, L).cuda(), torch.randn(N, L).cuda()
nb, eb, gb = nodes_h.cuda(), edges_h.cuda(), graph_h.cuda()
def one(mode):
    if mode == "same_graph_tensor":
        g_d = gb
    else:
        gb.add_(0)  # bumps the version: plan_for sees a new graph, as after an upload into the same slot
        g_d = gb
    n_d, e_d = nb.detach().requires_grad_(True), eb.detach().requires_grad_(True)
    gp = GraphPlans(g_d, N, N)
    e2, agg = net.edge_step(n_d, e_d, gp.by_src, gp.by_dst)
    grads = torch.autograd.grad([e2, agg], [n_d, e_d] + params, [cot_e, cot_a])
    return float((e2.sum() + agg.sum()).item())
for mode in ("same_graph_tensor", "new_graph_version", "new_graph_version+gc", "new_graph_version+clear_plan_cache"):
    torch.cuda.synchronize(); gc.collect(); torch.cuda.empty_cache()
    base = torch.cuda.memory_allocated()
    line = []
    for i in range(6):
        one(mode.split("+")[0])
        if mode.endswith("+gc"): gc.collect()
        if mode.endswith("clear_plan_cache"): ops.clear_plan_cache()
        torch.cuda.synchronize()
        line.append("%.2f" % ((torch.cuda.memory_allocated() - base) / 2**30))
    print(f"{mode:40s} allocated growth after each step (GiB): {' '.join(line)}   plan cache entries {len(ops._PLAN_CACHE)}")
