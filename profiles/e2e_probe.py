import sys, time, torch
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
nodes_p, edges_p, graph_p = nodes_h.pin_memory(), edges_h.pin_memory(), graph_h.pin_memory()
dev = torch.device("cuda")
def T(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def up():
    return nodes_p.to(dev, non_blocking=True), edges_p.to(dev, non_blocking=True), graph_p.to(dev, non_blocking=True)
print("upload only            %.2f ms" % T(up))
n_d, e_d, g_d = up(); n_d.requires_grad_(True); e_d.requires_grad_(True)
print("plans only             %.2f ms" % T(lambda: (lambda p: (p.by_src, p.by_dst))(GraphPlans(g_d, N, N))))
gp = GraphPlans(g_d, N, N); gp.by_src; gp.by_dst
def step(gp):
    e2, agg = net.edge_step(n_d, e_d, gp.by_src, gp.by_dst)
    return e2, agg, torch.autograd.grad([e2, agg], [n_d, e_d] + params, [cot_e, cot_a])
print("step only              %.2f ms" % T(lambda: step(gp)))
def full():
    gp = GraphPlans(g_d, N, N)
    e2, agg, grads = step(gp)
    m = torch.stack([e2.sum() + agg.sum(), grads[1].abs().sum()])
    return m.cpu()
print("plans+step+metric+D2H  %.2f ms" % T(full))
