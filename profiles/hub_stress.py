import sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.gnn_utils import GraphPlans, InteractionGNNCell
from hierarchicalgnn_b200.synth import synth_edge_problem
from hierarchicalgnn_b200.training_utils import kaiming_init
L, E = 128, 1_000_000
hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
for pl in (False, True):
    torch.manual_seed(0); cell = InteractionGNNCell(hp); kaiming_init(cell); cell.cuda()
    n, e, g = synth_edge_problem(E, L, power_law=pl)
    order = torch.argsort(g[1], stable=True); g, e = g[:, order].contiguous(), e[order].contiguous()
    n, e, g = n.cuda().requires_grad_(True), e.cuda().requires_grad_(True), g.cuda()
    N = n.shape[0]
    gp = GraphPlans(g, N, N, dst_sorted=True); gp.by_src; gp.by_dst
    ce, ca = torch.randn_like(e), torch.randn(N, L, device="cuda")
    ps = list(cell.edge_network.parameters())
    def step():
        e2, agg = cell.edge_network.edge_step(n, e, gp.by_src, gp.by_dst)
        torch.autograd.grad([e2, agg], [n, e] + ps, [ce, ca])
    for _ in range(3): step()
    ops.PROFILE = {}
    step(); torch.cuda.synchronize()
    print("power-law" if pl else "uniform", "max degree", int(torch.bincount(g[1]).max()), {k: round(sum(a.elapsed_time(b) for a, b in v), 3) for k, v in ops.PROFILE.items()})
    ops.PROFILE = None
