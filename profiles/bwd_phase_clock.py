"""Per-phase cycle breakdown of the tensor-core backward kernel (CTA 0), via the per-call hgnn_tc_edge_params.debug_phase_clock pointer."""
import sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops, _lib
from hierarchicalgnn_b200.gnn_utils import GraphPlans, InteractionGNNCell
from hierarchicalgnn_b200.synth import synth_edge_problem
from hierarchicalgnn_b200.training_utils import kaiming_init
L, E = 128, 1_000_000
hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
torch.manual_seed(0); cell = InteractionGNNCell(hp); kaiming_init(cell); cell.cuda()
n, e, g = synth_edge_problem(E, L)
order = torch.argsort(g[1], stable=True); g, e = g[:, order].contiguous(), e[order].contiguous()  # as bench.py / the models
n, e, g = n.cuda().requires_grad_(True), e.cuda().requires_grad_(True), g.cuda()
gp = GraphPlans(g, n.shape[0], n.shape[0]); gp.by_src; gp.by_dst
cot_e, cot_a = torch.randn_like(e), torch.randn(n.shape[0], L, device="cuda")
def step():
    e2, agg = cell.edge_network.edge_step(n, e, gp.by_src, gp.by_dst)
    torch.autograd.grad([e2, agg], [n, e] + list(cell.edge_network.parameters()), [cot_e, cot_a])
for _ in range(3): step()
clk = torch.zeros(16, dtype=torch.int64, device="cuda")
# the pointer travels in hgnn_tc_edge_params, which both kernels of a step receive: clock them in separate passes
e2, agg = cell.edge_network.edge_step(n, e, gp.by_src, gp.by_dst)
torch.cuda.synchronize()
ops.DEBUG_PHASE_CLOCK["ptr"] = clk.data_ptr()
torch.autograd.grad([e2, agg], [n, e] + list(cell.edge_network.parameters()), [cot_e, cot_a]); torch.cuda.synchronize()
ops.DEBUG_PHASE_CLOCK["ptr"] = None
c = clk.cpu().tolist()
names = ["setup", "LOAD(gout)", "EPI-B", "GEMM3", "EPI-C", "GEMM4", "EPI-D"]
tiles = (E + 127) // 128 // 296 + 1  # two CTAs per SM
tot = sum(c[:7])
for nm, v in zip(names, c[:7]):
    print(f"{nm:14s} {v / tiles:9.0f} cyc/tile  {100 * v / tot:5.1f}%")
print(f"total {tot / tiles:.0f} cycles/tile over ~{tiles} tiles")
# forward kernel
clk.zero_()
ops.DEBUG_PHASE_CLOCK["ptr"] = clk.data_ptr()
e2, agg = cell.edge_network.edge_step(n, e, gp.by_src, gp.by_dst); torch.cuda.synchronize()
ops.DEBUG_PHASE_CLOCK["ptr"] = None
c = clk.cpu().tolist()
names = ["setup", "GEMM1(gather)", "EPI1", "GEMM2", "EPI2", "store pass", "aggregate"]
tiles_f = (E + 127) // 128 // 296 + 1
tot = sum(c[:7])
print("forward kernel (2 CTAs / SM):")
for nm, v in zip(names, c[:7]):
    print(f"{nm:14s} {v / tiles_f:9.0f} cyc/tile  {100 * v / tot:5.1f}%")
print(f"total {tot / tiles_f:.0f} cycles/tile over ~{tiles_f} tiles")
