"""Trace-build helper (make EXTRA=-DHGNN_TRACE): timeline of CTA 0 of the ping-pong forward kernel.
Usage: HGNN_FWD_PP=2 python profiles/pp_trace.py [E] [infer|train]"""
import sys, ctypes, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.gnn_utils import GraphPlans, InteractionGNNCell
from hierarchicalgnn_b200.synth import synth_edge_problem
from hierarchicalgnn_b200.training_utils import kaiming_init
E = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 16 * 128
train = len(sys.argv) > 2 and sys.argv[2] == "train"
L = 128
hp = dict(latent=L, hidden=2 * L, nb_edge_layer=2, nb_node_layer=3, layernorm=True, hidden_activation="GELU")
torch.manual_seed(0)
cell = InteractionGNNCell(hp); kaiming_init(cell); cell.cuda()
n, e, g = synth_edge_problem(E, L)
order = torch.argsort(g[1], stable=True)
g, e = g[:, order].contiguous(), e[order].contiguous()
n, e, g = n.cuda().requires_grad_(train), e.cuda().requires_grad_(train), g.cuda()
gp = GraphPlans(g, n.shape[0], n.shape[0], dst_sorted=True); gp.by_src; gp.by_dst
lib = ctypes.CDLL("hierarchicalgnn_b200/libhgnn_b200.so")
buf = (ctypes.c_ulonglong * (3 * 8192))()
def step():
    if train:
        return cell.edge_network.edge_step(n, e, gp.by_src, gp.by_dst)
    with torch.no_grad():
        return cell.edge_network.edge_step(n, e, gp.by_src, gp.by_dst)
step(); step()
lib.hgnn_tc_debug_trace(buf, 8192)  # drop the warm-up records
step()
cnt = lib.hgnn_tc_debug_trace(buf, 8192)
recs = sorted((buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(cnt))
t0 = recs[0][0] if recs else 0
names = {(1, 1): "mma  G1 first piece", (1, 2): "mma  G1 committed", (1, 3): "mma  G2 first piece", (1, 4): "mma  G2 committed",
         (2, 1): "gath tile begin", (2, 2): "gath tile end",
         (3, 0): "epi  wait D1", (3, 1): "epi  D1 ready -> EPI1", (3, 2): "epi  EPI1 done (A2 full)", (3, 3): "epi  D2 ready -> EPI2",
         (3, 4): "epi  EPI2 done -> store", (3, 5): "epi  store done -> agg", (3, 6): "epi  tile done"}
names.update({(1, 10): "mma   piece scheduled (it*16+p)", (1, 11): "mma   piece head (it*16+p)", (1, 12): "mma   A0 ready", (1, 13): "mma   W ready", (1, 14): "mma   issued"})
for b in range(6):
    names[(2, 10 + b)] = f"gath  slot for block {b} acquired"; names[(2, 20 + b)] = f"gath  block {b} published"
for t, code, it in recs:
    role, ev, grp = code >> 16, (code >> 8) & 0xff, code & 0xff
    print(f"{t - t0:9d}  it {it:3d} g{grp}  {names.get((role, ev), (role, ev))}")
