"""Where an end-to-end step's time goes: H2D copy time (events on the copy stream), compute time (events on the compute
stream), host wall per step. Mirrors bench.py's pipelined e2e loop."""
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
t0 = time.perf_counter()
nodes_p, edges_p, graph_p = nodes_h.pin_memory(), edges_h.pin_memory(), graph_h.pin_memory()
print("pin_memory %.1f ms; is_pinned %s" % ((time.perf_counter() - t0) * 1e3, edges_p.is_pinned()))
dev = torch.device("cuda")
copy_stream = torch.cuda.Stream()
slots = [tuple(torch.empty_like(t, device=dev) for t in (nodes_p, edges_p, graph_p)) for _ in range(2)]
out_host = torch.empty(2).pin_memory()
def upload(i):
    bufs = slots[i % 2]
    with torch.cuda.stream(copy_stream):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(copy_stream)
        for d, s in zip(bufs, (nodes_p, edges_p, graph_p)): d.copy_(s, non_blocking=True)
        b.record(copy_stream)
    return bufs, a, b
def run(k):
    rec = []
    cur = upload(0)
    for i in range(k):
        w0 = time.perf_counter(); st0 = torch.cuda.memory_stats()
        nxt = upload(i + 1) if i + 1 < k else None
        (n_b, e_b, g_d), a, b = cur
        torch.cuda.current_stream().wait_event(b)
        c0 = torch.cuda.Event(enable_timing=True); c1 = torch.cuda.Event(enable_timing=True)
        c0.record()
        n_d, e_d = n_b.detach().requires_grad_(True), e_b.detach().requires_grad_(True)
        gp = GraphPlans(g_d, N, N)
        e2, agg = net.edge_step(n_d, e_d, gp.by_src, gp.by_dst)
        grads = torch.autograd.grad([e2, agg], [n_d, e_d] + params, [cot_e, cot_a])
        out_host.copy_(torch.stack([e2.detach().sum() + agg.detach().sum(), grads[1].abs().sum()]), non_blocking=True)
        c1.record()
        w1 = time.perf_counter()
        torch.cuda.current_stream().synchronize()
        w2 = time.perf_counter()
        st1 = torch.cuda.memory_stats()
        rec.append((a.elapsed_time(b), c0.elapsed_time(c1), (w1 - w0) * 1e3, (w2 - w0) * 1e3,
                    st1.get('num_device_alloc', 0) - st0.get('num_device_alloc', 0), st1.get('num_device_free', 0) - st0.get('num_device_free', 0),
                    st1['reserved_bytes.all.current'] / 2**30, st1['allocated_bytes.all.current'] / 2**30, st1['allocated_bytes.all.peak'] / 2**30))
        cur = nxt
    return rec
run(3)
for h2d, comp, host_issue, wall, na, nf, res, al, pk in run(8):
    print(f"H2D {h2d:6.2f} ms | compute stream {comp:6.2f} ms | host issue {host_issue:6.2f} ms | wall {wall:6.2f} ms | cudaMalloc {na} cudaFree {nf} reserved {res:.2f} allocated {al:.2f} peak {pk:.2f} GiB")
