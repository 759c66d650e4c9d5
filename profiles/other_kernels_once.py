"""One call of each HBM- / SIMT-bound kernel family at BASELINE sizes, for `ncu --set full` captures of the kernels that are not
the fused edge step: segment reduce, row gather, brute-force kNN, the row-layer backward and the narrow (fan-in 3) first layer.
Usage: python profiles/other_kernels_once.py  (under ncu: -k regex:'k_segment_reduce|k_gather_rows|k_knn_radius|k_tc_row_bwd|k_narrow_in')"""
import sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops
from hierarchicalgnn_b200.utils import make_mlp
DEV = 'cuda'
g = torch.Generator().manual_seed(0)
E, N, L = 1_000_000, 100_000, 128
src = torch.randn(E, L, generator=g).to(DEV)
idx = torch.randint(0, N, (E,), generator=g).to(DEV)
x = torch.randn(N, L, generator=g).to(DEV)
plan_sorted, plan_rand = ops.plan_for(idx.sort().values, N), ops.plan_for(idx, N)
for _ in range(2):
    ops.segment_reduce_raw(src, plan_sorted)          # destination-sorted rows (the models' layout)
    ops.segment_reduce_raw(src, plan_rand)            # arbitrary order (rows followed through perm)
    ops.gather_rows_raw(x, plan_rand.keys32, None, E)
# kNN at the full pile-up shapes (hits -> supernodes, k = 5; supernode graph, k = 10) and the 1 GeV supergraph
for nq, nr, k in [(120_000, 12_000, 5), (12_000, 12_000, 10), (1_200, 1_200, 10)]:
    q = torch.nn.functional.normalize(torch.randn(nq, 8, generator=g)).to(DEV)
    r = torch.nn.functional.normalize(torch.randn(nr, 8, generator=g)).to(DEV)
    for _ in range(2):
        ops.knn_radius(q, r, k, 1.5)
# node network (3 row layers on [x | agg]) forward + backward at N = 100 k rows; encoder (fan-in 3 first layer) at 1 M rows
old = ops.set_precision("auto")
net = make_mlp(2 * L, 2 * L, L, 3, layer_norm=True, output_activation="GELU", hidden_activation="GELU").to(DEV)
a, b = x.clone().requires_grad_(True), torch.randn(N, L, generator=g).to(DEV).requires_grad_(True)
enc = make_mlp(3, 2 * L, L, 3, layer_norm=True, output_activation="GELU", hidden_activation="GELU").to(DEV)
pts = torch.randn(1_000_000, 3, generator=g).to(DEV)
for _ in range(2):
    out = net.fused([a, b], skip=0)
    torch.autograd.grad(out.sum(), [a, b] + list(net.parameters()))
    o2 = enc(pts)
    torch.autograd.grad(o2.sum(), list(enc.parameters()))
torch.cuda.synchronize()
print("ok")
