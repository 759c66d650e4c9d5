import sys, torch
sys.path.insert(0, '.')
from hierarchicalgnn_b200 import ops, _lib
import oracle.hgnn_oracle as O
nq, nr, dim, k = 1200, 1200, 8, 10
g = torch.Generator().manual_seed(nq + nr)
q = torch.randn(nq, dim, generator=g); r = torch.randn(nr, dim, generator=g)
r[nr // 2:nr // 2 + nr // 8] = r[:nr // 8]
L = _lib.lib()
qd, rd = q.cuda(), r.cuda()
one = torch.empty((nq, k), dtype=torch.int64, device='cuda')
L.hgnn_knn_radius(qd.data_ptr(), nq, rd.data_ptr(), nr, dim, k, 1.9, one.data_ptr(), None)
split = ops.knn_radius(qd, rd, k, 1.9)
torch.cuda.synchronize()
want = O.knn_radius(q, r, k, 1.9)
print("split==one", (split == one).float().mean().item(), "one==want", (one.cpu() == want).float().mean().item(), "split==want", (split.cpu() == want).float().mean().item())
bad = (split != one).any(1).nonzero().squeeze(1)[:5]
for b in bad.tolist():
    print(b, one[b].tolist(), split[b].tolist(), want[b].tolist())
