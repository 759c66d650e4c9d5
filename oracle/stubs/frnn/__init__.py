"""ORACLE STUB (test infrastructure only).

CPU restatement of lxxue/FRNN `frnn_grid_points` as consumed by the reference
(utils.py:228-239, gnn_utils.py:194-202): for each row of points1 the up-to-K
rows of points2 with Euclidean distance < r, ascending by distance, -1 padded.
Distances are evaluated in float64 with the direct sum-of-squared-differences
form; ties break towards the smaller index (the contract this repo defines,
SURVEY.md B.2).
"""
import torch


def knn_radius_reference(points1, points2, K, r, chunk=2048):
    p1 = points1.detach().to(torch.float64)
    p2 = points2.detach().to(torch.float64)
    r2 = float(r) ** 2
    n1, n2 = p1.shape[0], p2.shape[0]
    kk = min(K, n2)
    idxs = torch.full((n1, K), -1, dtype=torch.long)
    dists = torch.full((n1, K), -1.0, dtype=torch.float64)
    for s in range(0, n1, chunk):
        q = p1[s:s + chunk]
        d2 = (q[:, None, :] - p2[None, :, :]).square().sum(-1)
        # stable sort => smaller index first among equal distances
        order = torch.sort(d2, dim=1, stable=True)
        dk, ik = order.values[:, :kk], order.indices[:, :kk]
        ok = dk < r2
        idxs[s:s + chunk, :kk] = torch.where(ok, ik, torch.full_like(ik, -1))
        dists[s:s + chunk, :kk] = torch.where(ok, dk, torch.full_like(dk, -1.0))
    return dists, idxs


def frnn_grid_points(points1, points2, lengths1=None, lengths2=None, K=10, r=1.0,
                     grid=None, return_nn=False, return_sorted=True, radius_cell_ratio=2.0):
    assert points1.dim() == 3 and points1.shape[0] == 1
    if torch.is_tensor(r):
        r = float(r.reshape(-1)[0])
    dists, idxs = knn_radius_reference(points1[0], points2[0], K, r)
    return dists.to(points1.dtype)[None], idxs[None].to(points1.device), None, None
