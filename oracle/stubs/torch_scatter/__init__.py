"""ORACLE STUB (test infrastructure only, never shipped on the product path).

CPU restatement of the torch_scatter 2.0.9 entry points the reference calls
(gnn_utils.py:50,124-125,142-143; BC/Models/HGNN_GMM.py:251,269;
bipartite_classification_base.py:158; tracking_utils.py:41).
"""
import torch


def _expand(index, src):
    shape = [src.shape[0]] + [1] * (src.dim() - 1)
    return index.reshape(shape).expand_as(src)


def scatter_add(src, index, dim=0, dim_size=None):
    assert dim == 0
    n = int(dim_size) if dim_size is not None else int(index.max()) + 1
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.scatter_add_(0, _expand(index, src), src)


scatter_sum = scatter_add


def scatter_mean(src, index, dim=0, dim_size=None):
    assert dim == 0
    n = int(dim_size) if dim_size is not None else int(index.max()) + 1
    total = scatter_add(src, index, 0, n)
    count = torch.zeros(n, dtype=src.dtype, device=src.device)
    count.scatter_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    count = count.clamp(min=1).reshape([n] + [1] * (src.dim() - 1))
    return total / count


def _scatter_extreme(src, index, dim_size, reduce):
    n = int(dim_size) if dim_size is not None else int(index.max()) + 1
    fill = float("inf") if reduce == "amin" else float("-inf")
    out = torch.full((n,) + tuple(src.shape[1:]), fill, dtype=src.dtype, device=src.device)
    out.scatter_reduce_(0, _expand(index, src), src, reduce=reduce, include_self=True)
    hit = src == out[index]
    pos = torch.arange(src.shape[0], device=src.device)
    pos = _expand(pos, src).clone()
    pos[~hit] = src.shape[0]
    arg = torch.full_like(out, src.shape[0], dtype=torch.long)
    arg.scatter_reduce_(0, _expand(index, src), pos, reduce="amin", include_self=True)
    out[torch.isinf(out)] = 0
    return out, arg


def scatter_min(src, index, dim=0, dim_size=None):
    return _scatter_extreme(src, index, dim_size, "amin")


def scatter_max(src, index, dim=0, dim_size=None):
    return _scatter_extreme(src, index, dim_size, "amax")
