"""ORACLE STUB: union of an edge list with its transpose, duplicates removed
(gnn_utils.py:198-199). Output order is canonical (lexicographic)."""
import torch


def symmetrize(src, dst):
    s = torch.as_tensor(getattr(src, "data", src))
    d = torch.as_tensor(getattr(dst, "data", dst))
    e = torch.stack([torch.cat([s, d]), torch.cat([d, s])], 0)
    e = torch.unique(e, dim=1)
    from cudf import Series
    return Series(e[0]), Series(e[1])
