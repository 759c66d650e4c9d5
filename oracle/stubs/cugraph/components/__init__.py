"""ORACLE STUB: weakly connected components over the vertices that occur in
the edge list; label = smallest vertex id of the component (canonical form,
SURVEY.md B.3)."""
import numpy as np
import torch
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components as _cc


def connected_components(G):
    src, dst = G.src, G.dst
    n = int(max(src.max(), dst.max())) + 1
    adj = coo_matrix((np.ones(len(src), dtype=np.int8), (src, dst)), shape=(n, n))
    _, lab = _cc(adj, directed=False)
    present = np.zeros(n, dtype=bool)
    present[src] = True
    present[dst] = True
    vertex = np.nonzero(present)[0]
    lab = lab[vertex]
    # canonical label: min vertex id per component
    first = np.full(lab.max() + 1, n, dtype=np.int64)
    np.minimum.at(first, lab, vertex)
    return {"vertex": torch.from_numpy(vertex.astype(np.int64)),
            "labels": torch.from_numpy(first[lab])}
