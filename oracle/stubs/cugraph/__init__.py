"""ORACLE STUB (test infrastructure only): the slice of cugraph 22.04 the
reference uses (gnn_utils.py:198; BC/Models/HGNN_GMM.py:215-232;
edge_classifier_base.py:157-165), restated on scipy."""
import numpy as np
import torch
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components as _cc

from . import structure, components  # noqa: F401


class Graph:
    def __init__(self):
        self.src = None
        self.dst = None

    def from_cudf_edgelist(self, df, source="src", destination="dst", edge_attr=None):
        self.src = np.asarray(torch.as_tensor(df[source]).cpu())
        self.dst = np.asarray(torch.as_tensor(df[destination]).cpu())
        if self.src.size == 0:
            raise ValueError("empty edge list")
