"""Tracking-performance metrics of the evaluation steps (reference: Modules/tracking_utils.py:18-83, called from
``shared_evaluation`` of both task bases). Same inputs, same four numbers; the reference builds a cupy sparse
particle x candidate count matrix and compares dense slices of it, here the non-zero (particle, candidate) cells are
kept as a sorted list (``torch.unique`` of the pair keys) and every criterion is a per-cell test against per-particle /
per-candidate vectors — plain torch ops on whatever device the event lives on, no host loop, no cupy.
"""
from __future__ import annotations

import torch

default_response = {"track_eff": 0, "track_pur": 0, "hit_eff": 0, "hit_pur": 0}


def _get(event, name):
    if isinstance(event, dict):
        return event.get(name)
    return getattr(event, name, None)


def eval_metrics(bipartite_graph, event, pt_cut=1.0, nhits_cut=5, majority_cut=0.5, primary=True):
    """bipartite_graph[2, A]: (hit, track candidate) assignments; event: ``pid``, ``pt`` (and ``primary``) per hit.
    Returns {"track_eff", "track_pur", "hit_eff", "hit_pur"} (tracking_utils.py:18-83):
      * candidates with fewer than nhits_cut * majority_cut hits are dropped, the rest renumbered;
      * a particle matches a candidate when it owns >= majority_cut of the candidate's hits, the candidate holds
        >= majority_cut of the particle's hits, and the candidate is the particle's best one (ties: the higher candidate
        id, which is what the reference's increasing ``cluster_hashing`` factor selects);
      * matches to noise (pid 0) or with <= majority_cut * nhits_cut shared hits do not count;
      * reconstructable = pT > pt_cut and >= nhits_cut hits (and primary, when asked for and present)."""
    bipartite_graph = bipartite_graph.long()
    pid_all, pt_all = _get(event, "pid"), _get(event, "pt")
    dev = pid_all.device
    if bipartite_graph.shape[1] == 0:
        return dict(default_response)
    _, cand, counts = bipartite_graph[1].unique(return_inverse=True, return_counts=True)
    keep = counts[cand] >= (nhits_cut * majority_cut)
    hits = bipartite_graph[0][keep]
    if hits.numel() == 0:
        return dict(default_response)
    cand = bipartite_graph[1][keep].unique(return_inverse=True)[1]
    n_cand = int(cand.max()) + 1

    original_pid, pid, nhits = torch.unique(pid_all, return_inverse=True, return_counts=True)
    n_part = original_pid.numel()
    pt = torch.full((n_part,), float("inf"), device=dev, dtype=pt_all.dtype).scatter_reduce(0, pid, pt_all, "amin")
    primary_mask = None
    prim = _get(event, "primary")
    if primary and prim is not None:
        primary_mask = torch.zeros(n_part, device=dev, dtype=torch.float64).index_add_(0, pid, prim.double()) > 0

    # non-zero cells of the particle x candidate count matrix
    cell, shared = torch.unique(pid[hits] * n_cand + cand, return_counts=True)
    row, col = torch.div(cell, n_cand, rounding_mode="floor"), cell % n_cand
    cand_size = torch.bincount(cand, minlength=n_cand)
    # the particle's best candidate: most shared hits, ties to the higher candidate id
    key = shared * n_cand + col
    best = torch.zeros(n_part, dtype=key.dtype, device=dev).scatter_reduce(0, row, key, "amax")
    matching = (shared >= majority_cut * cand_size[col]) & (shared >= majority_cut * nhits[row]) & (key == best[row])
    row, col, shared = row[matching], col[matching], shared[matching]
    if row.numel() == 0:
        return dict(default_response)

    matching_mask = (shared > majority_cut * nhits_cut) & (original_pid[row] != 0)
    n_rejected = int((~matching_mask).sum())
    row, col, shared = row[matching_mask], col[matching_mask], shared[matching_mask]
    if row.numel() == 0:
        return dict(default_response)

    mask = (pt[row] > pt_cut) & (nhits[row] >= nhits_cut)
    truth_mask = (pt > pt_cut) & (nhits >= nhits_cut)
    if primary and primary_mask is not None:
        mask = mask & primary_mask[row]
        truth_mask = truth_mask & primary_mask

    n_good = mask.sum().double()
    track_eff = n_good / truth_mask.sum().double()
    hit_pur = (shared.double() / cand_size[col].double()).mean()
    track_pur = n_good / (n_cand - n_rejected - int((~mask).sum()))
    hit_eff = (shared[mask].double() / nhits[row][mask].double()).mean()
    return {"track_eff": track_eff.item(), "track_pur": track_pur.item(), "hit_eff": hit_eff.item(), "hit_pur": hit_pur.item()}
