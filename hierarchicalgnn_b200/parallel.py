"""Multi-GPU execution of the message-passing path on one NVLink/NVSwitch box
(SURVEY.md §8e). The reference has no distributed code at all
(``Trainer(gpus=1)`` everywhere, README.md:65); both modes here are new.

One process per GPU, ``torch.distributed`` (NCCL) as plumbing:

* **Data parallel over events** (BASELINE config 4): every rank runs whole
  events; the only exchange is the all-reduce of the weight gradients
  (``allreduce_gradients``), followed by the global-norm clip the reference
  driver applies (``Trainer(gradient_clip_val=0.5)``, Notebooks/script.py:35).
* **One large event, partitioned by destination node** (BASELINE config 5):
  rank g owns the node block [g*B, (g+1)*B) and every directed edge whose
  destination lies in it. The incoming-edge sum of an owned node is then
  complete locally (the reduce-scatter of partial aggregates degenerates to
  nothing in the forward pass); what must travel is
    forward : all-gather of the updated node rows (each rank needs x[src] of
              arbitrary nodes for its edge step),
    backward: reduce-scatter (sum) of the node-gradient partials — the adjoint
              of that all-gather, i.e. the reduce-scatter of partial node
              aggregates the north star names,
  plus one all-reduce of the weight gradients per step.

The partition driver is written against three callables (node step, edge step,
segment sum) so the same bookkeeping is exercised on CPU/gloo in the tests
(with the oracle's functions) and on GPUs/NCCL with the CUDA modules.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

Tensor = torch.Tensor


# ---------------------------------------------------------------------------
# data parallel
# ---------------------------------------------------------------------------
def allreduce_gradients(params: Sequence[Tensor], world_size: Optional[int] = None, average: bool = True,
                        bucket_bytes: int = 64 << 20, group=None) -> None:
    """In-place all-reduce of ``p.grad`` over the group in flat fp32 buckets (NVSwitch: size buckets for launch
    latency, not for link count). Parameters without a gradient contribute zeros so every rank reduces the same
    layout (the last HGNN cell's dead edge networks)."""
    if not dist.is_initialized():
        return
    world = world_size or dist.get_world_size(group)
    if world == 1:
        return
    params = [p for p in params if p.requires_grad]
    bucket: List[Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket])
        dist.all_reduce(flat, group=group)
        if average:
            flat /= world
        off = 0
        for p in bucket:
            n = p.numel()
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        bucket, size = [], 0

    for p in params:
        bucket.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()


def clip_grad_norm_(params: Sequence[Tensor], max_norm: float) -> Tensor:
    """Global-norm clip on already-reduced gradients (identical on every rank)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return torch.zeros(())
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.float()) for g in grads]))
    scale = (max_norm / (total + 1e-6)).clamp(max=1.0)
    for g in grads:
        g.mul_(scale)
    return total


def data_parallel_step(model, batch, optimizer, clip: Optional[float] = 0.5, group=None) -> Tensor:
    """One DP training step of a LightningModule-style model: local ``training_step`` on this rank's event,
    gradient all-reduce (mean), clip, ``optimizer_step`` hook."""
    optimizer.zero_grad(set_to_none=True)
    loss = model.training_step(batch, 0)
    loss.backward()
    params = [p for p in model.parameters()]
    allreduce_gradients(params, group=group)
    if clip is not None:
        clip_grad_norm_(params, clip)
    model.optimizer_step(optimizer=optimizer)
    model.trainer.global_step += 1
    return loss.detach()


# ---------------------------------------------------------------------------
# destination-partitioned single event
# ---------------------------------------------------------------------------
@dataclass
class EdgePartition:
    rank: int
    world: int
    n_nodes: int            # global node count
    block: int              # padded node block size B (equal on every rank)
    node_lo: int            # first owned node
    node_hi: int            # one past the last owned node (<= n_nodes)
    edge_ids: Tensor        # [E_g] ids (into the global directed edge list) of the owned edges, ordered by destination
    graph: Tensor           # [2, E_g] owned edges, GLOBAL node ids
    dst_local: Tensor       # [E_g] destination ids relative to node_lo (segment ids of the local aggregate)

    @property
    def n_owned(self) -> int:
        return self.node_hi - self.node_lo


def partition_by_destination(graph: Tensor, n_nodes: int, world: int, rank: int) -> EdgePartition:
    """Rank ``rank`` owns nodes [rank*B, (rank+1)*B) with B = ceil(n_nodes / world) and all edges pointing into them."""
    block = (n_nodes + world - 1) // world
    lo, hi = rank * block, min(n_nodes, (rank + 1) * block)
    dst = graph[1]
    mine = ((dst >= lo) & (dst < hi)).nonzero().squeeze(1)
    mine = mine[torch.argsort(dst[mine], stable=True)]  # destination-sorted: the kernels stream rows in place
    g = graph[:, mine].contiguous()
    return EdgePartition(rank, world, n_nodes, block, lo, hi, mine, g, (g[1] - lo).contiguous())


class _AllGatherRows(torch.autograd.Function):
    """Row all-gather of equal-sized blocks; the adjoint is a sum reduce-scatter of the gradient."""

    @staticmethod
    def forward(ctx, local: Tensor, group):
        ctx.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return local.clone()
        out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, grad: Tensor):
        world = dist.get_world_size(ctx.group) if dist.is_initialized() else 1
        if world == 1:
            return grad, None
        grad = grad.contiguous()
        out = grad.new_empty((grad.shape[0] // world,) + tuple(grad.shape[1:]))
        if dist.get_backend(ctx.group) == "gloo":  # gloo has no reduce_scatter_tensor: all-reduce + slice (CPU tests only)
            full = grad.clone()
            dist.all_reduce(full, group=ctx.group)
            r = dist.get_rank(ctx.group)
            out.copy_(full[r * out.shape[0]:(r + 1) * out.shape[0]])
        else:
            dist.reduce_scatter_tensor(out, grad, group=ctx.group)
        return out, None


def all_gather_rows(local: Tensor, group=None) -> Tensor:
    return _AllGatherRows.apply(local, group)


def pad_rows(t: Tensor, rows: int) -> Tensor:
    if t.shape[0] == rows:
        return t
    return torch.cat([t, t.new_zeros((rows - t.shape[0],) + tuple(t.shape[1:]))], 0)


def partitioned_interaction_cell(part: EdgePartition, nodes_full: Tensor, edges_local: Tensor,
                                 node_fn: Callable[[Tensor, Tensor], Tensor],
                                 edge_fn: Callable[[Tensor, Tensor, Tensor], Tensor],
                                 segment_sum: Callable[[Tensor, Tensor, int], Tensor], group=None):
    """One InteractionGNNCell (gnn_utils.py:66-71) on a destination partition.

    nodes_full  [world*B, L]  replicated node latents (rows >= n_nodes are padding)
    edges_local [E_g, L]      latents of the owned edges
    node_fn(x_owned, agg_owned) -> x_owned'       (node MLP + skip on the owned block)
    edge_fn(x_full, e_local, graph_local) -> e_local'
    segment_sum(rows, seg_ids, n_seg) -> [n_seg, L]
    Returns (nodes_full', edges_local')."""
    lo = part.node_lo
    agg = segment_sum(edges_local, part.dst_local, part.block)         # complete for owned nodes: no exchange
    x_owned = nodes_full[lo:lo + part.block]
    x_new = node_fn(x_owned, agg)
    if part.n_owned < part.block:                                       # keep padding rows inert
        keep = (torch.arange(part.block, device=x_new.device) < part.n_owned).unsqueeze(1)
        x_new = torch.where(keep, x_new, torch.zeros_like(x_new))
    nodes_new = all_gather_rows(x_new, group)                           # fwd: all-gather; bwd: reduce-scatter(sum)
    edges_new = edge_fn(nodes_new, edges_local, part.graph)
    return nodes_new, edges_new


def cuda_cell_callables(cell):
    """Binds a hierarchicalgnn_b200 InteractionGNNCell to the partition driver."""
    from . import ops
    from .gnn_utils import GraphPlans

    def node_fn(x_owned, agg):
        return cell.node_network.fused([x_owned, agg], skip=0)

    def edge_fn(x_full, e_local, graph_local):
        gp = GraphPlans(graph_local, x_full.shape[0], x_full.shape[0], dst_sorted=True)  # partition_by_destination sorts
        return cell.edge_network.fused([x_full, x_full, e_local], [gp.by_src, gp.by_dst, None], skip=2)

    def segment_sum(rows, seg, n):
        return ops.scatter_add(rows, seg, dim_size=n)

    return node_fn, edge_fn, segment_sum
