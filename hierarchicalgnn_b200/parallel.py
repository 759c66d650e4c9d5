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
    latency, not for link count). A parameter without a gradient on this rank contributes zeros so every rank reduces
    the same layout; one that has a gradient on NO rank (the last HGNN cell's dead edge networks) keeps
    ``grad = None``, so the optimizer skips it exactly as on one GPU."""
    if not dist.is_initialized():
        return
    world = world_size or dist.get_world_size(group)
    if world == 1:
        return
    params = [p for p in params if p.requires_grad]
    if not params:
        return
    has = torch.tensor([0.0 if p.grad is None else 1.0 for p in params], device=params[0].device)
    dist.all_reduce(has, op=dist.ReduceOp.MAX, group=group)
    has = has.cpu()
    params = [p for p, h in zip(params, has) if h > 0]
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


class GradientBuckets:
    """Bucketed gradient all-reduce OVERLAPPED with the backward pass.

    Parameters are packed, in reverse registration order (the order the backward produces them), into flat fp32
    buckets of ~``bucket_bytes``; every parameter's ``.grad`` is a view into its bucket, so autograd accumulates
    straight into the buffer that travels. A post-accumulate hook counts arrivals; when a bucket is complete its
    all-reduce is issued asynchronously (NCCL's own stream) while the backward keeps producing the earlier layers'
    gradients. ``finish()`` reduces what is left (buckets holding parameters that received no gradient on this rank),
    waits, averages, and restores ``.grad = None`` for parameters that received a gradient on NO rank (the last HGNN
    cell's dead edge / superedge networks: the optimizer must skip them exactly as on one GPU)."""

    def __init__(self, params: Sequence[Tensor], bucket_bytes: int = 4 << 20, group=None, average: bool = True):
        self.group, self.average = group, average
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets: List[dict] = []
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self._close(cur)
                cur, size = [], 0
        if cur:
            self._close(cur)
        self._where = {}
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._where[id(p)] = bi
        self._hooks = [p.register_post_accumulate_grad_hook(self._arrived) for p in self.params]
        self._fired = [0.0] * len(self.params)
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._dead = None  # indices of parameters that get a gradient on no rank (structural: decided on the first step)

    def _close(self, plist):
        n = sum(p.numel() for p in plist)
        flat = torch.zeros(n, dtype=torch.float32, device=plist[0].device)
        views, off = [], 0
        for p in plist:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.buckets.append(dict(params=plist, flat=flat, views=views, pending=len(plist), work=None))

    def prepare(self):
        """Before the backward: zero the buckets and point every .grad at its view."""
        self._fired = [0.0] * len(self.params)
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            b["work"] = None
            for p, v in zip(b["params"], b["views"]):
                p.grad = v

    def _arrived(self, p):
        b = self.buckets[self._where[id(p)]]
        self._fired[self._index[id(p)]] = 1.0
        b["pending"] -= 1
        if b["pending"] == 0 and self.world > 1:
            b["work"] = dist.all_reduce(b["flat"], group=self.group, async_op=True)

    def finish(self):
        """After the backward: reduce the incomplete buckets, wait for all, average; un-set dead gradients."""
        if self.world > 1:
            for b in self.buckets:
                if b["work"] is None:
                    b["work"] = dist.all_reduce(b["flat"], group=self.group, async_op=True)
            fired = None
            if self._dead is None:  # one small MAX all-reduce + host read, first step only
                fired = torch.tensor(self._fired, device=self.buckets[0]["flat"].device)
                dist.all_reduce(fired, op=dist.ReduceOp.MAX, group=self.group)
            for b in self.buckets:
                b["work"].wait()
                if self.average:
                    b["flat"].div_(self.world)
            if fired is not None:
                self._dead = [i for i, f in enumerate(fired.cpu().tolist()) if f == 0]
        else:
            self._dead = [i for i, f in enumerate(self._fired) if f == 0]
        for i in self._dead:
            self.params[i].grad = None

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def sync_buffers(model, src: int = 0, group=None) -> None:
    """Broadcast the floating-point module buffers (BatchNorm running statistics, ``knn_radius``, ``score_cut``) from one
    rank, so that replicas keep building the same graphs and a checkpoint does not depend on the rank that wrote it."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    bufs = [b for b in model.buffers() if b.is_floating_point() and b.numel() > 0]
    if not bufs:
        return
    flat = torch.cat([b.reshape(-1).float() for b in bufs])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for b in bufs:
        b.copy_(flat[off:off + b.numel()].view_as(b))
        off += b.numel()


class DataParallelTrainer:
    """Drives ``training_step`` / ``optimizer_step`` of a LightningModule-style model data-parallel over events (the
    reference trains with batch_size = 1 on one GPU, edge_classifier_base.py:41; Notebooks/script.py:35 clips at 0.5).
    Per step: local forward + backward with the bucketed all-reduce overlapped, global-norm clip on the reduced
    gradients, optimizer step, buffer broadcast. The step counter lives here, not in ``model.trainer``."""

    def __init__(self, model, optimizer, clip: Optional[float] = 0.5, bucket_bytes: int = 4 << 20, group=None,
                 sync_buffers_every: int = 1):
        self.model, self.optimizer, self.clip, self.group = model, optimizer, clip, group
        self.buckets = GradientBuckets(list(model.parameters()), bucket_bytes, group)
        self.global_step = 0
        self.sync_every = sync_buffers_every

    def step(self, batch) -> Tensor:
        self.buckets.prepare()
        loss = self.model.training_step(batch, 0)
        loss.backward()
        self.buckets.finish()
        params = self.buckets.params
        if self.clip is not None:
            clip_grad_norm_(params, self.clip)
        if hasattr(self.model, "trainer") and hasattr(self.model.trainer, "global_step"):
            try:
                self.model.trainer.global_step = self.global_step  # the warm-up schedule of optimizer_step reads it
            except AttributeError:
                pass  # a real pytorch_lightning Trainer owns its counter
        self.model.optimizer_step(optimizer=self.optimizer)
        self.global_step += 1
        if self.sync_every and self.global_step % self.sync_every == 0:
            sync_buffers(self.model, 0, self.group)
        return loss.detach()


def data_parallel_step(model, batch, optimizer, clip: Optional[float] = 0.5, group=None) -> Tensor:
    """One DP training step without persistent state: local ``training_step``, gradient all-reduce (mean) after the
    backward, clip, ``optimizer_step`` hook. ``DataParallelTrainer`` is the overlapped version."""
    optimizer.zero_grad(set_to_none=True)
    loss = model.training_step(batch, 0)
    loss.backward()
    params = [p for p in model.parameters()]
    allreduce_gradients(params, group=group)
    if clip is not None:
        clip_grad_norm_(params, clip)
    model.optimizer_step(optimizer=optimizer)
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
