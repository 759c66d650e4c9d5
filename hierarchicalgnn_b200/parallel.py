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

import os
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
    """Global-norm clip on already-reduced gradients (identical on every rank). Multi-tensor kernels: two launches for
    the norms and one for the scaling instead of two per parameter (5.5 ms -> 0.3 ms on the 370 tensors of BC-HGNN-GMM)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return torch.zeros(())
    if all(g.is_cuda and g.dtype == torch.float32 for g in grads):
        total = torch.linalg.vector_norm(torch.stack(torch._foreach_norm(grads)))
        scale = (max_norm / (total + 1e-6)).clamp(max=1.0)
        torch._foreach_mul_(grads, scale)
        return total
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.float()) for g in grads]))
    scale = (max_norm / (total + 1e-6)).clamp(max=1.0)
    for g in grads:
        g.mul_(scale)
    return total


class GradientBuckets:
    """Bucketed gradient all-reduce OVERLAPPED with the backward pass.

    Parameters are packed, in reverse registration order (the order the backward produces them), into flat fp32
    buckets of ~``bucket_bytes``. A post-accumulate hook counts arrivals; when a bucket is complete its gradients are
    packed into the flat buffer with one multi-tensor copy, every ``.grad`` becomes a view into it, and its all-reduce is
    issued asynchronously (NCCL's own stream) while the backward keeps producing the earlier layers' gradients. ``finish()`` reduces what is left (buckets holding parameters that received no gradient on this rank),
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
        self._next = 0

    def _close(self, plist):
        n = sum(p.numel() for p in plist)
        flat = torch.zeros(n, dtype=torch.float32, device=plist[0].device)
        views, off = [], 0
        for p in plist:
            views.append(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.buckets.append(dict(params=plist, flat=flat, views=views, pending=len(plist), work=None))

    def prepare(self):
        """Before the backward: forget the previous step's gradients. The backward then leaves every gradient in its own
        tensor (autograd hands the first contribution over without a copy); a bucket is packed — one multi-tensor copy —
        when its last parameter has arrived. (Pointing .grad at zeroed bucket views instead made autograd run one small
        in-place add per parameter: ~400 launches per BC-HGNN step.)"""
        self._fired = [0.0] * len(self.params)
        self._next = 0  # buckets are reduced strictly in index order: every rank issues the same sequence of collectives
        for b in self.buckets:
            b["pending"] = len(b["params"])
            b["work"] = None
            b["packed"] = False
            b["ready"] = False
            for p in b["params"]:
                p.grad = None

    def _pack(self, b):
        """Gradients of the bucket's parameters -> its flat buffer; .grad becomes the view (what travels is what the optimizer
        reads). Parameters without a gradient on this rank contribute zeros."""
        if b["packed"]:
            return
        dst, src = [], []
        for p, v in zip(b["params"], b["views"]):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(p.grad.detach().reshape(v.shape))
        if dst:
            torch._foreach_copy_(dst, src)
        for p, v in zip(b["params"], b["views"]):
            if p.grad is not None:
                p.grad = v
        b["packed"] = True

    def _arrived(self, p):
        b = self.buckets[self._where[id(p)]]
        self._fired[self._index[id(p)]] = 1.0
        b["pending"] -= 1
        if b["pending"] == 0 and self.world > 1:
            b["ready"] = True
            self._launch(only_ready=True)

    def _launch(self, only_ready: bool):
        """Issue the all-reduce of the next buckets IN INDEX ORDER (a bucket that completed early waits for its predecessors:
        ranks whose backward completes buckets in different orders — a branch taken on one rank only — would otherwise pair
        different buckets in the same collective)."""
        while self._next < len(self.buckets):
            b = self.buckets[self._next]
            if only_ready and not b["ready"]:
                return
            self._pack(b)
            b["work"] = dist.all_reduce(b["flat"], group=self.group, async_op=True)
            self._next += 1

    def finish(self):
        """After the backward: reduce the incomplete buckets, wait for all, average; un-set dead gradients. On one rank
        nothing is packed or copied: the gradients stay where autograd left them."""
        if self.world > 1:
            self._launch(only_ready=False)
            fired = None
            if self._dead is None:  # one small MAX all-reduce + host read, first step only
                fired = torch.tensor(self._fired, device=self.buckets[0]["flat"].device)
                dist.all_reduce(fired, op=dist.ReduceOp.MAX, group=self.group)
            for b in self.buckets:
                b["work"].wait()
                if self.average:
                    b["flat"].div_(self.world)
                for p, v in zip(b["params"], b["views"]):
                    p.grad = v  # also for parameters that got their gradient from another rank only
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


class SymmetricRows:
    """The two [world * block, width] fp32 tables of the destination-partitioned event (replicated node rows; node-gradient
    partials) as SYMMETRIC memory: one allocation per rank, mapped into every peer (``torch.distributed._symmetric_memory``
    does the handle exchange), so that the row collectives are this repo's own peer-memory kernels
    (``csrc/p2p.cu``: ``multimem.st`` / ``multimem.ld_reduce`` through the NVSwitch, or plain peer stores / loads) instead of
    NCCL calls between kernels. ``all_gather`` / ``reduce_scatter`` bracket the kernels with the symmetric-memory barrier:
    before a table is overwritten (every peer has finished with its previous contents) and after it is written (before
    anybody reads it)."""

    def __init__(self, block: int, width: int, device, group=None, use_multicast: Optional[bool] = None, slots: int = 1):
        import ctypes
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.block, self.width, self.slots = int(block), int(width), max(1, int(slots))
        gname = self.group.group_name
        rows = self.world * self.block
        # tables [0, slots): all-gather destinations (one per cell of a stack, so that a cell's gathered table can stay
        # alive as an autograd saved tensor while the next cell gathers); table [slots]: reduce-scatter source
        self.tables = symm.empty((self.slots + 1, rows, self.width), dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.tables, gname)
        self.table_bytes = rows * self.width * 4
        mc = 0
        if use_multicast is None:
            use_multicast = os.environ.get("HGNN_P2P_MULTICAST", "1") != "0"
        if use_multicast:  # 0 / None when the box has no NVLS multicast object for this allocation
            mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.mc_base = mc
        self._peers = [(ctypes.c_uint64 * self.world)(*[int(b) + t * self.table_bytes for b in self.hdl.buffer_ptrs])
                       for t in range(self.slots + 1)]
        self.channel = 0
        # one independent tensor per table over the same memory (own TensorImpl, own version counter — NOT views of
        # self.tables): a gathered table is an autograd saved tensor while the reduce-scatter table is copied into in place
        n_el = rows * self.width
        self._alias = [self.hdl.get_buffer(self.rank, (rows, self.width), torch.float32, t * n_el) for t in range(self.slots + 1)]

    def _barrier(self):
        self.hdl.barrier(channel=self.channel)

    def _mc(self, t: int):
        return (self.mc_base + t * self.table_bytes) if self.mc_base else None

    def all_gather(self, local: Tensor, slot: Optional[int] = None) -> Tensor:
        """local [block, width] -> [world * block, width] holding every rank's block. With ``slot`` the result IS the
        symmetric table of that slot (valid until the same slot is gathered into again: one slot per cell of a stack, and
        the caller's step boundary — a collective every rank takes part in — separates the last reader from the next
        writer); without it the shared slot 0 is used under a leading barrier and the result is a private copy."""
        from . import ops, _lib
        local = ops._f32(local)
        assert tuple(local.shape) == (self.block, self.width), (tuple(local.shape), self.block, self.width)
        t = 0 if slot is None else int(slot) % self.slots
        if slot is None:
            self._barrier()                   # every peer has copied the previous table out
        ops.check(_lib.lib().hgnn_p2p_all_gather_rows(local.data_ptr(), self.block, self.width, self._mc(t), self._peers[t],
                                                      self.world, self.rank, ops._stream()), "p2p_all_gather_rows")
        ops._count()
        self._barrier()                       # every block has landed everywhere
        return self._alias[t].clone() if slot is None else self._alias[t]

    def reduce_scatter(self, full: Tensor) -> Tensor:
        """full [world * block, width] partials -> [block, width] = sum over ranks of this rank's block."""
        from . import ops, _lib
        full = ops._f32(full)
        t = self.slots
        self._barrier()                       # every peer has reduced the previous partials
        self._alias[t].copy_(full)
        self._barrier()                       # every rank's partials are in place
        out = torch.empty((self.block, self.width), dtype=torch.float32, device=full.device)
        ops.check(_lib.lib().hgnn_p2p_reduce_scatter_rows(out.data_ptr(), self.block, self.width, self._mc(t), self._peers[t],
                                                          self.world, self.rank, ops._stream()), "p2p_reduce_scatter_rows")
        ops._count()
        return out


    def all_reduce_(self, flat: Tensor) -> Tensor:
        """In-place sum over ranks of a flat fp32 tensor (the weight gradients of a step) through the reduce-scatter table:
        one kernel, each rank reduces 1/world of it inside the switch and stores the result to every rank."""
        from . import ops, _lib
        n = flat.numel()
        t = self.slots
        buf = self._alias[t].view(-1)
        n_pad = (n + 3) // 4 * 4
        if flat.dtype != torch.float32 or n_pad > buf.numel():
            dist.all_reduce(flat, group=self.group)
            return flat
        self._barrier()                       # the table is free (previous partials reduced everywhere)
        buf[:n].copy_(flat.reshape(-1))
        if n_pad > n:
            buf[n:n_pad].zero_()
        self._barrier()                       # every rank's addend is in place
        ops.check(_lib.lib().hgnn_p2p_all_reduce(n_pad, self._mc(t), self._peers[t], self.world, self.rank, ops._stream()),
                  "p2p_all_reduce")
        ops._count()
        self._barrier()                       # every slice has been stored everywhere
        flat.reshape(-1).copy_(buf[:n])
        return flat


class _SymmAllGatherRows(torch.autograd.Function):
    """Row all-gather through the symmetric tables (peer-memory kernels); the adjoint is the sum reduce-scatter."""

    @staticmethod
    def forward(ctx, local: Tensor, sr: SymmetricRows, slot):
        ctx.sr = sr
        out = sr.all_gather(local, slot)
        return out

    @staticmethod
    def backward(ctx, grad: Tensor):
        return ctx.sr.reduce_scatter(grad.contiguous()), None, None


def all_gather_rows(local: Tensor, group=None, symmetric: Optional[SymmetricRows] = None, slot: Optional[int] = None) -> Tensor:
    if symmetric is not None:
        return _SymmAllGatherRows.apply(local, symmetric, slot)
    return _AllGatherRows.apply(local, group)


def pad_rows(t: Tensor, rows: int) -> Tensor:
    if t.shape[0] == rows:
        return t
    return torch.cat([t, t.new_zeros((rows - t.shape[0],) + tuple(t.shape[1:]))], 0)


def partitioned_interaction_cell(part: EdgePartition, nodes_full: Tensor, edges_local: Tensor,
                                 node_fn: Callable[[Tensor, Tensor], Tensor],
                                 edge_fn: Callable[[Tensor, Tensor, Tensor], Tensor],
                                 segment_sum: Callable[[Tensor, Tensor, int], Tensor], group=None,
                                 symmetric: Optional[SymmetricRows] = None, agg_owned: Optional[Tensor] = None,
                                 return_agg: bool = False, x_owned: Optional[Tensor] = None, slot: Optional[int] = None,
                                 return_owned: bool = False):
    """One InteractionGNNCell (gnn_utils.py:66-71) on a destination partition.

    nodes_full  [world*B, L]  replicated node latents (rows >= n_nodes are padding)
    edges_local [E_g, L]      latents of the owned edges
    node_fn(x_owned, agg_owned) -> x_owned'       (node MLP + skip on the owned block)
    edge_fn(x_full, e_local, graph_local) -> e_local'  or  (e_local', agg_full') when it fuses the next cell's aggregate
    segment_sum(rows, seg_ids, n_seg) -> [n_seg, L]
    agg_owned   [B, L]        incoming-edge sums of the owned block if a previous cell's edge step already produced them
    x_owned     [B, L]        the owned block of nodes_full as its own tensor (the previous cell's block before it was
                              gathered): its gradient then stays a [B, L] block instead of a zero-padded [world*B, L]
                              table added to the gathered table's gradient
    Returns (nodes_full', edges_local'), then — with ``return_agg`` — the owned block of scatter_add(edges_local') or
    None, then — with ``return_owned`` — the owned block x_owned' that was gathered."""
    lo = part.node_lo
    if agg_owned is None:
        agg_owned = segment_sum(edges_local, part.dst_local, part.block)  # complete for owned nodes: no exchange
    if x_owned is None:
        x_owned = nodes_full[lo:lo + part.block]
    x_new = node_fn(x_owned, agg_owned)
    if part.n_owned < part.block:                                       # keep padding rows inert
        keep = (torch.arange(part.block, device=x_new.device) < part.n_owned).unsqueeze(1)
        x_new = torch.where(keep, x_new, torch.zeros_like(x_new))
    nodes_new = all_gather_rows(x_new, group, symmetric, slot)          # fwd: all-gather; bwd: reduce-scatter(sum)
    res = edge_fn(nodes_new, edges_local, part.graph)
    edges_new, agg_next = (res if isinstance(res, tuple) else (res, None))
    if agg_next is not None:
        agg_next = agg_next[lo:lo + part.block]
    out = [nodes_new, edges_new]
    if return_agg:
        out.append(agg_next)
    if return_owned:
        out.append(x_new)
    return tuple(out)


class _AllReduceSum(torch.autograd.Function):
    """Sum over ranks of per-rank partials whose consumers are REPLICATED (every rank goes on to compute the same thing
    from the sum): the gradient arriving at the sum is already identical on every rank and complete, so the adjoint is
    the identity."""

    @staticmethod
    def forward(ctx, partial: Tensor, group):
        out = partial.clone()
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(out, group=group)
        return out

    @staticmethod
    def backward(ctx, grad: Tensor):
        return grad, None


class _ReplicatedToLocal(torch.autograd.Function):
    """Marks the point where a REPLICATED tensor (identical on every rank, computed redundantly) enters rank-local work:
    forward is the identity; each rank's gradient is only its share, so the adjoint sums over ranks — every replica then
    back-propagates the complete gradient and the replicas stay identical."""

    @staticmethod
    def forward(ctx, t: Tensor, group):
        ctx.group = group
        return t.view_as(t)

    @staticmethod
    def backward(ctx, grad: Tensor):
        g = grad.contiguous().clone()
        if dist.is_initialized() and dist.get_world_size(ctx.group) > 1:
            dist.all_reduce(g, group=ctx.group)
        return g, None


@dataclass
class BipartitePartition:
    """The rows of the hit -> supernode assignment graph (HGNN_GMM.py:236-271) whose HIT lies in this rank's node block."""
    ids: Tensor          # [E_b,g] positions in the global bipartite edge list
    node_local: Tensor   # [E_b,g] hit id relative to node_lo
    supernode: Tensor    # [E_b,g] supernode id (global: supernodes are replicated)


def partition_bipartite(bipartite_graph: Tensor, part: EdgePartition) -> BipartitePartition:
    node = bipartite_graph[0]
    mine = ((node >= part.node_lo) & (node < part.node_hi)).nonzero().squeeze(1)
    return BipartitePartition(mine, (node[mine] - part.node_lo).contiguous(), bipartite_graph[1][mine].contiguous())


def partitioned_hierarchical_cell(part: EdgePartition, bpart: BipartitePartition, nodes_full: Tensor, edges_local: Tensor,
                                  supernodes: Tensor, superedges: Tensor, bweights_local: Tensor, super_graph: Tensor,
                                  super_edge_weights: Tensor, fns: dict, group=None, symmetric: Optional[SymmetricRows] = None,
                                  agg_owned: Optional[Tensor] = None, x_owned: Optional[Tensor] = None,
                                  slot: Optional[int] = None, skip_edge_updates: bool = False):
    """One HierarchicalGNNCell (gnn_utils.py:119-169) on a destination partition. Hits and hit-hit edges are partitioned as
    in ``partitioned_interaction_cell``; the supernode side (S << N rows: supernodes, superedges, the supernode graph and
    its weights) is REPLICATED — every rank computes it redundantly from identical inputs — and the hit <-> supernode
    messages cross between the two worlds:

      up    hits -> supernodes   each rank sums its owned hits' weighted rows per supernode; the partials are all-reduced
      down  supernodes -> hits   rank-local (every rank holds all supernodes); the adjoint all-reduces the supernode gradient

    fns: supernode(supernodes, attention, up), node(x_owned, agg_owned, down_owned), superedge(supernodes, superedges,
    super_graph), edge(x_full, e_local, graph_local) [-> e' or (e', agg_full)], segment_sum(rows, ids, n),
    weighted_sum(rows, weights, gather_ids, seg_ids, n) = scatter_add(weights * rows[gather_ids], seg_ids, n).
    Returns dict(nodes, edges, supernodes, superedges, agg_owned, x_owned). The weight gradients of fns["node"] / fns["edge"]
    are per-rank partials (sum them over ranks); those of fns["supernode"] / fns["superedge"] are complete on every rank."""
    lo, B, S = part.node_lo, part.block, supernodes.shape[0]
    if x_owned is None:
        x_owned = nodes_full[lo:lo + B]
    # supernode update (gnn_utils.py:137-145)
    up = _AllReduceSum.apply(fns["weighted_sum"](x_owned, bweights_local, bpart.node_local, bpart.supernode, S), group)
    attention = fns["weighted_sum"](superedges, super_edge_weights, None, super_graph[1], S)
    supernodes_new = fns["supernode"](supernodes, attention, up)
    # node update (gnn_utils.py:119-127): incoming hit-hit edge sums (local), messages from the supernodes (local reads of
    # the replicated table)
    sn_local = _ReplicatedToLocal.apply(supernodes_new, group)
    down = fns["weighted_sum"](sn_local, bweights_local, bpart.supernode, bpart.node_local, B)
    if agg_owned is None:
        agg_owned = fns["segment_sum"](edges_local, part.dst_local, B)
    x_new = fns["node"](x_owned, agg_owned, down)
    if part.n_owned < B:
        keep = (torch.arange(B, device=x_new.device) < part.n_owned).unsqueeze(1)
        x_new = torch.where(keep, x_new, torch.zeros_like(x_new))
    nodes_new = all_gather_rows(x_new, group, symmetric, slot)
    edges_new, superedges_new, agg_next = edges_local, superedges, None
    if not skip_edge_updates:  # (the last cell of a block: HGNN_GMM.py skips the dead edge updates)
        superedges_new = fns["superedge"](supernodes_new, superedges, super_graph)
        res = fns["edge"](nodes_new, edges_local, part.graph)
        edges_new, agg_next = (res if isinstance(res, tuple) else (res, None))
        if agg_next is not None:
            agg_next = agg_next[lo:lo + B]
    return dict(nodes=nodes_new, edges=edges_new, supernodes=supernodes_new, superedges=superedges_new,
                agg_owned=agg_next, x_owned=x_new)


def cuda_hier_cell_callables(cell, fuse_aggregate: bool = False) -> dict:
    """Binds a hierarchicalgnn_b200 HierarchicalGNNCell to ``partitioned_hierarchical_cell``."""
    from . import ops
    from .gnn_utils import GraphPlans

    def weighted_sum(rows, weights, gather_ids, seg_ids, n):
        seg_plan = ops.plan_for(seg_ids, n)
        if gather_ids is None:
            return ops._GatherScatter.apply(rows, weights, None, seg_plan, False)
        return ops.gather_scatter(rows, weights, ops.plan_for(gather_ids, rows.shape[0]), seg_plan)

    def edge(x_full, e_local, graph_local):
        gp = GraphPlans(graph_local, x_full.shape[0], x_full.shape[0], dst_sorted=True)
        if fuse_aggregate:
            e_new, agg = cell.edge_network.edge_step(x_full, e_local, gp.by_src, gp.by_dst)
            return (e_new, agg) if agg is not None else e_new
        return cell.edge_network.fused([x_full, x_full, e_local], [gp.by_src, gp.by_dst, None], skip=2)

    def superedge(supernodes, superedges, super_graph):
        S = supernodes.shape[0]
        gp = GraphPlans(super_graph, S, S)
        return cell.superedge_network.fused([supernodes, supernodes, superedges], [gp.by_src, gp.by_dst, None], skip=2)

    return dict(supernode=lambda sn, att, up: cell.supernode_network.fused([sn, att, up], skip=0),
                node=lambda x, agg, down: cell.node_network.fused([x, agg, down], skip=0),
                superedge=superedge, edge=edge, weighted_sum=weighted_sum,
                segment_sum=lambda rows, seg, n: ops.scatter_add(rows, seg, dim_size=n))


def cuda_cell_callables(cell, fuse_aggregate: bool = False):
    """Binds a hierarchicalgnn_b200 InteractionGNNCell to the partition driver. With ``fuse_aggregate`` the edge step also
    returns scatter_add(e', dst) over ALL node rows (its own fused segmented reduce: rows outside the owned block stay zero),
    which the driver hands to the next cell instead of running a separate segment sum."""
    from . import ops
    from .gnn_utils import GraphPlans

    def node_fn(x_owned, agg):
        return cell.node_network.fused([x_owned, agg], skip=0)

    def edge_fn(x_full, e_local, graph_local):
        gp = GraphPlans(graph_local, x_full.shape[0], x_full.shape[0], dst_sorted=True)  # partition_by_destination sorts
        if fuse_aggregate:
            e_new, agg = cell.edge_network.edge_step(x_full, e_local, gp.by_src, gp.by_dst)
            if agg is not None:
                return e_new, agg
            return e_new
        return cell.edge_network.fused([x_full, x_full, e_local], [gp.by_src, gp.by_dst, None], skip=2)

    def segment_sum(rows, seg, n):
        return ops.scatter_add(rows, seg, dim_size=n)

    return node_fn, edge_fn, segment_sum
