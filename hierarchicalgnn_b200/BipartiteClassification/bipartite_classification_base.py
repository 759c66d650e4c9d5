"""Task base of the bipartite classifier (reference:
BipartiteClassification/bipartite_classification_base.py).

Hook names, optimiser, and the two-term loss (pT-weighted hinge embedding loss
+ assignment BCE against a minimum-weight particle<->supernode matching, sine
loss schedule) follow the reference. The matching itself stays scipy on the
host exactly as in the reference (loss-side CPU stage, next
in line, not on the message-passing path).
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn.functional as F

from ..EdgeClassifier.edge_classifier_base import balanced_edge_weights, pt_weighting
from ..lightning_compat import LightningModule


class BipartiteClassificationBase(LightningModule):
    def __init__(self, hparams):
        super().__init__()
        self.save_hyperparameters(hparams)

    def configure_optimizers(self):
        params = list(self.parameters())
        # same optimiser and hyper-parameters as the reference; on CUDA parameters the fused multi-tensor implementation
        # (one launch per step instead of a dozen per parameter group chunk)
        fused = bool(params) and all(p.is_cuda for p in params)
        opt = torch.optim.AdamW(params, lr=self.hparams["lr"], betas=(0.9, 0.999), eps=1e-08, amsgrad=True,
                                **({"fused": True} if fused else {}))
        sched = torch.optim.lr_scheduler.StepLR(opt, step_size=self.hparams["patience"], gamma=self.hparams["factor"])
        return [opt], [{"scheduler": sched, "interval": "epoch", "frequency": 1}]

    # ---- embedding (hinge) term: bipartite_classification_base.py:141-150,196-204 ----
    def embedding_loss(self, batch, embeddings):
        graph = batch.edge_index
        y = batch.pid[graph[0]] == batch.pid[graph[1]]
        weights = balanced_edge_weights(batch.pt, graph, y, self.hparams)
        hinge = torch.where(y, 1, -1)
        dist = ((embeddings[graph[0]] - embeddings[graph[1]]).square().sum(-1) + 1e-12).sqrt()
        loss = F.hinge_embedding_loss(dist / self.hparams["train_r"], hinge, margin=1, reduction="none").square()
        return torch.dot(loss, weights)

    # ---- assignment term: bipartite_classification_base.py:152-191 ----
    def assignment_loss(self, batch, bipartite_graph, bipartite_scores):
        from scipy.sparse.csgraph import min_weight_full_bipartite_matching
        dev = bipartite_scores.device
        original_pid, pid = torch.unique(batch.pid, return_inverse=True)
        event = getattr(batch, "batch", None)
        row_ptr = None
        if event is not None:
            # a batch of events: particle ids are per event, so a particle is an (event, pid) pair. The score table is then
            # block diagonal (supernodes never span events): the matching is run block by block (its cost grows faster than
            # linearly with the table: 4 events matched as one table took 5x the time of 4 single matchings)
            n_ids = original_pid.numel()
            pair, pid = torch.unique(event * n_ids + pid, return_inverse=True)
            original_pid = original_pid[pair % n_ids]
            n_events = getattr(batch, "num_graphs", None) or int(event[-1]) + 1
            edges = torch.arange(n_events + 1, device=dev) * n_ids
            row_ptr = torch.searchsorted(pair, edges).tolist()  # particles (table rows) of event b: [row_ptr[b], row_ptr[b+1])
        n_p = int(pid.max()) + 1
        n_s = int(bipartite_graph[1].max()) + 1
        pt = torch.full((n_p,), float("inf"), device=dev).scatter_reduce(0, pid, batch.pt.float(), "amin")
        with torch.no_grad():
            # virtual supernodes (one per particle, epsilon score) guarantee a full matching exists
            rows = torch.cat([pid[bipartite_graph[0]], torch.arange(n_p, device=dev)])
            cols = torch.cat([bipartite_graph[1], torch.arange(n_s, n_s + n_p, device=dev)])
            vals = torch.cat([bipartite_scores.detach(), torch.full((n_p,), 1e-12, device=dev)])
            table = self._score_table(rows, cols, vals, n_p, n_s + n_p)
            if row_ptr is None:
                rm, cm = min_weight_full_bipartite_matching(table, maximize=True)
            else:
                rm, cm = self._match_blocks(table, row_ptr)
            rm, cm = torch.as_tensor(rm, device=dev).long(), torch.as_tensor(cm, device=dev).long()
            real = (original_pid[rm] != 0) & (cm < n_s)
            rm, cm = rm[real], cm[real]
            assigned = torch.full((n_p,), -1, dtype=torch.long, device=dev)
            assigned[rm] = cm
            truth = assigned[pid[bipartite_graph[0]]] == bipartite_graph[1]
            sn_pt = torch.zeros(n_s, device=dev)
            sn_pt[cm] = pt[rm]
            w = torch.maximum(pt_weighting(batch.pt[bipartite_graph[0]], self.hparams),
                              pt_weighting(sn_pt[bipartite_graph[1]], self.hparams))
            ratio = torch.as_tensor(float(self.hparams["log_weight_ratio"]), device=dev)
            ts, fs = (w * truth).sum().clamp(min=1e-30), (w * ~truth).sum().clamp(min=1e-30)
            w = torch.where(truth, w / ts * torch.sigmoid(ratio), w / fs * torch.sigmoid(-ratio)).float()
        return torch.dot(F.binary_cross_entropy(bipartite_scores, truth.float(), reduction="none"), w)

    @staticmethod
    def _match_blocks(table, row_ptr):
        """min_weight_full_bipartite_matching(table, maximize=True) of a block-diagonal CSR table, block b = rows
        [row_ptr[b], row_ptr[b+1]): the blocks are independent assignment problems, solved side by side on host threads by
        hgnn_match_blocks_max (csrc/matching.cu; scipy holds the GIL and, given the whole table, takes 5x the time of the
        blocks one by one). Returns (rows, columns) of the whole table, like scipy."""
        import numpy as np
        from .. import _lib
        table.sort_indices()
        n = table.shape[0]
        indptr = np.ascontiguousarray(table.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(table.indices, dtype=np.int32)
        data = np.ascontiguousarray(table.data, dtype=np.float32)
        ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        cols = np.empty(n, dtype=np.int64)
        # 0 = one thread per block up to the hardware threads. (Dividing the cores between the ranks of a box was measured
        # slower at 8 ranks on 16 cores — 94.5 vs 88.5 ms per training step: the ranks' matchings do not coincide.)
        threads = int(os.environ.get("HGNN_MATCH_THREADS", "0"))
        _lib.check(_lib.lib().hgnn_match_blocks_max(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, n, ptr.ctypes.data,
                                                    ptr.shape[0] - 1, cols.ctypes.data, threads), "match_blocks_max")
        return np.arange(n, dtype=np.int64), cols

    @staticmethod
    def _score_table(rows, cols, vals, n_rows, n_cols):
        """The particle x (supernode + virtual supernode) score table of the matching as a canonical scipy CSR matrix.
        The reference hands scipy the COO triplets and lets it sort them and sum the duplicates on the host
        (bipartite_classification_base.py:163-173: 3 ms of the 7 ms the matching stage takes on a 1 GeV event); on CUDA
        tensors the same table is assembled on the device — sorted unique (row, column) keys, duplicate scores summed per key
        in a fixed order (segment reduce, no float atomics), row pointers from a histogram — and scipy receives
        (data, indices, indptr) as they are."""
        from scipy.sparse import csr_matrix
        if not vals.is_cuda:
            return csr_matrix((vals.numpy(), (rows.numpy(), cols.numpy())), shape=(n_rows, n_cols))
        from .. import ops
        keys = rows * n_cols + cols
        uniq, inverse = torch.unique(keys, return_inverse=True)
        data = ops.scatter_add(vals.float().unsqueeze(1).contiguous(), inverse, dim_size=uniq.numel()).squeeze(1)
        r = torch.div(uniq, n_cols, rounding_mode="floor")
        indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=vals.device)
        indptr[1:] = torch.bincount(r, minlength=n_rows).cumsum(0)
        idx = torch.cat([(uniq - r * n_cols), indptr]).to(torch.int32).cpu().numpy()  # one transfer for both index arrays
        m = csr_matrix((data.cpu().numpy(), idx[:uniq.numel()], idx[uniq.numel():]), shape=(n_rows, n_cols))
        m.has_sorted_indices = True
        return m

    def loss_schedule(self):
        if self.hparams.get("loss_schedule") is not None:
            return self.hparams["loss_schedule"]
        ep, emb_ep = self.trainer.current_epoch, self.hparams["emb_epoch"]
        return 1 - math.sin(ep / 2 / emb_ep * math.pi) if ep < emb_ep else 0

    def _forward_batch(self, batch):
        """self(x, edge_index), plus the event vector when the loader collated several events into one batch (the
        reference's loaders use batch_size=1, bipartite_classification_base.py:42, and its forward has no such argument)."""
        event = getattr(batch, "batch", None)
        if event is None:
            return self(batch.x, batch.edge_index)
        return self(batch.x, batch.edge_index, batch=event, n_events=getattr(batch, "num_graphs", None))

    def training_step(self, batch, batch_idx=0):
        bipartite_graph, bipartite_scores, embeddings = self._forward_batch(batch)
        emb_loss = self.embedding_loss(batch, embeddings)
        asgmt_loss = self.assignment_loss(batch, bipartite_graph, bipartite_scores)
        s = self.loss_schedule()
        loss = s * emb_loss + (1 - s) * asgmt_loss
        self.log_dict({"training_loss": loss, "embedding_loss": emb_loss, "assignment_loss": asgmt_loss})
        return loss

    def shared_evaluation(self, batch, batch_idx=0, log=False):
        """Validation / test step body (bipartite_classification_base.py:226-287): both loss terms (the schedule counts only
        while training, else the assignment term alone), then the hit -> supernode assignments scoring >= ``score_cut`` as
        track candidates for ``eval_metrics``. Returns (bipartite_graph, loss)."""
        from ..EdgeClassifier.edge_classifier_base import _evaluation_event, _original_hits
        from ..tracking_utils import default_response, eval_metrics
        with torch.no_grad():
            bipartite_graph, bipartite_scores, embeddings = self._forward_batch(batch)
            emb_loss = self.embedding_loss(batch, embeddings)
            asgmt_loss = self.assignment_loss(batch, bipartite_graph, bipartite_scores)
            s = self.loss_schedule() if (self.training and hasattr(self.trainer, "current_epoch")) else 0
            loss = s * emb_loss + (1 - s) * asgmt_loss
            self.log_dict({"val_loss": loss, "val_embedding_loss": emb_loss, "val_assignment_loss": asgmt_loss})
            bipartite_graph = bipartite_graph[:, bipartite_scores >= self.hparams["score_cut"]]
            try:
                metrics = eval_metrics(_original_hits(bipartite_graph, batch), _evaluation_event(batch, self.device),
                                       pt_cut=self.hparams["ptcut"], nhits_cut=self.hparams["n_hits"],
                                       majority_cut=self.hparams["majority_cut"], primary=False)
            except (RuntimeError, IndexError, ValueError):  # the reference falls back to zeros on any failure here
                metrics = dict(default_response)
        if log:
            self.log_dict(metrics)
        return bipartite_graph, loss

    def validation_step(self, batch, batch_idx=0):
        return self.shared_evaluation(batch, batch_idx, log=True)[1]

    def test_step(self, batch, batch_idx=0):
        return self.shared_evaluation(batch, batch_idx, log=True)[1]

    def optimizer_step(self, epoch=None, batch_idx=None, optimizer=None, optimizer_idx=None, optimizer_closure=None,
                       on_tpu=False, using_native_amp=False, using_lbfgs=False):
        warm = self.hparams.get("warmup")
        if warm and self.trainer.global_step < warm:
            scale = min(1.0, float(self.trainer.global_step + 1) / warm)
            for group in optimizer.param_groups:
                group["lr"] = scale * self.hparams["lr"]
        optimizer.step(closure=optimizer_closure)
        optimizer.zero_grad()
