"""Task base of the edge classifier (reference: EdgeClassifier/edge_classifier_base.py).

Keeps the Lightning hook names and the training contract — ``training_step`` =
pT-weighted binary cross-entropy on ``self(batch.x, batch.edge_index)`` with the
AdamW(amsgrad) + StepLR optimiser and manual LR warm-up — so the data-parallel
driver (hierarchicalgnn_b200.parallel) can step it, and the evaluation steps
(``shared_evaluation`` / ``validation_step`` / ``test_step`` with the tracking metrics of
tracking_utils.eval_metrics). Dataset IO (``setup``, the dataloaders) is outside the hot path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ..lightning_compat import LightningModule


def pt_weighting(pt, hparams):
    """w = w_min + (1-w_min) * clip((pt - (ptcut - interval)) / interval, 0, 1) + leak * max(pt - ptcut, 0)
    (edge_classifier_base.py:82-97); NaN pT counts as 0."""
    pt = torch.nan_to_num(pt, nan=0.0)
    lo = hparams["ptcut"] - hparams["pt_interval"]
    cap = hparams["ptcut"]
    ramp = ((pt - lo) / (cap - lo)).clamp(0.0, 1.0)
    return hparams["weight_min"] + (1 - hparams["weight_min"]) * ramp + hparams["weight_leak"] * (pt - cap).clamp(min=0.0)


def balanced_edge_weights(pt, graph, y, hparams):
    """Per-edge weight = sum of the two end-point pT weights, normalised so true and
    fake edges each sum to sigmoid(+-log_weight_ratio) (edge_classifier_base.py:99-111)."""
    w = pt_weighting(pt[graph[0]], hparams) + pt_weighting(pt[graph[1]], hparams)
    y = y.bool()
    ratio = torch.as_tensor(float(hparams["log_weight_ratio"]), device=w.device)
    true_sum = (w * y).sum().clamp(min=1e-30)
    fake_sum = (w * ~y).sum().clamp(min=1e-30)
    return torch.where(y, w / true_sum * torch.sigmoid(ratio), w / fake_sum * torch.sigmoid(-ratio)).float()


def _evaluation_event(batch, device):
    """The event the tracking metrics are computed against: the unmodified event file the batch was cut from when the loader
    recorded its path (``batch.dir``, edge_classifier_base.py:170-176), else the batch itself."""
    path = getattr(batch, "dir", None)
    if path is None:
        return batch
    event = torch.load(path[0] if isinstance(path, (list, tuple)) else path, map_location=torch.device(device), weights_only=False)
    event.pt[event.pid == 0] = 0
    return event


def _original_hits(bipartite_graph, batch):
    """Hit ids of the unmodified event (``inverse_mask`` undoes the loader's removal of isolated hits)."""
    inverse = getattr(batch, "inverse_mask", None)
    if inverse is None:
        return bipartite_graph
    return torch.stack([inverse[bipartite_graph[0]], bipartite_graph[1]], dim=0)


class EdgeClassifierBase(LightningModule):
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

    def get_training_weight(self, batch, graph, y):
        return balanced_edge_weights(batch.pt, graph, y, self.hparams)

    def training_step(self, batch, batch_idx=0):
        scores = self(batch.x, batch.edge_index)
        if self.hparams.get("true_edges", "pid_true_edges") == "modulewise_true_edges":
            keep = (batch.y_pid == 0) | (batch.y == 1)
            graph, y, scores = batch.edge_index[:, keep], batch.y[keep], scores[keep]
        else:
            graph, y = batch.edge_index, batch.y_pid
        weights = self.get_training_weight(batch, graph, y)
        loss = torch.dot(F.binary_cross_entropy(scores, y.float(), reduction="none"), weights)
        self.log("training_loss", loss)
        return loss

    def shared_evaluation(self, batch, batch_idx=0, log=False):
        """Validation / test step body (edge_classifier_base.py:135-191): weighted BCE on the scores, then track candidates =
        connected components of the edges scoring >= ``score_cut`` (cugraph in the reference, the union-find kernel here),
        scored by ``eval_metrics`` against the event. Returns (bipartite_graph[2, V] = (hit, candidate), loss)."""
        from .. import ops
        from ..tracking_utils import eval_metrics
        with torch.no_grad():
            scores = self(batch.x, batch.edge_index)
            if self.hparams.get("true_edges", "pid_true_edges") == "modulewise_true_edges":
                keep = (batch.y_pid == 0) | (batch.y == 1)
                cut_graph, cut_y, cut_scores = batch.edge_index[:, keep], batch.y[keep], scores[keep]
            else:
                cut_graph, cut_y, cut_scores = batch.edge_index, batch.y_pid, scores
            weights = self.get_training_weight(batch, cut_graph, cut_y.bool())
            loss = torch.dot(F.binary_cross_entropy(cut_scores, cut_y.float(), reduction="none"), weights)
            passed = scores >= self.hparams["score_cut"]
            labels = ops.connected_components(batch.edge_index, batch.x.shape[0], passed if bool(passed.any()) else None).long()
            vertex = (labels >= 0).nonzero().squeeze(1)
            bipartite_graph = torch.stack([vertex, labels[vertex]], dim=0)
            metrics = eval_metrics(_original_hits(bipartite_graph, batch), _evaluation_event(batch, self.device),
                                   pt_cut=self.hparams["ptcut"], nhits_cut=self.hparams["n_hits"],
                                   majority_cut=self.hparams["majority_cut"], primary=False)
        if log:
            self.log_dict({**metrics, "val_loss": loss})
        return bipartite_graph, loss

    def validation_step(self, batch, batch_idx=0):
        return self.shared_evaluation(batch, batch_idx, log=True)[1]

    def test_step(self, batch, batch_idx=0):
        return self.shared_evaluation(batch, batch_idx, log=True)[1]

    def optimizer_step(self, epoch=None, batch_idx=None, optimizer=None, optimizer_idx=None, optimizer_closure=None,
                       on_tpu=False, using_native_amp=False, using_lbfgs=False):
        """Manual linear LR warm-up over the first ``warmup`` steps (edge_classifier_base.py:207-236)."""
        warm = self.hparams.get("warmup")
        if warm and self.trainer.global_step < warm:
            scale = min(1.0, float(self.trainer.global_step + 1) / warm)
            for group in optimizer.param_groups:
                group["lr"] = scale * self.hparams["lr"]
        optimizer.step(closure=optimizer_closure)
        optimizer.zero_grad()
