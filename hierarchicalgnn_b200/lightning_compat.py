"""LightningModule compatibility base (SURVEY.md §8b "Lightning API").

pytorch_lightning is used when importable; otherwise a minimal nn.Module base
exposes exactly what the reference models touch: ``save_hyperparameters``,
``hparams`` with item *and* attribute access, ``log`` / ``log_dict``,
``device``, ``trainer.current_epoch`` / ``trainer.global_step``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

try:  # pragma: no cover - not installed in the build image
    from pytorch_lightning import LightningModule as _PLModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    _PLModule = None
    HAVE_LIGHTNING = False


class AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class _MiniTrainer:
    def __init__(self):
        self.current_epoch = 0
        self.global_step = 0


class _CompatModule(nn.Module):
    def __init__(self):
        super().__init__()
        object.__setattr__(self, "_hparams", AttrDict())
        object.__setattr__(self, "trainer", _MiniTrainer())
        object.__setattr__(self, "logged_metrics", {})
        self.sync_free_logging = False

    def save_hyperparameters(self, hparams=None, **kw):
        hp = AttrDict(dict(hparams or {}))
        hp.update(kw)
        object.__setattr__(self, "_hparams", hp)

    @property
    def hparams(self):
        return self._hparams

    def log(self, name, value, *args, **kwargs):
        self.logged_metrics[name] = value

    def log_dict(self, metrics, *args, **kwargs):
        self.logged_metrics.update(metrics)

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")


LightningModule = _PLModule if HAVE_LIGHTNING else _CompatModule
