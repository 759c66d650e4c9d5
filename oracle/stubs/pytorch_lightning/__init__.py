"""ORACLE STUB: the subset of pytorch_lightning 1.6.3 LightningModule the
reference touches (save_hyperparameters / hparams item access / log /
log_dict / device / trainer)."""
import torch


class _Trainer:
    current_epoch = 0
    global_step = 0


class LightningModule(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self._hp = {}
        self.trainer = _Trainer()
        self.logged = {}

    def save_hyperparameters(self, hparams):
        self._hp = dict(hparams)

    @property
    def hparams(self):
        return self._hp

    def log(self, name, value, *a, **k):
        self.logged[name] = value

    def log_dict(self, d, *a, **k):
        self.logged.update(d)

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")
