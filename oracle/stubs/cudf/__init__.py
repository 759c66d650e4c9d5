"""ORACLE STUB: pure plumbing (zero-copy views in the real package)."""
import torch


class Series:
    def __init__(self, data):
        self.data = torch.as_tensor(data)

    def to_cupy(self):
        return self.data


class DataFrame(dict):
    def __getitem__(self, key):
        if torch.is_tensor(key):  # boolean row mask
            return DataFrame({k: v[key] for k, v in self.items()})
        return dict.__getitem__(self, key)
