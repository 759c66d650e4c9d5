"""ORACLE STUB: placeholders for dataset/evaluation-only names."""


class Data:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    @classmethod
    def from_dict(cls, d):
        return cls(**d)

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)


class DataLoader:
    def __init__(self, dataset, **kw):
        self.dataset = dataset

    def __iter__(self):
        return iter(self.dataset)
