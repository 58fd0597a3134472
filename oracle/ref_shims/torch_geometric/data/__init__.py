from . import data  # noqa: F401


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def __contains__(self, key):
        return hasattr(self, key)

    def to(self, device):
        import torch

        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self


class Batch(Data):
    pass


class Dataset:
    pass
