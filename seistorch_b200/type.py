"""TensorList -- the container WaveRNN.forward returns (seistorch/type.py:7-97)."""
from __future__ import annotations

import torch

from .utils import to_tensor


class TensorList(list):
    """A list of torch.Tensors with the reference's helper methods."""

    def __init__(self, input_list=()):
        super().__init__()
        self.data = []
        for item in input_list:
            self.append(item)

    @property
    def device(self):
        return self.data[0].device

    @property
    def shape(self):
        return (len(self.data),)

    def append(self, item):
        self.data.append(item if isinstance(item, torch.Tensor) else to_tensor(item))

    def cuda(self):
        self.data = [t.cuda() for t in self.data]
        return self

    def has_nan(self):
        """type.py:41-46: raises ValueError when any record contains NaN."""
        for t in self.data:
            if isinstance(t, torch.Tensor) and torch.isnan(t).any():
                raise ValueError("The tensor list contains NaN values.")
        return False

    def numpy(self):
        self.data = [t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t for t in self.data]
        return self

    def stack(self):
        """type.py:54-57: zero-pad every record to the largest (nt, nrec) then stack."""
        max_shape = max([t.shape for t in self.data])
        padded = [torch.nn.functional.pad(t, (0, max_shape[1] - t.shape[1], 0, max_shape[0] - t.shape[0]))
                  for t in self.data]
        return torch.stack(padded, dim=0)

    def tensor(self):
        return self.data

    def to(self, device):
        self.data = [t.to(device) for t in self.data]
        return self

    def tolist(self):
        return self

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return self.data[index]

    def __iter__(self):
        return iter(self.data)

    def __mul__(self, other):
        if isinstance(other, TensorList) and len(self.data) == len(other.data):
            return TensorList([a * b for a, b in zip(self.data, other.data)])
        raise ValueError("Multiplication is only defined between two instances of TensorList with the same length.")

    def __pow__(self, exponent):
        self.data = [t ** exponent for t in self.data]
        return self

    def __str__(self):
        return str(self.data)
