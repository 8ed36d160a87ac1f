"""Module templates with the reference's protocol (reference: behavenet/models/base.py).

``BaseModel.save`` / ``get_parameters`` are what ``fitting/training.py:284,390`` call;
``DiagLinear`` is the label head of the PS-VAE (base.py:70-103).  ``CustomDataParallel`` (the
reference's inert nn.DataParallel wrapper, base.py:106-116) is replaced by one process per GPU
with an NCCL all-reduce, see ``behavenet_b200.parallel``.
"""

import math

import torch
from torch import nn

__all__ = ['BaseModule', 'BaseModel', 'DiagLinear']


class BaseModule(nn.Module):
    """Template for encoder / decoder modules."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    def __str__(self):
        raise NotImplementedError

    def build_model(self):
        raise NotImplementedError

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def freeze(self):
        for p in self.parameters():
            p.requires_grad = False

    def unfreeze(self):
        for p in self.parameters():
            p.requires_grad = True


class BaseModel(nn.Module):
    """Template for models."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    def __str__(self):
        raise NotImplementedError

    def build_model(self):
        raise NotImplementedError

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def loss(self, *args, **kwargs):
        raise NotImplementedError

    def save(self, filepath):
        """Save the state_dict (same file format as the reference, base.py:61-63)."""
        torch.save(self.state_dict(), filepath)

    def get_parameters(self):
        """Parameters with gradient updates turned on (consumed by Adam, training.py:284)."""
        return filter(lambda p: p.requires_grad, self.parameters())


class DiagLinear(nn.Module):
    """y = x * w + b with a diagonal weight (state_dict keys ``weight``, ``bias``)."""

    def __init__(self, features, bias=True):
        super().__init__()
        self.features = features
        bound = 1 / math.sqrt(features)
        self.weight = nn.Parameter(torch.empty(features).uniform_(-bound, bound))
        if bias:
            self.bias = nn.Parameter(torch.empty(features).uniform_(-bound, bound))
        else:
            self.register_parameter('bias', None)

    def forward(self, input):
        out = input * self.weight
        return out if self.bias is None else out + self.bias

    def extra_repr(self):
        return 'features={}, bias={}'.format(self.features, self.bias is not None)
