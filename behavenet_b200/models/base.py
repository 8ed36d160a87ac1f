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


def _abstract(name):
    """A method every concrete encoder / decoder / model has to provide."""
    def method(self, *args, **kwargs):
        raise NotImplementedError('%s.%s' % (type(self).__name__, name))
    method.__name__ = name
    return method


class _Template(nn.Module):
    """What the training code expects of both modules and models: hparams-style constructors that ignore
    extra arguments, a printable architecture, ``build_model`` and ``forward``."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    __str__ = _abstract('__str__')
    build_model = _abstract('build_model')
    forward = _abstract('forward')


class BaseModule(_Template):
    """Template for encoder / decoder modules (reference base.py:12-39)."""

    def _set_trainable(self, flag):
        for p in self.parameters():
            p.requires_grad = flag

    def freeze(self):
        self._set_trainable(False)

    def unfreeze(self):
        self._set_trainable(True)


class BaseModel(_Template):
    """Template for models (reference base.py:42-67)."""

    loss = _abstract('loss')

    def save(self, filepath):
        """state_dict in the reference's file format (base.py:61-63)."""
        torch.save(self.state_dict(), filepath)

    def get_parameters(self):
        """The trainable parameters, as an iterator (handed to Adam at training.py:284)."""
        return (p for p in self.parameters() if p.requires_grad)


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
