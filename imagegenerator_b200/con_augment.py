"""Conditioning augmentation -- drop-in for the reference's ``con_augment.py``.

Same constructor and ``state_dict`` keys (``h``, ``mu``, ``sigma`` dense layers, con_augment.py:7-11);
``forward(tem)`` returns ``(c_hat, mu, sigma)`` like con_augment.py:18-22, and additionally accepts
the reparameterisation noise ``eps`` (the reference draws it from the global RNG at :20; ``None``
keeps that behaviour).  ``sigma`` is the raw linear output used as a standard deviation.
The arithmetic runs in the CUDA kernels behind ``imagegenerator_b200.ops`` (no CPU path).
"""
import torch
from torch import nn

from .layers import DenseParams, no_autograd


class ConditioningAugmentation(nn.Module):
    def __init__(self, tem_size, h_dim, c_dim):
        super().__init__()
        self.tem_size, self.h_dim, self.c_dim = tem_size, h_dim, c_dim
        self.h = DenseParams(tem_size, h_dim)
        self.mu = DenseParams(h_dim, c_dim)
        self.sigma = DenseParams(h_dim, c_dim)
        self._rt = None

    def runtime(self, ops=None):
        from .engine import CART, default_ops
        if self._rt is None or (ops is not None and self._rt.ops is not ops):
            self._rt = CART(ops or default_ops(), self)
        return self._rt

    def encode(self, tem):
        rt = self.runtime()
        st = rt.forward(tem, None, None)
        return st.mu.clone(), st.sigma.clone()

    def forward(self, tem, eps=None):
        rt = self.runtime()
        if eps is None:
            eps = torch.randn(tem.shape[0], self.c_dim, device=tem.device, dtype=torch.float32)
        st = rt.forward(tem, eps, None)
        return tuple(no_autograd(t.clone(), self) for t in (st.c_hat, st.mu, st.sigma))
