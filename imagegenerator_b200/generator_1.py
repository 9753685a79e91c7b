"""Stage-I generator -- drop-in for the reference's ``generator_1.py``.

[B, c_dim+z_dim] -> [B,3,64,64]: ConvT(k4,s1,p0)+BN+ReLU on the 1x1 input, three
ConvT(k4,s2,p1)+BN+ReLU, ConvT(24->3,k4,s2,p1,bias)+Tanh (generator_1.py:9-22).  Same
``state_dict`` keys (``upsampling.{0..3}.{0,1}.*``, ``upsampling.4.*``).  Every transposed
convolution runs as the data-gradient direction of the implicit-GEMM conv kernels (4 output
parities x 2x2 taps); BN statistics/apply/ReLU and Tanh are CUDA kernels (imagegenerator_b200/csrc).
"""
import torch
from torch import nn

from .layers import ConvParams, Slot, block, no_autograd

G1_CHANNELS = (192, 96, 48, 24)


class StageIGenerator(nn.Module):
    def __init__(self, c_dim, z_dim):
        super().__init__()
        self.c_dim, self.z_dim = c_dim, z_dim
        seq, cin = [], c_dim + z_dim
        for i, co in enumerate(G1_CHANNELS):
            s, p = (1, 0) if i == 0 else (2, 1)
            seq.append(block(ConvParams(cin, co, 4, s, p, transposed=True), co))
            cin = co
        seq.append(ConvParams(cin, 3, 4, 2, 1, bias=True, transposed=True))
        seq.append(Slot())
        self.upsampling = nn.Sequential(*seq)
        self._rt = {}

    def conv_layers(self):
        """[(ConvParams, BNParams|None)] in forward order."""
        out = [(self.upsampling[i][0], self.upsampling[i][1]) for i in range(4)]
        out.append((self.upsampling[4], None))
        return out

    def runtime(self, batch, ops=None):
        from .engine import GenRT, default_ops
        ops = ops or default_ops()
        key = (batch, id(ops))
        if key not in self._rt:
            self._rt[key] = GenRT(ops, self, batch)
        return self._rt[key]

    def forward(self, x):
        rt = self.runtime(x.shape[0])
        rt.refresh_weights()
        rt.set_input(x)
        rt.forward(training=self.training)
        out = torch.empty(x.shape[0], 3, 64, 64, device=x.device, dtype=torch.float32)
        rt.ops.nhwc_to_nchw(rt.out, out)
        return no_autograd(out, self)
