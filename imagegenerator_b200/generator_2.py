"""Stage-II generator -- drop-in for the reference's ``generator_2.py``.

img_64 [B,3,64,64], c_hat [B,128] -> [B,3,256,256] (generator_2.py:59-67): conv(3->128,k4,s2,p1,bias)+
LeakyReLU(0.1), conv(128->512,k4,s2,p1)+BN+LeakyReLU(0.1), c_hat replicated 16x16 and concatenated
(640 channels), four ``ResidualBlock(640, 320)`` (:5-39: three conv3x3+BN with ReLU between, identity
added before the last ReLU), three ConvT(k4,s2,p1)+BN+ReLU (640->320->160->80) and
ConvT(80->3,bias)+Tanh.  Same ``state_dict`` keys (``down_sampler.*``, ``residual_blocks.{r}.layer{1,2,3}.*``,
``up_sampler.*``).  All arithmetic runs in the CUDA kernels of libsgb200 (3x3 and strided convs on the
tcgen05 implicit-GEMM kernel, ConvT as its data-gradient direction).
"""
import torch
from torch import nn

from .layers import ConvParams, BNParams, Slot, block, no_autograd


class ResidualBlock(nn.Module):
    def __init__(self, in_channels, intermediate_channels):
        super().__init__()
        mk = lambda ci, co: nn.Sequential(ConvParams(ci, co, 3, 1, 1), BNParams(co))
        self.layer1 = mk(in_channels, intermediate_channels)
        self.layer2 = mk(intermediate_channels, intermediate_channels)
        self.layer3 = mk(intermediate_channels, in_channels)
        self.relu = Slot()

    def conv_layers(self):
        return [(self.layer1[0], self.layer1[1]), (self.layer2[0], self.layer2[1]), (self.layer3[0], self.layer3[1])]


class StageIIGenerator(nn.Module):
    C_TEXT = 128

    def __init__(self):
        super().__init__()
        self.down_sampler = nn.Sequential(ConvParams(3, 128, 4, 2, 1, bias=True), Slot(),
                                          block(ConvParams(128, 512, 4, 2, 1), 512))
        self.residual_blocks = nn.Sequential(*[ResidualBlock(640, 320) for _ in range(4)])
        ups, cin = [], 640
        for co in (320, 160, 80):
            ups.append(block(ConvParams(cin, co, 4, 2, 1, transposed=True), co))
            cin = co
        ups.append(ConvParams(cin, 3, 4, 2, 1, bias=True, transposed=True))
        ups.append(Slot())
        self.up_sampler = nn.Sequential(*ups)
        self._rt = {}

    def runtime(self, batch, ops=None):
        from .engine2 import Gen2RT
        from .engine import default_ops
        ops = ops or default_ops()
        key = (batch, id(ops))
        if key not in self._rt:
            self._rt[key] = Gen2RT(ops, self, batch)
        return self._rt[key]

    def forward(self, img_64, c_hat):
        B = img_64.shape[0]
        rt = self.runtime(B)
        rt.refresh_weights()
        rt.ops.nchw_to_nhwc(img_64.contiguous().float(), rt.x_in)
        rt.forward(c_hat.contiguous().float(), training=self.training)
        out = torch.empty(B, 3, 256, 256, device=img_64.device, dtype=torch.float32)
        rt.ops.nhwc_to_nchw(rt.out, out)
        return no_autograd(out, self)
