"""Stage-I critic -- drop-in for the reference's ``discrminator_1.py`` (sic: the typo is API).

img [B,3,64,64], tem [B,512] -> [B,1]: conv(3->64,k4,s2,p1,bias)+LeakyReLU(0.1), three
conv(k4,s2,p1)+BN+LeakyReLU(0.1), text embedding compressed to Nd, replicated 4x4 and
concatenated, 1x1 conv to 128 channels, flatten, linear to one score (discrminator_1.py:9-52).
Same ``state_dict`` keys.  The convs are implicit-GEMM CUDA kernels; the replicate+concat+1x1+linear
head is affine in (features, compressed text) and is evaluated without materialising the concat.
"""
import torch
from torch import nn

from .layers import ConvParams, DenseParams, Slot, block, no_autograd


class _CriticBase(nn.Module):
    def _build(self, tem_size, Nd, chs, head_ch, in_hw):
        self.tem_size, self.Nd, self.chs, self.head_ch, self.in_hw = tem_size, Nd, tuple(chs), head_ch, in_hw
        seq = [ConvParams(3, chs[0], 4, 2, 1, bias=True), Slot()]
        cin = chs[0]
        for co in chs[1:]:
            seq.append(block(ConvParams(cin, co, 4, 2, 1), co))
            cin = co
        self.down_sampler = nn.Sequential(*seq)
        self.compress = DenseParams(tem_size, Nd)
        self.channel_resize = ConvParams(cin + Nd, head_ch, 1, 1, 0, bias=True)
        self.critic_score = DenseParams(head_ch * 16, 1)
        self._rt = {}

    def conv_layers(self):
        out = [(self.down_sampler[0], None)]
        for j in range(2, 2 + len(self.chs) - 1):
            out.append((self.down_sampler[j][0], self.down_sampler[j][1]))
        return out

    def runtime(self, batch, ops=None):
        from .engine import CriticRT, default_ops
        ops = ops or default_ops()
        key = (batch, id(ops))
        if key not in self._rt:
            self._rt[key] = CriticRT(ops, self, batch)
        return self._rt[key]

    def forward(self, img, tem):
        rt = self.runtime(img.shape[0])
        rt.refresh_weights()
        rt.ops.nchw_to_nhwc(img.contiguous().float(), rt.group_view(rt.a[0], 0, 1))
        rt.set_text(tem.contiguous().float(), None)
        rt.forward(0, 1, dup_first=1, training=self.training)
        return no_autograd(rt.score[0].clone().reshape(-1, 1), self)


class StageIDiscriminator(_CriticBase):
    def __init__(self, tem_size, Nd):
        super().__init__()
        self._build(tem_size, Nd, (64, 128, 256, 512), 128, 64)
