"""Stage-II critic -- drop-in for the reference's ``discriminator_2.py``.

img [B,3,256,256], tem [B,512] -> [B,1]: conv(3->16,bias)+LeakyReLU(0.1), five conv(k4,s2,p1)+BN+
LeakyReLU(0.1) (16->32->64->128->256->512), then the same text head with 160 channels and
Linear(2560->1) (discriminator_2.py:8-25).  The reference's ``forward`` reads an unbound ``x`` at
:28; this implementation feeds ``img`` to the down-sampler, the one-token fix SURVEY.md section 0 records.
"""
from .discrminator_1 import _CriticBase


class StageIIDiscriminator(_CriticBase):
    def __init__(self, tem_size, Nd):
        super().__init__()
        self._build(tem_size, Nd, (16, 32, 64, 128, 256, 512), 160, 256)
