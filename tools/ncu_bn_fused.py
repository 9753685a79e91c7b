import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from imagegenerator_b200.ops import CudaOps, ACT_LRELU
ops = CudaOps("bf16")
rpg, C, G = 128 * 64, 256, 3
n = rpg * G
mk = lambda: (torch.randn(n, C, device="cuda")).to(torch.bfloat16)
y, da, dy = mk(), mk(), mk()
mr = torch.rand(G, C, 2, device="cuda") + 0.5
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
sums = torch.zeros(G, C, 2, dtype=torch.float64, device="cuda")
for _ in range(2):
    ops.bn_bwd_reduce(da, None, y, mr, sums, G, ACT_LRELU, gamma=gamma, beta=beta)
    ops.bn_bwd_apply(da, None, y, mr, gamma, sums, dy, G, ACT_LRELU, beta=beta)
    ops.bn_bwd(da, None, y, mr, gamma, sums, dy, G, ACT_LRELU, beta=beta)
torch.cuda.synchronize()
