"""One launch set of the HBM-bound BatchNorm kernels on the G2 up2 activation (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from imagegenerator_b200.ops import CudaOps
ops = CudaOps("bf16")
rows, C = 64 * 128 * 128, 80
mk = lambda: torch.randn(rows, C, device="cuda").to(torch.bfloat16)
da, a, y, dy = mk(), mk(), mk(), mk()
mr = torch.rand(1, C, 2, device="cuda") + 0.5
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
sums = torch.zeros(1, C, 2, dtype=torch.float64, device="cuda")
for _ in range(3):
    ops.bn_act(y, mr, gamma, beta, a, 1, 1)
    ops.bn_bwd_reduce(da, a, y, mr, sums, 1, 1)
    ops.bn_bwd_apply(da, a, y, mr, gamma, sums, dy, 1, 1)
torch.cuda.synchronize()
print("ok")
