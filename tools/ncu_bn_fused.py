"""Two launches each of the one-launch BatchNorm kernels on a Stage-I critic tensor (12.6 MB, 3 groups / 4.2 MB, 1 group), for
   ncu --set full --clock-control none --import-source on -k regex:fused8 -c 4 -o gpurun_out/bn_fused python tools/ncu_bn_fused.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from imagegenerator_b200.ops import CudaOps, ACT_LRELU
ops = CudaOps("bf16")
ops.set_option("bn_fused", 1)
rpg, C, G = 128 * 64, 256, 3
mk = lambda n: (torch.randn(n, C, device="cuda")).to(torch.bfloat16)
y, da, dy = mk(rpg * G), mk(rpg * G), mk(rpg * G)
mr = torch.rand(G, C, 2, device="cuda") + 0.5
gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
sums = torch.zeros(G, C, 2, dtype=torch.float64, device="cuda")
v, a, w, gy = mk(rpg), mk(rpg), mk(rpg), mk(rpg)
ts, dg = torch.zeros(C, 3, dtype=torch.float64, device="cuda"), torch.zeros(C, device="cuda")
for _ in range(2):
    ops.bn_bwd(da, None, y, mr, gamma, sums, dy, G, ACT_LRELU, beta=beta)
    ops.gp_bn(v, da[:rpg], a, y[:rpg], mr[:1], gamma, sums[:1], ts, w, gy, dg, ACT_LRELU)
torch.cuda.synchronize()
