"""Run Stage-I outer steps eagerly (no CUDA graph) so ncu sees every kernel launch."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import build_modules
from imagegenerator_b200.engine import Stage1Engine
from imagegenerator_b200.ops import CudaOps

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
ops = CudaOps(mode)
ca, d1, g1 = build_modules()
eng = Stage1Engine(ca, d1, g1, B, ops=ops)
g = torch.Generator().manual_seed(0)
dev = "cuda"
real = torch.randn(B, 3, 64, 64, generator=g).clamp_(-1, 1).to(dev)
tem = torch.randn(B, 512, generator=g).to(dev)
tem_mis = tem[torch.randperm(B, generator=g).to(dev)].contiguous()
z = torch.randn(5, B, 100, generator=g).to(dev); eca = torch.randn(5, B, 128, generator=g).to(dev); egp = torch.rand(5, B, generator=g).to(dev)
for s in range(steps):
    n0 = ops.launch_count()
    eng.step(real, tem, tem_mis, z, eca, egp, use_graph=False)
    torch.cuda.synchronize()
    print("step", s, "launches", ops.launch_count() - n0, "losses", eng.losses.tolist())
