"""Stage-II 256x256 outer step timing (BASELINE.json configs[2]): python tools/bench_stage2.py [B] [steps] [mode]."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from imagegenerator_b200.con_augment import ConditioningAugmentation
from imagegenerator_b200.discriminator_2 import StageIIDiscriminator
from imagegenerator_b200.generator_1 import StageIGenerator
from imagegenerator_b200.generator_2 import StageIIGenerator
from imagegenerator_b200.engine2 import Stage2Engine
from imagegenerator_b200.ops import CudaOps

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
graph = (sys.argv[4] != "eager") if len(sys.argv) > 4 else True
torch.manual_seed(42)
ca1, g1 = ConditioningAugmentation(512, 256, 128), StageIGenerator(128, 100)
ca2, d2, g2 = ConditioningAugmentation(512, 256, 128), StageIIDiscriminator(512, 128), StageIIGenerator()
ops = CudaOps(mode)
for kv in filter(None, os.environ.get("SG_OPTS", "").split(",")):      # e.g. SG_OPTS=narrow=0
    ops.set_option(kv.split("=")[0], int(kv.split("=")[1]))
eng = Stage2Engine(ca1, g1, ca2, d2, g2, B, ops=ops)
g = torch.Generator().manual_seed(0)
dev = "cuda"
real = torch.randn(B, 3, 256, 256, generator=g).clamp_(-1, 1).to(dev)
tem = torch.randn(B, 512, generator=g).to(dev)
tem_mis = tem[torch.randperm(B, generator=g).to(dev)].contiguous()
z = torch.randn(5, B, 100, generator=g).to(dev)
e1 = torch.randn(5, B, 128, generator=g).to(dev); e2 = torch.randn(5, B, 128, generator=g).to(dev)
egp = torch.rand(5, B, generator=g).to(dev)
for _ in range(2):
    eng.step(real, tem, tem_mis, z, e1, e2, egp, use_graph=graph)
torch.cuda.synchronize()
ts = []
for _ in range(steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.step(real, tem, tem_mis, z, e1, e2, egp, use_graph=graph); b.record(); b.synchronize()
    ts.append(a.elapsed_time(b))
ms = sum(ts) / len(ts)
F_D2, F_G2, F_G1 = 0.36412e9, 15.14563e9, 0.03207e9
flops = 5 * (F_G1 + 3 * F_G2 + 12 * F_D2 + F_D2) + (2 * F_D2 + 2 * F_G2)   # + dgrad to images in the critic step
print(json.dumps({"metric": "stackgan_stage2_train_images_per_sec", "value": round(B / (ms * 1e-3), 2), "ms_per_step": round(ms, 3),
                  "batch": B, "mode": mode, "launches_per_step": eng.launches_per_step, "cuda_graph": graph,
                  "step_tflops": round(flops * B / (ms * 1e-3) / 1e12, 1), "losses": eng.losses.tolist(),
                  "mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))
