"""Debug tool: run one critic iteration on CUDA and on the fp64 emulator, compare every buffer."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import stackgan_oracle as O
from emu_ops import EmuOps
from imagegenerator_b200.con_augment import ConditioningAugmentation
from imagegenerator_b200.discrminator_1 import StageIDiscriminator
from imagegenerator_b200.generator_1 import StageIGenerator
from imagegenerator_b200.engine import Stage1Engine
from imagegenerator_b200.ops import CudaOps

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
its = int(sys.argv[3]) if len(sys.argv) > 3 else 1

def mk():
    torch.manual_seed(42)
    return ConditioningAugmentation(512, 256, 128), StageIDiscriminator(512, 128), StageIGenerator(128, 100)

b = O.synthetic_batch(B, 1, 0)
engs = []
for ops in (EmuOps(torch.float64), CudaOps(mode)):
    ca, d1, g1 = mk()
    eng = Stage1Engine(ca, d1, g1, B, ops=ops)
    dv = lambda t: t.to(ops.device).to(ops.f32).contiguous()
    eng.load_batch(dv(b["real"]), dv(b["tem"]), dv(b["tem"][b["perm"]]))
    for it in range(its):
        eng.critic_iteration(dv(b["z"][it]), dv(b["eps_ca"][it]), dv(b["eps_gp"][it]))
    engs.append((eng, d1))
torch.cuda.synchronize()
(e0, m0), (e1, m1) = engs

def rep(name, a, c):
    a = a.double().cpu(); c = c.double().cpu()
    den = a.norm().item()
    print(f"{name:28s} rel_l2 {(a - c).norm().item() / max(den, 1e-300):10.3e}   norm {den:10.3e}  maxerr {(a-c).abs().max().item():.3e}")

d0, d1_ = e0.d, e1.d
for name in ["a", "y", "mr", "da", "dy", "gda", "gdy", "gsums", "sums", "v", "tsums", "w", "gy"]:
    L0, L1 = getattr(d0, name), getattr(d1_, name)
    for i, (x, y) in enumerate(zip(L0, L1)):
        if x is not None:
            rep(f"{name}[{i}]", x, y)
for name in ["g", "sq", "v0", "A", "dA", "Bv", "dBv", "score", "ce"]:
    rep(name, getattr(d0, name), getattr(d1_, name))
print("losses", e0.losses.tolist(), e1.losses.tolist())
for (k, p0), (_, p1) in zip(m0.named_parameters(), m1.named_parameters()):
    rep("grad " + k, p0.grad, p1.grad)
