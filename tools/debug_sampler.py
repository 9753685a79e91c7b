import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import stackgan_oracle as O
from test_engine2_emulated import build_all
from test_sampler_emulated import _randomise_running_stats
from emu_ops import EmuOps
from imagegenerator_b200.ops import CudaOps
from imagegenerator_b200.sampler import StackGANSampler
B = 4; dt = torch.float64
def make(ops):
    ms = build_all(); ps = O.init_all(42)
    p = {k: O.to_dtype(ps[k], dt) for k in ps}
    _randomise_running_stats(p["gen_1"], 3, dt); _randomise_running_stats(p["gen_2"], 4, dt)
    for key, m in (("gen_1", "g1"), ("gen_2", "g2"), ("con_augment_1", "ca1"), ("con_augment_2", "ca2")):
        ms[m].load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in p[key].items()})
    return StackGANSampler(ms["ca1"], ms["g1"], ms["ca2"], ms["g2"], B, ops=ops, bn_batch_stats=True), p
g = torch.Generator().manual_seed(0)
tem = torch.randn(B, 512, generator=g); z, e1, e2 = (torch.randn(B, n, generator=g) for n in (100, 128, 128))
sc, p = make(CudaOps("bf16")); se, _ = make(EmuOps(torch.bfloat16))
sc.sample(tem, z, e1, e2, use_graph=False); se.sample(tem, z, e1, e2)
torch.cuda.synchronize()
ref64, ref256 = O.sample(p["con_augment_1"], p["gen_1"], p["con_augment_2"], p["gen_2"], tem.double(), z.double(), e1.double(), e2.double(), g2_training=True)
rl = lambda a, b: ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()
print("fake_64  cuda-vs-emu(bf16) %.3e  cuda-vs-fp64 %.3e  emu-vs-fp64 %.3e" % (rl(sc.out_64, se.out_64), rl(sc.out_64, ref64), rl(se.out_64, ref64)))
print("fake_256 cuda-vs-emu(bf16) %.3e  cuda-vs-fp64 %.3e  emu-vs-fp64 %.3e" % (rl(sc.out_256, se.out_256), rl(sc.out_256, ref256), rl(se.out_256, ref256)))
a, b = sc.g2rt, se.g2rt
names = [("a1", a.a1, b.a1), ("b2.y", a.b2.y, b.b2.y), ("b2.a", a.b2.a, b.b2.a)]
for r in range(4):
    for j in range(3):
        names.append((f"rb{r}.{j}.y", a.rb[r][j].y, b.rb[r][j].y))
    names.append((f"X{r+1}", a.X[r + 1], b.X[r + 1]))
for i in range(3):
    names += [(f"ub{i}.y", a.ub[i].y, b.ub[i].y), (f"ub{i}.a", a.ub[i].a, b.ub[i].a)]
for n, x, y in names:
    print(f"{n:10s} rel-L2 cuda vs ideal-bf16 emu: {rl(x, y):.3e}   max|ref| {y.abs().max().item():.3e}")
