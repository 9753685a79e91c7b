"""Per-op timing of one outer step, labelled by op name and tensor shapes.

    python tools/profile_ops.py {1|2} [B] [top]

Every ``CudaOps`` call of one eager outer step is bracketed by CUDA events on the launch stream (the GPU
is kept busy by the previous calls, so an event pair measures the kernel, not the launch).  Lines are
aggregated by (op, shapes) and sorted by total time."""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from imagegenerator_b200.ops import CudaOps  # noqa: E402


class TimedOps:
    def __init__(self, ops):
        self._ops, self.rec, self.on = ops, [], False

    def __getattr__(self, name):
        attr = getattr(self._ops, name)
        if not callable(attr) or name in ("empty", "zeros", "launch_count", "set_option", "_st", "_ck", "_c", "_dt_of"):
            return attr

        def wrapped(*a, **k):
            if not self.on:
                return attr(*a, **k)
            key = name + " " + " ".join("x".join(map(str, t.shape)) for t in a if torch.is_tensor(t) and t.dim() >= 3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = attr(*a, **k)
            e1.record()
            self.rec.append((key, e0, e1))
            return r
        return wrapped


def main():
    stage = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (128 if stage == 1 else 64)
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    from imagegenerator_b200.con_augment import ConditioningAugmentation
    from imagegenerator_b200.generator_1 import StageIGenerator
    torch.manual_seed(42)
    ops = TimedOps(CudaOps("bf16"))
    g = torch.Generator().manual_seed(0)
    dev = "cuda"
    tem = torch.randn(B, 512, generator=g).to(dev)
    tem_mis = tem[torch.randperm(B, generator=g).to(dev)].contiguous()
    z = torch.randn(5, B, 100, generator=g).to(dev)
    e1 = torch.randn(5, B, 128, generator=g).to(dev)
    e2 = torch.randn(5, B, 128, generator=g).to(dev)
    egp = torch.rand(5, B, generator=g).to(dev)
    if stage == 1:
        from imagegenerator_b200.discrminator_1 import StageIDiscriminator
        from imagegenerator_b200.engine import Stage1Engine
        ca, d1, g1 = ConditioningAugmentation(512, 256, 128), StageIDiscriminator(512, 128), StageIGenerator(128, 100)
        eng = Stage1Engine(ca, d1, g1, B, ops=ops)
        real = torch.randn(B, 3, 64, 64, generator=g).clamp_(-1, 1).to(dev)
        step = lambda: eng.step(real, tem, tem_mis, z, e1, egp, use_graph=False)
    else:
        from imagegenerator_b200.discriminator_2 import StageIIDiscriminator
        from imagegenerator_b200.generator_2 import StageIIGenerator
        from imagegenerator_b200.engine2 import Stage2Engine
        ca1, g1 = ConditioningAugmentation(512, 256, 128), StageIGenerator(128, 100)
        ca2, d2, g2 = ConditioningAugmentation(512, 256, 128), StageIIDiscriminator(512, 128), StageIIGenerator()
        eng = Stage2Engine(ca1, g1, ca2, d2, g2, B, ops=ops)
        real = torch.randn(B, 3, 256, 256, generator=g).clamp_(-1, 1).to(dev)
        step = lambda: eng.step(real, tem, tem_mis, z, e1, e2, egp, use_graph=False)
    step()
    torch.cuda.synchronize()
    ops.on = True
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    step()
    t1.record()
    torch.cuda.synchronize()
    agg, cnt = defaultdict(float), defaultdict(int)
    for key, a, b in ops.rec:
        agg[key] += a.elapsed_time(b)
        cnt[key] += 1
    total = sum(agg.values())
    print(f"stage {stage} B={B}: step {t0.elapsed_time(t1):.2f} ms, sum of ops {total:.2f} ms, {len(ops.rec)} calls")
    byname = defaultdict(float)
    for k, v in agg.items():
        byname[k.split(" ")[0]] += v
    print("-- by op")
    for k, v in sorted(byname.items(), key=lambda x: -x[1])[:25]:
        print(f"{v:9.3f} ms {100 * v / total:5.1f}%  {k}")
    print("-- by op and shapes")
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:top]:
        print(f"{v:9.3f} ms {100 * v / total:5.1f}%  {cnt[k]:3d}x {1e3 * v / cnt[k]:8.1f} us  {k}")


if __name__ == "__main__":
    main()
