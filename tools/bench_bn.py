"""GB/s of the BatchNorm / elementwise kernels on the StackGAN tensor shapes.  python tools/bench_bn.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from imagegenerator_b200.ops import CudaOps, ACT_LRELU, ACT_RELU

SHAPES = [("s2.G2.up2", 64 * 128 * 128, 80, 1), ("s2.G2.up1", 64 * 64 * 64, 160, 1), ("s2.G2.up0", 64 * 32 * 32, 320, 1),
          ("s2.G2.res640", 64 * 256, 640, 1), ("s2.G2.res320", 64 * 256, 320, 1), ("s2.D2.ds2", 64 * 64 * 64, 32, 3),
          ("s2.D2.ds3", 64 * 32 * 32, 64, 3), ("s1.D1.ds2", 128 * 256, 128, 3), ("s1.D1.ds3", 128 * 64, 256, 3),
          ("s1.D1.ds4", 128 * 16, 512, 3), ("s1.G1.up3", 128 * 1024, 24, 1)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ops = CudaOps("bf16")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


print(f"{'tensor':14s} {'MB':>7s} | {'bn_act us':>9s} {'GB/s':>6s} | {'bwd_reduce':>10s} {'GB/s':>6s} | {'bwd_apply':>9s} {'GB/s':>6s} | {'act_bwd':>8s} {'GB/s':>6s}")
for name, rpg, C, G in SHAPES:
    n = rpg * G
    mk = lambda: (torch.randn(n, C, device="cuda")).to(torch.bfloat16)
    y, a, da, dy = mk(), mk(), mk(), mk()
    mr = torch.rand(G, C, 2, device="cuda") + 0.5
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    sums = torch.zeros(G, C, 2, dtype=torch.float64, device="cuda")
    mb = n * C * 2 / 1e6
    t1 = timeit(lambda: ops.bn_act(y, mr, gamma, beta, a, G, ACT_LRELU))
    t2 = timeit(lambda: ops.bn_bwd_reduce(da, a, y, mr, sums, G, ACT_LRELU))
    t3 = timeit(lambda: ops.bn_bwd_apply(da, a, y, mr, gamma, sums, dy, G, ACT_LRELU))
    t4 = timeit(lambda: ops.act_bwd(da, a, dy, ACT_LRELU))
    print(f"{name:14s} {mb:7.1f} | {t1*1e3:9.1f} {2*mb/t1/1e3:6.0f} | {t2*1e3:10.1f} {3*mb/t2/1e3:6.0f} | {t3*1e3:9.1f} {4*mb/t3/1e3:6.0f} | {t4*1e3:8.1f} {3*mb/t4/1e3:6.0f}", flush=True)
