"""BatchNorm backward: reduce + apply (two launches) against sg_bn_bwd (one launch, ranges parked in shared memory) on the
BatchNorm layers of both stages.  Two timings each: isolated after an L2 flush, and 20 calls back to back between one event
pair (operands L2-resident, like inside the step where the conv that produced `da` has just finished).
python tools/bench_bn_fused.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from imagegenerator_b200.ops import CudaOps, ACT_LRELU

SHAPES = [("s1.D1.ds1", 128 * 256, 128, 3), ("s1.D1.ds2", 128 * 64, 256, 3), ("s1.D1.ds3", 128 * 16, 512, 3),
          ("s1.D1.gp.ds1", 128 * 256, 128, 1), ("s1.D1.gp.ds2", 128 * 64, 256, 1),
          ("s1.G1.up0", 128 * 16, 192, 1), ("s1.G1.up1", 128 * 64, 96, 1), ("s1.G1.up2", 128 * 256, 48, 1), ("s1.G1.up3", 128 * 1024, 24, 1),
          ("s2.G2.res640", 64 * 256, 640, 1), ("s2.G2.res320", 64 * 256, 320, 1), ("s2.G2.up0", 64 * 32 * 32, 320, 1),
          ("s2.D2.ds3", 64 * 32 * 32, 64, 3), ("s2.D2.ds4", 64 * 16 * 16, 128, 3), ("s2.G2.up1", 64 * 64 * 64, 160, 1)]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
ops = CudaOps("bf16")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def isolated(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


def warm(fn, n=20):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    ts = []
    for _ in range(max(3, reps // 2)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3 / n


print(f"{'tensor':14s} {'MB':>6s} | isolated (L2 flushed): {'2 launches':>10s} {'1 launch':>9s} | back to back in a graph: {'2 launches':>10s} {'1 launch':>9s} {'keep50/100':>10s}")
for name, rpg, C, G in SHAPES:
    n = rpg * G
    mk = lambda: (torch.randn(n, C, device="cuda")).to(torch.bfloat16)
    y, da, dy = mk(), mk(), mk()
    mr = torch.rand(G, C, 2, device="cuda") + 0.5
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    sums = torch.zeros(G, C, 2, dtype=torch.float64, device="cuda")
    mb = n * C * 2 / 1e6

    def two():
        ops.bn_bwd_reduce(da, None, y, mr, sums, G, ACT_LRELU, gamma=gamma, beta=beta)
        ops.bn_bwd_apply(da, None, y, mr, gamma, sums, dy, G, ACT_LRELU, beta=beta)

    def one():
        ops.bn_bwd(da, None, y, mr, gamma, sums, dy, G, ACT_LRELU, beta=beta)

    ops.set_option("bn_fused", 1); ops.set_option("bn_fused_keep_pct", 0)
    n0 = ops.launch_count(); one(); fused = ops.launch_count() - n0 == 1
    a, b = isolated(two), isolated(one)
    c, d = warm(two), warm(one)
    ops.set_option("bn_fused", 0)
    print(f"{name:14s} {mb:6.1f} | {'':22s} {a:10.1f} {b:9.1f} | {'':24s} {c:10.1f} {d:9.1f}   {'fused' if fused else 'NOT fused'}", flush=True)

# ---- where the one-launch kernel spends its time: globaltimer stamps of its first and last CTA (option bn_fused_dbg)
print("\ntimeline of one isolated launch, us after the first CTA's start: claimed+copies issued | range reduced | sums added | done posted | "
      "rendezvous passed | applied   (first CTA / last CTA)")
ops.set_option("bn_fused_dbg", 1); ops.set_option("bn_fused", 1)
for name, rpg, C, G in SHAPES[:9]:
    n = rpg * G
    mk = lambda: (torch.randn(n, C, device="cuda")).to(torch.bfloat16)
    y, da, dy = mk(), mk(), mk()
    mr = torch.rand(G, C, 2, device="cuda") + 0.5
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    sums = torch.zeros(G, C, 2, dtype=torch.float64, device="cuda")
    for _ in range(3):
        flush.zero_()
        ops.bn_bwd(da, None, y, mr, gamma, sums, dy, G, ACT_LRELU, beta=beta)
    torch.cuda.synchronize()
    slot = ops._bn_slot[sums.data_ptr()]
    st = ops._bn_work[256 * slot + 200: 256 * slot + 232].view(torch.int64).cpu().tolist()
    a, b = st[:7], st[8:15]
    t0 = min(a[0], b[0])
    print(f"{name:14s} " + " | ".join(f"{(a[i] - t0) / 1e3:5.1f}/{(b[i] - t0) / 1e3:5.1f}" for i in range(7)), flush=True)
