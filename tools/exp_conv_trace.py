"""Timeline of one persistent conv launch from the kernel's own %globaltimer stamps (sg_debug_conv_trace).
    python tools/exp_conv_trace.py [N H Ci Co k s p] [stats]     default: critic ds3 at the bench batch"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from imagegenerator_b200.ops import CudaOps  # noqa: E402

NAMES = ["entry", "setup done", "pdl passed", "first operands", "item0 issued", "last item issued", "item0 acc ready",
         "item0 drained", "last acc ready", "last drained", "exit"]


def main():
    nums = [v for v in sys.argv[1:] if v.lstrip("-").isdigit()]
    a = [int(v) for v in nums[:7]] if len(nums) >= 7 else [384, 16, 128, 256, 4, 2, 1]
    with_stats = "stats" in sys.argv
    N, H, Ci, Co, k, s, p = a
    ops = CudaOps("bf16")
    for kv in filter(None, os.environ.get("SG_OPTS", "").split(",")):
        key, val = kv.split("=")
        ops.set_option(key, int(val))
    Ho = (H + 2 * p - k) // s + 1
    x = (torch.randn(N, H, H, Ci, device="cuda") * 0.5).to(torch.bfloat16)
    y = torch.empty(N, Ho, Ho, Co, device="cuda", dtype=torch.bfloat16)
    pf = (torch.randn(Co, k, k, Ci, device="cuda") * 0.05).to(torch.bfloat16)
    st = torch.zeros(1, Co, 2, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn = (lambda: ops.conv_fprop_stats(x, pf, y, st, 1, k, s, p)) if with_stats else (lambda: ops.conv_fprop(x, pf, None, y, k, s, p))
    for _ in range(3):
        fn()
    buf = torch.zeros(296 * 16, dtype=torch.int64, device="cuda")
    flush.zero_()
    torch.cuda.synchronize()
    ops.conv_trace(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); e1.synchronize()
    ops.conv_trace(None)
    t = buf.cpu().view(296, 16)
    used = t[:, 0] > 0
    t = t[used].double()
    t0 = t[:, 0].min()
    print(f"N={N} H={H} Ci={Ci} Co={Co} k={k} s={s} p={p} stats={with_stats}: {int(used.sum())} CTAs, event time {e0.elapsed_time(e1) * 1e3:.1f} us")
    print(f"{'stamp':20s} {'min us':>9s} {'median':>9s} {'max us':>9s}   (relative to the first CTA's entry)")
    for i, nm in enumerate(NAMES):
        col = t[:, i]
        col = col[col > 0]
        if col.numel() == 0:
            continue
        v = (col - t0) / 1e3
        print(f"{nm:20s} {v.min().item():9.2f} {v.median().item():9.2f} {v.max().item():9.2f}")


if __name__ == "__main__":
    main()
